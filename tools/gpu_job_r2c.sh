# Round 2, third GPU call: (1) GPU suite, (2) fused shade+sample kernel A/B (3 / 4 / 5 CTAs per SM) against the split kernels on all four
# workloads, (3) ncu --set full with source of the bounce-1 traversal launches (where the rays are incoherent) and of the shade kernels.
tag=${1:-r2c}
out=gpurun_out
mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -30 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -5 $out/${tag}_pytest_gpu.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-330
}
for w in c2_full c4_stress c3_full c1; do
  ab $w default LF_DUMMY=1
  for v in fused3 fused4 fused5; do [ -f ab/$v.so ] && ab $w $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_trace -s 2 -c 2 -o $out/${tag}_trace_b1 -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-llvmpipe --no-c5 > $out/${tag}_ncu_trace.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_shade\|k_sample -s 1 -c 2 -o $out/${tag}_shade_b1 -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-llvmpipe --no-c5 > $out/${tag}_ncu_shade.log 2>&1
ls -la $out/*.ncu-rep
