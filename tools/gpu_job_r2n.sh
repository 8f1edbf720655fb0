# Round 2 (1 GPU): GPU suite, then the light tests moved out of k_trace into the dense kernels + occupancy-sized persistent grids (any-hit kernel:
# 48 registers -> 10 CTAs per SM) against the commit before (ab/prev.so); LF_CTAS_PER_SM=9 separates the two effects.
tag=${1:-r2n}
out=gpurun_out
mkdir -p $out
( time LF_TEST_EXPERIMENTS=1 timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -14 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -5 $out/${tag}_pytest_gpu.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-330
}
for w in c2_full c4_stress c3_full c1; do
  ab $w default LF_DUMMY=1
  ab $w prev LF_LFCUDA_SO=$PWD/ab/prev.so
  ab $w cap9 LF_CTAS_PER_SM=9
done
