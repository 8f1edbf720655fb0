# Round 2 (1 GPU): k_sample re-reads the material from the table instead of being handed it through the path state (LF_SAMPLE_REMAT):
# the GPU suite on the new default, then A/B against the previous layout on every workload.
tag=${1:-r3a}
out=gpurun_out
mkdir -p $out
( time LF_TEST_EXPERIMENTS=1 timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -4 $out/${tag}_pytest_gpu.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-260
}
for w in c2_full c4_stress c3_full c1; do
  ab $w default LF_DUMMY=1
  ab $w remat0 LF_LFCUDA_SO=$PWD/ab/remat0.so
  ab $w default2 LF_DUMMY=2
done
