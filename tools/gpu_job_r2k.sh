# Round 2, eleventh GPU call (1 GPU): GPU suite, then the path-state diet (initial state not stored, `stale` only with lights, hit point formed in the
# shade kernels) against the commit before (ab/prev.so) on all four workloads.
tag=${1:-r2k}
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
  ab $w diet0 LF_LFCUDA_SO=$PWD/ab/diet0.so
  ab $w default2 LF_DUMMY=2
done
