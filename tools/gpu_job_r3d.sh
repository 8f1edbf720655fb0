# Round 2 (1 GPU): whole-record writes where a half was written unread (LF_FULL_RECORDS): GPU suite, A/B against the halves-only build.
tag=${1:-r3d}
out=gpurun_out
mkdir -p $out
( time timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -4 $out/${tag}_pytest_gpu.txt | head -2
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-250
}
for w in c3_full c2_full c4_stress; do
  ab $w default LF_DUMMY=1
  ab $w half LF_LFCUDA_SO=$PWD/ab/half.so
done
