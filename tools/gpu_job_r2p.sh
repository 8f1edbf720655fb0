# Round 2 (1 GPU): traversal micro-variants - the phase-end test every second inner step; a separate leaf-gather threshold for the any-hit kernel.
tag=${1:-r2p}
out=gpurun_out
mkdir -p $out
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-300
}
for w in c2_full c4_stress c3_full; do
  ab $w default LF_DUMMY=1
  for v in steps2 any13 any23 steps2any23; do [ -f ab/$v.so ] && ab $w $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
done
