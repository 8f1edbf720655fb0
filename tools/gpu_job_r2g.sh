# Round 2, seventh GPU call (1 GPU): GPU suite (EXR export), the price of bit-exact arithmetic (north_star's three bars for the -fmad=true
# build against the llvmpipe goldens + its speed), re-sweep of the traversal tunables (refill threshold, leaf-gather fraction).
tag=${1:-r2g}
out=gpurun_out
mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -12 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -5 $out/${tag}_pytest_gpu.txt
python tools/fmad_bars.py > $out/${tag}_bars_exact.txt 2>&1; cat $out/${tag}_bars_exact.txt
LF_LFCUDA_SO=$PWD/ab/fmad.so python tools/fmad_bars.py > $out/${tag}_bars_fmad.txt 2>&1; cat $out/${tag}_bars_fmad.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-330
}
for w in c2_full c4_stress c3_full; do
  ab $w default LF_DUMMY=1
  for v in fmad refill4 refill16 gather2 gather4; do [ -f ab/$v.so ] && ab $w $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
done
