# Development aid (run through gpurun): verification of the final binary (k_shade at 7 CTAs per SM) + the last occupancy A/B.
tag=${1:-r1l}
out=gpurun_out
mkdir -p $out
( time timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -4 $out/${tag}_pytest_gpu.txt
( timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2 ) > $out/${tag}_smoke.txt; cat $out/${tag}_smoke.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 100 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-llvmpipe --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-300
}
ab c2_full default LF_DUMMY=1
for v in shade6 shade5 shade8 sample7 sample6; do ab c2_full $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
ab c3_full default LF_DUMMY=1
ab c3_full shade6 LF_LFCUDA_SO=$PWD/ab/shade6.so
ab c3_full shade8 LF_LFCUDA_SO=$PWD/ab/shade8.so
ab c4_stress default LF_DUMMY=1
ab c4_stress shade6 LF_LFCUDA_SO=$PWD/ab/shade6.so
ab c4_stress shade8 LF_LFCUDA_SO=$PWD/ab/shade8.so
timeout 120 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
python tools/bench_brief.py < $out/${tag}_bench_c2.json | cut -c1-300
