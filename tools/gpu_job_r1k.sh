# Development aid (run through gpurun): A/B of the resident CTAs per SM of k_shade / k_sample (LF_SHADE_MINBLOCKS, LF_SAMPLE_MINBLOCKS) on C2.
tag=${1:-r1k}
out=gpurun_out
mkdir -p $out
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 100 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-llvmpipe --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-300
}
ab c2_full default LF_DUMMY=1
for v in shade7 shade9 shade10 shade6 sample9 sample10 sample12; do ab c2_full $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
