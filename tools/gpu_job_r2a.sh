# Development aid for the NEXT round's first gpurun call: (1) the GPU suite with the experiment tests on (verifies the two
# arithmetic fixes made after round 1's GPU budget ran out, and that LF_SORT_RAYS changes no pixel), (2) A/B of ray sorting
# (LF_SORT_RAYS=1 extend, 2 shadow, 3 both) on every workload, (3) k_shade at 5 / 6 / 7 CTAs per SM on every workload.
# Before the call, in the authoring container:
#   for n in 5 6; do bash tools/build_variant.sh shade$n -DLF_SHADE_MINBLOCKS=$n; done; bash tools/build_variant.sh sample6 -DLF_SAMPLE_MINBLOCKS=6
#   (and remove `ab` from .gpurunignore for that call)
# usage: bash tools/gpu_job_r2a.sh <tag>
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
( time LF_TEST_EXPERIMENTS=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -4 $out/${tag}_pytest_gpu.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 150 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-llvmpipe --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-300
}
for w in c2_full c4_stress c3_full c1; do
  ab $w default LF_DUMMY=1
  for m in 1 2 3; do ab $w sort$m LF_SORT_RAYS=$m; done
  for v in shade5 shade6 sample6; do [ -f ab/$v.so ] && ab $w $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
done
