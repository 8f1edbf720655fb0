# Round 2, first GPU call: (1) the GPU suite with the experiment tests on (new: whole-frame radiance vs oracle at full size, 18 stress
# seeds, instance edits vs oracle, NaN slab KAT, distant light, accumulation survives uniform changes), smoke, (2) default bench line,
# (3) ray-sort A/B (LF_SORT_RAYS 1/2/3), k_shade at 5/6 CTAs per SM and the -fmad=true build on C2 and C4.
# Before the call: LF_FMAD=true bash tools/build_variant.sh fmad; for n in 5 6; do bash tools/build_variant.sh shade$n -DLF_SHADE_MINBLOCKS=$n; done
# (and remove `ab` from .gpurunignore for that call).   usage: bash tools/gpu_job_r2a.sh <tag>
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt
( time LF_TEST_EXPERIMENTS=1 timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -45 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -5 $out/${tag}_pytest_gpu.txt
( timeout 300 python __graft_entry__.py smoke ) > $out/${tag}_smoke.txt 2>&1; tail -2 $out/${tag}_smoke.txt
timeout 600 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; python tools/bench_brief.py < $out/${tag}_bench_c2.json | cut -c1-400
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-llvmpipe --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-300
}
for w in c2_full c4_stress; do
  ab $w default LF_DUMMY=1
  for m in 1 2 3; do ab $w sort$m LF_SORT_RAYS=$m; done
  for v in shade5 shade6 fmad; do [ -f ab/$v.so ] && ab $w $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
done
