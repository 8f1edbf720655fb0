# Development aid: A/B of liblfcuda.so variants built by tools/build_variant.sh (run on the GPU box through gpurun).
# usage: bash tools/ab_run.sh "<workloads>" "<variants>"   (variant "default" = the in-tree library)
for w in $1; do
for v in $2; do
  if [ $v = default ]; then unset LF_LFCUDA_SO; else export LF_LFCUDA_SO=$PWD/ab/$v.so; fi
  case $v in *8) export LF_CTAS_PER_SM=8;; *) unset LF_CTAS_PER_SM;; esac   # variants named *8 are built for 8 CTAs per SM
  echo "== $w $v"
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe --workload $w 2>gpurun_out/ab_${w}_$v.err | tee gpurun_out/ab_${w}_$v.json | python tools/bench_brief.py
done; done
