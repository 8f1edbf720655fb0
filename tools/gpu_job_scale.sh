# Development aid (run through gpurun --gpus G): per-scene throughput at N GPUs, spp split + NCCL sum (bench.py under torchrun).
# usage: bash tools/gpu_job_scale.sh "<N list>" "<workloads>"      outputs: gpurun_out/scale_<workload>_n<N>.json
# c4_stress is run as C5: 4096 spp in total, i.e. 128 / N steps of 32 spp per rank.
port=29500
for n in $1; do
for w in $2; do
  steps=8
  if [ $w = c4_stress ]; then steps=$((128 / n)); fi
  if [ $w = c2_full ]; then steps=$((32 / n)); fi
  port=$((port + 1))
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps $steps --warmup 3 --workload $w --no-llvmpipe > gpurun_out/scale_${w}_n$n.json 2> gpurun_out/scale_${w}_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps $steps --warmup 3 \
        --workload $w --no-cpu-baseline --no-llvmpipe > gpurun_out/scale_${w}_n$n.json 2> gpurun_out/scale_${w}_n$n.err
  fi
  echo "== $w N=$n steps=$steps"; python tools/bench_brief.py < gpurun_out/scale_${w}_n$n.json | cut -c1-160
done; done
