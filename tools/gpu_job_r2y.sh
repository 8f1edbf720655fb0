# Round 2 (8 GPUs), the shipped kernels: bench.py under torchrun at N = 8 (reduce_check, c5_strong = BASELINE config 5) and the single renderer
# on 8 GPUs at 4096 spp.
tag=${1:-r2y}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/${tag}_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 32 --warmup 3 \
    > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err
python tools/bench_brief.py < $out/${tag}_bench_n8.json | cut -c1-200
python -c "
import json; j=json.loads([l for l in open('$out/${tag}_bench_n8.json') if l.startswith('{')][-1]); print('reduce_check', j.get('reduce_check')); c=j.get('c5_strong'); print('c5_strong', {k: c[k] for k in ('n_gpus','seconds','value') if k in c} if c else c); print('e2e', j['e2e']['value'])"
SC=$(ls scenes/_gen/c4_stress/assets/*.scene | head -1)
timeout 300 lavaframe_b200/bin/lf_render $SC --spp 4096 --gpus 8 2>&1 | tail -1 | tee $out/${tag}_lfrender_g8_4096.json
