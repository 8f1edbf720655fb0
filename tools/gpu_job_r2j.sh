# Development aid (run through gpurun --gpus 8): round 2 on eight B200s of one box - the device-group and NCCL tests across 8 devices, bench.py
# under torchrun at N = 8 and N = 4 (reduce_check, c5_strong = BASELINE config 5), and the single renderer on 8 GPUs (lf_render --gpus 8).
tag=${1:-r2j}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/${tag}_gpus.txt
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_group_gpu.py tests/test_multigpu_gpu.py tests/test_cuda_renderer.py -m gpu -q -x -s -k "group or multi or c5 or two_gpu" 2>&1 | grep -v "^$" | tail -25 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -6 $out/${tag}_pytest_gpu.txt
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 32 --warmup 3 \
      > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
  python tools/bench_brief.py < $out/${tag}_bench_n$n.json | cut -c1-200
  python -c "
import json; j=json.loads([l for l in open('$out/${tag}_bench_n$n.json') if l.startswith('{')][-1]); print('reduce_check', j.get('reduce_check')); c=j.get('c5_strong'); print('c5_strong', {k: c[k] for k in ('n_gpus','seconds','value') if k in c} if c else c); print('e2e', j['e2e']['value'])"
done
python - <<'PY'
import os, sys
sys.path.insert(0, '.')
from scenes import gen_scenes
print(gen_scenes.SCENES['c4_stress'](os.path.join('scenes', '_gen', 'c4_stress')))
PY
SC=$(ls scenes/_gen/c4_stress/assets/*.scene | head -1)
timeout 600 lavaframe_b200/bin/lf_render $SC --spp 4096 --gpus 8 2>&1 | tail -1 | tee $out/${tag}_lfrender_g8_4096.json
timeout 600 lavaframe_b200/bin/lf_render $SC --spp 1024 --gpus 8 --out $out/${tag}_g8.f32 2>&1 | tail -1 | tee $out/${tag}_lfrender_g8_1024.json
timeout 600 lavaframe_b200/bin/lf_render $SC --spp 1024 --gpus 1 --out $out/${tag}_g1.f32 2>&1 | tail -1 | tee $out/${tag}_lfrender_g1_1024.json
python - <<PY | tee $out/${tag}_lfrender_compare.txt
import numpy as np
a = np.fromfile('$out/${tag}_g1.f32', np.float32); b = np.fromfile('$out/${tag}_g8.f32', np.float32)
err = np.abs(a - b); print('lf_render 1 vs 8 GPUs, 1024 spp: max abs', err.max(), 'allclose(rtol 2e-5, atol 1e-6):', bool(np.allclose(a, b, rtol=2e-5, atol=1e-6)), 'mean', a.mean(), b.mean())
PY
rm -f $out/${tag}_g1.f32 $out/${tag}_g8.f32
