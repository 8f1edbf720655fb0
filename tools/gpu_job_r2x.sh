# Round 2 (2 GPUs), the shipped kernels: everything that needs two devices (NCCL reduce test, C5-shaped tile through NCCL and through a
# peer-access group, multi-GPU CudaRenderer), bench.py under torchrun at N = 2 (reduce_check, c5_strong, e2e through lfcuda_reduce),
# lf_render --gpus 1 / 2 on the 4K scene.
tag=${1:-r2x}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/${tag}_gpus.txt
( time timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_group_gpu.py tests/test_cuda_renderer.py tests/test_tlas_device_gpu.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -30 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -5 $out/${tag}_pytest_gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 32 --warmup 3 \
    > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
python tools/bench_brief.py < $out/${tag}_bench_n2.json | cut -c1-300
python -c "
import json; j=json.loads([l for l in open('$out/${tag}_bench_n2.json') if l.startswith('{')][-1]); print('reduce_check', j.get('reduce_check')); print('c5_strong', j.get('c5_strong')); print('e2e', j.get('e2e'))"
python - <<'PY'
import os, sys
sys.path.insert(0, '.')
from scenes import gen_scenes
print(gen_scenes.SCENES['c4_stress'](os.path.join('scenes', '_gen', 'c4_stress')))
PY
SC=$(ls scenes/_gen/c4_stress/assets/*.scene | head -1)
for g in 1 2; do
  timeout 600 lavaframe_b200/bin/lf_render $SC --spp 256 --gpus $g --out $out/${tag}_lfrender_g$g.f32 2>&1 | tail -1 | tee $out/${tag}_lfrender_g$g.json
done
python - <<PY
import numpy as np
a = np.fromfile('$out/${tag}_lfrender_g1.f32', np.float32); b = np.fromfile('$out/${tag}_lfrender_g2.f32', np.float32)
err = np.abs(a - b); print('lf_render 1 vs 2 GPUs: max abs', err.max(), 'allclose(rtol 2e-5, atol 1e-6):', bool(np.allclose(a, b, rtol=2e-5, atol=1e-6)), 'mean', a.mean(), b.mean())
PY
rm -f $out/${tag}_lfrender_g*.f32
