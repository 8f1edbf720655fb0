# Round 2 (1 GPU): device BLAS build with the zero-sign keys (tests + timing), compute-sanitizer memcheck over the builder on the synthetic inputs.
tag=${1:-r2w}
out=gpurun_out
mkdir -p $out
( time timeout 600 python -m pytest tests/test_blas_device_gpu.py -m gpu -q -x -s 2>&1 | tail -25 ) > $out/${tag}_pytest_blas.txt 2>&1
grep -E "passed|failed|device BLAS|mesh BVH" $out/${tag}_pytest_blas.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "
import sys; sys.path.insert(0, 'tests')
import numpy as np, lavaframe_b200 as lf
from blas_cases import synthetic_cases, signed_zero_cases
for name, b in synthetic_cases() + signed_zero_cases()[:12]:
    if len(b) > 3000: continue
    r = lf.build_blas(b, 0)
    print(name, r[3]['num_nodes'], r[3]['launches'])
pack = lf.ScenePack('tests/golden/cornell.lfpack')
pt = lf.PathTracer(0); pt.upload_pack(pack); pt.render_frames(2, 2); img = pt.read_accum(); pt.close()
print('render ok', float(img.mean()))
" > $out/${tag}_sanitizer.txt 2>&1
echo "sanitizer rc $?"; tail -5 $out/${tag}_sanitizer.txt
