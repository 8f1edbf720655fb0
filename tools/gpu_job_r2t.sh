# Round 2 (1 GPU): the device BLAS build (tests first, then timing), the GPU suite, and the two-stream overlap probe.
tag=${1:-r2t}
out=gpurun_out
mkdir -p $out
( time timeout 600 python -m pytest tests/test_blas_device_gpu.py -m gpu -q -x -s 2>&1 | tail -25 ) > $out/${tag}_pytest_blas.txt 2>&1
tail -12 $out/${tag}_pytest_blas.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_blas_device_gpu.py 2>&1 | tail -6 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -4 $out/${tag}_pytest_gpu.txt
for w in c2_full c4_stress; do
  timeout 400 python tools/overlap_probe.py $w 32 > $out/${tag}_overlap_$w.json 2> $out/${tag}_overlap_$w.err
  cat $out/${tag}_overlap_$w.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_blas_launches.csv \
    python -c "
import sys; sys.path.insert(0, 'tests')
import lavaframe_b200 as lf, bench
from blas_cases import pack_meshes
m = max(pack_meshes(lf.ScenePack(bench.ensure_pack('c2_full'))), key=lambda q: len(q['bounds']))
print(lf.build_blas(m['bounds'], 0)[3])" > $out/${tag}_blas_ncu.log 2>&1
tail -2 $out/${tag}_blas_ncu.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/r2t_blas_launches.csv') if l.startswith('"'))]
if rows:
    h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
    t = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(',', '')); v = v / 1e3 if r[ui] == 'ns' else v
        k = r[ki].split('<')[-1].split('>')[0] if 'k_blas_step' in r[ki] else r[ki][:40]
        t[k][0] += 1; t[k][1] += v
    for k, (c, us) in sorted(t.items(), key=lambda kv: -kv[1][1]): print(f'{k:40s} {c:5d} launches {us/1e3:9.3f} ms')
PY
