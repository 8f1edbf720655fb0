# Round 2, last one-GPU check of the committed state: the GPU suite, smoke, the default bench line and the reference arm (what the driver runs).
tag=${1:-r2o}
out=gpurun_out
mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -12 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -4 $out/${tag}_pytest_gpu.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $out/${tag}_smoke.txt 2>&1; tail -1 $out/${tag}_smoke.txt
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err ) 2>&1 | grep real
( time timeout 900 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err ) 2>&1 | grep real
python tools/bench_brief.py < $out/${tag}_bench_c2.json | cut -c1-400
python -c "
import json; j=json.load(open('$out/${tag}_bench_c2.json')); print({k: j[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','higher_is_better','scaling','vs_baseline','dtype','data','gpu_launches')}); print(j['config']); print(j['e2e']); print(j['clocks']); print('c5', j['c5_strong']['seconds'])"
