# Round 2, last check of the committed state on one GPU: the GPU suite, smoke, the default bench line (what the driver runs).
tag=${1:-r2z}
out=gpurun_out
mkdir -p $out
( time LF_TEST_EXPERIMENTS=1 timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -12 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -4 $out/${tag}_pytest_gpu.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $out/${tag}_smoke.txt 2>&1; tail -1 $out/${tag}_smoke.txt
( time timeout 900 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err ) 2>&1 | grep real
python tools/bench_brief.py < $out/${tag}_bench_c2.json | cut -c1-400
