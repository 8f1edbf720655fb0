# Round 2 (1 GPU): texture coordinates fetched only for hits on textured materials: GPU suite, A/B against the build before, the default bench line.
tag=${1:-r3c}
out=gpurun_out
mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -4 $out/${tag}_pytest_gpu.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-260
}
for w in c2_full c3_full c4_stress; do
  ab $w default LF_DUMMY=1
  ab $w prev LF_LFCUDA_SO=$PWD/ab/prev.so
done
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $out/${tag}_smoke.txt 2>&1; tail -1 $out/${tag}_smoke.txt
( time timeout 900 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err ) 2>&1 | grep real
python tools/bench_brief.py < $out/${tag}_bench_c2.json | cut -c1-400
python -c "
import json; j=json.load(open('$out/${tag}_bench_c2.json')); print('c5_strong', j['c5_strong']['seconds'], 'e2e', j['e2e']['value'])"
