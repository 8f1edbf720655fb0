# Round 2, fourth GPU call (1 GPU): GPU suite (new: C5-shaped converged tile through a device group), the default bench line with
# c5_strong + cpu_baseline, k_shade occupancy A/B (5 = default, 4, 3 CTAs per SM; zero spills from 4) on all four workloads.
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -30 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -5 $out/${tag}_pytest_gpu.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-330
}
for w in c2_full c4_stress c3_full c1; do
  ab $w default LF_DUMMY=1
  for v in shade4 shade3 shade4s6 nofuse; do [ -f ab/$v.so ] && ab $w $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
done
timeout 900 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; python tools/bench_brief.py < $out/${tag}_bench_c2.json | cut -c1-400
python -c "
import json; j=json.load(open('$out/${tag}_bench_c2.json')); print('c5_strong', j.get('c5_strong')); print('cpu_baseline', j.get('cpu_baseline')); print('roofline', {k: v for k, v in j['roofline'].items() if k != 'stage_ms'})"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; cut -c1-600 $out/${tag}_bench_ref.json
