# Round 2, final one-GPU evidence set of the shipped state (after the material re-read in k_sample):
# GPU suite (experiment tests included), smoke, the default bench line (C2 + c5_strong + cpu_baseline), the other workloads, the reference arm,
# launch lists (render step, BLAS build), DRAM traffic of the traversal launches, one `--set full` capture of the bounce-1 kernels.
tag=${1:-r3b}
out=gpurun_out
mkdir -p $out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $out/${tag}_smoke.txt 2>&1; tail -1 $out/${tag}_smoke.txt
( time timeout 900 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err ) 2>&1 | grep real
for w in c4_stress c3_full c1; do
  timeout 300 python bench.py --steps 8 --warmup 3 --workload $w --no-llvmpipe --no-cpu-baseline --no-c5 > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err ) 2>&1 | grep real
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_c2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-llvmpipe --no-c5 > $out/${tag}_launches.log 2>&1
for w in c2_full c4_stress; do
  timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace -c 200 --csv \
      --log-file $out/${tag}_traffic_$w.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_traffic_$w.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_trace|k_shade|k_sample" -s 4 -c 4 -o $out/${tag}_prof_c2 -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-llvmpipe --no-c5 > $out/${tag}_ncu_c2.log 2>&1
ncu -i $out/${tag}_prof_c2.ncu-rep --page raw --csv > $out/${tag}_c2_ncu_raw.csv 2>/dev/null
ncu -i $out/${tag}_prof_c2.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_trace > $out/${tag}_c2_trace_src.csv 2>/dev/null
python tools/ncu_lines.py $out/${tag}_c2_ncu_raw.csv $out/${tag}_c2_trace_src.csv 400 > $out/${tag}_c2_trace_ncu_lines.txt 2>&1
rm -f $out/${tag}_prof_c2.ncu-rep $out/${tag}_c2_trace_src.csv
for w in c2 c4_stress c3_full c1; do echo "== $w"; python tools/bench_brief.py < $out/${tag}_bench_$w.json | cut -c1-330; done
python -c "
import json; j=json.load(open('$out/${tag}_bench_c2.json')); print('c5_strong', j.get('c5_strong')); print('cpu_baseline', j.get('cpu_baseline')); print('roofline', {k: v for k, v in j['roofline'].items() if k != 'stage_ms'}); print(j['clocks'], j['gpu_launches'])"
head -c 600 $out/${tag}_bench_ref.json
