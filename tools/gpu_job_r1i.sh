# Development aid (run through gpurun): final-state evidence of round 1 on one B200 - GPU tests, bench line of every workload,
# the reference arm, the ncu launch list and one `--set full` capture.  usage: bash tools/gpu_job_r1i.sh <tag>
tag=${1:-r1i}
out=gpurun_out
mkdir -p $out
( time timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $out/${tag}_pytest_gpu.txt 2>&1
timeout 400 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
for w in c4_stress c3_full c1; do
  timeout 200 python bench.py --steps 8 --warmup 3 --workload $w --no-llvmpipe --no-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_c2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-llvmpipe > $out/${tag}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_trace|k_shade|k_sample" -s 12 -c 4 -o $out/${tag}_prof_c2 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-llvmpipe > $out/${tag}_ncu_c2.log 2>&1
ncu -i $out/${tag}_prof_c2.ncu-rep --page raw --csv > $out/${tag}_c2_ncu_raw.csv 2>/dev/null
rm -f $out/${tag}_prof_c2.ncu-rep
tail -3 $out/${tag}_pytest_gpu.txt
for w in c2 c4_stress c3_full c1; do echo "== $w"; python tools/bench_brief.py < $out/${tag}_bench_$w.json; done
head -c 700 $out/${tag}_bench_ref.json
