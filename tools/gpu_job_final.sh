# Development aid (run through gpurun): the evidence set of a round on one B200 - GPU tests, bench lines of every
# workload, the reference arm, the ncu launch list, DRAM traffic per traversal launch and one `--set full` capture.
# usage: bash tools/gpu_job_final.sh <tag>       (outputs: gpurun_out/<tag>_*)
tag=${1:-final}
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $out/${tag}_pytest_gpu.txt
python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
for w in c1 c3_full c4_stress; do
  python bench.py --steps 8 --warmup 3 --workload $w --no-llvmpipe > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
python bench.py --impl reference --steps 4 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_c2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-llvmpipe > $out/${tag}_launches.log 2>&1
for w in c2_full c4_stress; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace --csv \
      --log-file $out/${tag}_traffic_$w.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-llvmpipe --workload $w > $out/${tag}_traffic_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k "regex:k_trace|k_shade|k_sample" -s 12 -c 4 -o $out/${tag}_prof_c2 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-llvmpipe > $out/${tag}_ncu_c2.log 2>&1
ncu -i $out/${tag}_prof_c2.ncu-rep --page raw --csv > $out/${tag}_c2_ncu_raw.csv 2>/dev/null
ls -la $out | grep ${tag}_
