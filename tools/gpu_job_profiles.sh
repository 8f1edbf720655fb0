# Development aid (run through gpurun): DRAM traffic per traversal launch and one full ncu capture of the C4 traversal.
set -x
for w in c2_full c4_stress; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace --csv \
      --log-file gpurun_out/traffic_$w.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-llvmpipe --workload $w > gpurun_out/traffic_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 8 -c 2 -o gpurun_out/prof_trace_c4 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-llvmpipe --workload c4_stress > gpurun_out/ncu_trace_c4.log 2>&1
ls -la gpurun_out | tail -5
