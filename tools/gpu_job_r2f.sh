# Round 2, sixth GPU call (1 GPU): GPU suite, then the classify stage (k_classify ends misses / emitter hits and compacts surface hits for
# k_shade) against the previous pipeline (ab/prev.so = the commit before) on all four workloads, and an ncu counter pass of the shade-side kernels.
tag=${1:-r2f}
out=gpurun_out
mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -20 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -5 $out/${tag}_pytest_gpu.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-330
}
for w in c2_full c4_stress c3_full c1; do
  ab $w default LF_DUMMY=1
  ab $w prev LF_LFCUDA_SO=$PWD/ab/prev.so
done
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
for v in default prev; do
  if [ $v = prev ]; then export LF_LFCUDA_SO=$PWD/ab/prev.so; else unset LF_LFCUDA_SO; fi
  timeout 600 ncu --metrics $M --clock-control none -k regex:k_shade\|k_sample\|k_classify -c 18 --csv --log-file $out/${tag}_ncu_shade_$v.csv \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-llvmpipe --no-c5 > $out/${tag}_ncu_shade_$v.log 2>&1
  echo "ncu $v: $(wc -l < $out/${tag}_ncu_shade_$v.csv) lines"
done
