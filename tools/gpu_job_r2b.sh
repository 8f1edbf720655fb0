# Round 2, second GPU call: (1) GPU suite (new: device groups, multi-GPU CudaRenderer), (2) shared-memory / L1 carve-out A/B of the
# traversal kernels: LF_SH_STACK (entries of the stack kept in shared memory) x LF_WRAY_SHARED on C2 and C4, (3) ncu counters of the
# traversal launches for the default and the old layout (L1 hit rate, wavefronts, lanes per instruction), and of LF_SORT_RAYS=1.
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -30 ) > $out/${tag}_pytest_gpu.txt 2>&1
tail -5 $out/${tag}_pytest_gpu.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-330
}
for w in c2_full c4_stress; do
  ab $w default LF_DUMMY=1
  for v in sh32w1 sh12w1 sh8w0 sh16w0 sh12w0c10; do [ -f ab/$v.so ] && ab $w $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
done
ab c3_full default LF_DUMMY=1
ab c1 default LF_DUMMY=1
# ncu: per-launch counters of the traversal kernels (one bench step), default vs old layout vs sorted rays
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,launch__shared_mem_config_size
prof() {
  name=$1; shift
  env "$@" timeout 600 ncu --metrics $M --clock-control none -k regex:k_trace\|k_sort -c 60 --csv --log-file $out/${tag}_ncu_$name.csv \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-llvmpipe --no-c5 > $out/${tag}_ncu_$name.log 2>&1
  echo "ncu $name: $(wc -l < $out/${tag}_ncu_$name.csv) lines"
}
prof default LF_DUMMY=1
[ -f ab/sh32w1.so ] && prof sh32w1 LF_LFCUDA_SO=$PWD/ab/sh32w1.so
prof sort1 LF_SORT_RAYS=1
