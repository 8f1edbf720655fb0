# Development aid (run through gpurun): second verification pass (post-process parity tests) + A/B of two knobs + DRAM traffic
# of the traversal launches with the final kernels.  usage: bash tools/gpu_job_r1h.sh <tag>
tag=${1:-r1h}
out=gpurun_out
mkdir -p $out
( time timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $out/${tag}_pytest_gpu.txt 2>&1
ab() {  # name, env assignments...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe > $out/${tag}_ab_c2_$name.json 2> $out/${tag}_ab_c2_$name.err
  echo "== $name"; python tools/bench_brief.py < $out/${tag}_ab_c2_$name.json
}
ab default LF_DUMMY=1
ab l2p48 LF_L2_PERSIST_MB=48
ab l2p80 LF_L2_PERSIST_MB=80
ab inlmath LF_LFCUDA_SO=$PWD/ab/inlmath.so
ab default2 LF_DUMMY=1
ab inlmath_l2p48 LF_LFCUDA_SO=$PWD/ab/inlmath.so LF_L2_PERSIST_MB=48
for w in c2_full c4_stress; do
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace --csv \
      --log-file $out/${tag}_traffic_$w.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-llvmpipe --workload $w > $out/${tag}_traffic_$w.log 2>&1
done
tail -4 $out/${tag}_pytest_gpu.txt
