# Development aid (run through gpurun): verification + evidence after the llvmpipe-exact arithmetic change, most important first
# because the GPU budget may cut the job short.  usage: bash tools/gpu_job_r1g.sh <tag>   (outputs: gpurun_out/<tag>_*)
tag=${1:-r1g}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $out/${tag}_gpu.txt 2>&1
( time timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $out/${tag}_pytest_gpu.txt 2>&1
timeout 400 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
timeout 200 python bench.py --steps 8 --warmup 3 --workload c4_stress --no-llvmpipe --no-cpu-baseline > $out/${tag}_bench_c4_stress.json 2> $out/${tag}_bench_c4_stress.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_c2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-llvmpipe > $out/${tag}_launches.log 2>&1
LF_LFCUDA_SO=$PWD/ab/inlmath.so timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe > $out/${tag}_ab_c2_inlmath.json 2> $out/${tag}_ab_c2_inlmath.err
timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe > $out/${tag}_ab_c2_default.json 2> $out/${tag}_ab_c2_default.err
for w in c3_full c1; do
  timeout 150 python bench.py --steps 8 --warmup 3 --workload $w --no-llvmpipe --no-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_trace|k_shade|k_sample" -s 12 -c 4 -o $out/${tag}_prof_c2 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-llvmpipe > $out/${tag}_ncu_c2.log 2>&1
ncu -i $out/${tag}_prof_c2.ncu-rep --page raw --csv > $out/${tag}_c2_ncu_raw.csv 2>/dev/null
rm -f $out/${tag}_prof_c2.ncu-rep.tmp
ls -la $out | grep ${tag}_
tail -3 $out/${tag}_pytest_gpu.txt; cat $out/${tag}_bench_c2.json | head -c 600
