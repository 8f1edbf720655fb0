# Round 2 (1 GPU): device BLAS build after the scan / bin-layout changes (tests + timing + launch list), branchless triangle-test A/B,
# frames-in-flight sweep on the 4K scene (C4 runs 3-frame batches with the 32 M-slot default).
tag=${1:-r2u}
out=gpurun_out
mkdir -p $out
( time timeout 600 python -m pytest tests/test_blas_device_gpu.py -m gpu -q -x -s 2>&1 | tail -25 ) > $out/${tag}_pytest_blas.txt 2>&1
grep -E "passed|failed|device BLAS|mesh BVH" $out/${tag}_pytest_blas.txt
ab() {
  w=$1; name=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe --no-c5 --workload $w $EXTRA > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-200
}
for w in c2_full c4_stress c3_full; do
  EXTRA= ab $w default LF_DUMMY=1
  EXTRA= ab $w tribl LF_LFCUDA_SO=$PWD/ab/tribl.so
done
for f in 8 16 32; do EXTRA="--frames-in-flight $f" ab c4_stress fif$f LF_DUMMY=1; done
EXTRA="--frames-in-flight 32" ab c3_full fif32 LF_DUMMY=1
EXTRA="--frames-in-flight 64" ab c2_full fif64 LF_DUMMY=1
EXTRA= ab c4_stress default2 LF_DUMMY=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_blas_launches.csv \
    python -c "
import sys; sys.path.insert(0, 'tests')
import lavaframe_b200 as lf, bench
from blas_cases import pack_meshes
m = max(pack_meshes(lf.ScenePack(bench.ensure_pack('c2_full'))), key=lambda q: len(q['bounds']))
print(lf.build_blas(m['bounds'], 0)[3])" > $out/${tag}_blas_ncu.log 2>&1
python tools/blas_launch_summary.py $out/${tag}_blas_launches.csv | tee $out/${tag}_blas_launch_summary.txt
