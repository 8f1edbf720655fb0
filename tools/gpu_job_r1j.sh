# Development aid (run through gpurun): A/B of node / triangle fetches through the texture unit (LF_NODE_TEX, LF_TRI_TEX) on C2 and C4,
# plus the parity tests of the two kernel files with the best-looking variant.  usage: bash tools/gpu_job_r1j.sh <tag>
tag=${1:-r1j}
out=gpurun_out
mkdir -p $out
ab() {  # workload, name, env assignments...
  w=$1; name=$2; shift; shift
  env "$@" timeout 150 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-llvmpipe --workload $w > $out/${tag}_ab_${w}_$name.json 2> $out/${tag}_ab_${w}_$name.err
  echo "== $w $name"; python tools/bench_brief.py < $out/${tag}_ab_${w}_$name.json | cut -c1-330
}
ab c2_full default LF_DUMMY=1
for v in nodetex1 nodetex2 tritex nodetex2_tritex nodetex1_tritex; do ab c2_full $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
ab c2_full default2 LF_DUMMY=1
ab c4_stress default LF_DUMMY=1
for v in nodetex1 nodetex2 nodetex2_tritex; do ab c4_stress $v LF_LFCUDA_SO=$PWD/ab/$v.so; done
for v in nodetex2 nodetex1_tritex; do
  ( LF_LFCUDA_SO=$PWD/ab/$v.so timeout 300 python -m pytest tests/test_cuda_parity.py tests/test_edge_cases.py -m gpu -q 2>&1 | tail -5 ) > $out/${tag}_pytest_$v.txt 2>&1
  tail -2 $out/${tag}_pytest_$v.txt
done
