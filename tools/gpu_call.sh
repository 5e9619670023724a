cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/t12.log 2>&1; tail -n 3 $O/t12.log
for v in head s16m3 s16m2 s8m3 s20m2; do
  B2PT_LIB=build/variants/libb2pt_$v.so timeout 300 python tools/exp_walk.py $v 2>&1 | tail -n 3
done
