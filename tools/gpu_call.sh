cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/t17.log 2>&1; tail -n 3 $O/t17.log
for v in base fin pf s12 s20; do
  B2PT_LIB=build/variants/libb2pt_$v.so timeout 300 python tools/exp_walk.py $v 2>&1 | tail -n 3
done
