cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/t5.log 2>&1; tail -3 $O/t5.log
for v in v3_ticket v4 v4_L3 v4_L4 v4_g8 v4_r4 v4_r12; do
  B2PT_LIB=$PWD/build/variants/libb2pt_$v.so timeout 200 python tools/exp_walk.py $v 2>&1 | tail -3
done
