cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/t16.log 2>&1; tail -n 4 $O/t16.log
timeout 300 python tools/exp_walk.py side 2>&1 | tail -n 3
B2PT_SIDE_BY_SIDE=0 timeout 300 python tools/exp_walk.py seq 2>&1 | tail -n 3
