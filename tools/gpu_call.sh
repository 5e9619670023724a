cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/t8.log 2>&1; tail -8 $O/t8.log
timeout 300 python tools/exp_walk.py head 2>&1 | tail -3
