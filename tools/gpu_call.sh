cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_prims.py tests/test_gpu_parity.py -m gpu -x -q > $O/t14.log 2>&1; tail -n 3 $O/t14.log
timeout 300 python tools/exp_walk.py v9 2>&1 | tail -n 3
