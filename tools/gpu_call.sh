cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_prims.py tests/test_gpu_parity.py tests/test_gpu_pipe.py -m gpu -x -q > $O/t15.log 2>&1; tail -n 3 $O/t15.log
timeout 300 python tools/exp_walk.py counted 2>&1 | tail -n 3
B2PT_SORT_COUNTED=0 timeout 300 python tools/exp_walk.py lookback 2>&1 | tail -n 3
timeout 300 python tools/exp_two_ships.py 2>&1 | tail -n 4
