cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/t13.log 2>&1; tail -n 4 $O/t13.log
timeout 300 python tools/exp_walk.py v8 2>&1 | tail -n 3
timeout 900 ncu --set full --clock-control none --import-source on -s 104 -c 6 -f -o $O/r02_depth1_shared_v8 python tools/exp_one.py 250000 2 4 > $O/r02_depth1_shared_v8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 98 -c 6 -f -o $O/r02_depth0_shared_v8 python tools/exp_one.py 250000 2 4 > $O/r02_depth0_shared_v8.log 2>&1
ls -la $O/*v8*
