# scratch: the command of the last ad-hoc gpurun call of the session (kept so that the calls in profiles/r02_notes.md can be repeated)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; tail -n 4 $O/tests.log
timeout 300 python tools/exp_walk.py head 2>&1 | tail -n 3
