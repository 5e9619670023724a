cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
for c in 1 2 3 4 5; do
  timeout 600 python bench.py --config $c --steps 100 --warmup 8 --no-cpu --no-extras > $O/cfg$c.json 2> $O/cfg$c.err
done
python - <<'P'
import json
for c in (1,2,3,4,5):
    try:
        d=json.load(open(f"gpurun_out/cfg{c}.json"))
        print(c, d["config"]["workload"][:70], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],4), "iter", round(d["roofline_iter"]["frac"],3), d["segments_per_iteration"], d["e2e"]["d2h_bytes_per_step"])
    except Exception as e:
        print(c, "FAILED", e)
P
