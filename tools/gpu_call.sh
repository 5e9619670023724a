cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
for k in 4 5 6; do python bench.py --steps 200 --warmup 10 --no-cpu --no-extras --streams $k > $O/lanes$k.json 2> $O/lanes$k.err; python - $k <<'P'
import json,sys
d=json.load(open(f"gpurun_out/lanes{sys.argv[1]}.json")); print("lanes",sys.argv[1],round(d["value"],1),round(d["ms_per_step"],4),"e2e",round(d["e2e"]["value"],1),round(d["e2e"]["ms_per_step"],4), d["windows_ms"]["e2e"])
P
done
python bench.py --steps 50 --warmup 5 --no-cpu > $O/issue_check.json 2> $O/issue_check.err; python - <<'P'
import json
d=json.load(open("gpurun_out/issue_check.json")); print(json.dumps(d.get("issue_iter"))); print({k:(round(v["frac"],4), round(v.get("issue",{}).get("frac",0),3)) for k,v in d["roofline_kernels"].items()}); print(json.dumps(d["roofline"])[:700])
P
