#!/bin/bash
# Round-end measurements on one GPU (run under gpurun): bench line, ncu launch list of the same command,
# the five BASELINE configurations, two other stand-in mesh sizes.  Outputs in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
python bench.py --steps 200 --warmup 10 > $O/final_n1.json 2> $O/final_n1.err
python bench.py --steps 2 --warmup 3 --no-cpu > $O/final_plain.json 2> $O/final_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/final_ncu_bench.json 2> $O/final_ncu_bench.err
python tools/exp_configs.py > $O/final_configs.jsonl 2> $O/final_configs.err
python bench.py --steps 100 --warmup 5 --no-cpu --triangles 50000 > $O/final_50k.json 2> $O/final_50k.err
python bench.py --steps 100 --warmup 5 --no-cpu --triangles 1000000 > $O/final_1m.json 2> $O/final_1m.err
python - <<'P'
import json
for f in ("final_n1", "final_50k", "final_1m"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4),
              "single", round(d["e2e_single_context"]["value"], 1), "host", d.get("e2e_reference_host", {}).get("ms_per_call"))
    except Exception as e:
        print(f, "FAILED", e)
P
cat $O/final_configs.jsonl | cut -c1-160
