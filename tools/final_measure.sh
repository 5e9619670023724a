#!/bin/bash
# Round-end measurements on one GPU (run under gpurun): the bench line, the ncu launch list of the same command,
# warp-instruction counts per kernel, the five BASELINE configurations.  Outputs in gpurun_out/ (tag = $1, default r02).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
T=${1:-r02}
python bench.py --steps 200 --warmup 10 > $O/${T}_final_n1.json 2> $O/${T}_final_n1.err
python bench.py > $O/${T}_final_default.json 2> $O/${T}_final_default.err
python bench.py --steps 2 --warmup 3 --no-cpu > $O/${T}_final_plain.json 2> $O/${T}_final_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${T}_final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/${T}_final_ncu_bench.json 2> $O/${T}_final_ncu_bench.err
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread
ncu --metrics $M --clock-control none --csv --log-file $O/${T}_final_inst_shared.csv python tools/exp_one.py 250000 2 4 > $O/${T}_final_inst_shared.log 2>&1
python tools/exp_configs.py > $O/${T}_final_configs.jsonl 2> $O/${T}_final_configs.err
python tools/exp_walk.py $T 2>&1 | tail -n 3 > $O/${T}_final_walk.txt
python - "$T" <<'P'
import json, sys
T = sys.argv[1]
for f in (f"{T}_final_n1", f"{T}_final_default"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4),
              "single", round(d["e2e_single_context"]["value"], 1), "host", d.get("e2e_reference_host", {}).get("ms_per_call"),
              "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), "iter", round(d["roofline_iter"]["frac"], 3),
              "ref_gpu", (d.get("reference_gpu") or {}).get("ms_per_call_median"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
P
cat $O/${T}_final_walk.txt
cut -c1-170 $O/${T}_final_configs.jsonl
