"""Per-kernel totals of an `ncu --metrics ... --csv` log: launches, time, warp instructions, lanes, issue-slot use.
    python tools/ncu_inst.py gpurun_out/r02_inst_shared.csv [first_id last_id]"""
import collections, csv, io, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(io.StringIO("".join(rows)))
per = collections.OrderedDict()
for r in rd:
    i = int(r["ID"])
    per.setdefault(i, {"name": r["Kernel Name"]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
ids = sorted(per)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (ids[0], ids[-1])
agg = collections.OrderedDict()
for i in ids:
    if not lo <= i <= hi: continue
    k = per[i]; n = k["name"].split("(")[0].replace("b2pt::", "").replace("void ", "")
    a = agg.setdefault(n, dict(n=0, us=0.0, inst=0.0, thr=0.0, issue=0.0))
    a["n"] += 1; a["us"] += k["gpu__time_duration.sum"] / 1e3; a["inst"] += k["smsp__inst_executed.sum"]
    a["thr"] += k["smsp__thread_inst_executed_per_inst_executed.ratio"] * k["smsp__inst_executed.sum"]
    a["issue"] += k["smsp__issue_active.avg.pct_of_peak_sustained_active"] * k["gpu__time_duration.sum"]
tot_i = sum(a["inst"] for a in agg.values()); tot_t = sum(a["us"] for a in agg.values())
print(f"| kernel | launches | us | % time | M warp inst | % inst | lanes/inst | issue % |\n|---|---:|---:|---:|---:|---:|---:|---:|")
for n, a in agg.items():
    print(f"| `{n}` | {a['n']} | {a['us']:.1f} | {100*a['us']/tot_t:.1f} | {a['inst']/1e6:.2f} | {100*a['inst']/tot_i:.1f} | {a['thr']/max(a['inst'],1):.1f} | {a['issue']/max(a['us']*1e3,1):.1f} |")
print(f"| total | | {tot_t:.1f} | | {tot_i/1e6:.1f} | | | |")
