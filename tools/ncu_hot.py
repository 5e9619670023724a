"""Hot source lines of an `ncu --page source --csv --print-source cuda,sass` dump, per kernel."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
kern = f = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': kern = r[1][:40]; continue
    if r[0] == 'Line No': continue
    if r[0] != '' and len(r) > 8:
        try: ln = int(r[0]); samples = int(r[4]); inst = int(r[7]); tinst = int(r[8])
        except ValueError: continue
        agg.setdefault(kern, []).append((samples, inst, tinst, f, ln, r[1][:100]))
for k, v in agg.items():
    ts = sum(x[0] for x in v) or 1; ti = sum(x[1] for x in v) or 1
    print(k, 'samples', ts, 'inst', ti)
    for x in sorted(v, reverse=True)[:top]:
        print(f"  {100*x[0]/ts:5.1f}% smp {100*x[1]/ti:5.1f}% inst thr/inst {x[2]/max(x[1],1):5.1f} {x[3]}:{x[4]} {x[5]}")
