"""Opcode histogram (and loop structure) of one kernel's SASS: python tools/sass_hist.py <kernel substring> [lib]."""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "mygpuraytracer_b200/libb2pt.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if sys.argv[1] not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", b)
    ops = collections.Counter(o.split(".")[0] for _a, o in ins)
    print(f"{name}: {len(ins)} instructions")
    print("  " + "  ".join(f"{k} {v}" for k, v in ops.most_common(40)))
    # backward branches = loops
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?BRA[A-Z.]*\s+(?:[A-Z!0-9]+,\s*)?`?\(?\.?L_x_\d+\)?|/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?BRA\s+0x([0-9a-f]+)", b):
        pass
    brs = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?BRA[.A-Z]*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", b)
    loops = [(int(t, 16), int(a, 16)) for a, t in brs if int(t, 16) <= int(a, 16)]
    for t, a in sorted(loops):
        n = sum(1 for x, _o in ins if t <= int(x, 16) <= a)
        print(f"  loop {t:#06x}..{a:#06x}: {n} instructions")
