#!/usr/bin/env python
"""Static schedule of a kernel's SASS: decodes the per-instruction stall counts (control bits [105:109)) from
`cuobjdump -sass` and sums them between two addresses, with an opcode histogram. Usage:
  python tools/sass_sched.py file.o 'mangled_kernel_name' [start_addr end_addr]"""
import collections
import re
import subprocess
import sys

obj, fun = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout.split("\n")
ins = []
i = 0
while i < len(txt):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", txt[i])
    if m and i + 1 < len(txt):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", txt[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ins.append((int(m.group(1), 16), m.group(2), (hi >> 41) & 0xF))
            i += 2
            continue
    i += 1
lo, hi_ = (int(sys.argv[3], 16), int(sys.argv[4], 16)) if len(sys.argv) > 4 else (0, 1 << 30)
sel = [x for x in ins if lo <= x[0] <= hi_]
hist = collections.Counter()
stall = collections.Counter()
for a, t, st in sel:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("MUFU", "F2FP", "SYNCS", "BAR")) else "")
    hist[op] += 1
    stall[op] += st
print(f"{len(sel)} instructions, static stall sum {sum(x[2] for x in sel)} cycles")
for op, n in hist.most_common(25):
    print(f"  {op:20s} n={n:4d}  stall_sum={stall[op]:5d}  avg={stall[op] / n:.2f}")
if "--list" in sys.argv:
    for a, t, st in sel:
        print(f"{a:05x} st={st:2d} {t[:100]}")
