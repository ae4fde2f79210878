#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS opcodes in the built library (`cuobjdump -sass`): tcgen05 MMA
(UTCHMMA, .2CTA = cta_group::2), TMEM loads/stores (LDTM/STTM), TMA (UTMALDG/UTMASTG/UBLKCP), tcgen05 commit barriers
(UTCBAR), mbarrier waits (SYNCS), cluster ops (UCGABAR), MUFU.EX2. Usage:
  python tools/sass_opcodes.py [frameino_b200/libframeino_b200.so] > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "frameino_b200/libframeino_b200.so"
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "MUFU.EX2",
       "HMMA", "FFMA2"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
per = collections.OrderedDict()
cur = None
for line in txt.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                if o == "UTCHMMA" and ".2CTA" in op:
                    continue  # counted in its own column
                cur[o] += 1
print(f"# {lib}: {len(per)} kernels; columns = instruction counts in the SASS of each kernel (static, not executed counts)")
print("# kernels with no tcgen05 / TMA opcode are the HBM-bound row kernels and helpers")
hdr = ["total"] + OPS
print(" | ".join(f"{h:>12s}" for h in hdr) + " | kernel")
tot = collections.Counter()
for name, c in per.items():
    row = [c["_total"]] + [c[o] for o in OPS]
    for o in OPS:
        tot[o] += c[o]
    d = demangle(name)
    d = re.sub(r"\s+", " ", d)
    print(" | ".join(f"{v:12d}" for v in row) + " | " + d[:150])
print("# totals: " + ", ".join(f"{o}={tot[o]}" for o in OPS))
