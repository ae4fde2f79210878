#!/usr/bin/env python
"""Top stall sites of one kernel from an .ncu-rep captured with --import-source on (per-SASS-instruction warp-stall
samples). Usage: python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out, tot = [], 0
for r in data:
    try:
        n = int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    tot += n
    out.append((n, r[ix["Address"]], r[ix["Source"]], {s: int(r[ix[s]] or 0) for s in stalls},
                r[ix["Instructions Executed"]]))
print("total samples", tot)
agg = {s: sum(o[3][s] for o in out) for s in stalls}
print({k[6:]: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
for n, a, src, st, ie in sorted(out, key=lambda o: -o[0])[:top_n]:
    big = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{n:7d} {100 * n / tot:5.1f}% {a[-5:]} {src[:64]:64s} {ie:>10s} " +
          " ".join(f"{k[6:]}={v}" for k, v in big if v))
