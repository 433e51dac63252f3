#!/usr/bin/env python
"""Stall-reason summary of an ncu report's source page (SASS level).

    python tools/ncu_stalls.py report.ncu-rep [top N]

Prints the kernel's warp-stall samples by reason, then the N SASS instructions with the most samples (with their
dominant reasons), so that a profile can be read without the GUI."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
ns = col["# Samples"]
for r in data:
    for s in stalls:
        tot[s] += int(r[col[s]] or 0)
allsm = sum(tot.values())
print(f"# {rep}: {len(data)} SASS instructions, {allsm} stall samples")
for s, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {s:26s} {v:8d} {100.0 * v / allsm:5.1f} %")
print(f"\n# top {top} instructions by samples (index, samples, executed, instruction, dominant reasons)")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ns] or 0))[:top]
for i in sorted(order):
    r = data[i]
    rs = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{i:5d} {int(r[ns] or 0):6d} {r[col['Instructions Executed']]:>9s}  {r[col['Source']].strip()[:70]:70s} " +
          " ".join(f"{n}:{v}" for v, n in rs if v))
