# SPDX-License-Identifier: MIT
"""Aggregate warp-stall samples of an .ncu-rep: totals per stall reason and per CUDA source line
(`--print-source cuda`) or the top SASS instructions."""
import csv, subprocess, sys
rep, view, topn = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "cuda"), int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)  # sass view only
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def num(x):
    try: return float(x)
    except ValueError: return 0.0
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot = {c: sum(num(r[hdr.index(c)]) for r in body) for c in stalls}
allsamp = sum(tot.values()) or 1
print("stall reasons (% of samples):", ", ".join(f"{c[6:]} {100*v/allsamp:.1f}" for c, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / allsamp > 0.005))
si, ni = hdr.index("Source"), hdr.index("# Samples")
key0 = hdr[0]
ns = sum(num(r[ni]) for r in body) or 1
for r in sorted(body, key=lambda r: -num(r[ni]))[:topn]:
    top = sorted(((num(r[hdr.index(c)]), c[6:]) for c in stalls), reverse=True)[:2]
    print(f"{100*num(r[ni])/ns:5.1f}%  {r[0][-6:]:>6}  {r[si].strip()[:110]:110s} {top[0][1]}/{top[1][1]}")
