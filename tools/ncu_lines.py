# SPDX-License-Identifier: MIT
"""Warp-stall samples per CUDA source line from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ni, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
st = [k for k, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
tot, lines, reasons = 0, [], {}
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[0]:
        continue
    try:
        n = float(r[ni])
    except ValueError:
        continue
    tot += n
    for k in st:
        reasons[hdr[k][6:]] = reasons.get(hdr[k][6:], 0) + float(r[k] or 0)
    top = sorted(((float(r[k] or 0), hdr[k][6:]) for k in st), reverse=True)[:2]
    lines.append((n, int(r[0]), r[1].strip()[:100], float(r[ii] or 0), top))
rs = sum(reasons.values()) or 1
print("stall reasons:", ", ".join(f"{k} {100*v/rs:.1f}" for k, v in sorted(reasons.items(), key=lambda kv: -kv[1]) if v / rs > 0.01))
for n, ln, src, ins, top in sorted(lines, reverse=True)[:topn]:
    print(f"{100*n/tot:5.1f}% L{ln:4d} inst={ins/1e6:8.1f}M {top[0][1]:>10}/{top[1][1]:<10} {src}")
