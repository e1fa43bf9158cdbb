# SPDX-License-Identifier: MIT
"""Print the key counters of an .ncu-rep (raw page) and, optionally, the hottest source lines."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "lts__t_bytes.sum.per_second", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print(f"  {h} [{u}] = {v}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    # find header row
    for i, r in enumerate(rows):
        if "Source" in r and any("Sampl" in c for c in r):
            hdr = r; body = rows[i + 1:]; break
    else:
        print("no source page"); sys.exit()
    si = hdr.index("Source"); ci = [k for k, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
    ci = ci[0] if ci else None
    li = hdr.index("#") if "#" in hdr else 0
    def num(x):
        try: return float(x)
        except: return 0.0
    body = [r for r in body if len(r) == len(hdr)]
    tot = sum(num(r[ci]) for r in body) or 1
    top = sorted(body, key=lambda r: -num(r[ci]))[: int(sys.argv[2])]
    for r in top:
        print(f"{100*num(r[ci])/tot:5.1f}%  L{r[li]:>4}  {r[si].strip()[:140]}")
