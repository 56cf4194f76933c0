#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed): headline metrics, dynamic SASS op mix and the
hottest source lines, normalised per output pixel.  Usage: ncu_summary.py prof.ncu-rep [pixels_per_launch]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
npx = float(sys.argv[2]) if len(sys.argv) > 2 else 3840 * 2160
W = npx / 32.0

def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout

raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, data = raw[0], raw[1], raw[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers",
        "sm__maximum_warps_per_active_cycle_pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct"]
for r in data[:1]:
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print("%-75s %s %s" % (w, r[i], units[i]))
    for i, h in enumerate(hdr):
        if h == "smsp__inst_executed.sum":
            print("warp instructions per pixel: %.1f" % (float(r[i]) / W))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; ci = h.index("Instructions Executed"); si = h.index("Source")
tot = 0; byop = collections.Counter()
for r in rows[hi + 1:]:
    try: c = int(r[ci])
    except Exception: continue
    t = r[si].split()
    if not t: continue
    op = t[1] if t[0].startswith("@") else t[0]
    tot += c; byop[op.split(".")[0]] += c
print("\ndynamic SASS mix (warp-instructions per pixel), total %.1f:" % (tot / W))
print("  " + "  ".join("%s:%.1f" % (op, c / W) for op, c in byop.most_common(40)))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
cur = None; out = []; seen = set()
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] not in ("", "Line No") and r[2] == "-":
        try: c = int(r[7])
        except Exception: continue
        if (cur, r[0]) in seen: continue
        seen.add((cur, r[0])); out.append((c, cur, r[0], r[1].strip()[:110]))
print("\nhottest source lines (instructions per pixel; inlined lines are attributed at every level):")
for c, f, l, s in sorted(out, reverse=True)[:40]:
    print("%7.1f %s:%s  %s" % (c / W, f, l, s))
