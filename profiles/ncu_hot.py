#!/usr/bin/env python3
"""Summarise an .ncu-rep of the decode kernel: headline metrics, instructions per command and the
hottest source lines (executed instructions and stall samples).  Usage: ncu_hot.py REPORT [commands]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
ncmd = float(sys.argv[2]) if len(sys.argv) > 2 else None
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], stdout=subprocess.PIPE, text=True).stdout
keys = ("Duration", "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Eligible Warps", "Warp Cycles Per Issued", "Registers Per",
        "Achieved Occupancy", "Theoretical Occupancy", "L1/TEX Hit", "L2 Hit", "DRAM Throughput", "Mem Busy", "Avg. Active Threads", "Branch Eff",
        "SM Frequency", "Local Load", "Shared Load", "Bank conflict")
for line in det.splitlines():
    if any(k in line for k in keys):
        print(line.rstrip())
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
seen, cur, curfile = {}, None, None
for r in rows:
    if len(r) == 2:
        if r[0] == "File Path":
            curfile = r[1].split("/")[-1]
        continue
    if r and r[0].isdigit():
        cur = (curfile, int(r[0]), r[1].strip()[:90])
        continue
    if r and r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
        a = int(r[2], 16)
        if a not in seen:
            seen[a] = (int(r[7]) if r[7].isdigit() else 0, int(r[6]) if r[6].isdigit() else 0, cur, r[3].strip())
tot = sum(v[0] for v in seen.values())
tots = sum(v[1] for v in seen.values())
print("warp instructions executed: %.4g   stall samples: %d" % (tot, tots))
if ncmd:
    print("instructions per command: %.1f" % (tot / ncmd))
agg = collections.OrderedDict()
for a in sorted(seen):
    c, s, cur, _ = seen[a]
    k = cur
    x = agg.setdefault(k, [0, 0])
    x[0] += c; x[1] += s
print("%-7s %-7s  source line" % ("inst%", "stall%"))
for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%6.2f%% %6.2f%%  %s:%d  %s" % (100.0 * c / tot, 100.0 * s / max(tots, 1), k[0], k[1], k[2]))
if len(sys.argv) > 3:  # dump hot SASS in address order
    thr = float(sys.argv[3])
    for a in sorted(seen):
        c, s, cur, sass = seen[a]
        if ncmd and c / ncmd >= thr:
            print("%06x %6.2f s%6d %s:%d  %s" % (a & 0xFFFFFF, c / ncmd, s, cur[0][:10], cur[1], sass[:80]))
