#!/usr/bin/env python3
"""Split an .ncu-rep of the lane kernel into its command loop and the per-metablock (header) code: instruction share,
stall-sample share (= share of warp time), active threads and the top stall reasons of each, then the hottest
source lines of the header code.   python profiles/ncu_regions.py REPORT [lines]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
data, cur, curfile = [], None, None
for r in rows:
    if len(r) == 2:
        if r[0] == "File Path":
            curfile = r[1].split("/")[-1]
        continue
    if r and r[0] == "#":
        continue
    if r and r[0] == "Address" or (len(r) > 3 and r[2] == "Address"):
        hdr = r
        continue
    if r and r[0].isdigit():
        cur = (curfile, int(r[0]), r[1].strip()[:100])
        continue
    if r and r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
        data.append((int(r[2], 16), r, cur))
ix = {h: i for i, h in enumerate(hdr)} if hdr else {}
seen = {}
for a, r, c in data:
    seen.setdefault(a, (r, c))
addrs = sorted(seen)
def f(r, k, default_idx=None):
    try:
        return float(r[ix[k]])
    except (KeyError, ValueError, IndexError):
        return 0.0
# Regions by source line: the command loop is run_commands (two instances are inlined into the kernel); instructions of
# inlined helpers (lines above the per-metablock code, other files) belong to the region of the code around them.
import os
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rust-brotli-decompressor_b200", "csrc", "brotli_decode_lane.cuh")
lines = open(SRC).read().split("\n")
loop_lo = next(i + 1 for i, l in enumerate(lines) if "BD_DEV void run_commands(" in l)
loop_hi = next(i + 1 for i, l in enumerate(lines) if l.startswith("// Stream header: bit window and output cursor set-up"))
helpers_hi = next(i + 1 for i, l in enumerate(lines) if "per-metablock (cold) code" in l)
# the loop instances by address: from the first instruction of `while (warp_any(run))` to the last instruction of the
# loop's final statement, for every instance (heads further than 4 KB apart belong to different instances)
l_head = next(i + 1 for i, l in enumerate(lines) if "while (warp_any(run)) {" in l and i + 1 > loop_lo)
l_tail = next(i + 1 for i, l in enumerate(lines) if "if (run && ev != kStCommands) {" in l and i + 1 > l_head)
heads = sorted(a for a in addrs if seen[a][1] and seen[a][1][0] == "brotli_decode_lane.cuh" and seen[a][1][1] == l_head)
tails = sorted(a for a in addrs if seen[a][1] and seen[a][1][0] == "brotli_decode_lane.cuh" and l_tail <= seen[a][1][1] <= l_tail + 3)
starts = [h for i, h in enumerate(heads) if i == 0 or h - heads[i - 1] > 4096]
ranges = []
for i, h in enumerate(starts):
    nxt = starts[i + 1] if i + 1 < len(starts) else 1 << 62
    t = [x for x in tails if h < x < nxt]
    if t:
        ranges.append((h, max(t)))
region_of = {a: ("loop" if any(lo <= a <= hi for lo, hi in ranges) else "header") for a in addrs}
print("loop instances (address ranges): %s" % ", ".join("%x-%x" % (lo & 0xFFFFFF, hi & 0xFFFFFF) for lo, hi in ranges))
top, end = 0, 0
stall_keys = [k for k in ix if k.startswith("stall_") and "Not Issued" not in k]
tot_i = sum(f(seen[a][0], "Instructions Executed") for a in addrs)
tot_s = sum(f(seen[a][0], "# Samples") for a in addrs)
def region(name, sel):
    rs = [seen[a][0] for a in addrs if sel(a)]
    inst = sum(f(r, "Instructions Executed") for r in rs)
    thr = sum(f(r, "Thread Instructions Executed") for r in rs)
    samp = sum(f(r, "# Samples") for r in rs)
    st = {k: sum(f(r, k) for r in rs) for k in stall_keys}
    t = sum(st.values()) or 1
    print("%-28s inst %5.1f %%  time (stall samples) %5.1f %%  active threads %4.1f   %s" % (
        name, 100 * inst / tot_i, 100 * samp / tot_s, thr / max(inst, 1),
        ", ".join("%s %.0f%%" % (k[6:], 100 * v / t) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:6])))
region("command loop", lambda a: region_of[a] == "loop")
region("per-metablock code + rest", lambda a: region_of[a] != "loop")
agg = collections.defaultdict(lambda: [0.0, 0.0])
for a in addrs:
    if region_of[a] != "loop":
        r, c = seen[a]
        agg[c][0] += f(r, "Instructions Executed"); agg[c][1] += f(r, "# Samples")
print("hottest lines outside the command loop (inst %, time %):")
for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:nlines]:
    print("%6.2f %6.2f  %s:%s  %s" % (100 * c / tot_i, 100 * s / tot_s, k[0] if k else "?", k[1] if k else "?", k[2] if k else ""))
