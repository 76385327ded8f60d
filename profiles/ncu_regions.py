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
votes = [a for a in addrs if seen[a][0][3].strip().startswith("VOTE.ANY")]
top, end = votes[1], votes[2]
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
region("command loop", lambda a: top <= a <= end)
region("per-metablock code + rest", lambda a: a < top or a > end)
agg = collections.defaultdict(lambda: [0.0, 0.0])
for a in addrs:
    if a < top or a > end:
        r, c = seen[a]
        agg[c][0] += f(r, "Instructions Executed"); agg[c][1] += f(r, "# Samples")
print("hottest lines outside the command loop (inst %, time %):")
for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:nlines]:
    print("%6.2f %6.2f  %s:%s  %s" % (100 * c / tot_i, 100 * s / tot_s, k[0] if k else "?", k[1] if k else "?", k[2] if k else ""))
