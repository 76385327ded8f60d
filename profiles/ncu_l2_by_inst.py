#!/usr/bin/env python3
"""Top SASS instructions of a kernel by L2 sectors: python profiles/ncu_l2_by_inst.py <source.csv from `ncu -i X.ncu-rep --page source --csv`>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_g = sum(f(r, "L2 Theoretical Sectors Global") for r in data); tot_l = sum(f(r, "L2 Theoretical Sectors Local") for r in data)
print("total L2 theoretical sectors: global %.4g local %.4g" % (tot_g, tot_l))
top = sorted(data, key=lambda r: -(f(r, "L2 Theoretical Sectors Global") + f(r, "L2 Theoretical Sectors Local")))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for r in top:
    print(r[ix["Address"]][-6:], "%-64s" % r[ix["Source"]][:64], "inst %.3g" % f(r, "Instructions Executed"),
          "thr/inst %.1f" % f(r, "Avg. Predicated-On Threads Executed"), "L2 global %.3g" % f(r, "L2 Theoretical Sectors Global"),
          "local %.3g" % f(r, "L2 Theoretical Sectors Local"))
