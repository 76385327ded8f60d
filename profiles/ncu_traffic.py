#!/usr/bin/env python3
"""profiles/current_traffic.json from the raw-metrics CSV of one `ncu --set full` capture of the dominant kernel:
DRAM bytes per launch, the launch shape it was taken at, and the fingerprint of the kernel sources (bench.py quotes the
capture only for the code it was taken from).  Usage: ncu_traffic.py RAW.csv STREAMS_PER_LAUNCH KERNEL_NAME [OUT.json]"""
import csv, datetime, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, zip(units, vals)))


def bytes_of(name):
    unit, v = m[name]
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


out = {"kernel": sys.argv[3], "streams_per_launch": int(sys.argv[2]),
       "dram_bytes_read": int(bytes_of("dram__bytes_read.sum")), "dram_bytes_write": int(bytes_of("dram__bytes_write.sum")),
       "duration_ms_under_ncu": float(m["gpu__time_duration.sum"][1].replace(",", "")) / {"ns": 1e6, "us": 1e3, "ms": 1, "s": 1e-3}.get(m["gpu__time_duration.sum"][0], 1e6),
       "kernel_source_hash": bench.kernel_source_hash(), "captured": datetime.date.today().isoformat(), "source": os.path.basename(sys.argv[1])}
path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "current_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out))
