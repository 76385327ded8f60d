#!/usr/bin/env python3
"""Distribution of the lane kernel's cp.async waits and of its round time (BD_LANE_WAIT_HIST build of the library, selected with
BROTLI_B200_LIB): every warp times its three waits per round with clock64().  One JSON line per site.
    BROTLI_B200_LIB=$PWD/rust-brotli-decompressor_b200/variants/libbrotli_b200_whist.so python profiles/gpu_wait_hist.py [n]"""
import ctypes, importlib, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("rust-brotli-decompressor_b200")
corpus = importlib.import_module("tools.corpus")
lib = pkg.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 189440
U = 1024
comp, orig, desc = corpus.make_config("C2", U, size=65536)
idx = np.arange(n) % U
sizes = np.array([len(c) for c in comp], dtype=np.uint64)
in_off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(sizes[idx], out=in_off[1:])
out_off = np.arange(n + 1, dtype=np.uint64) * 65536
h_in = np.concatenate([np.frombuffer(comp[i], dtype=np.uint8) for i in idx])
d_in = torch.from_numpy(h_in).cuda()
d_in_off = torch.from_numpy(in_off.view(np.int64)).cuda(); d_out_off = torch.from_numpy(out_off.view(np.int64)).cuda()
d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device="cuda")
d_len = torch.zeros(n, dtype=torch.int64, device="cuda"); d_codes = torch.zeros(n, dtype=torch.int32, device="cuda")
buf = (ctypes.c_ulonglong * (4 * 64 + 8))()
probe = ctypes.CDLL(pkg.LIB_PATH).BrotliB200ProbeWaitHist
probe.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
for _ in range(2):
    pkg.decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_len, d_codes)
torch.cuda.synchronize()
assert probe(buf)  # clear
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); pkg.decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_len, d_codes); e1.record(); torch.cuda.synchronize()
assert probe(buf)
ok = bool((d_codes == 1).all())
h = np.array(list(buf), dtype=np.float64)  # [0..255] counts, [256..259] cycle sums
names = ["wait at the top of the round (next-A look-ahead)", "wait before phase P (copy chunk)", "wait before C1 (distance look-ahead)", "whole round"]
def edge(b):  # lower edge of half-octave bucket b, in cycles
    lg, half = b // 2, b % 2
    return (1 << lg) * (1.5 if half and lg > 0 else 1.0) - 1
print(json.dumps({"n": n, "ms": round(e0.elapsed_time(e1), 2), "all_decoded": ok, "note": "cycles at 1.965 GHz; instrumented build (clock64 + atomics)"}))
for s in range(4):
    c = h[s * 64:(s + 1) * 64]; tot = c.sum(); cum = np.cumsum(c)
    pct = {p: edge(int(np.searchsorted(cum, tot * p / 100.0))) for p in (50, 75, 90, 95, 99, 99.9)}
    over = {str(t): round(float(c[[b for b in range(64) if edge(b) >= t]].sum() / max(tot, 1)), 4) for t in (256, 1024, 2048, 4096, 8192)}
    # share of the site's total cycles spent in waits at least that long (bucket lower edges as weights: a lower bound)
    w = np.array([max(edge(b), 0) for b in range(64)]) * c
    share = {str(t): round(float(w[[b for b in range(64) if edge(b) >= t]].sum() / max(w.sum(), 1)), 3) for t in (1024, 2048, 4096, 8192)}
    print(json.dumps({"site": names[s], "events": int(tot), "mean_cycles": round(float(h[256 + s] / max(tot, 1)), 1),
                      "percentile_lower_edge_cycles": pct, "fraction_of_events_at_least": over, "share_of_cycles_in_events_at_least": share}))
