#!/usr/bin/env python3
"""Latency / throughput of the device-resident batch API against the batch size (headline streams: 64 KiB, q5 text).
One JSON line per n.  Geometry knobs come from the environment (BROTLI_B200_LANE_WARPS, BROTLI_B200_LANE, ...)."""
import importlib, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("rust-brotli-decompressor_b200")
corpus = importlib.import_module("tools.corpus")
pkg.lib()
sizes_arg = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 32, 256, 1024, 4096, 8192, 16384, 32768, 65536]
U = 512
comp, orig, desc = corpus.make_config("C2", U, size=65536)
tag = os.environ.get("LAT_TAG", "default")
for n in sizes_arg:
    idx = np.arange(n) % U
    sizes = np.array([len(c) for c in comp], dtype=np.uint64)
    in_off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(sizes[idx], out=in_off[1:])
    out_off = np.arange(n + 1, dtype=np.uint64) * 65536
    h_in = np.concatenate([np.frombuffer(comp[i], dtype=np.uint8) for i in idx])
    d_in = torch.from_numpy(h_in).cuda()
    d_in_off = torch.from_numpy(in_off.view(np.int64)).cuda(); d_out_off = torch.from_numpy(out_off.view(np.int64)).cuda()
    d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int64, device="cuda"); d_codes = torch.zeros(n, dtype=torch.int32, device="cuda")
    for _ in range(2):
        pkg.decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_len, d_codes)
    torch.cuda.synchronize()
    reps = 5
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        pkg.decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_len, d_codes)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    ok = bool((d_codes == 1).all())
    out = d_out.cpu().numpy()
    for j in range(0, n, max(1, n // 64)):
        ok = ok and out[j * 65536:(j + 1) * 65536].tobytes() == orig[idx[j]]
    print(json.dumps({"tag": tag, "n": n, "ms": round(ms, 3), "GBps": round(n * 65536 / ms / 1e6, 2), "bit_exact": ok}), flush=True)
    del d_in, d_out
