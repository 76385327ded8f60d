#!/usr/bin/env python3
"""Throughput of the device-resident batch API on the other BASELINE configs (C3, C5, C4-like), bit-exact
against the originals; prints one JSON line per config.  Not the headline bench (bench.py)."""
import importlib, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("rust-brotli-decompressor_b200")
corpus = importlib.import_module("tools.corpus")
pkg.lib()
ONLY = [c for c in os.environ.get("CONFIGS", "").split(",") if c]  # e.g. CONFIGS=C3,C5
for cfg, n_unique, n, size in (("C3", 4096, 1 << 20, 4096), ("C5", 2200, 131072, 65536), ("C4", 8, 2048, 4 << 20)):
    if ONLY and cfg not in ONLY:
        continue
    comp, orig, desc = corpus.make_config(cfg, n_unique, size=size)
    idx = np.random.default_rng(1).integers(0, n_unique, size=n)
    sizes = np.array([len(c) for c in comp], dtype=np.uint64)
    osz = np.array([len(o) for o in orig], dtype=np.uint64)
    in_off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(sizes[idx], out=in_off[1:])
    out_off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(osz[idx], out=out_off[1:])
    h_in = np.concatenate([np.frombuffer(comp[i], dtype=np.uint8) for i in idx])
    d_in = torch.from_numpy(h_in).cuda()
    d_in_off = torch.from_numpy(in_off.view(np.int64)).cuda(); d_out_off = torch.from_numpy(out_off.view(np.int64)).cuda()
    d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int64, device="cuda"); d_codes = torch.zeros(n, dtype=torch.int32, device="cuda")
    for _ in range(2):
        pkg.decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_len, d_codes)
    torch.cuda.synchronize()
    pkg.kernel_times(reset=True)
    t0 = time.perf_counter()
    for _ in range(3):
        pkg.decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_len, d_codes)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    kt = pkg.kernel_times()
    ok = bool((d_codes == 1).all())
    out = d_out.cpu().numpy()
    for j in range(0, n, max(1, n // 512)):
        ok = ok and out[int(out_off[j]):int(out_off[j + 1])].tobytes() == orig[idx[j]]
    print(json.dumps({"config": cfg, "desc": desc, "streams": n, "decompressed_GB": round(float(out_off[-1]) / 1e9, 3),
                      "GBps": round(float(out_off[-1]) / dt / 1e9, 2), "ms": round(dt * 1e3, 2), "lane_ms": round(kt["lane_ms"] / 3, 2),
                      "exact_ms": round(kt["exact_ms"] / 3, 2), "bailed_to_exact": kt["bailed"], "bit_exact": ok}), flush=True)
    del d_in, d_out
