#!/usr/bin/env python3
"""Kernel time of one batch of small payloads compressed against ONE shared custom dictionary (the shared-dictionary
serving case) through BrotliB200DecompressBatchPackedWithDictionary: lane kernel's dictionary instance (default
geometry) vs exact kernel only (run with BROTLI_B200_LANE=0).  Prints one JSON line; bit-exact against the originals."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("rust-brotli-decompressor_b200")
corpus = importlib.import_module("tools.corpus")
pool = corpus.text_pool()
d = pool[500000:560000]
rng = np.random.default_rng(5)
U, N = 256, 65536
# a payload: three pieces of the shared document interleaved with three pieces of new text
def payload():
    parts = []
    for _ in range(3):
        parts.append(pool[500000 + int(rng.integers(0, 58000)):][:int(rng.integers(600, 2000))])
        parts.append(pool[int(rng.integers(0, 400000)):][:int(rng.integers(600, 2000))])
    return (b"".join(parts) + pool[:8192])[:8192]
origs = [payload() for _ in range(U)]
comp = [corpus.compress_with_dictionary(o, d, 5) for o in origs]
idx = rng.integers(0, U, size=N)
in_bytes, in_off = corpus.pack([comp[i] for i in idx])
out_off = np.arange(N + 1, dtype=np.uint64) * np.uint64(8192)
out_bytes = np.zeros(N * 8192 + 1, dtype=np.uint8)
out_len = np.zeros(N, dtype=np.uint64); codes = np.zeros(N, dtype=np.int32)
for _ in range(2):
    pkg.decompress_batch_packed_custom_dict(in_bytes, in_off, out_bytes, out_off, out_len, codes, d)
pkg.kernel_times(reset=True)
for _ in range(3):
    pkg.decompress_batch_packed_custom_dict(in_bytes, in_off, out_bytes, out_off, out_len, codes, d)
kt = pkg.kernel_times()
ok = bool((codes == 1).all()) and all(out_bytes[j * 8192:(j + 1) * 8192].tobytes() == origs[idx[j]] for j in range(0, N, 97))
ms = (kt["lane_ms"] + kt["exact_ms"]) / 3
print(json.dumps({"lane": os.environ.get("BROTLI_B200_LANE", "1"), "streams": N, "payload_bytes": 8192, "dictionary_bytes": len(d),
                  "compressed_ratio": round(float(in_off[-1]) / (N * 8192), 4), "kernel_ms": round(ms, 2), "lane_ms": round(kt["lane_ms"] / 3, 2),
                  "exact_ms": round(kt["exact_ms"] / 3, 2), "bailed_to_exact": kt["bailed"], "kernel_GBps": round(N * 8192 / ms / 1e6, 2), "bit_exact": ok}))
