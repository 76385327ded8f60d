#!/usr/bin/env python3
"""Streaming sessions on the GPU:
  (1) one multi-metablock stream fed through BrotliDecoderDecompressStream in 64 KiB / 8 KiB / 1 KiB pieces;
  (2) N concurrent sessions (default 1000), each fed 8 KiB of compressed input per round through
      BrotliB200DecoderDecompressStreamBatch -- one decode launch per round -- checked call by call against the streaming
      oracle on a sample of the sessions; reports aggregate decompressed GB/s (wall clock, Python driver included).
One JSON line per experiment."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
pkg = importlib.import_module("rust-brotli-decompressor_b200")
corpus = importlib.import_module("tools.corpus")
pool = corpus.text_pool()
pkg.brotli_decode(corpus.compress(pool[:1000], 5), 4096)  # initialise the device context

data = (pool * 4)[:4000000]
comp = corpus.compress(data, 2)
res = {"experiment": "one stream in pieces", "compressed": len(comp), "decompressed": len(data)}
for in_chunk in (65536, 8192, 1024):
    st = pkg.DecoderState()
    pos, r, n_calls, got = 0, 2, 0, 0
    pkg.kernel_times(reset=True)
    t0 = time.perf_counter()
    while r not in (0, 1):
        r, used, out = st.decompress_stream(comp[pos:pos + in_chunk], 1 << 20)
        pos += used; got += len(out); n_calls += 1
    dt = time.perf_counter() - t0
    kt = pkg.kernel_times()
    assert r == 1 and got == len(data)
    st.close()
    res["chunk_%d" % in_chunk] = {"calls": n_calls, "wall_ms": round(dt * 1e3, 1), "kernel_ms_last_64_launches": round(kt["exact_ms"] + kt["lane_ms"], 1),
                                  "MBps": round(len(data) / dt / 1e6, 1)}
print(json.dumps(res), flush=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
piece = 8192
rng = np.random.default_rng(7)
U = 64
datas = [bytes(pool[int(a):int(a) + 400000]) for a in rng.integers(0, len(pool) - 400000, size=U)]
comps = [corpus.compress(d, 5) for d in datas]
oracle = helpers.Oracle()
check = list(range(0, n, max(1, n // 16)))  # sessions followed call by call by the streaming oracle
ostreams = {i: oracle.stream() for i in check}
states = [pkg.DecoderState() for _ in range(n)]
pos = [0] * n; pend = [b""] * n; res_code = [2] * n; total = 0
launch0 = pkg.kernel_launch_count()
rounds = 0
round_ms = []
t0 = time.perf_counter()
t_oracle = 0.0
while any(r in (2, 3) for r in res_code):
    idx = [i for i in range(n) if res_code[i] in (2, 3)]
    for i in idx:
        if res_code[i] == 2:
            c = comps[i % U]
            pend[i] = c[pos[i]:pos[i] + piece]; pos[i] += len(pend[i])
    tr = time.perf_counter()
    got = pkg.decompress_stream_batch([states[i] for i in idx], [pend[i] for i in idx], [1 << 17] * len(idx))
    round_ms.append(round((time.perf_counter() - tr) * 1e3, 1))
    for i, g in zip(idx, got):
        if i in ostreams:
            t1 = time.perf_counter()
            e = ostreams[i].call(pend[i], 1 << 17)
            t_oracle += time.perf_counter() - t1
            assert g == e, (i, rounds, g[:2], e[:2])
        res_code[i] = g[0]; pend[i] = pend[i][g[1]:]; total += len(g[2])
    rounds += 1
dt = time.perf_counter() - t0 - t_oracle
assert all(r == 1 for r in res_code) and total == sum(len(datas[i % U]) for i in range(n))
print(json.dumps({"experiment": "multiplexed sessions", "sessions": n, "piece_bytes": piece, "rounds": rounds, "decompressed_GB": round(total / 1e9, 3),
                  "wall_s": round(dt, 3), "GBps": round(total / dt / 1e9, 3), "kernel_launches": pkg.kernel_launch_count() - launch0,
                  "oracle_checked_sessions": len(check), "call_by_call_equal": True, "call_ms_per_round": round_ms}), flush=True)
for s in states:
    s.close()
