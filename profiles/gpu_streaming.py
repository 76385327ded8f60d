#!/usr/bin/env python3
"""Wall time of feeding one multi-metablock stream through BrotliDecoderDecompressStream in small pieces, with the
device-resident session (ResumeState) and with the re-submission path (BROTLI_B200_STREAM_SESSION=0): run once per
setting, prints one JSON line."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("rust-brotli-decompressor_b200")
corpus = importlib.import_module("tools.corpus")
pool = corpus.text_pool()
data = (pool * 4)[:4000000]
comp = corpus.compress(data, 2)
pkg.brotli_decode(comp[:1000], 4096)  # initialise the device context
res = {"session": os.environ.get("BROTLI_B200_STREAM_SESSION", "default(1)"), "compressed": len(comp), "decompressed": len(data)}
for in_chunk in (65536, 8192):
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
    res["chunk_%d" % in_chunk] = {"calls": n_calls, "wall_ms": round(dt * 1e3, 1), "kernel_ms": round(kt["exact_ms"] + kt["lane_ms"], 1)}
print(json.dumps(res))
