#!/usr/bin/env python3
"""Builds and runs profiles/probes/session_bench.cpp (multiplexed streaming sessions through the C ABI, C++ driver):
   python profiles/gpu_sessions.py [sessions=1000] [piece=8192] [stream_bytes=400000]"""
import importlib, os, struct, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
corpus = importlib.import_module("tools.corpus")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
piece = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
size = int(sys.argv[3]) if len(sys.argv) > 3 else 400000
pool = corpus.text_pool()
rng = np.random.default_rng(7)
U = 64
datas = [bytes(pool[int(a):int(a) + size]) for a in rng.integers(0, len(pool) - size, size=U)]
comps = [corpus.compress(d, 5) for d in datas]
path = "/tmp/sess_streams.bin"
with open(path, "wb") as f:
    f.write(struct.pack("<I", U))
    for c, d in zip(comps, datas):
        f.write(struct.pack("<II", len(c), len(d)))
    for c in comps:
        f.write(c)
pkg_dir = os.path.join(ROOT, "rust-brotli-decompressor_b200")
exe = "/tmp/session_bench"
subprocess.check_call(["g++", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "profiles", "probes", "session_bench.cpp"), "-L", pkg_dir,
                       "-l:libbrotli_b200.so", "-Wl,-rpath," + pkg_dir, "-o", exe])
for cap in (1 << 17,):
    subprocess.check_call([exe, path, str(n), str(piece), str(cap)])
