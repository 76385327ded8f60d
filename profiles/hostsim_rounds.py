#!/usr/bin/env python3
"""Rounds per stream of the lane kernel's command loop, counted on the HOST build of its logic (no GPU), and what
they imply for a warp of 32 streams: lane-rounds used / (32 x the slowest lane's rounds) for batch order, for the
size-bucket order the library uses, and for an oracle order by the true round count.  Also the share of commands
that take the loop's rare paths (each is issued for the whole warp whenever one lane needs it).
    python profiles/hostsim_rounds.py [config] [n_unique]        (DESIGN.md section 6 quotes headline / 1024)"""
import ctypes, importlib, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
corpus = importlib.import_module("tools.corpus")
cfg = sys.argv[1] if len(sys.argv) > 1 else "headline"
n_unique = int(sys.argv[2]) if len(sys.argv) > 2 else 256
d = os.path.join(ROOT, "build_tmp")
os.makedirs(d, exist_ok=True)
src = os.path.join(d, "rounds_paths.cpp")
open(src, "w").write('#include <stdint.h>\nextern "C" { uint64_t g_rounds = 0; uint64_t g_cnt[16] = {0}; }\n'
                     '#define BD_LANE_ROUND_STATS(ph, pa) (g_rounds++)\n#define BD_LANE_COUNT(i) (g_cnt[i]++)\n'
                     '#include "../tests/hostsim/hostsim_lane.cpp"\n')
so = os.path.join(d, "librounds_paths.so")
subprocess.check_call(["g++", "-O2", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", so, src, os.path.join(ROOT, "tables", "brotli_dictionary.c"),
                       '-DBROTLI_DICT_PATH="%s"' % os.path.join(ROOT, "tables", "brotli_dictionary.bin")])
L = ctypes.CDLL(so)
L.hostsim_lane_decode.restype = ctypes.c_int
L.hostsim_lane_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32,
                                  ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
g = ctypes.c_uint64.in_dll(L, "g_rounds")
cnt = (ctypes.c_uint64 * 16).in_dll(L, "g_cnt")
comp, orig, _ = corpus.make_config(cfg, n_unique)
per = []
for c, o in zip(comp, orig):
    buf = ctypes.create_string_buffer(len(o) + 80)
    dd, u, g0 = ctypes.c_uint64(0), ctypes.c_uint64(0), g.value
    L.hostsim_lane_decode(c, len(c), ctypes.addressof(buf) + (-ctypes.addressof(buf)) % 8, len(o), 178, ctypes.byref(dd), ctypes.byref(u))
    per.append(g.value - g0)
per, size = np.array(per), np.array([len(c) for c in comp])
def eff(order):
    grp = per[order][:len(per) // 32 * 32].reshape(-1, 32)
    return float(grp.sum() / (32 * grp.max(axis=1).sum()))
print("%s, %d streams: rounds per stream mean %.0f std %.0f; corr(rounds, compressed size) %.3f" % (cfg, len(per), per.mean(), per.std(), np.corrcoef(per, size)[0, 1]))
print("lane-round efficiency of a 32-stream warp: batch order %.3f, size buckets (library) %.3f, exact size %.3f, oracle (true rounds) %.3f" % (
    eff(np.arange(len(per))), eff(np.argsort(-(size >> 8), kind="stable")), eff(np.argsort(-size, kind="stable")), eff(np.argsort(-per))))
names = ["short-distance copies", "short-distance loop steps", "static-dictionary words", "copy chunks > 8 bytes", "command extras beyond one peek",
         "distance extras beyond one peek", "phase-A decode without look-ahead", "phase-C decode without look-ahead", "commands"]
tot = max(int(cnt[8]), 1)
for i, nm in enumerate(names):
    p = min(1.0, cnt[i] / tot)
    print("  %-36s %6.2f %% of commands -> some lane of 32 needs it in %3.0f %% of the rounds" % (nm, 100.0 * cnt[i] / tot, 100 * (1 - (1 - p) ** 32)))
