#!/usr/bin/env python3
"""Table-space statistics of the lane kernel on the HOST build of its logic (no GPU): per metablock the tree counts,
root widths and which groups left the shared slot; per symbol kind the share of decodes that needed a second-level
entry, for several slot sizes E (u16 entries).
    python profiles/hostsim_tables.py [config] [n_unique] [E ...]"""
import ctypes, importlib, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
corpus = importlib.import_module("tools.corpus")
cfg = sys.argv[1] if len(sys.argv) > 1 else "headline"
n_unique = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Es = [int(x) for x in sys.argv[3:]] or [178, 146, 103, 73]
d = os.path.join(ROOT, "build_tmp")
os.makedirs(d, exist_ok=True)
src = os.path.join(d, "tables_stats.cpp")
open(src, "w").write(r'''#include <stdint.h>
extern "C" { uint64_t g_mb[64] = {0}; uint64_t g_la[32] = {0}; }
// g_mb: 0 metablocks, 1 nontrivial, 2..4 sum ntrees lit/cmd/dist, 5..7 sum rbits, 8..10 groups in arena, 11 sum cold entries used
#define BD_LANE_MB_STATS(c, L) do { g_mb[0]++; g_mb[1] += (L.e_tab != c.E); g_mb[2] += L.n_lit; g_mb[3] += L.nbt[1]; g_mb[4] += L.n_dist; \
  for (int g_ = 0; g_ < 3; g_++) { g_mb[5 + g_] += L.rbits[g_]; g_mb[8 + g_] += (L.root[g_] >= c.E); } g_mb[11] += L.cold_next - c.E; \
  g_mb[12] += (L.nbt[0] > 1); g_mb[13] += (L.nbt[1] > 1); g_mb[14] += (L.nbt[2] > 1); } while (0)
// look-ahead: SLOT 32 = next-A symbol (command or literal), 48 = distance
#define BD_LANE_LA_STATS(SLOT, IN, TWO) do { g_la[((SLOT) == 48u ? 0 : 4) + 0]++; g_la[((SLOT) == 48u ? 0 : 4) + 1] += (IN) ? 0 : 1; g_la[((SLOT) == 48u ? 0 : 4) + 2] += (TWO) ? 1 : 0; } while (0)
#define BD_LANE_FALLBACK_STATS(kind, is_lit) (g_la[8 + (kind) * 2 + ((is_lit) ? 1 : 0)]++)
#include "../tests/hostsim/hostsim_lane.cpp"
''')
so = os.path.join(d, "libtables_stats.so")
subprocess.check_call(["g++", "-O2", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", so, src, os.path.join(ROOT, "tables", "brotli_dictionary.c"),
                       '-DBROTLI_DICT_PATH="%s"' % os.path.join(ROOT, "tables", "brotli_dictionary.bin")])
L = ctypes.CDLL(so)
L.hostsim_lane_decode.restype = ctypes.c_int
L.hostsim_lane_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32,
                                  ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
mb = (ctypes.c_uint64 * 64).in_dll(L, "g_mb")
la = (ctypes.c_uint64 * 32).in_dll(L, "g_la")
comp, orig, _ = corpus.make_config(cfg, n_unique)
for E in Es:
    for i in range(64): mb[i] = 0
    for i in range(32): la[i] = 0
    bails = 0
    for c, o in zip(comp, orig):
        buf = ctypes.create_string_buffer(len(o) + 80)
        dd, u = ctypes.c_uint64(0), ctypes.c_uint64(0)
        r = L.hostsim_lane_decode(c, len(c), ctypes.addressof(buf) + (-ctypes.addressof(buf)) % 8, len(o), E, ctypes.byref(dd), ctypes.byref(u))
        bails += r != 1
    n = max(1, mb[0])
    print("%s E=%d: %d streams, %d metablocks (%.2f/stream), bails %d, context-modelled literals in %.0f %% of metablocks; multi-block-type lit/cmd/dist %.0f/%.0f/%.0f %%" % (
        cfg, E, len(comp), mb[0], mb[0] / len(comp), bails, 100 * mb[1] / n, 100 * mb[12] / n, 100 * mb[13] / n, 100 * mb[14] / n))
    print("   trees lit/cmd/dist %.2f / %.2f / %.2f; root bits %.2f / %.2f / %.2f; group in arena %.0f / %.0f / %.0f %%; arena entries used %.0f" % (
        mb[2] / n, mb[3] / n, mb[4] / n, mb[5] / n, mb[6] / n, mb[7] / n, 100 * mb[8] / n, 100 * mb[9] / n, 100 * mb[10] / n, mb[11] / n))
    for nm, b in (("distance look-ahead", 0), ("next-A look-ahead", 4)):
        t = max(1, la[b])
        print("   %-20s %9d: root outside slot %.1f %%, second level %.1f %%" % (nm, la[b], 100 * la[b + 1] / t, 100 * la[b + 2] / t))
    print("   synchronous decodes: phase A cmd %d lit %d, phase C %d" % (la[8], la[9], la[10]))
