"""Call-by-call parity of the product's streaming sessions (csrc/brotli_b200_session.h over the host build of the exact
kernel) with the streaming oracle (the reference's BrotliDecompressStream, src/decode.rs:2779-2896): for every call the
same result, bytes consumed, bytes produced and total_out -- over the reference's buffer matrix
(src/bin/integration_tests.rs:439-463, C1 = alice29), the testdata fixtures, random chunkings, truncated and corrupted
streams, custom dictionaries, large windows, and many sessions multiplexed into one launch.  The GPU suite runs the same
comparisons through BrotliDecoderDecompressStream (tests/test_gpu_streaming.py)."""
import hashlib

import numpy as np
import pytest

import helpers

MATRIX = [(65536, 65536), (1, 65536), (65536, 1), (1, 1), (3, 3), (1024, 1024)]


def lockstep(oracle, hostsim, comp, in_chunk, out_chunk, large_window=False, custom_dict=None, expect=None):
    so = oracle.stream(large_window=large_window, custom_dict=custom_dict)
    sh = helpers.HostSimStream(hostsim, large_window=large_window, custom_dict=custom_dict)
    ro, oo, to = helpers.drive_stream(so.call, comp, in_chunk, out_chunk)
    rh, oh, th = helpers.drive_stream(sh.call, comp, in_chunk, out_chunk)
    if to != th:
        k = next(i for i, (a, b) in enumerate(zip(to + [None], th + [None])) if a != b)
        raise AssertionError("call %d of %d/%d differs: oracle %r, session %r (in %d out %d)" % (k, len(to), len(th), to[k] if k < len(to) else None,
                                                                                              th[k] if k < len(th) else None, in_chunk, out_chunk))
    assert (ro, oo) == (rh, oh)
    if ro == 0:
        assert so.error_code() == sh.error_code()
    if expect is not None:
        assert ro == 1 and oo == expect
    so.close(); sh.close()
    return ro, oo, to


@pytest.mark.parametrize("pair", MATRIX, ids=lambda p: "%dx%d" % p)
def test_c1_alice29_buffer_matrix(oracle, hostsim, pair):
    comp = helpers.golden_fixture("alice29.txt.compressed")
    man = helpers.golden_manifest()["alice29.txt.compressed"]
    r, out, trace = lockstep(oracle, hostsim, comp, pair[0], pair[1], large_window=True)
    assert r == 1 and hashlib.sha256(out).hexdigest() == man["original_sha256"]


def test_fixture_matrix(oracle, hostsim):
    man = helpers.golden_manifest()
    names = sorted(n for n, m in man.items() if "original_size" in m and m["original_size"] <= 400000)
    for name in names:
        comp = helpers.golden_fixture(name)
        for pair in MATRIX:
            if 1 in pair and man[name]["original_size"] > 20000:
                continue
            r, out, _ = lockstep(oracle, hostsim, comp, pair[0], pair[1], large_window=True)
            assert r == 1 and hashlib.sha256(out).hexdigest() == man[name]["original_sha256"], (name, pair)


def test_random_chunkings_all_qualities(oracle, hostsim, corpus):
    rng = np.random.default_rng(11)
    pool = corpus.text_pool()
    mix = b"".join(v for k, v in sorted(corpus.mix_pools().items()) if k != "text")
    for q in (0, 1, 2, 4, 5, 9, 11):
        for src in (pool, mix):
            a = int(rng.integers(0, len(src) - 300000))
            data = bytes(src[a:a + int(rng.integers(20000, 250000))])
            comp = corpus.compress(data, q, lgwin=int(rng.integers(16, 23)))
            for _ in range(2):
                lockstep(oracle, hostsim, comp, int(rng.integers(1, 9000)), int(rng.integers(1, 70000)), expect=data)


def test_small_windows_wrap_many_times(oracle, hostsim, corpus):
    """lgwin 10..14: the ring buffer wraps every 1..16 KiB, so every flush rule of WriteRingBuffer is exercised; output
    buffers smaller and larger than the ring."""
    pool = corpus.text_pool()
    rng = np.random.default_rng(12)
    for lgwin in (10, 11, 12, 14):
        data = bytes(pool[50000:50000 + 90000])
        comp = corpus.compress(data, 5, lgwin=lgwin)
        for pair in ((4096, 100), (77, 3000), (100000, 1 << lgwin), (100000, (1 << lgwin) - 1), (100000, (1 << lgwin) + 1), (513, 517)):
            lockstep(oracle, hostsim, comp, pair[0], pair[1], expect=data)


def test_truncated_and_corrupt_streams(oracle, hostsim, corpus):
    rng = np.random.default_rng(13)
    pool = corpus.text_pool()
    base = [helpers.golden_fixture(n) for n in ("alice29.txt.compressed", "random_org_10k.bin.compressed", "compressed_repeated.compressed")]
    base.append(corpus.compress(bytes(pool[1000:120000]), 5, lgwin=12))
    base.append(corpus.compress(bytes(pool[1000:60000]), 1, lgwin=16))
    kinds = set()
    for comp in base:
        for m in helpers.mutations(comp, rng, 14):
            r, out, trace = lockstep(oracle, hostsim, m, int(rng.integers(1, 5000)), int(rng.integers(1, 40000)), large_window=True)
            kinds.add(r)
    assert 0 in kinds and 2 in kinds


def test_custom_dictionary_and_large_window(oracle, hostsim, corpus):
    pool = corpus.text_pool()
    d = bytes(pool[200000:230000])
    data = bytes(pool[210000:225000]) + bytes(pool[400000:430000])
    comp = corpus.compress_with_dictionary(data, d, 5, lgwin=18) if hasattr(corpus, "compress_with_dictionary") else None
    if comp is not None:
        for pair in ((300, 1000), (65536, 65536), (7, 7)):
            lockstep(oracle, hostsim, comp, pair[0], pair[1], large_window=True, custom_dict=d, expect=data)
    man = helpers.golden_manifest()
    if "rnd_chunk.br" in man:  # large-window fixture (src/bin/integration_tests.rs:998-1006): refused unless the state allows it
        comp = helpers.golden_fixture("rnd_chunk.br")
        so = oracle.stream(large_window=False); sh = helpers.HostSimStream(hostsim, large_window=False)
        a = so.call(comp[:4096], 65536); b = sh.call(comp[:4096], 65536)
        assert a == b and a[0] == 0 and so.error_code() == sh.error_code() == -13


def test_multiplexed_sessions_one_launch(oracle, hostsim, corpus):
    """Many open sessions fed in lock-step: one (simulated) launch per round decodes all of them; each session's
    call-by-call trace equals its own oracle stream."""
    rng = np.random.default_rng(14)
    pool = corpus.text_pool()
    n = 24
    datas = [bytes(pool[int(a):int(a) + int(rng.integers(3000, 90000))]) for a in rng.integers(0, len(pool) - 100000, size=n)]
    comps = [corpus.compress(d, int(rng.integers(1, 10)), lgwin=int(rng.integers(12, 22))) for d in datas]
    oracles = [oracle.stream() for _ in range(n)]
    sims = [helpers.HostSimStream(hostsim) for _ in range(n)]
    pos = [0] * n; pend = [b""] * n; res = [2] * n; outs = [bytearray() for _ in range(n)]
    before = helpers.hostsim_stream_stats(hostsim)["launches"]
    rounds = 0
    while any(r in (2, 3) for r in res):
        idx = [i for i in range(n) if res[i] in (2, 3)]
        for i in idx:
            if res[i] == 2:
                pend[i] = comps[i][pos[i]:pos[i] + 2048]; pos[i] += len(pend[i])
        caps = [int(rng.integers(1, 6000)) for _ in idx]
        got = helpers.stream_calls(hostsim.lib, [sims[i] for i in idx], [pend[i] for i in idx], caps)
        for i, cap, g in zip(idx, caps, got):
            e = oracles[i].call(pend[i], cap)
            assert g == e, (i, rounds, g[:2], e[:2])
            res[i] = g[0]; pend[i] = pend[i][g[1]:]; outs[i] += g[2]
        rounds += 1
    launches = helpers.hostsim_stream_stats(hostsim)["launches"] - before
    assert all(r == 1 for r in res) and all(bytes(o) == d for o, d in zip(outs, datas))
    assert launches <= rounds + 8  # one launch per round (plus a few window-growth repeats), not one per session
    for s in sims:
        s.close()


def test_session_memory_is_bounded_by_the_window(oracle, hostsim, corpus):
    """A long stream through a small window: the session's device buffers stay O(window), not O(stream)."""
    pool = corpus.text_pool()
    data = bytes(pool[:1100000])
    comp = corpus.compress(data, 3, lgwin=14)
    base = helpers.hostsim_stream_stats(hostsim)["live_bytes"]
    sh = helpers.HostSimStream(hostsim)
    so = oracle.stream()
    out = bytearray(); pos = 0; peak = 0
    while pos < len(comp):
        piece = comp[pos:pos + 3000]; pos += len(piece)
        while True:
            g = sh.call(piece, 8192); e = so.call(piece, 8192)
            assert g == e
            out += g[2]; piece = piece[g[1]:]
            peak = max(peak, helpers.hostsim_stream_stats(hostsim)["live_bytes"] - base)
            if g[0] != 3:
                break
    assert bytes(out) == data
    arena = 1500000  # the session's table arena
    assert peak < arena + 600000, peak  # windows: 2 x (2 x 16 KiB + growth) output, a few KB of input -- not 1.1 MB of stream
    sh.close()
