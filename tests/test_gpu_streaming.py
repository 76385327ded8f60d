"""GPU streaming sessions through the C ABI (BrotliDecoderDecompressStream / BrotliB200DecoderDecompressStreamBatch)
in lock-step with the streaming oracle (the reference's BrotliDecompressStream, src/decode.rs:2779-2896): every call
must give the same result, bytes consumed, bytes produced and total_out.  Config C1 (alice29) runs the reference's
full buffer matrix (src/bin/integration_tests.rs:439-463)."""
import hashlib

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

MATRIX = [(65536, 65536), (1, 65536), (65536, 1), (1, 1), (3, 3), (1024, 1024)]


def lockstep(oracle, pkg, comp, in_chunk, out_chunk, large_window=False, custom_dict=None):
    so = oracle.stream(large_window=large_window, custom_dict=custom_dict)
    st = pkg.DecoderState(large_window=large_window, custom_dict=custom_dict)
    ro, oo, to = helpers.drive_stream(so.call, comp, in_chunk, out_chunk)
    rg, og, tg = helpers.drive_stream(st.call, comp, in_chunk, out_chunk)
    if to != tg:
        k = next(i for i, (a, b) in enumerate(zip(to + [None], tg + [None])) if a != b)
        raise AssertionError("call %d of %d/%d differs: oracle %r, gpu %r (in %d out %d)" % (k, len(to), len(tg), to[k] if k < len(to) else None,
                                                                                          tg[k] if k < len(tg) else None, in_chunk, out_chunk))
    assert (ro, oo) == (rg, og)
    if ro == 0:
        assert so.error_code() == st.error_code()
    if ro == 1:
        assert st.is_finished()
    so.close(); st.close()
    return ro, oo, to


@pytest.mark.parametrize("pair", MATRIX, ids=lambda p: "%dx%d" % p)
def test_c1_alice29_buffer_matrix(gpu_lib, pkg, oracle, pair):
    comp = helpers.golden_fixture("alice29.txt.compressed")
    man = helpers.golden_manifest()["alice29.txt.compressed"]
    r, out, trace = lockstep(oracle, pkg, comp, pair[0], pair[1])
    assert r == 1 and len(out) == 152089 and hashlib.sha256(out).hexdigest() == man["original_sha256"]
    assert sum(t[1] for t in trace) == len(comp)


def test_fixtures_random_chunkings(gpu_lib, pkg, oracle):
    rng = np.random.default_rng(21)
    man = helpers.golden_manifest()
    names = sorted(n for n, m in man.items() if "original_size" in m and m["original_size"] <= 1200000)
    for name in names:
        comp = helpers.golden_fixture(name)
        for _ in range(2):
            r, out, _ = lockstep(oracle, pkg, comp, int(rng.integers(64, 40000)), int(rng.integers(64, 200000)), large_window=True)
            assert r == 1 and hashlib.sha256(out).hexdigest() == man[name]["original_sha256"], name


def test_small_windows_and_errors(gpu_lib, pkg, oracle, corpus):
    rng = np.random.default_rng(22)
    pool = corpus.text_pool()
    data = bytes(pool[50000:50000 + 120000])
    for lgwin in (10, 12, 16):
        comp = corpus.compress(data, 5, lgwin=lgwin)
        for pair in ((4096, 100), (100000, 1 << lgwin), (100000, (1 << lgwin) + 1), (513, 517)):
            r, out, _ = lockstep(oracle, pkg, comp, pair[0], pair[1])
            assert r == 1 and out == data
        for m in helpers.mutations(comp, rng, 8):
            lockstep(oracle, pkg, m, int(rng.integers(100, 5000)), int(rng.integers(100, 40000)))


def test_custom_dictionary_session(gpu_lib, pkg, oracle, corpus):
    pool = corpus.text_pool()
    d = bytes(pool[200000:230000])
    data = bytes(pool[210000:225000]) + bytes(pool[400000:430000])
    comp = corpus.compress_with_dictionary(data, d, 5, lgwin=18)
    for pair in ((300, 1000), (65536, 65536)):
        r, out, _ = lockstep(oracle, pkg, comp, pair[0], pair[1], large_window=True, custom_dict=d)
        assert r == 1 and out == data


def test_multiplexed_sessions_one_launch(gpu_lib, pkg, oracle, corpus):
    """256 open sessions fed in lock-step through BrotliB200DecoderDecompressStreamBatch: one decode launch per round;
    each session's trace equals its own oracle stream."""
    rng = np.random.default_rng(23)
    pool = corpus.text_pool()
    n = 256
    datas = [bytes(pool[int(a):int(a) + int(rng.integers(3000, 120000))]) for a in rng.integers(0, len(pool) - 130000, size=n)]
    comps = [corpus.compress(d, int(rng.integers(1, 10)), lgwin=int(rng.integers(12, 22))) for d in datas]
    oracles = [oracle.stream() for _ in range(n)]
    states = [pkg.DecoderState() for _ in range(n)]
    pos = [0] * n; pend = [b""] * n; res = [2] * n; outs = [bytearray() for _ in range(n)]
    rounds = 0
    launches0 = pkg.kernel_launch_count()
    while any(r in (2, 3) for r in res):
        idx = [i for i in range(n) if res[i] in (2, 3)]
        for i in idx:
            if res[i] == 2:
                pend[i] = comps[i][pos[i]:pos[i] + 8192]; pos[i] += len(pend[i])
        caps = [int(rng.integers(1000, 60000)) for _ in idx]
        got = pkg.decompress_stream_batch([states[i] for i in idx], [pend[i] for i in idx], caps)
        for i, cap, g in zip(idx, caps, got):
            e = oracles[i].call(pend[i], cap)
            assert g == e, (i, rounds, g[:2], e[:2], g[3], e[3])
            res[i] = g[0]; pend[i] = pend[i][g[1]:]; outs[i] += g[2]
        rounds += 1
    assert all(r == 1 for r in res) and all(bytes(o) == d for o, d in zip(outs, datas))
    launches = pkg.kernel_launch_count() - launches0
    # per round: scatter + decode (+ empty-bail bookkeeping) + gather + window moves, a repeat for windows that had to grow --
    # a constant number of kernels per round, not one per session (256 sessions x 11 rounds would be ~3000 decode launches)
    assert launches <= 12 * rounds + 64, (launches, rounds)
    for s in states:
        s.close()


def test_reader_accepts_large_window(gpu_lib, pkg):
    """Decompressor<R> builds its state with new_with_custom_dictionary: large-window streams decode (src/reader.rs:226,
    src/state.rs:400-411; fixture of src/bin/integration_tests.rs:998-1006)."""
    import io
    comp = helpers.golden_fixture("rnd_chunk.br")
    out = pkg.Decompressor(io.BytesIO(comp), 65536).read()
    info, ref = pkg.brotli_decode(comp, len(out) + 64)  # the one-shot entry accepts large windows (src/state.rs:394)
    assert info.code == 1 and out == ref and len(out) > 1 << 20
    st = pkg.DecoderState()  # BrotliDecoderCreateInstance: large_window false (src/ffi/mod.rs:127)
    r, used, _ = st.decompress_stream(comp[:4096], 65536)
    assert r == 0 and st.error_code() == -13
    st.close()
