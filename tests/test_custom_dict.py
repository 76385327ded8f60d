"""Custom LZ77 dictionaries (BrotliState::new_with_custom_dictionary, src/state.rs:400-411): the reference's own
known-answer tests (src/test.rs:438-508, tests/golden/dict_vectors.json) pin the oracle; the exact kernel's logic
(host build) and, with -m gpu, the kernel itself through the C ABI are compared with the oracle on those vectors,
on streams compressed against a dictionary (libbrotlienc raw shared dictionary; same addressing while
dictionary + data fit the window), on truncations / corruptions / small output regions, and on the reference's
corner cases: dictionaries longer than the window (only the last (1 << WBITS) - 16 bytes are reachable) and the
literal-context seed at positions 0 and 1."""
import json
import os

import numpy as np
import pytest

import helpers

VEC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "dict_vectors.json")))["vectors"]


def dict_cases(corpus, seed, count):
    """(compressed, dictionary, original) triples; dictionary and data overlap in the pool so that references into the
    dictionary are plentiful."""
    rng = np.random.default_rng(seed)
    pool = corpus.text_pool()
    out = []
    for i in range(count):
        q = int(rng.choice([2, 5, 9, 11]))
        a = int(rng.integers(0, len(pool) - 200000))
        dlen = int(rng.integers(16, 60000))
        n = int(rng.integers(1, 40000))
        d = pool[a:a + dlen]
        data = pool[a + dlen // 2:a + dlen // 2 + n]
        lgwin = int(rng.choice([17, 18, 22]))
        out.append((corpus.compress_with_dictionary(data, d, q, lgwin), d, data))
    return out


@pytest.mark.parametrize("v", VEC, ids=[v["name"] for v in VEC])
def test_reference_known_answers(oracle, hostsim, v):
    data, d, want = (bytes.fromhex(v[k]) for k in ("input_hex", "dict_hex", "output_hex"))
    res, code, out = oracle.decode(data, 1024, True, d)
    assert (res, code, out) == (1, 1, want)
    assert hostsim.decode(data, 1024, True, d) == (1, want)
    # without the dictionary the stream decodes to something else (or fails): the dictionary is really used
    assert oracle.decode(data, 1024, True)[2] != want


def test_exact_core_matches_oracle_on_generated_streams(oracle, hostsim, corpus):
    rng = np.random.default_rng(21)
    used_dict = 0
    for comp, d, data in dict_cases(corpus, 20, 24):
        res, code, out = oracle.decode(comp, len(data), True, d)
        assert (code, out) == (1, data)
        assert hostsim.decode(comp, len(data), True, d) == (1, data)
        used_dict += oracle.decode(comp, len(data), True)[2] != data
        for cap in (0, len(data) // 2, len(data) - 1, len(data) + 9):
            _, oc, oo = oracle.decode(comp, cap, True, d)
            assert hostsim.decode(comp, cap, True, d) == (oc, oo), cap
        for m in helpers.mutations(comp, rng, 12):
            _, oc, oo = oracle.decode(m, len(data) + 16, True, d)
            assert hostsim.decode(m, len(data) + 16, True, d) == (oc, oo), m.hex()[:80]
    assert used_dict >= 12


def test_dictionary_longer_than_the_window(oracle, hostsim, corpus):
    # WBITS 16 (lgwin 16): only the last 65520 dictionary bytes are reachable (src/decode.rs:1831-1838), and the
    # distance limit is max_backward from the first byte on (:2954-2955 uses the unclipped length)
    pool = corpus.text_pool()
    tail = pool[300000:300000 + 20000]
    data = pool[310000:310000 + 9000]
    comp = corpus.compress_with_dictionary(data, tail, 9, 16)
    for pad in (0, 65520 - len(tail), 65536 - len(tail), 200000):
        d = pool[:pad] + tail
        _, oc, oo = oracle.decode(comp, len(data), True, d)
        assert hostsim.decode(comp, len(data), True, d) == (oc, oo), pad
        if pad == 0:
            assert (oc, oo) == (1, data)


def test_context_seed_with_dictionary(oracle, hostsim, corpus):
    # streams whose literal contexts matter from the first byte (q10/q11 model literal contexts)
    pool = corpus.text_pool()
    for a in (1000, 50000, 400000):
        d = pool[a:a + 3000]
        data = pool[a + 7000:a + 7000 + 6000]
        for q in (10, 11):
            comp = corpus.compress_with_dictionary(data, d, q)
            _, oc, oo = oracle.decode(comp, len(data), True, d)
            assert (oc, oo) == (1, data)
            assert hostsim.decode(comp, len(data), True, d) == (1, data)


@pytest.mark.gpu
def test_gpu_custom_dictionary(gpu_lib, pkg, oracle, corpus):
    for v in VEC:
        data, d, want = (bytes.fromhex(v[k]) for k in ("input_hex", "dict_hex", "output_hex"))
        info, out = pkg.brotli_decode_custom_dict(data, 1024, d)
        assert (info.result, info.code, out) == (1, 1, want)
    rng = np.random.default_rng(23)
    for comp, d, data in dict_cases(corpus, 22, 12):
        info, out = pkg.brotli_decode_custom_dict(comp, len(data), d)
        assert (info.code, out) == (1, data)
        for cap in (0, len(data) - 1):
            _, oc, oo = oracle.decode(comp, cap, True, d)
            info, out = pkg.brotli_decode_custom_dict(comp, cap, d)
            assert (info.code, out) == (oc, oo)
        for m in helpers.mutations(comp, rng, 6):
            _, oc, oo = oracle.decode(m, len(data) + 16, True, d)
            info, out = pkg.brotli_decode_custom_dict(m, len(data) + 16, d)
            assert (info.code, out) == (oc, oo)


@pytest.mark.gpu
def test_gpu_custom_dictionary_batch(gpu_lib, pkg, oracle, corpus):
    # one dictionary shared by a batch (shared-dictionary serving: many small responses against one base document)
    pool = corpus.text_pool()
    d = pool[500000:560000]
    rng = np.random.default_rng(24)
    origs = [pool[500000 + int(rng.integers(0, 50000)):][:int(rng.integers(1, 9000))] for _ in range(300)]
    comp = [corpus.compress_with_dictionary(o, d, int(rng.choice([5, 9]))) for o in origs]
    comp[7] = comp[7][:len(comp[7]) // 2]  # one truncated stream
    caps = [len(o) for o in origs]
    caps[11] -= 1                            # one region too small
    in_bytes, in_off = corpus.pack(comp)
    out_off = np.zeros(len(comp) + 1, dtype=np.uint64)
    out_off[1:] = np.cumsum(caps)
    out_bytes = np.zeros(int(out_off[-1]) + 1, dtype=np.uint8)
    out_len = np.zeros(len(comp), dtype=np.uint64)
    codes = np.zeros(len(comp), dtype=np.int32)
    before = pkg.kernel_launch_count()
    pkg.decompress_batch_packed_custom_dict(in_bytes, in_off, out_bytes, out_off, out_len, codes, d)
    assert pkg.kernel_launch_count() > before
    for i, (c, cap) in enumerate(zip(comp, caps)):
        _, oc, oo = oracle.decode(c, cap, True, d)
        got = out_bytes[int(out_off[i]):int(out_off[i]) + int(out_len[i])].tobytes()
        assert (int(codes[i]), got) == (oc, oo), i
    assert sum(int(c) == 1 for c in codes) == len(comp) - 2
    # the same call without a dictionary must not be affected by the previous one
    plain = [corpus.compress(o, 5) for o in origs[:40]]
    res = pkg.decompress_batch(plain, [len(o) for o in origs[:40]])
    assert all(r[1] == 1 and r[2] == o for r, o in zip(res, origs[:40]))


def test_lane_core_with_dictionary(oracle, hostsim, corpus):
    """The lane kernel's dictionary instance (host build): it decodes a stream exactly as the oracle does or gives it up
    (copies that start in the dictionary and run on into the output, anything malformed) -- and it must decode most
    well-formed streams itself."""
    rng = np.random.default_rng(31)
    decoded = total = 0
    for comp, d, data in dict_cases(corpus, 30, 40):
        total += 1
        code, out, used = hostsim.lane_decode(comp, len(data), int(rng.choice([146, 178])), int(rng.integers(0, 4)), custom_dict=d)
        if code == 1:
            assert out == data and used == len(comp)
            decoded += 1
        else:
            assert code == hostsim.LANE_BAIL
        for m in helpers.mutations(comp, rng, 10):
            cap = len(data) + int(rng.integers(0, 32))
            code, out, _ = hostsim.lane_decode(m, cap, 178, int(rng.integers(0, 4)), custom_dict=d)
            if code == 1:
                _, ocode, oout = oracle.decode(m, cap, True, d)
                assert ocode == 1 and out == oout
    for v in VEC:
        data, d, want = (bytes.fromhex(v[k]) for k in ("input_hex", "dict_hex", "output_hex"))
        code, out, _ = hostsim.lane_decode(data, 1024, 178, 0, custom_dict=d)
        assert code == hostsim.LANE_BAIL or out == want
    # dictionaries longer than the window and the zero context seed (q10/q11 model literal contexts from byte 0)
    pool = corpus.text_pool()
    tail = pool[300000:320000]
    payload = pool[310000:319000]
    comp = corpus.compress_with_dictionary(payload, tail, 9, 16)
    for pad in (0, 65520 - len(tail), 200000):
        dd = pool[:pad] + tail
        _, ocode, oout = oracle.decode(comp, len(payload), True, dd)
        code, out, _ = hostsim.lane_decode(comp, len(payload), 178, 0, custom_dict=dd)
        assert code == hostsim.LANE_BAIL or (ocode == 1 and out == oout), pad
    for q in (10, 11):
        d = pool[1000:4000]; payload = pool[8000:14000]
        comp = corpus.compress_with_dictionary(payload, d, q)
        code, out, _ = hostsim.lane_decode(comp, len(payload), 178, 0, custom_dict=d)
        assert code == hostsim.LANE_BAIL or out == payload
    assert decoded >= total * 3 // 4, (decoded, total)


@pytest.mark.gpu
def test_gpu_streaming_with_custom_dictionary(gpu_lib, pkg, oracle, corpus):
    """Decompressor::new_with_custom_dict / BrotliDecompressCustomDict semantics through the streaming state."""
    import io
    pool = corpus.text_pool()
    d = pool[600000:650000]
    data = pool[610000:610000 + 500000]
    comp = corpus.compress_with_dictionary(data, d, 2)
    st = pkg.DecoderState(custom_dict=d)
    got, pos, r = [], 0, 2
    while r not in (0, 1):
        r, used, out = st.decompress_stream(comp[pos:pos + 30000], 1 << 18)
        pos += used
        got.append(out)
    assert r == 1 and b"".join(got) == data
    st.close()
    rd = pkg.Decompressor(io.BytesIO(comp), 4096, custom_dict=d)
    out = b""
    while True:
        b = rd.read(100000)
        if not b:
            break
        out += b
    assert out == data
    for v in VEC:  # the reference's own dictionary tests run through a streaming state (src/test.rs:120-150)
        st = pkg.DecoderState(custom_dict=bytes.fromhex(v["dict_hex"]))
        r, used, out = st.decompress_stream(bytes.fromhex(v["input_hex"]), 1024)
        assert r == 1 and out == bytes.fromhex(v["output_hex"])
        assert not pkg.lib().BrotliB200DecoderSetCustomDictionary(st._s, b"x", 1)  # only before the first byte
        st.close()
