"""Pins the CPU oracle (oracle/brotli_oracle.c) to the reference's own vectors: every testdata
fixture (tests/golden/manifest.json), the inline vectors of its unit tests, the 256 one-byte
streams, and -- differentially -- the system libbrotlidec 1.1.0 (the C decoder the crate ports)."""
import hashlib

import numpy as np
import pytest

import helpers

MAN = helpers.golden_manifest()
SMALL = [n for n, e in MAN.items() if "original_sha256" in e]


@pytest.mark.parametrize("name", SMALL)
def test_fixture_decodes_to_original(oracle, name):
    e = MAN[name]
    result, code, out = oracle.decode(helpers.golden_fixture(name), e["original_size"])
    assert (result, code) == (1, 1)
    assert len(out) == e["original_size"]
    assert hashlib.sha256(out).hexdigest() == e["original_sha256"]


def test_borked_fails(oracle):  # src/bin/integration_tests.rs:972
    result, code, _ = oracle.decode(helpers.golden_fixture("borked.compressed"), 1 << 20)
    assert result == 0 and code < 0


def test_large_window_rnd_chunk(oracle):  # src/bin/integration_tests.rs:465-507,998-1006
    e = MAN["rnd_chunk.br"]
    data = helpers.golden_fixture("rnd_chunk.br")
    result, code, out = oracle.decode(data, e["original_size"])
    assert (result, code, len(out)) == (1, 1, e["original_size"])
    pre, post = bytes.fromhex(e["prefix_hex"]), bytes.fromhex(e["postfix_hex"])
    rep_gap = 100000000  # the prefix repeats after this many zero bytes
    assert out[:len(pre)] == pre and out[-len(post):] == post
    assert out[len(pre) + rep_gap:2 * len(pre) + rep_gap] == pre
    nulls_in_input = 2 * pre.count(0) + post.count(0)
    assert out.count(0) - nulls_in_input == len(out) - 2 * len(pre) - len(post)
    # a non-large-window state rejects it (src/ffi/mod.rs:127)
    result, code, _ = oracle.decode(data, e["original_size"], large_window=False)
    assert (result, code) == (0, -13)


@pytest.mark.parametrize("v", helpers.inline_vectors()["vectors"], ids=lambda v: v["name"])
def test_inline_vector(oracle, v):
    data = bytes.fromhex(v["input_hex"])
    result, code, out = oracle.decode(data, 1 << 18)
    if v["expect"] == "failure":
        assert result == 0
        if "code" in v:
            assert code == v["code"]
    elif v["expect"] == "success_prefix":
        exp = bytes.fromhex(v["output_hex"])
        assert result == 1 and out[:len(exp)] == exp
    else:
        assert result == 1 and out == bytes.fromhex(v["output_hex"])


def test_one_byte_streams(oracle):  # src/bin/tests.rs:76-158
    ok = set(helpers.inline_vectors()["one_byte_ok"])
    for b in range(256):
        result, code, out = oracle.decode(bytes([b]), 64)
        assert (result == 1) == (b in ok), b
        assert out == b""


def test_output_too_small_and_truncation(oracle):
    data = helpers.golden_fixture("alice29.txt.compressed")
    size = MAN["alice29.txt.compressed"]["original_size"]
    result, code, out = oracle.decode(data, size - 1)
    assert (result, code, len(out)) == (3, 3, size - 1)  # NeedsMoreOutput, App. D-3
    result, code, out = oracle.decode(data[:-100], size)
    assert (result, code) == (2, 2) and 0 < len(out) < size  # truncated input: partial output, App. D-5


def test_differential_vs_system_decoder(oracle, corpus):
    """Generated q0..11 streams and their corruptions: oracle result must agree with libbrotlidec
    (success/failure, and bytes on success)."""
    rng = np.random.default_rng(1234)
    pools = list(corpus.mix_pools().values())
    checked = 0
    for q in range(0, 12):
        for size in (0, 1, 17, 700, 5000, 70000 if q < 10 else 20000):
            pool = pools[int(rng.integers(0, len(pools)))]
            orig = corpus.cut_windows(pool, 1, size, rng)[0] if size else b""
            comp = corpus.compress(orig, q, int(rng.integers(10, 25)))
            result, code, out = oracle.decode(comp, len(orig))
            assert (result, out) == (1, orig), (q, size)
            for m in helpers.mutations(comp, rng, 6):
                ok, ref = corpus.system_decompress(m, len(orig) + 64)
                result, code, out = oracle.decode(m, len(orig) + 64)
                assert (result == 1) == ok, (q, size, code)
                if ok:
                    assert out == ref
                checked += 1
    assert checked > 300
