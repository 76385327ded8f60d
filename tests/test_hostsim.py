"""CPU check of the CUDA decoder's control flow: csrc/brotli_decode_core.cuh compiled for the host
with warp width 1 (tests/hostsim) against the oracle -- same result code, same decoded_size, same
bytes -- on the reference's fixtures, generated q0..11 corpora, truncations and corruptions.
The GPU tests (-m gpu) repeat this through the C ABI with real warps."""
import hashlib

import numpy as np
import pytest

import helpers

MAN = helpers.golden_manifest()
SMALL = [n for n, e in MAN.items() if "original_sha256" in e]


def same(oracle, hostsim, data, cap, large_window=True, roomy=True):
    """roomy=False: the buffer may be too small for what the stream describes.  The reference decodes
    into its ring buffer and meets a later corruption before it notices the full output buffer; the
    GPU path has no ring buffer and stops at the capacity, so for a stream that is BOTH corrupt and
    given too little room it may report NEEDS_MORE_OUTPUT where the reference names the corruption
    (both are BROTLI_DECODER_RESULT_ERROR through BrotliDecoderDecompress; DESIGN.md, deviations)."""
    result, code, out = oracle.decode(data, cap, large_window=large_window)
    hcode, hout = hostsim.decode(data, cap, large_window=large_window)
    if not roomy and hcode == 3 and code != 1 and code != 3:
        return code
    assert hcode == code, (hcode, code, len(data), cap)
    assert hout == out, (code, len(hout), len(out))
    return code


@pytest.mark.parametrize("name", SMALL)
def test_fixture(oracle, hostsim, name):
    e = MAN[name]
    code, out = hostsim.decode(helpers.golden_fixture(name), e["original_size"])
    assert code == 1 and hashlib.sha256(out).hexdigest() == e["original_sha256"]


def test_fixture_capacity_matrix(oracle, hostsim):
    """Too-small / oversized output buffers around the exact size, and truncated inputs."""
    for name in ("10x10y.compressed", "quickfox_repeated.compressed", "alice29.txt.compressed", "backward65536.compressed",
                 "zeros.compressed", "ukkonooa.compressed", "random_org_10k.bin.compressed", "metablock_reset.compressed",
                 "compressed_repeated.compressed", "empty.compressed.17", "x.compressed.03"):
        data = helpers.golden_fixture(name)
        size = MAN[name]["original_size"]
        for cap in sorted({0, 1, size // 3, size - 1, size, size + 1, size + 100} - {-1}):
            same(oracle, hostsim, data, cap)
        for cut in sorted({1, 2, 3, len(data) // 2, len(data) - 2, len(data) - 1}):
            if 0 < cut < len(data):
                same(oracle, hostsim, data[:cut], size + 16)


def test_borked_and_large_window_flag(oracle, hostsim):
    same(oracle, hostsim, helpers.golden_fixture("borked.compressed"), 1 << 20)
    data = helpers.golden_fixture("rnd_chunk.br")
    assert same(oracle, hostsim, data, 4096, large_window=False) == -13


def test_large_window_rnd_chunk(oracle, hostsim):
    e = MAN["rnd_chunk.br"]
    code, out = hostsim.decode(helpers.golden_fixture("rnd_chunk.br"), e["original_size"])
    result, ocode, oout = oracle.decode(helpers.golden_fixture("rnd_chunk.br"), e["original_size"])
    assert code == 1 and ocode == 1 and out == oout


@pytest.mark.parametrize("v", helpers.inline_vectors()["vectors"], ids=lambda v: v["name"])
def test_inline_vector(oracle, hostsim, v):
    same(oracle, hostsim, bytes.fromhex(v["input_hex"]), 1 << 18)


def test_one_byte_streams(oracle, hostsim):
    for b in range(256):
        same(oracle, hostsim, bytes([b]), 64)
    same(oracle, hostsim, b"", 64)


@pytest.mark.parametrize("q", range(0, 12))
def test_generated_and_mutated(oracle, hostsim, corpus, q):
    rng = np.random.default_rng(77 + q)
    pools = list(corpus.mix_pools().values())
    n_fail = 0
    for size in (0, 1, 33, 900, 4096, 65536 if q < 10 else 16384, 200000 if q in (1, 5, 9) else 3000):
        pool = pools[int(rng.integers(0, len(pools)))]
        orig = corpus.cut_windows(pool, 1, size, rng)[0] if size else b""
        comp = corpus.compress(orig, q, int(rng.integers(10, 25)))
        assert same(oracle, hostsim, comp, len(orig)) == 1
        same(oracle, hostsim, comp, max(len(orig) - 1, 0))
        same(oracle, hostsim, comp, len(orig) // 2)
        for m in helpers.mutations(comp, rng, 12):
            if same(oracle, hostsim, m, len(orig) + 32) != 1:
                n_fail += 1
            same(oracle, hostsim, m, max(len(orig) - 7, 0), roomy=False)
    assert n_fail > 10


def test_differential_fuzz_both_cores(oracle, hostsim, corpus):
    """Host builds of BOTH kernels' logic against the oracle on a few thousand seeded mutations of streams of every
    family and quality: the exact core must reproduce code and decoded bytes (up to the documented corrupt-and-too-small
    deviation, see `same`), the lane core must either decode exactly what the oracle decodes or give the stream up."""
    n = lane_ok = tolerated = 0
    for seed in range(2):
        rng = np.random.default_rng(7000 + seed)
        for cfg, cnt, size in (("C5", 22, int(rng.integers(3000, 30000))), ("C3", 16, 4096), ("headline", 3, 65536)):
            comp, orig, _ = corpus.make_config(cfg, cnt, size=size, seed=900 + seed)
            for c, o in zip(comp, orig):
                for m in [c] + helpers.mutations(c, rng, 20):
                    cap = len(o) + int(rng.integers(0, 40)) if rng.random() < 0.8 else int(rng.integers(0, len(o) + 1))
                    _, ocode, oout = oracle.decode(m, cap)
                    hcode, hout = hostsim.decode(m, cap)
                    n += 1
                    if (hcode, hout) != (ocode, oout):
                        assert hcode == 3 and ocode not in (1, 3), (cfg, len(m), cap, hcode, ocode)
                        tolerated += 1
                    lcode, lout, _ = hostsim.lane_decode(m, cap, int(rng.choice([146, 178, 230])), int(rng.integers(0, 4)))
                    if lcode == 1:
                        assert ocode == 1 and lout == oout, (cfg, len(m), cap)
                        lane_ok += 1
                    else:
                        assert lcode == hostsim.LANE_BAIL
    assert n > 1500 and lane_ok > 200 and tolerated < n // 5, (n, lane_ok, tolerated)
