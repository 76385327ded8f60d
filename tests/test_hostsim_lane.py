"""CPU check of the lane-per-stream kernel's logic (csrc/brotli_decode_lane.cuh compiled for the host):
the optimistic path must either decode a stream exactly as the oracle does (SUCCESS, same bytes, same
input consumption) or give it up -- never accept what the reference rejects, never write outside its
output region.  The GPU tests (-m gpu) repeat this through the C ABI with 32 streams per warp."""
import hashlib

import numpy as np
import pytest

import helpers

MAN = helpers.golden_manifest()
SMALL = [n for n, e in MAN.items() if "original_sha256" in e]
# fixtures the lane path is expected to decode itself (compressed metablocks only, regular window)
MUST_DECODE = ["alice29.txt.compressed", "asyoulik.txt.compressed", "lcet10.txt.compressed", "plrabn12.txt.compressed",
               "backward65536.compressed", "10x10y.compressed", "64x.compressed", "monkey.compressed", "ukkonooa.compressed",
               "mapsdatazrh.compressed", "quickfox_repeated.compressed", "zeros.compressed"]


@pytest.fixture(autouse=True, params=[0, 1], ids=["throughput", "latency"])
def lane_config(request, hostsim):
    """Both configurations of the command loop: what the large geometries run, and the latency configuration of the small ones
    (more literals per round; BD_LANE_*_LAT in csrc/brotli_decode_lane.cuh)."""
    hostsim.lib.hostsim_lane_set_latency_config(request.param)
    yield
    hostsim.lib.hostsim_lane_set_latency_config(0)


@pytest.mark.parametrize("name", SMALL)
def test_fixture(hostsim, name):
    e = MAN[name]
    data = helpers.golden_fixture(name)
    for entries in (430, 178, 146, 103, 73, 53, 34, 100000):
        for mis in (0, 1, 2, 3, 13, 16, 28, 31):
            code, out, used = hostsim.lane_decode(data, e["original_size"], entries, mis)
            if code == 1:
                assert hashlib.sha256(out).hexdigest() == e["original_sha256"], (name, entries, mis)
                assert used <= len(data)
            else:
                assert code == hostsim.LANE_BAIL
                assert not (name in MUST_DECODE and entries >= 146), name


def test_generated_configs_and_capacity(hostsim, oracle, corpus):
    rng = np.random.default_rng(11)
    decoded = {}
    for cfg, n, size in (("headline", 12, None), ("C3", 32, None), ("C5", 44, None), ("C4", 1, 1 << 19)):
        comp, orig, _ = corpus.make_config(cfg, n, size=size)
        ok = 0
        for c, o in zip(comp, orig):
            code, out, used = hostsim.lane_decode(c, len(o), int(rng.choice([178, 146, 103, 73, 34])), int(rng.integers(0, 32)))
            if code == 1:
                assert out == o and used == len(c)
                ok += 1
                assert hostsim.lane_decode(c, len(o) - 1)[0] == hostsim.LANE_BAIL  # too small a region is the exact kernel's case
                assert hostsim.lane_decode(c, len(o) + 7)[1] == o
        decoded[cfg] = ok
    assert decoded["headline"] == 12 and decoded["C3"] == 32 and decoded["C5"] >= 36, decoded


def test_mutations_never_accept_what_the_oracle_rejects(hostsim, oracle, corpus):
    rng = np.random.default_rng(12)
    comp, orig, _ = corpus.make_config("C5", 22)
    comp2, orig2, _ = corpus.make_config("C3", 12)
    accepted = 0
    for c, o in list(zip(comp, orig)) + list(zip(comp2, orig2)):
        for m in helpers.mutations(c, rng, 25):
            if not m:
                continue
            cap = len(o) + int(rng.integers(0, 64))
            code, out, _ = hostsim.lane_decode(m, cap, 178, int(rng.integers(0, 32)))
            if code == 1:
                _, ocode, oout = oracle.decode(m, cap)
                assert ocode == 1 and out == oout
                accepted += 1
    assert accepted > 0


def test_inline_vectors_and_one_byte_streams(hostsim, oracle):
    vec = helpers.inline_vectors()
    streams = [bytes.fromhex(v["input_hex"]) for v in vec["vectors"] if v.get("input_hex")] + [bytes([b]) for b in range(256)]
    for data in streams:
        for cap in (0, 16, 4096):
            code, out, _ = hostsim.lane_decode(data, cap)
            if code == 1:
                _, ocode, oout = oracle.decode(data, cap)
                assert ocode == 1 and out == oout, data.hex()
