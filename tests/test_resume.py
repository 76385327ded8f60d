"""Resumable streaming state of the exact kernel: a session decodes growing prefixes of a
stream into one persistent output buffer and continues behind the last metablock boundary an earlier call
reached (ResumeState, csrc/brotli_decode_core.cuh).  Every call must report exactly what a from-scratch decode of
the same prefix reports (the oracle: code and decoded bytes), whatever the chunking."""
import numpy as np
import pytest

import helpers


def multi_metablock_streams(corpus):
    pool = corpus.text_pool()
    out = []
    # low qualities emit a metablock every few hundred KB; the concatenation of two streams' worth of data at q2/q3
    for q, n, a in ((2, 900000, 1000), (3, 700000, 200000), (1, 400000, 50000), (5, 1100000, 0)):
        data = pool[a:a + n]
        out.append((corpus.compress(data, q), data))
    # uncompressed and metadata metablocks between compressed ones (reference fixtures)
    man = helpers.golden_manifest()
    for name in ("random_org_10k.bin.compressed", "compressed_repeated.compressed", "mapsdatazrh.compressed"):
        if name in man and "original_size" in man[name]:
            out.append((helpers.golden_fixture(name), None))
    return out


def test_resume_matches_from_scratch_decode(oracle, hostsim, corpus):
    rng = np.random.default_rng(77)
    boundaries_used = 0
    for comp, data in multi_metablock_streams(corpus):
        full = oracle.decode(comp, 1 << 22)
        cap = len(full[2]) + 17
        for chunking in range(3):
            sess = helpers.HostSimSession(hostsim, cap)
            cuts = sorted(set(int(x) for x in rng.integers(1, len(comp), size=int(rng.integers(3, 14)))) | {len(comp)})
            last_valid_bit = 0
            for c in cuts:
                code, out = sess.decode(comp[:c])
                _, ocode, oout = oracle.decode(comp[:c], cap)
                assert (code, out) == (ocode, oout), (len(comp), c, code, ocode, len(out), len(oout))
                valid, bitpos, pos = sess.resumed_at()
                if valid:
                    assert bitpos >= last_valid_bit and bitpos <= 8 * c and pos <= len(out)
                    boundaries_used += bitpos > last_valid_bit
                    last_valid_bit = bitpos
            assert code == 1 and (data is None or out == data)
    assert boundaries_used >= 6  # the sessions really resumed behind metablock boundaries


def test_resume_after_output_growth_and_errors(oracle, hostsim, corpus):
    pool = corpus.text_pool()
    data = pool[300000:300000 + 800000]
    comp = corpus.compress(data, 2)
    # the output buffer is too small at first: NEEDS_MORE_OUTPUT, then a larger session buffer with the same bytes
    sess = helpers.HostSimSession(hostsim, len(data))
    code, out = sess.decode(comp[:len(comp) // 2])
    assert code == 2 and sess.resumed_at()[0] == 1
    small = helpers.HostSimSession(hostsim, len(out) + 5)
    small.state = sess.state                     # same boundary ...
    small.buf[:len(out)] = out                   # ... and the bytes decoded so far (what the host copies on growth)
    code2, out2 = small.decode(comp)
    _, ocode, oout = oracle.decode(comp, len(out) + 5)
    assert (code2, out2) == (ocode, oout) and code2 == 3
    # a corrupted tail after a good boundary: the error code of a from-scratch decode
    bad = bytearray(comp); bad[len(comp) * 3 // 4] ^= 0x5A
    sess = helpers.HostSimSession(hostsim, len(data) + 64)
    sess.decode(bytes(bad[:len(comp) // 2]))
    code3, out3 = sess.decode(bytes(bad))
    _, ocode, oout = oracle.decode(bytes(bad), len(data) + 64)
    assert (code3, out3) == (ocode, oout)


def test_resume_with_custom_dictionary(oracle, hostsim, corpus):
    """A session whose state carries a custom dictionary (Decompressor::new_with_custom_dict): multi-metablock stream,
    arbitrary chunking, every call equal to a from-scratch decode with the same dictionary."""
    rng = np.random.default_rng(78)
    pool = corpus.text_pool()
    d = pool[600000:650000]
    data = pool[610000:610000 + 700000]
    comp = corpus.compress_with_dictionary(data, d, 2)
    assert oracle.decode(comp, len(data), True, d)[1:] == (1, data)
    for _ in range(3):
        sess = helpers.HostSimSession(hostsim, len(data) + 9, custom_dict=d)
        cuts = sorted(set(int(x) for x in rng.integers(1, len(comp), size=9)) | {len(comp)})
        for c in cuts:
            code, out = sess.decode(comp[:c])
            _, ocode, oout = oracle.decode(comp[:c], len(data) + 9, True, d)
            assert (code, out) == (ocode, oout), c
        assert code == 1 and out == data and sess.resumed_at()[0] == 1
