"""The reference's known-answer tests for the pieces under the decoder, run against the oracle's restatements
(oracle/brotli_oracle.h "pieces exported for the reference's known-answer tests"): the 11 Huffman-table dumps of
src/huffman/tests.rs (:5, :149, :293, :315, :1684, :3485, :4582-:7724) and the 10 bit-reader cases of
src/bit_reader/mod.rs:450-632.  Vectors: tests/golden/kat_vectors.json (made by tests/golden/make_kat_vectors.py)."""
import ctypes
import json
import os

import pytest

import helpers

KATS = json.load(open(os.path.join(helpers.ROOT, "tests", "golden", "kat_vectors.json")))


def table_of(buf, n):
    return [[buf[i].bits, buf[i].value] for i in range(n)]


@pytest.mark.parametrize("case", KATS["huffman"], ids=lambda c: c["name"])
def test_huffman_table_builders(oracle, case):
    L = oracle.lib
    n = len(case["end_table"])
    table = (helpers.OracleHuffmanCode * max(n, 1100))()
    if case["function"] == "BrotliBuildCodeLengthsHuffmanTable":  # src/huffman/mod.rs:196-271
        cl = (ctypes.c_uint8 * len(case["code_lengths"]))(*case["code_lengths"])
        count = (ctypes.c_uint16 * 16)(*case["count"])
        L.oracle_build_code_lengths_huffman_table(table, cl, count)
    elif case["function"] == "BrotliBuildHuffmanTable":           # src/huffman/mod.rs:273-386
        sym = (ctypes.c_uint16 * len(case["symbol_array"]))(*case["symbol_array"])
        counts = (ctypes.c_uint16 * 16)(*case["counts"])
        L.oracle_build_huffman_table.restype = ctypes.c_uint32
        L.oracle_build_huffman_table.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        size = L.oracle_build_huffman_table(table, case["root_bits"], sym, case["symbol_lists_offset"], counts)
        assert size == case["size"] <= n  # (the table of `singlelevel` has four untouched entries behind the 256 built)
        assert list(counts)[:len(case["end_counts"])] == case["end_counts"]
    else:                                                         # BrotliBuildSimpleHuffmanTable, src/huffman/mod.rs:390-471
        val = (ctypes.c_uint16 * len(case["val"]))(*case["val"])
        L.oracle_build_simple_huffman_table.restype = ctypes.c_uint32
        L.oracle_build_simple_huffman_table.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32]
        size = L.oracle_build_simple_huffman_table(table, case["root_bits"], val, len(case["val"]), case["num_symbols"])
        assert size == case["size"] <= n  # (the table of `singlelevel` has four untouched entries behind the 256 built)
    assert table_of(table, n) == case["end_table"]


@pytest.mark.parametrize("case", KATS["bit_reader"], ids=lambda c: "%s-%s" % (c["test"], c["function"]))
def test_bit_reader(oracle, case):
    L = oracle.lib
    data = bytes(case["data"])
    st = case["state"]
    br = helpers.OracleBitReader(st["val_"], st["bit_pos_"], st["next_in"], st["avail_in"])
    exp = dict(case["expect"])
    fn = case["function"]
    if fn == "BrotliWarmupBitReader":
        L.oracle_br_warmup.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        ret = L.oracle_br_warmup(ctypes.byref(br), data)
    elif fn == "BrotliSafeReadBits":
        val = ctypes.c_uint32(case.get("val_in", 0))
        L.oracle_br_safe_read_bits.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_char_p]
        ret = L.oracle_br_safe_read_bits(ctypes.byref(br), case["n_bits"], ctypes.byref(val), data)
        assert val.value == exp.pop("val")
    else:
        f = {"BrotliReadBits": L.oracle_br_read_bits, "BrotliReadConstantNBits": L.oracle_br_read_constant_n_bits,
             "BrotliGet16BitsUnmasked": L.oracle_br_get16_bits_unmasked}[fn]
        f.restype = ctypes.c_uint32
        if fn == "BrotliGet16BitsUnmasked":
            f.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
            ret = f(ctypes.byref(br), data)
        else:
            f.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p]
            ret = f(ctypes.byref(br), case["n_bits"], data)
    assert ret == exp.pop("ret")
    for k, v in exp.items():
        assert getattr(br, k) == v, (k, getattr(br, k), v)
