#!/usr/bin/env python3
"""Generate the RFC 7932 constant tables used by both the CPU oracle and the CUDA path.

Nothing is transcribed from the reference crate: every table is either derived from the
RFC 7932 formulas here, or read out of the system `libbrotlicommon.so.1` (Google's C brotli,
which exports the RFC appendix constants), and then fingerprint-checked:

  * dictionary            RFC 7932 App. A   CRC-32 0x5136cb04   (= /root/reference/src/dictionary/mod.rs:18)
  * context lookup        RFC 7932 s7.1     CRC-32 0x6c1497b8   (= src/context.rs:112, order LSB6,MSB6,UTF8,SIGNED)
  * transforms            RFC 7932 App. B   (= src/transform.rs:32-716)
  * insert&copy LUT       RFC 7932 s5       (= src/prefix.rs:116, kCmdLut[704])
  * block length codes    RFC 7932 s6       (= src/prefix.rs:9-113)

`tests/test_tables.py` re-checks the fingerprints and, when /root/reference is present,
compares every table against the values parsed out of the reference sources.

Outputs (committed): tables/brotli_dictionary.bin, tables/brotli_tables.h
"""
import ctypes
import os
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))

INS_BASE = [0, 1, 2, 3, 4, 5, 6, 8, 10, 14, 18, 26, 34, 50, 66, 98, 130, 194, 322, 578, 1090, 2114, 6210, 22594]
INS_EXTRA = [0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 12, 14, 24]
COPY_BASE = [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 18, 22, 30, 38, 54, 70, 102, 134, 198, 326, 582, 1094, 2118]
COPY_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 24]
# RFC 7932 section 5: insert-and-copy cell (code >> 6) -> (insert code base, copy code base)
CELLS = [(0, 0), (0, 8), (0, 0), (0, 8), (8, 0), (8, 8), (0, 16), (16, 0), (8, 16), (16, 8), (16, 16)]


def block_length_codes():
    """RFC 7932 section 6: 26 block count codes; base grows by 1 << nbits."""
    nbits = [2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12, 13, 24]
    off, out = 1, []
    for n in nbits:
        out.append((off, n))
        off += 1 << n
    return out


def cmd_lut():
    out = []
    for code in range(704):
        cell = code >> 6
        ibase, cbase = CELLS[cell]
        ic = ibase + ((code >> 3) & 7)
        cc = cbase + (code & 7)
        out.append(dict(
            insert_len_extra_bits=INS_EXTRA[ic], copy_len_extra_bits=COPY_EXTRA[cc],
            distance_code=0 if cell < 2 else -1, context=min(COPY_BASE[cc] - 2, 3),
            insert_len_offset=INS_BASE[ic], copy_len_offset=COPY_BASE[cc]))
    return out


class _Dict(ctypes.Structure):
    _fields_ = [("size_bits_by_length", ctypes.c_uint8 * 32),
                ("offsets_by_length", ctypes.c_uint32 * 32),
                ("data_size", ctypes.c_size_t),
                ("data", ctypes.POINTER(ctypes.c_uint8))]


class _Transforms(ctypes.Structure):
    _fields_ = [("prefix_suffix_size", ctypes.c_uint16),
                ("prefix_suffix", ctypes.POINTER(ctypes.c_uint8)),
                ("prefix_suffix_map", ctypes.POINTER(ctypes.c_uint16)),
                ("num_transforms", ctypes.c_uint32),
                ("transforms", ctypes.POINTER(ctypes.c_uint8)),
                ("params", ctypes.POINTER(ctypes.c_uint8)),
                ("cutOffTransforms", ctypes.c_int16 * 10)]


def from_libbrotlicommon():
    lib = ctypes.CDLL("libbrotlicommon.so.1")
    lib.BrotliGetDictionary.restype = ctypes.POINTER(_Dict)
    d = lib.BrotliGetDictionary().contents
    data = bytes(d.data[:d.data_size])
    size_bits = list(d.size_bits_by_length)[:25]
    offsets = list(d.offsets_by_length)[:25]
    ctx = bytes((ctypes.c_uint8 * 2048).in_dll(lib, "_kBrotliContextLookupTable"))
    lib.BrotliGetTransforms.restype = ctypes.POINTER(_Transforms)
    t = lib.BrotliGetTransforms().contents
    pool = bytes(t.prefix_suffix[:t.prefix_suffix_size])

    def pstr(idx):  # libbrotli stores length-prefixed strings, indexed through a map
        off = t.prefix_suffix_map[idx]
        n = pool[off]
        return pool[off + 1: off + 1 + n]
    transforms = []
    for i in range(t.num_transforms):
        p, ty, s = t.transforms[3 * i], t.transforms[3 * i + 1], t.transforms[3 * i + 2]
        transforms.append((pstr(p), ty, pstr(s)))
    return data, size_bits, offsets, ctx, transforms


def c_array(name, ctype, values, per_line=16, fmt="{}"):
    lines = []
    for i in range(0, len(values), per_line):
        lines.append("  " + ", ".join(fmt.format(v) for v in values[i:i + per_line]) + ",")
    return "BROTLI_TABLE_ATTR %s %s[%d] = {\n%s\n};\n" % (ctype, name, len(values), "\n".join(lines))


def main():
    data, size_bits, offsets, ctx, transforms = from_libbrotlicommon()
    assert len(data) == 122784 and zlib.crc32(data) == 0x5136cb04, "dictionary fingerprint"
    assert zlib.crc32(ctx) == 0x6c1497b8, "context LUT fingerprint"
    assert len(transforms) == 121
    assert transforms[0] == (b"", 0, b"") and transforms[1] == (b"", 0, b" ")
    assert transforms[3] == (b"", 12, b"") and transforms[49] == (b"", 1, b"ing ")
    assert transforms[120] == (b" ", 10, b"='")
    with open(os.path.join(HERE, "brotli_dictionary.bin"), "wb") as f:
        f.write(data)

    # NUL-terminated prefix/suffix string pool, deduplicated, empty string at offset 0.
    pool, where = bytearray(b"\0"), {b"": 0}
    for p, _, s in transforms:
        for x in (p, s):
            if x not in where:
                where[x] = len(pool)
                pool += x + b"\0"
    assert len(pool) < 256
    tr_flat = []
    for p, ty, s in transforms:
        tr_flat += [where[p], ty, where[s]]

    lut = cmd_lut()
    blk = block_length_codes()
    assert blk[-1] == (16625, 24) and blk[12] == (113, 5)

    h = []
    h.append("/* GENERATED by tables/gen_tables.py -- do not edit.\n"
             " * RFC 7932 constants shared by oracle/ (CPU checker) and the CUDA decoder.\n"
             " * Reference counterparts: src/prefix.rs, src/context.rs, src/transform.rs:32-716,\n"
             " * src/dictionary/mod.rs:3-15 of dropbox/rust-brotli-decompressor. */\n"
             "#ifndef BROTLI_B200_TABLES_H_\n#define BROTLI_B200_TABLES_H_\n#include <stdint.h>\n"
             "/* storage class of the arrays; the CUDA TU defines it as `__device__ const` */\n"
             "#ifndef BROTLI_TABLE_ATTR\n#define BROTLI_TABLE_ATTR static const\n#endif\n\n")
    h.append("#define BROTLI_DICTIONARY_SIZE 122784\n"
             "#define BROTLI_MIN_DICTIONARY_WORD_LENGTH 4\n"
             "#define BROTLI_MAX_DICTIONARY_WORD_LENGTH 24\n"
             "#define BROTLI_NUM_TRANSFORMS 121\n\n")
    h.append(c_array("kBrotliDictOffsetsByLength", "uint32_t", offsets, 8))
    h.append(c_array("kBrotliDictSizeBitsByLength", "uint8_t", size_bits, 25))
    h.append("/* [mode*512 + i]: i<256 -> f(p1), i>=256 -> g(p2); modes LSB6, MSB6, UTF8, SIGNED */\n")
    h.append(c_array("kBrotliContextLookup", "uint8_t", list(ctx), 32))
    h.append("/* NUL-terminated prefix/suffix strings; kBrotliTransforms[i] = {prefix_off, type, suffix_off} */\n")
    h.append(c_array("kBrotliPrefixSuffix", "uint8_t", list(pool), 16, "0x{:02x}"))
    h.append(c_array("kBrotliTransforms", "uint8_t", tr_flat, 15))
    h.append("enum { BROTLI_TRANSFORM_IDENTITY = 0, BROTLI_TRANSFORM_OMIT_LAST_1 = 1, BROTLI_TRANSFORM_OMIT_LAST_9 = 9,\n"
             "       BROTLI_TRANSFORM_UPPERCASE_FIRST = 10, BROTLI_TRANSFORM_UPPERCASE_ALL = 11,\n"
             "       BROTLI_TRANSFORM_OMIT_FIRST_1 = 12, BROTLI_TRANSFORM_OMIT_FIRST_9 = 20 };\n\n")
    h.append(c_array("kBrotliBlockLengthOffset", "uint16_t", [o for o, _ in blk], 13))
    h.append(c_array("kBrotliBlockLengthNBits", "uint8_t", [n for _, n in blk], 26))
    h.append(c_array("kBrotliInsertBase", "uint16_t", INS_BASE, 12))
    h.append(c_array("kBrotliInsertExtra", "uint8_t", INS_EXTRA, 24))
    h.append(c_array("kBrotliCopyBase", "uint16_t", COPY_BASE, 12))
    h.append(c_array("kBrotliCopyExtra", "uint8_t", COPY_EXTRA, 24))
    h.append("typedef struct BrotliCmdLutElement {\n"
             "  uint8_t insert_len_extra_bits;\n  uint8_t copy_len_extra_bits;\n"
             "  int8_t distance_code; /* 0: reuse last distance (no distance symbol), -1: read one */\n"
             "  uint8_t context;      /* distance context 0..3 from copy length */\n"
             "  uint16_t insert_len_offset;\n  uint16_t copy_len_offset;\n} BrotliCmdLutElement;\n")
    rows = ["  {%d, %d, %d, %d, %d, %d}," % (e["insert_len_extra_bits"], e["copy_len_extra_bits"],
                                            e["distance_code"], e["context"], e["insert_len_offset"],
                                            e["copy_len_offset"]) for e in lut]
    h.append("BROTLI_TABLE_ATTR BrotliCmdLutElement kBrotliCmdLut[704] = {\n" + "\n".join(rows) + "\n};\n")
    h.append("\n#endif  /* BROTLI_B200_TABLES_H_ */\n")
    with open(os.path.join(HERE, "brotli_tables.h"), "w") as f:
        f.write("".join(h))
    print("wrote brotli_dictionary.bin (%d B) and brotli_tables.h" % len(data))


if __name__ == "__main__":
    main()
