"""The generated constant tables (tables/brotli_tables.h, tables/brotli_dictionary.bin; built by tables/gen_tables.py from the
RFC 7932 formulas and libbrotlicommon) against the reference's own tables: digests of the canonical forms parsed out of
src/prefix.rs, src/context.rs, src/dictionary/mod.rs and src/transform.rs (tests/golden/table_digests.json, made by
tests/golden/make_table_digests.py), plus the CRC fingerprints of SURVEY.md App. C."""
import hashlib
import json
import os
import re
import zlib

import helpers

TABLES = os.path.join(helpers.ROOT, "tables")


def c_array(src, name):
    m = re.search(r"%s\[\d+\] = \{(.*?)\};" % re.escape(name), src, flags=re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    return [int(x, 0) for x in re.findall(r"-?0x[0-9a-fA-F]+|-?\d+", body)]


def digest(obj):
    return hashlib.sha256(json.dumps(obj, separators=(",", ":")).encode()).hexdigest()


def test_tables_match_the_reference():
    ref = json.load(open(os.path.join(helpers.ROOT, "tests", "golden", "table_digests.json")))
    src = open(os.path.join(TABLES, "brotli_tables.h")).read()
    data = open(os.path.join(TABLES, "brotli_dictionary.bin"), "rb").read()
    assert len(data) == ref["dictionary_len"] == 122784 and hashlib.sha256(data).hexdigest() == ref["dictionary_sha256"]
    assert zlib.crc32(data) == 0x5136cb04
    assert digest(c_array(src, "kBrotliDictOffsetsByLength")) == ref["dict_offsets"]
    assert digest(c_array(src, "kBrotliDictSizeBitsByLength")) == ref["dict_size_bits"]
    ctx = c_array(src, "kBrotliContextLookup")
    assert digest(ctx) == ref["context_lookup"] and zlib.crc32(bytes(ctx)) == 0x6c1497b8
    blk = list(map(list, zip(c_array(src, "kBrotliBlockLengthOffset"), c_array(src, "kBrotliBlockLengthNBits"))))
    assert digest(blk) == ref["block_length"]
    lut = c_array(src, "kBrotliCmdLut")
    rows = [lut[i:i + 6] for i in range(0, len(lut), 6)]
    assert len(rows) == 704 and digest(rows) == ref["cmd_lut"]
    pool = bytes(c_array(src, "kBrotliPrefixSuffix"))
    tr = c_array(src, "kBrotliTransforms")

    def s(off):
        return pool[off:pool.index(0, off)].hex()
    transforms = [[s(tr[3 * i]), tr[3 * i + 1], s(tr[3 * i + 2])] for i in range(121)]
    assert digest(transforms) == ref["transforms"]


def test_oracle_and_product_embed_the_same_dictionary(oracle):
    """Both libraries incbin tables/brotli_dictionary.bin: word 0 of length 4 is 'time' (RFC 7932 App. A)."""
    data = open(os.path.join(TABLES, "brotli_dictionary.bin"), "rb").read()
    assert data[:16] == b"timedownlifeleft"
