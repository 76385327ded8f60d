#!/usr/bin/env python3
"""SHA-256 digests of the reference's constant tables (parsed out of its Rust sources) in a canonical encoding, so that
tests/test_tables.py can check the generated tables/brotli_tables.h + brotli_dictionary.bin against the reference
without the reference being present.  Run in the build container; writes tests/golden/table_digests.json."""
import hashlib
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ints(body):
    body = re.sub(r"/\*.*?\*/|//[^\n]*", "", body, flags=re.S)
    return [int(x, 0) for x in re.findall(r"-?0x[0-9a-fA-F]+|-?\d+", body)]


def digest(obj):
    return hashlib.sha256(json.dumps(obj, separators=(",", ":")).encode()).hexdigest()


def canonical_tables_from_reference():
    prefix = open(os.path.join(REF, "src/prefix.rs")).read()
    blk = re.search(r"kBlockLengthPrefixCode: \[PrefixCodeRange; 26\] = \[(.*?)\];", prefix, flags=re.S).group(1)
    block_length = [[int(o), int(n)] for o, n in re.findall(r"offset: (\d+),\s*nbits: (\d+)", blk)]
    lut = re.search(r"kCmdLut: \[CmdLutElement; 704\] = \[(.*?)\];", prefix, flags=re.S).group(1)
    cmd = [[int(a, 0), int(b, 0), int(c, 0), int(d, 0), int(e, 0), int(f, 0)] for a, b, c, d, e, f in re.findall(
        r"insert_len_extra_bits: (\w+),\s*copy_len_extra_bits: (\w+),\s*distance_code: (-?\w+),\s*context: (\w+),\s*insert_len_offset: (\w+),\s*copy_len_offset: (\w+)", lut)]
    ctx_src = open(os.path.join(REF, "src/context.rs")).read()
    ctx = ints(ctx_src[ctx_src.index("pub static kContextLookup:[[u8;512];4] = [") + len("pub static kContextLookup:[[u8;512];4] = ["):])[:2048]
    d = open(os.path.join(REF, "src/dictionary/mod.rs")).read()
    offsets = ints(re.search(r"kBrotliDictionaryOffsetsByLength: \[u32; 25\] =\s*\[(.*?)\];", d, flags=re.S).group(1))
    size_bits = ints(re.search(r"kBrotliDictionarySizeBitsByLength: \[u8; 25\] =\s*\[(.*?)\];", d, flags=re.S).group(1))
    data = bytes(ints(re.search(r"kBrotliDictionary: \[u8; 122784\] =\s*\[(.*?)\];", d, flags=re.S).group(1)))
    t = open(os.path.join(REF, "src/transform.rs")).read()
    consts = {k: int(v, 0) for k, v in re.findall(r"const (k\w+): u8 = (\w+);", t)}
    pool = bytes(ints(re.search(r"const kPrefixSuffix: \[u8; \d+\] =\s*\[(.*?)\];", t, flags=re.S).group(1)))

    def s(off):
        return pool[off:pool.index(0, off)].hex()
    tr = re.search(r"kTransforms: \[Transform; kNumTransforms as usize\] = \[(.*?)\];", t, flags=re.S).group(1)
    transforms = [[s(consts[p]), consts[ty], s(consts[sx])] for p, ty, sx in re.findall(r"prefix_id: (\w+),\s*transform: (\w+),\s*suffix_id: (\w+)", tr)]
    return {"block_length": block_length, "cmd_lut": cmd, "context_lookup": ctx, "dict_offsets": offsets, "dict_size_bits": size_bits,
            "dictionary_sha256": hashlib.sha256(data).hexdigest(), "dictionary_len": len(data), "transforms": transforms}


if __name__ == "__main__":
    t = canonical_tables_from_reference()
    assert len(t["block_length"]) == 26 and len(t["cmd_lut"]) == 704 and len(t["context_lookup"]) == 2048 and len(t["transforms"]) == 121
    out = {k: (v if isinstance(v, (str, int)) else digest(v)) for k, v in t.items()}
    out["_provenance"] = "sha256 of json.dumps(canonical form, separators=(',',':')) of the tables parsed from src/prefix.rs, src/context.rs, src/dictionary/mod.rs, src/transform.rs"
    json.dump(out, open(os.path.join(HERE, "table_digests.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))
