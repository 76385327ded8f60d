#!/usr/bin/env python3
"""Extracts the reference's known-answer tests for the Huffman-table builders (src/huffman/tests.rs: 11 table dumps)
and the bit reader (src/bit_reader/mod.rs:450-632: 10 cases in 5 tests) into tests/golden/kat_vectors.json.
Run in the build container (needs /root/reference); the JSON is committed, this script is its provenance."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ints(body):
    return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", re.sub(r"/\*.*?\*/|//[^\n]*", "", body, flags=re.S))]


def huffman_kats():
    src = open(os.path.join(REF, "src/huffman/tests.rs")).read()
    out = []
    for m in re.finditer(r"#\[test\]\nfn (\w+)\(\) \{\n(.*?)\n\}\n", src, flags=re.S):
        name, body = m.group(1), m.group(2)
        case = {"name": name, "source": "src/huffman/tests.rs"}
        for am in re.finditer(r"let (?:mut )?(\w+): \[(u8|u16); [^\]]+\]\s*=\s*\[([^\]]*)\];", body, flags=re.S):
            case[am.group(1)] = ints(am.group(3))
        tm = re.search(r"let end_table: \[HuffmanCode; [^\]]+\]\s*=\s*\[(.*?)\];", body, flags=re.S)
        entries = [[int(b), int(v)] for b, v in re.findall(r"HuffmanCode \{\s*bits: (\d+),\s*value: (\d+),?\s*\}", tm.group(1))]
        rep = re.search(r"\}\s*;\s*(\d+|1 << BROTLI_HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH)\s*$", tm.group(1).strip())
        if rep:
            entries = entries * (32 if "<<" in rep.group(1) else int(rep.group(1)))  # 1 << BROTLI_HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH = 32
        case["end_table"] = entries  # [bits, value]
        cm = re.search(r"(BrotliBuild\w+)\((.*?)\);", body, flags=re.S)
        case["function"] = cm.group(1)
        args = [a.strip() for a in cm.group(2).split(",")]
        if case["function"] == "BrotliBuildHuffmanTable":
            case["root_bits"], case["symbol_lists_offset"] = int(args[1]), int(args[3])
        if case["function"] == "BrotliBuildSimpleHuffmanTable":
            case["root_bits"], case["num_symbols"] = int(args[1]), int(args[3])
        sm = re.search(r"assert_eq!\((?:size|goal_size), (\d+)\)", body)
        if sm:
            case["size"] = int(sm.group(1))
        out.append(case)
    return out


def bitreader_kats():
    src = open(os.path.join(REF, "src/bit_reader/mod.rs")).read()
    src = src[src.index("mod tests {"):]
    out = []
    for tm in re.finditer(r"#\[test\]\n  fn (\w+)\(\) \{\n(.*?)\n  \}\n", src, flags=re.S):
        test, body = tm.group(1), tm.group(2)
        for bm in re.finditer(r"let data: \[u8; \d+\] = \[(.*?)\];(.*?)(?=let data: |\Z)", body, flags=re.S):
            blk = bm.group(2)
            st = re.search(r"BrotliBitReader \{\s*val_: (\w+),\s*bit_pos_: (\w+),\s*avail_in: (\w+),\s*next_in: (\w+),", blk)
            call = re.search(r"let ret = (\w+)\(&mut bit_reader(?:, (\d+))?(?:, &mut val)?, &data\[\.\.\]\);", blk)
            case = {"test": test, "source": "src/bit_reader/mod.rs", "data": ints(bm.group(1)),
                    "state": {"val_": int(st.group(1), 0), "bit_pos_": int(st.group(2), 0), "avail_in": int(st.group(3), 0), "next_in": int(st.group(4), 0)},
                    "function": call.group(1), "n_bits": int(call.group(2)) if call.group(2) else None}
            iv = re.search(r"let mut val: u32 = (\w+);", blk)
            if iv:
                case["val_in"] = int(iv.group(1), 0)
            exp = {}
            for am in re.finditer(r"assert_eq!\((?:bit_reader\.)?(\w+), (\w+)\);", blk):
                v = am.group(2)
                exp[am.group(1)] = (1 if v == "true" else 0) if v in ("true", "false") else int(v, 0)
            case["expect"] = exp
            out.append(case)
    return out


if __name__ == "__main__":
    h, b = huffman_kats(), bitreader_kats()
    assert len(h) == 11 and len(b) == 10, (len(h), len(b))
    json.dump({"huffman": h, "bit_reader": b}, open(os.path.join(HERE, "kat_vectors.json"), "w"), separators=(",", ":"))
    print("huffman", [(c["name"], c["function"], len(c["end_table"])) for c in h])
    print("bit_reader", [(c["test"], c["function"]) for c in b])
