#!/usr/bin/env python3
"""Writes tests/golden/dict_vectors.json: the two custom-dictionary known-answer tests of the reference
(src/test.rs:438-508, `test_dict` and `test_dict_medium`: compressed patch + custom LZ77 dictionary ->
exact output through BrotliState::new_with_custom_dictionary).  Run in the build container, where
/root/reference is mounted; the JSON travels with the repo."""
import json, os, re
HERE = os.path.dirname(os.path.abspath(__file__))
SRC = open("/root/reference/src/test.rs").read()


def array(body, name):
    m = re.search(r"let\s+" + name + r"\s*:\s*&\[u8\]\s*=\s*&\[(.*?)\];", body, re.S)
    return bytes(int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1)))


def fn_body(name):
    i = SRC.index("fn %s()" % name)
    j = SRC.index("#[test]", i) if "#[test]" in SRC[i:] else len(SRC)
    return SRC[i:j]


V = []
b = fn_body("test_dict")
V.append({"name": "test_dict", "source": "src/test.rs:438-479", "input_hex": array(b, "patch").hex(),
          "dict_hex": array(b, "dict").hex(), "output_hex": array(b, "expected").hex()})
b = fn_body("test_dict_medium")
V.append({"name": "test_dict_medium", "source": "src/test.rs:482-508", "input_hex": array(b, "br").hex(),
          "dict_hex": bytes(range(256)).hex(), "output_hex": array(b, "expected").hex(),
          "note": "dictionary = bytes 0..255 (built by a loop in the test)"})
json.dump({"vectors": V}, open(os.path.join(HERE, "dict_vectors.json"), "w"), indent=1)
print([(v["name"], len(v["input_hex"]) // 2, len(v["dict_hex"]) // 2, len(v["output_hex"]) // 2) for v in V])
