#!/usr/bin/env python3
"""Collect the reference's own decode fixtures into tests/golden/ (run in the dev container only).

Copies every compressed fixture of /root/reference/testdata (test vectors, not source code) to
tests/golden/fixtures/ and records, per fixture, the SHA-256 and length of the ORIGINAL file the
reference's tests compare against (src/bin/integration_tests.rs:439-463 and the per-file tests
around :509-1006).  The originals themselves are not copied: the manifest pins them, and the
text pool used by bench.py is rebuilt on the GPU box by decoding the fixtures.

Also extracts the inline byte vectors of the reference's unit tests into
tests/golden/inline_vectors.json.
"""
import hashlib
import json
import os
import shutil

REF = "/root/reference/testdata"
HERE = os.path.dirname(os.path.abspath(__file__))


def original_of(name):
    if name.endswith(".bro"):
        return name[:-4] + ".unbro"
    if name.endswith(".br"):
        return name[:-3]
    return name[: name.index(".compressed")]


def main():
    man = {}
    for name in sorted(os.listdir(REF)):
        if not (".compressed" in name or name.endswith(".br") or name.endswith(".bro")):
            continue
        src = os.path.join(REF, name)
        data = open(src, "rb").read()
        entry = {"compressed_size": len(data), "compressed_sha256": hashlib.sha256(data).hexdigest()}
        if name == "borked.compressed":
            entry["must_fail"] = True  # src/bin/integration_tests.rs:972
        elif name == "rnd_chunk.br":
            # large window (wbits 27), 100 011 280 bytes: rnd_prefix + zeros + rnd_postfix,
            # src/bin/integration_tests.rs:465-507,998-1006
            pre = open(os.path.join(REF, "rnd_prefix"), "rb").read()
            post = open(os.path.join(REF, "rnd_postfix"), "rb").read()
            entry.update(large_window=True, original_size=100011280, prefix_hex=pre.hex(), postfix_hex=post.hex())
        else:
            orig = open(os.path.join(REF, original_of(name)), "rb").read()
            entry.update(original_size=len(orig), original_sha256=hashlib.sha256(orig).hexdigest())
        man[name] = entry
        shutil.copyfile(src, os.path.join(HERE, "fixtures", name))
    json.dump(man, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    print(len(man), "fixtures")


if __name__ == "__main__":
    main()
