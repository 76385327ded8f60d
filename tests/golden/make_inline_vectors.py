#!/usr/bin/env python3
"""Writes tests/golden/inline_vectors.json: the byte vectors the reference's own unit tests
decode inline.  `expect` is "success" with the exact output, or "failure"
(optionally with the BrotliDecoderErrorCode the reference's test asserts)."""
import json, os
HERE = os.path.dirname(os.path.abspath(__file__))
H = bytes.fromhex
V = []
def add(name, src, data, expect, out=None, code=None, note=None):
    e = {"name": name, "source": src, "input_hex": bytes(data).hex(), "expect": expect}
    if out is not None: e["output_hex"] = bytes(out).hex()
    if code is not None: e["code"] = code
    if note: e["note"] = note
    V.append(e)

add("10x10y", "src/test.rs:176-194", [0x1b,0x13,0x00,0x00,0xa4,0xb0,0xb2,0xea,0x81,0x47,0x02,0x8a], "success", b"X"*10+b"Y"*10)
add("x", "src/test.rs:199-211", [0x0b,0x00,0x80,0x58,0x03], "success", b"X")
add("empty", "src/test.rs:226-237", [0x06], "success", b"")
add("corrupt_large_distance_code", "src/test.rs:214-223",
    [17,139,32,255,8,0,136,255,32,46,146,32,255,255,255,255]+[32]*30, "failure")
QF = [0x5B,0xFF,0xAF,0x02,0xC0,0x22,0x79,0x5C,0xFB,0x5A,0x8C,0x42,0x3B,0xF4,0x25,0x55,0x19,0x5A,0x92,0x99,0xB1,0x35,0xC8,0x19,
      0x9E,0x9E,0x0A,0x7B,0x4B,0x90,0xB9,0x3C,0x98,0xC8,0x09,0x40,0xF3,0xE6,0xD9,0x4D,0xE4,0x6D,0x65,0x1B,0x27,0x87,0x13,0x5F,
      0xA6,0xE9,0x30,0x96,0x7B,0x3C,0x15,0xD8,0x53,0x1C]
add("quickfox_repeated", "src/test.rs:244-290", QF, "success", b"The quick brown fox jumps over the lazy dog"*4096)
add("emergency_broadcast", "c/main.c:16-53",
    H("1b3000e08dd4592d3937b5024810952a9aea420e51a416b9cbf5f85c64b92fc96a3fb1dca8e03507"), "success_prefix",
    b"THIS IS A TEST OF THE EMERGENCY BROADCAST SYSTEM", note="c/main.c compares the first sizeof(key)-1 bytes")
add("emergency_broadcast_corrupt", "c/main.c:55-77",
    H("1b3000e08dd4592d39ffb5024810952a9aea420e51a416b9cbf5f85c64b92fc96a3fb1dca8e03507"), "failure", code=-8)
add("himselfself", "src/bin/integration_tests.rs:900-910",
    H("1b0a00000000" "80e3b40d0000" "075b26314002" "00e04e1ba180" "2000"), "success", b"himselfself")
add("scrollroll", "src/bin/integration_tests.rs:912-923",
    H("1b0900000000" "80e3b40d0000" "075b26314002" "00e04e1b21a0" "2000"), "success", b"scrollroll")
add("leftdatadataleft", "src/bin/integration_tests.rs:925-936",
    H("1b0f00000000" "80e3b40d0000" "075b26314002" "00e04e1b4180" "205010240806"), "success", b"leftdatadataleft")
add("ff_x8", "src/bin/ffi_stream_tests.rs:77", H("1f0700f827fe43840000"), "success", b"\xff"*8)
add("hello_trailing", "src/reader.rs:359,397; src/writer.rs:385", H("8f028068656c6c6f0a03") + b"trailing garbage", "success", b"hello\n",
    note="one-shot ignores trailing bytes (App. D-2)")
FOX69 = H("1b4a0000c4f4a469bd79252d22b452ea830d38706" "8b271c041761e36c6ce1384e836f22a0ce789687a04492faaf731a19b0d48b7f01f483342a59c312697a9c6be67855202")
add("fox69", "src/bin/error_handling_tests.rs:184-197", FOX69, "success",
    b"the quick brown fox jumps over the lazy dog twice for redundancy and length")
for off, x in [(13, 0x01), (23, 0x01), (33, 0x55)]:
    c = bytearray(FOX69); c[off] ^= x
    add("fox69_flip_%d_%02x" % (off, x), "src/bin/error_handling_tests.rs:202-219", c, "failure")
add("mlen_overflow", "src/bin/error_handling_tests.rs:231-256", H(
    "0b2b01008cd4484d73bb8171bab646fb2203e581794e0cb52237983782d1e0880c00200849a0f3a41ab32b9"
    "8909e575ad3dc35383ab3bc757871fff6ab3a7135eeae1f988262c82ca1426ff34692b30a2beadb7a8589f9"
    "b5dd93eba74fd1f0b376393efe188bc0d3b812b5c91988a7e5965437750e8c4c2f6d1e9cdfbdfe2876540ed"
    "bcb1b222168125320d7593de1a4cc82f2bad61e7e7c6e75e7f8eaf1e35ff7d2663edc7f2803c751396295d1"
    "e18fa5e614573576f40f4f2d6eec9fddbe7ccb5658f49bf30b24c02822832fd35adca1c48cfcb2da966e6e6"
    "c7665fbe8f2e1fd4f73937adadfbe080dc352d822a5c1ee8ba664175536b4f70d4d2eacef9dde3c7f496690"
    "77ebe909e02024813e3000b81200c0488dd434b70b8000286308abc0fb1d453300b81200c0488dd434b70b8"
    "0002863088be0b73600b81200c0488dd434b70b80002863088be0b73600b81200c0488dd434b70b80002863"
    "088be0b73600b81200c0488dd434b70b80c88fc6428cc0bb0103"), "failure")
# src/bin/tests.rs:76-80: of the 256 one-byte streams exactly these decode (to nothing)
json.dump({"vectors": V, "one_byte_ok": [6, 26, 51, 53, 55, 57, 59, 61, 63]},
          open(os.path.join(HERE, "inline_vectors.json"), "w"), indent=1)
print(len(V), "vectors")
