"""CPU-side checks of the product library: it loads, exports every symbol include/brotli_b200/decode.h
declares, keeps the reference's struct layout and enum values, and -- having no CPU decode path --
fails loudly when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "brotli_b200", "decode.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"BROTLI_B200_API[^;(]*?\b(Brotli[A-Za-z0-9]+)\s*\(", text)))


def test_header_symbols_are_bound(pkg):
    assert declared_symbols() == sorted(pkg.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    L = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), name


def test_library_does_not_link_the_oracle(pkg):
    import subprocess
    out = subprocess.run(["nm", "-D", pkg.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle_" not in out
    needed = subprocess.run(["ldd", pkg.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "liboracle" not in needed and "libbrotlidec" not in needed


def test_return_info_layout(pkg):  # c/brotli/decode.h:127-132
    RI = pkg.BrotliDecoderReturnInfo
    assert ctypes.sizeof(RI) == 8 + 256 + 4 + 4
    assert RI.error.offset == 8 and RI.result.offset == 264 and RI.code.offset == 268


def test_version_and_error_strings(pkg):
    L = pkg.lib()
    assert L.BrotliDecoderVersion() == 0x1000F00  # src/ffi/mod.rs:588-590
    assert pkg.error_string(-8) == "ERROR_FORMAT_CONTEXT_MAP_REPEAT"
    assert pkg.error_string(-6) == "ERROR_FORMAT_FL_SPACE"  # sic, src/state.rs:547
    assert pkg.error_string(1) == "SUCCESS" and pkg.error_string(-31) == "ERROR_UNREACHABLE"


def test_instance_lifecycle_without_decoding(pkg):
    L = pkg.lib()
    s = L.BrotliDecoderCreateInstance(None, None, None)
    assert s
    assert L.BrotliDecoderIsUsed(s) == 0 and L.BrotliDecoderIsFinished(s) == 0 and L.BrotliDecoderHasMoreOutput(s) == 0
    assert L.BrotliDecoderSetParameter(s, 1, 1) == 1
    L.BrotliDecoderDestroyInstance(s)
    # both callbacks or neither, src/ffi/mod.rs:132-135
    cb = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)(lambda o, n: None)
    assert not L.BrotliDecoderCreateInstance(ctypes.cast(cb, ctypes.c_void_p), None, None)


def test_argument_validation_needs_no_device(pkg):
    L = pkg.lib()
    # decoded_size == NULL -> ERROR, src/ffi/mod.rs:269-271
    assert L.BrotliDecoderDecompress(1, b"\x06", None, None) == 0
    # NULL input with non-zero length -> INVALID_ARGUMENTS, src/ffi/mod.rs:612-647
    buf = ctypes.create_string_buffer(16)
    info = L.BrotliDecoderDecompressWithReturnInfo(4, None, 16, buf)
    assert (info.result, info.code) == (0, -20)
    info = L.BrotliDecoderDecompressPrealloc(1, b"\x06", 16, buf, 8, None, 0, None, 0, None)
    assert (info.result, info.code) == (0, -20)
    misaligned = ctypes.c_void_p(ctypes.addressof(buf) + 1)
    info = L.BrotliDecoderDecompressPrealloc(1, b"\x06", 16, buf, 0, None, 2, misaligned, 0, None)
    assert (info.result, info.code) == (0, -20)


def test_no_cpu_fallback(pkg):
    """Without a GPU the decode entry points must fail, not decode on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present; covered by the gpu tests")
    info, out = pkg.brotli_decode(b"\x0b\x00\x80\x58\x03", 16)
    assert info.result == 0 and info.code == -31 and out == b""
    assert b"no CUDA device" in info.error
    r, out = pkg.BrotliDecoderDecompress(b"\x06", 16)
    assert r == 0
    with pytest.raises(pkg.BrotliB200Error):
        pkg.decompress_batch([b"\x06"], [16])
