"""Synthetic Brotli batches for tests/ and bench.py.

Measurement/test infrastructure, not part of the decode path.  Inputs are cut from pools that
are rebuilt from the reference's own fixtures committed under tests/golden/fixtures (decoded
with the system libbrotlidec and pinned by the SHA-256 in tests/golden/manifest.json), then
compressed with the system libbrotlienc 1.1.0 -- the only encoder available offline.
"""
import ctypes
import hashlib
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(_ROOT, "tests", "golden")

TEXT_FILES = ["lcet10.txt", "plrabn12.txt", "alice29.txt", "asyoulik.txt"]
MIX_FILES = ["metablock_reset", "mapsdatazrh", "reducetostream.map", "random_then_unicode"]

_enc = _dec = None


def _encoder():
    global _enc
    if _enc is None:
        _enc = ctypes.CDLL("libbrotlienc.so.1")
        _enc.BrotliEncoderCompress.restype = ctypes.c_int
        _enc.BrotliEncoderCompress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p,
                                               ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p]
        _enc.BrotliEncoderMaxCompressedSize.restype = ctypes.c_size_t
        _enc.BrotliEncoderMaxCompressedSize.argtypes = [ctypes.c_size_t]
    return _enc


def _system_decoder():
    global _dec
    if _dec is None:
        _dec = ctypes.CDLL("libbrotlidec.so.1")
        _dec.BrotliDecoderDecompress.restype = ctypes.c_int
        _dec.BrotliDecoderDecompress.argtypes = [ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p]
    return _dec


def compress(data, quality=5, lgwin=22):
    """libbrotlienc BrotliEncoderCompress(quality, lgwin, BROTLI_MODE_GENERIC)."""
    data = bytes(data)
    enc = _encoder()
    cap = enc.BrotliEncoderMaxCompressedSize(len(data)) or (len(data) + 1024)
    buf = ctypes.create_string_buffer(cap)
    size = ctypes.c_size_t(cap)
    if not enc.BrotliEncoderCompress(int(quality), int(lgwin), 0, len(data), data, ctypes.byref(size), buf):
        raise RuntimeError("BrotliEncoderCompress failed")
    return buf.raw[:size.value]


def compress_with_dictionary(data, dictionary, quality=5, lgwin=22):
    """libbrotlienc streaming encoder with a raw LZ77 dictionary attached (BrotliEncoderPrepareDictionary +
    BrotliEncoderAttachPreparedDictionary, libbrotli >= 1.1).  While dictionary + data fit the window, the
    stream is what the reference decodes with BrotliState::new_with_custom_dictionary: distances past the start
    of the output reach back into the dictionary, and static-dictionary references come after it."""
    enc = _encoder()
    data, dictionary = bytes(data), bytes(dictionary)
    if len(data) + len(dictionary) + 16 > (1 << lgwin):
        raise ValueError("dictionary + data must fit the window for custom-dictionary semantics")
    enc.BrotliEncoderCreateInstance.restype = ctypes.c_void_p
    enc.BrotliEncoderCreateInstance.argtypes = [ctypes.c_void_p] * 3
    enc.BrotliEncoderSetParameter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32]
    enc.BrotliEncoderPrepareDictionary.restype = ctypes.c_void_p
    enc.BrotliEncoderPrepareDictionary.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_int] + [ctypes.c_void_p] * 3
    enc.BrotliEncoderAttachPreparedDictionary.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    enc.BrotliEncoderDestroyPreparedDictionary.argtypes = [ctypes.c_void_p]
    enc.BrotliEncoderDestroyInstance.argtypes = [ctypes.c_void_p]
    enc.BrotliEncoderCompressStream.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_void_p),
                                                ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
    st = enc.BrotliEncoderCreateInstance(None, None, None)
    prepared = enc.BrotliEncoderPrepareDictionary(0, len(dictionary), dictionary, int(quality), None, None, None)  # 0: BROTLI_SHARED_DICTIONARY_RAW
    try:
        if not st or not prepared:
            raise RuntimeError("libbrotlienc: cannot create the encoder / prepare the dictionary")
        enc.BrotliEncoderSetParameter(st, 1, int(quality))  # BROTLI_PARAM_QUALITY
        enc.BrotliEncoderSetParameter(st, 2, int(lgwin))    # BROTLI_PARAM_LGWIN
        if not enc.BrotliEncoderAttachPreparedDictionary(st, prepared):
            raise RuntimeError("BrotliEncoderAttachPreparedDictionary failed")
        cap = enc.BrotliEncoderMaxCompressedSize(len(data)) or (len(data) + 1024)
        out = ctypes.create_string_buffer(cap + 64)
        src = ctypes.create_string_buffer(data, len(data) or 1)
        avail_in, next_in = ctypes.c_size_t(len(data)), ctypes.c_void_p(ctypes.addressof(src))
        avail_out, next_out = ctypes.c_size_t(cap + 64), ctypes.c_void_p(ctypes.addressof(out))
        if not enc.BrotliEncoderCompressStream(st, 2, ctypes.byref(avail_in), ctypes.byref(next_in), ctypes.byref(avail_out),
                                               ctypes.byref(next_out), None) or avail_in.value != 0:  # 2: BROTLI_OPERATION_FINISH
            raise RuntimeError("BrotliEncoderCompressStream failed")
        return out.raw[:cap + 64 - avail_out.value]
    finally:
        if prepared:
            enc.BrotliEncoderDestroyPreparedDictionary(prepared)
        if st:
            enc.BrotliEncoderDestroyInstance(st)


def system_decompress(data, capacity):
    """Google C decoder (libbrotlidec 1.1.0) one-shot; returns (ok, bytes)."""
    data = bytes(data)
    buf = ctypes.create_string_buffer(max(capacity, 1))
    size = ctypes.c_size_t(capacity)
    r = _system_decoder().BrotliDecoderDecompress(len(data), data, ctypes.byref(size), buf)
    return r == 1, buf.raw[:size.value]


def manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def fixture(name):
    return open(os.path.join(GOLDEN, "fixtures", name), "rb").read()


def original(name, man=None):
    """The original file of fixture `<name>.compressed`, rebuilt by decoding it and checked against the manifest."""
    man = man or manifest()
    e = man[name + ".compressed"]
    ok, data = system_decompress(fixture(name + ".compressed"), e["original_size"])
    if not ok or hashlib.sha256(data).hexdigest() != e["original_sha256"]:
        raise RuntimeError("fixture %s does not reproduce its original" % name)
    return data


def text_pool():
    """"enwik-like" text pool: lcet10 | plrabn12 | alice29 | asyoulik (1 185 883 bytes)."""
    man = manifest()
    return b"".join(original(n, man) for n in TEXT_FILES)


def mix_pools():
    """Families of the "Silesia-mix" pool available offline (bb.binast is not shipped: 12 MB, uncompressed only)."""
    man = manifest()
    fams = {"text": text_pool()}
    for n in MIX_FILES:
        fams[n] = original(n, man)
    return fams


def cut_windows(pool, n, size, rng):
    pool = np.frombuffer(pool, dtype=np.uint8)
    offs = rng.integers(0, max(len(pool) - size, 1), size=n)
    return [pool[o:o + size].tobytes() for o in offs]


def compress_many(originals, qualities, lgwin=22, threads=None):
    threads = threads or os.cpu_count() or 1
    if np.isscalar(qualities):
        qualities = [int(qualities)] * len(originals)
    with ThreadPoolExecutor(max_workers=threads) as ex:  # ctypes releases the GIL inside the encoder
        return list(ex.map(lambda a: compress(a[0], a[1], lgwin), zip(originals, qualities)))


def make_config(config, n_unique, size=None, seed=None, threads=None):
    """Unique streams of one BASELINE config.  Returns (compressed list, originals list, description).

    headline/C2: text pool, 64 KiB, q5, lgwin 22.   C3: 4 KiB, q4, 50% text / 50% linker map + maps tile.
    C4: long streams, q5, lgwin 24, concatenated windows of 64 KiB-1 MiB from the mix pool.
    C5: mix pool, 64 KiB, quality 1 + (i mod 11)."""
    ids = {"headline": 0, "C2": 2, "C3": 3, "C4": 4, "C5": 5}
    rng = np.random.default_rng(seed if seed is not None else 0xB2000000 + ids[config])
    if config in ("headline", "C2"):
        size = size or 65536
        originals = cut_windows(text_pool(), n_unique, size, rng)
        q, lgwin, desc = 5, 22, "text pool, %d B windows, q5 lgwin 22" % size
        comp = compress_many(originals, q, lgwin, threads)
    elif config == "C3":
        size = size or 4096
        fams = mix_pools()
        web = fams["reducetostream.map"] + fams["mapsdatazrh"]
        originals = cut_windows(fams["text"], n_unique // 2, size, rng) + cut_windows(web, n_unique - n_unique // 2, size, rng)
        order = rng.permutation(len(originals))
        originals = [originals[i] for i in order]
        q, lgwin, desc = 4, 22, "50%% text / 50%% linker-map+maps, %d B windows, q4 lgwin 22" % size
        comp = compress_many(originals, q, lgwin, threads)
    elif config == "C4":
        size = size or (16 << 20)
        fams = mix_pools()
        pool = np.frombuffer(b"".join(fams.values()), dtype=np.uint8)
        originals = []
        for _ in range(n_unique):
            parts, total = [], 0
            while total < size:
                w = int(rng.integers(65536, 1 << 20))
                o = int(rng.integers(0, len(pool) - w))
                parts.append(pool[o:o + w]); total += w
            originals.append(np.concatenate(parts)[:size].tobytes())
        q, lgwin, desc = 5, 24, "mix pool windows 64 KiB-1 MiB concatenated to %d B, q5 lgwin 24" % size
        comp = compress_many(originals, q, lgwin, threads)
    elif config == "C5":
        size = size or 65536
        fams = list(mix_pools().values())
        originals = []
        for i in range(n_unique):
            r = rng.random()
            if r < 0.01:
                originals.append(bytes(size))
            elif r < 0.02:
                originals.append(rng.integers(0, 256, size=size, dtype=np.uint8).tobytes())
            else:
                originals.append(cut_windows(fams[int(rng.integers(0, len(fams)))], 1, size, rng)[0])
        qs = [1 + (i % 11) for i in range(n_unique)]
        lgwin, desc = 22, "mix pool, %d B windows, quality 1+(i mod 11), lgwin 22" % size
        comp = compress_many(originals, qs, lgwin, threads)
    else:
        raise ValueError(config)
    return comp, originals, desc


def pack(blobs):
    """-> (bytes u8[total], off u64[n+1])"""
    sizes = np.fromiter((len(b) for b in blobs), dtype=np.uint64, count=len(blobs))
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    np.cumsum(sizes, out=off[1:])
    return np.frombuffer(b"".join(blobs), dtype=np.uint8), off
