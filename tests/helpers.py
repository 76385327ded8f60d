"""Shared test plumbing: ctypes bindings of the CPU oracle (oracle/liboracle.so) and of the
host-simulation build of the kernel logic (tests/hostsim), plus stream mutation helpers.
Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("%s failed:\n%s" % (" ".join(cmd), r.stdout))


class OracleReturnInfo(ctypes.Structure):
    _fields_ = [("decoded_size", ctypes.c_size_t), ("error", ctypes.c_char * 256), ("result", ctypes.c_int), ("error_code", ctypes.c_int)]


class OracleHuffmanCode(ctypes.Structure):
    _fields_ = [("value", ctypes.c_uint16), ("bits", ctypes.c_uint8)]


class OracleBitReader(ctypes.Structure):
    _fields_ = [("val_", ctypes.c_uint64), ("bit_pos_", ctypes.c_uint32), ("next_in", ctypes.c_uint32), ("avail_in", ctypes.c_uint32)]


class Oracle:
    """oracle/liboracle.so: CPU restatement of the reference decoder (the parity checker)."""

    def __init__(self):
        _run(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        L = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L.oracle_brotli_decode.restype = OracleReturnInfo
        L.oracle_brotli_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.oracle_brotli_decode_ex.restype = OracleReturnInfo
        L.oracle_brotli_decode_ex.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                              ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_brotli_decode_batch.restype = ctypes.c_int
        L.oracle_brotli_decode_batch.argtypes = [ctypes.c_size_t] + [ctypes.c_void_p] * 6 + [ctypes.c_int]
        L.oracle_error_string.restype = ctypes.c_char_p
        L.oracle_error_string.argtypes = [ctypes.c_int]
        L.oracle_stream_create.restype = ctypes.c_void_p
        L.oracle_stream_create.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_stream_decompress.restype = ctypes.c_int
        L.oracle_stream_decompress.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p,
                                               ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p,
                                               ctypes.POINTER(ctypes.c_size_t)]
        L.oracle_stream_error_code.restype = ctypes.c_int
        L.oracle_stream_error_code.argtypes = [ctypes.c_void_p]
        L.oracle_stream_destroy.restype = None
        L.oracle_stream_destroy.argtypes = [ctypes.c_void_p]
        self.lib = L

    def stream(self, large_window=False, custom_dict=None):
        """A resumable decoder state: the reference's BrotliDecompressStream call by call (OracleStream)."""
        return OracleStream(self.lib, large_window, custom_dict)

    def decode(self, data, capacity, large_window=True, custom_dict=None):
        data = bytes(data)
        buf = ctypes.create_string_buffer(max(int(capacity), 1))
        if custom_dict is None and large_window:
            r = self.lib.oracle_brotli_decode(data, len(data), buf, int(capacity))
        else:
            cd = bytes(custom_dict) if custom_dict else None
            r = self.lib.oracle_brotli_decode_ex(data, len(data), buf, int(capacity), 1 if large_window else 0, cd, len(cd) if cd else 0)
        return r.result, r.error_code, buf.raw[:r.decoded_size]

    def decode_batch(self, in_bytes, in_off, out_bytes, out_off, out_len, codes, threads=1):
        n = len(out_len)
        return self.lib.oracle_brotli_decode_batch(n, in_bytes.ctypes.data, in_off.ctypes.data, out_bytes.ctypes.data, out_off.ctypes.data,
                                                   out_len.ctypes.data, codes.ctypes.data, int(threads))


class OracleStream:
    """One BrotliState of the oracle across many BrotliDecompressStream calls (src/decode.rs:2779-2790).
    call(data, out_cap) feeds `data` with an output buffer of out_cap bytes and returns
    (result, consumed, produced_bytes, total_out) exactly as the reference moves its arguments."""

    def __init__(self, lib, large_window=False, custom_dict=None):
        self.lib = lib
        cd = bytes(custom_dict) if custom_dict else None
        self.h = lib.oracle_stream_create(1 if large_window else 0, cd, len(cd) if cd else 0)
        self.total_out = ctypes.c_size_t(0)

    def call(self, data, out_cap):
        data = bytes(data)
        buf = ctypes.create_string_buffer(max(int(out_cap), 1))
        ain, ioff = ctypes.c_size_t(len(data)), ctypes.c_size_t(0)
        aout, ooff = ctypes.c_size_t(int(out_cap)), ctypes.c_size_t(0)
        r = self.lib.oracle_stream_decompress(self.h, ctypes.byref(ain), ctypes.byref(ioff), data, ctypes.byref(aout), ctypes.byref(ooff), buf,
                                              ctypes.byref(self.total_out))
        assert ioff.value + ain.value == len(data) and ooff.value + aout.value == int(out_cap)
        return r, ioff.value, buf.raw[:ooff.value], self.total_out.value

    def error_code(self):
        return self.lib.oracle_stream_error_code(self.h)

    def close(self):
        if self.h:
            self.lib.oracle_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def drive_stream(call, comp, in_chunk, out_chunk, max_calls=10_000_000):
    """The reference's own streaming driver (decompress_internal, src/bin/integration_tests.rs:122-216) over any
    decoder exposing call(data, out_cap) -> (result, consumed, produced, total_out): new input (<= in_chunk bytes)
    only after NeedsMoreInput, a fresh out_chunk-byte output buffer for every call.
    Returns (final result, output bytes, trace) with trace = [(result, consumed, n_produced, total_out), ...]."""
    comp = bytes(comp)
    pos = 0
    pending = b""
    result = 2
    out = bytearray()
    trace = []
    for _ in range(max_calls):
        if result == 2:
            if pos >= len(comp):
                return 2, bytes(out), trace  # "Read EOF"
            pending = comp[pos:pos + in_chunk]
            pos += len(pending)
        elif result != 3:
            break
        result, consumed, produced, total = call(pending, out_chunk)
        trace.append((result, consumed, len(produced), total))
        pending = pending[consumed:]
        out += produced
    return result, bytes(out), trace


class HostSim:
    """tests/hostsim: csrc/brotli_decode_core.cuh compiled for the host with warp width 1."""

    def __init__(self):
        d = os.path.join(ROOT, "tests", "hostsim")
        so = os.path.join(d, "libhostsim.so")
        srcs = [os.path.join(d, "hostsim.cpp"), os.path.join(d, "hostsim_lane.cpp"),
                os.path.join(ROOT, "rust-brotli-decompressor_b200", "csrc", "brotli_decode_core.cuh"),
                os.path.join(ROOT, "rust-brotli-decompressor_b200", "csrc", "brotli_decode_lane.cuh"),
                os.path.join(ROOT, "tables", "brotli_tables.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            _run(["g++", "-O2", "-g", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", so, srcs[0], srcs[1], os.path.join(ROOT, "tables", "brotli_dictionary.c"),
                  '-DBROTLI_DICT_PATH="%s"' % os.path.join(ROOT, "tables", "brotli_dictionary.bin")])
        L = ctypes.CDLL(so)
        L.hostsim_decode.restype = ctypes.c_int
        L.hostsim_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_uint64)]
        L.hostsim_decode_dict.restype = ctypes.c_int
        L.hostsim_decode_dict.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                          ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
        L.hostsim_lane_decode_dict.restype = ctypes.c_int
        L.hostsim_lane_decode_dict.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32,
                                               ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.c_size_t]
        L.hostsim_stream_create.restype = ctypes.c_void_p
        L.hostsim_stream_create.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
        L.hostsim_stream_destroy.restype = None
        L.hostsim_stream_destroy.argtypes = [ctypes.c_void_p]
        L.hostsim_stream_calls.restype = ctypes.c_int
        L.hostsim_stream_calls.argtypes = [ctypes.c_size_t] + [ctypes.c_void_p] * 10
        L.hostsim_stream_stats.restype = None
        L.hostsim_stream_stats.argtypes = [ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_uint64)]
        L.hostsim_lane_set_latency_config.restype = None
        L.hostsim_lane_set_latency_config.argtypes = [ctypes.c_int]
        L.hostsim_lane_decode.restype = ctypes.c_int
        L.hostsim_lane_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32,
                                          ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        self.lib = L

    LANE_BAIL = 1000

    def lane_decode(self, data, capacity, table_entries=178, misalign=0, custom_dict=None):
        """Lane-per-stream (optimistic) path -> (code, bytes, input bytes used); code 1 = decoded, LANE_BAIL = the
        path gave the stream up (the exact kernel decodes it on the device; bytes are then meaningless).
        table_entries = u16 entries of the lane's shared-memory slot; misalign = output address modulo 32 (the
        path writes whole 32-byte sectors; the head and tail of an unaligned region are written byte-wise)."""
        data = bytes(data)
        buf = ctypes.create_string_buffer(max(int(capacity), 1) + 136)
        addr = ctypes.addressof(buf)
        addr += (-addr) % 32 + 32 + misalign
        off = addr - ctypes.addressof(buf)
        n = ctypes.c_uint64(0)
        used = ctypes.c_uint64(0)
        cd = bytes(custom_dict) if custom_dict else b""
        code = self.lib.hostsim_lane_decode_dict(data, len(data), addr, int(capacity), int(table_entries), ctypes.byref(n), ctypes.byref(used),
                                                 cd, len(cd))
        raw = buf.raw
        assert raw[:off] == bytes(off) and raw[off + int(capacity):] == bytes(len(raw) - off - int(capacity)), "wrote outside the region"
        return code, raw[off:off + n.value], used.value

    def decode(self, data, capacity, large_window=True, custom_dict=None):
        """-> (code, bytes): code is the BrotliDecoderErrorCode, bytes the reference's decoded_size prefix."""
        data = bytes(data)
        buf = ctypes.create_string_buffer(max(int(capacity), 1))
        n = ctypes.c_uint64(0)
        cd = bytes(custom_dict) if custom_dict else b""
        code = self.lib.hostsim_decode_dict(data, len(data), buf, int(capacity), 1 if large_window else 0, cd, len(cd), ctypes.byref(n))
        return code, buf.raw[:n.value]


class HostSimStream:
    """A streaming session of the product's host logic (csrc/brotli_b200_session.h) over the host build of the exact
    kernel: call(data, out_cap) has BrotliDecoderDecompressStream's semantics and returns
    (result, consumed, produced_bytes, total_out) like OracleStream.call."""

    def __init__(self, hostsim, large_window=False, custom_dict=None):
        self.lib = hostsim.lib
        cd = bytes(custom_dict) if custom_dict else None
        self.h = self.lib.hostsim_stream_create(1 if large_window else 0, cd, len(cd) if cd else 0)
        self.code = 0

    def call(self, data, out_cap):
        (r,) = stream_calls(self.lib, [self], [data], [out_cap])
        return r

    def error_code(self):
        return self.code

    def close(self):
        if self.h:
            self.lib.hostsim_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def stream_calls(lib, streams, datas, out_caps):
    """One batched call for several HostSimStream sessions (one simulated launch)."""
    n = len(streams)
    datas = [bytes(d) for d in datas]
    handles = (ctypes.c_void_p * n)(*[s.h for s in streams])
    ins = (ctypes.c_char_p * n)(*datas)
    in_size = (ctypes.c_size_t * n)(*[len(d) for d in datas])
    bufs = [ctypes.create_string_buffer(max(int(c), 1)) for c in out_caps]
    outs = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in bufs])
    caps = (ctypes.c_size_t * n)(*[int(c) for c in out_caps])
    consumed = (ctypes.c_size_t * n)(); produced = (ctypes.c_size_t * n)(); total = (ctypes.c_size_t * n)()
    results = (ctypes.c_int * n)(); codes = (ctypes.c_int * n)()
    rc = lib.hostsim_stream_calls(n, handles, ins, in_size, outs, caps, consumed, produced, total, results, codes)
    assert rc == 0
    out = []
    for i in range(n):
        streams[i].code = codes[i]
        out.append((results[i], consumed[i], bufs[i].raw[:produced[i]], total[i]))
    return out


def hostsim_stream_stats(hostsim):
    a, b, c = ctypes.c_size_t(0), ctypes.c_size_t(0), ctypes.c_uint64(0)
    hostsim.lib.hostsim_stream_stats(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return {"live_bytes": a.value, "peak_bytes": b.value, "launches": c.value}


def result_of(code):
    """BrotliDecoderErrorCode -> BrotliResult (src/decode.rs:33-40)."""
    return code if code in (1, 2, 3) else 0


def mutations(stream, rng, count):
    """Truncations and bit/byte corruptions of a valid stream (deterministic for a seeded rng)."""
    out = []
    s = bytes(stream)
    for _ in range(count):
        kind = int(rng.integers(0, 4))
        if kind == 0 and len(s) > 1:
            out.append(s[: int(rng.integers(0, len(s)))])
        elif kind == 1:
            b = bytearray(s); i = int(rng.integers(0, len(b))); b[i] ^= 1 << int(rng.integers(0, 8)); out.append(bytes(b))
        elif kind == 2:
            b = bytearray(s); i = int(rng.integers(0, len(b))); b[i] = int(rng.integers(0, 256)); out.append(bytes(b))
        else:
            b = bytearray(s)
            for _ in range(3):
                i = int(rng.integers(0, len(b))); b[i] ^= int(rng.integers(1, 256))
            out.append(bytes(b))
    return out


def golden_manifest():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))


def golden_fixture(name):
    return open(os.path.join(ROOT, "tests", "golden", "fixtures", name), "rb").read()


def inline_vectors():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "inline_vectors.json")))
