"""brotli_b200 -- B200-native batched Brotli decoder (Python binding of libbrotli_b200.so).

The product is the shared library built from ``csrc/`` (hand-written sm_100a CUDA kernels + C++
host runtime, C ABI in ``include/brotli_b200/decode.h``).  This module is only the ctypes binding
used by ``tests/`` and ``bench.py`` plus thin Python mirrors of the reference crate's surface:

* ``BrotliDecoderDecompress`` / ``brotli_decode``   -> src/ffi/mod.rs:262-292, src/lib.rs:446-468
* ``BrotliDecompressStream`` via ``DecoderState``   -> src/decode.rs:2779-2790, src/ffi/mod.rs:389-463
* ``Decompressor(reader, buffer_size).read()``      -> src/reader.rs:91-130,299-350
* ``decompress_batch*``                             -> the batch extension (one warp per stream)

There is no CPU decode path: if the library is missing or no CUDA device is usable every call
raises.  Import with ``importlib.import_module("rust-brotli-decompressor_b200")``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BROTLI_B200_LIB selects another build of the same library (tuning experiments: warps per CTA etc.)
LIB_PATH = os.environ.get("BROTLI_B200_LIB") or os.path.join(_HERE, "libbrotli_b200.so")

RESULT_ERROR, RESULT_SUCCESS, RESULT_NEEDS_MORE_INPUT, RESULT_NEEDS_MORE_OUTPUT = 0, 1, 2, 3

# every symbol include/brotli_b200/decode.h declares
EXPORTED_SYMBOLS = [
    "BrotliDecoderDecompress", "BrotliDecoderDecompressWithReturnInfo", "BrotliDecoderDecompressPrealloc",
    "BrotliDecoderCreateInstance", "BrotliDecoderSetParameter", "BrotliDecoderDestroyInstance",
    "BrotliDecoderDecompressStream", "BrotliDecoderDecompressStreaming", "BrotliDecoderHasMoreOutput",
    "BrotliDecoderTakeOutput", "BrotliDecoderIsUsed", "BrotliDecoderIsFinished", "BrotliDecoderGetErrorCode",
    "BrotliDecoderGetErrorString", "BrotliDecoderErrorString", "BrotliDecoderVersion", "BrotliDecoderMallocU8",
    "BrotliDecoderFreeU8", "BrotliDecoderMallocUsize", "BrotliDecoderFreeUsize", "BrotliB200DecompressBatchDevice",
    "BrotliB200DecompressBatchPacked", "BrotliB200DecompressBatch", "BrotliB200ChecksumBatchDevice",
    "BrotliB200DecompressWithDictionary", "BrotliB200DecompressBatchPackedWithDictionary", "BrotliB200DecoderSetCustomDictionary",
    "BrotliB200DecoderDecompressStreamBatch", "BrotliB200SetTuning",
    "BrotliB200KernelLaunchCount", "BrotliB200LastKernelMs", "BrotliB200KernelTimes", "BrotliB200LastError", "BrotliB200ResidentWarps", "BrotliB200LastLaneGeometry",
    "BrotliB200Shutdown",
]


class BrotliDecoderReturnInfo(ctypes.Structure):
    """c/brotli/decode.h:127-132 == src/lib.rs:336-342"""
    _fields_ = [("decoded_size", ctypes.c_size_t), ("error", ctypes.c_char * 256), ("result", ctypes.c_int),
                ("code", ctypes.c_int)]


class BrotliB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded libbrotli_b200.so; raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BrotliB200Error("libbrotli_b200.so is not built (%s); run `make -C rust-brotli-decompressor_b200/csrc`. "
                              "There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p
    L.BrotliDecoderDecompress.restype = ctypes.c_int
    L.BrotliDecoderDecompress.argtypes = [sz, vp, ctypes.POINTER(sz), vp]
    L.BrotliDecoderDecompressWithReturnInfo.restype = BrotliDecoderReturnInfo
    L.BrotliDecoderDecompressWithReturnInfo.argtypes = [sz, vp, sz, vp]
    L.BrotliDecoderDecompressPrealloc.restype = BrotliDecoderReturnInfo
    L.BrotliDecoderDecompressPrealloc.argtypes = [sz, vp, sz, vp, sz, vp, sz, vp, sz, vp]
    L.BrotliDecoderCreateInstance.restype = vp
    L.BrotliDecoderCreateInstance.argtypes = [vp, vp, vp]
    L.BrotliDecoderSetParameter.restype = ctypes.c_int
    L.BrotliDecoderSetParameter.argtypes = [vp, ctypes.c_int, ctypes.c_uint32]
    L.BrotliDecoderDestroyInstance.restype = None
    L.BrotliDecoderDestroyInstance.argtypes = [vp]
    L.BrotliDecoderDecompressStream.restype = ctypes.c_int
    L.BrotliDecoderDecompressStream.argtypes = [vp, ctypes.POINTER(sz), ctypes.POINTER(vp), ctypes.POINTER(sz),
                                                ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.BrotliDecoderDecompressStreaming.restype = ctypes.c_int
    L.BrotliDecoderDecompressStreaming.argtypes = [vp, ctypes.POINTER(sz), vp, ctypes.POINTER(sz), vp]
    for name in ("BrotliDecoderHasMoreOutput", "BrotliDecoderIsUsed", "BrotliDecoderIsFinished", "BrotliDecoderGetErrorCode"):
        getattr(L, name).restype = ctypes.c_int
        getattr(L, name).argtypes = [vp]
    L.BrotliDecoderTakeOutput.restype = vp
    L.BrotliDecoderTakeOutput.argtypes = [vp, ctypes.POINTER(sz)]
    L.BrotliDecoderGetErrorString.restype = u8p
    L.BrotliDecoderGetErrorString.argtypes = [vp]
    L.BrotliDecoderErrorString.restype = u8p
    L.BrotliDecoderErrorString.argtypes = [ctypes.c_int]
    L.BrotliDecoderVersion.restype = ctypes.c_uint32
    L.BrotliDecoderVersion.argtypes = []
    L.BrotliB200DecompressBatchDevice.restype = ctypes.c_int
    L.BrotliB200DecompressBatchDevice.argtypes = [sz, vp, vp, vp, vp, vp, vp, vp]
    L.BrotliB200DecompressBatchPacked.restype = ctypes.c_int
    L.BrotliB200DecompressBatchPacked.argtypes = [sz, vp, vp, vp, vp, vp, vp]
    L.BrotliB200DecompressBatch.restype = ctypes.c_int
    L.BrotliB200DecompressBatch.argtypes = [sz, vp, vp, vp, vp, vp, vp]
    L.BrotliB200DecompressWithDictionary.restype = BrotliDecoderReturnInfo
    L.BrotliB200DecompressWithDictionary.argtypes = [sz, vp, sz, vp, vp, sz]
    L.BrotliB200DecompressBatchPackedWithDictionary.restype = ctypes.c_int
    L.BrotliB200DecompressBatchPackedWithDictionary.argtypes = [sz, vp, vp, vp, vp, vp, vp, vp, sz]
    L.BrotliB200DecoderSetCustomDictionary.restype = ctypes.c_int
    L.BrotliB200DecoderSetCustomDictionary.argtypes = [vp, vp, sz]
    L.BrotliB200DecoderDecompressStreamBatch.restype = ctypes.c_int
    L.BrotliB200DecoderDecompressStreamBatch.argtypes = [sz] + [vp] * 7
    L.BrotliB200SetTuning.restype = ctypes.c_int
    L.BrotliB200SetTuning.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
    L.BrotliB200ChecksumBatchDevice.restype = ctypes.c_int
    L.BrotliB200ChecksumBatchDevice.argtypes = [sz, vp, vp, vp, vp, vp]
    L.BrotliB200KernelLaunchCount.restype = ctypes.c_uint64
    L.BrotliB200LastKernelMs.restype = ctypes.c_double
    L.BrotliB200KernelTimes.restype = ctypes.c_int
    L.BrotliB200KernelTimes.argtypes = [vp, vp, vp, vp, ctypes.c_int]
    L.BrotliB200LastError.restype = u8p
    L.BrotliB200ResidentWarps.restype = ctypes.c_int
    L.BrotliB200LastLaneGeometry.restype = ctypes.c_int
    L.BrotliB200Shutdown.restype = None
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise BrotliB200Error("%s failed (%d): %s" % (what, rc, lib().BrotliB200LastError().decode("utf-8", "replace")))


def error_string(code):
    """BrotliDecoderErrorString, c/brotli/decode.h:377"""
    return lib().BrotliDecoderErrorString(int(code)).decode()


def kernel_times(reset=False):
    """-> dict(lane_ms, exact_ms, launches, bailed): summed device time of the two decode kernels over the decode
    calls since the last reset (CUDA events on the launching stream), and the bail count of the last call."""
    lane, exact = ctypes.c_double(0), ctypes.c_double(0)
    n, bailed = ctypes.c_uint32(0), ctypes.c_uint32(0)
    _check(lib().BrotliB200KernelTimes(ctypes.byref(lane), ctypes.byref(exact), ctypes.byref(n), ctypes.byref(bailed), 1 if reset else 0),
           "BrotliB200KernelTimes")
    return {"lane_ms": lane.value, "exact_ms": exact.value, "launches": n.value, "bailed": bailed.value}


def kernel_launch_count():
    return int(lib().BrotliB200KernelLaunchCount())


# ---------------------------------------------------------------------------------------------
# one-shot (batch of one)
# ---------------------------------------------------------------------------------------------
def brotli_decode(data, capacity):
    """``brotli_decode(input, output)`` (src/lib.rs:446-468) via BrotliDecoderDecompressWithReturnInfo.

    Returns ``(info, output_bytes)`` with ``info.result`` the BrotliResult, ``info.code`` the
    BrotliDecoderErrorCode and ``output_bytes`` the first ``info.decoded_size`` bytes."""
    data = bytes(data)
    buf = ctypes.create_string_buffer(max(int(capacity), 1))
    info = lib().BrotliDecoderDecompressWithReturnInfo(len(data), data, int(capacity), buf)
    return info, buf.raw[:info.decoded_size]


def brotli_decode_custom_dict(data, capacity, custom_dictionary):
    """One stream decoded with a custom LZ77 dictionary, the one-shot form of ``BrotliDecompressCustomDict``
    (src/lib.rs:105-131; ``BrotliState::new_with_custom_dictionary``, src/state.rs:400-411).  Returns
    ``(info, output_bytes)`` like :func:`brotli_decode`."""
    data, cd = bytes(data), bytes(custom_dictionary)
    buf = ctypes.create_string_buffer(max(int(capacity), 1))
    info = lib().BrotliB200DecompressWithDictionary(len(data), data, int(capacity), buf, cd, len(cd))
    return info, buf.raw[:info.decoded_size]


def BrotliDecoderDecompress(data, capacity):
    """C one-shot (src/ffi/mod.rs:262-292): returns ``(result, output_bytes)``; result is 1 or 0."""
    data = bytes(data)
    buf = ctypes.create_string_buffer(max(int(capacity), 1))
    size = ctypes.c_size_t(int(capacity))
    r = lib().BrotliDecoderDecompress(len(data), data, ctypes.byref(size), buf)
    return r, buf.raw[:size.value]


# ---------------------------------------------------------------------------------------------
# streaming surface
# ---------------------------------------------------------------------------------------------
class DecoderState:
    """BrotliDecoderState driven through BrotliDecoderDecompressStream (src/ffi/mod.rs:389-463)."""

    def __init__(self, large_window=False, custom_dict=None):
        self._s = lib().BrotliDecoderCreateInstance(None, None, None)
        if not self._s:
            raise BrotliB200Error("BrotliDecoderCreateInstance failed")
        if large_window:
            lib().BrotliDecoderSetParameter(self._s, 1, 1)
        if custom_dict:  # BrotliState::new_with_custom_dictionary, src/state.rs:400-411
            cd = bytes(custom_dict)
            if not lib().BrotliB200DecoderSetCustomDictionary(self._s, cd, len(cd)):
                raise BrotliB200Error("BrotliB200DecoderSetCustomDictionary failed")

    def close(self):
        if self._s:
            lib().BrotliDecoderDestroyInstance(self._s)
            self._s = None

    __del__ = close

    def decompress_stream(self, data, out_capacity):
        """One BrotliDecoderDecompressStream call.  Returns (result, consumed, produced_bytes)."""
        data = bytes(data)
        inbuf = ctypes.create_string_buffer(data, max(len(data), 1))
        outbuf = ctypes.create_string_buffer(max(int(out_capacity), 1))
        avail_in, avail_out = ctypes.c_size_t(len(data)), ctypes.c_size_t(int(out_capacity))
        next_in = ctypes.c_void_p(ctypes.addressof(inbuf))
        next_out = ctypes.c_void_p(ctypes.addressof(outbuf))
        total = ctypes.c_size_t(0)
        r = lib().BrotliDecoderDecompressStream(self._s, ctypes.byref(avail_in), ctypes.byref(next_in), ctypes.byref(avail_out),
                                                ctypes.byref(next_out), ctypes.byref(total))
        produced = int(out_capacity) - avail_out.value
        self.total_out = total.value
        return r, len(data) - avail_in.value, outbuf.raw[:produced]

    def call(self, data, out_capacity):
        """decompress_stream with the running total: (result, consumed, produced_bytes, total_out)."""
        r, used, out = self.decompress_stream(data, out_capacity)
        return r, used, out, self.total_out

    def is_finished(self):
        return bool(lib().BrotliDecoderIsFinished(self._s))

    def is_used(self):
        return bool(lib().BrotliDecoderIsUsed(self._s))

    def has_more_output(self):
        return bool(lib().BrotliDecoderHasMoreOutput(self._s))

    def error_code(self):
        return int(lib().BrotliDecoderGetErrorCode(self._s))

    def error_string(self):
        return lib().BrotliDecoderGetErrorString(self._s).decode()


def last_lane_geometry():
    """BrotliB200LastLaneGeometry: warps per SM of the lane kernel that decoded the most recent batch (0: it was not used)."""
    return int(lib().BrotliB200LastLaneGeometry())


def set_tuning(name, value):
    """BrotliB200SetTuning: "lane_min_streams", "small_geometry", "sort_streams", "lane_slot_bytes"."""
    return bool(lib().BrotliB200SetTuning(name.encode(), int(value)))


def decompress_stream_batch(states, datas, out_capacities):
    """One BrotliDecoderDecompressStream call on each of ``states`` (DecoderState objects), served by ONE decode launch
    (BrotliB200DecoderDecompressStreamBatch).  Returns [(result, consumed, produced_bytes, total_out), ...]."""
    n = len(states)
    if n == 0:
        return []
    datas = [bytes(d) for d in datas]
    inbufs = [ctypes.create_string_buffer(d, max(len(d), 1)) for d in datas]
    outbufs = [ctypes.create_string_buffer(max(int(c), 1)) for c in out_capacities]
    handles = (ctypes.c_void_p * n)(*[s._s for s in states])
    avail_in = (ctypes.c_size_t * n)(*[len(d) for d in datas])
    avail_out = (ctypes.c_size_t * n)(*[int(c) for c in out_capacities])
    next_in = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in inbufs])
    next_out = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in outbufs])
    total = (ctypes.c_size_t * n)()
    results = (ctypes.c_int * n)()
    _check(lib().BrotliB200DecoderDecompressStreamBatch(n, handles, avail_in, next_in, avail_out, next_out, total, results),
           "BrotliB200DecoderDecompressStreamBatch")
    out = []
    for i in range(n):
        produced = int(out_capacities[i]) - avail_out[i]
        states[i].total_out = total[i]
        out.append((results[i], len(datas[i]) - avail_in[i], outbufs[i].raw[:produced], total[i]))
    return out


class Decompressor:
    """``Decompressor<R: Read>`` (src/reader.rs:91-130): wraps a reader of compressed bytes.

    ``read(n)`` returns up to n decompressed bytes, b"" at the end of the stream, and raises
    ``ValueError("Invalid Data")`` for a corrupt stream or, on the read after the end, when
    bytes remain after the final metablock (src/reader.rs:299-350)."""

    def __init__(self, reader, buffer_size=4096, custom_dict=None):
        """``custom_dict``: ``Decompressor::new_with_custom_dict(r, buffer_size, dict)`` (src/reader.rs:105)."""
        self._r = reader
        self._bufsize = max(int(buffer_size), 1)
        # the reference builds this state with BrotliState::new_with_custom_dictionary (src/reader.rs:226), which accepts
        # large-window streams (src/state.rs:400-411)
        self._state = DecoderState(large_window=True, custom_dict=custom_dict)
        self._pending = b""
        self._eof = False
        self._done = False

    def read(self, n=-1):
        if n is None or n < 0:
            chunks = []
            while True:
                c = self.read(65536)
                if not c:
                    return b"".join(chunks)
                chunks.append(c)
        if n == 0:
            return b""
        while True:
            if self._done:
                # src/reader.rs:335-344: leftover input after the end of the stream is an error
                if self._pending or (not self._eof and self._fill()):
                    raise ValueError("Invalid Data")
                return b""
            r, used, out = self._state.decompress_stream(self._pending, n)
            self._pending = self._pending[used:]
            if r == RESULT_ERROR:
                raise ValueError("Invalid Data")
            if r == RESULT_SUCCESS:
                self._done = True
            if out:
                return out
            if r == RESULT_NEEDS_MORE_INPUT:
                if not self._fill():
                    raise ValueError("Invalid Data")  # truncated: UnexpectedEof in the reference (src/reader.rs:318-322)
            elif r == RESULT_SUCCESS:
                continue

    def _fill(self):
        chunk = self._r.read(self._bufsize)
        if not chunk:
            self._eof = True
            return False
        self._pending += chunk
        return True


# ---------------------------------------------------------------------------------------------
# batch extension
# ---------------------------------------------------------------------------------------------
def decompress_batch_packed(in_bytes, in_off, out_bytes, out_off, out_len, codes):
    """Host-resident packed batch (numpy arrays or torch CPU tensors exposing ``ctypes``/``data_ptr``).

    in_bytes u8[...], in_off u64[n+1], out_bytes u8[...], out_off u64[n+1], out_len u64[n], codes i32[n]."""
    n = len(out_len)
    _check(lib().BrotliB200DecompressBatchPacked(n, _ptr(in_bytes), _ptr(in_off), _ptr(out_bytes), _ptr(out_off), _ptr(out_len),
                                                 _ptr(codes)), "BrotliB200DecompressBatchPacked")


def decompress_batch_packed_custom_dict(in_bytes, in_off, out_bytes, out_off, out_len, codes, custom_dictionary):
    """:func:`decompress_batch_packed` with one custom LZ77 dictionary shared by every stream of the batch."""
    n = len(out_len)
    cd = bytes(custom_dictionary)
    _check(lib().BrotliB200DecompressBatchPackedWithDictionary(n, _ptr(in_bytes), _ptr(in_off), _ptr(out_bytes), _ptr(out_off),
                                                               _ptr(out_len), _ptr(codes), cd, len(cd)),
           "BrotliB200DecompressBatchPackedWithDictionary")


def decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_out_len, d_codes, stream=None):
    """Device-resident packed batch; arguments are torch CUDA tensors (u8, i64/u64, u8, i64/u64, i64/u64, i32).
    Asynchronous on ``stream`` (a torch.cuda.Stream; default: the current stream)."""
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    _check(lib().BrotliB200DecompressBatchDevice(int(n), d_in.data_ptr(), d_in_off.data_ptr(), d_out.data_ptr(),
                                                 d_out_off.data_ptr(), d_out_len.data_ptr(), d_codes.data_ptr(),
                                                 ctypes.c_void_p(s.cuda_stream)), "BrotliB200DecompressBatchDevice")


def checksum_batch_device(n, d_bytes, d_off, d_len, d_sums, stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    _check(lib().BrotliB200ChecksumBatchDevice(int(n), d_bytes.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_sums.data_ptr(),
                                               ctypes.c_void_p(s.cuda_stream)), "BrotliB200ChecksumBatchDevice")


def decompress_batch(streams, capacities):
    """Scattered host batch == n x BrotliDecoderDecompress.  Returns a list of (result, code, bytes)."""
    n = len(streams)
    if n == 0:
        return []
    streams = [bytes(s) for s in streams]
    ins = (ctypes.c_char_p * n)(*streams)
    in_size = (ctypes.c_size_t * n)(*[len(s) for s in streams])
    bufs = [ctypes.create_string_buffer(max(int(c), 1)) for c in capacities]
    outs = (ctypes.c_void_p * n)(*[ctypes.addressof(b) for b in bufs])
    out_size = (ctypes.c_size_t * n)(*[int(c) for c in capacities])
    results = (ctypes.c_int * n)()
    codes = (ctypes.c_int * n)()
    _check(lib().BrotliB200DecompressBatch(n, ins, in_size, outs, out_size, results, codes), "BrotliB200DecompressBatch")
    return [(results[i], codes[i], bufs[i].raw[:out_size[i]]) for i in range(n)]


def _ptr(a):
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    return ctypes.c_void_p(a.ctypes.data)


def checksum_reference(data):
    """numpy restatement of brotli_checksum_batch_kernel for one stream's bytes (used by the checkers)."""
    import numpy as np
    b = np.frombuffer(bytes(data), dtype=np.uint8).astype(np.uint64)
    j = np.arange(len(b), dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (b + np.uint64(1)) * (np.uint64(0x9E3779B97F4A7C15) + np.uint64(2) * j)
        x ^= x >> np.uint64(29)
        acc = np.sum(x * np.uint64(0xBF58476D1CE4E5B9), dtype=np.uint64)
        return int(acc ^ (np.uint64(len(b)) * np.uint64(0x94D049BB133111EB)))
