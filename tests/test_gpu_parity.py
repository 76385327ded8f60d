"""Parity tests proper: the CUDA path, called through the C ABI of libbrotli_b200.so, against the
CPU oracle on the same inputs -- bit-exact bytes, same decoded_size, same BrotliDecoderErrorCode.
Run on the B200 box:  python -m pytest tests -m gpu"""
import hashlib
import io

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

MAN = helpers.golden_manifest()
SMALL = [n for n, e in MAN.items() if "original_sha256" in e]


def oracle_batch(oracle, streams, caps):
    return [oracle.decode(s, c) for s, c in zip(streams, caps)]


def check_batch(pkg, oracle, streams, caps, roomy=True):
    """Decode the batch on the GPU (scattered host API) and compare every stream with the oracle."""
    got = pkg.decompress_batch(streams, caps)
    n_ok = 0
    for i, ((gres, gcode, gout), s, c) in enumerate(zip(got, streams, caps)):
        ores, ocode, oout = oracle.decode(s, c)
        # (roomy is kept for the call sites; corrupt-AND-too-small streams get the reference's code as well: the host
        # decodes them again with the capacity as a budget, csrc/brotli_b200_host.cpp redo_needs_more_output)
        assert gcode == ocode, (i, gcode, ocode, len(s), c)
        assert gres == (1 if ocode == 1 else 0)
        assert gout == oout, (i, ocode, len(gout), len(oout))
        n_ok += ocode == 1
    return n_ok


def test_fixtures_one_shot(gpu_lib, pkg, oracle):
    for name in SMALL:
        e = MAN[name]
        info, out = pkg.brotli_decode(helpers.golden_fixture(name), e["original_size"])
        assert (info.result, info.code) == (1, 1), (name, info.code, info.error)
        assert len(out) == e["original_size"] and hashlib.sha256(out).hexdigest() == e["original_sha256"], name
        assert info.error == b"SUCCESS"


def test_fixtures_as_one_batch(gpu_lib, pkg, oracle):
    streams = [helpers.golden_fixture(n) for n in SMALL] + [helpers.golden_fixture("borked.compressed")]
    caps = [MAN[n]["original_size"] for n in SMALL] + [1 << 16]
    assert check_batch(pkg, oracle, streams, caps) == len(SMALL)


def test_c_one_shot_semantics(gpu_lib, pkg):
    """BrotliDecoderDecompress: SUCCESS only for a complete decode (src/ffi/mod.rs:279-283)."""
    data = helpers.golden_fixture("alice29.txt.compressed")
    size = MAN["alice29.txt.compressed"]["original_size"]
    r, out = pkg.BrotliDecoderDecompress(data, size)
    assert r == 1 and len(out) == size
    r, out = pkg.BrotliDecoderDecompress(data, size - 1)          # too small -> ERROR, decoded_size = written (App. D-3)
    assert r == 0 and len(out) == size - 1
    r, out = pkg.BrotliDecoderDecompress(data[:-10], size)        # truncated -> ERROR
    assert r == 0
    r, out = pkg.BrotliDecoderDecompress(data + b"trailing", size)  # trailing bytes ignored (App. D-2)
    assert r == 1 and len(out) == size
    info, out = pkg.brotli_decode(helpers.golden_fixture("borked.compressed"), 1 << 16)
    assert info.result == 0 and info.code < 0 and info.error.startswith(b"ERROR_")


def test_inline_vectors_and_one_byte_streams(gpu_lib, pkg, oracle):
    vec = helpers.inline_vectors()
    streams = [bytes.fromhex(v["input_hex"]) for v in vec["vectors"]] + [bytes([b]) for b in range(256)] + [b""]
    caps = [1 << 18] * len(vec["vectors"]) + [64] * 257
    check_batch(pkg, oracle, streams, caps)
    got = pkg.decompress_batch(streams[len(vec["vectors"]):-1], [64] * 256)
    assert sorted(b for b in range(256) if got[b][0] == 1) == vec["one_byte_ok"]
    info, _ = pkg.brotli_decode(bytes.fromhex(vec["vectors"][6]["input_hex"]), 4096)  # c/main.c:55-77
    assert info.code == -8 and info.error == b"ERROR_FORMAT_CONTEXT_MAP_REPEAT"


def test_large_window(gpu_lib, pkg, oracle):
    e = MAN["rnd_chunk.br"]
    info, out = pkg.brotli_decode(helpers.golden_fixture("rnd_chunk.br"), e["original_size"])
    assert (info.result, len(out)) == (1, e["original_size"])
    _, _, oout = oracle.decode(helpers.golden_fixture("rnd_chunk.br"), e["original_size"])
    assert out == oout


@pytest.mark.parametrize("config,n,size", [("C2", 256, 65536), ("C3", 1024, 4096), ("C5", 264, 65536), ("C4", 6, 3 << 20)])
def test_config_samples(gpu_lib, pkg, oracle, corpus, config, n, size):
    """Seeded samples of every BASELINE config, ragged capacities included."""
    comp, orig, _ = corpus.make_config(config, n, size=size)
    caps = [len(o) for o in orig]
    got = pkg.decompress_batch(comp, caps)
    for (res, code, out), o in zip(got, orig):
        assert (res, code) == (1, 1) and out == o
    caps2 = [len(o) + (i % 5) * 7 for i, o in enumerate(orig)]
    assert check_batch(pkg, oracle, comp, caps2) == n


def test_corrupt_truncated_and_small_buffers(gpu_lib, pkg, oracle, corpus):
    rng = np.random.default_rng(99)
    comp, orig, _ = corpus.make_config("C5", 66, size=20000)
    streams, caps = [], []
    for c, o in zip(comp, orig):
        for m in helpers.mutations(c, rng, 8):
            streams.append(m); caps.append(len(o) + 32)
    n_ok = check_batch(pkg, oracle, streams, caps)
    assert n_ok < len(streams)
    small = [max(len(o) - 1 - i % 9, 0) for i, o in enumerate(orig)]
    check_batch(pkg, oracle, comp, small)                       # valid streams, too little room: NEEDS_MORE_OUTPUT parity
    check_batch(pkg, oracle, streams, [max(c - 40, 0) for c in caps], roomy=False)


def test_empty_and_ragged_batches(gpu_lib, pkg, oracle):
    assert pkg.decompress_batch([], []) == []
    x = bytes.fromhex("0b00805803")
    streams = [b"", x, b"\x06", x[:2], x + b"junk", helpers.golden_fixture("zeros.compressed")]
    caps = [0, 1, 0, 10, 1, 262144]
    check_batch(pkg, oracle, streams, caps)


def test_device_batch_and_checksums(gpu_lib, pkg, oracle, corpus):
    """Device-resident packed API (the path bench.py times) + the on-device checksum used at full size."""
    import torch
    comp, orig, _ = corpus.make_config("C2", 512, size=65536)
    reps = 8
    blobs = comp * reps
    in_bytes, in_off = corpus.pack(blobs)
    out_off = np.arange(len(blobs) + 1, dtype=np.uint64) * np.uint64(65536)
    d_in = torch.from_numpy(in_bytes.copy()).cuda()
    d_in_off = torch.from_numpy(in_off.view(np.int64)).cuda()
    d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device="cuda")
    d_out_off = torch.from_numpy(out_off.view(np.int64)).cuda()
    d_len = torch.zeros(len(blobs), dtype=torch.int64, device="cuda")
    d_codes = torch.zeros(len(blobs), dtype=torch.int32, device="cuda")
    d_sums = torch.zeros(len(blobs), dtype=torch.int64, device="cuda")
    before = pkg.kernel_launch_count()
    pkg.decompress_batch_device(len(blobs), d_in, d_in_off, d_out, d_out_off, d_len, d_codes)
    pkg.checksum_batch_device(len(blobs), d_out, d_out_off, d_len, d_sums)
    torch.cuda.synchronize()
    assert pkg.kernel_launch_count() >= before + 2  # decode (lane pass + exact pass) + checksum
    assert bool((d_codes == 1).all()) and bool((d_len == 65536).all())
    out = d_out.cpu().numpy()
    assert out.tobytes() == b"".join(orig) * reps
    want = np.array([pkg.checksum_reference(o) for o in orig] * reps, dtype=np.uint64)
    assert np.array_equal(d_sums.cpu().numpy().view(np.uint64), want)


def test_host_packed_batch_pipeline(gpu_lib, pkg, corpus):
    """Host packed API with enough data to span several pipeline chunks."""
    comp, orig, _ = corpus.make_config("C2", 256, size=65536)
    reps = 40  # 10240 streams, 640 MiB of output
    blobs = comp * reps
    in_bytes, in_off = corpus.pack(blobs)
    out_off = np.arange(len(blobs) + 1, dtype=np.uint64) * np.uint64(65536)
    out = np.zeros(int(out_off[-1]), dtype=np.uint8)
    out_len = np.zeros(len(blobs), dtype=np.uint64)
    codes = np.zeros(len(blobs), dtype=np.int32)
    pkg.decompress_batch_packed(in_bytes, in_off, out, out_off, out_len, codes)
    assert (codes == 1).all() and (out_len == 65536).all()
    ref = np.frombuffer(b"".join(orig), dtype=np.uint8)
    assert np.array_equal(out.reshape(reps, -1), np.broadcast_to(ref, (reps, len(ref))))


def test_streaming_api_buffer_matrix(gpu_lib, pkg):
    """BrotliDecoderDecompressStream with the reference's buffer-size pairs
    (src/bin/integration_tests.rs:439-463: (65536,65536), (1,65536)-like small chunks, ...)."""
    for name, pairs in (("alice29.txt.compressed", [(65536, 65536), (4093, 70001), (50096, 1000)]),
                        ("quickfox_repeated.compressed", [(1, 65536), (3, 3000), (58, 1)]),
                        ("ukkonooa.compressed", [(1, 1), (3, 3), (1024, 1024)])):
        data = helpers.golden_fixture(name)
        e = MAN[name]
        for in_chunk, out_chunk in pairs:
            if name.startswith("quickfox_repeated") and out_chunk == 1:
                out_chunk = 4096
            st = pkg.DecoderState()
            got, pos, r = [], 0, 2
            for _ in range(10000000):
                chunk = data[pos:pos + in_chunk]
                r, used, out = st.decompress_stream(chunk, out_chunk)
                pos += used
                got.append(out)
                if r == 1 or r == 0:
                    break
                if r == 2:
                    assert pos < len(data), "decoder wants input past the end"
            assert r == 1 and st.is_finished() and st.is_used()
            assert hashlib.sha256(b"".join(got)).hexdigest() == e["original_sha256"], (name, in_chunk, out_chunk)
            st.close()


def test_streaming_errors_and_trailing_bytes(gpu_lib, pkg):
    st = pkg.DecoderState()
    bad = bytes.fromhex(helpers.inline_vectors()["vectors"][6]["input_hex"])
    r, used, out = st.decompress_stream(bad, 0)  # c/main.c:55-77
    assert r == 0 and st.error_code() == -8 and st.error_string() == "ERROR_FORMAT_CONTEXT_MAP_REPEAT"
    st.close()
    st = pkg.DecoderState()
    r, used, out = st.decompress_stream(bytes.fromhex("1f0700f827fe43840000"), 64)  # src/bin/ffi_stream_tests.rs:77
    assert r == 1 and out == b"\xff" * 8 and used == 10 and st.is_used()
    st.close()
    st = pkg.DecoderState()  # bytes after the end of the stream are left unconsumed (src/ffi/mod.rs:452-453)
    r, used, out = st.decompress_stream(bytes.fromhex("8f028068656c6c6f0a03") + b"trailing garbage", 64)
    assert r == 1 and out == b"hello\n" and used == 10
    st.close()
    st = pkg.DecoderState()  # large window needs the parameter on a streaming instance (src/ffi/mod.rs:127)
    r, used, out = st.decompress_stream(helpers.golden_fixture("rnd_chunk.br"), 4096)
    assert r == 0 and st.error_code() == -13
    st.close()


def test_decompressor_reader(gpu_lib, pkg):
    """Decompressor<R>::read (src/reader.rs:299-350) incl. the trailing-garbage rule (:353-389)."""
    name = "asyoulik.txt.compressed"
    d = pkg.Decompressor(io.BytesIO(helpers.golden_fixture(name)), 4096)
    out = d.read()
    assert hashlib.sha256(out).hexdigest() == MAN[name]["original_sha256"]
    assert d.read(10) == b""
    d = pkg.Decompressor(io.BytesIO(bytes.fromhex("8f028068656c6c6f0a03") + b"trailing garbage"), 4096)
    assert d.read(100) == b"hello\n"
    with pytest.raises(ValueError):
        d.read(100)
    d = pkg.Decompressor(io.BytesIO(helpers.golden_fixture(name)[:1000]), 256)
    with pytest.raises(ValueError):
        d.read()


def test_lane_kernel_takes_the_headline_streams(gpu_lib, pkg, corpus):
    """The lane-per-stream kernel must decode every well-formed q5 text stream itself (nothing handed to the
    exact kernel), and BrotliB200KernelTimes must see both kernels of the call."""
    import torch
    comp, orig, _ = corpus.make_config("headline", 320)
    in_bytes, in_off = corpus.pack(comp)
    out_off = np.arange(len(comp) + 1, dtype=np.uint64) * np.uint64(65536)
    d_in = torch.from_numpy(in_bytes.copy()).cuda()
    d_in_off = torch.from_numpy(in_off.view(np.int64)).cuda()
    d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device="cuda")
    d_out_off = torch.from_numpy(out_off.view(np.int64)).cuda()
    d_len = torch.zeros(len(comp), dtype=torch.int64, device="cuda")
    d_codes = torch.zeros(len(comp), dtype=torch.int32, device="cuda")
    pkg.kernel_times(reset=True)
    pkg.decompress_batch_device(len(comp), d_in, d_in_off, d_out, d_out_off, d_len, d_codes)
    t = pkg.kernel_times()
    assert t["launches"] == 1 and t["lane_ms"] > 0 and t["bailed"] == 0, t
    assert bool((d_codes == 1).all()) and d_out.cpu().numpy().tobytes() == b"".join(orig)


def test_bails_reach_the_exact_kernel(gpu_lib, pkg, oracle, corpus):
    """Streams the optimistic kernel gives up (too small a region, truncation, corruption, a large-window header) come
    back with the exact kernel's codes and sizes; the bail count says they took that path.  Uncompressed and metadata
    metablocks do NOT bail any more: the lane kernel copies / skips them itself."""
    comp, orig, _ = corpus.make_config("C2", 40)
    rnd = np.random.default_rng(5).integers(0, 256, size=3000, dtype=np.uint8).tobytes()
    bad = bytearray(comp[1]); bad[len(bad) // 3] ^= 0x41
    lw = helpers.golden_fixture("rnd_chunk.br")
    streams = list(comp) + [corpus.compress(rnd, 5), comp[0][: len(comp[0]) // 2], helpers.golden_fixture("x.compressed"), bytes(bad), lw]
    caps = [len(o) for o in orig] + [3000, 65536, 1, 65536, len(oracle.decode(lw, 1 << 24)[2])]
    caps[3] -= 1  # a valid stream with too little room
    pkg.kernel_times(reset=True)
    check_batch(pkg, oracle, streams, caps)
    bailed = pkg.kernel_times()["bailed"]
    assert 4 <= bailed <= 6, bailed   # too small, truncated, corrupt, large window (+ x.compressed into 1 byte); not the raw-bytes stream


def test_raw_and_metadata_metablocks_stay_on_the_lane_kernel(gpu_lib, pkg, oracle, corpus):
    """ISUNCOMPRESSED and metadata metablocks (src/decode.rs:1754-1806, :3031-3045) decoded by the lane kernel: random data at
    every quality (the encoder stores it raw), fixtures with raw / empty / metadata metablocks, unaligned regions."""
    rng = np.random.default_rng(6)
    streams, caps = [], []
    for size in (1, 17, 3000, 65536, 200000):
        for q in (0, 1, 5, 9, 11):
            data = rng.integers(0, 256, size=size, dtype=np.uint8).tobytes()
            streams.append(corpus.compress(data, q)); caps.append(size)
    for name in SMALL:
        streams.append(helpers.golden_fixture(name)); caps.append(MAN[name]["original_size"])
    pkg.kernel_times(reset=True)
    n_ok = check_batch(pkg, oracle, streams, caps)
    assert n_ok == len(streams) and pkg.kernel_times()["bailed"] <= 1  # (rnd_chunk.br is in SMALL only if it has an original)


@pytest.mark.parametrize("mode", ["exact_only", "lane_warps_8", "lane_warps_14", "lane_warps_24", "lane_warps_32", "small_slots"])
def test_other_kernel_configurations(gpu_lib, mode):
    """The same parity run with the lane kernel switched off (every stream through the exact warp-per-stream
    kernel) and with other lane-kernel geometries (smaller / larger shared-memory table slots per lane)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ)
    if mode == "exact_only":
        env["BROTLI_B200_LANE"] = "0"
    elif mode == "small_slots":  # 34 table entries per lane: most tree groups live in the arena (asynchronous root look-ups)
        env["BROTLI_B200_LANE_SLOT_BYTES"] = "84"
    else:
        env["BROTLI_B200_LANE_WARPS"] = mode.rsplit("_", 1)[1]
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(helpers.ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                        "-k", "config_samples or corrupt_truncated or fixtures_one_shot or empty_and_ragged"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]


def test_batch_size_routing(gpu_lib, pkg, corpus):
    """Default routing by batch size (DESIGN.md latency table): a small batch never touches the lane kernel, a large
    one does, and both give the originals."""
    import torch
    comp, orig, _ = corpus.make_config("C2", 96, size=65536)
    try:
        assert pkg.set_tuning("lane_min_streams", 6000)
        for n in (96, 6144):
            idx = np.arange(n) % 96
            blobs = [comp[i] for i in idx]
            in_bytes, in_off = corpus.pack(blobs)
            out_off = np.arange(n + 1, dtype=np.uint64) * 65536
            d_in = torch.from_numpy(in_bytes).cuda(); d_in_off = torch.from_numpy(in_off.view(np.int64)).cuda()
            d_out_off = torch.from_numpy(out_off.view(np.int64)).cuda()
            d_out = torch.zeros(n * 65536, dtype=torch.uint8, device="cuda")
            d_len = torch.zeros(n, dtype=torch.int64, device="cuda"); d_codes = torch.zeros(n, dtype=torch.int32, device="cuda")
            pkg.kernel_times(reset=True)
            pkg.decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_len, d_codes)
            torch.cuda.synchronize()
            kt = pkg.kernel_times()
            assert bool((d_codes == 1).all())
            out = d_out.cpu().numpy().reshape(n, 65536)
            for j in range(0, n, 97):
                assert out[j].tobytes() == orig[idx[j]]
            if n < 6000:
                assert kt["lane_ms"] < 0.05 and kt["exact_ms"] > 1.0, kt   # exact kernel only
            else:
                assert kt["lane_ms"] > 5.0 and kt["exact_ms"] < 1.0, kt    # lane kernel, empty bail list
    finally:
        pkg.set_tuning("lane_min_streams", 0)


def test_exact_kernel_wide_geometry(gpu_lib, pkg, oracle, corpus):
    """The warp-per-stream kernel has two geometries (24 and 28 warps per SM, csrc/brotli_b200_kernels.cu); a batch that
    needs fewer waves with the second one takes it (config C4's 4096 streams do).  4096 streams of every kind -- valid,
    truncated, corrupt, too little room -- through the exact kernel alone: codes, sizes and bytes equal to the oracle's."""
    rng = np.random.default_rng(21)
    comp, orig, _ = corpus.make_config("C5", 64, size=6000)
    streams, caps = [], []
    for i in range(4096):
        c, o = comp[i % 64], orig[i % 64]
        if i % 16 == 3:
            m = helpers.mutations(c, rng, 1)[0] or c
            streams.append(m); caps.append(len(o) + 16)
        elif i % 16 == 7:
            streams.append(c); caps.append(max(len(o) - 1 - i % 5, 0))
        else:
            streams.append(c); caps.append(len(o))
    try:
        assert pkg.set_tuning("lane_min_streams", 1 << 30)   # nothing goes to the lane kernel
        pkg.kernel_times(reset=True)
        got = pkg.decompress_batch(streams, caps)
        kt = pkg.kernel_times()
        assert kt["lane_ms"] < 0.05 and kt["exact_ms"] > 0.5, kt
        memo = {}
        for i, ((gres, gcode, gout), s_, c_) in enumerate(zip(got, streams, caps)):
            key = (s_, c_)
            if key not in memo:
                memo[key] = oracle.decode(s_, c_)
            ores, ocode, oout = memo[key]
            assert gcode == ocode and gout == oout, (i, gcode, ocode)
    finally:
        pkg.set_tuning("lane_min_streams", 0)


def test_lane_geometry_by_wave_fit(gpu_lib, pkg, corpus):
    """A uniform batch that fits one wave of the 14-warp geometry but not the 8-warp one is decoded by the geometry the
    device picks after the sort (csrc/brotli_b200_lane_kernel.cu, brotli_lane_geometry_kernel): every candidate geometry
    is launched, all but the chosen one exit at once.  Same bytes as the originals either way; a mixed batch of the same
    size is not fitted (one lane-kernel launch)."""
    import torch
    comp, orig, _ = corpus.make_config("C2", 48, size=4096)          # uniform: 4 KiB text windows
    compm, origm, _ = corpus.make_config("C5", 48, size=4096)        # mixed qualities and families
    n = 40000
    for blobs, originals, uniform in ((comp, orig, True), (compm, origm, False)):
        idx = np.arange(n) % len(blobs)
        in_bytes, in_off = corpus.pack([blobs[i] for i in idx])
        osz = np.array([len(originals[i]) for i in idx], dtype=np.uint64)
        out_off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(osz, out=out_off[1:])
        d_in = torch.from_numpy(in_bytes.copy()).cuda(); d_in_off = torch.from_numpy(in_off.view(np.int64)).cuda()
        d_out_off = torch.from_numpy(out_off.view(np.int64)).cuda()
        d_out = torch.zeros(int(out_off[-1]), dtype=torch.uint8, device="cuda")
        d_len = torch.zeros(n, dtype=torch.int64, device="cuda"); d_codes = torch.zeros(n, dtype=torch.int32, device="cuda")
        before = pkg.kernel_launch_count()
        pkg.kernel_times(reset=True)
        pkg.decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_len, d_codes)
        torch.cuda.synchronize()
        launches = pkg.kernel_launch_count() - before
        kt = pkg.kernel_times()
        assert bool((d_codes == 1).all()) and kt["lane_ms"] > 0.5
        out = d_out.cpu().numpy()
        for j in range(0, n, 331):
            assert out[int(out_off[j]):int(out_off[j + 1])].tobytes() == originals[idx[j]]
        # the uniform batch fits one wave of the 14-warp geometry (66 304 lanes); the mixed one stays on the default
        sizes = np.array([len(b) for b in blobs])
        is_uniform = 2 * (sizes.min() >> 8) >= (sizes.max() >> 8) and (sizes.min() >> 8) != 0
        assert is_uniform == uniform, (sizes.min(), sizes.max())
        assert launches == 7, launches  # sort (2), choice, three lane launches (two exit at once), exact kernel
        assert pkg.last_lane_geometry() == (14 if uniform else 20)


def test_differential_fuzz_error_codes(gpu_lib, pkg, oracle, corpus):
    """every BrotliDecoderErrorCode and decoded_size the GPU path reports for malformed input
    equals the oracle's, over a few thousand seeded mutations (truncations, bit flips, byte smashes, multi-byte
    corruption) of streams of every quality and every stream family, decoded as ONE batch so that damaged and intact
    streams share warps of the lane kernel."""
    rng = np.random.default_rng(4242)
    streams, caps = [], []
    for cfg, n, size in (("C5", 44, 12000), ("C3", 40, 4096), ("headline", 6, 65536)):
        comp, orig, _ = corpus.make_config(cfg, n, size=size)
        for c, o in zip(comp, orig):
            streams.append(c); caps.append(len(o))
            for m in helpers.mutations(c, rng, 30):
                streams.append(m); caps.append(len(o) + int(rng.integers(0, 48)))
    # roomy=False: a damaged stream may describe more output than its region holds (documented deviation, tests/test_hostsim.py::same)
    n_ok = check_batch(pkg, oracle, streams, caps, roomy=False)
    codes = {}
    for s, c in zip(streams, caps):
        code = oracle.decode(s, c)[1]
        codes[code] = codes.get(code, 0) + 1
    assert n_ok >= 90 and len([k for k in codes if k < 0]) >= 12, codes  # a dozen distinct format errors at least


def test_streaming_session_resumes_behind_metablocks(gpu_lib, pkg, corpus):
    """a BrotliDecoderState keeps its stream and its output on the device and each
    BrotliDecoderDecompressStream call continues behind the last complete metablock (ResumeState), so feeding a
    multi-metablock stream in small pieces costs far less device time than re-decoding every prefix."""
    import time
    pool = corpus.text_pool()
    data = pool[100000:100000 + 1000000]
    comp = corpus.compress(data, 2)          # q2: a metablock every few hundred KB of input
    for in_chunk, out_chunk in ((len(comp) // 40 + 1, 1 << 20), (8192, 65536), (len(comp), 4096)):
        st = pkg.DecoderState()
        got, pos, r = [], 0, 2
        pkg.kernel_times(reset=True)
        for _ in range(1000000):
            r, used, out = st.decompress_stream(comp[pos:pos + in_chunk], out_chunk)
            pos += used
            got.append(out)
            if r in (0, 1):
                break
        assert r == 1 and st.is_finished() and b"".join(got) == data, (in_chunk, out_chunk, r)
        st.close()
    # truncated stream: NEEDS_MORE_INPUT with everything decodable handed out, then the rest completes it
    st = pkg.DecoderState()
    half = len(comp) // 2
    r, used, out1 = st.decompress_stream(comp[:half], 1 << 21)
    assert r == 2 and used == half and data.startswith(out1) and len(out1) > 0
    r, used, out2 = st.decompress_stream(comp[half:], 1 << 21)
    assert r == 1 and out1 + out2 == data
    st.close()
    # corruption behind a completed metablock: the reference's error code for the whole stream
    bad = bytearray(comp); bad[len(comp) * 3 // 4] ^= 0x5A
    info, _ = pkg.brotli_decode(bytes(bad), len(data) + 64)
    st = pkg.DecoderState()
    r, used, _ = st.decompress_stream(bytes(bad[:half]), 1 << 21)
    r, used, _ = st.decompress_stream(bytes(bad[half:]), 1 << 21)
    assert (r == 0) == (info.code < 0) and (info.code >= 0 or st.error_code() == info.code)
    st.close()


def test_corrupt_and_too_small_reports_the_corruption(gpu_lib, pkg, oracle, corpus):
    """The reference decodes a whole ring buffer ahead of the caller's buffer: a stream that is corrupt (or truncated)
    beyond a too small capacity reports the corruption / NeedsMoreInput, not NeedsMoreOutput (src/decode.rs:1693-1738).
    Same code, same decoded_size, same bytes here -- for every mutation, through the batch and the one-shot entry."""
    rng = np.random.default_rng(31)
    pool = corpus.text_pool()
    streams, caps = [], []
    for q, lgwin, size in ((5, 22, 65536), (5, 16, 200000), (1, 18, 150000), (9, 10, 30000), (11, 22, 40000)):
        a = int(rng.integers(0, len(pool) - size))
        comp = corpus.compress(bytes(pool[a:a + size]), q, lgwin=lgwin)
        for m in [comp] + helpers.mutations(comp, rng, 40):
            for frac in (0.02, 0.3, 0.7, 0.97):
                streams.append(m); caps.append(int(size * frac))
    n = len(streams)
    got = pkg.decompress_batch(streams, caps)
    seen = {}
    for i, ((gres, gcode, gout), s, c) in enumerate(zip(got, streams, caps)):
        ores, ocode, oout = oracle.decode(s, c)
        assert (gcode, gout) == (ocode, oout), (i, gcode, ocode, len(gout), len(oout), c)
        seen[ocode] = seen.get(ocode, 0) + 1
    assert seen.get(3, 0) > 50 and sum(v for k, v in seen.items() if k < 0) > 50 and seen.get(2, 0) > 5, seen
    # the one-shot entry takes the same path
    for i in range(0, n, 37):
        info, out = pkg.brotli_decode(streams[i], caps[i])
        ores, ocode, oout = oracle.decode(streams[i], caps[i])
        assert (info.code, out) == (ocode, oout)


def test_c_main_acceptance(gpu_lib, pkg, tmp_path):
    """The reference's own C example (c/main.c:16-78: one-shot round trip of a literal stream, then a streaming call on
    a corrupt stream expecting BROTLI_DECODER_ERROR_FORMAT_CONTEXT_MAP_REPEAT), compiled UNMODIFIED against this
    library's header and linked against libbrotli_b200.so, must print what it prints with the reference."""
    import os, shutil, subprocess
    src = os.path.join(helpers.ROOT, "tests", "golden", "c_main.c")  # verbatim copy of the reference's c/main.c (test vector)
    inc = tmp_path / "brotli"
    inc.mkdir()
    # c/main.c includes "brotli/decode.h": this library's header under the reference's include path
    shutil.copy(os.path.join(helpers.ROOT, "include", "brotli_b200", "decode.h"), inc / "decode.h")
    exe = tmp_path / "c_main"
    libdir = os.path.dirname(pkg.LIB_PATH)
    r = subprocess.run(["gcc", "-O1", "-I", str(tmp_path), src, "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(pkg.LIB_PATH),
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # after its three self-checks (asserts) main() is a stdin -> stdout decompressor over BrotliDecoderDecompressStream
    comp = helpers.golden_fixture("alice29.txt.compressed")
    r = subprocess.run([str(exe)], input=comp, capture_output=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr)
    assert hashlib.sha256(r.stdout).hexdigest() == MAN["alice29.txt.compressed"]["original_sha256"]
    r = subprocess.run([str(exe)], input=comp[:20000], capture_output=True, timeout=300)  # truncated: "Unexpected EOF", exit 1
    assert r.returncode == 1 and b"Unexpected EOF" in r.stderr


def test_c4_real_shape(gpu_lib, pkg, oracle, corpus):
    """BASELINE config C4 at its real shape: 16 MiB streams, window bits 24 (long-distance backreferences, hundreds of
    prefix codes per metablock), a batch of 8 -- bit-exact against the originals and the oracle's result fields."""
    comp, orig, desc = corpus.make_config("C4", 8, size=16 << 20)
    assert all(len(o) == 16 << 20 for o in orig)
    got = pkg.decompress_batch(comp, [len(o) for o in orig])
    for (gres, gcode, gout), o in zip(got, orig):
        assert (gres, gcode) == (1, 1) and gout == o
    ores, ocode, oout = oracle.decode(comp[0], len(orig[0]))
    assert ocode == 1 and oout == orig[0]
