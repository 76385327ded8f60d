#!/usr/bin/env python3
"""bench.py -- decompressed GB/s of the batched Brotli decoder on the BASELINE.json headline workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`): 262 144 independent 64 KiB text streams, brotli -q5 lgwin 22,
split evenly over the N GPUs of one box (total work fixed -> "strong" scaling; no data-path
collective, torch.distributed only reduces the timing).  A step is one decode pass over the whole
batch.  `value` is measured with the batch resident in HBM (CUDA events, max over ranks); `e2e`
goes through the host-buffer C-ABI call BrotliB200DecompressBatchPacked with pinned host buffers,
so every step pays the H2D copy of the compressed bytes and the D2H copy of the decoded bytes.

The unique-stream cap bounds host compression time: U unique streams are
compressed with libbrotlienc and tiled by seeded permutations into physically distinct copies, so
the device still reads 262 144 x C and writes 262 144 x D bytes per step (far larger than L2).
Every step's output is verified: per-stream checksums of ALL streams against the originals plus
a full byte compare of a 4096-stream sample.

`--impl reference`: the reference crate is Rust and cannot be built here (no rustc); its decode path
is timed as the CPU port in oracle/ (checked against the reference's fixtures) on all host threads.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decompressed GB/s on 256Kx64KiB brotli batch"
N_STREAMS = 262144
STREAM_BYTES = 65536
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per stream of the dominant kernel, from the committed
# `ncu --set full` capture of a headline-shaped launch (profiles/r01/zz_ncu_lane_kernel*); None = no capture yet.
NCU_TRAFFIC_BYTES_PER_STREAM = 1267320  # profiles/r01/zz_ncu_lane_kernel_raw.csv: (155.07 + 11.04) GB over a 131 072-stream launch


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=N_STREAMS, help="total streams over all GPUs (headline: 262144)")
    ap.add_argument("--unique", type=int, default=4096, help="unique compressed streams (tiled into distinct copies)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-seconds budget of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    return ap.parse_args()


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    return helpers.Oracle()


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].strip().lower() == "active"})
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


def build_unique(corpus, pkg, n_unique, threads):
    comp, orig, desc = corpus.make_config("headline", n_unique, size=STREAM_BYTES, threads=threads)
    sums = np.array([pkg.checksum_reference(o) for o in orig], dtype=np.uint64)
    return comp, orig, sums, desc


def tile_indices(n_total, n_unique, seed=0xB2000000):
    """Global stream j decodes unique stream idx[j]: consecutive seeded permutations of range(U)."""
    rng = np.random.default_rng(seed)
    blocks = [rng.permutation(n_unique) for _ in range((n_total + n_unique - 1) // n_unique)]
    return np.concatenate(blocks)[:n_total]


def shard(n_total, world, rank):
    """Contiguous per-GPU slice; streams are equal-sized here so an even count split is byte-balanced."""
    sharding = importlib.import_module("rust-brotli-decompressor_b200.sharding")
    return sharding.even_split(n_total, world, rank)


def run_reference(args, rank, world):
    """CPU port of the reference decode path (oracle/), all host threads, bounded sample per step."""
    if rank != 0:
        return
    corpus = importlib.import_module("tools.corpus")
    oracle = load_oracle()
    threads = os.cpu_count() or 1
    n_unique = min(args.unique, 1024)
    comp, orig, _ = corpus.make_config("headline", n_unique, size=STREAM_BYTES)
    in_bytes, in_off = corpus.pack(comp)
    out_off = np.arange(n_unique + 1, dtype=np.uint64) * np.uint64(STREAM_BYTES)
    out = np.zeros(int(out_off[-1]), dtype=np.uint8)
    out_len = np.zeros(n_unique, dtype=np.uint64)
    codes = np.zeros(n_unique, dtype=np.int32)
    # calibrate so one step is ~1-2 s of wall clock
    t0 = time.perf_counter()
    oracle.decode_batch(in_bytes, in_off, out, out_off, out_len, codes, threads)
    dt = time.perf_counter() - t0
    assert (codes == 1).all() and out.tobytes() == b"".join(orig), "oracle output differs from the originals"
    reps = max(1, min(64, int(1.5 / max(dt, 1e-3))))
    sample_streams = reps * n_unique

    def step():
        for _ in range(reps):
            oracle.decode_batch(in_bytes, in_off, out, out_off, out_len, codes, threads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    gbs = sample_streams * STREAM_BYTES / dt / 1e9
    sample = "%d of %d streams per step (%d unique x %d), oracle port of src/decode.rs, %d threads, %s" % (
        sample_streams, args.streams, n_unique, reps, threads, cpu_model())
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "256Kx64KiB text q5 lgwin22 (bounded sample: %s)" % sample},
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference crate is Rust; rustc/cargo absent -> timed as the C port in oracle/ (pinned to the reference's fixtures)",
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args, corpus, comp, orig):
    """Oracle port on all host cores over a bounded sample (about args.cpu_seconds CPU-seconds)."""
    oracle = load_oracle()
    threads = os.cpu_count() or 1
    n_u = len(comp)
    in_bytes, in_off = corpus.pack(comp)
    out_off = np.arange(n_u + 1, dtype=np.uint64) * np.uint64(STREAM_BYTES)
    out = np.zeros(int(out_off[-1]), dtype=np.uint8)
    out_len = np.zeros(n_u, dtype=np.uint64)
    codes = np.zeros(n_u, dtype=np.int32)
    k = min(n_u, 256)
    t0 = time.perf_counter()  # single-thread calibration on k streams
    oracle.decode_batch(in_bytes, in_off[:k + 1], out, out_off[:k + 1], out_len[:k], codes[:k], 1)
    per_core = k * STREAM_BYTES / (time.perf_counter() - t0)
    reps = max(1, int(args.cpu_seconds * per_core / (n_u * STREAM_BYTES)))
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.decode_batch(in_bytes, in_off, out, out_off, out_len, codes, threads)
    dt = time.perf_counter() - t0
    ok = bool((codes == 1).all()) and out.tobytes() == b"".join(orig)
    res = {"value": round(reps * n_u * STREAM_BYTES / dt / 1e9, 4), "unit": "GB/s", "cores": threads, "kind": "port",
           "sample": "%d streams (%d unique x %d passes) of the same workload; oracle/ port of src/decode.rs; %s; single-thread %.3f GB/s"
                     % (reps * n_u, n_u, reps, cpu_model(), per_core / 1e9),
           "output_matches_originals": ok}
    # context: Google's C decoder (upper bound on the Rust crate per its README.md:83-87), single thread
    try:
        t0 = time.perf_counter()
        for c in comp[:k]:
            corpus.system_decompress(c, STREAM_BYTES)
        res["libbrotlidec_1thread_gbs"] = round(k * STREAM_BYTES / (time.perf_counter() - t0) / 1e9, 4)
    except OSError:
        pass
    return res


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # Everything libraries print to stdout (NCCL's version banner, ...) goes to stderr: stdout carries the one JSON line.
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the decoder has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module("rust-brotli-decompressor_b200")
    corpus = importlib.import_module("tools.corpus")
    pkg.lib()

    # ---- workload ----
    threads = max(1, (os.cpu_count() or 1) // world)
    n_unique = min(args.unique, args.streams)
    comp, orig, usums, desc = build_unique(corpus, pkg, n_unique, threads)
    idx = tile_indices(args.streams, n_unique)
    lo, hi = shard(args.streams, world, rank)
    my = idx[lo:hi]
    n = len(my)
    usize = np.array([len(c) for c in comp], dtype=np.uint64)
    in_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(usize[my], out=in_off[1:])
    out_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(STREAM_BYTES)
    c_bytes, d_bytes = int(in_off[-1]), int(out_off[-1])

    h_in = torch.empty(c_bytes, dtype=torch.uint8).pin_memory()
    h_in_np = h_in.numpy()
    ucomp = [np.frombuffer(c, dtype=np.uint8) for c in comp]
    for b0 in range(0, n, 4096):
        sel = my[b0:b0 + 4096]
        h_in_np[int(in_off[b0]):int(in_off[min(b0 + 4096, n)])] = np.concatenate([ucomp[i] for i in sel])
    d_in = h_in.cuda(non_blocking=True)
    d_in_off = torch.from_numpy(in_off.view(np.int64)).cuda()
    d_out_off = torch.from_numpy(out_off.view(np.int64)).cuda()
    d_out = torch.empty(d_bytes, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int64, device="cuda")
    d_codes = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_sums = torch.zeros(n, dtype=torch.int64, device="cuda")
    want_sums = torch.from_numpy(usums[my].view(np.int64)).cuda()
    torch.cuda.synchronize()

    def step():
        pkg.decompress_batch_device(n, d_in, d_in_off, d_out, d_out_off, d_len, d_codes)

    def verify():
        pkg.checksum_batch_device(n, d_out, d_out_off, d_len, d_sums)
        torch.cuda.synchronize()
        ok = bool((d_codes == 1).all()) and bool((d_len == STREAM_BYTES).all()) and bool((d_sums == want_sums).all())
        ns = min(n, 4096)  # full byte compare of a sample
        sample = d_out[:ns * STREAM_BYTES].cpu().numpy().reshape(ns, STREAM_BYTES)
        for j in range(ns):
            if sample[j].tobytes() != orig[my[j]]:
                return False
        return ok

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: batch resident in HBM ----
    for _ in range(max(args.warmup, 3)):
        d_out.zero_()
        step()
    bit_exact = verify()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = pkg.kernel_launch_count()
    pkg.kernel_times(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = pkg.kernel_launch_count() - launches0
    ktimes = pkg.kernel_times()  # CUDA events recorded by the library on the launching stream around each kernel
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1) / args.steps
    bit_exact = bit_exact and verify()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(d_bytes), float(c_bytes), float(n), float(1 if bit_exact else 0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exact_min = tot[3:].clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(exact_min, op=dist.ReduceOp.MIN)
        tot[3] = exact_min[0]
    ms_max = float(t.item())
    all_d, all_c, all_n, all_exact = [float(x) for x in tot.tolist()]
    value = all_d / (ms_max * 1e-3) / 1e9

    # ---- e2e: host buffers through the C ABI (H2D + decode + D2H per step) ----
    e2e = None
    if not args.no_e2e:
        h_out = torch.empty(d_bytes, dtype=torch.uint8).pin_memory()
        h_len = np.zeros(n, dtype=np.uint64)
        h_codes = np.zeros(n, dtype=np.int32)
        h_out_np = h_out.numpy()

        def e2e_step():
            pkg.decompress_batch_packed(h_in_np, in_off, h_out_np, out_off, h_len, h_codes)

        for _ in range(2):
            e2e_step()
        ok = bool((h_codes == 1).all()) and bool((h_len == STREAM_BYTES).all())
        for j in range(0, n, max(1, n // 2048)):
            ok = ok and h_out_np[j * STREAM_BYTES:(j + 1) * STREAM_BYTES].tobytes() == orig[my[j]]
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([dt, 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": round(all_d / te[0].item() / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": int(all_c + 16 * (all_n + world)),
               "d2h_bytes_per_step": int(all_d + 12 * all_n), "ms_per_step": round(te[0].item() * 1e3, 3),
               "api": "BrotliB200DecompressBatchPacked (pinned host buffers, chunked H2D/decode/D2H pipeline)",
               "bit_exact": te[1].item() == 0.0, "last_kernel_span_ms": round(pkg.lib().BrotliB200LastKernelMs(), 3)}
        del h_out

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        algo = (c_bytes + d_bytes) / 1e9  # per launch on this GPU: compressed bytes read once + decoded bytes written once
        kn = max(ktimes["launches"], 1)
        lane_ms, exact_ms = ktimes["lane_ms"] / kn, ktimes["exact_ms"] / kn
        dominant = "brotli_decode_lane_kernel" if lane_ms >= exact_ms else "brotli_decode_batch_kernel"
        achieved = algo / (max(lane_ms, exact_ms) * 1e-3)
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_max, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "%d x 64 KiB streams (%s), %d unique tiled into distinct copies; %d streams per GPU" % (
                           int(all_n), desc, n_unique, n),
                       "compressed_bytes": int(all_c), "decompressed_bytes": int(all_d), "l2_policy": "inputs+outputs per step (%.1f GB/GPU) >> 126 MB L2" % algo,
                       "parallelism": "independent streams split evenly over %d GPU(s), no data-path collective" % world},
            "bit_exact": bool(all_exact >= 1.0),  # minimum over ranks of each rank's verdict
            "verification": "per-stream 64-bit checksums of all streams vs originals + full byte compare of 4096 streams per GPU",
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
                         "traffic": NCU_TRAFFIC_BYTES_PER_STREAM * n if NCU_TRAFFIC_BYTES_PER_STREAM else None, "kernel": dominant,
                         "kernel_ms": round(max(lane_ms, exact_ms), 3), "other_kernel_ms": round(min(lane_ms, exact_ms), 3),
                         "streams_bailed_to_exact_kernel": ktimes["bailed"],
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                         "algorithmic_bytes_per_launch": int(c_bytes + d_bytes)},
            "clocks": clocks,
            "e2e": e2e,
        }
        line["cpu_baseline"] = cpu_baseline(args, corpus, comp[:min(n_unique, 2048)], orig[:min(n_unique, 2048)]) if world == 1 and not args.no_cpu else None
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
