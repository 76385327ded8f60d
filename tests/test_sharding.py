"""Host-side multi-GPU logic on CPU: the per-rank batch split and the final counter reduction, with a
real world-size-2 `gloo` process group (no GPU needed).  Each rank decodes ITS slice with the oracle
(the checker standing in for the device here) so the test also proves the slices tile the batch."""
import importlib
import os
import socket

import numpy as np
import pytest

sharding = importlib.import_module("rust-brotli-decompressor_b200.sharding")


def test_even_split_tiles_the_batch():
    for n in (0, 1, 7, 262144):
        for world in (1, 2, 4, 8):
            cuts = [sharding.even_split(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.even_split(10, 2, 2)


def test_byte_balanced_split():
    rng = np.random.default_rng(7)
    c = rng.integers(100, 30000, size=1000)
    d = rng.integers(0, 70000, size=1000)
    in_off = np.concatenate([[0], np.cumsum(c)]).astype(np.uint64)
    out_off = np.concatenate([[0], np.cumsum(d)]).astype(np.uint64)
    for world in (1, 2, 4, 8):
        cuts = sharding.byte_balanced_split(in_off, out_off, world)
        assert cuts[0] == 0 and cuts[-1] == 1000 and all(a <= b for a, b in zip(cuts, cuts[1:]))
        cost = [(in_off[cuts[r + 1]] - in_off[cuts[r]]) + (out_off[cuts[r + 1]] - out_off[cuts[r]]) for r in range(world)]
        total = float(sum(cost))
        assert max(cost) <= total / world + 100000  # within one stream of the ideal share
    # ragged: empty batch and a batch with one huge stream
    assert sharding.byte_balanced_split(np.zeros(1, np.uint64), np.zeros(1, np.uint64), 4) == [0, 0, 0, 0, 0]
    cuts = sharding.byte_balanced_split(np.array([0, 10, 20, 1000020], np.uint64), np.array([0, 1, 2, 3], np.uint64), 2)
    assert cuts[0] == 0 and cuts[-1] == 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    import torch.distributed as dist
    import helpers
    sh = importlib.import_module("rust-brotli-decompressor_b200.sharding")
    corpus = importlib.import_module("tools.corpus")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comp, orig, _ = corpus.make_config("C3", 24, size=4096)  # same seeded batch on every rank
        lo, hi = sh.even_split(len(comp), world, rank)
        oracle = helpers.Oracle()
        ok, c_bytes, d_bytes, failed = True, 0, 0, 0
        for s, o in zip(comp[lo:hi], orig[lo:hi]):
            res, code, out = oracle.decode(s, len(o))
            ok = ok and out == o
            failed += code != 1
            c_bytes += len(s)
            d_bytes += len(out)
        r = sh.reduce_run(dist, elapsed_ms=10.0 * (rank + 1), n_streams=hi - lo, c_bytes=c_bytes, d_bytes=d_bytes,
                          n_failed=failed, bit_exact=ok)
        r["want_c"] = sum(len(s) for s in comp)
        r["want_d"] = sum(len(o) for o in orig)
        q.put((rank, r))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_split_and_reduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        r = got[rank]
        assert r["streams"] == 24 and r["failed"] == 0 and r["bit_exact"]
        assert r["compressed_bytes"] == r["want_c"] and r["decompressed_bytes"] == r["want_d"]
        assert r["ms"] == 20.0  # max over ranks
        assert abs(r["decompressed_gbs"] - r["want_d"] / 0.020 / 1e9) < 1e-9
