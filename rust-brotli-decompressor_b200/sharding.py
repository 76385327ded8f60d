"""Per-GPU batch split and the end-of-run counter reduction.

Streams are independent, so N GPUs each take a contiguous slice of the batch and nothing is
exchanged on the data path.  torch.distributed (NCCL on the GPU box, gloo in the CPU tests) only
carries the final reduction: SUM{streams, compressed bytes, decoded bytes, failures}, MAX{elapsed},
MIN{bit-exact flag}.
"""
import numpy as np


def even_split(n_streams, world, rank):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one stream."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n_streams * rank // world, n_streams * (rank + 1) // world


def byte_balanced_split(in_off, out_off, world):
    """Contiguous slices balanced by the algorithmic bytes sum(C_i + D_i) (prefix-sum split into `world`
    parts).  in_off/out_off are the u64[n+1] offset arrays of a packed batch.  Returns world+1 cut
    points: rank r decodes streams [cuts[r], cuts[r+1])."""
    in_off = np.asarray(in_off, dtype=np.uint64)
    out_off = np.asarray(out_off, dtype=np.uint64)
    n = len(in_off) - 1
    if n < 0 or len(out_off) != n + 1:
        raise ValueError("offset arrays must both have n+1 entries")
    cost = (in_off - in_off[0]).astype(np.float64) + (out_off - out_off[0]).astype(np.float64)  # prefix sums, cost[n] = total
    cuts = [0]
    for r in range(1, world):
        target = cost[n] * r / world
        c = int(np.searchsorted(cost, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return cuts


def reduce_run(dist, elapsed_ms, n_streams, c_bytes, d_bytes, n_failed, bit_exact, device="cpu"):
    """All-reduce the per-rank results of one timed run.  `dist` is torch.distributed (initialised) or
    None for a single process.  Returns a dict with the whole-job aggregates."""
    import torch
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    s = torch.tensor([float(n_streams), float(c_bytes), float(d_bytes), float(n_failed)], dtype=torch.float64, device=device)
    e = torch.tensor([1.0 if bit_exact else 0.0], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(e, op=dist.ReduceOp.MIN)
    n, c, d, f = [float(x) for x in s.tolist()]
    ms = float(t.item())
    return {"ms": ms, "streams": int(n), "compressed_bytes": int(c), "decompressed_bytes": int(d), "failed": int(f),
            "bit_exact": bool(e.item() == 1.0), "decompressed_gbs": (d / (ms * 1e-3) / 1e9) if ms > 0 else 0.0}
