"""Multi-GPU plumbing of the detect path: image-wise frame sharding and the final gather.

Frames are independent (`TagDetector::detect` is stateless, src/detector.rs:505-540), so a batch
is split into contiguous frame ranges, one per rank / GPU, and no collective is needed on the
data path.  The only exchange is the optional final gather of the fixed-size detection records
to rank 0 and the max-over-ranks of the elapsed time.  Works with any torch.distributed backend
(NCCL on GPUs; the CPU tests use gloo).
"""
import numpy as np


def shard_range(n_frames, rank, world):
    """Contiguous range [lo, hi) of rank `rank`: [g*B/G, (g+1)*B/G)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return lo, hi


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (elapsed time: a job is as slow as its slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_detections(tags, counts, n_total, device=None):
    """Gather per-rank results to rank 0.

    tags: (n_local, cap) structured TAG_DTYPE array, counts: (n_local,) int32 for this rank's
    contiguous shard of an n_total-frame batch.  Returns (tags_all, counts_all) in global frame
    order on rank 0 and (None, None) elsewhere.  Shards may have different sizes."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tags, counts
    world, rank = dist.get_world_size(), dist.get_rank()
    cap = tags.shape[1]
    rec_words = tags.dtype.itemsize // 4
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    # pad every shard to n_max frames so that a plain all_gather works on every backend
    buf = np.zeros((n_max, cap * rec_words + 1), np.int32)
    n_local = tags.shape[0]
    buf[:n_local, :-1] = tags.view(np.int32).reshape(n_local, cap * rec_words)
    buf[:n_local, -1] = counts
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    if rank != 0:
        return None, None
    tags_all = np.zeros((n_total, cap), tags.dtype)
    counts_all = np.zeros(n_total, np.int32)
    for r, (lo, hi) in enumerate(sizes):
        a = out[r].cpu().numpy()[: hi - lo]
        tags_all[lo:hi] = a[:, :-1].copy().view(tags.dtype).reshape(hi - lo, cap)
        counts_all[lo:hi] = a[:, -1]
    return tags_all, counts_all
