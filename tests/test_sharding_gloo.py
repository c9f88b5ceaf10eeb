"""N > 1 host logic on CPU: world_size-2 gloo.  Frame sharding, the final gather of detection
records to rank 0 and the max-over-ranks timing reduction (no GPU, no kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as entry


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_results(pkg, lo, hi, cap):
    """Deterministic fake detections for global frames lo..hi-1."""
    tags = np.zeros((hi - lo, cap), pkg.TAG_DTYPE)
    counts = np.zeros(hi - lo, np.int32)
    for j, f in enumerate(range(lo, hi)):
        n = f % (cap + 1)
        counts[j] = n
        for k in range(n):
            tags[j, k]["id"] = (f * 7 + k) % 587
            tags[j, k]["xy"] = np.arange(8, dtype=np.float32) + f + 0.25 * k
    return tags, counts


def _worker(rank, world, port, n_total, cap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = entry.load_package()
        from aprilgrid_rs_b200 import shard
        lo, hi = shard.shard_range(n_total, rank, world)
        tags, counts = _fake_results(pkg, lo, hi, cap)
        all_tags, all_counts = shard.gather_detections(tags, counts, n_total)
        slow = shard.max_over_ranks(1.0 + rank)
        if rank == 0:
            want_t, want_c = _fake_results(pkg, 0, n_total, cap)
            ok = np.array_equal(all_counts, want_c) and all_tags.tobytes() == want_t.tobytes()
            q.put((ok, slow, int(all_counts.sum())))
        else:
            assert all_tags is None and all_counts is None
            q.put((True, slow, -1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8, 1])
def test_shard_gather_world2(n_total):
    world, cap = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, cap, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in res)
    assert all(abs(r[1] - 2.0) < 1e-12 for r in res)  # max over ranks of (1 + rank)


def test_shard_ranges_cover_and_are_contiguous(pkg):
    from aprilgrid_rs_b200 import shard
    for n in (0, 1, 7, 1024, 1000):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            for a, b in zip(r, r[1:]):
                assert a[1] == b[0]
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)
