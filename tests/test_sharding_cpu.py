"""CPU, world_size 2 over gloo: the N > 1 path shards independent columns (no data-path collective);
the only exchange is gathering the per-column results back in transcript order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from halo2_vectordb_b200 import sharding


def test_column_shard_partition():
    for n_cols in (0, 1, 7, 64, 617):
        for world in (1, 2, 4, 8):
            owned = [sharding.column_shard(n_cols, r, world) for r in range(world)]
            flat = sorted(i for o in owned for i in o)
            assert flat == list(range(n_cols))
            assert max(map(len, owned)) - min(map(len, owned)) <= 1
            assert sharding.shard_sizes(n_cols, world) == [len(o) for o in owned]
    with pytest.raises(ValueError):
        sharding.column_shard(4, 2, 2)


def _fake_commit(col_idx):
    # stands in for a rank-local commitment: any deterministic function of the column index
    return np.array([col_idx * 8 + k + (1 << 63) for k in range(8)], dtype=np.uint64)


def _worker(rank, world, port, n_cols, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.column_shard(n_cols, rank, world)
        local = np.stack([_fake_commit(i) for i in mine]) if mine else np.zeros((0, 8), dtype=np.uint64)
        full = sharding.gather_in_column_order(local, n_cols, rank, world)
        q.put((rank, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cols", [1, 5, 12])
def test_gather_in_column_order_gloo(n_cols):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.stack([_fake_commit(i) for i in range(n_cols)])
    for _, full in results:
        assert (full == expect).all()


# ---- one multiexp split by index range: slices + one 64-byte all-gather + a sum of `world` points -------------------
def test_index_slice_partition():
    for n in (0, 1, 5, 64, 1000, 1 << 16):
        for world in (1, 2, 3, 4, 8):
            sl = [sharding.index_slice(n, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            assert max(h - l for l, h in sl) - min(h - l for l, h in sl) <= max(1, world - 1)
    with pytest.raises(ValueError):
        sharding.index_slice(8, 2, 2)


def _msm_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O    # the checker stands in for the rank-local GPU MSM in this CPU test
        bases, scalars = O.gen_bases(n, threads=1), O.fr_fill(n, 77)

        def local(lo, hi):
            return O.best_multiexp_affine(scalars[lo:hi], bases[lo:hi], threads=1) if hi > lo else np.zeros(8, dtype=np.uint64)

        def combine(parts):
            acc = np.zeros(12, dtype=np.uint64)
            for p in parts:
                acc = O.g1_add_mixed(acc, p)
            return O.g1_to_affine(acc)

        got = sharding.sliced_multiexp(local, combine, n, rank, world)
        q.put((rank, got, O.msm_closed_form(scalars)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 257, 4096])
def test_sliced_multiexp_gloo(n):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_msm_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, got, want in results:
        assert (got == want).all()
