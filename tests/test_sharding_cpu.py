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
