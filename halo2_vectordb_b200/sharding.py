"""Column sharding across GPUs: one process per GPU, column j -> rank j mod G.

A proof is hundreds of independent length-2^k columns that share the SRS bases and the domain
twiddles, so each rank holds a full SRS/twiddle replica and transforms / commits only its own
columns; there is NO collective on the math path.  The only exchange is returning the 64-byte
commitments (and 32-byte evaluations) to the host transcript *in the original column order*,
which is what `gather_in_column_order` does over torch.distributed (gloo on CPU, nccl on GPU).
"""
import numpy as np


def column_shard(n_cols, rank, world):
    """Indices of the columns rank `rank` of `world` owns (round-robin keeps per-rank work balanced
    when column kinds -- advice, lookup, permutation products -- arrive in blocks)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_cols, world))


def shard_sizes(n_cols, world):
    return [len(range(r, n_cols, world)) for r in range(world)]


def gather_in_column_order(local, n_cols, rank, world, width=8, device=None):
    """all-gather per-rank results (len(column_shard) x width uint64) and restore transcript order."""
    local = np.ascontiguousarray(local, dtype=np.uint64).reshape(-1, width)
    if local.shape[0] != len(column_shard(n_cols, rank, world)):
        raise ValueError("local result count does not match this rank's shard")
    if world == 1:
        return local.copy()
    import torch
    import torch.distributed as dist

    per = max(shard_sizes(n_cols, world))
    buf = torch.zeros((per, width), dtype=torch.int64, device=device)
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(local.view(np.int64)).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    res = np.zeros((n_cols, width), dtype=np.uint64)
    for r in range(world):
        idx = column_shard(n_cols, r, world)
        res[idx] = out[r][: len(idx)].cpu().numpy().view(np.uint64)
    return res


# ---- one standalone multiexp split over the GPUs by index range -----------------------
def index_slice(n, rank, world):
    """[lo, hi) of the scalar / base index range rank `rank` owns: contiguous, near-equal slices."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    per = -(-n // world) if n else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_partials(partial_affine, rank, world, device=None):
    """The one exchange step of a sliced multiexp: all-gather of the `world` affine partial sums (64 B each).
    Returns a (world, 8) uint64 array, identical on every rank."""
    partial_affine = np.ascontiguousarray(partial_affine, dtype=np.uint64).reshape(8)
    if world == 1:
        return partial_affine.reshape(1, 8).copy()
    import torch
    import torch.distributed as dist

    mine = torch.from_numpy(partial_affine.view(np.int64).copy())
    if device is not None:
        mine = mine.to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return np.stack([o.cpu().numpy().view(np.uint64) for o in out])


def sliced_multiexp(local_msm, combine, n, rank, world, device=None):
    """sum_i s_i B_i with the index range split over `world` ranks.  `local_msm(lo, hi)` returns this rank's affine
    partial sum over [lo, hi) (on a GPU: ParamsKZG.commit on a handle holding bases[lo:hi], or best_multiexp);
    `combine(partials)` adds the gathered partial sums (on a GPU: halo2_vectordb_b200.g1_sum).  Every rank returns
    the same affine point."""
    lo, hi = index_slice(n, rank, world)
    part = local_msm(lo, hi)
    return combine(gather_partials(part, rank, world, device))
