"""Short MSM workload for ncu / tuning: commit_lagrange of COLS columns at 2^K (device-resident), twice.
DIST=uniform|witness selects the scalar distribution (witness: 60% {0,1}, 30% < 2^LOOKUP_BITS, 5% full
width, 5% r - small -- the shape of FixedPointChip columns, /root/reference/src/gadget/fixed_point.rs:68-119)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import halo2_vectordb_b200 as h

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R2 = 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7


def limbs(x):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def witness_like(cols, n, lookup_bits, seed):
    """(cols, n, 4) uint64 Montgomery scalars with the skew of real witness columns (to_mont on the device)."""
    rng = np.random.default_rng(seed)
    tot = cols * n
    canon = np.zeros((tot, 4), dtype=np.uint64)
    sel = rng.integers(0, 100, tot)
    canon[:, 0] = np.where(sel < 60, rng.integers(0, 2, tot), rng.integers(0, 1 << lookup_bits, tot)).astype(np.uint64)
    full = sel >= 90
    nf = int(full.sum())
    canon[full] = rng.integers(0, 1 << 62, (nf, 4)).astype(np.uint64)          # < 2^254 < r
    neg = sel >= 95
    small = rng.integers(1, 1 << 40, int(neg.sum()))
    rl = np.array(limbs(R_MOD), dtype=np.uint64)
    negv = np.tile(rl, (int(neg.sum()), 1))
    negv[:, 0] = negv[:, 0] - small.astype(np.uint64)                           # low limb of r is > 2^40: no borrow
    canon[neg] = negv
    r2 = np.tile(np.array(limbs(R2), dtype=np.uint64), (tot, 1))
    return h.selftest_field(0, 0, canon, r2).reshape(cols, n, 4)


if __name__ == "__main__":
    k = int(os.environ.get("K", "16")); cols = int(os.environ.get("COLS", "32")); n = 1 << k
    dist = os.environ.get("DIST", "uniform")
    h.init(0)
    srs = h.ParamsKZG(k, None, h.synthetic_bases(n))
    if dist == "uniform":
        g = torch.Generator().manual_seed(1)
        a = torch.randint(-(1 << 63), (1 << 63) - 1, (cols, n, 4), dtype=torch.int64, generator=g)
        a[..., 3] &= (1 << 60) - 1
    else:
        a = torch.from_numpy(witness_like(cols, n, min(k - 1, 19), 1).view(np.int64))
    d = a.cuda(); out = torch.zeros((cols, 8), dtype=torch.int64, device="cuda")
    for _ in range(3):
        srs.commit_batch_dev(d.data_ptr(), n, cols, n, out.data_ptr())
        ms = h.last_kernel_ms()
        print(dist, k, cols, "total %.3f ms" % sum(ms.values()), {k_: round(v, 3) for k_, v in ms.items() if v})
