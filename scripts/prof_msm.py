"""Short MSM workload for ncu / tuning: commit_lagrange of COLS columns at 2^K (device-resident), twice.
DIST=uniform|witness selects the scalar distribution (witness: 60% {0,1}, 30% < 2^LOOKUP_BITS, 5% full
width, 5% r - small -- the shape of FixedPointChip columns, /root/reference/src/gadget/fixed_point.rs:68-119)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import halo2_vectordb_b200 as h

from halo2_vectordb_b200.synthetic import witness_like


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
