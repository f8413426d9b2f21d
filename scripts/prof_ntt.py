"""Short NTT workload for ncu: lagrange_to_coeff (2^k) and coeff_to_extended (2^k -> 2^(k+2)) of COLS columns."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import halo2_vectordb_b200 as h
k = int(os.environ.get("K", "16")); cols = int(os.environ.get("COLS", "32")); n = 1 << k
h.init(0)
dom = h.EvaluationDomain(4, k)
g = torch.Generator().manual_seed(1)
a = torch.randint(-(1 << 63), (1 << 63) - 1, (cols, n, 4), dtype=torch.int64, generator=g)
a[..., 3] &= (1 << 60) - 1
d = a.cuda(); o1 = torch.empty_like(d); o2 = torch.empty((cols, 4 * n, 4), dtype=torch.int64, device="cuda")
for _ in range(2):
    dom.transform_dev(h.OP_LAGRANGE_TO_COEFF, d.data_ptr(), n, o1.data_ptr(), n, cols)
    print("l2c", h.last_kernel_ms()["ntt"])
    dom.transform_dev(h.OP_COEFF_TO_EXTENDED, d.data_ptr(), n, o2.data_ptr(), 4 * n, cols)
    print("c2e", h.last_kernel_ms()["ntt"])
