"""Short evaluate_h workload for ncu: gates / permutation / lookup row loops at k = K over COLS extended columns."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import halo2_vectordb_b200 as h
k = int(os.environ.get("K", "16")); cols = int(os.environ.get("COLS", "128"))
h.init(0)
dom = h.EvaluationDomain(4, k)
ne = 1 << dom.extended_k
g = torch.Generator(device="cuda").manual_seed(1)
def mk(c):
    a = torch.randint(-(1 << 63), (1 << 63) - 1, (c, ne, 4), dtype=torch.int64, generator=g, device="cuda")
    a[..., 3] &= (1 << 60) - 1
    return a
q, a, sg, z, m = mk(cols), mk(cols), mk(cols), mk(cols // 2), mk(9)
dh = torch.zeros((ne, 4), dtype=torch.int64, device="cuda")
y = m[0, :3].cpu().numpy().view(np.uint64)
p = lambda i: m[i].data_ptr()
for _ in range(2):
    dom.quotient_gates(dh.data_ptr(), y[0], cols, q.data_ptr(), ne, a.data_ptr(), ne)
    print("gates", h.last_kernel_ms()["ntt"])
    dom.quotient_permutation(dh.data_ptr(), y[0], y[1], y[2], cols, 2, a.data_ptr(), ne, sg.data_ptr(), ne, z.data_ptr(), ne, p(1), p(2), p(3), 5)
    print("perm", h.last_kernel_ms()["ntt"])
    dom.quotient_lookup(dh.data_ptr(), y[0], y[1], y[2], p(4), p(5), p(6), p(7), p(8), p(1), p(2), p(3))
    print("lookup", h.last_kernel_ms()["ntt"])
