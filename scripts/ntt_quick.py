"""NTT kernel timings for before/after comparisons (1 GPU, CUDA events on the launching stream, median of 5 after 3 warm-ups):
lagrange_to_coeff / coeff_to_extended / extended_to_coeff at 2^16 x 64 columns and 2^20 x 8 columns, device-resident."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import halo2_vectordb_b200 as h
h.init(0)
g = torch.Generator(device="cuda").manual_seed(3)
def cols_dev(c, n):
    a = torch.randint(-(1 << 63), (1 << 63) - 1, (c, n, 4), dtype=torch.int64, generator=g, device="cuda")
    a[..., 3] &= (1 << 60) - 1
    return a
def timed(fn):
    ms = []
    for i in range(8):
        fn()
        if i >= 3: ms.append(h.last_kernel_ms()["ntt"])
    return statistics.median(ms)
out = []
for k, cols in ((16, 64), (20, 8)):
    n = 1 << k; en = 4 * n
    dom = h.EvaluationDomain(4, k)
    d_in = cols_dev(cols, n); d_out = torch.empty_like(d_in)
    d_ext = torch.empty((cols, en, 4), dtype=torch.int64, device="cuda"); d_back = torch.empty_like(d_ext)
    t1 = timed(lambda: dom.transform_dev(h.OP_LAGRANGE_TO_COEFF, d_in.data_ptr(), n, d_out.data_ptr(), n, cols))
    t2 = timed(lambda: dom.transform_dev(h.OP_COEFF_TO_EXTENDED, d_in.data_ptr(), n, d_ext.data_ptr(), en, cols))
    t3 = timed(lambda: dom.transform_dev(h.OP_EXTENDED_TO_COEFF, d_ext.data_ptr(), en, d_back.data_ptr(), en, cols))
    out.append(f"2^{k} x {cols}: l2c {t1:.3f} ms  c2e {t2:.3f} ms  e2c {t3:.3f} ms")
    dom.close()
print(" | ".join(out))
