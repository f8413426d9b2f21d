"""BASELINE config 5, batched NTT variants: lagrange_to_coeff / coeff_to_extended / extended_to_coeff at
n = 2^13, 2^16, 2^20 with 8 / 64 / 512 independent columns, device-resident (CUDA events, median of 3)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import halo2_vectordb_b200 as h
h.init(0)
peak = max(h.imad_peak(), h.op_rate(1) * 136)
g = torch.Generator(device="cuda").manual_seed(3)
def cols_dev(c, n):
    a = torch.randint(-(1 << 63), (1 << 63) - 1, (c, n, 4), dtype=torch.int64, generator=g, device="cuda")
    a[..., 3] &= (1 << 60) - 1
    return a
def timed(fn):
    ms = []
    for i in range(5):
        fn()
        if i >= 2: ms.append(h.last_kernel_ms()["ntt"])
    return statistics.median(ms)
print(f"# integer-pipe peak {peak/1e12:.2f} T wide-MAC/s; k | columns | op: ms, Gelem/s (input elements), frac of the integer peak "
      "((N/2) log2 N x 136 wide-MACs per transform of size N, + N x 136 for the scalings), HBM GB/s by 64 B x N x passes")
for k in (13, 16, 20):
    n = 1 << k
    dom = h.EvaluationDomain(4, k)
    ek = dom.extended_k; en = 1 << ek
    for cols in (8, 64, 512):
        if cols * en * 32 > (48 << 30):      # keep the extended buffers under 48 GB
            cols_e = (48 << 30) // (en * 32)
        else:
            cols_e = cols
        d_in = cols_dev(cols, n); d_out = torch.empty_like(d_in)
        t1 = timed(lambda: dom.transform_dev(h.OP_LAGRANGE_TO_COEFF, d_in.data_ptr(), n, d_out.data_ptr(), n, cols))
        d_ext = torch.empty((cols_e, en, 4), dtype=torch.int64, device="cuda")
        t2 = timed(lambda: dom.transform_dev(h.OP_COEFF_TO_EXTENDED, d_in.data_ptr(), n, d_ext.data_ptr(), en, cols_e))
        d_back = torch.empty((cols_e, en, 4), dtype=torch.int64, device="cuda")
        t3 = timed(lambda: dom.transform_dev(h.OP_EXTENDED_TO_COEFF, d_ext.data_ptr(), en, d_back.data_ptr(), en, cols_e))
        def line(name, t, c, N, L):
            macs = c * ((N // 2) * L + N) * 136
            passes = -(-L // 9)
            return (f"{name} {t:8.3f} ms {c * (n if name != 'e2c' else N) / t / 1e6:6.2f} Gelem/s int {macs / (t * 1e-3) / peak:.2f} "
                    f"hbm {c * N * 64 * passes / (t * 1e-3) / 1e9:6.0f} GB/s")
        print(f"2^{k} x {cols:3d}: {line('l2c', t1, cols, n, k)} | x {cols_e:3d}: {line('c2e', t2, cols_e, en, ek)} | {line('e2c', t3, cols_e, en, ek)}", flush=True)
        del d_in, d_out, d_ext, d_back
        torch.cuda.empty_cache()
    dom.close()
