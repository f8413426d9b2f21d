"""BASELINE config 5: standalone BN254 G1 MSM and Fr NTT sweep 2^16 .. 2^24 on one GPU (single transform /
single MSM, device-resident; plus best_multiexp through the host ABI with arbitrary bases)."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import halo2_vectordb_b200 as h
from halo2_vectordb_b200.synthetic import uniform_scalars, witness_like
h.init(0)
peak = max(h.imad_peak(), h.op_rate(1) * 136)
WMAC = {16: 27200, 17: 27200, 18: 23120, 19: 21760, 20: 20400, 21: 20400, 22: 20400, 23: 17680, 24: 17680}
print(f"# integer-pipe peak {peak/1e12:.2f} T wide-MAC/s; columns: lg n | MSM ms uniform (Mpts/s, frac of peak by SURVEY 8d work) | "
      f"MSM ms witness | best_multiexp e2e ms (raw bases) | NTT ms (Gelem/s, int frac) | coset NTT n->4n ms")
for lg in range(16, 25):
    n = 1 << lg
    bases = h.synthetic_bases(n)
    srs = h.ParamsKZG(lg, None, bases)
    out = torch.zeros((1, 8), dtype=torch.int64, device="cuda")
    res = []
    for arr in (uniform_scalars(1, n, lg), witness_like(1, n, min(lg - 1, 20), lg)):
        d = torch.from_numpy(arr.view(np.int64)).cuda()
        ms = []
        for i in range(5):
            srs.commit_batch_dev(d.data_ptr(), n, 1, n, out.data_ptr())
            if i >= 2: ms.append(sum(h.last_kernel_ms().values()))
        res.append(statistics.median(ms))
    srs.close()
    u = uniform_scalars(1, n, lg)[0]
    t0 = time.perf_counter(); h.best_multiexp(u, bases); t_raw = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); h.best_multiexp(u, bases); t_raw = min(t_raw, (time.perf_counter() - t0) * 1e3)
    del bases
    dom = h.EvaluationDomain(4, lg) if lg <= 24 else None
    d_in = torch.from_numpy(uniform_scalars(1, n, 7).view(np.int64)).cuda()
    d_o = torch.empty_like(d_in)
    ms = []
    for i in range(5):
        dom.transform_dev(h.OP_COEFF_TO_LAGRANGE, d_in.data_ptr(), n, d_o.data_ptr(), n, 1)
        if i >= 2: ms.append(h.last_kernel_ms()["ntt"])
    t_ntt = statistics.median(ms)
    t_c2e = float("nan")
    if lg <= 22:
        d_e = torch.empty((1, 4 * n, 4), dtype=torch.int64, device="cuda")
        ms = []
        for i in range(5):
            dom.transform_dev(h.OP_COEFF_TO_EXTENDED, d_in.data_ptr(), n, d_e.data_ptr(), 4 * n, 1)
            if i >= 2: ms.append(h.last_kernel_ms()["ntt"])
        t_c2e = statistics.median(ms)
        del d_e
    dom.close()
    msm_frac = n * WMAC[lg] / (res[0] * 1e-3) / peak
    ntt_frac = (n // 2) * lg * 136 / (t_ntt * 1e-3) / peak
    print(f"2^{lg}: MSM {res[0]:8.3f} ms ({n/res[0]/1e3:6.1f} Mpts/s, {msm_frac:.2f}) | witness {res[1]:8.3f} ms | raw e2e {t_raw:8.2f} ms | "
          f"NTT {t_ntt:7.3f} ms ({n/t_ntt/1e6:5.2f} Gelem/s, {ntt_frac:.2f}) | coset {t_c2e:7.3f} ms", flush=True)
    torch.cuda.empty_cache()
