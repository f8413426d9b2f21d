"""Single-call latencies through the host-facing C ABI (what a per-column drop-in pays)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import halo2_vectordb_b200 as h
from halo2_vectordb_b200.synthetic import uniform_scalars, witness_like
h.init(0)
def med(f, n=15):
    f(); f()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return sorted(ts)[len(ts) // 2] * 1e3
for k in (13, 16, 20):
    n = 1 << k
    srs = h.ParamsKZG(k, None, h.synthetic_bases(n))
    u = torch.from_numpy(uniform_scalars(1, n, 1).view(np.int64)).pin_memory().numpy().view(np.uint64)[0]
    w = torch.from_numpy(witness_like(1, n, min(k - 1, 19), 2).view(np.int64)).pin_memory().numpy().view(np.uint64)[0]
    t_u = med(lambda: srs.commit_lagrange(u)); ms_u = h.last_kernel_ms()
    t_w = med(lambda: srs.commit_lagrange(w))
    cols8 = [u] * 8
    t_8 = med(lambda: srs.commit_batch(cols8))
    d = h.EvaluationDomain(4, k)
    t_l2c = med(lambda: d.lagrange_to_coeff(u))
    t_c2e = med(lambda: d.coeff_to_extended(u))
    print(f"k={k}: commit_lagrange uniform {t_u:.3f} ms, witness {t_w:.3f} ms, batch of 8 {t_8:.3f} ms; "
          f"lagrange_to_coeff {t_l2c:.3f} ms, coeff_to_extended {t_c2e:.3f} ms", flush=True)
    srs.close(); d.close()
print("--- device-resident single column, per kernel class (ms)")
for k in (13, 16, 20):
    n = 1 << k
    srs = h.ParamsKZG(k, None, h.synthetic_bases(n))
    out = torch.zeros((1, 8), dtype=torch.int64, device="cuda")
    for nm, arr in (("uniform", uniform_scalars(1, n, 1)), ("witness", witness_like(1, n, min(k - 1, 19), 2))):
        d = torch.from_numpy(arr.view(np.int64)).cuda()
        for _ in range(3):
            srs.commit_batch_dev(d.data_ptr(), n, 1, n, out.data_ptr())
        ms = h.last_kernel_ms()
        print(k, nm, "sum %.3f" % sum(ms.values()), {a: round(b, 3) for a, b in ms.items() if b}, flush=True)
    srs.close()
print("--- concurrent per-column callers (k=16, uniform), commits per second")
import threading
k = 16; n = 1 << k
srs = h.ParamsKZG(k, None, h.synthetic_bases(n))
cols = [torch.from_numpy(uniform_scalars(1, n, 10 + i).view(np.int64)).pin_memory().numpy().view(np.uint64)[0] for i in range(16)]
for nthreads in (1, 2, 4, 8):
    per = 64 // nthreads
    def work(t):
        for i in range(per):
            srs.commit_lagrange(cols[(t * per + i) % 16])
    ths = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    work(0)
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    print(f"{nthreads} threads: {nthreads * per / dt:.0f} commits/s", flush=True)
srs.close()
