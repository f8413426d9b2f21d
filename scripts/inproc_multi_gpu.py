"""In-process multi-GPU strong scaling (one process, h2v_init(devices)): ONE fixed phase of the kmeans k = 16 proof --
commit_lagrange of 1 150 host columns of 2^16, then lagrange_to_coeff of the same columns -- over 1, 2, 4, 8 devices.
Run with gpurun --gpus 8."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import halo2_vectordb_b200 as h
from halo2_vectordb_b200.synthetic import uniform_scalars, witness_like

K, COLS, BASE = 16, 1150, 96
N = 1 << K
ndev = h.device_count()
h.init(0)
bases = h.synthetic_bases(N)
hu = torch.from_numpy(uniform_scalars(BASE, N, 1).view(np.int64)).pin_memory().numpy().view(np.uint64)
hw = torch.from_numpy(witness_like(BASE, N, 15, 2).view(np.int64)).pin_memory().numpy().view(np.uint64)
outc = torch.empty((BASE, N, 4), dtype=torch.int64).pin_memory().numpy().view(np.uint64)
res = {}
for G in (1, 2, 4, 8):
    if G > ndev:
        break
    h.init(list(range(G)))
    srs = h.ParamsKZG(K, None, bases)
    dom = h.EvaluationDomain(4, K)
    out = np.zeros((COLS, 8), dtype=np.uint64)
    for name, host in (("uniform", hu), ("witness", hw)):
        ptrs = (C.c_void_p * COLS)(*[host[j % BASE].ctypes.data for j in range(COLS)])
        ts = []
        for it in range(4):
            t0 = time.perf_counter()
            h._check(h.lib().h2v_commit_batch(srs._h, 1, ptrs, COLS, N, out.ctypes.data_as(C.c_void_p)))
            ts.append(time.perf_counter() - t0)
        t = min(ts[1:])
        res[(name, G)] = t
        base = res[(name, 1)]
        print(f"commit_lagrange {COLS} x 2^{K} {name:8s} G={G}: {t*1e3:8.2f} ms  {COLS*N/t/1e6:8.1f} Mpts/s  speed-up {base/t:5.2f}  efficiency {base/t/G:4.2f}", flush=True)
        if G > 1:      # the split must not change results
            h.init(0); s1 = h.ParamsKZG(K, None, bases); o1 = np.zeros((8, 8), dtype=np.uint64)
            p8 = (C.c_void_p * 8)(*[host[j % BASE].ctypes.data for j in range(8)])
            h._check(h.lib().h2v_commit_batch(s1._h, 1, p8, 8, N, o1.ctypes.data_as(C.c_void_p)))
            assert (o1 == out[:8]).all(), "multi-device result differs"
            s1.close(); h.init(list(range(G)))
    ia = (C.c_void_p * COLS)(*[hu[j % BASE].ctypes.data for j in range(COLS)])
    oa = (C.c_void_p * COLS)(*[outc[j % BASE].ctypes.data for j in range(COLS)])
    ts = []
    for it in range(3):
        t0 = time.perf_counter()
        h._check(h.lib().h2v_domain_transform_batch(dom._h, h.OP_LAGRANGE_TO_COEFF, ia, oa, COLS))
        ts.append(time.perf_counter() - t0)
    t = min(ts[1:])
    res[("l2c", G)] = t
    print(f"lagrange_to_coeff {COLS} x 2^{K} host in/out   G={G}: {t*1e3:8.2f} ms  {COLS*N/t/1e9:6.2f} Gelem/s  speed-up {res[('l2c',1)]/t:5.2f}", flush=True)
    srs.close(); dom.close()
