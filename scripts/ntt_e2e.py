import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import halo2_vectordb_b200 as h
from halo2_vectordb_b200.synthetic import uniform_scalars
h.init(0)
k = 16; N = 1 << k
dom = h.EvaluationDomain(4, k)
for cols in (1, 4, 16, 32, 33, 64):
    hin = torch.from_numpy(uniform_scalars(cols, N, 1).view(np.int64)).pin_memory()
    hout = torch.empty((cols, N, 4), dtype=torch.int64).pin_memory()
    ia = (C.c_void_p * cols)(*[hin[i].data_ptr() for i in range(cols)])
    oa = (C.c_void_p * cols)(*[hout[i].data_ptr() for i in range(cols)])
    ts = []
    for i in range(6):
        t0 = time.perf_counter()
        h._check(h.lib().h2v_domain_transform_batch(dom._h, h.OP_LAGRANGE_TO_COEFF, ia, oa, cols))
        ts.append(time.perf_counter() - t0)
    print(cols, "cols l2c e2e ms:", [round(t * 1e3, 2) for t in ts], flush=True)
