import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import halo2_vectordb_b200 as h
from halo2_vectordb_b200.synthetic import uniform_scalars
h.init(0)
k = 16; N = 1 << k; cols = 96
srs = h.ParamsKZG(k, None, h.synthetic_bases(N))
host = torch.from_numpy(uniform_scalars(cols, N, 1).view(np.int64)).pin_memory()
ptrs = (C.c_void_p * cols)(*[host[i].data_ptr() for i in range(cols)])
out = np.zeros((cols, 8), dtype=np.uint64)
ts = []
for i in range(8):
    t0 = time.perf_counter()
    h._check(h.lib().h2v_commit_batch(srs._h, 1, ptrs, cols, N, out.ctypes.data_as(C.c_void_p)))
    ts.append(time.perf_counter() - t0)
t = sorted(ts[2:])[len(ts[2:]) // 2]
print(os.environ.get("H2V_FIRST_MB", "32"), os.environ.get("H2V_SUB_MB", "160"), "MiB first/sub: %.2f ms -> %.1f Mpts/s" % (t * 1e3, cols * N / t / 1e6))
