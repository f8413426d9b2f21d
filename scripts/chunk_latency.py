import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import halo2_vectordb_b200 as h
from halo2_vectordb_b200.synthetic import uniform_scalars, witness_like
h.init(0)
for k in (13, 16):
    n = 1 << k
    srs = h.ParamsKZG(k, None, h.synthetic_bases(n))
    out = torch.zeros((8, 8), dtype=torch.int64, device="cuda")
    for nm, arr in (("uniform", uniform_scalars(8, n, 1)), ("witness", witness_like(8, n, min(k - 1, 19), 2))):
        d = torch.from_numpy(arr.view(np.int64)).cuda()
        for cols in (1, 8):
            res = []
            for chunk in (4, 8, 12, 16, 24, 32):
                h.set_tuning(chunk, -1)
                for _ in range(3):
                    srs.commit_batch_dev(d.data_ptr(), n, cols, n, out.data_ptr())
                ms = h.last_kernel_ms()
                res.append("%d:%.3f(acc %.3f fin %.3f)" % (chunk, sum(ms.values()), ms["msm_accumulate"], ms["msm_finish"]))
            print(k, nm, cols, "cols:", " ".join(res), flush=True)
    h.set_tuning(-1, -1)
    srs.close()
