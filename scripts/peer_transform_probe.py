"""Device-resident transforms split over the devices of h2v_init (h2v_domain_transform_dev): wall time of lagrange_to_coeff
and coeff_to_extended batches at k = 16 and k = 20 on 1 device and on all of them.  Run with gpurun --gpus 2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import halo2_vectordb_b200 as h
ndev = h.device_count()
for k, cols in ((16, 512), (20, 96)):
    n = 1 << k
    for devs in ([0], list(range(ndev))):
        h.init(devs if len(devs) > 1 else 0)
        dom = h.EvaluationDomain(4, k)
        torch.cuda.set_device(0)
        g = torch.Generator(device="cuda:0").manual_seed(1)
        a = torch.randint(-(1 << 63), (1 << 63) - 1, (cols, n, 4), dtype=torch.int64, generator=g, device="cuda:0")
        a[..., 3] &= (1 << 60) - 1
        o1 = torch.empty_like(a)
        o2 = torch.empty((cols, 4 * n, 4), dtype=torch.int64, device="cuda:0")
        for name, op, dst, stride in (("l2c", h.OP_LAGRANGE_TO_COEFF, o1, n), ("c2e", h.OP_COEFF_TO_EXTENDED, o2, 4 * n)):
            ts = []
            for _ in range(4):
                torch.cuda.synchronize(0)
                t0 = time.perf_counter()
                dom.transform_dev(op, a.data_ptr(), n, dst.data_ptr(), stride, cols)
                ts.append((time.perf_counter() - t0) * 1e3)
            print(f"k={k} cols={cols} devices={len(devs)} {name}: " + " ".join(f"{t:8.2f}" for t in ts) + " ms", flush=True)
        dom.close()
        del a, o1, o2
        torch.cuda.empty_cache()
