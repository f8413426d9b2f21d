"""permute_expression_pair timing: U usable rows of a range lookup, device-resident."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import halo2_vectordb_b200 as h
from halo2_vectordb_b200.synthetic import to_mont
h.init(0)
for k, bits in ((13, 12), (16, 15), (20, 19)):
    u = (1 << k) - 6
    rng = np.random.default_rng(k)
    canon = lambda v: np.stack([v, np.zeros_like(v), np.zeros_like(v), np.zeros_like(v)], axis=1)
    fi = to_mont(canon(rng.integers(0, 1 << bits, u, dtype=np.uint64)))
    ft = to_mont(canon(np.concatenate([np.arange(1 << bits, dtype=np.uint64), np.zeros(u - (1 << bits), dtype=np.uint64)])))
    bufs = [h.DeviceBuffer(u * 32) for _ in range(4)]
    bufs[0].upload(fi); bufs[1].upload(ft)
    ms = []
    for i in range(6):
        h.permute_expression_pair_dev(bufs[0].ptr, bufs[1].ptr, u, bufs[2].ptr, bufs[3].ptr)
        if i >= 2: ms.append(h.last_kernel_ms()["ntt"])
    print(f"k={k}: {u} rows, {statistics.median(ms):.3f} ms")
