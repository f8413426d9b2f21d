import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench, halo2_vectordb_b200 as h
h.init(0); dev = torch.device("cuda", 0)
srs = h.ParamsKZG(bench.K, None, h.synthetic_bases(bench.N, bench.SYN_A, bench.SYN_B))
for bc in (32, 96, 192):   # 0.459 / 0.425 / 0.418 s on 1x B200
    os.environ["H2V_BENCH_BC"] = str(bc)
    print(bc, bench.prove_shaped_resident(h, torch, dev, srs), flush=True)
