"""ONE standalone multiexp of 2^LOG_N points split over the GPUs by index range:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29517 scripts/sliced_msm.py [LOG_N ...]
Every rank holds bases[lo:hi] (+ window tables) and the matching scalars; the only exchange is an all-gather of the
G affine partial sums (64 B each, NCCL) followed by h2v_g1_sum.  Timed on the device, max over ranks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import halo2_vectordb_b200 as h
from halo2_vectordb_b200 import sharding, synthetic

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
h.init(local)
sizes = [int(a) for a in sys.argv[1:]] or [20, 22, 24]
for L in sizes:
    n = 1 << L
    lo, hi = sharding.index_slice(n, rank, world)
    m = hi - lo
    bases = np.ascontiguousarray(h.synthetic_bases(n)[lo:hi])      # B_i = (a i + b) G for the global index i
    srs = h.ParamsKZG(L - (world.bit_length() - 1), None, bases)
    s_all = synthetic.uniform_scalars(1, n, seed=L)[0]            # every rank draws the same column, keeps its slice
    d_s = torch.from_numpy(s_all[lo:hi].view(np.int64)).to(dev)
    d_o = torch.zeros(8, dtype=torch.int64, device=dev)
    ms = []
    for it in range(5):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        srs.commit_batch_dev(d_s.data_ptr(), m, 1, m, d_o.data_ptr())
        parts = sharding.gather_partials(d_o.cpu().numpy().view(np.uint64), rank, world, dev)
        total = h.g1_sum(parts)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it >= 2:
            ms.append(dt.item() * 1e3)
    if rank == 0:
        from oracle import oracle as O        # checker only: closed form of the WHOLE sum
        ok = bool((total == O.msm_closed_form(s_all, h.SYN_A, h.SYN_B)).all())
        t = sorted(ms)[len(ms) // 2]
        print(f"2^{L} over {world} GPU(s): {t:8.3f} ms  {n / t / 1e3:8.1f} Mpts/s  window c={srs.info()[0]}  closed-form check {'ok' if ok else 'MISMATCH'}", flush=True)
    srs.close()
if world > 1:
    dist.destroy_process_group()
