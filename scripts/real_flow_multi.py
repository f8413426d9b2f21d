"""One real proof of the kmeans example (BASELINE configs[2]: k = 16, LOOKUP_BITS = 15, data/kmeans.in) with the commit
phases of create_proof spread over G = 1, 2, 4, 8 devices in process (h2v_init(devices)): wall-clock per proof, per phase.
The proof bytes must not depend on G."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import halo2_vectordb_b200 as h
from halo2_vectordb_b200 import circuit as Z

name = os.environ.get("EXAMPLE", "kmeans")
k, bits = Z.EXAMPLE_PARAMS[name]
n = 1 << k
builder = Z.GateThreadBuilder(bits)
pub = []
Z.EXAMPLES[name](builder.main(0), Z.example_input(name), pub)
builder.make_public(pub)
rc = Z.RangeCircuit(builder, k)
A = len(rc.advice)
pinned = torch.empty((A, n, 4), dtype=torch.int64).pin_memory()
adv = pinned.numpy().view(np.uint64)
for i, c in enumerate(rc.advice):
    adv[i] = c
cols = [adv[i] for i in range(A)]
ref = None
ndev = h.device_count()
for G in (1, 2, 4, 8):
    if G > ndev:
        break
    h.init(list(range(G)))
    srs = h.ParamsKZG.gen_srs(k)
    pk = h.ProvingKey(srs, rc.cs, rc.fixed, rc.sigma, np.array([k, 0, 0, 0], dtype=np.uint64))
    proof = pk.create_proof(cols, rc.instances, bytes(32))
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        p2 = pk.create_proof(cols, rc.instances, bytes(32))
        ts.append(time.perf_counter() - t0)
        assert p2 == proof
    if ref is None:
        ref = proof
    assert proof == ref, "proof bytes depend on the device count"
    print(json.dumps({"example": name, "k": k, "devices": G, "prove_s": round(min(ts), 4),
                      "phase_ms": {kk: round(v, 1) for kk, v in pk.last_phase_ms().items()}, "proof_bytes": len(proof)}), flush=True)
    pk.close()
    srs.close()
h.init(0)
