"""BASELINE.json configs[3]: a synthetic SIFT-shaped `nearest_vector` + Poseidon `merkle_commitment` over 1024 random
128-dimensional fixed-point vectors at k = 20 (LOOKUP_BITS = 19) -- ONE real proof, end to end:

  witness   the restated chips (halo2_vectordb_b200.circuit.exhaustive_merkle = /root/reference/examples/query.rs:32-73)
  columns   the restated halo2-base layouter (advice / fixed / sigma / instance)
  proof     h2v_create_proof on 1 or, in process, several GPUs (h2v_init(devices): commit / transform batches split by column)
  check     the restated verifier (oracle/plonk.py: plonk/verifier.rs + SHPLONK, pairing replaced by the known setup secret);
            a tampered public input must be rejected; the N-GPU proof must equal the 1-GPU proof byte for byte

Run under gpurun:  python scripts/sift_k20_proof.py [--vectors 1024 --dim 128 --k 20 --bits 19 --gpus 1]
The oracle is used as the checker only.  Prints one JSON line."""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SECRET = 0x1CE1CEBABE5EED0123456789ABCDEF0FEDCBA9876543210
SEED = bytes(range(32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vectors", type=int, default=1024)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--bits", type=int, default=19)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()

    import psutil
    import torch

    import halo2_vectordb_b200 as h
    from halo2_vectordb_b200 import circuit as Z
    from oracle import oracle as O
    from oracle import plonk as PL
    from oracle import pyref as P

    avail = psutil.virtual_memory().available / 2**30
    need = 0.25 * args.vectors * args.dim / 1024      # ~ GiB of host memory for the trace + the laid-out columns (measured at 128 x 128)
    out = {"config": f"nearest_vector + merkle_commitment, {args.vectors} x {args.dim}, k={args.k}, LOOKUP_BITS={args.bits}",
           "host": {"cores": os.cpu_count(), "mem_available_gib": round(avail, 1)}}
    if avail < need + 8:
        out["skipped"] = f"host memory: {avail:.0f} GiB available, about {need + 8:.0f} GiB needed"
        print(json.dumps(out))
        return
    k, n = args.k, 1 << args.k
    rnd = random.Random(20)
    inp = dict(query=[rnd.uniform(0.0, 2.0) for _ in range(args.dim)],
               database=[[rnd.uniform(0.0, 2.0) for _ in range(args.dim)] for _ in range(args.vectors)])
    # the trace has about 2 200 cells per vector coordinate: reserve its address space once (h2v_builder_new reads this)
    os.environ.setdefault("H2V_TRACE_RESERVE", str(int(args.vectors * args.dim * 2200)))
    t0 = time.perf_counter()
    builder = Z.GateThreadBuilder(args.bits)
    public = []
    Z.exhaustive_merkle(builder.main(0), inp, public)
    builder.make_public(public)
    t_wit = time.perf_counter() - t0
    st = builder.stats()
    t0 = time.perf_counter()
    rc = Z.RangeCircuit(builder, k)
    t_lay = time.perf_counter() - t0
    A = len(rc.advice)
    out.update({"advice_cells": st["advice_cells"], "lookup_cells": st["lookup_cells"], "advice_columns": rc.num_advice,
                "lookup_advice_columns": rc.num_lookup_advice, "fixed_columns": len(rc.fixed),
                "permutation_columns": len(rc.cs["permutation"]), "public_inputs": rc.num_instances,
                "host_s": {"witness_generation": round(t_wit, 2), "layout_and_sigma": round(t_lay, 2)}})
    print(json.dumps(out), file=sys.stderr, flush=True)

    pinned = torch.empty((A, n, 4), dtype=torch.int64).pin_memory()
    adv = pinned.numpy().view(np.uint64)
    for i, c in enumerate(rc.advice):
        adv[i] = c
    cols = [adv[i] for i in range(A)]
    s = O.fr_from_ints([SECRET])[0]
    vk_repr = O.fr_from_ints([0x5eed])[0]
    ref_proof = None
    runs = {}
    for devs in ([[0]] if args.gpus == 1 else [[0], list(range(args.gpus))]):
        h.init(devs if len(devs) > 1 else devs[0])
        t0 = time.perf_counter()
        g, gl = h.srs_setup(k, s)
        srs = h.ParamsKZG(k, g, gl)
        t_srs = time.perf_counter() - t0
        t0 = time.perf_counter()
        pk = h.ProvingKey(srs, rc.cs, rc.fixed, rc.sigma, vk_repr)
        t_pk = time.perf_counter() - t0
        ts = []
        proof = None
        for _ in range(args.reps):
            t0 = time.perf_counter()
            p2 = pk.create_proof(cols, rc.instances, SEED)
            ts.append(time.perf_counter() - t0)
            assert proof is None or p2 == proof, "create_proof is not deterministic in the seed"
            proof = p2
        run = {"prove_s": round(min(ts), 4), "prove_s_all": [round(t, 4) for t in ts], "proof_bytes": len(proof),
               "phase_ms": {kk: round(v, 1) for kk, v in pk.last_phase_ms().items()},
               "setup_s": {"srs_setup": round(t_srs, 2), "pk_load": round(t_pk, 2)},
               "hbm_used_gib": round((torch.cuda.mem_get_info(0)[1] - torch.cuda.mem_get_info(0)[0]) / 2**30, 1)}
        if ref_proof is None:
            ref_proof = proof
            if not args.no_verify:
                t0 = time.perf_counter()
                vk = {"fixed": [O.g1_affine_to_ints(p) for p in srs.commit_batch(rc.fixed)],
                      "sigma": [O.g1_affine_to_ints(p) for p in srs.commit_batch(rc.sigma)]}
                params = PL.Params(k, None, None, SECRET)
                inst = [O.fr_to_ints(np.ascontiguousarray(c)) for c in rc.instances]
                run["verifier_accepts"] = bool(PL.verify_proof(params, rc.cs, vk, 0x5eed, inst, proof))
                bad = [[(inst[0][0] + 1) % P.R] + inst[0][1:]] + inst[1:]
                run["verifier_rejects_tampered_instance"] = not PL.verify_proof(params, rc.cs, vk, 0x5eed, bad, proof)
                run["verify_s"] = round(time.perf_counter() - t0, 2)
        else:
            run["same_bytes_as_1_gpu"] = proof == ref_proof
        runs[f"{len(devs)}_gpu"] = run
        pk.close()
        srs.close()
        del g, gl
    out["runs"] = runs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
