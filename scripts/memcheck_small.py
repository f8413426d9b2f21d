"""Small end-to-end pass for compute-sanitizer: raw multiexp, commit (both tables), NTT, one proof."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import halo2_vectordb_b200 as h
from oracle import oracle as O
h.init(0)
n = 700
b = O.gen_bases(1024)
s = O.fr_fill(n, 9, mode=1)
assert (O.g1_to_affine(h.best_multiexp(s, b[:n])) == O.best_multiexp_affine(s, b[:n])).all()
print("raw ok", flush=True)
srs = h.ParamsKZG(10, None, b)
cols = [O.fr_fill(1024, i, mode=i % 2) for i in range(5)]
for table in (-1, 0, 1):
    h.set_tuning(-1, table)
    got = srs.commit_batch(cols)
    for g, c in zip(got, cols):
        assert (g == O.best_multiexp_affine(c, b)).all()
h.set_tuning(-1, -1)
print("commit ok", flush=True)
d = h.EvaluationDomain(4, 10)
od = O.EvaluationDomain(4, 10)
assert (d.coeff_to_extended(cols[0]) == od.coeff_to_extended(cols[0])).all()
print("ntt ok", flush=True)
from toy_circuit import Toy
from oracle import plonk as PL
from common import fr_arr
t = Toy(6, seed=5)
params = PL.Params.setup(6, 12345)
s6 = h.ParamsKZG(6, params.g, params.g_lagrange)
pk = h.ProvingKey(s6, t.cs, [fr_arr(c) for c in t.fixed], [fr_arr(c) for c in t.sigma], fr_arr([t.vk_repr])[0])
proof = pk.create_proof([fr_arr(c) for c in t.advice], [fr_arr(c) for c in t.instances], bytes(32))
assert proof == PL.create_proof(params, t.cs, t.fixed, t.sigma, t.vk_repr, t.advice, t.instances, bytes(32))
print("proof ok", flush=True)
