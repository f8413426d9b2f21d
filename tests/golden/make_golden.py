"""Generate tests/golden/bn254_golden.json from first principles (oracle/pyref.py: Python big ints,
naive DFT, double-and-add).  The reference itself cannot run here (Rust, un-vendored deps), and it
ships no vectors for this boundary, so these are *derived* known answers -- they pin the C oracle and
the CUDA path to the mathematics, and include the SURVEY.md App. B values as literal cross-checks.

    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyref as P  # noqa: E402

rnd = random.Random(0x5EED0001)
hx = lambda v: hex(v)
pt = lambda p: None if p is None else [hex(p[0]), hex(p[1])]
out = {}

out["scalar_mul"] = [{"k": hx(k), "point": pt(P.g1_mul(P.G1_GEN, k))} for k in
                     [1, 2, 3, 30, 0xDEADBEEF, P.R - 1, P.R, rnd.randrange(P.R), rnd.randrange(P.R)]]

msms = []
for n in (1, 2, 5, 17, 64):
    ks = [rnd.randrange(1, P.R) for _ in range(n)]
    bases = [P.g1_mul(P.G1_GEN, k) for k in ks]
    for dist in ("uniform", "small", "edge"):
        if dist == "uniform":
            s = [rnd.randrange(P.R) for _ in range(n)]
        elif dist == "small":
            s = [rnd.choice([0, 1, 2, rnd.randrange(4096)]) for _ in range(n)]
        else:
            s = [rnd.choice([0, P.R - 1, P.R - 2, 1, (1 << 253)]) for _ in range(n)]
        msms.append({"scalars": [hx(x) for x in s], "bases": [pt(b) for b in bases], "dist": dist,
                     "result": pt(P.g1_msm(s, bases))})
# duplicate / opposite / identity bases: exercises P+P, P+(-P) and identity handling in every add formula
g5 = P.g1_mul(P.G1_GEN, 5)
special = [g5, g5, P.g1_neg(g5), None, g5, P.g1_mul(P.G1_GEN, 10), P.g1_neg(P.g1_mul(P.G1_GEN, 10)), g5]
s = [1, 1, 1, 7, 2, 1, 1, P.R - 4]
msms.append({"scalars": [hx(x) for x in s], "bases": [pt(b) for b in special], "dist": "special",
             "result": pt(P.g1_msm(s, special))})
out["msm"] = msms

ffts = []
for log_n in (0, 1, 2, 3, 4, 6):
    a = [rnd.randrange(P.R) for _ in range(1 << log_n)]
    w = P.omega_for(log_n)
    ffts.append({"log_n": log_n, "omega": hx(w), "a": [hx(x) for x in a], "out": [hx(x) for x in P.dft_naive(a, w)]})
ffts.append({"log_n": 2, "omega": hx(P.omega_for(2)), "a": ["0x1", "0x2", "0x3", "0x4"],
             "out": [hx(x) for x in P.dft_naive([1, 2, 3, 4], P.omega_for(2))]})
out["fft"] = ffts

doms = []
for j, k in ((4, 3), (4, 5), (3, 4), (5, 3), (2, 4)):
    d = P.Domain(j, k)
    a = [rnd.randrange(P.R) for _ in range(d.n)]
    ext = d.coeff_to_extended(a)
    h = [rnd.randrange(P.R) for _ in range(1 << d.extended_k)]
    doms.append({
        "j": j, "k": k, "extended_k": d.extended_k,
        "omega": hx(d.omega), "omega_inv": hx(d.omega_inv), "extended_omega": hx(d.omega_ext),
        "extended_omega_inv": hx(d.omega_ext_inv), "g_coset": hx(d.g_coset), "g_coset_inv": hx(d.g_coset_inv),
        "ifft_divisor": hx(d.ifft_divisor), "extended_ifft_divisor": hx(d.extended_ifft_divisor),
        "t_evaluations": [hx(x) for x in d.t_evaluations],
        "a": [hx(x) for x in a],
        "lagrange_to_coeff": [hx(x) for x in d.lagrange_to_coeff(a)],
        "coeff_to_lagrange": [hx(x) for x in d.coeff_to_lagrange(a)],
        "coeff_to_extended": [hx(x) for x in ext],
        "h": [hx(x) for x in h],
        "divide_by_vanishing_poly": [hx(x) for x in d.divide_by_vanishing_poly(h)],
        "extended_to_coeff": [hx(x) for x in d.extended_to_coeff(h)],
    })
out["domain"] = doms
out["omega_by_k"] = {str(k): hx(P.omega_for(k)) for k in (13, 16, 20)}

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bn254_golden.json")
with open(path, "w") as f:
    json.dump(out, f, indent=0)
print(path, os.path.getsize(path), "bytes")
