"""Known answers that this repository did not produce (tests/golden/external_vectors.json): EIP-196 ecAdd / ecMul
precompile vectors, the RFC 7539 ChaCha20 key stream, the Poseidon reference implementation's permutation vectors,
halo2curves' documented Fr constants, the EIP-197 G2 generator.  CPU: both oracles (pyref big integers, liboracle C)
and the product's host-side code (Poseidon, ChaCha20).  GPU: the device group law and MSM through the C ABI."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as P
from oracle import transcript as T

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "external_vectors.json")) as f:
    EXT = json.load(f)


def _pt(p):
    return (int(p[0], 16), int(p[1], 16))


def test_eip196_add_oracles():
    for c in EXT["eip196_bn256_add"]["cases"]:
        a, b, s = _pt(c["a"]), _pt(c["b"]), _pt(c["sum"])
        assert P.g1_is_on_curve(a) and P.g1_is_on_curve(b) and P.g1_is_on_curve(s), c["name"]
        assert P.g1_add(a, b) == s, c["name"]
        ja = np.concatenate([O.g1_affine_from_ints(a), O.to_mont(O.ints_to_limbs([1]), O.FQ)[0]])
        assert O.g1_affine_to_ints(O.g1_to_affine(O.g1_add_mixed(ja, O.g1_affine_from_ints(b)))) == s, c["name"]


def test_eip196_mul_oracles():
    for c in EXT["eip196_bn256_scalar_mul"]["cases"]:
        p, k, out = _pt(c["p"]), int(c["k"], 16), _pt(c["out"])
        assert P.g1_mul(p, k) == out, c["name"]            # reduces k mod r, as the precompile does
        assert O.g1_affine_to_ints(O.g1_mul(O.g1_affine_from_ints(p), k % P.R)) == out, c["name"]
        # the same through the oracle's best_multiexp (one term, padded with zero scalars)
        sc = O.fr_from_ints([k % P.R, 0, 0, 0])
        bs = np.stack([O.g1_affine_from_ints(p)] * 4)
        assert O.g1_affine_to_ints(O.best_multiexp_affine(sc, bs)) == out, c["name"]


def test_chacha20_keystream():
    import halo2_vectordb_b200 as h

    v = EXT["chacha20_keystream"]
    key = bytes.fromhex(v["key"])
    for blk in v["blocks"]:
        words = T.chacha20_block([int.from_bytes(key[4 * i:4 * i + 4], "little") for i in range(8)], blk["counter"])
        assert b"".join(w.to_bytes(4, "little") for w in words).hex() == blk["out"]
        assert h.chacha20_block(key, blk["counter"]).hex() == blk["out"]          # product, host side
    # Fr::random = the 512-bit little-endian integer of 64 key-stream bytes, mod r
    stream = bytes.fromhex(v["blocks"][0]["out"]) + bytes.fromhex(v["blocks"][1]["out"])
    want = [int.from_bytes(stream[64 * i:64 * i + 64], "little") % P.R for i in range(2)]
    rng = T.ChaCha20Rng(key)
    assert [rng.fr_random(), rng.fr_random()] == want
    assert O.fr_to_ints(h.chacha20_fr_random(key, 2)) == want


def test_poseidon_reference_vectors():
    import halo2_vectordb_b200 as h

    for c in EXT["poseidon_permutation"]["cases"]:
        want = [int(x, 16) for x in c["output"]]
        assert T.poseidon_permutation(c["input"], c["r_f"], c["r_p"]) == want
        got = h.poseidon_permutation(O.fr_from_ints(c["input"]), c["r_f"], c["r_p"])   # product, host side
        assert O.fr_to_ints(got) == want


def test_halo2curves_constants():
    c = EXT["halo2curves_bn256_fr_constants"]
    assert int(c["modulus"], 16) == P.R and c["s"] == P.S and c["multiplicative_generator"] == P.GEN
    assert int(c["root_of_unity"], 16) == P.ROOT_OF_UNITY == pow(P.GEN, (P.R - 1) >> P.S, P.R)
    assert int(c["delta"], 16) == pow(P.GEN, 1 << P.S, P.R) == O.fr_to_ints(O.fr_delta())[0]
    assert int(c["zeta"], 16) == P.ZETA and pow(P.ZETA, 3, P.R) == 1 and P.ZETA != 1
    d = O.EvaluationDomain(4, 5)
    assert O.fr_to_ints(d.g_coset)[0] == P.ZETA


def test_g2_generator_on_twist():
    from oracle import g2

    g = EXT["eip197_bn256_g2_generator"]
    pt = ((int(g["x_c0"]), int(g["x_c1"])), (int(g["y_c0"]), int(g["y_c1"])))
    assert pt == g2.G2_GEN and g2.is_on_curve(pt)
    assert g2.mul(pt, P.R) is None                       # prime order r


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_group_law_eip196(h2v):
    cases = EXT["eip196_bn256_add"]["cases"]
    p = np.stack([O.g1_affine_from_ints(_pt(c["a"])) for c in cases])
    q = np.stack([O.g1_affine_from_ints(_pt(c["b"])) for c in cases])
    for mode in (0, 1):                                   # XYZZ mixed add, full add
        got = h2v.selftest_group(mode, p, q)
        for i, c in enumerate(cases):
            assert O.g1_affine_to_ints(got[i]) == _pt(c["sum"]), (mode, c["name"])
    got = h2v.selftest_group(2, p[2:3], q[2:3])           # doubling of the generator
    assert O.g1_affine_to_ints(got[0]) == _pt(cases[2]["sum"])


@pytest.mark.gpu
def test_gpu_msm_eip196(h2v):
    for c in EXT["eip196_bn256_scalar_mul"]["cases"]:
        p, k, out = _pt(c["p"]), int(c["k"], 16) % P.R, _pt(c["out"])
        # k * P as an MSM: (k - 5) * P + 2 * P + 3 * P, through best_multiexp and through a commit handle
        sc = O.fr_from_ints([(k - 5) % P.R, 2, 3, 0])
        bs = np.stack([O.g1_affine_from_ints(p)] * 4)
        assert O.g1_affine_to_ints(O.g1_to_affine(h2v.best_multiexp(sc, bs))) == out, c["name"]
        srs = h2v.ParamsKZG(2, bs, None)
        assert O.g1_affine_to_ints(srs.commit(sc)) == out, c["name"]
        srs.close()
