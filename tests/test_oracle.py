"""CPU: pin the oracle.  (1) SURVEY.md App. B constants / known answers as literals, (2) the C oracle
against the first-principles Python model and the committed golden vectors, (3) algorithm-independent
properties at larger sizes (closed-form MSM, Horner spot checks, round trips)."""
import random

import numpy as np
import pytest

from common import fr_arr, g1_arr, golden, ipt, ival
from oracle import oracle as O
from oracle import pyref as P


def test_constants_app_b():
    assert P.P == 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
    assert P.R == 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    assert pow(2, 256, P.R) == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert pow(2, 512, P.R) == 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7
    assert (-pow(P.R, -1, 1 << 64)) % (1 << 64) == 0xC2E1F593EFFFFFFF
    assert (-pow(P.P, -1, 1 << 64)) % (1 << 64) == 0x87D20782E4866389
    assert pow(2, 256, P.P) == 0x0E0A77C19A07DF2F666EA36F7879462C0A78EB28F5C70B3DD35D438DC58F0D9D
    assert P.ROOT_OF_UNITY == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert pow(P.ROOT_OF_UNITY, 1 << 28, P.R) == 1 and pow(P.ROOT_OF_UNITY, 1 << 27, P.R) == P.R - 1
    assert pow(P.ZETA, 3, P.R) == 1 and P.ZETA != 1
    assert P.omega_for(13) == 0x10E3D295C1599FF535A1BB49F23D81AA03BD0ED25881F9ED12B179AF67F67AE1
    assert P.omega_for(16) == 0x09D2CC4B5782FBE923E49ACE3F647643A5F5D8FB89091C3ABABD582133584B29
    assert P.omega_for(20) == 0x2A14464F1FF42DE3856402B62520E670745E39FADA049D5B2F0E1E3182673378


def test_curve_kats_app_b():
    G = P.G1_GEN
    assert P.g1_mul(G, 2) == (0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3,
                              0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4)
    assert P.g1_mul(G, 3) == (0x0769BF9AC56BEA3FF40232BCB1B6BD159315D84715B8E679F2D355961915ABF0,
                              0x2AB799BEE0489429554FDB7C8D086475319E63B40B9C5B57CDF1FF3DD9FE2261)
    thirty = (0x036083BFA420B15A4C11F66A3CFFD55318B019FEB45F833A876E93848625F5AE,
              0x2630C348C019C3EDB74FE62A7E921361AAE9621988223514D56CA8B36ADC9E36)
    assert P.g1_mul(G, 30) == thirty
    assert P.g1_msm([1, 2, 3, 4], [P.g1_mul(G, i) for i in (1, 2, 3, 4)]) == thirty   # toy MSM KAT
    assert P.g1_mul(G, P.R - 1) == (1, P.P - 2) and P.g1_mul(G, P.R) is None
    assert P.g1_mul(G, 0xDEADBEEF) == (0x1FD9BF9C6C9FC892F0B4F856657CD9309F43E2F1CFA3ED4724C40BD74EA13803,
                                       0x18EE06DE0E49DEAF292D55F31FD13E603489F81BFA4EC6F2443BA2274621703F)
    # NTT toy KAT: DFT_4([1,2,3,4])
    out = P.dft_naive([1, 2, 3, 4], P.omega_for(2))
    assert out[0] == 10 and out[2] == P.R - 2
    assert out[1] == 0x00000000000000016789AF3A83522EB1969386A2F88C094A419FE246C11F9394
    assert out[3] == 0x30644E72E131A02850C6967BFE2F29AB91A061A5812D67470242134D2EE06C69


def test_oracle_field_ops():
    rnd = random.Random(3)
    for field, mod, nm in ((O.FR, P.R, "fr"), (O.FQ, P.P, "fq")):
        for t in range(100):
            a, b = rnd.randrange(mod), rnd.randrange(mod)
            if t == 0:
                a = b = mod - 1
            am, bm = O.to_mont(O.ints_to_limbs([a]), field)[0], O.to_mont(O.ints_to_limbs([b]), field)[0]
            f = lambda x: O.limbs_to_ints(O.from_mont(x, field))[0]
            assert f(O.field_op(nm + "_mul", am, bm)) == a * b % mod
            assert f(O.field_op(nm + "_add", am, bm)) == (a + b) % mod
            assert f(O.field_op(nm + "_sub", am, bm)) == (a - b) % mod
            if a:
                assert f(O.field_op(nm + "_inv", am)) == pow(a, -1, mod)
        assert O.limbs_to_ints(O.to_mont(O.ints_to_limbs([1]), field))[0] == pow(2, 256, mod)


def test_oracle_scalar_mul_golden():
    G = O.g1_generator()
    assert O.g1_affine_to_ints(G) == (1, 2)
    for v in golden()["scalar_mul"]:
        got = O.g1_affine_to_ints(O.g1_mul(G, ival(v["k"])))
        assert got == ipt(v["point"])


def test_oracle_compress():
    G = O.g1_generator()
    for k in (1, 2, 3, 0xDEADBEEF):
        aff = O.g1_mul(G, k)
        assert O.g1_compress(aff) == P.g1_compress(O.g1_affine_to_ints(aff))
    assert O.g1_compress(np.zeros(8, dtype=np.uint64)) == bytes(32)


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_oracle_multiexp_golden(threads):
    for v in golden()["msm"]:
        s = fr_arr([ival(x) for x in v["scalars"]])
        b = g1_arr([ipt(p) for p in v["bases"]])
        got = O.g1_affine_to_ints(O.best_multiexp_affine(s, b, threads))
        assert got == ipt(v["result"]), v["dist"]


def test_oracle_multiexp_empty():
    out = O.best_multiexp(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 8), dtype=np.uint64), 4)
    assert O.g1_affine_to_ints(O.g1_to_affine(out)) is None


def test_oracle_multiexp_closed_form():
    for n, mode in ((1000, 0), (4096, 1), (1 << 13, 0)):
        b = O.gen_bases(n)
        assert all(O.g1_is_on_curve(b[i]) for i in (0, 1, n // 2, n - 1))
        s = O.fr_fill(n, 77 + n, mode=mode)
        assert (O.best_multiexp_affine(s, b) == O.msm_closed_form(s)).all()


def test_oracle_fft_golden():
    for v in golden()["fft"]:
        a = fr_arr([ival(x) for x in v["a"]])
        w = fr_arr([ival(v["omega"])])[0]
        for threads in (1, 4, 16):
            got = O.fr_to_ints(O.best_fft(a, w, v["log_n"], threads))
            assert got == [ival(x) for x in v["out"]]


def test_oracle_fft_properties_large():
    rnd = random.Random(9)
    for L in (10, 14):
        a = O.fr_fill(1 << L, 5 + L)
        w = fr_arr([P.omega_for(L)])[0]
        out = O.best_fft(a, w, L)
        for j in rnd.sample(range(1 << L), 4):
            x = fr_arr([pow(P.omega_for(L), j, P.R)])[0]
            assert (O.fr_eval_poly(a, x) == out[j]).all()
        # serial and threaded paths agree
        assert (O.best_fft(a, w, L, 1) == out).all()


def test_oracle_domain_golden():
    for v in golden()["domain"]:
        d = O.EvaluationDomain(v["j"], v["k"])
        assert d.extended_k == v["extended_k"]
        for nm in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv",
                   "ifft_divisor", "extended_ifft_divisor"):
            assert O.fr_to_ints(getattr(d, nm))[0] == ival(v[nm]), nm
        assert [O.fr_to_ints(t)[0] for t in d.t_evaluations] == [ival(x) for x in v["t_evaluations"]]
        a = fr_arr([ival(x) for x in v["a"]])
        h = fr_arr([ival(x) for x in v["h"]])
        assert O.fr_to_ints(d.lagrange_to_coeff(a)) == [ival(x) for x in v["lagrange_to_coeff"]]
        assert O.fr_to_ints(d.coeff_to_lagrange(a)) == [ival(x) for x in v["coeff_to_lagrange"]]
        assert O.fr_to_ints(d.coeff_to_extended(a)) == [ival(x) for x in v["coeff_to_extended"]]
        assert O.fr_to_ints(d.divide_by_vanishing_poly(h)) == [ival(x) for x in v["divide_by_vanishing_poly"]]
        assert O.fr_to_ints(d.extended_to_coeff(h)) == [ival(x) for x in v["extended_to_coeff"]]


def test_oracle_domain_roundtrip_k13():
    d = O.EvaluationDomain(4, 13)
    assert O.fr_to_ints(d.omega)[0] == P.omega_for(13)
    a = O.fr_fill(d.n, 41, mode=1)
    assert (d.coeff_to_lagrange(d.lagrange_to_coeff(a)) == a).all()
    back = d.extended_to_coeff(d.coeff_to_extended(a))
    assert (back[: d.n] == a).all() and not back[d.n:].any()


def test_oracle_row2_primitives():
    """batch_invert / grand product / kate_division restatements against first principles."""
    rnd = random.Random(21)
    n = 70
    a = [rnd.randrange(P.R) for _ in range(n)]
    a[3] = a[40] = 0
    assert O.fr_to_ints(O.fr_batch_invert(fr_arr(a))) == [pow(x, -1, P.R) if x else 0 for x in a]
    num = [rnd.randrange(1, P.R) for _ in range(n)]
    den = [rnd.randrange(1, P.R) for _ in range(n)]
    exp = [1]
    for i in range(n - 1):
        exp.append(exp[-1] * num[i] * pow(den[i], -1, P.R) % P.R)
    assert O.fr_to_ints(O.fr_grand_product(fr_arr(num), fr_arr(den))) == exp
    for b in (rnd.randrange(P.R), 0, 1):
        q = O.fr_to_ints(O.fr_kate_division(fr_arr(a), fr_arr([b])[0]))
        back = [0] * n                      # (X - b) q(X) + a(b) == a(X)
        for i, c in enumerate(q):
            back[i + 1] = (back[i + 1] + c) % P.R
            back[i] = (back[i] - b * c) % P.R
        back[0] = (back[0] + P.eval_poly(a, b)) % P.R
        assert back == a


def test_oracle_srs_setup_is_lagrange_basis():
    """ParamsKZG::setup restatement: g[i] = s^i G and g_lagrange[i] = L_i(s) G (Lagrange polynomials of the
    2^k-th roots of unity), checked against the product formula."""
    k, n, s = 3, 8, 123456789123456789
    g, gl = O.srs_setup(k, fr_arr([s])[0])
    w = P.omega_for(k)
    for i in range(n):
        assert O.g1_affine_to_ints(g[i]) == P.g1_mul(P.G1_GEN, pow(s, i, P.R))
        num = den = 1
        for j in range(n):
            if j != i:
                num = num * (s - pow(w, j, P.R)) % P.R
                den = den * (pow(w, i, P.R) - pow(w, j, P.R)) % P.R
        assert O.g1_affine_to_ints(gl[i]) == P.g1_mul(P.G1_GEN, num * pow(den, -1, P.R) % P.R)
    # KZG consistency on the CPU path: commit(coeffs) == commit_lagrange(evals)
    d = O.EvaluationDomain(4, k)
    ev = O.fr_fill(n, 9)
    assert (O.best_multiexp_affine(ev, gl) == O.best_multiexp_affine(d.lagrange_to_coeff(ev), g)).all()
