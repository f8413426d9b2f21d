"""GPU parity: the CUDA path (through the C ABI of libh2v.so) against the CPU oracle, the committed
golden vectors, and size-independent properties.  Bit-exact everywhere (integer arithmetic)."""
import random

import numpy as np
import pytest

from common import fr_arr, g1_arr, golden, ipt, ival
from oracle import oracle as O
from oracle import pyref as P

pytestmark = pytest.mark.gpu

FQ_ONE = O.to_mont(O.ints_to_limbs([1]), O.FQ)[0]


def jac(aff):
    z = FQ_ONE if np.asarray(aff).any() else np.zeros(4, dtype=np.uint64)
    return np.concatenate([aff, z])


# ----------------------------------------------------------------------------- field / group layer
@pytest.mark.parametrize("field,nm", [(0, "fr"), (1, "fq")])
def test_device_field_ops(h2v, field, nm):
    mod = P.R if field == 0 else P.P
    rnd = random.Random(field)
    vals = [0, 1, mod - 1, mod - 2, 2, (1 << 253), mod >> 1] + [rnd.randrange(mod) for _ in range(2041)]
    a = O.to_mont(O.ints_to_limbs(vals), field)
    b = O.to_mont(O.ints_to_limbs(list(reversed(vals))), field)
    for op, opn in ((0, "mul"), (1, "add"), (2, "sub")):
        got = h2v.selftest_field(field, op, a, b)
        exp = np.stack([O.field_op(f"{nm}_{opn}", a[i], b[i]) for i in range(len(a))])
        assert (got == exp).all(), opn
    sq = np.stack([O.field_op(f"{nm}_mul", a[i], a[i]) for i in range(len(a))])
    for op in (5, 6):           # the dedicated Montgomery squaring on canonical inputs and on lazily reduced ones (x + m)
        assert (h2v.selftest_field(field, op, a) == sq).all(), op
    exp = np.stack([O.field_op(f"{nm}_inv", a[i]) for i in range(1, 65)])
    for op in (3, 4):           # Fermat and binary-Euclid inversions
        assert (h2v.selftest_field(field, op, a[1:65]) == exp).all(), op
    assert not h2v.selftest_field(field, 4, a[0:1]).any()      # 0 -> 0


def test_device_shoup_product(h2v):
    """The NTT's twiddle product (shoup.cuh): ANY 256-bit a (the butterflies hold lazily reduced values) times a twiddle,
    with the truncated high product and the top-limb correction; compared with big-integer arithmetic."""
    rnd = random.Random(7)
    edge = [0, 1, P.R - 1, P.R, 2 * P.R, 4 * P.R - 1, (1 << 256) - 1, (1 << 256) - 2, 1 << 255, (1 << 254) - 1, 1 << 254]
    a_int = edge + [rnd.randrange(1 << 256) for _ in range(4096 - len(edge))]
    w_int = [0, 1, P.R - 1, P.R - 2, 2, P.R >> 1] + [rnd.randrange(P.R) for _ in range(4090)]
    a = O.ints_to_limbs(a_int)                               # taken as they are: not reduced, not converted
    b = O.to_mont(O.ints_to_limbs(w_int), 0)                 # the twiddle arrives in Montgomery form
    got = h2v.selftest_field(0, 7, a, b)
    exp = O.ints_to_limbs([(x * w) % P.R for x, w in zip(a_int, w_int)])
    assert (got == exp).all()
    # rows that force the worst case of the quotient estimate: a = 2^256 - 1 against every twiddle
    a2 = O.ints_to_limbs([(1 << 256) - 1] * len(w_int))
    assert (h2v.selftest_field(0, 7, a2, b) == O.ints_to_limbs([(((1 << 256) - 1) * w) % P.R for w in w_int])).all()


def test_device_group_law(h2v):
    pts = O.gen_bases(64)
    p, q = pts[:32].copy(), pts[32:].copy()
    q[1] = p[1]                                            # P + P
    x, y = O.g1_affine_to_ints(p[2])
    q[2] = O.g1_affine_from_ints((x, (-y) % P.P))          # P + (-P)
    q[3] = 0                                               # P + 0
    p[4] = 0                                               # 0 + Q
    p[5] = 0
    q[5] = 0                                               # 0 + 0
    for mode in (0, 1):
        got = h2v.selftest_group(mode, p, q)
        for i in range(32):
            assert (got[i] == O.g1_to_affine(O.g1_add_mixed(jac(p[i]), q[i]))).all(), (mode, i)
    got = h2v.selftest_group(2, p, q)
    for i in range(32):
        assert (got[i] == O.g1_to_affine(O.g1_double(jac(p[i])))).all(), i


# ----------------------------------------------------------------------------- best_fft
def test_best_fft_golden(h2v):
    for v in golden()["fft"]:
        a = fr_arr([ival(x) for x in v["a"]])
        w = fr_arr([ival(v["omega"])])[0]
        got = O.fr_to_ints(h2v.best_fft(a, w, v["log_n"]))
        assert got == [ival(x) for x in v["out"]], v["log_n"]


@pytest.mark.parametrize("log_n", list(range(0, 15)) + [16, 17, 18, 19])
def test_best_fft_vs_oracle(h2v, log_n):
    a = O.fr_fill(1 << log_n, 1000 + log_n, mode=log_n % 2)
    w = fr_arr([P.omega_for(log_n)])[0]
    assert (h2v.best_fft(a, w, log_n) == O.best_fft(a, w, log_n)).all()


def test_best_fft_other_root_and_linearity(h2v):
    L = 12
    w = fr_arr([pow(P.omega_for(L), 5, P.R)])[0]           # another primitive 2^L-th root
    a, b = O.fr_fill(1 << L, 1), O.fr_fill(1 << L, 2)
    fa, fb = h2v.best_fft(a, w, L), h2v.best_fft(b, w, L)
    assert (fa == O.best_fft(a, w, L)).all()
    s = fr_arr([(x + y) % P.R for x, y in zip(O.fr_to_ints(a), O.fr_to_ints(b))])
    fs = O.fr_to_ints(h2v.best_fft(s, w, L))
    assert fs == [(x + y) % P.R for x, y in zip(O.fr_to_ints(fa), O.fr_to_ints(fb))]


@pytest.mark.parametrize("log_n", [20, 22, 24])
def test_best_fft_large_properties(h2v, log_n):
    """Full-size sweep sizes: Horner spot checks against the definition + inverse round trip."""
    n = 1 << log_n
    a = O.fr_fill(n, 31 + log_n)
    wi = P.omega_for(log_n)
    w = fr_arr([wi])[0]
    out = h2v.best_fft(a, w, log_n)
    rnd = random.Random(log_n)
    for j in [0, 1, n - 1] + rnd.sample(range(n), 3):
        x = fr_arr([pow(wi, j, P.R)])[0]
        assert (O.fr_eval_poly(a, x) == out[j]).all(), j
    back = h2v.best_fft(out, fr_arr([pow(wi, -1, P.R)])[0], log_n)
    ninv = fr_arr([pow(n, -1, P.R)])[0]
    idx = [0, 5, n // 2, n - 1]
    for j in idx:
        assert (O.field_op("fr_mul", back[j], ninv) == a[j]).all()


def test_best_fft_rejects_bad_length(h2v):
    with pytest.raises(ValueError):
        h2v.best_fft(O.fr_fill(12, 1), fr_arr([P.omega_for(4)])[0], 4)


# ----------------------------------------------------------------------------- EvaluationDomain
def test_domain_golden(h2v):
    for v in golden()["domain"]:
        d = h2v.EvaluationDomain(v["j"], v["k"])
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
        d.close()


@pytest.mark.parametrize("k", [3, 6, 9, 10, 13, 16])
def test_domain_vs_oracle(h2v, k):
    d, od = h2v.EvaluationDomain(4, k), O.EvaluationDomain(4, k)
    a = O.fr_fill(1 << k, 70 + k, mode=1)
    assert (d.lagrange_to_coeff(a) == od.lagrange_to_coeff(a)).all()
    assert (d.coeff_to_lagrange(a) == od.coeff_to_lagrange(a)).all()
    e, oe = d.coeff_to_extended(a), od.coeff_to_extended(a)
    assert (e == oe).all()
    h = O.fr_fill(1 << d.extended_k, 90 + k)
    assert (d.extended_to_coeff(h) == od.extended_to_coeff(h)).all()
    assert (d.divide_by_vanishing_poly(h) == od.divide_by_vanishing_poly(h)).all()
    fused = d.transform_batch(h2v.OP_DIVIDE_BY_VANISHING, [h, e])
    assert (fused[0] == od.extended_to_coeff(od.divide_by_vanishing_poly(h))).all()
    assert (fused[1] == od.extended_to_coeff(od.divide_by_vanishing_poly(oe))).all()
    d.close()


def test_domain_batch_matches_single(h2v):
    k = 11
    d, od = h2v.EvaluationDomain(4, k), O.EvaluationDomain(4, k)
    cols = [O.fr_fill(1 << k, 200 + i, mode=i % 2) for i in range(7)]
    for op, f in ((h2v.OP_LAGRANGE_TO_COEFF, od.lagrange_to_coeff), (h2v.OP_COEFF_TO_LAGRANGE, od.coeff_to_lagrange),
                  (h2v.OP_COEFF_TO_EXTENDED, od.coeff_to_extended)):
        outs = d.transform_batch(op, cols)
        for c, o in zip(cols, outs):
            assert (o == f(c)).all()
    assert d.transform_batch(h2v.OP_LAGRANGE_TO_COEFF, []) == []
    d.close()


def test_domain_roundtrip_k20(h2v):
    """config 4 size: coset round trip extended_to_coeff(coeff_to_extended(a)) = a || 0 and iNTT(NTT) = id."""
    d = h2v.EvaluationDomain(4, 20)
    a = O.fr_fill(d.n, 2020, mode=1)
    assert (d.coeff_to_lagrange(d.lagrange_to_coeff(a)) == a).all()
    ext = d.coeff_to_extended(a)
    back = d.extended_to_coeff(ext)
    assert (back[: d.n] == a).all() and not back[d.n:].any()
    # spot-check the coset evaluations against the definition: ext[i] = a(zeta * w_ext^i)
    zeta, wext = O.fr_to_ints(d.g_coset)[0], O.fr_to_ints(d.extended_omega)[0]
    for i in (0, 1, 12345, d.extended_n - 1):
        x = fr_arr([zeta * pow(wext, i, P.R) % P.R])[0]
        assert (O.fr_eval_poly(a, x) == ext[i]).all()
    d.close()


def test_domain_other_degrees_and_errors(h2v):
    for j, k in ((2, 5), (3, 5), (5, 4), (9, 3)):
        d, od = h2v.EvaluationDomain(j, k), O.EvaluationDomain(j, k)
        assert d.extended_k == od.extended_k
        a = O.fr_fill(1 << k, j * 100 + k)
        e = d.coeff_to_extended(a)
        assert (e == od.coeff_to_extended(a)).all()
        assert (d.extended_to_coeff(e) == od.extended_to_coeff(e)).all()
        d.close()
    with pytest.raises(ValueError):
        h2v.EvaluationDomain(4, 27)
    d = h2v.EvaluationDomain(4, 4)
    with pytest.raises(ValueError):
        d.lagrange_to_coeff(O.fr_fill(15, 1))
    d.close()


# ----------------------------------------------------------------------------- best_multiexp
def test_best_multiexp_golden(h2v):
    for v in golden()["msm"]:
        s = fr_arr([ival(x) for x in v["scalars"]])
        b = g1_arr([ipt(p) for p in v["bases"]])
        got = O.g1_affine_to_ints(O.g1_to_affine(h2v.best_multiexp(s, b)))
        assert got == ipt(v["result"]), v["dist"]


@pytest.mark.parametrize("n,mode", [(1, 0), (3, 0), (31, 1), (32, 0), (257, 1), (1000, 0), (4096, 1), (1 << 13, 0), (12345, 1)])
def test_best_multiexp_vs_oracle(h2v, n, mode):
    b = O.gen_bases(n)
    s = O.fr_fill(n, 300 + n, mode=mode)
    got = O.g1_to_affine(h2v.best_multiexp(s, b))
    assert (got == O.best_multiexp_affine(s, b)).all()
    assert (got == O.msm_closed_form(s)).all()


def test_best_multiexp_edge_cases(h2v):
    n = 512
    b = O.gen_bases(n)
    ident = np.zeros(8, dtype=np.uint64)
    # empty input -> identity
    e = h2v.best_multiexp(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 8), dtype=np.uint64))
    assert (O.g1_to_affine(e) == ident).all()
    # all-zero scalars -> identity
    z = np.zeros((n, 4), dtype=np.uint64)
    assert (O.g1_to_affine(h2v.best_multiexp(z, b)) == ident).all()
    # all scalars r-1 (every signed digit path), all ones, one hot
    for val in (P.R - 1, 1, (1 << 253) + 12345):
        s = fr_arr([val] * n)
        assert (O.g1_to_affine(h2v.best_multiexp(s, b)) == O.msm_closed_form(s)).all(), hex(val)
    # all bases equal (P + P inside every bucket), bases containing identity and opposite pairs
    g7 = O.g1_mul(O.g1_generator(), 7)
    same = np.tile(g7, (n, 1))
    s = O.fr_fill(n, 5, mode=1)
    tot = sum(O.fr_to_ints(s)) % P.R
    assert (O.g1_to_affine(h2v.best_multiexp(s, same)) == O.g1_mul(g7, tot)).all()
    mixed = b.copy()
    mixed[::3] = 0
    x, y = O.g1_affine_to_ints(b[1])
    mixed[4] = O.g1_affine_from_ints((x, (-y) % P.P))
    assert (O.g1_to_affine(h2v.best_multiexp(s, mixed)) == O.best_multiexp_affine(s, mixed)).all()
    with pytest.raises(ValueError):
        h2v.best_multiexp(s[:10], b[:11])


# ----------------------------------------------------------------------------- ParamsKZG commit path
@pytest.mark.parametrize("k,ncols,mode", [(3, 2, 0), (8, 5, 1), (12, 6, 0), (13, 12, 1), (16, 3, 0)])
def test_commit_vs_closed_form(h2v, k, ncols, mode):
    n = 1 << k
    b = O.gen_bases(n)
    gm = O.gen_bases(n, a=12345677, b=987654321)
    srs = h2v.ParamsKZG(k, gm, b)
    cols = [O.fr_fill(n, 4000 + 17 * i + k, mode=mode, lookup_bits=12) for i in range(ncols)]
    got = srs.commit_batch(cols)
    for i, c in enumerate(cols):
        assert (got[i] == O.msm_closed_form(c)).all(), i
    # single-column entry points, both bases, short polynomial, oracle best_multiexp on the same inputs
    assert (srs.commit_lagrange(cols[0]) == got[0]).all()
    assert (srs.commit(cols[0]) == O.msm_closed_form(cols[0], a=12345677, b=987654321)).all()
    ln = n // 2 + 1
    assert (srs.commit_lagrange(cols[1][:ln]) == O.best_multiexp_affine(cols[1][:ln], b[:ln])).all()
    assert (srs.commit_lagrange(np.zeros((n, 4), dtype=np.uint64)) == 0).all()
    with pytest.raises(ValueError):
        srs.commit_lagrange(np.zeros((n + 1, 4), dtype=np.uint64))
    srs.close()


def test_commit_linearity_and_checksum_k16(h2v):
    """kmeans config size (n = 2^16): commit(a) + commit(b) = commit(a + b), and the batch agrees with itself."""
    k, n = 16, 1 << 16
    b = O.gen_bases(n)
    srs = h2v.ParamsKZG(k, None, b)
    a1, a2 = O.fr_fill(n, 1, mode=0), O.fr_fill(n, 2, mode=1, lookup_bits=15)
    s = fr_arr([(x + y) % P.R for x, y in zip(O.fr_to_ints(a1), O.fr_to_ints(a2))])
    c = srs.commit_batch([a1, a2, s, a1])
    assert (c[0] == c[3]).all()
    sum_pt = O.g1_to_affine(O.g1_add_mixed(jac(c[0]), c[1]))
    assert (sum_pt == c[2]).all()
    assert (c[0] == O.msm_closed_form(a1)).all() and (c[1] == O.msm_closed_form(a2)).all()
    with pytest.raises(ValueError):
        srs.commit(a1)   # monomial basis was not loaded
    srs.close()


def test_commit_k20_closed_form(h2v):
    """config 4 size (n = 2^20), uniform and witness-like scalars."""
    k, n = 20, 1 << 20
    b = O.gen_bases(n)
    srs = h2v.ParamsKZG(k, None, b)
    cols = [O.fr_fill(n, 20, mode=0), O.fr_fill(n, 21, mode=1, lookup_bits=19)]
    got = srs.commit_batch(cols)
    for i, c in enumerate(cols):
        assert (got[i] == O.msm_closed_form(c)).all(), i
    srs.close()


@pytest.mark.parametrize("log_n,mode", [(22, 0), (24, 1)])
def test_best_multiexp_sweep_sizes_closed_form(h2v, log_n, mode):
    """config 5 sweep sizes: device-generated bases (a*i+b)*G, closed-form check on the host (O(n) field ops)."""
    n = 1 << log_n
    bases = h2v.synthetic_bases(n)
    chk = O.gen_bases(64)
    assert (bases[:64] == chk).all() and O.g1_is_on_curve(bases[n - 1])
    s = O.fr_fill(n, 77 + log_n, mode=mode, lookup_bits=20)
    got = O.g1_to_affine(h2v.best_multiexp(s, bases))
    assert (got == O.msm_closed_form(s)).all()


def test_synthetic_bases_match_oracle(h2v):
    for n in (1, 5, 300):
        assert (h2v.synthetic_bases(n) == O.gen_bases(n)).all()
    assert (h2v.synthetic_bases(17, 12345677, 987654321) == O.gen_bases(17, a=12345677, b=987654321)).all()


def test_witness_columns_with_giant_buckets(h2v):
    """A column that is 90% ones and small values puts most points into a handful of buckets: exercises the
    chunk-straddling merge and the warp-per-bucket path for long spans."""
    k, n = 14, 1 << 14
    b = O.gen_bases(n)
    srs = h2v.ParamsKZG(k, None, b)
    rnd = random.Random(5)
    vals = [1 if rnd.random() < 0.9 else rnd.choice([0, 2, P.R - 1, P.R - 5, rnd.randrange(P.R)]) for _ in range(n)]
    col = fr_arr(vals)
    ones = fr_arr([1] * n)
    neg = fr_arr([P.R - 3] * n)
    got = srs.commit_batch([col, ones, neg])
    for g, c in zip(got, (col, ones, neg)):
        assert (g == O.msm_closed_form(c)).all()
    srs.close()


def test_cpp_mirror(h2v, tmp_path):
    """include/h2v.hpp: compile, link against libh2v.so and run the round-trip check on the GPU."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cpp_mirror_check")
    libdir = os.path.join(root, "halo2_vectordb_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "tests", "cpp_mirror_check.cpp"), "-L" + libdir, "-lh2v", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "cpp mirror ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("table,chunk", [(1, 4), (0, 7), (1, 64), (0, 5), (-1, 1), (1, 1)])
def test_msm_tuning_invariance(h2v, table, chunk):
    """The window table (main / small window) and the chunk length change the schedule, never the result: force
    them on small inputs (uniform, witness-like with giant buckets, duplicate / opposite / identity bases)."""
    ba_rounds = table + 1
    try:
        h2v.set_tuning(chunk, table)
        k, n = 11, 1 << 11
        b = O.gen_bases(n)
        srs = h2v.ParamsKZG(k, None, b)
        rnd = random.Random(ba_rounds * 100 + chunk)
        cols = [O.fr_fill(n, 1, mode=0), O.fr_fill(n, 2, mode=1), fr_arr([1] * n),
                fr_arr([rnd.choice([0, 1, 2, P.R - 1]) for _ in range(n)])]
        got = srs.commit_batch(cols)
        for g, c in zip(got, cols):
            assert (g == O.msm_closed_form(c)).all()
        srs.close()
        # raw path with pathological bases
        m = 700
        g7 = O.g1_mul(O.g1_generator(), 7)
        x, y = O.g1_affine_to_ints(g7)
        neg7 = O.g1_affine_from_ints((x, (-y) % P.P))
        bases = np.stack([rnd.choice([g7, g7, neg7, np.zeros(8, dtype=np.uint64), b[3]]) for _ in range(m)])
        s = O.fr_fill(m, 9, mode=1)
        assert (O.g1_to_affine(h2v.best_multiexp(s, bases)) == O.best_multiexp_affine(s, bases)).all()
        for v in golden()["msm"]:
            sc = fr_arr([ival(x) for x in v["scalars"]])
            bs = g1_arr([ipt(p) for p in v["bases"]])
            assert O.g1_affine_to_ints(O.g1_to_affine(h2v.best_multiexp(sc, bs))) == ipt(v["result"])
    finally:
        h2v.set_tuning(-1, -1)


def test_tiny_srs_and_pinned_buffers(h2v):
    """k = 0, 1, 2 (degenerate SRS sizes) and page-locked caller buffers."""
    for k in (0, 1, 2):
        n = 1 << k
        b = O.gen_bases(n)
        srs = h2v.ParamsKZG(k, b, b)
        s = O.fr_fill(n, 40 + k)
        assert (srs.commit(s) == O.best_multiexp_affine(s, b)).all()
        assert (srs.commit_lagrange(s) == O.msm_closed_form(s)).all()
        srs.close()
    k, n = 12, 1 << 12
    d, od = h2v.EvaluationDomain(4, k), O.EvaluationDomain(4, k)
    a = O.fr_fill(n, 3)
    h2v.host_register(a)
    try:
        assert (d.lagrange_to_coeff(a) == od.lagrange_to_coeff(a)).all()
        srs = h2v.ParamsKZG(k, None, O.gen_bases(n))
        assert (srs.commit_batch([a] * 30)[29] == O.msm_closed_form(a)).all()
        srs.close()
    finally:
        h2v.host_unregister(a)
    d.close()


# ----------------------------------------------------------------------------- "next" row 2 primitives
@pytest.mark.parametrize("n,n_polys,n_points", [(1, 1, 1), (7, 2, 3), (256, 3, 2), (1000, 2, 2), (1 << 13, 5, 3), (1 << 16, 2, 4)])
def test_eval_polynomial_vs_oracle(h2v, n, n_polys, n_points):
    polys = [O.fr_fill(n, 60 + i + n, mode=i % 2) for i in range(n_polys)]
    pts = O.fr_fill(n_points, 5 + n)
    pts[0] = 0 if n_points > 1 else pts[0]          # x = 0 -> constant term
    got = h2v.eval_polynomial_batch(polys, pts)
    for i, p in enumerate(polys):
        for j in range(n_points):
            assert (got[i, j] == O.fr_eval_poly(p, pts[j])).all(), (i, j)
    assert (h2v.eval_polynomial(polys[0], pts[-1]) == O.fr_eval_poly(polys[0], pts[-1])).all()


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 300, 4097, 1 << 16])
def test_batch_invert_and_grand_product_vs_oracle(h2v, n):
    a = O.fr_fill(n, 800 + n, mode=1)              # witness-like: contains zeros and ones
    assert (h2v.batch_invert(a) == O.fr_batch_invert(a)).all()
    num, den = O.fr_fill(n, 900 + n), O.fr_fill(n, 901 + n)
    assert (h2v.grand_product(num, den) == O.fr_grand_product(num, den)).all()


@pytest.mark.parametrize("n", [2, 3, 9, 2048, 2049, 5000, 1 << 16])
def test_kate_division_vs_oracle(h2v, n):
    a = O.fr_fill(n, 700 + n)
    for b in (O.fr_fill(1, n)[0], np.zeros(4, dtype=np.uint64), O.fr_from_ints([1])[0]):
        assert (h2v.kate_division(a, b) == O.fr_kate_division(a, b)).all()


def test_row2_properties_k20(h2v):
    """config-4 size: z[n-1] * r[n-1] telescopes, a(x) = (x - b) q(x) + a(b) at a random x."""
    n = 1 << 20
    a = O.fr_fill(n, 2024)
    b, x = O.fr_fill(2, 99)
    q = h2v.kate_division(a, b)
    ax, ab, qx = (h2v.eval_polynomial(a, x), h2v.eval_polynomial(a, b), h2v.eval_polynomial(q, x))
    lhs = O.field_op("fr_add", O.field_op("fr_mul", O.field_op("fr_sub", x, b), qx), ab)
    assert (lhs == ax).all() and (ax == O.fr_eval_poly(a, x)).all()
    num = O.fr_fill(n, 1)
    z = h2v.grand_product(num, num)                 # num / num = 1 everywhere
    one = O.fr_from_ints([1])[0]
    assert (z == one).all()
    inv = h2v.batch_invert(a)
    for i in (0, 12345, n - 1):
        assert (O.field_op("fr_mul", inv[i], a[i]) == one).all()


def test_concurrent_callers(h2v):
    """Entry points are thread-safe: worker threads commit / transform column by column at the same time
    (what stock create_proof does from its rayon pool); every result must still be exact."""
    import threading

    k, n = 12, 1 << 12
    b = O.gen_bases(n)
    srs = h2v.ParamsKZG(k, None, b)
    dom, od = h2v.EvaluationDomain(4, k), O.EvaluationDomain(4, k)
    cols = [O.fr_fill(n, 3000 + i, mode=i % 2) for i in range(24)]
    exp_c = [O.msm_closed_form(c) for c in cols]
    exp_t = [od.lagrange_to_coeff(c) for c in cols]
    errors = []

    def worker(tid):
        try:
            for i in range(tid, len(cols), 8):
                if not (srs.commit_lagrange(cols[i]) == exp_c[i]).all():
                    errors.append(("commit", i))
                if not (dom.lagrange_to_coeff(cols[i]) == exp_t[i]).all():
                    errors.append(("l2c", i))
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    srs.close()
    dom.close()


# ----------------------------------------------------------------------------- ParamsKZG::setup + KZG consistency
def test_srs_setup_vs_oracle(h2v):
    s = O.fr_from_ints([0x1234567890ABCDEF1234567890ABCDEF % P.R])[0]
    for k in (0, 1, 4, 7):
        g, gl = h2v.srs_setup(k, s)
        og, ogl = O.srs_setup(k, s)
        assert (g == og).all() and (gl == ogl).all(), k
    with pytest.raises(ValueError):
        h2v.srs_setup(3, O.fr_from_ints([P.omega_for(3)])[0])       # s must not be a root of unity


@pytest.mark.parametrize("k", [6, 12, 16])
def test_kzg_consistency_monomial_vs_lagrange(h2v, k):
    """With a real SRS the two bases describe the same commitment scheme: for any polynomial,
    commit(coefficients) == commit_lagrange(evaluations).  This ties MSM (both bases, window tables),
    the NTT (lagrange_to_coeff / coeff_to_lagrange) and the setup together at the kmeans size."""
    n = 1 << k
    s = O.fr_fill(1, 4242 + k)[0]
    srs = h2v.ParamsKZG.setup(k, s)
    dom = h2v.EvaluationDomain(4, k)
    one = O.fr_from_ints([1])[0]
    G = O.g1_generator()
    # sum_i L_i(s) = 1  =>  commit_lagrange(1, 1, ..., 1) = G ;  commit(X) = s G = g[1]
    assert (srs.commit_lagrange(np.tile(one, (n, 1))) == G).all()
    xpoly = np.zeros((n, 4), dtype=np.uint64)
    xpoly[1] = one
    assert (srs.commit(xpoly) == srs.g[1]).all()
    assert (srs.g[0] == G).all()
    for mode in (0, 1):
        evals = O.fr_fill(n, 77 + mode, mode=mode, lookup_bits=max(k - 1, 2))
        coeffs = dom.lagrange_to_coeff(evals)
        c_l = srs.commit_lagrange(evals)
        c_m = srs.commit(coeffs)
        assert (c_l == c_m).all(), mode
        assert O.g1_is_on_curve(c_l)
        # and the commitment is p(s) G
        ps = O.fr_to_ints(O.fr_eval_poly(coeffs, s))[0]
        assert (c_m == O.g1_mul(G, ps)).all()
    srs.close()
    dom.close()


def test_kzg_opening_identity(h2v):
    """A KZG opening assembled from the library's own pieces: q = kate_division(p, z), then
    commit(q) * (s - z) == commit(p) - p(z) G, checked through the known secret."""
    k, n = 12, 1 << 12
    s = O.fr_fill(1, 31337)[0]
    srs = h2v.ParamsKZG.setup(k, s)
    p = O.fr_fill(n, 5)
    z = O.fr_fill(1, 6)[0]
    q = h2v.kate_division(p, z)
    qpad = np.zeros((n, 4), dtype=np.uint64)
    qpad[: n - 1] = q
    pz = h2v.eval_polynomial(p, z)
    cq = srs.commit(qpad)
    si, zi = O.fr_to_ints(s)[0], O.fr_to_ints(z)[0]
    ps, pzi = O.fr_to_ints(O.fr_eval_poly(p, s))[0], O.fr_to_ints(pz)[0]
    G = O.g1_generator()
    assert (cq == O.g1_mul(G, (ps - pzi) * pow(si - zi, -1, P.R) % P.R)).all()
    assert (srs.commit(p) == O.g1_mul(G, ps)).all()
    srs.close()


def test_device_resident_phase_through_the_abi(h2v):
    """One prover phase with the columns uploaded once and kept in HBM: commit_lagrange, lagrange_to_coeff,
    coeff_to_extended and evaluations, all through `_dev` entry points and h2v_dev_* buffers (no torch)."""
    k, n, cols = 10, 1 << 10, 5
    srs = h2v.ParamsKZG(k, None, O.gen_bases(n))
    dom, od = h2v.EvaluationDomain(4, k), O.EvaluationDomain(4, k)
    data = np.stack([O.fr_fill(n, 50 + i, mode=i % 2) for i in range(cols)])
    d_cols = h2v.DeviceBuffer(cols * n * 32)
    d_coef = h2v.DeviceBuffer(cols * n * 32)
    d_ext = h2v.DeviceBuffer(cols * 4 * n * 32)
    d_commit = h2v.DeviceBuffer(cols * 64)
    pts = O.fr_fill(2, 8)
    d_pts = h2v.DeviceBuffer(2 * 32)
    d_ev = h2v.DeviceBuffer(cols * 2 * 32)
    d_cols.upload(data)
    d_pts.upload(pts)
    srs.commit_batch_dev(d_cols.ptr, n, cols, n, d_commit.ptr)
    dom.transform_dev(h2v.OP_LAGRANGE_TO_COEFF, d_cols.ptr, n, d_coef.ptr, n, cols)
    dom.transform_dev(h2v.OP_COEFF_TO_EXTENDED, d_coef.ptr, n, d_ext.ptr, 4 * n, cols)
    h2v._check(h2v.lib().h2v_eval_polynomial_dev(d_coef.ptr, n, cols, n, d_pts.ptr, 2, d_ev.ptr))
    commits = d_commit.download((cols, 8))
    coefs = d_coef.download((cols, n, 4))
    exts = d_ext.download((cols, 4 * n, 4))
    evs = d_ev.download((cols, 2, 4))
    for i in range(cols):
        assert (commits[i] == O.msm_closed_form(data[i])).all()
        oc = od.lagrange_to_coeff(data[i])
        assert (coefs[i] == oc).all()
        assert (exts[i] == od.coeff_to_extended(oc)).all()
        for j in range(2):
            assert (evs[i, j] == O.fr_eval_poly(oc, pts[j])).all()
    with pytest.raises(ValueError):
        dom.transform_dev(h2v.OP_COEFF_TO_EXTENDED, d_coef.ptr, n, d_ext.ptr, n, cols)   # out stride too short
    for b in (d_cols, d_coef, d_ext, d_commit, d_pts, d_ev):
        b.free()
    srs.close()
    dom.close()


def test_large_batches_take_the_pipelined_paths(h2v):
    """Batches larger than one staging buffer: double-buffered commit (two streams), pipelined transforms, and
    more columns than one MSM launch holds (workspace-sized sub-batches)."""
    k, n = 16, 1 << 16
    b = h2v.synthetic_bases(n)
    srs = h2v.ParamsKZG(k, None, b)
    base_cols = [O.fr_fill(n, 600 + i, mode=i % 2, lookup_bits=15) for i in range(5)]
    exp = [O.msm_closed_form(c) for c in base_cols]
    cols = [base_cols[i % 5] for i in range(40)]                 # 80 MiB of scalars -> 2 sub-batches
    got = srs.commit_batch(cols)
    for i in range(40):
        assert (got[i] == exp[i % 5]).all(), i
    # 700 identical device-resident columns (stride 0): exceeds the columns one launch holds
    d_col = h2v.DeviceBuffer(n * 32)
    d_col.upload(base_cols[0])
    d_out = h2v.DeviceBuffer(700 * 64)
    srs.commit_batch_dev(d_col.ptr, 0, 700, n, d_out.ptr)
    outs = d_out.download((700, 8))
    assert (outs == exp[0]).all()
    d_col.free()
    d_out.free()
    srs.close()
    dom, od = h2v.EvaluationDomain(4, k), O.EvaluationDomain(4, k)
    tcols = [base_cols[i % 5] for i in range(12)]                # 96 MiB of extended output -> pipelined path
    ext = dom.transform_batch(h2v.OP_COEFF_TO_EXTENDED, tcols)
    oext = [od.coeff_to_extended(c) for c in base_cols]
    for i in range(12):
        assert (ext[i] == oext[i % 5]).all(), i
    dom.close()


def test_sparse_input_short_columns(h2v):
    """6 columns of 14 scalars against a 2^10 SRS -- far fewer entries than buckets -- with one entry per accumulate
    thread, on both window tables."""
    for table in (0, 1):
        _sparse_short_columns(h2v, table)


def _sparse_short_columns(h2v, table):
    try:
        h2v.set_tuning(1, table)
        k, n, ln = 10, 1 << 10, 14
        b = O.gen_bases(n)
        srs = h2v.ParamsKZG(k, None, b)
        cols = [O.fr_fill(ln, 70 + i, mode=i % 2) for i in range(6)]
        got = srs.commit_batch(cols)
        for g, c in zip(got, cols):
            assert (g == O.best_multiexp_affine(c, b[:ln])).all()
        srs.close()
    finally:
        h2v.set_tuning(-1, -1)


def test_plain_c_example(h2v, tmp_path):
    """examples/commit_example.c: the ABI from plain C99 -- setup, both commits agree, wire format."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "commit_example")
    libdir = os.path.join(root, "halo2_vectordb_b200")
    subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "commit_example.c"),
                           "-L" + libdir, "-lh2v", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "== commit(coeffs)" in out.stdout, out.stdout + out.stderr


# ---- one multiexp split over GPUs by index range: slices + g1_sum --------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 2, 3, 8, 100, 1000])
def test_g1_sum_vs_oracle(h2v, n):
    rng = random.Random(n)
    pts = O.gen_bases(max(n, 1), a=rng.randrange(1, 1 << 30), b=rng.randrange(1, 1 << 30))[:n]
    if n >= 3:      # identity, a duplicate and an opposite pair
        pts[0] = 0
        pts[1] = pts[2]
        xy = O.g1_affine_to_ints(pts[n - 1])
        pts[n - 2] = O.g1_affine_from_ints((xy[0], (-xy[1]) % P.P))
    acc = np.zeros(12, dtype=np.uint64)
    for p in pts:
        acc = O.g1_add_mixed(acc, p)
    assert np.array_equal(h2v.g1_sum(pts), O.g1_to_affine(acc))


@pytest.mark.gpu
@pytest.mark.parametrize("k,world", [(4, 2), (10, 2), (12, 4), (14, 8), (12, 3)])
def test_sliced_multiexp_equals_whole(h2v, k, world):
    """every 'rank' holds only its slice of the bases; the partial sums add up to the unsliced commitment"""
    from halo2_vectordb_b200 import sharding
    n = 1 << k
    bases = h2v.synthetic_bases(n)
    s = O.fr_fill(n, 900 + k, mode=k & 1)
    parts = []
    for r in range(world):
        lo, hi = sharding.index_slice(n, r, world)
        if world & (world - 1):         # uneven slices: the handle-free entry point
            parts.append(O.g1_to_affine(h2v.best_multiexp(s[lo:hi], bases[lo:hi])))
        else:                           # power-of-two slices: a handle over the slice's bases
            srs = h2v.ParamsKZG(k - (world.bit_length() - 1), None, bases[lo:hi])
            parts.append(srs.commit_lagrange(s[lo:hi]))
            srs.close()
    got = h2v.g1_sum(np.stack(parts))
    whole = h2v.ParamsKZG(k, None, bases)
    assert np.array_equal(got, whole.commit_lagrange(s))
    assert np.array_equal(got, O.msm_closed_form(s))
    whole.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,cols", [(1, 1), (5, 3), (2048, 2), (2049, 5), (1 << 13, 7), (1 << 16, 4)])
def test_grand_product_dev_batch_vs_oracle(h2v, n, cols):
    num = O.fr_fill(n * cols, 40 + cols).reshape(cols, n, 4)
    den = O.fr_fill(n * cols, 50 + cols).reshape(cols, n, 4)
    d_num, d_den, d_out = (h2v.DeviceBuffer(n * cols * 32) for _ in range(3))
    d_num.upload(num); d_den.upload(den)
    h2v.grand_product_dev(d_num.ptr, d_den.ptr, n, cols, d_out.ptr)
    got = d_out.download((cols, n, 4))
    for c in range(cols):
        assert np.array_equal(got[c], O.fr_grand_product(num[c], den[c])), (n, c)
        assert np.array_equal(got[c], h2v.grand_product(num[c], den[c]))


@pytest.mark.gpu
def test_grand_product_zero_denominator_is_skipped_like_batch_invert(h2v):
    """ff's BatchInvert (which the permutation / lookup provers call) leaves a zero element as zero instead of poisoning
    the whole batch: one zero denominator in ONE of several columns only zeroes that column's running product from
    there on; the other columns are untouched"""
    n, cols = 300, 3
    num = O.fr_fill(n * cols, 61).reshape(cols, n, 4)
    den = O.fr_fill(n * cols, 62).reshape(cols, n, 4)
    den[1, 7] = 0
    d_num, d_den, d_out = (h2v.DeviceBuffer(n * cols * 32) for _ in range(3))
    d_num.upload(num); d_den.upload(den)
    h2v.grand_product_dev(d_num.ptr, d_den.ptr, n, cols, d_out.ptr)
    got = d_out.download((cols, n, 4))
    for c in range(cols):
        ni, di = O.fr_to_ints(num[c]), O.fr_to_ints(den[c])
        z, want = 1, []
        for i in range(n):
            want.append(z)
            z = z * ni[i] % P.R * (pow(di[i], P.R - 2, P.R) if di[i] else 0) % P.R
        assert O.fr_to_ints(got[c]) == want, c
        assert np.array_equal(got[c], h2v.grand_product(num[c], den[c]))
    assert any(got[1, 8]) is False or not got[1, 8].any()
    assert got[0, 8].any() and got[2, 8].any()
