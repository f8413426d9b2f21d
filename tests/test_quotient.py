"""Quotient evaluation row loops: evaluate_h's gates / permutation / lookup terms.

CPU: the oracle restatement against the defining property on a satisfied toy circuit (plonk_model.py).
GPU: the CUDA kernels through the C ABI -- bit-exact against the oracle on random columns, and the same
property test with every column device-resident from coeff_to_extended to the divided quotient."""
import random

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as P
from common import fr_arr
from plonk_model import ToyCircuit

R = P.R


def _names(circ):
    return list(circ.lagrange_columns().keys())


def _oracle_quotient(circ):
    """columns -> coefficients -> extended coset -> folded h -> h / (X^n - 1) in coefficient form, all through the oracle."""
    dom = O.EvaluationDomain(circ.degree, circ.k)
    lag = circ.lagrange_columns()
    coeff = {k: dom.lagrange_to_coeff(fr_arr(v)) for k, v in lag.items()}
    ext = {k: dom.coeff_to_extended(v) for k, v in coeff.items()}
    ne = 1 << dom.extended_k
    nc = len(circ.cols)
    y, beta, gamma = (fr_arr([v])[0] for v in (circ.y, circ.beta, circ.gamma))
    h = np.zeros((ne, 4), dtype=np.uint64)
    h = dom.quotient_gates(h, y, np.stack([ext[f"q{j}"] for j in range(circ.n_gates)]),
                           np.stack([ext[f"col{j}"] for j in range(circ.n_gates)]))
    h = dom.quotient_permutation(h, y, beta, gamma, circ.chunk_len, np.stack([ext[f"col{c}"] for c in range(nc)]),
                                 np.stack([ext[f"sigma{c}"] for c in range(nc)]),
                                 np.stack([ext[f"z{s}"] for s in range(circ.n_sets)]), ext["l0"], ext["l_last"], ext["l_active"], circ.bf)
    h = dom.quotient_lookup(h, y, beta, gamma, ext["l_input"], ext["l_table"], ext["perm_input"], ext["perm_table"],
                            ext["z_lookup"], ext["l0"], ext["l_last"], ext["l_active"])
    hq = dom.extended_to_coeff(dom.divide_by_vanishing_poly(h))
    return {k: O.fr_to_ints(v) for k, v in coeff.items()}, O.fr_to_ints(hq), h


def _check_identity(circ, coeff_ints, h_ints, seed, expect_ok=True):
    rng = random.Random(seed)
    ok = True
    for _ in range(2):
        x = rng.randrange(2, R)
        lhs = P.eval_poly(h_ints, x) * (pow(x, circ.n, R) - 1) % R
        ok = ok and lhs == circ.expected_at(coeff_ints, x)
    assert ok == expect_ok


@pytest.mark.parametrize("k,gate_cols,degree", [(5, 2, 4), (6, 2, 4), (6, 4, 4), (5, 1, 5), (6, 3, 6)])
def test_oracle_quotient_identity(k, gate_cols, degree):
    circ = ToyCircuit(k, seed=100 + k + gate_cols, n_gate_cols=gate_cols, degree=degree)
    coeff, hq, _ = _oracle_quotient(circ)
    assert len(hq) == circ.n * (degree - 1)
    _check_identity(circ, coeff, hq, seed=k)


def test_oracle_quotient_detects_unsatisfied_gate():
    circ = ToyCircuit(5, seed=7)
    circ.break_gate()
    coeff, hq, _ = _oracle_quotient(circ)
    _check_identity(circ, coeff, hq, seed=1, expect_ok=False)


def test_toy_circuit_vanishes_on_domain():
    """the model itself: E(omega^i) = 0 for every row (checked with big integers only)"""
    circ = ToyCircuit(5, seed=3)
    dom = circ.dom
    coeff = {k: dom.lagrange_to_coeff(v) for k, v in circ.lagrange_columns().items()}
    for i in (0, 1, 7, circ.u - 1, circ.u, circ.n - 1):
        assert circ.expected_at(coeff, pow(dom.omega, i, R)) == 0


# ----------------------------------------------------------------------------------------------- GPU
def _rand_ext(rng, n):
    return fr_arr([rng.randrange(R) for _ in range(n)])


class _Dev:
    """columns stacked in one device allocation, `stride` elements apart"""

    def __init__(self, h2v, arrs, stride):
        self.buf = h2v.DeviceBuffer(max(1, len(arrs)) * stride * 32)
        for i, a in enumerate(arrs):
            self.buf.upload(a, offset=i * stride * 32)
        self.ptr = self.buf.ptr


@pytest.mark.gpu
@pytest.mark.parametrize("k,degree,n_gates", [(4, 4, 1), (6, 4, 3), (8, 4, 5), (7, 3, 2), (6, 6, 2), (10, 4, 7)])
def test_gpu_quotient_gates_match_oracle(h2v, k, degree, n_gates):
    rng = random.Random(k * 31 + n_gates)
    gd, od = h2v.EvaluationDomain(degree, k), O.EvaluationDomain(degree, k)
    ne = 1 << od.extended_k
    stride = ne + 8      # padded stride: columns need not be contiguous
    q = [_rand_ext(rng, ne) for _ in range(n_gates)]
    a = [_rand_ext(rng, ne) for _ in range(n_gates)]
    h0, y = _rand_ext(rng, ne), _rand_ext(rng, 1)[0]
    dq, da, dh = _Dev(h2v, q, stride), _Dev(h2v, a, stride), _Dev(h2v, [h0], ne)
    gd.quotient_gates(dh.ptr, y, n_gates, dq.ptr, stride, da.ptr, stride)
    want = od.quotient_gates(h0, y, np.stack(q), np.stack(a))
    assert np.array_equal(dh.buf.download((ne, 4)), want)


@pytest.mark.gpu
@pytest.mark.parametrize("k,degree,n_cols,bf", [(4, 4, 1, 3), (6, 4, 4, 5), (6, 4, 5, 5), (8, 5, 7, 6), (7, 3, 3, 5), (10, 4, 9, 5)])
def test_gpu_quotient_permutation_matches_oracle(h2v, k, degree, n_cols, bf):
    rng = random.Random(k * 131 + n_cols)
    gd, od = h2v.EvaluationDomain(degree, k), O.EvaluationDomain(degree, k)
    ne = 1 << od.extended_k
    chunk = degree - 2
    n_sets = (n_cols + chunk - 1) // chunk
    cols = [_rand_ext(rng, ne) for _ in range(n_cols)]
    sig = [_rand_ext(rng, ne) for _ in range(n_cols)]
    z = [_rand_ext(rng, ne) for _ in range(n_sets)]
    l0, ll, la, h0 = (_rand_ext(rng, ne) for _ in range(4))
    y, beta, gamma = (_rand_ext(rng, 1)[0] for _ in range(3))
    dc, ds, dz = _Dev(h2v, cols, ne), _Dev(h2v, sig, ne + 4), _Dev(h2v, z, ne)
    dl, dh = _Dev(h2v, [l0, ll, la], ne), _Dev(h2v, [h0], ne)
    gd.quotient_permutation(dh.ptr, y, beta, gamma, n_cols, chunk, dc.ptr, ne, ds.ptr, ne + 4, dz.ptr, ne,
                            dl.ptr, dl.ptr + ne * 32, dl.ptr + 2 * ne * 32, bf)
    want = od.quotient_permutation(h0, y, beta, gamma, chunk, np.stack(cols), np.stack(sig), np.stack(z), l0, ll, la, bf)
    assert np.array_equal(dh.buf.download((ne, 4)), want)


@pytest.mark.gpu
@pytest.mark.parametrize("k,degree", [(4, 4), (6, 4), (9, 4), (7, 3), (6, 6)])
def test_gpu_quotient_lookup_matches_oracle(h2v, k, degree):
    rng = random.Random(k * 17 + degree)
    gd, od = h2v.EvaluationDomain(degree, k), O.EvaluationDomain(degree, k)
    ne = 1 << od.extended_k
    arrs = [_rand_ext(rng, ne) for _ in range(8)]     # input, table, A', S', z, l0, l_last, l_active
    h0 = _rand_ext(rng, ne)
    y, beta, gamma = (_rand_ext(rng, 1)[0] for _ in range(3))
    d, dh = _Dev(h2v, arrs, ne), _Dev(h2v, [h0], ne)
    gd.quotient_lookup(dh.ptr, y, beta, gamma, *[d.ptr + i * ne * 32 for i in range(8)])
    want = od.quotient_lookup(h0, y, beta, gamma, *arrs)
    assert np.array_equal(dh.buf.download((ne, 4)), want)


def _gpu_quotient(h2v, circ):
    """the device-resident flow a prover would run: only n-sized columns go up, only the (j-1)n quotient comes back"""
    dom = h2v.EvaluationDomain(circ.degree, circ.k)
    n, ne = circ.n, 1 << dom.extended_k
    lag = circ.lagrange_columns()
    names = list(lag.keys())
    ix = {nm: i for i, nm in enumerate(names)}
    d_lag = _Dev(h2v, [fr_arr(lag[nm]) for nm in names], n)
    d_coeff = h2v.DeviceBuffer(len(names) * n * 32)
    d_ext = h2v.DeviceBuffer(len(names) * ne * 32)
    dom.transform_dev(h2v.OP_LAGRANGE_TO_COEFF, d_lag.ptr, n, d_coeff.ptr, n, len(names))
    dom.transform_dev(h2v.OP_COEFF_TO_EXTENDED, d_coeff.ptr, n, d_ext.ptr, ne, len(names))
    e = lambda nm: d_ext.ptr + ix[nm] * ne * 32
    # the column order of lagrange_columns() keeps q_j apart and col_c / sigma_c interleaved: strides express that
    nc = len(circ.cols)
    y, beta, gamma = (fr_arr([v])[0] for v in (circ.y, circ.beta, circ.gamma))
    dh = _Dev(h2v, [np.zeros((ne, 4), dtype=np.uint64)], ne)
    assert ix["q1"] == ix["q0"] + 1 if circ.n_gates > 1 else True
    assert ix["col1"] == ix["col0"] + 2 and ix["sigma1"] == ix["sigma0"] + 2
    dom.quotient_gates(dh.ptr, y, circ.n_gates, e("q0"), ne, e("col0"), 2 * ne)
    dom.quotient_permutation(dh.ptr, y, beta, gamma, nc, circ.chunk_len, e("col0"), 2 * ne, e("sigma0"), 2 * ne, e("z0"), ne,
                             e("l0"), e("l_last"), e("l_active"), circ.bf)
    dom.quotient_lookup(dh.ptr, y, beta, gamma, e("l_input"), e("l_table"), e("perm_input"), e("perm_table"), e("z_lookup"),
                        e("l0"), e("l_last"), e("l_active"))
    h_ext = dh.buf.download((ne, 4))
    d_out = h2v.DeviceBuffer(ne * 32)
    dom.transform_dev(h2v.OP_DIVIDE_BY_VANISHING, dh.ptr, ne, d_out.ptr, ne, 1)
    hq = d_out.download((n * (circ.degree - 1), 4))
    coeff = d_coeff.download((len(names), n, 4))
    return {nm: O.fr_to_ints(coeff[ix[nm]]) for nm in names}, O.fr_to_ints(hq), h_ext


@pytest.mark.gpu
@pytest.mark.parametrize("k,gate_cols,degree", [(5, 2, 4), (6, 4, 4), (7, 3, 4), (5, 1, 5), (6, 3, 6)])
def test_gpu_quotient_identity_and_oracle(h2v, k, gate_cols, degree):
    circ = ToyCircuit(k, seed=200 + k + gate_cols, n_gate_cols=gate_cols, degree=degree)
    coeff, hq, h_ext = _gpu_quotient(h2v, circ)
    _check_identity(circ, coeff, hq, seed=k)                    # the verifier's equation, big integers
    o_coeff, o_hq, o_h_ext = _oracle_quotient(circ)             # and bit-exact against the oracle end to end
    assert np.array_equal(h_ext, o_h_ext) and hq == o_hq and coeff == o_coeff


@pytest.mark.gpu
def test_gpu_quotient_rejects_bad_arguments(h2v):
    dom = h2v.EvaluationDomain(4, 5)
    ne = 1 << dom.extended_k
    buf = h2v.DeviceBuffer(4 * ne * 32)
    y = fr_arr([3])[0]
    with pytest.raises(ValueError):
        dom.quotient_gates(buf.ptr, y, 2, buf.ptr, ne - 1, buf.ptr, ne)          # stride shorter than the column
    with pytest.raises(ValueError):
        dom.quotient_gates(None, y, 1, buf.ptr, ne, buf.ptr, ne)
    with pytest.raises(ValueError):
        dom.quotient_permutation(buf.ptr, y, y, y, 2, 0, buf.ptr, ne, buf.ptr, ne, buf.ptr, ne, buf.ptr, buf.ptr, buf.ptr, 5)
    with pytest.raises(ValueError):
        dom.quotient_lookup(buf.ptr, y, y, y, buf.ptr, buf.ptr, None, buf.ptr, buf.ptr, buf.ptr, buf.ptr, buf.ptr)
    dom.quotient_gates(buf.ptr, y, 0, None, 0, None, 0)                          # no gates: no-op, like upstream's empty loop


@pytest.mark.gpu
@pytest.mark.parametrize("k,gate_cols", [(5, 2), (7, 3)])
def test_gpu_lookup_and_products_feed_the_quotient(h2v, k, gate_cols):
    """A', S' from h2v.permute_expression_pair and every grand product from h2v.grand_product replace the model's own
    (the arrangement of S' differs from the model's, both are valid): the quotient built from them by the device
    kernels must still satisfy the verifier's identity."""
    from plonk_model import DELTA
    circ = ToyCircuit(k, seed=300 + k, n_gate_cols=gate_cols)
    u, n, w = circ.u, circ.n, circ.dom.omega
    beta, gamma = circ.beta, circ.gamma
    ga, gs = h2v.permute_expression_pair(fr_arr(circ.l_input[:u]), fr_arr(circ.l_table[:u]))
    oa, os_ = O.permute_expression_pair(fr_arr(circ.l_input[:u]), fr_arr(circ.l_table[:u]))
    assert np.array_equal(ga, oa) and np.array_equal(gs, os_)
    circ.perm_input[:u], circ.perm_table[:u] = O.fr_to_ints(ga), O.fr_to_ints(gs)
    num = [(circ.l_input[r] + beta) * (circ.l_table[r] + gamma) % R for r in range(u)] + [1]
    den = [(circ.perm_input[r] + beta) * (circ.perm_table[r] + gamma) % R for r in range(u)] + [1]
    zl = O.fr_to_ints(h2v.grand_product(fr_arr(num), fr_arr(den)))
    assert zl[0] == 1 and zl[u] == 1
    circ.z_lookup[:u + 1] = zl
    last, nc = 1, len(circ.cols)
    for s_ in range(circ.n_sets):
        cs = range(s_ * circ.chunk_len, min((s_ + 1) * circ.chunk_len, nc))
        num, den = [], []
        for r in range(u):
            a = b = 1
            for c in cs:
                a = a * (circ.cols[c][r] + beta * pow(DELTA, c, R) * pow(w, r, R) + gamma) % R
                b = b * (circ.cols[c][r] + beta * circ.sigma[c][r] + gamma) % R
            num.append(a)
            den.append(b)
        gp = O.fr_to_ints(h2v.grand_product(fr_arr(num + [1]), fr_arr(den + [1])))
        circ.z[s_][:u + 1] = [last * v % R for v in gp]
        last = circ.z[s_][u]
    assert last == 1
    coeff, hq, _ = _gpu_quotient(h2v, circ)
    _check_identity(circ, coeff, hq, seed=k + 50)
