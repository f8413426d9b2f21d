"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY; parity unpinned).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  See bn254_oracle.c for the reference citations of each function.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "liboracle.so")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_DIR, "bn254_oracle.c")):
        subprocess.check_call(["make", "-C", _DIR, "-s"] + (["-B"] if force else []))
    return _SO


def _load():
    build()
    try:
        return C.CDLL(_SO)
    except OSError:
        build(force=True)
        return C.CDLL(_SO)


_lib = _load()
_u64p = C.POINTER(C.c_uint64)


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _sig(name, *argtypes, res=None):
    f = getattr(_lib, name)
    f.argtypes = list(argtypes)
    f.restype = res
    return f


_best_multiexp = _sig("orc_best_multiexp", _u64p, _u64p, C.c_size_t, C.c_int, _u64p)
_to_affine = _sig("orc_g1_to_affine", _u64p, _u64p)
_g1_mul = _sig("orc_g1_mul", _u64p, _u64p, _u64p)
_g1_gen = _sig("orc_g1_generator", _u64p)
_g1_on_curve = _sig("orc_g1_is_on_curve", _u64p, res=C.c_int)
_g1_compress = _sig("orc_g1_compress", _u64p, C.POINTER(C.c_uint8))
_g1_add = _sig("orc_g1_add", _u64p, _u64p, _u64p)
_g1_add_mixed = _sig("orc_g1_add_mixed", _u64p, _u64p, _u64p)
_g1_double = _sig("orc_g1_double", _u64p, _u64p)
_gen_bases = _sig("orc_gen_bases", C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, _u64p)
_fr_fill = _sig("orc_fr_fill", C.c_uint64, C.c_int, C.c_uint, C.c_size_t, _u64p)
_to_mont = _sig("orc_to_mont", C.c_int, _u64p, C.c_size_t, _u64p)
_from_mont = _sig("orc_from_mont", C.c_int, _u64p, C.c_size_t, _u64p)
_dot = _sig("orc_fr_dot_affine_index", _u64p, C.c_size_t, C.c_uint64, C.c_uint64, _u64p)
_eval_poly = _sig("orc_fr_eval_poly", _u64p, C.c_size_t, _u64p, _u64p)
_batch_invert = _sig("orc_fr_batch_invert", _u64p, C.c_size_t)
_grand_product = _sig("orc_fr_grand_product", _u64p, _u64p, C.c_size_t, _u64p)
_kate = _sig("orc_fr_kate_division", _u64p, C.c_size_t, _u64p, _u64p)
_srs_setup = _sig("orc_srs_setup", C.c_uint32, _u64p, C.c_int, _u64p, _u64p, res=C.c_int)
_best_fft = _sig("orc_best_fft", _u64p, _u64p, C.c_uint32, C.c_int)
_domain_new = _sig("orc_domain_new", C.c_uint32, C.c_uint32, C.c_void_p, res=C.c_int)
_domain_sizeof = _sig("orc_domain_sizeof", res=C.c_size_t)
_domain_get = _sig("orc_domain_get", C.c_void_p, C.c_int, _u64p)
_domain_ek = _sig("orc_domain_extended_k", C.c_void_p, res=C.c_uint32)
_l2c = _sig("orc_lagrange_to_coeff", C.c_void_p, _u64p, C.c_int)
_c2l = _sig("orc_coeff_to_lagrange", C.c_void_p, _u64p, C.c_int)
_c2e = _sig("orc_coeff_to_extended", C.c_void_p, _u64p, _u64p, C.c_int)
_e2c = _sig("orc_extended_to_coeff", C.c_void_p, _u64p, _u64p, C.c_int)
_dvp = _sig("orc_divide_by_vanishing_poly", C.c_void_p, _u64p)
_q_gates = _sig("orc_quotient_gates", C.c_void_p, _u64p, _u64p, C.c_size_t, _u64p, C.c_size_t, _u64p, C.c_size_t)
_q_perm = _sig("orc_quotient_permutation", C.c_void_p, _u64p, _u64p, _u64p, _u64p, C.c_size_t, C.c_size_t, _u64p, C.c_size_t,
               _u64p, C.c_size_t, _u64p, C.c_size_t, _u64p, _u64p, _u64p, C.c_uint32)
_q_lookup = _sig("orc_quotient_lookup", C.c_void_p, *([_u64p] * 12))
_fr_delta = _sig("orc_fr_delta", _u64p)
_permute_pair = _sig("orc_permute_expression_pair", _u64p, _u64p, C.c_size_t, _u64p, _u64p, res=C.c_int)
for _n in ("fr_mul", "fr_add", "fr_sub", "fq_mul", "fq_add", "fq_sub"):
    _sig("orc_" + _n, _u64p, _u64p, _u64p)
for _n in ("fr_inv", "fq_inv"):
    _sig("orc_" + _n, _u64p, _u64p)

FR, FQ = 0, 1
NCPU = os.cpu_count() or 1


# ----------------------------------------------------------------- int <-> limb helpers
def ints_to_limbs(xs):
    out = np.empty((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        for k in range(4):
            out[i, k] = (x >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(v) << (64 * k) for k, v in enumerate(row)) for row in a]


def to_mont(canon, field=FR):
    canon = np.ascontiguousarray(canon, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(canon)
    _to_mont(field, _p(canon), len(canon), _p(out))
    return out


def from_mont(mont, field=FR):
    mont = np.ascontiguousarray(mont, dtype=np.uint64).reshape(-1, 4)
    out = np.empty_like(mont)
    _from_mont(field, _p(mont), len(mont), _p(out))
    return out


def fr_from_ints(xs):
    return to_mont(ints_to_limbs(xs), FR)


def fr_to_ints(a):
    return limbs_to_ints(from_mont(a, FR))


def field_op(name, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    o = np.empty(4, dtype=np.uint64)
    f = getattr(_lib, "orc_" + name)
    if b is None:
        f(_p(a), _p(o))
    else:
        b = np.ascontiguousarray(b, dtype=np.uint64)
        f(_p(a), _p(b), _p(o))
    return o


# ----------------------------------------------------------------- G1
def g1_generator():
    o = np.empty(8, dtype=np.uint64)
    _g1_gen(_p(o))
    return o


def g1_mul(base_aff, k_int):
    base_aff = np.ascontiguousarray(base_aff, dtype=np.uint64)
    k = ints_to_limbs([k_int]).reshape(4)
    o = np.empty(8, dtype=np.uint64)
    _g1_mul(_p(base_aff), _p(k), _p(o))
    return o


def g1_to_affine(jac):
    jac = np.ascontiguousarray(jac, dtype=np.uint64)
    o = np.empty(8, dtype=np.uint64)
    _to_affine(_p(jac), _p(o))
    return o


def g1_affine_to_ints(aff):
    """Montgomery affine limbs -> (x, y) python ints, or None for the identity."""
    c = limbs_to_ints(from_mont(np.asarray(aff, dtype=np.uint64).reshape(2, 4), FQ))
    return None if c == [0, 0] else (c[0], c[1])


def g1_affine_from_ints(pt):
    if pt is None:
        return np.zeros(8, dtype=np.uint64)
    return to_mont(ints_to_limbs(list(pt)), FQ).reshape(8)


def g1_is_on_curve(aff):
    aff = np.ascontiguousarray(aff, dtype=np.uint64)
    return bool(_g1_on_curve(_p(aff)))


def g1_compress(aff):
    aff = np.ascontiguousarray(aff, dtype=np.uint64)
    o = (C.c_uint8 * 32)()
    _g1_compress(_p(aff), o)
    return bytes(o)


def g1_add(a_jac, b_jac):
    o = np.empty(12, dtype=np.uint64)
    _g1_add(_p(np.ascontiguousarray(a_jac)), _p(np.ascontiguousarray(b_jac)), _p(o))
    return o


def g1_add_mixed(a_jac, b_aff):
    o = np.empty(12, dtype=np.uint64)
    _g1_add_mixed(_p(np.ascontiguousarray(a_jac)), _p(np.ascontiguousarray(b_aff)), _p(o))
    return o


def g1_double(a_jac):
    o = np.empty(12, dtype=np.uint64)
    _g1_double(_p(np.ascontiguousarray(a_jac)), _p(o))
    return o


def gen_bases(n, a=0x9E3779B97F4A7C15 >> 2, b=0x632BE59BD9B4E019 >> 2, threads=NCPU):
    """bases[i] = (a*i + b) * G, affine Montgomery (n, 8) u64 -- SURVEY.md 8d config 5."""
    out = np.empty((n, 8), dtype=np.uint64)
    _gen_bases(a, b, n, threads, _p(out))
    return out


def fr_fill(n, seed, mode=0, lookup_bits=12):
    """mode 0: uniform in [0, r); mode 1: witness-like skewed distribution. Montgomery (n, 4) u64."""
    out = np.empty((n, 4), dtype=np.uint64)
    _fr_fill(seed, mode, lookup_bits, n, _p(out))
    return out


def msm_closed_form(scalars_mont, a=0x9E3779B97F4A7C15 >> 2, b=0x632BE59BD9B4E019 >> 2):
    """(sum_i s_i (a i + b) mod r) * G as affine Montgomery limbs -- independent of any MSM algorithm."""
    scalars_mont = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    k = np.empty(4, dtype=np.uint64)
    _dot(_p(scalars_mont), len(scalars_mont), a, b, _p(k))
    o = np.empty(8, dtype=np.uint64)
    _g1_mul(_p(g1_generator()), _p(k), _p(o))
    return o


def best_multiexp(coeffs, bases, threads=NCPU):
    """halo2-axiom arithmetic.rs best_multiexp; returns the Jacobian point (12 u64)."""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    assert len(coeffs) == len(bases)
    o = np.empty(12, dtype=np.uint64)
    _best_multiexp(_p(coeffs), _p(bases), len(coeffs), threads, _p(o))
    return o


def best_multiexp_affine(coeffs, bases, threads=NCPU):
    return g1_to_affine(best_multiexp(coeffs, bases, threads))


# ----------------------------------------------------------------- FFT / domain
def best_fft(a, omega, log_n, threads=NCPU):
    """halo2-axiom arithmetic.rs best_fft: in place, natural order in and out. Returns a new array."""
    a = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    assert len(a) == 1 << log_n
    omega = np.ascontiguousarray(omega, dtype=np.uint64)
    _best_fft(_p(a), _p(omega), log_n, threads)
    return a


def fr_eval_poly(a, x):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    o = np.empty(4, dtype=np.uint64)
    _eval_poly(_p(a), len(a), _p(np.ascontiguousarray(x, dtype=np.uint64)), _p(o))
    return o


def srs_setup(k, s_mont, threads=NCPU):
    """ParamsKZG::setup(k, rng) with the secret s given: returns (g, g_lagrange), each (2^k, 8)."""
    n = 1 << k
    g = np.zeros((n, 8), dtype=np.uint64)
    gl = np.zeros((n, 8), dtype=np.uint64)
    if _srs_setup(k, _p(np.ascontiguousarray(s_mont, dtype=np.uint64)), threads, _p(g), _p(gl)) != 0:
        raise ValueError("unsupported k")
    return g, gl


def fr_batch_invert(a):
    """ff::BatchInvert::batch_invert: non-zero elements inverted, zeros untouched."""
    a = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 4)
    _batch_invert(_p(a), len(a))
    return a


def fr_grand_product(num, den):
    """z[0] = 1, z[i+1] = z[i] * num[i] / den[i] (permutation / lookup running product)."""
    num = np.ascontiguousarray(num, dtype=np.uint64).reshape(-1, 4)
    den = np.ascontiguousarray(den, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(num)
    _grand_product(_p(num), _p(den), len(num), _p(out))
    return out


def fr_kate_division(a, b):
    """arithmetic.rs kate_division: quotient of a(X) by (X - b)."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((max(len(a) - 1, 0), 4), dtype=np.uint64)
    if len(a) > 1:
        _kate(_p(a), len(a), _p(np.ascontiguousarray(b, dtype=np.uint64)), _p(out))
    return out


class EvaluationDomain:
    """halo2-axiom poly/domain.rs EvaluationDomain::new(j, k) and its transforms."""

    _NAMES = ["omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv",
              "ifft_divisor", "extended_ifft_divisor"]

    def __init__(self, j, k):
        self._buf = C.create_string_buffer(_domain_sizeof())
        self._d = C.cast(self._buf, C.c_void_p)
        if _domain_new(j, k, self._d) != 0:
            raise ValueError("unsupported domain")
        self.k = k
        self.n = 1 << k
        self.j = j
        self.extended_k = int(_domain_ek(self._d))
        for i, name in enumerate(self._NAMES):
            o = np.empty(4, dtype=np.uint64)
            _domain_get(self._d, i, _p(o))
            setattr(self, name, o)
        self.t_evaluations = []
        for i in range(1 << (self.extended_k - k)):
            o = np.empty(4, dtype=np.uint64)
            _domain_get(self._d, 8 + i, _p(o))
            self.t_evaluations.append(o)

    def lagrange_to_coeff(self, a, threads=NCPU):
        a = np.array(a, dtype=np.uint64, copy=True).reshape(self.n, 4)
        _l2c(self._d, _p(a), threads)
        return a

    def coeff_to_lagrange(self, a, threads=NCPU):
        a = np.array(a, dtype=np.uint64, copy=True).reshape(self.n, 4)
        _c2l(self._d, _p(a), threads)
        return a

    def coeff_to_extended(self, a, threads=NCPU):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(self.n, 4)
        out = np.empty((1 << self.extended_k, 4), dtype=np.uint64)
        _c2e(self._d, _p(a), _p(out), threads)
        return out

    def extended_to_coeff(self, a, threads=NCPU):
        a = np.array(a, dtype=np.uint64, copy=True).reshape(1 << self.extended_k, 4)
        out = np.empty((self.n * (self.j - 1), 4), dtype=np.uint64)
        _e2c(self._d, _p(a), _p(out), threads)
        return out

    def divide_by_vanishing_poly(self, a):
        a = np.array(a, dtype=np.uint64, copy=True).reshape(1 << self.extended_k, 4)
        _dvp(self._d, _p(a))
        return a

    # --- evaluate_h row loops (plonk/evaluation.rs [UPSTREAM]); h is updated and returned, columns are (n_cols, 2^extended_k, 4)
    def _ext(self, a, cols=None):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        shape = (1 << self.extended_k, 4) if cols is None else (cols, 1 << self.extended_k, 4)
        if a.shape != shape:
            raise ValueError(f"expected {shape}, got {a.shape}")
        return a

    def quotient_gates(self, h, y, q, a):
        h = np.array(self._ext(h), copy=True)
        n_gates = len(q)
        if n_gates:
            q, a = self._ext(q, n_gates), self._ext(a, n_gates)
            _q_gates(self._d, _p(h), _p(_one(y)), n_gates, _p(q), 1 << self.extended_k, _p(a), 1 << self.extended_k)
        return h

    def quotient_permutation(self, h, y, beta, gamma, chunk_len, cols, sigma, z, l0, l_last, l_active, blinding_factors):
        h = np.array(self._ext(h), copy=True)
        n_cols = len(cols)
        if n_cols:
            n_sets = (n_cols + chunk_len - 1) // chunk_len
            e = 1 << self.extended_k
            _q_perm(self._d, _p(h), _p(_one(y)), _p(_one(beta)), _p(_one(gamma)), n_cols, chunk_len, _p(self._ext(cols, n_cols)), e,
                    _p(self._ext(sigma, n_cols)), e, _p(self._ext(z, n_sets)), e, _p(self._ext(l0)), _p(self._ext(l_last)),
                    _p(self._ext(l_active)), blinding_factors)
        return h

    def quotient_lookup(self, h, y, beta, gamma, inp, table, perm_input, perm_table, z, l0, l_last, l_active):
        h = np.array(self._ext(h), copy=True)
        _q_lookup(self._d, _p(h), _p(_one(y)), _p(_one(beta)), _p(_one(gamma)), _p(self._ext(inp)), _p(self._ext(table)),
                  _p(self._ext(perm_input)), _p(self._ext(perm_table)), _p(self._ext(z)), _p(self._ext(l0)), _p(self._ext(l_last)),
                  _p(self._ext(l_active)))
        return h


def _one(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64).reshape(4))


def fr_delta():
    """Fr::DELTA (Montgomery limbs)."""
    o = np.empty(4, dtype=np.uint64)
    _fr_delta(_p(o))
    return o


def permute_expression_pair(inp, table):
    """plonk/lookup/prover.rs permute_expression_pair on the usable rows -> (permuted_input, permuted_table);
    ValueError when an input value is missing from the table (upstream: Error::ConstraintSystemFailure)."""
    inp = np.ascontiguousarray(inp, dtype=np.uint64).reshape(-1, 4)
    table = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1, 4)
    if inp.shape != table.shape:
        raise ValueError("input and table must have the same number of usable rows")
    a, s_ = np.zeros_like(inp), np.zeros_like(inp)
    if len(inp) and _permute_pair(_p(inp), _p(table), len(inp), _p(a), _p(s_)) != 0:
        raise ValueError("input value not in table")
    return a, s_
