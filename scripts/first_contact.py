"""First GPU contact: stage-by-stage diagnostics (field -> group -> NTT -> MSM), each isolated."""
import os, sys, time, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_vectordb_b200 as h
from oracle import oracle as O

def section(name, fn):
    t = time.time()
    try:
        r = fn()
        print(f"[{name}] OK {r if r is not None else ''} ({time.time()-t:.2f}s)", flush=True)
        return True
    except Exception as e:
        print(f"[{name}] FAIL {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
        return False

h.init(0)
print("devices", h.device_count(), flush=True)

def t_field():
    rng = np.random.default_rng(1)
    for field, nm in ((0, "fr"), (1, "fq")):
        a = O.fr_fill(4096, 11 + field); b = O.fr_fill(4096, 13 + field)
        if field == 1:  # reinterpret canonical-ish values below p: Fr values < r < p are valid Fq limbs
            pass
        for op, opn in ((0, "mul"), (1, "add"), (2, "sub")):
            got = h.selftest_field(field, op, a, b)
            exp = np.stack([O.field_op(f"{nm}_{opn}", a[i], b[i]) for i in range(len(a))])
            bad = np.nonzero((got != exp).any(axis=1))[0]
            assert len(bad) == 0, f"{nm}_{opn}: {len(bad)} mismatches, first {bad[:4]} got {got[bad[0]]} exp {exp[bad[0]]}"
        got = h.selftest_field(field, 3, a[:64])
        exp = np.stack([O.field_op(f"{nm}_inv", a[i]) for i in range(64)])
        assert (got == exp).all(), f"{nm}_inv mismatch"
section("field", t_field)

def t_group():
    G = O.g1_generator()
    pts = O.gen_bases(64)
    p, q = pts[:32].copy(), pts[32:].copy()
    q[1] = p[1]                      # doubling
    x, y = O.g1_affine_to_ints(p[2]); q[2] = O.g1_affine_from_ints((x, (-y) % O_P))  # inverse
    q[3] = 0; p[4] = 0
    for mode in (0, 1, 2):
        got = h.selftest_group(mode, p, q)
        for i in range(32):
            if mode == 2:
                e = O.g1_to_affine(O.g1_double(np.concatenate([p[i], O.to_mont(O.ints_to_limbs([1 if p[i].any() else 0]), 1).reshape(4)])))
            else:
                pj = np.concatenate([p[i], O.to_mont(O.ints_to_limbs([1 if p[i].any() else 0]), 1).reshape(4)])
                e = O.g1_to_affine(O.g1_add_mixed(pj, q[i]))
            assert (got[i] == e).all(), f"mode {mode} idx {i}"
from oracle import pyref
O_P = pyref.P
section("group", t_group)

print("imad peak", section("imad", lambda: f"{h.imad_peak()/1e12:.2f} T wide-MAC/s"), flush=True)

def t_ntt(L):
    def f():
        a = O.fr_fill(1 << L, 100 + L)
        om = O.to_mont(O.ints_to_limbs([pyref.omega_for(L)]))[0]
        t = time.time(); got = h.best_fft(a, om, L); t1 = time.time() - t
        if L <= 20:
            exp = O.best_fft(a, om, L)
            bad = np.nonzero((got != exp).any(axis=1))[0]
            assert len(bad) == 0, f"{len(bad)} mismatches first {bad[:8]}"
        else:
            xs = [3, 12345, (1 << L) - 1]
            for j in xs:
                x = O.to_mont(O.ints_to_limbs([pow(pyref.omega_for(L), j, pyref.R)]))[0]
                assert (O.fr_eval_poly(a, x) == got[j]).all(), f"spot {j}"
        return f"{t1*1e3:.1f} ms e2e"
    return f
for L in (0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 13, 16, 18, 20, 22):
    section(f"ntt 2^{L}", t_ntt(L))

def t_domain(k):
    def f():
        d = h.EvaluationDomain(4, k); od = O.EvaluationDomain(4, k)
        for nm in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv", "ifft_divisor", "extended_ifft_divisor"):
            assert (getattr(d, nm) == getattr(od, nm)).all(), nm
        for a, b in zip(d.t_evaluations, od.t_evaluations): assert (a == b).all()
        a = O.fr_fill(1 << k, 7 + k, mode=1)
        assert (d.lagrange_to_coeff(a) == od.lagrange_to_coeff(a)).all(), "l2c"
        assert (d.coeff_to_lagrange(a) == od.coeff_to_lagrange(a)).all(), "c2l"
        e = d.coeff_to_extended(a); oe = od.coeff_to_extended(a)
        assert (e == oe).all(), "c2e"
        assert (d.extended_to_coeff(e) == od.extended_to_coeff(oe)).all(), "e2c"
        assert (d.divide_by_vanishing_poly(e) == od.divide_by_vanishing_poly(oe)).all(), "dvp"
        fused = d.transform_batch(h.OP_DIVIDE_BY_VANISHING, [e, e])
        exp = od.extended_to_coeff(od.divide_by_vanishing_poly(oe))
        assert (fused[0] == exp).all() and (fused[1] == exp).all(), "fused"
    return f
for k in (3, 5, 8, 11, 13, 16):
    section(f"domain k={k}", t_domain(k))

def t_msm_raw(n, mode=0):
    def f():
        b = O.gen_bases(n); s = O.fr_fill(n, 500 + n, mode=mode)
        t = time.time(); got = O.g1_to_affine(h.best_multiexp(s, b)); t1 = time.time() - t
        exp = O.msm_closed_form(s)
        assert (got == exp).all(), f"got {got} exp {exp}"
        return f"{t1*1e3:.1f} ms e2e"
    return f
for n in (1, 2, 7, 64, 1000, 1 << 13, 1 << 16):
    section(f"msm_raw n={n}", t_msm_raw(n))
section("msm_raw skew 2^14", t_msm_raw(1 << 14, 1))

def t_commit(k, ncols, mode=0):
    def f():
        n = 1 << k
        b = O.gen_bases(n)
        t = time.time(); srs = h.ParamsKZG(k, None, b); t0 = time.time() - t
        cols = [O.fr_fill(n, 900 + i, mode=mode) for i in range(ncols)]
        t = time.time(); got = srs.commit_batch(cols); t1 = time.time() - t
        for i in range(ncols):
            assert (got[i] == O.msm_closed_form(cols[i])).all(), f"col {i}"
        short = srs.commit_lagrange(cols[0][: n // 2 + 3])
        assert (short == O.msm_closed_form(cols[0][: n // 2 + 3])).all(), "short poly"
        srs.close()
        return f"srs {t0*1e3:.0f} ms, batch {t1*1e3:.1f} ms"
    return f
section("commit k=4", t_commit(4, 3))
section("commit k=10", t_commit(10, 4))
section("commit k=13", t_commit(13, 8))
section("commit k=13 skew", t_commit(13, 4, 1))
section("commit k=16", t_commit(16, 4))
print("launches", h.launch_count(), flush=True)
