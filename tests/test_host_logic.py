"""CPU: host-side logic of the product -- the C ABI library loads and exports every symbol declared in
include/h2v.h, fails loudly without a GPU, the host instantiations of the shared field / curve code
match the oracle, and the Python models of the kernels' index arithmetic reproduce the definitions."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

import kernel_model as KM
from oracle import oracle as O
from oracle import pyref as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from halo2_vectordb_b200 import build

    build.build()
    return build


def test_abi_exports_every_declared_symbol(built):
    import halo2_vectordb_b200 as h

    hdr = open(os.path.join(ROOT, "include", "h2v.h")).read()
    declared = sorted(set(re.findall(r"\b(h2v_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = h.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(h.ABI_SYMBOLS) == declared
    assert b"sm_100a" in L.h2v_version()


def test_no_cpu_fallback(built):
    """Without a CUDA device every compute entry point must fail loudly (never fall back to the oracle)."""
    import halo2_vectordb_b200 as h

    if h.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(h.H2VError):
        h.init(0)
    with pytest.raises(h.H2VError):
        h.best_fft(np.zeros((4, 4), dtype=np.uint64), np.zeros(4, dtype=np.uint64), 2)
    with pytest.raises(h.H2VError):
        h.best_multiexp(np.zeros((2, 4), dtype=np.uint64), np.zeros((2, 8), dtype=np.uint64))
    with pytest.raises(h.H2VError):
        h.EvaluationDomain(4, 4)
    with pytest.raises(h.H2VError):
        h.ParamsKZG(2, np.zeros((4, 8), dtype=np.uint64), None)
    one = np.ones((3, 4), dtype=np.uint64)
    with pytest.raises(h.H2VError):
        h.permute_expression_pair(one, one)
    with pytest.raises(h.H2VError):
        h.grand_product(one, one)
    with pytest.raises(h.H2VError):
        h.g1_sum(np.zeros((2, 8), dtype=np.uint64))
    with pytest.raises(h.H2VError):
        h.eval_polynomial_batch([one], one[:1])
    with pytest.raises(h.H2VError):
        h.DeviceBuffer(64)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "halo2_vectordb_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "liboracle" not in src and "bn254_oracle" not in src, f
    hdr = open(os.path.join(ROOT, "include", "h2v.h")).read()
    assert "oracle" not in hdr.lower().replace("no cpu fallback", "")


def _hostcheck(built):
    L = C.CDLL(built.HOSTCHECK)
    u = C.POINTER(C.c_uint64)
    L.h2v_host_field.argtypes = [C.c_int, C.c_int, u, u, u]
    L.h2v_host_group.argtypes = [C.c_int, u, u, u]
    return L, (lambda a: np.ascontiguousarray(a, dtype=np.uint64).ctypes.data_as(u))


def test_shared_field_code_host_instantiation(built):
    L, p = _hostcheck(built)
    rnd = random.Random(1)
    for field, mod, nm in ((0, P.R, "fr"), (1, P.P, "fq")):
        for t in range(100):
            a, b = rnd.randrange(mod), rnd.randrange(mod)
            if t == 0:
                a = b = mod - 1
            if t == 1:
                a = 0
            am, bm = O.to_mont(O.ints_to_limbs([a]), field)[0], O.to_mont(O.ints_to_limbs([b]), field)[0]
            o = np.empty(4, dtype=np.uint64)
            for op, opn in enumerate(["mul", "add", "sub"]):
                L.h2v_host_field(field, op, p(am), p(bm), p(o))
                assert (o == O.field_op(f"{nm}_{opn}", am, bm)).all()
            if a:
                L.h2v_host_field(field, 3, p(am), None, p(o))
                assert (o == O.field_op(f"{nm}_inv", am)).all()
            L.h2v_host_field(field, 4, p(O.ints_to_limbs([a])[0]), None, p(o))
            assert (o == am).all()


def test_shared_group_law_host_instantiation(built):
    L, p = _hostcheck(built)
    rnd = random.Random(2)
    G = O.g1_generator()
    pts = [O.g1_mul(G, rnd.randrange(1, P.R)) for _ in range(12)]
    ident = np.zeros(8, dtype=np.uint64)
    neg = lambda a: O.g1_affine_from_ints((lambda xy: (xy[0], (-xy[1]) % P.P))(O.g1_affine_to_ints(a)))
    cases = [(pts[i], pts[i + 1]) for i in range(0, 10, 2)]
    cases += [(pts[0], pts[0]), (pts[1], neg(pts[1])), (ident, pts[2]), (pts[3], ident), (ident, ident)]
    for a, b in cases:
        ea = P.g1_add(O.g1_affine_to_ints(a), O.g1_affine_to_ints(b))
        for mode in (0, 1):
            o = np.empty(8, dtype=np.uint64)
            L.h2v_host_group(mode, p(a), p(b), p(o))
            assert O.g1_affine_to_ints(o) == ea
        o = np.empty(8, dtype=np.uint64)
        L.h2v_host_group(2, p(a), p(b), p(o))
        assert O.g1_affine_to_ints(o) == P.g1_add(O.g1_affine_to_ints(a), O.g1_affine_to_ints(a))
        o = np.empty(12, dtype=np.uint64)
        L.h2v_host_group(3, p(a), p(b), p(o))
        assert O.g1_affine_to_ints(O.g1_to_affine(o)) == ea


@pytest.mark.parametrize("L,max_s,max_log", [(3, 9, 11), (5, 9, 11), (9, 9, 11), (10, 9, 11), (6, 3, 5), (9, 3, 4),
                                             (8, 4, 6), (10, 4, 6), (10, 5, 7)])
def test_ntt_index_model(L, max_s, max_log):
    rnd = random.Random(L)
    a = [rnd.randrange(P.R) for _ in range(1 << L)]
    w = P.omega_for(L)
    assert KM.run_ntt(a, L, w, max_s=max_s, max_log=max_log) == P.dft_naive(a, w)


def test_ntt_model_coset_ops():
    rnd = random.Random(4)
    for k, max_s, max_log in ((4, 9, 11), (4, 3, 5), (5, 4, 6)):
        d = P.Domain(4, k)
        a = [rnd.randrange(P.R) for _ in range(d.n)]
        ext = KM.run_ntt(a, d.extended_k, d.omega_ext, n_in=d.n, pre=[1, d.g_coset, d.g_coset_inv], max_s=max_s, max_log=max_log)
        assert ext == d.coeff_to_extended(a)
        post = [d.extended_ifft_divisor, d.extended_ifft_divisor * d.g_coset_inv % P.R, d.extended_ifft_divisor * d.g_coset % P.R]
        fused = KM.run_ntt(ext, d.extended_k, d.omega_ext_inv, n_out=3 * d.n, pre=d.t_evaluations, post=post,
                           max_s=max_s, max_log=max_log)
        assert fused[: 3 * d.n] == d.extended_to_coeff(d.divide_by_vanishing_poly(ext))


@pytest.mark.parametrize("n,c,pre,chunk,ncols,dist", [(64, 4, True, 4, 1, "u"), (64, 4, False, 4, 2, "u"),
                                                       (100, 5, True, 7, 2, "skew"), (200, 6, False, 5, 1, "skew"),
                                                       (33, 3, True, 32, 3, "u"), (50, 7, True, 4, 1, "edge"),
                                                       (16, 13, True, 8, 1, "u"), (40, 2, False, 6, 1, "u")])
def test_msm_pipeline_model(n, c, pre, chunk, ncols, dist):
    rnd = random.Random(n * 31 + c)
    pts = [rnd.randrange(1, P.R) for _ in range(n)]
    cols = []
    for _ in range(ncols):
        if dist == "u":
            s = [rnd.randrange(P.R) for _ in range(n)]
        elif dist == "skew":
            s = [rnd.choice([0, 1, 1, 1, rnd.randrange(4096), P.R - rnd.randrange(1, 1000), rnd.randrange(P.R)]) for _ in range(n)]
        else:
            s = [rnd.choice([0, P.R - 1, 1, P.R - 2, 1 << 253]) for _ in range(n)]
        cols.append(s)
    exp = [sum(a * b for a, b in zip(s, pts)) % P.R for s in cols]
    for log_seg in (3, 5):
        assert KM.msm_model(cols, pts, c, pre, chunk, ncols, log_seg) == exp


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(built, tmp_path):
    import subprocess

    import halo2_vectordb_b200 as h

    exe = str(tmp_path / "cpp_mirror_check")
    libdir = os.path.join(ROOT, "halo2_vectordb_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cpp_mirror_check.cpp"), "-L" + libdir, "-lh2v", "-Wl,-rpath," + libdir])
    if h.device_count() == 0:
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and "no device" in out.stdout


def test_batch_inverse_tree_model():
    rnd = random.Random(11)
    for n, g in ((1, 4), (5, 4), (17, 4), (64, 4), (100, 8), (33, 32), (257, 16)):
        xs = [rnd.randrange(1, P.R) for _ in range(n)]
        inv = KM.batch_inverse_tree(xs, g)
        assert all(a * b % P.R == 1 for a, b in zip(xs, inv)), (n, g)


@pytest.mark.parametrize("n,c,pre,chunk,ncols,dist,rounds,K,G", [
    (64, 4, True, 4, 1, "u", 3, 4, 4), (64, 4, False, 4, 2, "u", 2, 3, 2), (100, 5, True, 7, 2, "skew", 4, 5, 4),
    (200, 6, False, 5, 1, "skew", 6, 8, 4), (33, 3, True, 32, 3, "u", 1, 2, 8), (40, 2, False, 6, 1, "dup", 5, 4, 3)])
def test_msm_batch_affine_rounds_model(n, c, pre, chunk, ncols, dist, rounds, K, G):
    """pair rounds (msm_affine.cuh): slot <-> bucket walking in both directions, the inversion tree, and the
    identity / doubling / cancelling cases, on the integer model of the group."""
    rnd = random.Random(n * 7 + rounds)
    pts = [rnd.randrange(1, P.R) for _ in range(n)]
    if dist == "dup":
        pts = [rnd.choice([5, 5, (-5) % P.R, None, 10, (-10) % P.R, 7]) for _ in range(n)]
    cols = []
    for _ in range(ncols):
        if dist in ("u", "dup"):
            s = [rnd.randrange(P.R) for _ in range(n)]
        else:
            s = [rnd.choice([0, 1, 1, 1, rnd.randrange(4096), P.R - rnd.randrange(1, 1000), rnd.randrange(P.R)]) for _ in range(n)]
        cols.append(s)
    exp = [sum(a * (b or 0) for a, b in zip(s, pts)) % P.R for s in cols]
    assert KM.msm_model_affine(cols, pts, c, pre, chunk, ncols, rounds, K, G) == exp


def test_wire_format_helpers(built):
    """h2v_g1_to_bytes / h2v_fr_to_repr run on the host (no GPU needed): against the oracle and first principles."""
    import halo2_vectordb_b200 as h

    G = O.g1_generator()
    pts = [O.g1_mul(G, k) for k in (1, 2, 3, 0xDEADBEEF, P.R - 1)] + [np.zeros(8, dtype=np.uint64)]
    got = h.g1_to_bytes(np.stack(pts))
    for g, p in zip(got, pts):
        assert g == O.g1_compress(p) == P.g1_compress(O.g1_affine_to_ints(p))
    assert got[0] == (1).to_bytes(32, "little")                      # G = (1, 2): y even, no flag
    assert got[-1] == bytes(32)
    vals = [0, 1, P.R - 1, 0x1234567890ABCDEF << 100]
    assert h.fr_to_repr(O.fr_from_ints(vals)) == [v.to_bytes(32, "little") for v in vals]


def test_shoup_product_model():
    """Line-by-line integer model of csrc/shoup.cuh: the companion w' = low256(w_mont * (-r^-1 mod 2^256)) equals
    floor(w 2^256 / r); the quotient estimate built from the partial products a_i w'_j with i + j >= 6 is q or q - 1;
    t = low256(a w + q_hat (2^256 - r)) is a w - q_hat r < 3r, and the top-limb correction leaves t < 2r, t = a w (mod r)."""
    import random

    from oracle import pyref as P
    r, B, M = P.R, 1 << 256, (1 << 32) - 1
    ninv = (-pow(r, -1, B)) % B
    lit = [0xefffffff, 0xc2e1f593, 0x4c6911b3, 0x6586864b, 0x99062391, 0xe39a9828, 0x0d8341b2, 0x73f82f1d]   # fr_ninv256
    assert ninv == sum(v << (32 * i) for i, v in enumerate(lit))
    assert r < (1 << 254) <= 2 * r                   # the correction threshold lies in (r, 2r]
    rnd = random.Random(11)
    ws = [0, 1, r - 1, r - 2, r >> 1] + [rnd.randrange(r) for _ in range(300)]
    As = [0, 1, B - 1, B - 2, 4 * r - 1, 1 << 255, r, 2 * r] + [rnd.randrange(B) for _ in range(300)]
    worst = 0
    for w in ws:
        wp = ((w * B % r) * ninv) % B
        assert wp == (w * B) // r
        for a in As[:40] if w > 5 else As:
            al = [(a >> (32 * i)) & M for i in range(8)]
            wl = [(wp >> (32 * i)) & M for i in range(8)]
            s_trunc = sum(al[i] * wl[j] << (32 * (i + j)) for i in range(8) for j in range(8) if i + j >= 6)
            q_hat, q = s_trunc >> 256, (a * wp) >> 256
            assert q - 1 <= q_hat <= q
            t = (a * w + q_hat * (B - r)) % B
            assert t == a * w - q_hat * r and t < 3 * r
            worst = max(worst, t / r)
            if (t >> 224) >= 0x40000000:
                t -= r
            assert 0 <= t < 2 * r and t % r == (a * w) % r
    assert worst < 2.76
