"""CPU restatement of the reference's prove / verify flow for halo2-base-shaped circuits (TEST INFRASTRUCTURE ONLY;
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import anything under oracle/).

PARITY UNPINNED: the flow lives in un-vendored, un-locked git dependencies of the reference
(/root/reference/Cargo.toml:19-28) and there is no Rust toolchain here, so this file restates, from the published
algorithms, what `gen_snark_shplonk` (/root/reference/src/scaffold/mod.rs:296) and `verify_proof`
(mod.rs:312-319) run [UPSTREAM, halo2-axiom = PSE halo2 v2023_02_02 lineage]:

  create_proof   plonk/prover.rs; plonk/lookup/prover.rs (commit_permuted, commit_product, evaluate, open);
                 plonk/permutation/prover.rs (commit, evaluate, open); plonk/vanishing/prover.rs;
                 plonk/evaluation.rs evaluate_h; poly/kzg/multiopen/shplonk/{prover.rs, ../shplonk.rs}
  verify_proof   plonk/verifier.rs; plonk/{lookup,permutation,vanishing}/verifier.rs;
                 poly/kzg/multiopen/shplonk/verifier.rs.  The final pairing check e(W', [s]_2) = e(F + u W', [1]_2) is
                 done in G1 with the setup secret s (known for gen_srs' "unsafe" seed-zero setup): s * W' == F + u * W'.

The heavy vector work goes through the C restatement (liboracle: best_multiexp, best_fft, evaluate_h row loops,
permute_expression_pair); the control flow, the transcript, the RNG draw order and the multi-open point sets are
plain Python integers here.  The product's implementation of the same flow is C++/CUDA (csrc/prover.cu) -- written
separately; tests compare the two byte streams.

Constraint-system description (`cs`, a dict; identical to halo2_vectordb_b200.ProvingKey's):
  k, degree, blinding_factors, n_advice, n_fixed, n_instance,
  gates [(advice column, selector fixed column)]      q * (a + a(wX) a(w^2 X) - a(w^3 X))
  lookups [(input advice column, table fixed column)]
  permutation [(kind, index)]  kind 0 advice / 1 fixed / 2 instance, in cs.permutation order
  advice_queries / fixed_queries [(column, rotation)] in cs order; instance_queries likewise (verifier only)
"""
import numpy as np

from . import oracle as O
from . import pyref as P
from .transcript import ChaCha20Rng, PoseidonTranscript

R = P.R
DELTA = pow(P.GEN, 1 << P.S, R)


def _np(ints):
    return O.fr_from_ints([int(v) % R for v in ints])


def _ints(arr):
    return O.fr_to_ints(arr)


def _one(v):
    return _np([v])[0]


def _commit(bases, col_ints):
    """ParamsKZG::commit / commit_lagrange -> affine point as ints (None = identity)"""
    return O.g1_affine_to_ints(O.best_multiexp_affine(_np(col_ints), bases[:len(col_ints)]))


def _lagrange_interpolate(xs, ys):
    m = len(xs)
    out = [0] * m
    for j in range(m):
        num = [1]
        den = 1
        for k in range(m):
            if k == j:
                continue
            nx = [0] * (len(num) + 1)
            for t, c in enumerate(num):
                nx[t + 1] = (nx[t + 1] + c) % R
                nx[t] = (nx[t] - c * xs[k]) % R
            num = nx
            den = den * (xs[j] - xs[k]) % R
        sc = ys[j] * pow(den, -1, R) % R
        for t in range(m):
            out[t] = (out[t] + num[t] * sc) % R
    return out


def _rotate(x, omega, rot):
    return x * pow(omega, rot, R) % R


def _intermediate_sets(queries):
    """poly/kzg/multiopen/shplonk.rs construct_intermediate_sets.  queries: [(commitment key, point, eval)].
    Returns (rotation sets [(sorted points, [(key, evals at the points)])], sorted super point set)."""
    com_points = []           # [(key, set of points)] in order of first appearance
    index = {}
    evals = {}
    for key, pt, ev in queries:
        if key not in index:
            index[key] = len(com_points)
            com_points.append((key, set()))
        com_points[index[key]][1].add(pt)
        evals[(key, pt)] = ev
    sets = []                 # [(frozenset of points, [keys])] in order of first appearance
    for key, pts in com_points:
        fs = frozenset(pts)
        for s in sets:
            if s[0] == fs:
                s[1].append(key)
                break
        else:
            sets.append((fs, [key]))
    out = []
    for fs, keys in sets:
        pts = sorted(fs)       # BTreeSet<Fr>: numeric order of the canonical values
        out.append((pts, [(key, [evals[(key, p)] for p in pts]) for key in keys]))
    return out, sorted({pt for _, pt, _ in queries})


class Params:
    """ParamsKZG<Bn256> as far as the G1 side goes: g (monomial), g_lagrange; `s` is the setup secret when known."""

    def __init__(self, k, g, g_lagrange, s=None):
        self.k, self.n, self.g, self.g_lagrange, self.s = k, 1 << k, g, g_lagrange, s

    @classmethod
    def setup(cls, k, s):
        g, gl = O.srs_setup(k, _one(s))
        return cls(k, g, gl, s)


def keygen_vk(params, cs, fixed, sigma):
    """the verifying key's commitments: `commit_lagrange` of every fixed column and every permutation polynomial"""
    return {"fixed": [_commit(params.g_lagrange, c) for c in fixed], "sigma": [_commit(params.g_lagrange, c) for c in sigma]}


# =================================================================================================== create_proof
def create_proof(params, cs, fixed, sigma, vk_repr, advice, instances, rng_seed=bytes(32), trace=None):
    """fixed / sigma / advice: lists of n ints per column (Lagrange basis); instances: list of lists of ints.
    Returns the proof bytes.  `trace` (a dict) receives intermediate values for debugging."""
    k, n, bf, degree = cs["k"], 1 << cs["k"], cs["blinding_factors"], cs["degree"]
    u = n - (bf + 1)
    dom = O.EvaluationDomain(degree, k)
    ne = 1 << dom.extended_k
    omega = _ints(dom.omega)[0]
    T = PoseidonTranscript()
    rng = ChaCha20Rng(rng_seed)
    A, L = cs["n_advice"], len(cs["lookups"])
    perm = cs["permutation"]
    chunk = degree - 2
    tr = trace if trace is not None else {}

    # --- vk, instances (KZG: QUERY_INSTANCE = false, the values go into the transcript)
    T.common_scalar(vk_repr)
    inst_cols = []
    for vals in instances:
        if len(vals) > u:
            raise ValueError("InstanceTooLarge")
        for v in vals:
            T.common_scalar(v)
        inst_cols.append([v % R for v in vals] + [0] * (n - len(vals)))
    # --- advice: blind the unusable rows, one (unused) Blind per column, commit
    adv = [list(c) for c in advice]
    for c in adv:
        for r in range(u, n):
            c[r] = rng.fr_random()
    for _ in adv:
        rng.fr_random()
    for c in adv:
        T.write_point(_commit(params.g_lagrange, c))
    theta = T.squeeze_challenge()
    tr["theta"] = theta

    def column(kind, idx):
        return adv[idx] if kind == 0 else fixed[idx] if kind == 1 else inst_cols[idx]

    # --- lookups: permuted input / table
    pa, ps = [], []
    for (li, lt) in cs["lookups"]:
        a_np, s_np = O.permute_expression_pair(_np(adv[li][:u]), _np(fixed[lt][:u]))
        a_, s_ = _ints(a_np), _ints(s_np)
        a_ += [rng.fr_random() for _ in range(bf + 1)]
        s_ += [rng.fr_random() for _ in range(bf + 1)]
        rng.fr_random()
        rng.fr_random()
        T.write_point(_commit(params.g_lagrange, a_))
        T.write_point(_commit(params.g_lagrange, s_))
        pa.append(a_)
        ps.append(s_)
    beta = T.squeeze_challenge()
    gamma = T.squeeze_challenge()
    tr["beta"], tr["gamma"] = beta, gamma
    # --- permutation grand products
    zs = []
    last_z = 1
    deltaomega_start = 1
    for s0 in range(0, len(perm), chunk):
        cols = perm[s0:s0 + chunk]
        mod = [1] * n
        for j, (kind, idx) in enumerate(cols):
            v, sg = column(kind, idx), sigma[s0 + j]
            for i in range(n):
                mod[i] = mod[i] * (beta * sg[i] + gamma + v[i]) % R
        mod = _ints(O.fr_batch_invert(_np(mod)))
        d = deltaomega_start
        for j, (kind, idx) in enumerate(cols):
            v = column(kind, idx)
            w = d
            for i in range(n):
                mod[i] = mod[i] * (w * beta + gamma + v[i]) % R
                w = w * omega % R
            d = d * DELTA % R
        deltaomega_start = d
        z = [last_z]
        for row in range(1, n):
            z.append(z[row - 1] * mod[row - 1] % R)
        for r in range(n - bf, n):
            z[r] = rng.fr_random()
        last_z = z[u]
        rng.fr_random()
        T.write_point(_commit(params.g_lagrange, z))
        zs.append(z)
    # --- lookup grand products
    zl = []
    for l, (li, lt) in enumerate(cs["lookups"]):
        prod = [(beta + pa[l][i]) * (gamma + ps[l][i]) % R for i in range(n)]
        prod = _ints(O.fr_batch_invert(_np(prod)))
        a, t = adv[li], fixed[lt]
        prod = [prod[i] * ((a[i] + beta) % R) % R * ((t[i] + gamma) % R) % R for i in range(n)]
        z = [1]
        for i in range(n - bf - 1):
            z.append(z[-1] * prod[i] % R)
        z += [rng.fr_random() for _ in range(bf)]
        rng.fr_random()
        T.write_point(_commit(params.g_lagrange, z))
        zl.append(z)
    # --- vanishing argument: random polynomial
    random_poly = [rng.fr_random() for _ in range(n)]
    rng.fr_random()
    T.write_point(_commit(params.g, random_poly))
    y = T.squeeze_challenge()
    tr["y"] = y

    # --- coefficient / extended forms, evaluate_h
    def l2c(col):
        return dom.lagrange_to_coeff(_np(col))

    adv_c = [l2c(c) for c in adv]
    inst_c = [l2c(c) for c in inst_cols]
    fixed_c = [l2c(c) for c in fixed]
    sigma_c = [l2c(c) for c in sigma]
    pa_c, ps_c = [l2c(c) for c in pa], [l2c(c) for c in ps]
    z_c, zl_c = [l2c(c) for c in zs], [l2c(c) for c in zl]
    ext = dom.coeff_to_extended
    adv_e = [ext(c) for c in adv_c]
    inst_e = [ext(c) for c in inst_c]
    fixed_e = [ext(c) for c in fixed_c]
    sigma_e = [ext(c) for c in sigma_c]
    l0 = ext(l2c([1] + [0] * (n - 1)))
    l_last = ext(l2c([1 if r == u else 0 for r in range(n)]))
    l_active = ext(l2c([1 if r < u else 0 for r in range(n)]))
    yn, bn, gn_ = _one(y), _one(beta), _one(gamma)
    h = np.zeros((ne, 4), dtype=np.uint64)
    if cs["gates"]:
        h = dom.quotient_gates(h, yn, np.stack([fixed_e[s] for (_, s) in cs["gates"]]), np.stack([adv_e[a] for (a, _) in cs["gates"]]))

    def col_e(kind, idx):
        return adv_e[idx] if kind == 0 else fixed_e[idx] if kind == 1 else inst_e[idx]

    if perm:
        h = dom.quotient_permutation(h, yn, bn, gn_, chunk, np.stack([col_e(kd, ix) for (kd, ix) in perm]), np.stack(sigma_e),
                                     np.stack([ext(c) for c in z_c]), l0, l_last, l_active, bf)
    for l, (li, lt) in enumerate(cs["lookups"]):
        h = dom.quotient_lookup(h, yn, bn, gn_, adv_e[li], fixed_e[lt], ext(pa_c[l]), ext(ps_c[l]), ext(zl_c[l]), l0, l_last, l_active)
    h_coeff = dom.extended_to_coeff(dom.divide_by_vanishing_poly(h))          # n * (degree - 1) coefficients
    pieces = [h_coeff[i * n:(i + 1) * n] for i in range(degree - 1)]
    for _ in pieces:
        rng.fr_random()
    for p_ in pieces:
        T.write_point(O.g1_affine_to_ints(O.best_multiexp_affine(p_, params.g)))
    x = T.squeeze_challenge()
    xn = pow(x, n, R)
    tr["x"] = x

    # --- evaluations
    def ev(coeff_np, pt):
        return _ints(O.fr_eval_poly(coeff_np, _one(pt)))[0]

    x_next, x_prev, x_last = _rotate(x, omega, 1), _rotate(x, omega, -1), _rotate(x, omega, -(bf + 1))
    adv_evals = [ev(adv_c[c], _rotate(x, omega, rot)) for (c, rot) in cs["advice_queries"]]
    fix_evals = [ev(fixed_c[c], _rotate(x, omega, rot)) for (c, rot) in cs["fixed_queries"]]
    for e in adv_evals + fix_evals:
        T.write_scalar(e)
    pieces_i = [_ints(p_) for p_ in pieces]
    h_poly = [0] * n
    for p_ in reversed(pieces_i):
        h_poly = [(a * xn + b) % R for a, b in zip(h_poly, p_)]
    random_eval = ev(_np(random_poly), x)
    T.write_scalar(random_eval)
    sigma_evals = [ev(c, x) for c in sigma_c]
    for e in sigma_evals:
        T.write_scalar(e)
    z_evals = []
    for s, c in enumerate(z_c):
        e0, e1 = ev(c, x), ev(c, x_next)
        T.write_scalar(e0)
        T.write_scalar(e1)
        e2 = None
        if s + 1 < len(z_c):
            e2 = ev(c, x_last)
            T.write_scalar(e2)
        z_evals.append((e0, e1, e2))
    lk_evals = []
    for l in range(L):
        es = (ev(zl_c[l], x), ev(zl_c[l], x_next), ev(pa_c[l], x), ev(pa_c[l], x_prev), ev(ps_c[l], x))
        for e in es:
            T.write_scalar(e)
        lk_evals.append(es)
    # --- multi-open argument (SHPLONK).  Query keys identify polynomials; coefficient vectors are looked up by key.
    polys = {}
    queries = []

    def q(key, coeff, pt, e):
        polys[key] = coeff
        queries.append((key, pt, e))

    for (c, rot), e in zip(cs["advice_queries"], adv_evals):
        q(("adv", c), adv_c[c], _rotate(x, omega, rot), e)
    for s, c in enumerate(z_c):
        q(("z", s), c, x, z_evals[s][0])
        q(("z", s), c, x_next, z_evals[s][1])
    for s in reversed(range(len(z_c) - 1)):
        q(("z", s), z_c[s], x_last, z_evals[s][2])
    for l in range(L):
        e = lk_evals[l]
        q(("zl", l), zl_c[l], x, e[0])
        q(("pa", l), pa_c[l], x, e[2])
        q(("ps", l), ps_c[l], x, e[4])
        q(("pa", l), pa_c[l], x_prev, e[3])
        q(("zl", l), zl_c[l], x_next, e[1])
    for (c, rot), e in zip(cs["fixed_queries"], fix_evals):
        q(("fix", c), fixed_c[c], _rotate(x, omega, rot), e)
    for c, e in enumerate(sigma_evals):
        q(("sig", c), sigma_c[c], x, e)
    h_np = _np(h_poly)
    q(("h",), h_np, x, ev(h_np, x))
    q(("rnd",), _np(random_poly), x, random_eval)

    ych = T.squeeze_challenge()
    vch = T.squeeze_challenge()
    rsets, super_pts = _intermediate_sets(queries)
    S_list, r_list = [], []
    hx = [0] * n
    vp = 1
    for pts, coms in rsets:
        S = [0] * n
        Rcomb = [0] * len(pts)
        yp = 1
        rs = []
        for key, evs in coms:
            coeffs = _ints(polys[key])
            S = [(a + yp * b) % R for a, b in zip(S, coeffs)]
            rpoly = _lagrange_interpolate(pts, evs)
            rs.append(rpoly)
            Rcomb = [(a + yp * b) % R for a, b in zip(Rcomb, rpoly)]
            yp = yp * ych % R
        S_list.append(S)
        r_list.append(rs)
        N = list(S)
        for t_, c in enumerate(Rcomb):
            N[t_] = (N[t_] - c) % R
        Q = _np(N)
        for p_ in pts:
            Q = O.fr_kate_division(Q, _one(p_))
        Q = _ints(Q) + [0] * len(pts)
        hx = [(a + vp * b) % R for a, b in zip(hx, Q)]
        vp = vp * vch % R
    T.write_point(_commit(params.g, hx))
    uch = T.squeeze_challenge()
    Lx = [0] * n
    vp = 1
    z0 = None
    for i, (pts, coms) in enumerate(rsets):
        zi = 1
        for p_ in super_pts:
            if p_ not in pts:
                zi = zi * (uch - p_) % R
        if i == 0:
            z0 = zi
        ri, yp = 0, 1
        for rpoly in r_list[i]:
            ri = (ri + yp * P.eval_poly(rpoly, uch)) % R
            yp = yp * ych % R
        c = vp * zi % R
        Lx = [(a + c * b) % R for a, b in zip(Lx, S_list[i])]
        Lx[0] = (Lx[0] - c * ri) % R
        vp = vp * vch % R
    zt = 1
    for p_ in super_pts:
        zt = zt * (uch - p_) % R
    Lx = [(a - zt * b) % R for a, b in zip(Lx, hx)]
    W = _ints(O.fr_kate_division(_np(Lx), _one(uch)))
    z0inv = pow(z0, -1, R)
    W = [w * z0inv % R for w in W]
    T.write_point(_commit(params.g, W))
    return T.finalize()


# =================================================================================================== verify_proof
def verify_proof(params, cs, vk, vk_repr, instances, proof):
    """plonk/verifier.rs verify_proof with the SHPLONK verifier; the final pairing equation is checked in G1 through
    the setup secret `params.s`.  Returns True / False (malformed encodings count as False)."""
    try:
        return _verify(params, cs, vk, vk_repr, instances, proof)
    except ValueError:
        return False


def _verify(params, cs, vk, vk_repr, instances, proof):
    if params.s is None:
        raise RuntimeError("verify_proof needs the setup secret (no pairing here)")
    k, n, bf, degree = cs["k"], 1 << cs["k"], cs["blinding_factors"], cs["degree"]
    dom = P.Domain(degree, k)
    omega = dom.omega
    A, L = cs["n_advice"], len(cs["lookups"])
    perm = cs["permutation"]
    chunk = degree - 2
    NS = (len(perm) + chunk - 1) // chunk if perm else 0
    T = PoseidonTranscript(proof)
    T.common_scalar(vk_repr)
    for vals in instances:
        for v in vals:
            T.common_scalar(v)
    adv_com = [T.read_point() for _ in range(A)]
    theta = T.squeeze_challenge()
    lk_perm = [(T.read_point(), T.read_point()) for _ in range(L)]
    beta = T.squeeze_challenge()
    gamma = T.squeeze_challenge()
    z_com = [T.read_point() for _ in range(NS)]
    zl_com = [T.read_point() for _ in range(L)]
    rnd_com = T.read_point()
    y = T.squeeze_challenge()
    h_com = [T.read_point() for _ in range(degree - 1)]
    x = T.squeeze_challenge()
    xn = pow(x, n, R)
    # instance evaluations by Lagrange interpolation (KZG does not open the instance columns)
    iq = cs.get("instance_queries", [])
    inst_evals = []
    if iq:
        min_rot = min(0, min(r for _, r in iq))
        max_rot = max(0, max(r for _, r in iq))
        max_len = max([len(v) for v in instances] + [0])
        lo, hi = -max_rot, max_len + abs(min_rot)
        l_i_s = [_l_i(x, xn, omega, n, i) for i in range(lo, hi)]
        for (c, rot) in iq:
            off = max_rot - rot
            inst_evals.append(sum(v * l for v, l in zip(instances[c], l_i_s[off:off + len(instances[c])])) % R)
    adv_evals = [T.read_scalar() for _ in cs["advice_queries"]]
    fix_evals = [T.read_scalar() for _ in cs["fixed_queries"]]
    random_eval = T.read_scalar()
    sigma_evals = [T.read_scalar() for _ in perm]
    z_evals = []
    for s in range(NS):
        e0, e1 = T.read_scalar(), T.read_scalar()
        e2 = T.read_scalar() if s + 1 < NS else None
        z_evals.append((e0, e1, e2))
    lk_evals = [tuple(T.read_scalar() for _ in range(5)) for _ in range(L)]   # z, z_next, a', a'_inv, s'
    # --- expected h(x)
    l_evals = [_l_i(x, xn, omega, n, i) for i in range(-(bf + 1), 1)]
    l_last, l_blind, l_0 = l_evals[0], sum(l_evals[1:1 + bf]) % R, l_evals[1 + bf]
    aq = {q_: i for i, q_ in enumerate(cs["advice_queries"])}
    fq = {q_: i for i, q_ in enumerate(cs["fixed_queries"])}
    iqm = {q_: i for i, q_ in enumerate(iq)}

    def ae(c, rot=0):
        return adv_evals[aq[(c, rot)]]

    def fe_(c, rot=0):
        return fix_evals[fq[(c, rot)]]

    def any_eval(kind, idx):
        return ae(idx) if kind == 0 else fe_(idx) if kind == 1 else inst_evals[iqm[(idx, 0)]]

    exprs = []
    for (a, s) in cs["gates"]:
        exprs.append(fe_(s) * (ae(a, 0) + ae(a, 1) * ae(a, 2) - ae(a, 3)) % R)
    active = (1 - (l_last + l_blind)) % R
    if NS:
        exprs.append(l_0 * (1 - z_evals[0][0]) % R)
        zl_ = z_evals[-1][0]
        exprs.append((zl_ * zl_ - zl_) * l_last % R)
        for s in range(1, NS):
            exprs.append((z_evals[s][0] - z_evals[s - 1][2]) * l_0 % R)
        for s in range(NS):
            cols = perm[s * chunk:(s + 1) * chunk]
            left = z_evals[s][1]
            for j, (kind, idx) in enumerate(cols):
                left = left * (any_eval(kind, idx) + beta * sigma_evals[s * chunk + j] + gamma) % R
            right = z_evals[s][0]
            cur = beta * x % R * pow(DELTA, s * chunk, R) % R
            for (kind, idx) in cols:
                right = right * (any_eval(kind, idx) + cur + gamma) % R
                cur = cur * DELTA % R
            exprs.append((left - right) * active % R)
    for l, (li, lt) in enumerate(cs["lookups"]):
        ze, zn, ap, ai, sp = lk_evals[l]
        exprs.append(l_0 * (1 - ze) % R)
        exprs.append(l_last * (ze * ze - ze) % R)
        left = zn * (ap + beta) % R * (sp + gamma) % R
        right = ze * (ae(li) + beta) % R * (fe_(lt) + gamma) % R      # single expressions: theta does not enter
        exprs.append((left - right) * active % R)
        exprs.append(l_0 * (ap - sp) % R)
        exprs.append((ap - sp) * (ap - ai) % R * active % R)
    expected_h = 0
    for e in exprs:
        expected_h = (expected_h * y + e) % R
    expected_h = expected_h * pow(xn - 1, -1, R) % R
    # --- queries (commitment given as a list of (scalar, point) terms so that h's MSM form fits)
    x_next, x_prev, x_last = _rotate(x, omega, 1), _rotate(x, omega, -1), _rotate(x, omega, -(bf + 1))
    coms = {}
    queries = []

    def q(key, terms, pt, e):
        coms[key] = terms
        queries.append((key, pt, e))

    for i, (c, rot) in enumerate(cs["advice_queries"]):
        q(("adv", c), [(1, adv_com[c])], _rotate(x, omega, rot), adv_evals[i])
    for s in range(NS):
        q(("z", s), [(1, z_com[s])], x, z_evals[s][0])
        q(("z", s), [(1, z_com[s])], x_next, z_evals[s][1])
    for s in reversed(range(NS - 1)):
        q(("z", s), [(1, z_com[s])], x_last, z_evals[s][2])
    for l in range(L):
        ze, zn, ap, ai, sp = lk_evals[l]
        q(("zl", l), [(1, zl_com[l])], x, ze)
        q(("pa", l), [(1, lk_perm[l][0])], x, ap)
        q(("ps", l), [(1, lk_perm[l][1])], x, sp)
        q(("pa", l), [(1, lk_perm[l][0])], x_prev, ai)
        q(("zl", l), [(1, zl_com[l])], x_next, zn)
    for i, (c, rot) in enumerate(cs["fixed_queries"]):
        q(("fix", c), [(1, vk["fixed"][c])], _rotate(x, omega, rot), fix_evals[i])
    for c in range(len(perm)):
        q(("sig", c), [(1, vk["sigma"][c])], x, sigma_evals[c])
    q(("h",), [(pow(xn, i, R), h_com[i]) for i in range(degree - 1)], x, expected_h)
    q(("rnd",), [(1, rnd_com)], x, random_eval)
    # --- SHPLONK verifier
    ych = T.squeeze_challenge()
    vch = T.squeeze_challenge()
    h1 = T.read_point()
    uch = T.squeeze_challenge()
    h2 = T.read_point()
    if T.pos != len(proof):
        return False
    rsets, super_pts = _intermediate_sets(queries)
    terms = []                 # the outer MSM as (scalar, point) pairs
    r_outer = 0
    z0 = z0_diff_inv = None
    vp = 1
    for i, (pts, cms) in enumerate(rsets):
        zdi = 1
        for p_ in super_pts:
            if p_ not in pts:
                zdi = zdi * (uch - p_) % R
        if i == 0:
            z0 = 1
            for p_ in pts:
                z0 = z0 * (uch - p_) % R
            z0_diff_inv = pow(zdi, -1, R)
            zdi = 1
        else:
            zdi = zdi * z0_diff_inv % R
        yp = 1
        r_inner = 0
        for key, evs in cms:
            r_inner = (r_inner + yp * P.eval_poly(_lagrange_interpolate(pts, evs), uch)) % R
            for (sc, pt) in coms[key]:
                terms.append((sc * yp % R * vp % R * zdi % R, pt))
            yp = yp * ych % R
        r_outer = (r_outer + vp * r_inner % R * zdi) % R
        vp = vp * vch % R
    terms.append(((-r_outer) % R, P.G1_GEN))
    terms.append(((-z0) % R, h1))
    terms.append((uch, h2))
    rhs = None
    for sc, pt in terms:
        rhs = P.g1_add(rhs, P.g1_mul(pt, sc))
    # e(h2, [s]_2) == e(rhs, [1]_2)   <=>   s * h2 == rhs
    return P.g1_mul(h2, params.s) == rhs


def _l_i(x, xn, omega, n, i):
    """EvaluationDomain::l_i_range entry: l_i(x) = w^i (x^n - 1) / (n (x - w^i))"""
    wi = pow(omega, i, R)
    return wi * (xn - 1) % R * pow(n * (x - wi) % R, -1, R) % R
