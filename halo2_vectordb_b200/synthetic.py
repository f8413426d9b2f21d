"""Synthetic inputs for benchmarks and full-size tests (no oracle involved).

* uniform_scalars: (cols, n, 4) uint64 whose 256-bit value is uniform below 2^252 < r -- every such value is a
  valid Montgomery residue, i.e. the Montgomery form of a uniformly distributed field element.
* witness_like: the skew of real halo2-base / FixedPointChip advice columns
  (/root/reference/src/gadget/fixed_point.rs:68-119: bits, small limbs < 2^LOOKUP_BITS, and "negative"
  fixed-point values r - small that are full width): 60% {0,1}, 30% < 2^lookup_bits, 5% full width, 5% r - small.
  Canonical values are converted to Montgomery form on the device (h2v_selftest_field: x * R^2 * R^-1).
"""
import numpy as np

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R2 = 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7


def _limbs(x):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def uniform_scalars(cols, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 64, (cols, n, 4), dtype=np.uint64)
    a[..., 3] &= np.uint64((1 << 60) - 1)
    return a


def witness_like(cols, n, lookup_bits, seed):
    from . import selftest_field

    rng = np.random.default_rng(seed)
    tot = cols * n
    canon = np.zeros((tot, 4), dtype=np.uint64)
    sel = rng.integers(0, 100, tot)
    canon[:, 0] = np.where(sel < 60, rng.integers(0, 2, tot), rng.integers(0, 1 << lookup_bits, tot)).astype(np.uint64)
    full = sel >= 90
    canon[full] = rng.integers(0, 1 << 62, (int(full.sum()), 4)).astype(np.uint64)     # < 2^254 < r
    neg = sel >= 95
    small = rng.integers(1, 1 << 40, int(neg.sum())).astype(np.uint64)
    negv = np.tile(np.array(_limbs(R_MOD), dtype=np.uint64), (int(neg.sum()), 1))
    negv[:, 0] = negv[:, 0] - small                                                     # low limb of r > 2^40: no borrow
    canon[neg] = negv
    r2 = np.tile(np.array(_limbs(R2), dtype=np.uint64), (tot, 1))
    return selftest_field(0, 0, canon, r2).reshape(cols, n, 4)


def to_mont(canon):
    """Canonical Fr values ((n, 4) uint64 limbs) -> Montgomery form, on the device (x * R^2 * R^-1)."""
    from . import selftest_field

    canon = np.ascontiguousarray(canon, dtype=np.uint64).reshape(-1, 4)
    r2 = np.tile(np.array(_limbs(R2), dtype=np.uint64), (canon.shape[0], 1))
    return selftest_field(0, 0, canon, r2)


# ----------------------------------------------------------------------------------------------- circuits
R_ONE = 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB      # R mod r: Montgomery form of 1
GEN, S_2ADIC = 7, 28


def _mul(a, b):
    from . import selftest_field
    return selftest_field(0, 0, a, b)


def _add(a, b):
    from . import selftest_field
    return selftest_field(0, 1, a, b)


def _witness_canonical(rng, count, lookup_bits):
    """`count` canonical values with the skew of FixedPointChip cells (see witness_like) as (count, 4) limbs"""
    canon = np.zeros((count, 4), dtype=np.uint64)
    sel = rng.integers(0, 100, count)
    canon[:, 0] = np.where(sel < 60, rng.integers(0, 2, count), rng.integers(0, 1 << lookup_bits, count)).astype(np.uint64)
    full = sel >= 90
    canon[full] = rng.integers(0, 1 << 62, (int(full.sum()), 4)).astype(np.uint64)
    neg = sel >= 95
    small = rng.integers(1, 1 << 40, int(neg.sum())).astype(np.uint64)
    negv = np.tile(np.array(_limbs(R_MOD), dtype=np.uint64), (int(neg.sum()), 1))
    negv[:, 0] = negv[:, 0] - small
    canon[neg] = negv
    return canon


def synthetic_circuit(k, n_gate_cols, n_lookup_cols, lookup_bits, seed=0, n_public=4, blinding_factors=5, degree=4):
    """A satisfied circuit with halo2-base's constraint-system shape at any size, built with numpy and the library's own
    device field arithmetic (bench / large tests; no oracle involved): `n_gate_cols` basic-gate advice columns
    (q * (a + b c - d) on rows 4g .. 4g + 3, witness-shaped a, b, c), `n_lookup_cols` lookup-advice columns checked against
    one fixed table column of 2^lookup_bits values, a constants column and one instance column; copy constraints pair up the
    b cells of every gate column and tie the public inputs to advice cells.  Returns Montgomery-form numpy columns:
    dict(cs, fixed, sigma, advice, instances, vk_repr) as halo2_vectordb_b200.ProvingKey / create_proof take them."""
    rng = np.random.default_rng(seed)
    n = 1 << k
    bf = blinding_factors
    u = n - (bf + 1)
    G, Lc = n_gate_cols, n_lookup_cols
    A, F, TABLE, CONST = G + Lc, G + 2, G, G + 1
    tsize = 1 << lookup_bits
    assert tsize <= u
    perm = [(1, CONST)] + [(0, c) for c in range(A)] + [(2, 0)]
    cs = dict(k=k, degree=degree, blinding_factors=bf, n_advice=A, n_fixed=F, n_instance=1,
              gates=[(c, c) for c in range(G)], lookups=[(G + l, TABLE) for l in range(Lc)], permutation=perm,
              advice_queries=[q for c in range(G) for q in ((c, 0), (c, 1), (c, 2), (c, 3))] + [(G + l, 0) for l in range(Lc)],
              fixed_queries=[(CONST, 0), (TABLE, 0)] + [(c, 0) for c in range(G)], instance_queries=[(0, 0)])
    n_g = len(range(0, u - 3, 4))
    one = np.array(_limbs(R_ONE), dtype=np.uint64)
    # ---- gate columns
    advice = []
    pairs = []                       # per gate column: (rows of first cells, rows of their partners)
    chunk = max(1, (1 << 22) // n_g)   # columns per device call
    for c0 in range(0, G, chunk):
        cc = min(chunk, G - c0)
        abc = to_mont(_witness_canonical(rng, 3 * cc * n_g, lookup_bits)).reshape(3, cc, n_g, 4)
        a, b, c = abc[0], abc[1].copy(), abc[2]
        for j in range(cc):
            p = rng.permutation(n_g)
            half = n_g // 2
            first, second = p[:half], p[half:2 * half]
            b[j, second] = b[j, first]
            pairs.append((4 * first + 1, 4 * second + 1))
        d = _add(a.reshape(-1, 4), _mul(b.reshape(-1, 4), c.reshape(-1, 4))).reshape(cc, n_g, 4)
        for j in range(cc):
            col = np.zeros((n, 4), dtype=np.uint64)
            col[0:4 * n_g:4], col[1:4 * n_g:4], col[2:4 * n_g:4], col[3:4 * n_g:4] = a[j], b[j], c[j], d[j]
            advice.append(col)
    # ---- lookup-advice columns: values below 2^lookup_bits on the usable rows
    for _ in range(Lc):
        canon = np.zeros((n, 4), dtype=np.uint64)
        canon[:u, 0] = rng.integers(0, tsize, u).astype(np.uint64)
        advice.append(to_mont(canon))
    # ---- fixed columns: selectors, table, constants
    fixed = []
    for _ in range(G):
        q = np.zeros((n, 4), dtype=np.uint64)
        q[0:4 * n_g:4] = one
        fixed.append(q)
    tab = np.zeros((n, 4), dtype=np.uint64)
    tab[:tsize, 0] = np.arange(tsize, dtype=np.uint64)
    fixed.append(to_mont(tab))
    fixed.append(np.zeros((n, 4), dtype=np.uint64))
    # ---- public inputs: the a cells of the first gates of column 0
    n_public = min(n_public, n_g)
    instances = [advice[0][0:4 * n_public:4].copy()]
    # ---- sigma polynomials: identity delta^col * omega^row, then the copy cycles
    r, root = R_MOD, pow(GEN, (R_MOD - 1) >> S_2ADIC, R_MOD)
    omega = pow(root, 1 << (S_2ADIC - k), r)
    delta = pow(GEN, 1 << S_2ADIC, r)
    wp = np.zeros((n, 4), dtype=np.uint64)
    cur = 1
    for i in range(n):
        wp[i] = _limbs(cur)
        cur = cur * omega % r
    wp = to_mont(wp)

    def ident(pi):
        d_ = to_mont(np.array([_limbs(pow(delta, pi, r))], dtype=np.uint64))
        return _mul(wp, np.tile(d_, (n, 1)))

    sigma = [ident(pi) for pi in range(len(perm))]
    for j in range(G):
        pi = 1 + j
        r1, r2 = pairs[j]
        idc = sigma[pi].copy()
        sigma[pi][r1], sigma[pi][r2] = idc[r2], idc[r1]
    p_inst, p_adv0 = len(perm) - 1, 1
    rows = 4 * np.arange(n_public)
    id_adv0 = ident(p_adv0)                      # the a cells are untouched by the b swaps, but keep the identity values explicit
    sigma[p_inst][:n_public], sigma[p_adv0][rows] = id_adv0[rows], ident(p_inst)[:n_public]
    vk_repr = to_mont(np.array([_limbs(int(rng.integers(1, 1 << 62)) ** 3 % r)], dtype=np.uint64))[0]
    return dict(cs=cs, fixed=fixed, sigma=sigma, advice=advice, instances=instances, vk_repr=vk_repr)
