"""TEST INFRASTRUCTURE (oracle side, never imported by the product): the checks of halo2's `MockProver::run(..)
.assert_satisfied()` as the reference's `mock` command runs them (/root/reference/src/scaffold/mod.rs:263-266), restated
for the constraint-system shape halo2-base builds, plus an independent cell count of the reference's chips.

* `mock_prover(cs, fixed, sigma, advice, instances)`: every enabled vertical gate q * (a + b c - d) holds, every
  lookup-advice value lies in the table column, the sigma columns are a permutation of the identity and every cell equals
  its image (copy constraints, constants, public inputs).  Plain Python integers; columns are lists of canonical ints.
* `CellCount`: the number of advice cells / lookup cells each chip call adds, derived from the cell layouts of the
  halo2-base primitives [UPSTREAM halo2-lib v0.3.0, recalled] and the call structure of
  /root/reference/src/gadget/{fixed_point,distance,vectordb}.rs -- no values, only counts, written independently of
  the product's builder (csrc/zk_builder.hpp, zk_chips.hpp) so that the two can be compared.

parity unpinned: nothing from the reference's crates can run here; what pins these is the algebra
(gates hold or they do not) and, for the chips' results, the f64 computations of the reference's own tests.
"""
from . import pyref as P

R = P.R


def _delta():
    return pow(7, 1 << 28, R)


def mock_prover(cs, fixed, sigma, advice, instances, max_failures=8):
    """-> list of failure descriptions (empty = satisfied)"""
    k = cs["k"]
    n = 1 << k
    bf = cs["blinding_factors"]
    usable = n - (bf + 1)
    fails = []

    def fail(msg):
        fails.append(msg)
        return len(fails) >= max_failures

    # ---- gates
    for (a_col, q_col) in cs["gates"]:
        a, q = advice[a_col], fixed[q_col]
        for row in range(n):
            if q[row]:
                if row + 3 >= usable:
                    if fail(f"gate on column {a_col} row {row} reaches into the unusable rows"):
                        return fails
                    continue
                if (q[row] * (a[row] + a[row + 1] * a[row + 2] - a[row + 3])) % R:
                    if fail(f"gate not satisfied: column {a_col} row {row}"):
                        return fails
    # ---- lookups
    for (in_col, t_col) in cs["lookups"]:
        table = set(fixed[t_col][:usable])
        col = advice[in_col]
        for row in range(usable):
            if col[row] not in table:
                if fail(f"lookup: column {in_col} row {row} holds {col[row]}, not in the table"):
                    return fails
    # ---- permutation
    perm = cs["permutation"]
    omega = P.omega_for(k)
    delta = _delta()
    col_of = {pow(pow(delta, c, R), n, R): c for c in range(len(perm))}
    row_of = {}
    w = 1
    for r_ in range(n):
        row_of[w] = r_
        w = w * omega % R
    dinv = [pow(pow(delta, c, R), R - 2, R) for c in range(len(perm))]

    def column(kind, idx):
        if kind == 0:
            return advice[idx]
        if kind == 1:
            return fixed[idx]
        inst = instances[idx]
        return list(inst) + [0] * (n - len(inst))

    cols = [column(kind, idx) for kind, idx in perm]
    seen = [bytearray(n) for _ in perm]
    for c, sg in enumerate(sigma):
        for row in range(n):
            v = sg[row]
            c2 = col_of.get(pow(v, n, R))
            r2 = row_of.get(v * dinv[c2] % R) if c2 is not None else None
            if r2 is None:
                if fail(f"sigma column {c} row {row} is not delta^i omega^j"):
                    return fails
                continue
            if seen[c2][r2]:
                if fail(f"sigma is not a permutation: ({c2}, {r2}) hit twice"):
                    return fails
            seen[c2][r2] = 1
            if (c2, r2) != (c, row):
                if row >= usable or r2 >= usable:
                    if fail(f"copy constraint touches an unusable row: ({c}, {row}) -> ({c2}, {r2})"):
                        return fails
                if cols[c][row] != cols[c2][r2]:
                    if fail(f"copy constraint violated: ({c}, {row}) = {cols[c][row]} vs ({c2}, {r2}) = {cols[c2][r2]}"):
                        return fails
    return fails


def copy_classes(cs, sigma):
    """number of cells that take part in some copy constraint (cells whose sigma is not the identity)"""
    k = cs["k"]
    n = 1 << k
    omega = P.omega_for(k)
    delta = _delta()
    moved = 0
    for c, sg in enumerate(sigma):
        w = pow(delta, c, R)
        for row in range(n):
            if sg[row] != w:
                moved += 1
            w = w * omega % R
    return moved


# ------------------------------------------------------------------------------------------------ cell counts
class CellCount:
    """(advice cells, lookup cells) added by each call; P = PRECISION_BITS, lb = LOOKUP_BITS"""

    def __init__(self, precision_bits=48, lookup_bits=12):
        self.P, self.lb = precision_bits, lookup_bits
        self.advice = 0
        self.lookup = 0

    # -- halo2-base GateChip: every basic operation is one 4-cell gate region
    def gate(self, times=1):
        self.advice += 4 * times

    add = sub = neg = mul = mul_add = assert_bit = gate

    def load(self, times=1):                 # load_witness / load_constant
        self.advice += times

    def select(self):                        # two gates
        self.advice += 8

    or_ = is_zero = select

    def is_equal(self):
        self.sub()
        self.is_zero()

    def sum(self, m):
        self.advice += 1 if m == 1 else 1 + 3 * (m - 1)

    def inner_product(self, m, starts_with_one):
        self.advice += 1 + 3 * (m - 1) if starts_with_one else 1 + 3 * m

    def num_to_bits(self, bits):
        self.inner_product(bits, True)       # pow_of_two[0] = 1
        self.advice += 4 * bits              # assert_bit each

    def select_from_idx(self, m):
        self.advice += m * (7 + 4)           # idx_to_indicator: one 7-cell region + assert_bit per index
        self.advice += 1 + 3 * m             # select_by_indicator

    def select_by_indicator(self, m):
        self.advice += 1 + 3 * m

    # -- RangeChip
    def range_check(self, bits):
        k = -(-bits // self.lb)
        rem = bits % self.lb
        if k > 1:
            self.inner_product(k, True)      # limb_bases[0] = 1
        self.lookup += k
        if rem == 1:
            self.assert_bit()
        elif rem > 1:
            self.mul()
            self.lookup += 1

    def check_less_than(self, bits):
        self.advice += 7
        self.range_check(bits)

    def check_big_less_than_safe(self, bound_bits):
        rb = -(-bound_bits // self.lb) * self.lb
        self.range_check(rb)
        self.check_less_than(rb)

    def is_less_than(self, bits):
        padded = -(-bits // self.lb) * self.lb
        self.advice += 7
        self.range_check(padded + self.lb)
        self.is_zero()

    def div_mod(self, divisor_bits_minus_1, a_bits):
        # divisor = 2^e: quotient bound 2^(a_bits - e) + 1 has a_bits - e + 1 bits, the divisor itself e + 1 bits
        e = divisor_bits_minus_1
        self.advice += 4
        self.check_big_less_than_safe(a_bits - e + 1)
        self.check_big_less_than_safe(e + 1)

    def div_mod_var(self, a_bits, b_bits):
        self.advice += 4
        self.range_check(a_bits)
        self.range_check(b_bits)
        self.check_less_than(b_bits)

    # -- FixedPointChip (fixed_point.rs)
    def is_neg(self):                        # :523-539
        self.div_mod(2 * self.P + 1, 254)
        self.is_zero()
        self.sub()                           # not

    def qabs(self):                          # :511-521
        self.neg()
        self.is_neg()
        self.select()

    def cond_neg(self):                      # :541-556
        self.neg()
        self.select()

    def signed_div_scale(self):              # :974-1016
        self.advice += 4
        self.check_big_less_than_safe(self.P + 1)
        self.qabs()
        self.check_big_less_than_safe(3 * self.P + 1)

    def qmul(self):                          # :588-604
        self.mul()
        self.signed_div_scale()

    def bit_xor(self):                       # :797-815
        self.add(2)
        self.assert_bit(2)
        self.add(2)
        self.is_equal()

    def qdiv(self):                          # :631-656
        self.is_neg()
        self.is_neg()
        self.qabs()
        self.qabs()
        self.mul()
        self.div_mod_var(4 * self.P, 2 * self.P)
        self.bit_xor()
        self.cond_neg()

    def qmod(self):                          # :606-629
        self.is_neg()
        self.is_neg()
        self.qabs()
        self.div_mod_var(4 * self.P, 2 * self.P)
        self.sub()
        self.select()

    def polynomial(self, n_coef):            # :658-686
        self.add()
        for i in range(n_coef):
            self.add()
            if i < n_coef - 1:
                self.qmul()

    def check_power_of_two(self):            # :688-708
        bits = 2 * self.P
        self.num_to_bits(bits)
        self.sum(bits)
        self.sub()
        self.is_zero()
        self.select_from_idx(bits)
        self.sub()
        self.is_zero()

    def qexp2(self):                         # :710-734
        self.qabs()
        self.div_mod(self.P, 2 * self.P)
        self.select_from_idx(254)
        self.polynomial(13)
        self.mul()
        self.qdiv()
        self.is_neg()
        self.select()

    def qlog2(self):                         # :736-795
        self.add()
        self.is_neg()
        self.is_zero()
        self.or_()
        self.add(2)
        self.check_power_of_two()
        self.mul()
        self.add()
        self.check_power_of_two()
        self.is_less_than(2 * self.P)
        self.is_less_than(2 * self.P)
        self.is_equal()
        self.or_()
        self.mul()                           # and
        self.sub()
        self.is_neg()
        self.qabs()
        self.add()
        self.check_power_of_two()
        self.mul()
        self.div_mod_var(2 * self.P, self.P + 1)
        self.select()
        self.polynomial(15)
        self.neg()
        self.mul()
        self.add()

    def qexp(self):                          # :876-886
        self.load()
        self.qdiv()
        self.qexp2()

    def qlog(self):                          # :954-964
        self.load()
        self.qlog2()
        self.qdiv()

    def qsqrt(self):                         # :966-972, :441-456
        self.load()
        self.qlog()
        self.qmul()
        self.qexp()

    def fp_inner_product(self, m):           # :854-874
        self.add()
        for _ in range(m):
            self.qmul()
            self.add()

    def qmin(self):                          # :936-952
        self.sub()
        self.is_neg()
        self.select()

    # -- DistanceChip (distance.rs)
    def euclidean(self, dim):                # :97-119
        self.sub(dim)
        self.fp_inner_product(dim)
        self.qsqrt()

    def cosine(self, dim):                   # :121-144
        for _ in range(3):
            self.fp_inner_product(dim)
        self.qsqrt()
        self.qsqrt()
        self.qmul()
        self.qdiv()
        self.load()
        self.sub()

    def hamming(self, dim):                  # :146-175
        for _ in range(dim):
            self.is_equal()
        self.sum(dim)
        self.load(2)
        self.qdiv()
        self.load()
        self.sub()

    def manhattan(self, dim):                # :177-195
        self.sub(dim)
        for _ in range(dim):
            self.qabs()
        self.sum(dim)

    # -- Poseidon chip, T = 3, RATE = 2
    def poseidon_permutation(self, n_inputs, r_f, r_p):
        T = 3
        self.add()                                           # state[0] + pre-constant
        for _ in range(n_inputs):
            self.sum(3)
        self.add(T - 1 - n_inputs)
        full = (12 * T) + T * (1 + 3 * T)                    # three x^5 (3 gates each) + three inner products of length T
        partial = 12 + (1 + 3 * T) + 4 * (T - 1)
        self.advice += r_f * full + r_p * partial

    def poseidon_hash(self, m, r_f, r_p):
        for i in range(0, m, 2):
            self.poseidon_permutation(min(2, m - i), r_f, r_p)
        if m % 2 == 0:
            self.poseidon_permutation(0, r_f, r_p)

    # -- VectorDBChip (vectordb.rs)
    def nearest_vector(self, n_vec, dim, distance):          # :122-163
        for _ in range(n_vec):
            distance(dim)
        for _ in range(n_vec - 1):
            self.qmin()
        for _ in range(n_vec):
            self.is_equal()
        for _ in range(dim):
            self.select_by_indicator(n_vec)

    def merkle_commitment(self, n_vec, dim, r_f, r_p):      # :165-223
        for _ in range(n_vec):
            self.poseidon_hash(dim, r_f, r_p)
        leaves = 1
        while leaves < n_vec:
            leaves <<= 1
        if leaves > n_vec:
            self.load()                                       # load_zero (cached afterwards)
        while leaves > 1:
            for _ in range(leaves // 2):
                self.poseidon_hash(2, r_f, r_p)
            leaves //= 2

    def kmeans(self, n_vec, dim, K, I, distance):            # :225-362
        self.load(2)                                          # quantised one, zero
        for _ in range(I):
            for _ in range(n_vec):
                for _ in range(K):
                    distance(dim)
                for _ in range(K - 1):
                    self.qmin()
                for _ in range(K):
                    self.is_equal()
                    self.select()
            self.add((n_vec - 1) * K)
            for _ in range(K):
                for _ in range(n_vec):
                    self.is_zero()
                    for _ in range(dim):
                        self.select()
                self.add((n_vec - 1) * dim)
                for _ in range(dim):
                    self.qdiv()


def example_cell_counts(name, inp, lookup_bits, precision_bits=48):
    """cells of /root/reference/examples/{distances,query,kmeans}.rs on the given input -> (advice, lookup)"""
    c = CellCount(precision_bits, lookup_bits)
    if name == "distances":
        dim = len(inp["a"])
        c.load(2 * dim)
        c.euclidean(dim)
        c.manhattan(dim)
        c.cosine(dim)
        c.hamming(dim)
    elif name == "query":
        dim, n_vec = len(inp["query"]), len(inp["database"])
        c.load(3)                                             # PoseidonChip::new: the initial state
        c.load(dim * (n_vec + 1))
        c.nearest_vector(n_vec, dim, c.cosine)
        c.merkle_commitment(n_vec, dim, 8, 57)
    elif name == "kmeans":
        dim, n_vec = len(inp["vectors"][0]), len(inp["vectors"])
        c.load(dim * n_vec)
        c.kmeans(n_vec, dim, 4, 10, c.cosine)
    else:
        raise ValueError(name)
    return c.advice, c.lookup
