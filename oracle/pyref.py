"""Pure-Python big-integer model of the BN254 hot path (TEST INFRASTRUCTURE ONLY).

This file is the *mathematical* ground truth used to pin the C oracle
(`oracle/bn254_oracle.c`): Python ints, no Montgomery form, no windows, no
butterflies -- a naive O(n^2) DFT and double-and-add scalar multiplication.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import anything under `oracle/`; the product path never does.

PARITY UNPINNED: the reference (/root/reference) contains no prover arithmetic
and no golden vectors for this boundary; the
arithmetic lives in un-vendored git dependencies (halo2-axiom / halo2curves,
Cargo.toml:19-28) and no Rust toolchain exists here.  What *is* pinned: the
constants and known answers of SURVEY.md App. B, which this module recomputes
from first principles.

Semantic spec followed: SURVEY.md App. A (halo2curves bn256 Fr/Fq/G1,
halo2-axiom arithmetic.rs best_multiexp/best_fft, poly/domain.rs
EvaluationDomain), reached from /root/reference/src/scaffold/mod.rs:260,273,296.
"""

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr
S = 28                      # two-adicity of Fr
GEN = 7                     # multiplicative generator of Fr
ROOT_OF_UNITY = pow(GEN, (R - 1) >> S, R)          # order 2^28
ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23  # Fr::ZETA
MONT_R = 1 << 256
G1_GEN = (1, 2)
B_COEFF = 3


# --------------------------------------------------------------------------- encodings
def to_mont(x, m):
    return (x * MONT_R) % m


def from_mont(x, m):
    return (x * pow(MONT_R, -1, m)) % m


def limbs4(x):
    """256-bit int -> 4 little-endian u64 limbs (the [u64;4] layout of halo2curves)."""
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs4(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


# --------------------------------------------------------------------------- G1 (affine, None = identity)
def g1_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % P == 0


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_mul(pt, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, pt)
        pt = g1_add(pt, pt)
        k >>= 1
    return acc


def g1_msm(scalars, points):
    acc = None
    for s, pt in zip(scalars, points):
        acc = g1_add(acc, g1_mul(pt, s))
    return acc


def g1_compress(pt):
    """halo2curves 0.3.x `to_bytes()`: x LE, top bit of byte 31 = lsb(y); identity = zeros
    (SURVEY.md App. A.2 -- recalled, unverified)."""
    if pt is None:
        return bytes(32)
    b = bytearray(pt[0].to_bytes(32, "little"))
    b[31] |= (pt[1] & 1) << 7
    return bytes(b)


# --------------------------------------------------------------------------- NTT / EvaluationDomain
def omega_for(log_n):
    """omega of order 2^log_n, as EvaluationDomain::new derives it (App. A.5)."""
    return pow(ROOT_OF_UNITY, 1 << (S - log_n), R)


def dft_naive(a, omega):
    n = len(a)
    out = []
    for j in range(n):
        wj = pow(omega, j, R)
        acc = 0
        w = 1
        for i in range(n):
            acc = (acc + a[i] * w) % R
            w = w * wj % R
        out.append(acc)
    return out


def eval_poly(a, x):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % R
    return acc


class Domain:
    """EvaluationDomain::new(j, k) scalars (App. A.5)."""

    def __init__(self, j, k):
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ek = k
        while (1 << ek) < self.n * (j - 1):
            ek += 1
        self.extended_k = ek
        self.omega_ext = omega_for(ek)
        self.omega = pow(self.omega_ext, 1 << (ek - k), R)
        self.omega_inv = pow(self.omega, -1, R)
        self.omega_ext_inv = pow(self.omega_ext, -1, R)
        self.ifft_divisor = pow(1 << k, -1, R)
        self.extended_ifft_divisor = pow(1 << ek, -1, R)
        self.g_coset = ZETA
        self.g_coset_inv = ZETA * ZETA % R
        self.t_evaluations = [
            pow((pow(ZETA * pow(self.omega_ext, i, R), self.n, R) - 1) % R, -1, R)
            for i in range(1 << (ek - k))
        ]

    def lagrange_to_coeff(self, a):
        return [x * self.ifft_divisor % R for x in dft_naive(a, self.omega_inv)]

    def coeff_to_lagrange(self, a):
        return dft_naive(a, self.omega)

    def coeff_to_extended(self, a):
        z = [1, self.g_coset, self.g_coset_inv]
        b = [x * z[i % 3] % R for i, x in enumerate(a)] + [0] * ((1 << self.extended_k) - len(a))
        return dft_naive(b, self.omega_ext)

    def extended_to_coeff(self, a):
        b = dft_naive(a, self.omega_ext_inv)
        z = [1, self.g_coset_inv, self.g_coset]
        b = [x * self.extended_ifft_divisor % R * z[i % 3] % R for i, x in enumerate(b)]
        return b[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, a):
        m = len(self.t_evaluations)
        return [x * self.t_evaluations[i % m] % R for i, x in enumerate(a)]
