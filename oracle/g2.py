"""BN254 G2 (the twist y^2 = x^3 + 3/(9 + i) over Fq2 = Fq[i]/(i^2 + 1)) in plain Python integers
(TEST INFRASTRUCTURE ONLY).  Needed for the verifier side of `ParamsKZG` ([UPSTREAM] poly/kzg/commitment.rs:
`g2`, `s_g2 = s * g2`, written after the G1 bases by `ParamsKZG::write`; reached from
/root/reference/src/scaffold/mod.rs:260 `gen_srs`).  Points are ((x0, x1), (y0, y1)) or None for the identity."""
from .pyref import P as Q

G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * n % Q, (-a[1]) * n % Q)


B2 = f2_mul((3, 0), f2_inv((9, 1)))


def is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), B2)


def add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_mul((2, 0), y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def mul(pt, k):
    acc = None
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc
