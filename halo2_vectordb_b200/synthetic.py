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
