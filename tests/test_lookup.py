"""permute_expression_pair (the lookup argument's A', S'; create_proof step 5).
CPU: the oracle against a line-by-line Python model of the upstream routine and against the argument's own
requirements.  GPU: the CUDA path (sort + scans + scatters) bit-exact against the oracle through the C ABI."""
import random
from collections import OrderedDict

import numpy as np
import pytest

from common import fr_arr
from oracle import oracle as O
from oracle import pyref as P

R = P.R


def model_permute(inp, table):
    """plonk/lookup/prover.rs permute_expression_pair [UPSTREAM], transliterated on Python ints"""
    a = sorted(inp)
    leftover = OrderedDict()
    for v in sorted(table):                       # BTreeMap<value, count>
        leftover[v] = leftover.get(v, 0) + 1
    s = [0] * len(a)
    repeated = []
    for row, v in enumerate(a):
        if row == 0 or v != a[row - 1]:
            s[row] = v
            if leftover.get(v, 0) == 0:
                raise ValueError("ConstraintSystemFailure")
            leftover[v] -= 1
        else:
            repeated.append(row)
    for v, cnt in leftover.items():
        for _ in range(cnt):
            s[repeated.pop()] = v
    assert not repeated
    return a, s


def make_case(rng, u, kind):
    if kind == "range":          # halo2-base range lookup: values < 2^bits against the table 0..2^bits-1 padded with zeros
        bits = max(1, min(u.bit_length() - 1, rng.randrange(1, 16)))
        table = list(range(1 << bits)) + [0] * (u - (1 << bits))
        inp = [rng.randrange(1 << bits) for _ in range(u)]
    elif kind == "wide":         # full-width field elements, table = a shuffle of the distinct inputs plus fillers
        pool = [rng.randrange(R) for _ in range(max(1, u // 3))]
        inp = [rng.choice(pool) for _ in range(u)]
        distinct = list(set(inp))
        table = distinct + [rng.randrange(R) for _ in range(u - len(distinct))]
        rng.shuffle(table)
    else:                        # constant input
        inp = [5] * u
        table = [5] + [rng.randrange(R) for _ in range(u - 1)]
        rng.shuffle(table)
    return inp, table


def check_argument(a, s, inp, table):
    assert a == sorted(inp) and sorted(s) == sorted(table)
    assert all(a[i] == s[i] or (i > 0 and a[i] == a[i - 1]) for i in range(len(a)))


@pytest.mark.parametrize("u,kind", [(1, "const"), (2, "range"), (7, "wide"), (58, "range"), (250, "wide"), (1018, "range"), (4090, "const")])
def test_oracle_permute_matches_model(u, kind):
    rng = random.Random(u)
    inp, table = make_case(rng, u, kind)
    a, s = O.permute_expression_pair(fr_arr(inp), fr_arr(table))
    ma, ms = model_permute(inp, table)
    assert O.fr_to_ints(a) == ma and O.fr_to_ints(s) == ms
    check_argument(ma, ms, inp, table)


def test_oracle_permute_rejects_missing_value():
    with pytest.raises(ValueError):
        O.permute_expression_pair(fr_arr([1, 2, 9]), fr_arr([1, 2, 3]))
    a, s = O.permute_expression_pair(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64))
    assert a.shape == (0, 4) and s.shape == (0, 4)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("u,kind", [(1, "const"), (2, "range"), (3, "wide"), (58, "range"), (511, "wide"), (512, "range"), (1023, "wide"),
                                    (1024, "range"), (1025, "wide"), (2047, "const"), (8186, "range"), (65530, "range"), (65530, "wide"),
                                    (100000, "wide")])
def test_gpu_permute_matches_oracle(h2v, u, kind):
    rng = random.Random(u * 7 + len(kind))
    inp, table = make_case(rng, u, kind)
    fi, ft = fr_arr(inp), fr_arr(table)
    ga, gs = h2v.permute_expression_pair(fi, ft)
    oa, os_ = O.permute_expression_pair(fi, ft)
    assert np.array_equal(ga, oa) and np.array_equal(gs, os_)


@pytest.mark.gpu
def test_gpu_permute_k20_properties_and_dev_entry(h2v):
    """2^20 - 6 usable rows (LOOKUP_BITS = 19 like the SIFT config): device-resident call, checked by the argument's
    own requirements (sorted, same multiset, equal-or-repeated) with numpy on the canonical values, and against the oracle"""
    u = (1 << 20) - 6
    rng = np.random.default_rng(5)
    bits = 19
    inp = rng.integers(0, 1 << bits, u, dtype=np.uint64)
    table = np.concatenate([np.arange(1 << bits, dtype=np.uint64), np.zeros(u - (1 << bits), dtype=np.uint64)])
    canon = lambda v: np.stack([v, np.zeros_like(v), np.zeros_like(v), np.zeros_like(v)], axis=1)
    fi, ft = O.to_mont(canon(inp)), O.to_mont(canon(table))
    d_in, d_t, d_a, d_s = (h2v.DeviceBuffer(u * 32) for _ in range(4))
    d_in.upload(fi); d_t.upload(ft)
    h2v.permute_expression_pair_dev(d_in.ptr, d_t.ptr, u, d_a.ptr, d_s.ptr)
    ga, gs = d_a.download((u, 4)), d_s.download((u, 4))
    ca, cs = O.from_mont(ga), O.from_mont(gs)
    assert not ca[:, 1:].any() and not cs[:, 1:].any()
    a, s = ca[:, 0], cs[:, 0]
    assert np.array_equal(a, np.sort(inp)) and np.array_equal(np.sort(s), np.sort(table))
    first = np.concatenate([[True], a[1:] != a[:-1]])
    assert np.array_equal(s[first], a[first])
    oa, os_ = O.permute_expression_pair(fi, ft)
    assert np.array_equal(ga, oa) and np.array_equal(gs, os_)


@pytest.mark.gpu
def test_gpu_permute_errors(h2v):
    with pytest.raises(ValueError):
        h2v.permute_expression_pair(fr_arr([1, 2, 9]), fr_arr([1, 2, 3]))
    with pytest.raises(ValueError):
        h2v.permute_expression_pair(fr_arr([1, 2]), fr_arr([1, 2, 3]))
    a, s = h2v.permute_expression_pair(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64))
    assert a.shape == (0, 4)
    # the library is still usable after the error
    a, s = h2v.permute_expression_pair(fr_arr([2, 2, 1]), fr_arr([3, 1, 2]))
    assert O.fr_to_ints(a) == [1, 2, 2] and O.fr_to_ints(s) == [1, 2, 3]


@pytest.mark.gpu
def test_gpu_permute_batch_matches_per_lookup_oracle(h2v):
    """all lookups of a proof phase in one set of launches: five inputs over two distinct tables (one shared by three of
    them, as halo2-base's single range table is), ragged sizes, against the oracle one lookup at a time"""
    for u in (37, 1000, 5000):
        rng = np.random.default_rng(u)
        bits = 5 if u < 100 else 9
        stride = u + 11
        canon = lambda v: np.stack([v, np.zeros_like(v), np.zeros_like(v), np.zeros_like(v)], axis=1)
        t0 = np.concatenate([np.arange(1 << bits, dtype=np.uint64), np.zeros(u - (1 << bits), dtype=np.uint64)])
        t1 = rng.permutation(np.concatenate([np.arange(3, 3 + (1 << bits), dtype=np.uint64), np.full(u - (1 << bits), 7, dtype=np.uint64)]))
        tables = [O.to_mont(canon(t0)), O.to_mont(canon(t1))]
        which = [0, 1, 0, 0, 1]
        inputs = []
        for l, w in enumerate(which):
            lo = 0 if w == 0 else 3
            v = rng.integers(lo, lo + (1 << bits), u, dtype=np.uint64)
            if l == 2:
                v[:] = v[0]                      # one value repeated on every row
            inputs.append(O.to_mont(canon(v)))
        d_t = [h2v.DeviceBuffer(u * 32) for _ in tables]
        for b, t in zip(d_t, tables):
            b.upload(t)
        d_i = [h2v.DeviceBuffer(u * 32) for _ in inputs]
        for b, v in zip(d_i, inputs):
            b.upload(v)
        L = len(inputs)
        d_a, d_s = h2v.DeviceBuffer(L * stride * 32), h2v.DeviceBuffer(L * stride * 32)
        h2v.permute_expression_pair_batch_dev([b.ptr for b in d_i], [d_t[w].ptr for w in which], u, d_a.ptr, stride, d_s.ptr, stride)
        ga, gs = d_a.download((L, stride, 4)), d_s.download((L, stride, 4))
        for l, w in enumerate(which):
            oa, os_ = O.permute_expression_pair(inputs[l], tables[w])
            assert np.array_equal(ga[l, :u], oa) and np.array_equal(gs[l, :u], os_), (u, l)
    # a value missing from ONE of the tables fails the whole batch, as the single call does
    bad = O.to_mont(canon(np.full(u, 9999, dtype=np.uint64)))
    d_i[3].upload(bad)
    with pytest.raises(ValueError):
        h2v.permute_expression_pair_batch_dev([b.ptr for b in d_i], [d_t[w].ptr for w in which], u, d_a.ptr, stride, d_s.ptr, stride)
