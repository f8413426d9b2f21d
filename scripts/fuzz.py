"""Randomised differential test: the CUDA path (through the C ABI) against the oracle on random sizes,
distributions, tuning knobs and entry points.  Not part of pytest (minutes); run on a GPU box:
    python scripts/fuzz.py [seconds] [seed]"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import halo2_vectordb_b200 as h
from oracle import oracle as O, pyref as P

BIG = "--big" in sys.argv
sys.argv = [a for a in sys.argv if a != "--big"]
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rnd = random.Random(seed)
h.init(0)
t_end = time.time() + budget
stats = {}

def scalars(n, mode):
    if mode == 2:   # pathological: few distinct values
        vals = [rnd.choice([0, 1, 2, P.R - 1, P.R - 2, 1 << 253, rnd.randrange(P.R)]) for _ in range(8)]
        return O.fr_from_ints([rnd.choice(vals) for _ in range(n)])
    return O.fr_fill(n, rnd.randrange(1 << 30), mode=mode, lookup_bits=rnd.randrange(1, 20))

def bases_for(n):
    b = O.gen_bases(n, a=rnd.randrange(1, 1 << 40), b=rnd.randrange(1, 1 << 40))
    if rnd.random() < 0.3:      # duplicates, opposites, identities
        for _ in range(max(1, n // 8)):
            i, j = rnd.randrange(n), rnd.randrange(n)
            r = rnd.random()
            if r < 0.4: b[i] = b[j]
            elif r < 0.7:
                xy = O.g1_affine_to_ints(b[j])
                b[i] = O.g1_affine_from_ints(None if xy is None else (xy[0], (-xy[1]) % P.P))
            else: b[i] = 0
    return b

BIG_SRS = {}
it = 0
while time.time() < t_end:
    it += 1
    kind = rnd.choice(["commit", "commit", "multiexp", "ntt", "domain", "row2", "quotient", "permute"] if not BIG else ["commit_big", "domain_big", "dev_big"])
    stats[kind] = stats.get(kind, 0) + 1
    tune = (rnd.choice([-1, -1, 1, 3, 5, 17, 64]), rnd.choice([-1, -1, 0, 1]))
    h.set_tuning(*tune)
    case = None
    try:
        if kind == "commit":
            k = rnd.randrange(0, 13); n = 1 << k
            b = bases_for(n)
            srs = h.ParamsKZG(k, b if rnd.random() < 0.5 else None, b)
            cols = rnd.randrange(1, 7); ln = rnd.choice([n, n, rnd.randrange(0, n + 1)])
            data = [scalars(n, rnd.randrange(3))[:ln] for _ in range(cols)]
            case = ("commit", k, ln, cols, tune)
            got = srs.commit_batch(data)
            for g, c in zip(got, data):
                exp = O.best_multiexp_affine(c, b[:ln]) if ln else np.zeros(8, dtype=np.uint64)
                assert (g == exp).all(), ("commit", k, ln, cols)
            srs.close()
        elif kind == "commit_big":      # staged / double-buffered host path, closed-form check on a few columns
            k = rnd.randrange(12, 17); n = 1 << k
            if k not in BIG_SRS:
                BIG_SRS[k] = h.ParamsKZG(k, None, h.synthetic_bases(n))
            cols = rnd.randrange(1, 60)
            distinct = [scalars(n, rnd.randrange(3)) for _ in range(min(cols, 4))]
            data = [distinct[i % len(distinct)] for i in range(cols)]
            case = ("commit_big", k, cols, tune)
            got = BIG_SRS[k].commit_batch(data)
            exp = [O.msm_closed_form(c) for c in distinct]
            for i in range(cols):
                assert (got[i] == exp[i % len(distinct)]).all(), case
        elif kind == "domain_big":
            k = rnd.randrange(12, 17); cols = rnd.randrange(1, 24)
            d, od = h.EvaluationDomain(4, k), O.EvaluationDomain(4, k)
            distinct = [scalars(1 << k, rnd.randrange(3)) for _ in range(min(cols, 2))]
            data = [distinct[i % len(distinct)] for i in range(cols)]
            op = rnd.choice([h.OP_LAGRANGE_TO_COEFF, h.OP_COEFF_TO_EXTENDED, h.OP_COEFF_TO_LAGRANGE])
            case = ("domain_big", k, cols, op)
            outs = d.transform_batch(op, data)
            f = {h.OP_LAGRANGE_TO_COEFF: od.lagrange_to_coeff, h.OP_COEFF_TO_EXTENDED: od.coeff_to_extended,
                 h.OP_COEFF_TO_LAGRANGE: od.coeff_to_lagrange}[op]
            exp = [f(c) for c in distinct]
            for i in range(cols):
                assert (outs[i] == exp[i % len(distinct)]).all(), case
            d.close()
        elif kind == "dev_big":         # device-resident entry points with strides
            k = rnd.randrange(10, 15); n = 1 << k; cols = rnd.randrange(1, 12)
            if k not in BIG_SRS:
                BIG_SRS[k] = h.ParamsKZG(k, None, h.synthetic_bases(n))
            stride = n + rnd.choice([0, 0, 4, 64])
            data = [scalars(n, rnd.randrange(3)) for _ in range(cols)]
            buf = np.zeros((cols, stride, 4), dtype=np.uint64)
            for i, c in enumerate(data):
                buf[i, :n] = c
            d_in = h.DeviceBuffer(buf.nbytes); d_in.upload(buf)
            d_out = h.DeviceBuffer(cols * 64)
            ln = rnd.choice([n, rnd.randrange(1, n + 1)])
            case = ("dev_big", k, cols, stride, ln, tune)
            BIG_SRS[k].commit_batch_dev(d_in.ptr, stride, cols, ln, d_out.ptr)
            got = d_out.download((cols, 8))
            for i in range(cols):
                assert (got[i] == O.msm_closed_form(data[i][:ln])).all(), case
            d_in.free(); d_out.free()
        elif kind == "multiexp":
            n = rnd.choice([rnd.randrange(1, 64), rnd.randrange(1, 3000), rnd.randrange(1, 20000)])
            b = bases_for(n); s = scalars(n, rnd.randrange(3))
            case = ("multiexp", n, tune)
            assert (O.g1_to_affine(h.best_multiexp(s, b)) == O.best_multiexp_affine(s, b)).all(), ("multiexp", n)
        elif kind == "ntt":
            L = rnd.randrange(0, 17); a = scalars(1 << L, rnd.randrange(3))
            w = O.fr_from_ints([pow(P.omega_for(L), rnd.randrange(1, 1 << (L + 1), 2), P.R)])[0]   # any primitive root
            assert (h.best_fft(a, w, L) == O.best_fft(a, w, L)).all(), ("ntt", L)
        elif kind == "domain":
            k = rnd.randrange(1, 13); j = rnd.choice([2, 3, 4, 4, 4, 5, 6, 9])
            d, od = h.EvaluationDomain(j, k), O.EvaluationDomain(j, k)
            a = scalars(1 << k, rnd.randrange(3)); ext = scalars(1 << d.extended_k, rnd.randrange(3))
            ncol = rnd.randrange(1, 5)
            outs = d.transform_batch(h.OP_LAGRANGE_TO_COEFF, [a] * ncol)
            assert all((o == od.lagrange_to_coeff(a)).all() for o in outs)
            assert (d.coeff_to_extended(a) == od.coeff_to_extended(a)).all(), ("c2e", j, k)
            assert (d.extended_to_coeff(ext) == od.extended_to_coeff(ext)).all(), ("e2c", j, k)
            f = d.transform_batch(h.OP_DIVIDE_BY_VANISHING, [ext])[0]
            assert (f == od.extended_to_coeff(od.divide_by_vanishing_poly(ext))).all(), ("dvp", j, k)
            d.close()
        elif kind == "permute":         # lookup argument A', S': random sizes, heavy repetition, full-width values
            u = rnd.choice([rnd.randrange(1, 40), rnd.randrange(1, 3000), rnd.randrange(1, 40000)])
            pool = [rnd.choice([0, 1, P.R - 1, rnd.randrange(P.R), rnd.randrange(1 << 16)]) for _ in range(rnd.randrange(1, max(2, u)))]
            inp = [rnd.choice(pool) for _ in range(u)]
            distinct = list(set(inp))
            table = distinct + [rnd.choice(pool + [rnd.randrange(P.R)]) for _ in range(u - len(distinct))]
            rnd.shuffle(table)
            bad = rnd.random() < 0.1 and u > 1
            if bad:
                table = [v for v in table if v != inp[0]] ; table += [(inp[0] + 1) % P.R] * (u - len(table))
                bad = inp[0] not in table
            fi, ft = O.fr_from_ints(inp), O.fr_from_ints(table)
            case = ("permute", u, bad)
            try:
                ga, gs = h.permute_expression_pair(fi, ft)
                assert not bad, case
                oa, os_ = O.permute_expression_pair(fi, ft)
                assert (ga == oa).all() and (gs == os_).all(), case
            except ValueError:
                assert bad, case
        elif kind == "quotient":        # evaluate_h row loops, random shapes / strides / pathological values
            k = rnd.randrange(1, 11); j = rnd.choice([3, 4, 4, 4, 5, 6, 9])
            d, od = h.EvaluationDomain(j, k), O.EvaluationDomain(j, k)
            ne = 1 << d.extended_k
            ng = rnd.randrange(0, 6); nc = rnd.randrange(0, 8); chunk = max(1, j - 2)
            bf = rnd.randrange(0, max(1, min(8, (1 << k) - 2)))
            ns = (nc + chunk - 1) // chunk if nc else 0
            stride = ne + rnd.choice([0, 0, 2, 32])
            mk = lambda c: [scalars(ne, rnd.randrange(3)) for _ in range(c)]
            q, a, sg, z, misc = mk(ng), mk(max(ng, nc)), mk(nc), mk(ns), mk(9)
            yv = scalars(3, 0)
            case = ("quotient", j, k, ng, nc, bf, stride)
            def dev(arrs):
                buf = h.DeviceBuffer(max(1, len(arrs)) * stride * 32)
                for i, c in enumerate(arrs):
                    buf.upload(c, offset=i * stride * 32)
                return buf
            dq, da, ds, dz, dm, dh = dev(q), dev(a), dev(sg), dev(z), dev(misc[1:]), dev(misc[:1])
            m = lambda i: dm.ptr + (i - 1) * stride * 32
            want = misc[0]
            d.quotient_gates(dh.ptr, yv[0], ng, dq.ptr, stride, da.ptr, stride)
            want = od.quotient_gates(want, yv[0], np.stack(q) if ng else [], np.stack(a[:ng]) if ng else [])
            if nc:
                d.quotient_permutation(dh.ptr, yv[0], yv[1], yv[2], nc, chunk, da.ptr, stride, ds.ptr, stride, dz.ptr, stride,
                                       m(1), m(2), m(3), bf)
                want = od.quotient_permutation(want, yv[0], yv[1], yv[2], chunk, np.stack(a[:nc]), np.stack(sg), np.stack(z),
                                               misc[1], misc[2], misc[3], bf)
            d.quotient_lookup(dh.ptr, yv[0], yv[1], yv[2], m(4), m(5), m(6), m(7), m(8), m(1), m(2), m(3))
            want = od.quotient_lookup(want, yv[0], yv[1], yv[2], misc[4], misc[5], misc[6], misc[7], misc[8], misc[1], misc[2], misc[3])
            assert (dh.download((ne, 4)) == want).all(), case
            for b_ in (dq, da, ds, dz, dm, dh):
                b_.free()
            d.close()
        else:
            n = rnd.choice([rnd.randrange(1, 40), rnd.randrange(1, 5000), rnd.randrange(1, 70000)])
            a = scalars(n, rnd.randrange(3))
            x = scalars(2, 0)
            assert (h.eval_polynomial_batch([a], x)[0, 1] == O.fr_eval_poly(a, x[1])).all(), ("eval", n)
            assert (h.batch_invert(a) == O.fr_batch_invert(a)).all(), ("binv", n)
            num, den = scalars(n, 0), scalars(n, 0)
            assert (h.grand_product(num, den) == O.fr_grand_product(num, den)).all(), ("gp", n)
            bb = rnd.choice([x[0], np.zeros(4, dtype=np.uint64)])
            assert (h.kate_division(a, bb) == O.fr_kate_division(a, bb)).all(), ("kate", n)
    except AssertionError as e:
        print("MISMATCH at iteration", it, e, "seed", seed, flush=True)
        sys.exit(1)
    except Exception as e:
        print("ERROR at iteration", it, kind, case, "seed", seed, repr(e)[:300], flush=True)
        sys.exit(2)
h.set_tuning(-1, -1)
print(f"fuzz ok: {it} iterations in {budget:.0f} s, seed {seed}: {stats}")
