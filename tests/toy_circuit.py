"""A satisfied circuit with the constraint-system shape halo2-base builds (TEST INFRASTRUCTURE ONLY).

[UPSTREAM] halo2-base gates/flex_gate.rs + gates/range.rs, as configured by the reference's scaffold
(/root/reference/src/scaffold/mod.rs:345-402): `n_gate_cols` "basic gate" advice columns with one selector each
(q * (a + b*c - d) on four consecutive rows), `n_lookup_cols` lookup-advice columns looked up in one fixed table column
of 2^lookup_bits values, one fixed constants column, one instance column; every advice column, the constants column and
the instance column take part in the permutation argument.  The witness is random but satisfies every gate, copy
constraint and lookup, so an honest proof verifies; `break_*` helpers make one constraint fail.

Produces plain Python integers (Lagrange basis, n = 2^k per column) and the `cs` description both provers take.
"""
import random

from oracle import pyref as P

R = P.R
DELTA = pow(P.GEN, 1 << P.S, R)


class Toy:
    def __init__(self, k, seed, n_gate_cols=2, n_lookup_cols=1, lookup_bits=4, n_public=3, blinding_factors=5, degree=4):
        rng = random.Random(seed)
        self.k, self.n = k, 1 << k
        n = self.n
        bf = blinding_factors
        u = self.u = n - (bf + 1)
        G, Lc = n_gate_cols, n_lookup_cols
        tsize = 1 << lookup_bits
        assert tsize <= u and 4 * 4 <= u
        omega = P.Domain(degree, k).omega
        A = G + Lc
        F = G + 2                       # selectors, table, constants
        TABLE, CONST = G, G + 1
        # permutation order: constants column, gate advice, lookup advice, instance (the order of enable_equality calls)
        perm = [(1, CONST)] + [(0, c) for c in range(A)] + [(2, 0)]
        pidx = {col: i for i, col in enumerate(perm)}
        self.cs = dict(
            k=k, degree=degree, blinding_factors=bf, n_advice=A, n_fixed=F, n_instance=1,
            gates=[(c, c) for c in range(G)], lookups=[(G + l, TABLE) for l in range(Lc)], permutation=perm,
            advice_queries=[q for c in range(G) for q in ((c, 0), (c, 1), (c, 2), (c, 3))] + [(G + l, 0) for l in range(Lc)],
            fixed_queries=[(CONST, 0), (TABLE, 0)] + [(c, 0) for c in range(G)],
            instance_queries=[(0, 0)])
        adv = [[0] * n for _ in range(A)]
        fixed = [[0] * n for _ in range(F)]
        fixed[TABLE][:tsize] = list(range(tsize))
        public = [rng.randrange(R) for _ in range(n_public)]
        inst = public + [0] * (n - n_public)
        # gates on rows 4g .. 4g+3; d of an even gate is copied into a of the next one (values flow down the column)
        n_g = (u - 3) // 4 + (1 if (u - 3) % 4 else 0)
        n_g = len(range(0, u - 3, 4))
        free = []                       # cells whose value is ours to choose: (kind, col, row)
        chained = set()
        for c in range(G):
            for g in range(n_g):
                r = 4 * g
                fixed[c][r] = 1
                if g % 2 == 1 and g >= 1:
                    chained.add((c, r))  # a = previous gate's d
                else:
                    free.append((0, c, r))
                free.append((0, c, r + 1))
                free.append((0, c, r + 2))
        for l in range(Lc):
            free += [(0, G + l, r) for r in range(u)]
        consts = [(1, CONST, r) for r in range(min(u, 64))]
        pub_cells = [(2, 0, r) for r in range(n_public)]
        # copy-constraint cycles
        rng.shuffle(free)
        cycles = []
        pos = 0
        for cell in pub_cells:          # every public input is copied into two advice cells
            cycles.append([cell, free[pos], free[pos + 1]])
            pos += 2
        for cell in consts[:16]:        # constants feed advice cells
            cycles.append([cell, free[pos]])
            pos += 1
        budget = len(free) // 3
        while pos + 4 <= budget:
            size = rng.randrange(2, 5)
            cycles.append(free[pos:pos + size])
            pos += size
        in_cycle = set()

        def setv(cell, v):
            kind, c, r = cell
            if kind == 0:
                adv[c][r] = v
            elif kind == 1:
                fixed[c][r] = v
            else:
                assert inst[r] == v

        for cell in free:
            kind, c, r = cell
            adv[c][r] = rng.randrange(tsize) if c >= G else rng.randrange(R)
        for cyc in cycles:
            small = any(kind == 0 and c >= G for (kind, c, r) in cyc)
            pubs = [cell for cell in cyc if cell[0] == 2]
            v = inst[pubs[0][2]] if pubs else (rng.randrange(tsize) if small else rng.randrange(R))
            if pubs and small:          # a public input cannot sit in a lookup cell unless it is small: swap the cell out
                cyc[:] = [cell for cell in cyc if not (cell[0] == 0 and cell[1] >= G)]
            for cell in cyc:
                setv(cell, v)
                in_cycle.add(cell)
        for c in range(G):
            for g in range(n_g):
                r = 4 * g
                if (c, r) in chained:
                    adv[c][r] = adv[c][r - 1]
                    cycles.append([(0, c, r - 1), (0, c, r)])
                adv[c][r + 3] = (adv[c][r] + adv[c][r + 1] * adv[c][r + 2]) % R
        # sigma polynomials: sigma[col][row] = delta^col' * w^row' for the next cell (col', row') of the cycle
        wp = [1] * n
        for i in range(1, n):
            wp[i] = wp[i - 1] * omega % R
        dp = [pow(DELTA, i, R) for i in range(len(perm))]
        sigma = [[dp[i] * wp[r] % R for r in range(n)] for i in range(len(perm))]
        for cyc in cycles:
            if len(cyc) < 2:
                continue
            for i, cell in enumerate(cyc):
                nk, nc, nr = cyc[(i + 1) % len(cyc)]
                sigma[pidx[(cell[0], cell[1])]][cell[2]] = dp[pidx[(nk, nc)]] * wp[nr] % R
        self.advice, self.fixed, self.sigma, self.instances = adv, fixed, sigma, [public]
        self.vk_repr = rng.randrange(R)
        self.G, self.Lc, self.tsize = G, Lc, tsize
        self._cycles = cycles

    # --- ways to make the statement false (the prover still runs; the verifier must reject)
    def break_gate(self):
        self.advice[0][3] = (self.advice[0][3] + 1) % R

    def break_copy(self):
        cyc = next(c for c in self._cycles if len(c) >= 2 and c[0][0] == 0 and c[0][1] < self.G and c[0][2] % 4 in (1, 2))
        _, c, r = cyc[0]
        self.advice[c][r] = (self.advice[c][r] + 1) % R
        g = r - r % 4
        self.advice[c][g + 3] = (self.advice[c][g] + self.advice[c][g + 1] * self.advice[c][g + 2]) % R

    def break_public_input(self):
        self.instances = [[(self.instances[0][0] + 1) % R] + list(self.instances[0][1:])]
