"""A small satisfied PLONKish circuit in plain Python integers, shaped like halo2-base's constraint system
(vertical gates q*(a + b*c - d), one range lookup, a permutation argument over several columns), used to check
the quotient-evaluation row loops by their defining property instead of by restating code:

  for a satisfied circuit the folded expression  E(X) = fold_y(terms)  vanishes on the 2^k domain, so
  h = E / (X^n - 1) is a polynomial of degree < (j-1) n and  h(x) * (x^n - 1) == E(x)  at any x.

`expected_at(x)` recomputes E(x) from the column polynomials with big integers (Horner at x * omega^rot), in the
order of halo2-axiom plonk/evaluation.rs evaluate_h [UPSTREAM]: gates, permutation, lookups.
TEST INFRASTRUCTURE ONLY.
"""
import random

from oracle import pyref as P

R = P.R
DELTA = pow(P.GEN, 1 << P.S, R)          # Fr::DELTA


class ToyCircuit:
    def __init__(self, k, seed, n_gate_cols=2, blinding_factors=5, degree=4, lookup_bits=4):
        rng = random.Random(seed)
        self.k, self.n = k, 1 << k
        n = self.n
        self.degree = degree                      # cs.degree(); EvaluationDomain::new(j = degree, k)
        self.chunk_len = degree - 2
        self.bf = blinding_factors
        u = self.u = n - (blinding_factors + 1)   # usable rows; l_last sits on row u
        self.dom = P.Domain(degree, k)
        w = self.dom.omega
        tsize = 1 << lookup_bits
        assert tsize <= u

        # --- columns: gate advice columns, one lookup advice column, one fixed column (all in the permutation)
        self.n_gates = n_gate_cols
        adv = [[rng.randrange(R) for _ in range(n)] for _ in range(n_gate_cols)]
        look = [rng.randrange(tsize) for _ in range(u)] + [rng.randrange(R) for _ in range(n - u)]
        fixed = [rng.randrange(tsize) for _ in range(u)] + [0] * (n - u)
        cols = adv + [look, fixed]
        nc = len(cols)
        # copy constraints over cells the gates leave free (rows = 1, 2 mod 4 of the gate columns; any usable row else)
        free = [(c, r) for c in range(n_gate_cols) for r in range(u) if r % 4 in (1, 2)]
        free += [(c, r) for c in (nc - 2, nc - 1) for r in range(u)]
        rng.shuffle(free)
        mapping = {(c, r): (c, r) for c in range(nc) for r in range(n)}
        pos = 0
        while pos + 4 <= len(free) // 2:
            size = rng.randrange(2, 5)
            cells = free[pos:pos + size]
            pos += size
            val = rng.randrange(tsize)
            for i, (c, r) in enumerate(cells):
                cols[c][r] = val
                mapping[(c, r)] = cells[(i + 1) % size]
        # gates: q = 1 on rows 0 mod 4
        q = [[0] * n for _ in range(n_gate_cols)]
        for c in range(n_gate_cols):
            for r in range(0, u - 3, 4):
                q[c][r] = 1
                cols[c][r + 3] = (cols[c][r] + cols[c][r + 1] * cols[c][r + 2]) % R
        self.q, self.cols = q, cols
        self.sigma = [[pow(DELTA, mapping[(c, r)][0], R) * pow(w, mapping[(c, r)][1], R) % R for r in range(n)] for c in range(nc)]

        # --- challenges
        self.theta, self.beta, self.gamma, self.y = (rng.randrange(1, R) for _ in range(4))
        beta, gamma = self.beta, self.gamma

        # --- permutation grand products, chunk_len columns each; the running product continues from set to set
        self.n_sets = (nc + self.chunk_len - 1) // self.chunk_len
        self.z = []
        last = 1
        for s in range(self.n_sets):
            cs = range(s * self.chunk_len, min((s + 1) * self.chunk_len, nc))
            z = [last]
            for r in range(u):
                num = den = 1
                for c in cs:
                    num = num * (cols[c][r] + beta * pow(DELTA, c, R) * pow(w, r, R) + gamma) % R
                    den = den * (cols[c][r] + beta * self.sigma[c][r] + gamma) % R
                z.append(z[-1] * num % R * pow(den, -1, R) % R)
            last = z[u]
            z += [rng.randrange(R) for _ in range(n - u - 1)]
            self.z.append(z)
        assert last == 1, "copy constraints not satisfied"

        # --- lookup: input = lookup advice column, table = fixed table column (halo2-base range lookup)
        table = list(range(tsize)) + [0] * (u - tsize) + [0] * (n - u)
        self.l_input, self.l_table = look, table
        a_sorted = sorted(look[:u])
        leftover = sorted(table[:u])
        s_perm = [None] * u
        for r in range(u):
            if r == 0 or a_sorted[r] != a_sorted[r - 1]:
                s_perm[r] = a_sorted[r]
                leftover.remove(a_sorted[r])
        it = iter(leftover)
        for r in range(u):
            if s_perm[r] is None:
                s_perm[r] = next(it)
        self.perm_input = a_sorted + [rng.randrange(R) for _ in range(n - u)]
        self.perm_table = s_perm + [rng.randrange(R) for _ in range(n - u)]
        zl = [1]
        for r in range(u):
            num = (look[r] + beta) * (table[r] + gamma) % R
            den = (self.perm_input[r] + beta) * (self.perm_table[r] + gamma) % R
            zl.append(zl[-1] * num % R * pow(den, -1, R) % R)
        assert zl[u] == 1, "lookup not satisfied"
        self.z_lookup = zl + [rng.randrange(R) for _ in range(n - u - 1)]

        self.l0 = [1] + [0] * (n - 1)
        self.l_last = [1 if r == u else 0 for r in range(n)]
        self.l_active = [1 if r < u else 0 for r in range(n)]

    # every column, Lagrange basis, in a fixed order (name -> list of n ints)
    def lagrange_columns(self):
        d = {}
        for j in range(self.n_gates):
            d[f"q{j}"] = self.q[j]
        for c, col in enumerate(self.cols):
            d[f"col{c}"] = col
            d[f"sigma{c}"] = self.sigma[c]
        for s, z in enumerate(self.z):
            d[f"z{s}"] = z
        d.update(l_input=self.l_input, l_table=self.l_table, perm_input=self.perm_input, perm_table=self.perm_table,
                 z_lookup=self.z_lookup, l0=self.l0, l_last=self.l_last, l_active=self.l_active)
        return d

    def break_gate(self):
        """Make one gate row unsatisfied (the quotient then stops being a polynomial of the right degree)."""
        self.cols[0][3] = (self.cols[0][3] + 1) % R

    def expected_at(self, coeffs, x):
        """E(x) from coefficient-form columns (name -> ints), folded in evaluate_h's order."""
        w = self.dom.omega
        y, beta, gamma = self.y, self.beta, self.gamma

        def ev(name, rot=0):
            return P.eval_poly(coeffs[name], x * pow(w, rot, R) % R)

        acc = 0

        def fold(t):
            nonlocal acc
            acc = (acc * y + t) % R

        for j in range(self.n_gates):
            c = f"col{j}"
            fold(ev(f"q{j}") * (ev(c) + ev(c, 1) * ev(c, 2) - ev(c, 3)))
        nc = len(self.cols)
        l0, l_last, l_active = ev("l0"), ev("l_last"), ev("l_active")
        last_rot = -(self.bf + 1)
        fold(l0 * (1 - ev("z0")))
        zl = ev(f"z{self.n_sets - 1}")
        fold(l_last * (zl * zl - zl))
        for s in range(1, self.n_sets):
            fold(l0 * (ev(f"z{s}") - ev(f"z{s - 1}", last_rot)))
        for s in range(self.n_sets):
            left, right = ev(f"z{s}", 1), ev(f"z{s}")
            for c in range(s * self.chunk_len, min((s + 1) * self.chunk_len, nc)):
                left = left * (ev(f"col{c}") + beta * ev(f"sigma{c}") + gamma) % R
                right = right * (ev(f"col{c}") + pow(DELTA, c, R) * beta * x + gamma) % R
            fold(l_active * (left - right))
        z, ap, sp = ev("z_lookup"), ev("perm_input"), ev("perm_table")
        fold(l0 * (1 - z))
        fold(l_last * (z * z - z))
        fold(l_active * (ev("z_lookup", 1) * (ap + beta) * (sp + gamma) - z * (ev("l_input") + beta) * (ev("l_table") + gamma)))
        fold(l0 * (ap - sp))
        fold(l_active * (ap - sp) * (ap - ev("perm_input", -1)))
        return acc % R
