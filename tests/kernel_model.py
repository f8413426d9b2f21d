"""Line-by-line Python models of the index arithmetic in the CUDA kernels (ntt.cuh, msm.cuh).

They run the same tiling / addressing / run-flush / reduction-tree logic over Python integers so
that the *structure* of the kernels can be checked on a CPU-only box (the kernels themselves are
checked on the GPU by the `-m gpu` parity tests).  Field = Z_r, "group" = (Z_r, +) with point k
standing for k*G, so an MSM result is sum s_i k_i mod r.
"""
from oracle import pyref as P

R = P.R


def brev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def swz(i):
    return i ^ ((i >> 3) & 7) ^ ((i >> 6) & 7) ^ ((i >> 9) & 7)


def ntt_plan(L, max_s=9):
    Pn = max(1, (L + max_s - 1) // max_s)
    base, rem = divmod(L, Pn)
    return [base + (1 if i < rem else 0) for i in range(Pn)]


def ntt_pass(src, dst, tw, L, t0, S, logT, first, last, n_in, n_out, pre, post):
    """mirror of ntt_pass_kernel over all tiles / threads of one column"""
    T = 1 << logT
    nelem = 1 << (S + logT)
    nthr = max(nelem >> 3, 1)
    tiles = 1 << (L - S - logT)
    for tile in range(tiles):
        sm = {}
        hi = lo_tile = 0
        if first:
            H = L - S
            for e in range(nelem):
                q, midr = e & (T - 1), e >> logT
                si = (midr << H) + tile * T + q
                x = 0
                if si < n_in:
                    x = src[si]
                    if pre:
                        x = x * pre[si % len(pre)] % R
                sm[swz((brev(midr, S) << logT) | q)] = x
        else:
            hi = tile >> (t0 - logT)
            lo_tile = tile & ((1 << (t0 - logT)) - 1)
            base = (hi << (t0 + S)) + lo_tile * T
            for e in range(nelem):
                q, mid = e & (T - 1), e >> logT
                sm[swz(e)] = dst[base + (mid << t0) + q]
        b = 0
        while b < S:
            bp = b if b + 3 <= S else S - 3
            u0 = b - bp
            for w in range(nelem >> 3):
                q, rest = w & (T - 1), w >> logT
                low = rest & ((1 << bp) - 1)
                mid_base = low | ((rest >> bp) << (bp + 3))
                x = [sm[swz(((mid_base + (k << bp)) << logT) | q)] for k in range(8)]
                lo = 0 if first else lo_tile * T + q
                jlow = lo + (low << t0)
                pad4 = first and 4 * n_in <= (1 << L) and b == 0
                if pad4:
                    assert all(x[k] == 0 for k in (1, 2, 3, 5, 6, 7))
                    x[1] = x[2] = x[3] = x[0]
                    x[5] = x[6] = x[7] = x[4]
                for U in range(3):
                    if U < u0 or (pad4 and U < 2):
                        continue
                    t = t0 + bp + U
                    e_base = jlow << (L - 1 - t)
                    for k in range(8):
                        if (k >> U) & 1:
                            continue
                        if U == 0 and t == 0:
                            tv = x[k + 1]
                        else:
                            wv = tw[e_base + ((k & ((1 << U) - 1)) << (L - 1 - U))]
                            tv = x[k + (1 << U)] * wv % R
                        x[k + (1 << U)] = (x[k] - tv) % R
                        x[k] = (x[k] + tv) % R
                for k in range(8):
                    sm[swz(((mid_base + (k << bp)) << logT) | q)] = x[k]
            b += 3
        if first:
            H = L - S
            for e in range(nelem):
                mid, q = e & ((1 << S) - 1), e >> S
                j = (brev(tile * T + q, H) << S) | mid
                if last and j >= n_out:
                    continue
                x = sm[swz((mid << logT) | q)]
                if last and post:
                    x = x * post[j % len(post)] % R
                dst[j] = x
        else:
            jbase = (hi << (t0 + S)) + lo_tile * T
            for e in range(nelem):
                q, mid = e & (T - 1), e >> logT
                j = jbase + (mid << t0) + q
                if last and j >= n_out:
                    continue
                x = sm[swz(e)]
                if last and post:
                    x = x * post[j % len(post)] % R
                dst[j] = x


def run_ntt(src, L, omega, n_in=None, n_out=None, pre=None, post=None, max_s=9, max_log=11):
    """mirror of run_ntt in h2v.cu (L >= 3)"""
    N = 1 << L
    n_in = N if n_in is None else n_in
    n_out = N if n_out is None else n_out
    tw = [pow(omega, i, R) for i in range(N // 2)]
    dst = [None] * N
    t0 = 0
    plan = ntt_plan(L, max_s)
    for i, S in enumerate(plan):
        first, last = i == 0, i == len(plan) - 1
        logT = min(max_log - S, L - S) if first else min(max_log - S, t0)
        ntt_pass(src, dst, tw, L, t0, S, logT, first, last, n_in, n_out, pre, post)
        t0 += S
    return dst


# ----------------------------------------------------------------------------- MSM model
def windows_for(c):
    return (255 + c - 1) // c


def msm_model(scalars, points, c, precomp, chunk, n_cols=1, log_seg=5):
    """scalars: list of columns of canonical ints; points: ints k_i (point = k_i * G).
    Returns per-column result as an int (multiple of G), following msm.cuh kernel by kernel."""
    n = len(points)
    W = windows_for(c)
    G = 1 if precomp else W
    nb = 1 << (c - 1)
    half = nb - 1
    K = sum(half << (j * c) for j in range(W))
    table = [(k << (j * c)) % R for j in range(W) for k in points] if precomp else list(points)
    pstride = n
    n_buckets = n_cols * G * nb
    counts = [0] * n_buckets
    keys = {}
    for col in range(n_cols):
        for i in range(n):
            t = scalars[col][i] + K
            assert t < 1 << 288
            for j in range(W):
                d = (t >> (j * c)) & ((1 << c) - 1)
                sd = d - half
                key = None
                if sd != 0:
                    mag = abs(sd) - 1
                    key = (mag, sd < 0)
                    g = j if G > 1 else 0
                    counts[(col * G + g) * nb + mag] += 1
                keys[(col, j, i)] = key
            assert (t >> (W * c)) == 0
    offsets = [0]
    for v in counts:
        offsets.append(offsets[-1] + v)
    cursor = offsets[:-1]
    M = offsets[-1]
    entries = [None] * M
    for col in range(n_cols):
        for j in range(W):
            for i in range(n):
                key = keys[(col, j, i)]
                if key is None:
                    continue
                mag, neg = key
                g = j if G > 1 else 0
                b = (col * G + g) * nb + mag
                pref = i if G > 1 else j * pstride + i
                entries[cursor[b]] = (pref, neg, b)
                cursor[b] += 1
    buckets = [None] * n_buckets
    nthreads = (n_cols * W * n + chunk - 1) // chunk
    edges = [None] * (2 * nthreads)
    for t in range(nthreads):
        start = t * chunk
        if start >= M:
            continue
        end = min(start + chunk, M)
        prev_b = entries[start - 1][2] if start > 0 else -1
        next_b = entries[end][2] if end < M else -1
        acc = 0
        cur_b = entries[start][2]
        first_run = True
        for e in range(start, end):
            nxt_b = entries[e + 1][2] if e + 1 < end else -1
            pref, neg, _ = entries[e]
            acc = (acc + (-table[pref] if neg else table[pref])) % R
            if nxt_b != cur_b:
                last_run = e + 1 == end
                starts_before = first_run and prev_b == cur_b
                continues_after = last_run and next_b == cur_b
                if not starts_before and not continues_after:
                    assert buckets[cur_b] is None
                    buckets[cur_b] = acc
                elif starts_before:
                    edges[2 * t] = acc
                else:
                    edges[2 * t + 1] = acc
                acc = 0
                cur_b = nxt_b
                first_run = False
    for b in range(n_buckets):
        s, e = offsets[b], offsets[b + 1]
        if s == e:
            buckets[b] = 0
            continue
        t0, t1 = s // chunk, (e - 1) // chunk
        if t0 == t1:
            continue
        acc = edges[2 * t0 + 1]
        for t in range(t0 + 1, t1 + 1):
            acc = (acc + edges[2 * t]) % R
        assert buckets[b] is None
        buckets[b] = acc
    assert all(v is not None for v in buckets)
    n_inst = n_cols * G
    S_in, A_in, cnt, shift = buckets, None, nb, 0
    while True:
        cnt_out = (cnt + (1 << log_seg) - 1) >> log_seg
        S_out = [0] * (n_inst * cnt_out)
        A_out = [0] * (n_inst * cnt_out)
        for inst in range(n_inst):
            for s in range(cnt_out):
                lo, hi = s << log_seg, min((s << log_seg) + (1 << log_seg), cnt)
                run = tz = 0
                r = hi
                while r > lo + 1:
                    r -= 1
                    run = (run + S_in[inst * cnt + r]) % R
                    tz = (tz + run) % R
                run = (run + S_in[inst * cnt + lo]) % R
                tz = (tz << shift) % R
                if A_in is not None:
                    for r in range(lo, hi):
                        tz = (tz + A_in[inst * cnt + r]) % R
                S_out[inst * cnt_out + s] = run
                A_out[inst * cnt_out + s] = tz
        S_in, A_in, cnt, shift = S_out, A_out, cnt_out, shift + log_seg
        if cnt <= 1:
            break
    out = []
    for col in range(n_cols):
        acc = 0
        for g in reversed(range(G)):
            if g + 1 != G:
                acc = (acc << c) % R
            acc = (acc + S_in[col * G + g] + A_in[col * G + g]) % R
        out.append(acc)
    return out


# ----------------------------------------------------------------------------- batch-affine pair rounds
def _upper_bound(arr, n, v):
    """first index i in [0, n] with arr[i] > v (arr non-decreasing, arr has n+1 entries)"""
    lo, hi = 0, n + 1
    while lo < hi:
        mid = (lo + hi) // 2
        if arr[mid] > v:
            hi = mid
        else:
            lo = mid + 1
    return lo


def batch_inverse_tree(X, G):
    """mirror of binv_up / binv_top / binv_down: inverses of all X (non-zero) with ONE modular inversion"""
    levels = [list(X)]
    prefixes = []
    while len(levels[-1]) > 1:
        cur = levels[-1]
        nxt, pre = [], [0] * len(cur)
        for g in range(0, len(cur), G):
            run = 1
            for i in range(g, min(g + G, len(cur))):
                pre[i] = run
                run = run * cur[i] % R
            nxt.append(run)
        prefixes.append(pre)
        levels.append(nxt)
    inv = [pow(levels[-1][0], -1, R)]          # the only true inversion (binv_top)
    for lvl in range(len(levels) - 2, -1, -1):
        cur, pre = levels[lvl], prefixes[lvl]
        out = [0] * len(cur)
        for g in range(0, len(cur), G):
            I = inv[g // G]
            for i in range(min(g + G, len(cur)) - 1, g - 1, -1):
                out[i] = I * pre[i] % R
                I = I * cur[i] % R
        inv = out
    return inv


def pair_round(points_in, off_in, n_buckets, K, G, first_entries=None, table=None):
    """One batch-affine round.  points_in: list of ints (None = identity) or, for the first round, entries
    (pref, neg, b) into `table`.  Returns (points_out, off_out, bid_out).  'Affine add' of k1*G and k2*G is
    modelled as (k1 + k2) mod R, with the x-coordinate stand-in x(k) = k*k+7 mod R (so that x(k) == x(-k),
    as on the curve) driving the same case analysis and the same batch-inversion tree as the kernels."""
    def load(idx):
        if first_entries is not None:
            pref, neg, _ = first_entries[idx]
            v = table[pref]
            return None if v is None else ((-v) % R if neg else v)
        return points_in[idx]

    def xcoord(v):
        return (v * v + 7) % R

    def classify(a, b):
        # returns (kind, d): kinds ADD, DOUBLE, CANCEL, TAKE_A, TAKE_B
        if a is None and b is None:
            return "CANCEL", 1
        if a is None:
            return "TAKE_B", 1
        if b is None:
            return "TAKE_A", 1
        if xcoord(a) != xcoord(b):
            return "ADD", (xcoord(b) - xcoord(a)) % R
        if a == b and a != 0:
            return "DOUBLE", (2 * a) % R or 1
        return "CANCEL", 1

    cnt = [(off_in[b + 1] - off_in[b] + 1) // 2 for b in range(n_buckets)]
    off_out = [0]
    for c in cnt:
        off_out.append(off_out[-1] + c)
    S = off_out[-1]
    n_thr = (S + K - 1) // K
    P0 = [None] * S
    X1 = [1] * n_thr
    # forward
    for t in range(n_thr):
        o0 = t * K
        b = _upper_bound(off_out, n_buckets, o0) - 1
        run = 1
        for k in range(K):
            o = o0 + k
            if o >= S:
                break
            while o >= off_out[b + 1]:
                b += 1
            i = o - off_out[b]
            in0 = off_in[b] + 2 * i
            P0[o] = run
            if in0 + 1 < off_in[b + 1]:
                _, d = classify(load(in0), load(in0 + 1))
                run = run * d % R
        X1[t] = run
    I1 = batch_inverse_tree(X1, G)
    out = [None] * S
    bid = [None] * S
    # backward
    for t in range(n_thr):
        o_last = min(t * K + K, S) - 1
        b = _upper_bound(off_out, n_buckets, o_last) - 1
        I = I1[t]
        for o in range(o_last, t * K - 1, -1):
            while o < off_out[b]:
                b -= 1
            i = o - off_out[b]
            in0 = off_in[b] + 2 * i
            bid[o] = b
            if in0 + 1 < off_in[b + 1]:
                a, c = load(in0), load(in0 + 1)
                kind, d = classify(a, c)
                inv_d = I * P0[o] % R
                I = I * d % R
                assert inv_d * d % R == 1
                if kind in ("ADD", "DOUBLE"):
                    out[o] = (a + c) % R
                elif kind == "TAKE_A":
                    out[o] = a
                elif kind == "TAKE_B":
                    out[o] = c
                else:
                    out[o] = None
            else:
                out[o] = load(in0)
    return out, off_out, bid


def msm_model_affine(scalars, points, c, precomp, chunk, n_cols=1, rounds=3, K=4, G=4):
    """msm_model with `rounds` batch-affine pair rounds in front of the chunked XYZZ accumulation."""
    n = len(points)
    W = windows_for(c)
    Gw = 1 if precomp else W
    nb = 1 << (c - 1)
    half = nb - 1
    Kadd = sum(half << (j * c) for j in range(W))
    table = [(k << (j * c)) % R for j in range(W) for k in points] if precomp else list(points)
    n_buckets = n_cols * Gw * nb
    lists = [[] for _ in range(n_buckets)]
    for col in range(n_cols):
        for j in range(W):
            for i in range(n):
                t = scalars[col][i] + Kadd
                sd = ((t >> (j * c)) & ((1 << c) - 1)) - half
                if sd:
                    g = j if Gw > 1 else 0
                    lists[(col * Gw + g) * nb + abs(sd) - 1].append((i if Gw > 1 else j * n + i, sd < 0))
    entries, off = [], [0]
    for b, l in enumerate(lists):
        entries += [(p, s, b) for p, s in l]
        off.append(len(entries))
    pts, first = None, entries
    for r in range(rounds):
        pts, off, bid = pair_round(pts, off, n_buckets, K, G, first_entries=first, table=table)
        first = None
    if rounds:
        entries = [(o, False, bid[o]) for o in range(len(pts))]
        table = pts
    # chunked accumulation + finish + reduction exactly as msm_model (identity points are skipped)
    M = len(entries)
    buckets = [None] * n_buckets
    nthreads = max((M + chunk - 1) // chunk, 1)
    edges = [None] * (2 * nthreads)
    for t in range(nthreads):
        start = t * chunk
        if start >= M:
            continue
        end = min(start + chunk, M)
        prev_b = entries[start - 1][2] if start > 0 else -1
        next_b = entries[end][2] if end < M else -1
        acc, cur_b, first_run = 0, entries[start][2], True
        for e in range(start, end):
            nxt_b = entries[e + 1][2] if e + 1 < end else -1
            pref, neg, _ = entries[e]
            v = table[pref]
            if v is not None:
                acc = (acc + (-v if neg else v)) % R
            if nxt_b != cur_b:
                last_run = e + 1 == end
                sb, ca = first_run and prev_b == cur_b, last_run and next_b == cur_b
                if not sb and not ca:
                    buckets[cur_b] = acc
                elif sb:
                    edges[2 * t] = acc
                else:
                    edges[2 * t + 1] = acc
                acc, cur_b, first_run = 0, nxt_b, False
    for b in range(n_buckets):
        s, e = off[b], off[b + 1]
        if s == e:
            buckets[b] = 0
            continue
        t0, t1 = s // chunk, (e - 1) // chunk
        if t0 == t1:
            continue
        acc = edges[2 * t0 + 1]
        for t in range(t0 + 1, t1 + 1):
            acc = (acc + edges[2 * t]) % R
        buckets[b] = acc
    out = []
    for col in range(n_cols):
        acc = 0
        for g in reversed(range(Gw)):
            if g + 1 != Gw:
                acc = (acc << c) % R
            base = (col * Gw + g) * nb
            acc = (acc + sum((m + 1) * buckets[base + m] for m in range(nb))) % R
        out.append(acc)
    return out
