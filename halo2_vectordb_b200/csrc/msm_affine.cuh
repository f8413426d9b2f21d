// msm_affine.cuh -- batch-affine pair rounds in front of the chunked XYZZ bucket accumulation.
//
// Same contract as msm.cuh (halo2-axiom `best_multiexp` / `multiexp_serial`): the
// points of one bucket are summed; only the cost per group addition changes.  An affine addition
// needs one field inversion; sharing ONE inversion among all additions of a round (Montgomery's trick,
// organised as a tree over the whole GPU) leaves 6 products per addition instead of the 10 of an XYZZ
// mixed add:  1 (prefix product) + 2 (unwinding) + 3 (lambda, x3, y3).
//
// A round halves every bucket: output slot o of bucket b holds in[off_in[b] + 2i] + in[off_in[b] + 2i + 1]
// (i = o - off_out[b]; an odd leftover is copied).  Round 0 reads the sorted (point_ref, sign) entries and
// the SRS tables, later rounds read the previous round's point list.  After R rounds (R ~ log2 of the mean
// bucket load) the survivors go through msm_accumulate_kernel unchanged, which keeps every property of the
// XYZZ path (arbitrary bucket lengths, giant buckets, identity / doubling / cancelling pairs).
//   ba_count_kernel     per-bucket output count ceil(len/2)           -> msm_scan_* -> off_out
//   ba_forward_kernel   thread = K consecutive slots: denominators on the fly, exclusive prefix products
//                       P0[slot], thread total X1[t]
//   binv_up/top/down    inverses of all X1[t] with a single Fermat inversion (radix-G tree)
//   ba_backward_kernel  unwinds the prefix products (1/d_k), does the additions, writes the next list
#pragma once
#include "msm.cuh"

namespace h2v {

enum { BA_ADD = 0, BA_DOUBLE = 1, BA_CANCEL = 2, BA_TAKE_A = 3, BA_TAKE_B = 4 };

struct BaRound {
    const uint2 *entries;      // first round: sorted (point_ref | sign<<31, bucket)
    const affine *table;       // first round: SRS window tables / raw bases
    const affine *list_in;     // later rounds: previous list
    const uint32_t *off_in;    // [n_buckets + 1]
    const uint32_t *off_out;   // [n_buckets + 1]
    uint32_t n_buckets;
    uint32_t K;                // slots per thread
    uint32_t n_threads;        // threads the inversion tree covers (idle ones contribute a total of 1)
    fe *P0;                    // [slots] exclusive prefix products inside a thread's group
    fe *X1;                    // [threads] group totals
    const fe *I1;              // [threads] inverses of the group totals (backward pass)
    affine *list_out;          // [slots]
    uint2 *entries_out;        // last round only: (slot, bucket) for msm_accumulate_kernel
};

__device__ __forceinline__ affine affine_ld(const affine *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1], c = q[2], d = q[3];
    affine r;
    r.x.v[0] = a.x; r.x.v[1] = a.y; r.x.v[2] = a.z; r.x.v[3] = a.w;
    r.x.v[4] = b.x; r.x.v[5] = b.y; r.x.v[6] = b.z; r.x.v[7] = b.w;
    r.y.v[0] = c.x; r.y.v[1] = c.y; r.y.v[2] = c.z; r.y.v[3] = c.w;
    r.y.v[4] = d.x; r.y.v[5] = d.y; r.y.v[6] = d.z; r.y.v[7] = d.w;
    return r;
}
__device__ __forceinline__ void affine_st(affine *p, const affine &v) {
    fe_st(&p->x, v.x);
    fe_st(&p->y, v.y);
}
template <bool FIRST> __device__ __forceinline__ affine ba_load(const BaRound &p, uint32_t idx) {
    if (FIRST) {
        uint2 e = p.entries[idx];
        affine pt = affine_load_ro(p.table + (e.x & 0x7fffffffu));
        if (e.x & 0x80000000u) pt.y = fe_neg<Fq>(pt.y);
        return pt;
    }
    return affine_ld(p.list_in + idx);
}
// what a + b needs: the kind of the pair and, for ADD / DOUBLE, the denominator of lambda
__device__ __forceinline__ int ba_classify(const affine &a, const affine &b, fe &d) {
    const bool ia = affine_is_identity(a), ib = affine_is_identity(b);
    if (ia || ib) return ia ? (ib ? BA_CANCEL : BA_TAKE_B) : BA_TAKE_A;
    d = fe_sub<Fq>(b.x, a.x);
    if (!fe_is_zero(d)) return BA_ADD;
    if (fe_eq(a.y, b.y) && !fe_is_zero(a.y)) {
        d = fe_dbl<Fq>(a.y);
        return BA_DOUBLE;
    }
    return BA_CANCEL;
}
// first index i in [0, n] with arr[i] > v   (arr non-decreasing with n + 1 entries)
__device__ __forceinline__ uint32_t ba_upper_bound(const uint32_t *__restrict__ arr, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n + 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (arr[mid] > v) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) ba_count_kernel(const uint32_t *__restrict__ off_in, uint32_t *__restrict__ cnt, uint32_t n_buckets) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_buckets) cnt[b] = (off_in[b + 1] - off_in[b] + 1) >> 1;
}

template <bool FIRST> __global__ void __launch_bounds__(128) ba_forward_kernel(BaRound p) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t S = p.off_out[p.n_buckets];
    const uint64_t o0_64 = (uint64_t)t * p.K;
    if (o0_64 >= S) {
        if (t < p.n_threads) fe_st(p.X1 + t, fe_one<Fq>());
        return;
    }
    const uint32_t o0 = (uint32_t)o0_64;
    uint32_t b = ba_upper_bound(p.off_out, p.n_buckets, o0) - 1;
    uint32_t ob = p.off_out[b], ob_end = p.off_out[b + 1], ib = p.off_in[b], ib_end = p.off_in[b + 1];
    fe run = fe_one<Fq>();
    for (uint32_t k = 0; k < p.K; ++k) {
        const uint32_t o = o0 + k;
        if (o >= S) break;
        while (o >= ob_end) {
            ++b;
            ob = ob_end;
            ob_end = p.off_out[b + 1];
            ib = p.off_in[b];
            ib_end = p.off_in[b + 1];
        }
        const uint32_t in0 = ib + 2 * (o - ob);
        fe_st(p.P0 + o, run);
        if (in0 + 1 < ib_end) {
            affine a = ba_load<FIRST>(p, in0), c = ba_load<FIRST>(p, in0 + 1);
            fe d;
            if (ba_classify(a, c, d) <= BA_DOUBLE) run = fe_mul<Fq>(run, d);
        }
    }
    fe_st(p.X1 + t, run);
}

template <bool FIRST> __global__ void __launch_bounds__(128) ba_backward_kernel(BaRound p) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t S = p.off_out[p.n_buckets];
    const uint64_t o0_64 = (uint64_t)t * p.K;
    if (o0_64 >= S) return;
    const uint32_t o0 = (uint32_t)o0_64;
    const uint32_t o_last = min(o0 + p.K, S) - 1;
    uint32_t b = ba_upper_bound(p.off_out, p.n_buckets, o_last) - 1;
    uint32_t ob = p.off_out[b], ib = p.off_in[b], ib_end = p.off_in[b + 1];
    fe I = fe_ld(p.I1 + t);
    for (uint32_t o = o_last + 1; o-- > o0;) {
        while (o < ob) {
            --b;
            ob = p.off_out[b];
            ib = p.off_in[b];
            ib_end = p.off_in[b + 1];
        }
        const uint32_t in0 = ib + 2 * (o - ob);
        affine a = ba_load<FIRST>(p, in0), r;
        if (in0 + 1 < ib_end) {
            affine c = ba_load<FIRST>(p, in0 + 1);
            fe d;
            const int kind = ba_classify(a, c, d);
            if (kind <= BA_DOUBLE) {
                fe inv_d = fe_mul<Fq>(I, fe_ld(p.P0 + o));
                I = fe_mul<Fq>(I, d);
                fe num;
                if (kind == BA_ADD) {
                    num = fe_sub<Fq>(c.y, a.y);
                } else {
                    fe xx = fe_sqr<Fq>(a.x);
                    num = fe_add<Fq>(fe_dbl<Fq>(xx), xx);
                }
                fe lam = fe_mul<Fq>(num, inv_d);
                r.x = fe_sub<Fq>(fe_sub<Fq>(fe_sqr<Fq>(lam), a.x), c.x);
                r.y = fe_sub<Fq>(fe_mul<Fq>(lam, fe_sub<Fq>(a.x, r.x)), a.y);
            } else if (kind == BA_TAKE_A) {
                r = a;
            } else if (kind == BA_TAKE_B) {
                r = c;
            } else {
                r.x = fe_zero();
                r.y = fe_zero();
            }
        } else {
            r = a;
        }
        affine_st(p.list_out + o, r);
        if (p.entries_out) p.entries_out[o] = make_uint2(o, b);
    }
}

// ------------------------------------------------------------------ inversion tree (all elements non-zero)
// P[i] = product of X[j] for j < i inside i's group of G;  Xn[g] = product of group g
template <class F> __global__ void __launch_bounds__(128) binv_up_kernel(const fe *__restrict__ X, fe *__restrict__ P, fe *__restrict__ Xn, uint32_t n, uint32_t G) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = (uint64_t)g * G;
    if (lo >= n) return;
    uint32_t hi = (uint32_t)min((uint64_t)n, lo + G);
    fe run = fe_one<F>();
    for (uint32_t i = (uint32_t)lo; i < hi; ++i) {
        fe_st(P + i, run);
        run = fe_mul<F>(run, fe_ld(X + i));
    }
    fe_st(Xn + g, run);
}
template <class F> __global__ void binv_top_kernel(const fe *__restrict__ X, fe *__restrict__ I) {
    if (blockIdx.x == 0 && threadIdx.x == 0) fe_st(I, fe_inv_fast<F>(fe_ld(X)));
}
// I[i] = 1 / X[i] from In[g] = 1 / (product of group g)
template <class F> __global__ void __launch_bounds__(128) binv_down_kernel(const fe *__restrict__ X, const fe *__restrict__ P, const fe *__restrict__ In,
                                                        fe *__restrict__ I, uint32_t n, uint32_t G) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = (uint64_t)g * G;
    if (lo >= n) return;
    uint32_t hi = (uint32_t)min((uint64_t)n, lo + G);
    fe inv = fe_ld(In + g);
    for (uint32_t i = hi; i-- > (uint32_t)lo;) {
        fe_st(I + i, fe_mul<F>(inv, fe_ld(P + i)));
        inv = fe_mul<F>(inv, fe_ld(X + i));
    }
}

}  // namespace h2v
