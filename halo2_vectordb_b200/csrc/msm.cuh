// msm.cuh -- BN254 G1 Pippenger multi-scalar multiplication kernels.
//
// Replaces halo2-axiom arithmetic.rs `best_multiexp` / `multiexp_serial` and, through them,
// poly/kzg/commitment.rs `ParamsKZG::{commit, commit_lagrange}` (SURVEY.md 8(a) rows a1-a4,
// App. A.3, A.6; reached from /root/reference/src/scaffold/mod.rs:273,296).  The result is the
// same group element sum_i s_i * B_i; it is compared with the reference after affine
// normalisation, which does not depend on window size, digit signs or summation order.
//
// Pipeline (one launch each, batched over `n_cols` scalar columns that share the bases):
//   msm_digits     scalar: Montgomery -> canonical, + K, split into W signed c-bit digits
//                  d in [-(2^(c-1)-1), 2^(c-1)]; per-bucket histogram
//   msm_scan_*     exclusive prefix sum of the histogram -> bucket offsets (tiles / top / apply)
//   msm_scatter    counting-sort the (point, sign) pairs by bucket
//   msm_accumulate fixed-size chunks of the sorted list, one thread each, XYZZ mixed adds;
//                  load balance is independent of the scalar distribution
//   msm_finish     merge the partial sums of buckets that straddle chunk boundaries (thread per bucket;
//                  a warp per bucket for the few that span many chunks)
//   msm_reduce     sum_m (m+1) * bucket[m] by a radix-8..32 tree of running sums
//   msm_final      fold window groups (Horner, c doublings each), normalise to affine
// Two layouts of the same kernels:
//   precomputed (SRS handles): tables T_j[i] = 2^(j c) B_i are built once per SRS, every digit
//       of every window lands in ONE bucket set per column (G = 1), so there is one reduction per
//       column instead of one per window and no doublings at the end;
//   raw (`best_multiexp` shape, arbitrary bases): G = W bucket sets, folded by Horner.
#pragma once
#include "ec.cuh"

namespace h2v {

typedef FrP Fr;

struct MsmShape {
    uint32_t n;        // points per column
    uint32_t n_cols;   // columns in this batch
    uint32_t c;        // window bits
    uint32_t W;        // windows
    uint32_t G;        // bucket groups per column: 1 (precomputed tables) or W (raw)
    uint32_t nb;       // buckets per group = 2^(c-1)
    uint32_t chunk;    // sorted entries per accumulate thread
    uint32_t pstride;  // table stride between window levels (precomputed layout), >= n
    uint32_t kadd[9];  // K = sum_j (2^(c-1)-1) 2^(jc), added before digit extraction (W*c <= 288 bits)
};

#define H2V_KEY_INVALID 0xffffffffu

__device__ __forceinline__ affine affine_load_ro(const affine *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    affine r;
    r.x.v[0] = a.x; r.x.v[1] = a.y; r.x.v[2] = a.z; r.x.v[3] = a.w;
    r.x.v[4] = b.x; r.x.v[5] = b.y; r.x.v[6] = b.z; r.x.v[7] = b.w;
    r.y.v[0] = c.x; r.y.v[1] = c.y; r.y.v[2] = c.z; r.y.v[3] = c.w;
    r.y.v[4] = d.x; r.y.v[5] = d.y; r.y.v[6] = d.z; r.y.v[7] = d.w;
    return r;
}
__device__ __forceinline__ void fe_st(fe *p, const fe &x) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ fe fe_ld(const fe *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void xyzz_st(xyzz *p, const xyzz &v) {
    fe_st(&p->x, v.x); fe_st(&p->y, v.y); fe_st(&p->zz, v.zz); fe_st(&p->zzz, v.zzz);
}
__device__ __forceinline__ xyzz xyzz_ld(const xyzz *p) {
    xyzz r;
    r.x = fe_ld(&p->x); r.y = fe_ld(&p->y); r.zz = fe_ld(&p->zz); r.zzz = fe_ld(&p->zzz);
    return r;
}

// ------------------------------------------------------------------ digits
// The W signed c-bit digits of s + K (K = sum_j (2^(c-1) - 1) 2^(jc)): d_j = digit_j(s + K) - (2^(c-1) - 1) lies in
// [-(2^(c-1) - 1), 2^(c-1)] and sum_j d_j 2^(jc) = s.  Both passes over the scalars (histogram, scatter) recompute them
// from the scalar -- one Montgomery product and W shifts -- instead of staging 4 W bytes per scalar in HBM.
struct MsmDigits {
    uint32_t t[9];
    __device__ __forceinline__ void load(const fe *p, const MsmShape &sh) {
        fe s = fe_from_mont<Fr>(fe_ld(p));
        uint64_t cy = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cy += (uint64_t)s.v[k] + sh.kadd[k];
            t[k] = (uint32_t)cy;
            cy >>= 32;
        }
        t[8] = (uint32_t)cy + sh.kadd[8];
    }
    // digit j: false for a zero digit, else magnitude - 1 and sign
    __device__ __forceinline__ bool get(uint32_t j, const MsmShape &sh, uint32_t &mag, uint32_t &neg) const {
        const uint32_t bit = j * sh.c, w = bit >> 5, sft = bit & 31;
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {       // t stays in registers: no dynamic indexing
            lo = (w == (uint32_t)k) ? t[k] : lo;
            hi = (w + 1 == (uint32_t)k) ? t[k] : hi;
        }
        const uint32_t d = (uint32_t)((((uint64_t)hi << 32) | lo) >> sft) & ((1u << sh.c) - 1);
        const int32_t sd = (int32_t)d - (int32_t)((1u << (sh.c - 1)) - 1);
        if (sd == 0) return false;
        neg = sd < 0 ? 0x80000000u : 0u;
        mag = (uint32_t)(sd < 0 ? -sd : sd) - 1;
        return true;
    }
};
// Lanes of a warp that hit the same counter are served by ONE atomic (skewed witness columns put a third of a column
// into one bucket: scalar 1, the shared high digits of r - small).  Returns this lane's slot: base + rank in its group.
__device__ __forceinline__ uint32_t warp_agg_add(uint32_t *counter_base, uint32_t idx, bool valid) {
    const uint32_t lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? idx : 0xffffffffu);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (valid && (int)lane == leader) base = atomicAdd(counter_base + idx, (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1));
}

// the same without a result: the leader's atomic is a fire-and-forget reduction (no round trip to wait for)
__device__ __forceinline__ void warp_agg_inc(uint32_t *counter_base, uint32_t idx, bool valid) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t key = valid ? idx : 0xffffffffu - lane;
    // cheap pre-test: uniform scalars practically never put two neighbouring lanes into one bucket; skewed columns do
    // (a third of a witness column is the scalar 1), and only then is the full match worth its cost
    const uint32_t n1 = __shfl_xor_sync(0xffffffffu, key, 1), n2 = __shfl_xor_sync(0xffffffffu, key, 2);      // all lanes, both
    const bool dup = n1 == key || n2 == key;
    if (!__any_sync(0xffffffffu, dup)) {
        if (valid) atomicAdd(counter_base + idx, 1u);
        return;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (valid && (int)lane == __ffs(peers) - 1) atomicAdd(counter_base + idx, (uint32_t)__popc(peers));
}

// ------------------------------------------------------------------ pass 1: histogram
// The window size is a template parameter (host dispatch over 2..24): after unrolling over the windows every limb
// index and shift is a constant, so a digit costs a funnel shift and a compare instead of a 9-way select.
template <int C> struct MsmDigitsC {
    static constexpr uint32_t W = (255 + C - 1) / C;
    uint32_t t[9];
    __device__ __forceinline__ void load(const fe *p, const MsmShape &sh) {
        fe s = fe_from_mont<Fr>(fe_ld(p));
        uint64_t cy = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cy += (uint64_t)s.v[k] + sh.kadd[k];
            t[k] = (uint32_t)cy;
            cy >>= 32;
        }
        t[8] = (uint32_t)cy + sh.kadd[8];
    }
    template <uint32_t J> __device__ __forceinline__ bool get(uint32_t &mag, uint32_t &neg) const {
        constexpr uint32_t bit = J * C, w = bit >> 5, sft = bit & 31;
        const uint32_t lo = w < 9 ? t[w < 9 ? w : 0] : 0u, hi = (w + 1) < 9 ? t[(w + 1) < 9 ? w + 1 : 0] : 0u;
        const uint32_t d = (sft ? __funnelshift_r(lo, hi, sft) : lo) & ((1u << C) - 1);
        const int32_t sd = (int32_t)d - (int32_t)((1u << (C - 1)) - 1);
        if (sd == 0) return false;
        neg = sd < 0 ? 0x80000000u : 0u;
        mag = (uint32_t)(sd < 0 ? -sd : sd) - 1;
        return true;
    }
};
template <int C, uint32_t J> struct MsmCountStep {
    static __device__ __forceinline__ void run(const MsmDigitsC<C> &dg, bool in, uint32_t col, uint32_t *counts, const MsmShape &sh) {
        uint32_t mag = 0, neg = 0;
        const bool ok = in && dg.template get<J>(mag, neg);
        const uint32_t g = sh.G > 1 ? J : 0;
        warp_agg_inc(counts, (col * sh.G + g) * sh.nb + mag, ok);
        MsmCountStep<C, J + 1>::run(dg, in, col, counts, sh);
    }
};
template <int C> struct MsmCountStep<C, MsmDigitsC<C>::W> {
    static __device__ __forceinline__ void run(const MsmDigitsC<C> &, bool, uint32_t, uint32_t *, const MsmShape &) {}
};
template <int C>
__global__ void __launch_bounds__(256) msm_count_kernel(const fe *__restrict__ scalars, size_t col_stride, uint32_t *__restrict__ counts,
                                                        MsmShape sh) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, col = blockIdx.y;
    const bool in = i < sh.n;
    MsmDigitsC<C> dg;
    dg.load(scalars + (size_t)col * col_stride + (in ? i : 0), sh);
    MsmCountStep<C, 0>::run(dg, in, col, counts, sh);
}
// estimate of the number of non-zero digits under two window sizes, from a strided sample of the batch (window choice)
__global__ void __launch_bounds__(256) msm_density_kernel(const fe *__restrict__ scalars, size_t col_stride, uint32_t n_cols, uint32_t n,
                                                          uint32_t samples, MsmShape sh0, MsmShape sh1, uint32_t *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c0 = 0, c1 = 0;
    if (t < samples) {
        const uint64_t tot = (uint64_t)n_cols * n;
        const uint64_t pos = ((uint64_t)t * 0x9E3779B97F4A7C15ull >> 11) % tot;       // spread over columns and rows
        const fe *p = scalars + (size_t)(pos / n) * col_stride + (pos % n);
        MsmDigits dg;
        uint32_t mag, neg;
        dg.load(p, sh0);
        for (uint32_t j = 0; j < sh0.W; ++j) c0 += dg.get(j, sh0, mag, neg) ? 1u : 0u;
        dg.load(p, sh1);
        for (uint32_t j = 0; j < sh1.W; ++j) c1 += dg.get(j, sh1, mag, neg) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        c0 += __shfl_down_sync(0xffffffffu, c0, o);
        c1 += __shfl_down_sync(0xffffffffu, c1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, c0);
        atomicAdd(out + 1, c1);
    }
}

// ------------------------------------------------------------------ exclusive scan (three launches)
// tile = 2048 counters per CTA (8 per thread).  (1) per-tile totals, (2) one CTA scans the totals,
// (3) every tile rescans its counters on top of its base: offsets[0..len], cursor[0..len), offsets[len] = total
#define H2V_SCAN_TILE 2048
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}
// exclusive prefix of `v` across a 256-thread CTA; *total receives the CTA sum
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t *total) {
    __shared__ uint32_t wsum[8];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v);
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t x = wsum[w];
        if (w < (int)wid) base += x;
        tot += x;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}
__global__ void __launch_bounds__(256) msm_scan_tiles_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ tile_sums,
                                                             uint32_t len) {
    const uint32_t base = blockIdx.x * H2V_SCAN_TILE + threadIdx.x * 8;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (base + k < len) s += counts[base + k];
    uint32_t tot;
    block_excl_scan_256(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
// in-place exclusive scan of tile_sums[0..ntiles) by one CTA; offsets_total[0] = grand total
__global__ void __launch_bounds__(256) msm_scan_top_kernel(uint32_t *__restrict__ tile_sums, uint32_t ntiles,
                                                           uint32_t *__restrict__ offsets_total) {
    const uint32_t per = (ntiles + 255) / 256;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, ntiles);
    uint32_t s = 0;
    for (uint32_t k = lo; k < hi; ++k) s += tile_sums[k];
    uint32_t tot;
    uint32_t run = block_excl_scan_256(s, &tot);
    for (uint32_t k = lo; k < hi; ++k) {
        uint32_t v = tile_sums[k];
        tile_sums[k] = run;
        run += v;
    }
    if (threadIdx.x == 0) offsets_total[0] = tot;
}
__global__ void __launch_bounds__(256) msm_scan_apply_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ tile_offs,
                                                             uint32_t *__restrict__ offsets, uint32_t *__restrict__ cursor, uint32_t len) {
    const uint32_t base = blockIdx.x * H2V_SCAN_TILE + threadIdx.x * 8;
    uint32_t c[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        c[k] = (base + k < len) ? counts[base + k] : 0u;
        s += c[k];
    }
    uint32_t tot;
    uint32_t run = block_excl_scan_256(s, &tot) + tile_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (base + k < len) {
            offsets[base + k] = run;
            cursor[base + k] = run;
        }
        run += c[k];
    }
}

// ------------------------------------------------------------------ pass 2: scatter
// entries[pos] = (point_ref | sign<<31, global bucket), counting sort by bucket.
// Only digits with magnitude in [mag_lo, mag_hi) are placed: for a very large MSM the host sweeps the bucket
// space in slices so that the randomly written part of `entries` stays L2-resident.
template <int C, uint32_t J> struct MsmScatterStep {
    static __device__ __forceinline__ void run(const MsmDigitsC<C> &dg, bool in, uint32_t i, uint32_t col, uint32_t *cursor, uint2 *entries,
                                               const MsmShape &sh, uint32_t mag_lo, uint32_t mag_hi) {
        uint32_t mag = 0, neg = 0;
        bool ok = in && dg.template get<J>(mag, neg);
        ok = ok && mag >= mag_lo && mag < mag_hi;
        const uint32_t g = sh.G > 1 ? J : 0;
        const uint32_t b = (col * sh.G + g) * sh.nb + mag;
        const uint32_t pos = warp_agg_add(cursor, b, ok);
        if (ok) entries[pos] = make_uint2((sh.G > 1 ? i : J * sh.pstride + i) | neg, b);
        MsmScatterStep<C, J + 1>::run(dg, in, i, col, cursor, entries, sh, mag_lo, mag_hi);
    }
};
template <int C> struct MsmScatterStep<C, MsmDigitsC<C>::W> {
    static __device__ __forceinline__ void run(const MsmDigitsC<C> &, bool, uint32_t, uint32_t, uint32_t *, uint2 *, const MsmShape &, uint32_t,
                                               uint32_t) {}
};
template <int C>
__global__ void __launch_bounds__(256) msm_scatter_kernel(const fe *__restrict__ scalars, size_t col_stride, uint32_t *__restrict__ cursor,
                                                          uint2 *__restrict__ entries, MsmShape sh, uint32_t mag_lo, uint32_t mag_hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, col = blockIdx.y;
    const bool in = i < sh.n;
    MsmDigitsC<C> dg;
    dg.load(scalars + (size_t)col * col_stride + (in ? i : 0), sh);
    MsmScatterStep<C, 0>::run(dg, in, i, col, cursor, entries, sh, mag_lo, mag_hi);
}

// ------------------------------------------------------------------ chunked accumulation
// Thread t sums entries [t*chunk, (t+1)*chunk).  A run (maximal same-bucket subsequence) whose
// bucket lies wholly inside the chunk goes straight to buckets[b]; a run of a bucket that began in
// an earlier chunk goes to edges[2t] ("head"), one that continues into a later chunk to
// edges[2t+1] ("tail").  msm_finish adds tail(t0) + head(t0+1..t1) for straddling buckets.
__global__ void __launch_bounds__(128, 4) msm_accumulate_kernel(const uint2 *__restrict__ entries,
                                                             const uint32_t *__restrict__ offsets, uint32_t n_buckets,
                                                             const affine *__restrict__ points, xyzz *__restrict__ buckets,
                                                             xyzz *__restrict__ edges, uint32_t chunk,
                                                             uint2 *__restrict__ strad_list, uint32_t *__restrict__ strad_count) {
    const uint32_t M = offsets[n_buckets];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t start64 = (uint64_t)t * chunk;
    bool opens = false;          // this chunk holds the first part of a bucket that continues into later chunks
    uint32_t open_b = 0;
    if (start64 < M) {
        const uint32_t start = (uint32_t)start64;
        const uint32_t end = min(start + chunk, M);
        const uint32_t prev_b = start > 0 ? entries[start - 1].y : 0xffffffffu;
        const uint32_t next_b = end < M ? entries[end].y : 0xffffffffu;
        xyzz acc = xyzz_identity();
        uint2 ent = entries[start];
        uint32_t cur_b = ent.y;
        bool first_run = true;
        for (uint32_t e = start; e < end; ++e) {
            uint2 nxt = (e + 1 < end) ? entries[e + 1] : make_uint2(0u, 0xffffffffu);
            affine pt = affine_load_ro(points + (ent.x & 0x7fffffffu));
            if (ent.x & 0x80000000u) pt.y = fe_neg<Fq>(pt.y);
            xyzz_add_mixed_lazy(acc, pt);
            if (nxt.y != cur_b) {
                // run ends here (bucket change or end of chunk)
                xyzz_canon(acc);
                const bool last_run = (e + 1 == end);
                const bool starts_before = first_run && (prev_b == cur_b);
                const bool continues_after = last_run && (next_b == cur_b);
                if (!starts_before && !continues_after) xyzz_st(buckets + cur_b, acc);
                else if (starts_before) xyzz_st(edges + 2 * (size_t)t, acc);
                else xyzz_st(edges + 2 * (size_t)t + 1, acc);
                if (continues_after && !starts_before) {
                    opens = true;
                    open_b = cur_b;
                }
                acc = xyzz_identity();
                cur_b = nxt.y;
                first_run = false;
            }
            ent = nxt;
        }
    }
    // the straddling buckets, one list entry each (first chunk, bucket): msm_finish works on this list only
    const unsigned m = __ballot_sync(0xffffffffu, opens);
    if (m) {
        const uint32_t lane = threadIdx.x & 31;
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(strad_count, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (opens) strad_list[base + __popc(m & ((1u << lane) - 1))] = make_uint2(t, open_b);
    }
}

// One thread per straddling bucket (the list msm_accumulate wrote): tail(t0) + heads(t0+1..t1).  Buckets that span
// more than H2V_LONG_SPAN chunks (skewed witness columns put a large share of a column into a few buckets:
// scalar 1, the shared high digits of r - small) are queued for msm_finish_long_kernel instead.  Empty buckets are
// never touched: the reduction skips them by their offsets.
#define H2V_LONG_SPAN 16
__global__ void __launch_bounds__(128) msm_finish_kernel(const uint32_t *__restrict__ offsets, const uint2 *__restrict__ strad_list,
                                                         const uint32_t *__restrict__ strad_count, const xyzz *__restrict__ edges,
                                                         xyzz *__restrict__ buckets, uint32_t chunk, uint32_t *__restrict__ long_list,
                                                         uint32_t *__restrict__ long_count) {
    const uint32_t count = *strad_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint2 it = strad_list[i];
        const uint32_t t0 = it.x, b = it.y;
        const uint32_t t1 = (offsets[b + 1] - 1) / chunk;
        if (t1 - t0 > H2V_LONG_SPAN) {
            long_list[atomicAdd(long_count, 1u)] = b;
            continue;
        }
        xyzz acc = xyzz_ld(edges + 2 * (size_t)t0 + 1);
        for (uint32_t t = t0 + 1; t <= t1; ++t) {
            xyzz h = xyzz_ld(edges + 2 * (size_t)t);
            xyzz_add(acc, h);
        }
        xyzz_st(buckets + b, acc);
    }
}
__device__ __forceinline__ xyzz xyzz_shfl_down(const xyzz &v, int off) {
    xyzz r;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        r.x.v[k] = __shfl_down_sync(0xffffffffu, v.x.v[k], off);
        r.y.v[k] = __shfl_down_sync(0xffffffffu, v.y.v[k], off);
        r.zz.v[k] = __shfl_down_sync(0xffffffffu, v.zz.v[k], off);
        r.zzz.v[k] = __shfl_down_sync(0xffffffffu, v.zzz.v[k], off);
    }
    return r;
}
__device__ __forceinline__ xyzz xyzz_warp_sum(xyzz v) {      // total in lane 0
#pragma unroll 1
    for (int off = 16; off >= 1; off >>= 1) {
        xyzz o = xyzz_shfl_down(v, off);
        xyzz_add(v, o);
    }
    return v;
}
// sum of `v` over the threads of a CTA of up to 8 warps (shuffle tree inside each warp, the warp partials
// through shared memory); the total is returned by thread 0 only
__device__ __forceinline__ xyzz xyzz_block_sum_256(xyzz v, xyzz *sm8) {
#pragma unroll 1
    for (int off = 16; off >= 1; off >>= 1) {
        xyzz o = xyzz_shfl_down(v, off);
        xyzz_add(v, o);
    }
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();                       // sm8 may still be read from a previous call
    if (lane == 0) sm8[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < (blockDim.x >> 5) ? sm8[lane] : xyzz_identity();
#pragma unroll 1
        for (int off = 4; off >= 1; off >>= 1) {
            xyzz o = xyzz_shfl_down(v, off);
            xyzz_add(v, o);
        }
    }
    return v;
}
// Queued long buckets.  Few of them (a single witness column): one CTA per bucket, 256 threads stride over the
// chunk heads and a tree folds the partials -- shortest critical path.  Many of them (a batch of columns):
// one warp per bucket -- same work with fewer redundant tree steps.
__global__ void __launch_bounds__(256) msm_finish_long_kernel(const uint32_t *__restrict__ offsets, const xyzz *__restrict__ edges,
                                                              xyzz *__restrict__ buckets, uint32_t chunk,
                                                              const uint32_t *__restrict__ long_list,
                                                              const uint32_t *__restrict__ long_count) {
    __shared__ xyzz sm8[8];
    const uint32_t count = *long_count;
    if (count <= 2 * gridDim.x) {
        for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
            const uint32_t b = long_list[i];
            const uint32_t t0 = offsets[b] / chunk, t1 = (offsets[b + 1] - 1) / chunk;
            xyzz acc = xyzz_identity();
            if (threadIdx.x == 0) acc = xyzz_ld(edges + 2 * (size_t)t0 + 1);
            for (uint32_t t = t0 + 1 + threadIdx.x; t <= t1; t += 256) {
                xyzz h = xyzz_ld(edges + 2 * (size_t)t);
                xyzz_add(acc, h);
            }
            acc = xyzz_block_sum_256(acc, sm8);
            if (threadIdx.x == 0) xyzz_st(buckets + b, acc);
        }
    } else {
        const uint32_t lane = threadIdx.x & 31;
        const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
        for (uint32_t i = warp; i < count; i += n_warps) {
            const uint32_t b = long_list[i];
            const uint32_t t0 = offsets[b] / chunk, t1 = (offsets[b + 1] - 1) / chunk;
            xyzz acc = xyzz_identity();
            if (lane == 0) acc = xyzz_ld(edges + 2 * (size_t)t0 + 1);
            for (uint32_t t = t0 + 1 + lane; t <= t1; t += 32) {
                xyzz h = xyzz_ld(edges + 2 * (size_t)t);
                xyzz_add(acc, h);
            }
            acc = xyzz_warp_sum(acc);
            if (lane == 0) xyzz_st(buckets + b, acc);
        }
    }
}

// ------------------------------------------------------------------ bucket reduction tree
// For each of `n_inst` instances: S_in[inst][0..cnt_in) (and optional A_in) -> cnt_out = ceil(cnt_in / seg),
// seg = 2^log_seg (8..32, chosen per level so that every level still fills the GPU):
//   S_out[s] = sum_{r} S_in[seg s + r]
//   A_out[s] = sum_{r} A_in[seg s + r] + 2^shift * sum_r r * S_in[seg s + r]     (shift = sum of earlier log_seg)
// Iterating until cnt == 1 gives  A = sum_m m * S0[m],  S = sum_m S0[m];  the group result is A + S
// (bucket m holds the points of digit magnitude m + 1).
// `occ` (first level only): the bucket offsets -- element r of instance `inst` is the bucket inst * cnt_in + r, empty
// (never written) when occ[b] == occ[b + 1].
__global__ void __launch_bounds__(128, 3) msm_reduce_kernel(const xyzz *__restrict__ S_in, const xyzz *__restrict__ A_in,
                                                            xyzz *__restrict__ S_out, xyzz *__restrict__ A_out,
                                                            uint32_t cnt_in, uint32_t cnt_out, uint32_t n_inst, uint32_t shift,
                                                            uint32_t log_seg, const uint32_t *__restrict__ occ) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_inst * cnt_out) return;
    uint32_t inst = gid / cnt_out, s = gid % cnt_out;
    uint32_t lo = s << log_seg, hi = min(lo + (1u << log_seg), cnt_in);
    const xyzz *Sin = S_in + (size_t)inst * cnt_in;
    const uint32_t *oc = occ ? occ + (size_t)inst * cnt_in : nullptr;
    xyzz run = xyzz_identity(), tz = xyzz_identity();
    for (uint32_t r = hi; r-- > lo + 1;) {
        if (!oc || oc[r] != oc[r + 1]) {
            xyzz v = xyzz_ld(Sin + r);
            xyzz_add(run, v);
        }
        xyzz_add(tz, run);
    }
    if (!oc || oc[lo] != oc[lo + 1]) {
        xyzz v = xyzz_ld(Sin + lo);
        xyzz_add(run, v);
    }
    for (uint32_t k = 0; k < shift; ++k) tz = xyzz_double(tz);
    if (A_in) {
        const xyzz *Ain = A_in + (size_t)inst * cnt_in;
        for (uint32_t r = lo; r < hi; ++r) {
            xyzz v = xyzz_ld(Ain + r);
            xyzz_add(tz, v);
        }
    }
    xyzz_st(S_out + (size_t)inst * cnt_out + s, run);
    xyzz_st(A_out + (size_t)inst * cnt_out + s, tz);
}

// Tail of the reduction, once <= H2V_TREE_MAX partial pairs (S_s, A_s) per instance are left: the serial radix levels
// are latency-bound there (a lone warp needs ~4 us per full addition), so ONE CTA per instance finishes the job with
// log-depth sums whose active lanes stay packed:
//   sum_s s * S_s = sum_b 2^b Z_b,  Z_b = sum over { s : bit b of s set } of S_s.
// With T the pairwise block sums of S (level l: blocks of 2^l), Z_l is the sum of the ODD blocks of level l.
// Phase A builds the T tree (log2 cnt steps) and keeps every level's odd blocks; phase B sums all levels' odd blocks and
// the A values at the same time (log2 cnt - 1 halving steps over packed lanes); phase C scales Z_l by 2^l and adds.
// cnt - 1 additions for T and cnt - 1 for the Z's in total -- the same work as the serial running sums.
// Shared memory: T (cnt / 2) + odd blocks of all levels (cnt) + A (cnt / 2) XYZZ values, 128 B each.
#define H2V_TREE_MAX 512
__global__ void __launch_bounds__(256) msm_reduce_tree_kernel(const xyzz *__restrict__ S_in, const xyzz *__restrict__ A_in, uint32_t cnt,
                                                              uint32_t shift, xyzz *__restrict__ S_out, xyzz *__restrict__ A_out) {
    extern __shared__ xyzz tsm[];
    const uint32_t inst = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const xyzz *S = S_in + (size_t)inst * cnt;
    const uint32_t half0 = cnt >> 1;              // cnt is a power of two >= 2
    uint32_t lg = 0;
    while ((1u << lg) < cnt) ++lg;                // levels 0 .. lg - 1; level l has cnt >> (l + 1) odd blocks
    xyzz *T = tsm, *O = tsm + half0, *AA = tsm + half0 + cnt;
    // odd blocks of level l start at O + (cnt - (cnt >> l))
    const bool haveA = A_in != nullptr;
    // ---- phase A, level 0 (from global memory): T[i] = S[2i] + S[2i+1], O_0[i] = S[2i+1], AA[i] = A[2i] + A[2i+1]
    for (uint32_t i = tid; i < half0; i += nth) {
        xyzz a = xyzz_ld(S + 2 * i), b = xyzz_ld(S + 2 * i + 1);
        O[i] = b;
        xyzz_add(a, b);
        T[i] = a;
        if (haveA) {
            const xyzz *A = A_in + (size_t)inst * cnt;
            xyzz c = xyzz_ld(A + 2 * i), d = xyzz_ld(A + 2 * i + 1);
            xyzz_add(c, d);
            AA[i] = c;
        }
    }
    __syncthreads();
    for (uint32_t l = 1; l < lg; ++l) {
        const uint32_t h = cnt >> (l + 1);        // h <= cnt / 4 <= blockDim.x (the host launches at least cnt / 4 threads)
        xyzz a = xyzz_identity(), b = xyzz_identity();
        if (tid < h) {
            a = T[2 * tid];
            b = T[2 * tid + 1];
        }
        __syncthreads();                          // the pairwise sums overwrite T in place
        if (tid < h) {
            O[cnt - (cnt >> l) + tid] = b;
            xyzz_add(a, b);
            T[tid] = a;
        }
        __syncthreads();
    }
    // ---- phase B: every level's odd blocks (and AA) are halved together until one value per level is left
    for (uint32_t step = 1; (half0 >> step) >= 1; ++step) {
        // work items of this step: for level l, the first (cnt >> (l + 1 + step)) elements; AA likewise (as level 0)
        for (uint32_t w = tid;; w += nth) {
            uint32_t rem = w;
            xyzz *seg = nullptr;
            uint32_t hcount = 0;
            for (uint32_t l = 0; l < lg; ++l) {
                const uint32_t hc = cnt >> (l + 1 + step);
                if (hc == 0) break;
                if (rem < hc) {
                    seg = O + (cnt - (cnt >> l));
                    hcount = hc;
                    break;
                }
                rem -= hc;
            }
            if (!seg && haveA) {
                const uint32_t hc = half0 >> step;
                if (rem < hc) {
                    seg = AA;
                    hcount = hc;
                }
            }
            if (!seg) break;
            xyzz a = seg[rem];
            xyzz_add(a, seg[rem + hcount]);
            seg[rem] = a;
        }
        __syncthreads();
    }
    // ---- phase C: lane l scales Z_l by 2^l (lockstep doublings), a shuffle tree adds them
    if (tid < 32) {
        xyzz acc = tid < lg ? O[cnt - (cnt >> tid)] : xyzz_identity();
        for (uint32_t k = 0; k + 1 < lg; ++k)
            if (k < tid && tid < lg) acc = xyzz_double(acc);
        acc = xyzz_warp_sum(acc);
        if (tid == 0) {
            for (uint32_t k = 0; k < shift; ++k) acc = xyzz_double(acc);
            if (haveA) xyzz_add(acc, AA[0]);
            xyzz_st(S_out + inst, T[0]);
            xyzz_st(A_out + inst, acc);
        }
    }
}

// one warp per column (lane 0 works: the data-dependent inversion would diverge across columns): group results
// R_g = A[g] + S[g]; fold sum_g 2^(g c) R_g; write affine or Jacobian
__global__ void msm_final_kernel(const xyzz *__restrict__ S, const xyzz *__restrict__ A, uint32_t n_cols, uint32_t G,
                                 uint32_t c, affine *__restrict__ out_affine, jacobian *__restrict__ out_jac) {
    uint32_t col = blockIdx.x;
    if (col >= n_cols || threadIdx.x != 0) return;
    xyzz acc = xyzz_identity();
    for (uint32_t g = G; g-- > 0;) {
        if (g + 1 != G)
            for (uint32_t k = 0; k < c; ++k) acc = xyzz_double(acc);
        xyzz r = xyzz_ld(S + (size_t)col * G + g);
        xyzz a = xyzz_ld(A + (size_t)col * G + g);
        xyzz_add(r, a);
        xyzz_add(acc, r);
    }
    if (out_affine) {
        affine o = xyzz_to_affine_fast(acc);
        fe_st(&out_affine[col].x, o.x);
        fe_st(&out_affine[col].y, o.y);
    }
    if (out_jac) {
        jacobian o = xyzz_to_jacobian(acc);
        fe_st(&out_jac[col].x, o.x);
        fe_st(&out_jac[col].y, o.y);
        fe_st(&out_jac[col].z, o.z);
    }
}

// ------------------------------------------------------------------ SRS table precomputation
// table[(j+1)*n + i] = 2^c * table[j*n + i], affine; one thread per point, one level per launch
__global__ void __launch_bounds__(128) msm_precompute_kernel(affine *__restrict__ table, uint32_t n, uint32_t level, uint32_t c) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine p = affine_load_ro(table + (size_t)(level - 1) * n + i);
    affine o;
    if (affine_is_identity(p)) {
        o = p;
    } else {
        xyzz q = xyzz_double_affine(p);
        for (uint32_t k = 1; k < c; ++k) q = xyzz_double(q);
        o = xyzz_to_affine(q);
    }
    fe_st(&table[(size_t)level * n + i].x, o.x);
    fe_st(&table[(size_t)level * n + i].y, o.y);
}

// ------------------------------------------------------------------ ParamsKZG::setup
// halo2-axiom poly/kzg/commitment.rs `ParamsKZG::setup`: g[i] = s^i G and, directly from s,
// g_lagrange[i] = ((s^n - 1)/n) w^i / (s - w^i) G.  One thread per point: the Fr scalar, then a 254-bit
// double-and-add from the generator and the affine normalisation.  One-time work, not on the prove path.
struct SetupParams {
    fe s;        // the secret, Montgomery Fr
    fe omega;    // 2^k-th root of unity
    fe mult;     // (s^n - 1) / n
    uint32_t n;
};
__global__ void __launch_bounds__(128) srs_setup_kernel(SetupParams p, int lagrange, affine *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    fe sc;
    if (!lagrange) {
        sc = fe_pow_small<Fr>(p.s, i);
    } else {
        fe w = fe_pow_small<Fr>(p.omega, i);
        sc = fe_mul<Fr>(fe_mul<Fr>(p.mult, w), fe_inv<Fr>(fe_sub<Fr>(p.s, w)));
    }
    sc = fe_from_mont<Fr>(sc);
    affine g;
    fe c = fe_zero();
    c.v[0] = 1;
    g.x = fe_to_mont<Fq>(c);
    c.v[0] = 2;
    g.y = fe_to_mont<Fq>(c);
    xyzz acc = xyzz_identity();
    for (int bit = 253; bit >= 0; --bit) {
        acc = xyzz_double(acc);
        if ((sc.v[bit >> 5] >> (bit & 31)) & 1u) xyzz_add_mixed(acc, g);
    }
    affine o = xyzz_to_affine(acc);
    fe_st(&out[i].x, o.x);
    fe_st(&out[i].y, o.y);
}

}  // namespace h2v
