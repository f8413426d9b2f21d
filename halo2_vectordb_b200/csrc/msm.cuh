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

// ------------------------------------------------------------------ digits + histogram
// keys[(col*W + j)*n + i] = (|d|-1) | sign<<31, or INVALID for d == 0
__global__ void __launch_bounds__(256) msm_digits_kernel(const fe *__restrict__ scalars, size_t col_stride,
                                                         uint32_t *__restrict__ keys, uint32_t *__restrict__ counts,
                                                         MsmShape sh) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t col = blockIdx.y;
    if (i >= sh.n) return;
    fe s = fe_from_mont<Fr>(fe_ld(scalars + (size_t)col * col_stride + i));
    uint32_t t[9];
    {   // s + K  (K can reach W*c <= 288 bits: 9 limbs)
        uint64_t cy = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cy += (uint64_t)s.v[k] + sh.kadd[k];
            t[k] = (uint32_t)cy;
            cy >>= 32;
        }
        t[8] = (uint32_t)cy + sh.kadd[8];
    }
    const uint32_t c = sh.c, mask = (1u << c) - 1, half = (1u << (c - 1)) - 1;
    for (uint32_t j = 0; j < sh.W; ++j) {
        uint32_t bit = j * c, w = bit >> 5, sft = bit & 31;
        uint32_t lo = w < 9 ? t[w] : 0u, hi = (w + 1) < 9 ? t[w + 1] : 0u;
        uint32_t d = (uint32_t)((((uint64_t)hi << 32) | lo) >> sft) & mask;
        int32_t sd = (int32_t)d - (int32_t)half;
        uint32_t key = H2V_KEY_INVALID;
        if (sd != 0) {
            uint32_t mag = (uint32_t)(sd < 0 ? -sd : sd) - 1;
            key = mag | (sd < 0 ? 0x80000000u : 0u);
            uint32_t g = sh.G > 1 ? j : 0;
            atomicAdd(&counts[((size_t)col * sh.G + g) * sh.nb + mag], 1u);
        }
        keys[((size_t)col * sh.W + j) * sh.n + i] = key;
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

// ------------------------------------------------------------------ scatter
// entries[pos] = (point_ref | sign<<31, global bucket)
// Only digits with magnitude in [mag_lo, mag_hi) are placed: for a very large MSM the host sweeps the bucket
// space in slices so that the randomly written part of `entries` stays L2-resident (the keys are re-read per
// slice, which is sequential and cheap).
__global__ void __launch_bounds__(256) msm_scatter_kernel(const uint32_t *__restrict__ keys, uint32_t *__restrict__ cursor,
                                                          uint2 *__restrict__ entries, MsmShape sh, uint32_t mag_lo, uint32_t mag_hi) {
    // one thread per scalar, looping over its W digits four at a time: four independent atomics in flight
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t col = blockIdx.y;
    if (i >= sh.n) return;
    const uint32_t *kp = keys + (size_t)col * sh.W * sh.n + i;
    for (uint32_t j0 = 0; j0 < sh.W; j0 += 4) {
        uint32_t key[4], b[4], pos[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) key[u] = (j0 + u < sh.W) ? kp[(size_t)(j0 + u) * sh.n] : H2V_KEY_INVALID;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint32_t mag = key[u] & 0x7fffffffu;
            if (key[u] == H2V_KEY_INVALID || mag < mag_lo || mag >= mag_hi) {
                key[u] = H2V_KEY_INVALID;
                continue;
            }
            uint32_t g = sh.G > 1 ? j0 + u : 0;
            b[u] = (col * sh.G + g) * sh.nb + mag;
            pos[u] = atomicAdd(&cursor[b[u]], 1u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (key[u] == H2V_KEY_INVALID) continue;
            uint32_t pref = (sh.G > 1 ? i : (j0 + u) * sh.pstride + i) | (key[u] & 0x80000000u);
            entries[pos[u]] = make_uint2(pref, b[u]);
        }
    }
}

// ------------------------------------------------------------------ chunked accumulation
// Thread t sums entries [t*chunk, (t+1)*chunk).  A run (maximal same-bucket subsequence) whose
// bucket lies wholly inside the chunk goes straight to buckets[b]; a run of a bucket that began in
// an earlier chunk goes to edges[2t] ("head"), one that continues into a later chunk to
// edges[2t+1] ("tail").  msm_finish adds tail(t0) + head(t0+1..t1) for straddling buckets.
__global__ void __launch_bounds__(128, 4) msm_accumulate_kernel(const uint2 *__restrict__ entries,
                                                             const uint32_t *__restrict__ offsets, uint32_t n_buckets,
                                                             const affine *__restrict__ points, xyzz *__restrict__ buckets,
                                                             xyzz *__restrict__ edges, uint32_t chunk) {
    const uint32_t M = offsets[n_buckets];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t start64 = (uint64_t)t * chunk;
    if (start64 >= M) return;
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
            acc = xyzz_identity();
            cur_b = nxt.y;
            first_run = false;
        }
        ent = nxt;
    }
}

// one thread per bucket: empty -> identity; straddling -> tail(t0) + heads(t0+1..t1).  Buckets that span
// more than `long_span` chunks (skewed witness columns put a large share of a column into a few buckets:
// scalar 1, the shared high digits of r - small) are queued for msm_finish_long_kernel instead.
#define H2V_LONG_SPAN 16
__global__ void __launch_bounds__(128) msm_finish_kernel(const uint32_t *__restrict__ offsets, uint32_t n_buckets,
                                                         const xyzz *__restrict__ edges, xyzz *__restrict__ buckets,
                                                         uint32_t chunk, uint32_t *__restrict__ long_list,
                                                         uint32_t *__restrict__ long_count) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    uint32_t s = offsets[b], e = offsets[b + 1];
    if (s == e) {
        xyzz_st(buckets + b, xyzz_identity());
        return;
    }
    uint32_t t0 = s / chunk, t1 = (e - 1) / chunk;
    if (t0 == t1) return;
    if (t1 - t0 > H2V_LONG_SPAN) {
        long_list[atomicAdd(long_count, 1u)] = b;
        return;
    }
    xyzz acc = xyzz_ld(edges + 2 * (size_t)t0 + 1);
    for (uint32_t t = t0 + 1; t <= t1; ++t) {
        xyzz h = xyzz_ld(edges + 2 * (size_t)t);
        xyzz_add(acc, h);
    }
    xyzz_st(buckets + b, acc);
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
__global__ void __launch_bounds__(128, 3) msm_reduce_kernel(const xyzz *__restrict__ S_in, const xyzz *__restrict__ A_in,
                                                            xyzz *__restrict__ S_out, xyzz *__restrict__ A_out,
                                                            uint32_t cnt_in, uint32_t cnt_out, uint32_t n_inst, uint32_t shift,
                                                            uint32_t log_seg) {
    uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_inst * cnt_out) return;
    uint32_t inst = gid / cnt_out, s = gid % cnt_out;
    uint32_t lo = s << log_seg, hi = min(lo + (1u << log_seg), cnt_in);
    const xyzz *Sin = S_in + (size_t)inst * cnt_in;
    xyzz run = xyzz_identity(), tz = xyzz_identity();
    for (uint32_t r = hi; r-- > lo + 1;) {
        xyzz v = xyzz_ld(Sin + r);
        xyzz_add(run, v);
        xyzz_add(tz, run);
    }
    {
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

// Tail of the reduction once a few thousand partials per instance are left: the serial radix levels would be
// latency-bound there, so the weighted sum is taken bit by bit with parallel tree sums instead:
//   sum_i i * S[i] = sum_b 2^b * T_b,   T_b = sum over { i : bit b of i set } of S[i].
// grid (nbits + 2, n_inst): CTA `which` < nbits computes T_which, nbits the plain sum of A, nbits + 1 that of S.
__global__ void __launch_bounds__(256) msm_reduce_bits_kernel(const xyzz *__restrict__ S_in, const xyzz *__restrict__ A_in,
                                                              uint32_t cnt, uint32_t nbits, xyzz *__restrict__ T) {
    __shared__ xyzz sm8[8];
    const uint32_t which = blockIdx.x, inst = blockIdx.y;
    const xyzz *src = (which == nbits ? A_in : S_in) + (size_t)inst * cnt;
    xyzz acc = xyzz_identity();
    if (which != nbits || A_in) {
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
            if (which < nbits && !((i >> which) & 1u)) continue;
            xyzz v = xyzz_ld(src + i);
            xyzz_add(acc, v);
        }
    }
    acc = xyzz_block_sum_256(acc, sm8);
    if (threadIdx.x == 0) xyzz_st(T + (size_t)inst * (nbits + 2) + which, acc);
}
// one warp per instance: lane b scales its bit sum by 2^b (b doublings, all lanes in lockstep), a shuffle tree
// adds them, then the common weight 2^shift and the plain sum of A
__global__ void msm_reduce_combine_kernel(const xyzz *__restrict__ T, uint32_t nbits, uint32_t shift, uint32_t n_inst,
                                          xyzz *__restrict__ S_out, xyzz *__restrict__ A_out) {
    const uint32_t inst = blockIdx.x, lane = threadIdx.x;
    if (inst >= n_inst) return;
    const xyzz *t = T + (size_t)inst * (nbits + 2);
    xyzz acc = lane < nbits ? xyzz_ld(t + lane) : xyzz_identity();
    for (uint32_t k = 0; k + 1 < nbits; ++k)
        if (k < lane && lane < nbits) acc = xyzz_double(acc);
    acc = xyzz_warp_sum(acc);
    if (lane != 0) return;
    for (uint32_t k = 0; k < shift; ++k) acc = xyzz_double(acc);
    xyzz a = xyzz_ld(t + nbits);
    xyzz_add(acc, a);
    xyzz_st(S_out + inst, xyzz_ld(t + nbits + 1));
    xyzz_st(A_out + inst, acc);
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
