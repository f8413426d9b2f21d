// ntt.cuh -- radix-2^s Fr NTT passes for EvaluationDomain.
//
// Replaces halo2-axiom arithmetic.rs `best_fft` / `recursive_butterfly_arithmetic` and the
// poly/domain.rs transforms built on it (reached
// from /root/reference/src/scaffold/mod.rs:273,296).  Semantics are those of best_fft:
// natural order in, natural order out, A[j] = sum_i a[i] w^(ij).
//
// Decomposition: the log2(N) = L radix-2 DIT stages are grouped into P = ceil(L/9) passes of
// S_p <= 9 stages.  One pass = one HBM round trip (64 B per element):
//   pass 0   reads the input at bit-reversed addresses (the permutation of best_fft is folded
//            into the addressing, together with zero padding and the coset pre-scaling of
//            coeff_to_extended / divide_by_vanishing_poly) and writes contiguous 2^S blocks;
//   pass p>0 works in place on column tiles: T consecutive elements (T*32 B contiguous) for each
//            of the 2^S strided rows.
// Inside a CTA the 2^(S+logT) <= 2048 elements sit in shared memory as two 16-byte planes with
// an XOR swizzle; each thread keeps 8 elements in registers and runs up to three stages
// (radix-8) between shared-memory exchanges.  Twiddles w^i (i < N/2) are precomputed once per
// domain (twiddle_kernel) and read through the read-only path; a stage-t butterfly on index j
// uses w^((j mod 2^t) << (L-1-t)), exactly the table entry best_fft would use.
#pragma once
#include "ff.cuh"
#include "shoup.cuh"

namespace h2v {

typedef FrP Fr;

struct NttPass {
    const fe *src;          // column 0 of the input  (pass 0) / unused (later passes)
    fe *dst;                // column 0 of the output; later passes work in place on dst
    size_t src_stride;      // elements between columns
    size_t dst_stride;
    const fe *tw;           // w^i, i < 2^(L-1), Montgomery
    const fe *tws;          // the same twiddles for fe_mul_shoup_lazy: (canonical w^i, floor(w^i 2^256 / r)) pairs
    const fe *pre;          // pass 0: input i is multiplied by pre[i % pre_mod]   (nullptr: none)
    const fe *post;         // last pass: output j is multiplied by post[j % post_mod] (nullptr: none)
    uint32_t pre_mod, post_mod;
    uint32_t post_shift;    // last pass, instead of `post`: output j is multiplied by 2^-post_shift (the 1/n of an inverse transform)
    uint32_t n_in;          // pass 0: inputs with index >= n_in read as zero
    uint32_t n_out;         // last pass: only outputs j < n_out are stored
    int L, t0, S, logT;
    int first, last;
};

// shared-memory index swizzle (a bijection on [0, 2048)): makes the stride-2^k accesses of every
// round hit distinct 16-byte bank groups within a quarter warp
__device__ __forceinline__ uint32_t ntt_swz(uint32_t i) { return i ^ ((i >> 3) & 7u) ^ ((i >> 6) & 7u) ^ ((i >> 9) & 7u); }

#define H2V_NTT_PLANE_PAD 4   // uint4 units: shifts plane 1 by half a 128-byte bank row

__device__ __forceinline__ fe ntt_sm_load(const uint4 *sm, uint32_t plane1, uint32_t idx) {
    uint32_t p = ntt_swz(idx);
    uint4 a = sm[p], b = sm[plane1 + p];
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void ntt_sm_store(uint4 *sm, uint32_t plane1, uint32_t idx, const fe &x) {
    uint32_t p = ntt_swz(idx);
    sm[p] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    sm[plane1 + p] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
// the same with the swizzled slot already computed
__device__ __forceinline__ fe ntt_sm_load_slot(const uint4 *sm, uint32_t plane1, uint32_t p) {
    uint4 a = sm[p], b = sm[plane1 + p];
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void ntt_sm_store_slot(uint4 *sm, uint32_t plane1, uint32_t p, const fe &x) {
    sm[p] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    sm[plane1 + p] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ fe fe_load_global(const fe *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ fe fe_load_ro(const fe *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fe r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void fe_store_global(fe *p, const fe &x) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// one radix-2 stage on the 8 register-resident elements: pairs (k, k + 2^U)
// unit: 0 = every butterfly has a table twiddle; 1 = all twiddles are 1 (stage 0); 2 = e_base is 0, so the butterflies
// whose in-thread exponent (k mod 2^U) is 0 have twiddle 1 (first round of pass 0: three of its eight products)
// SHOUP: `tw` is the table of (w, w') pairs and the product is fe_mul_shoup_lazy; else the Montgomery table and fe_mul_lazy
template <int U, bool SHOUP>
__device__ __forceinline__ void ntt_stage(fe (&x)[8], const fe *__restrict__ tw, uint32_t e_base, int L, int unit) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if ((k >> U) & 1) continue;
        // lazy reduction (Harvey): t < 2r; a is brought below theta (just above 2r, see fe_csub_2m_top), so a + t and
        // a + 2r - t stay below theta + 2r < 2^256
        fe t;
        if (unit == 1 || (unit == 2 && (k & ((1 << U) - 1)) == 0)) {
            t = x[k + (1 << U)];       // twiddle 1: no product, only the range step (values < 4r -> t < 2r)
            fe_csub_2m<Fr>(t);
        } else {
            const uint32_t e = e_base + ((uint32_t)(k & ((1 << U) - 1)) << (L - 1 - U));
            if (SHOUP) {
                // twiddle product by Shoup's method (shoup.cuh): 99 + 16 instead of 128 + 8 multiplier instructions, t < 2r
                const fe *wp = tw + 2 * (size_t)e;
                fe w = fe_load_ro(wp), ws = fe_load_ro(wp + 1);
                t = fe_mul_shoup_lazy<Fr>(x[k + (1 << U)], w, ws);
            } else {
                t = fe_mul_lazy<Fr>(x[k + (1 << U)], fe_load_ro(tw + e));
            }
        }
        fe a = x[k];
        fe_csub_2m_top<Fr>(a);
        x[k + (1 << U)] = fe_sub_plus_2m<Fr>(a, t);
        x[k] = fe_add_raw(a, t);
    }
}

// x * 2^-s mod r for 3 <= s <= 31 by ONE word of Montgomery reduction: q = -x r^-1 mod 2^s makes x + q r divisible by 2^s.
// 8 wide multiply-adds instead of the 136 of a product with the field element 2^-s; x < 2^256 gives a result below
// x / 2^s + r < 2r.  Montgomery form is preserved (the factor R rides along).
__device__ __forceinline__ fe fe_div_pow2_lazy(const fe &x, uint32_t s) {
    const uint32_t q = (x.v[0] * Fr::inv()) & ((1u << s) - 1u);
    uint32_t t[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c += (uint64_t)q * Fr::m(i) + x.v[i];
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    t[8] = (uint32_t)c;
    fe r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = __funnelshift_r(t[i], t[i + 1], s);
    return r;
}

// last pass: bring a lazily reduced value (< 4r) back to its canonical representative, folding in the post-scale
__device__ __forceinline__ fe ntt_finish(fe x, const fe *post, uint32_t shift) {
    if (shift) {
        x = fe_div_pow2_lazy(x, shift);
        fe_reduce_once<Fr>(x);
        return x;
    }
    if (post) return fe_mul<Fr>(x, fe_load_ro(post));     // x < 4r, post < r: product < 2r, one conditional subtraction
    fe_csub_2m<Fr>(x);      // x < theta + 2r, slightly above 4r: two exact steps of 2r, then one of r
    fe_csub_2m<Fr>(x);
    fe_reduce_once<Fr>(x);
    return x;
}

// Later passes (t0 > 0) read twiddles that are spread over the whole table -- every butterfly of a stage its own entry --
// and the round's products wait for them (ncu: a quarter of the stall samples of pass 1 were long-scoreboard waits on
// IMAD.WIDE).  The round's <= 7 twiddle lines are therefore requested with prefetch.global.L1 one barrier ahead: before
// the tile load for the first round, before the closing barrier of a round for the next one.  No registers are held.
template <bool SHOUP>
__device__ __forceinline__ void ntt_prefetch_twiddles(const NttPass &p, int bp, int u0, uint32_t jlow) {
    const fe *tw = SHOUP ? p.tws : p.tw;
#pragma unroll
    for (int U = 0; U < 3; ++U) {
        if (U < u0) continue;
        const int t = p.t0 + bp + U;
        const uint32_t e_base = jlow << (p.L - 1 - t);
#pragma unroll
        for (int kk = 0; kk < (1 << U); ++kk) {
            const uint32_t e = e_base + ((uint32_t)kk << (p.L - 1 - U));
            const fe *a = SHOUP ? tw + 2 * (size_t)e : tw + e;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
        }
    }
}

// SHOUP (transforms of up to 2^18 points): Shoup twiddle products from the (w, w') table -- 3-5 % faster there; larger
// transforms keep the Montgomery product, whose twiddle stream is half as wide (measured 4 % faster at 2^20 / 2^22)
template <bool SHOUP>
__global__ void __launch_bounds__(256, 2) ntt_pass_kernel(NttPass p) {
    extern __shared__ uint4 ntt_sm[];
    const int S = p.S, logT = p.logT, L = p.L, t0 = p.t0;
    const uint32_t T = 1u << logT;
    const uint32_t nelem = 1u << (S + logT);
    const uint32_t plane1 = nelem + H2V_NTT_PLANE_PAD;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t tile = blockIdx.x;
    fe *dst = p.dst + (size_t)blockIdx.y * p.dst_stride;

    // ------------------------------------------------------------------ load tile -> shared
    uint32_t hi = 0, lo_tile = 0;
    if (p.first) {
        // element (mid, q): source index rev_S(mid) * 2^H + tile*T + q, H = L - S
        const fe *src = p.src + (size_t)blockIdx.y * p.src_stride;
        const int H = L - S;
        // the launch has nelem / 8 threads (run_ntt), so every thread moves exactly 8 elements: all 8 global loads are issued
        // before the first shared-memory store (one DRAM latency per tile instead of eight dependent load -> store pairs)
        fe xs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const uint32_t e = tid + it * nthr;
            const uint32_t si = ((e >> logT) << H) + tile * T + (e & (T - 1));
            xs[it] = fe_zero();
            if (e < nelem && si < p.n_in) xs[it] = fe_load_global(src + si);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const uint32_t e = tid + it * nthr;
            if (e >= nelem) break;
            const uint32_t q = e & (T - 1), midr = e >> logT;
            const uint32_t si = (midr << H) + tile * T + q;
            fe x = xs[it];
            if (p.pre && si < p.n_in) x = fe_mul<Fr>(x, fe_load_ro(p.pre + (si % p.pre_mod)));
            ntt_sm_store(ntt_sm, plane1, (bitrev(midr, S) << logT) | q, x);
        }
    } else {
        // element (mid, q): index hi * 2^(t0+S) + mid * 2^t0 + lo_tile*T + q, in place on dst
        hi = tile >> (t0 - logT);
        lo_tile = tile & ((1u << (t0 - logT)) - 1);
        if (tid < (nelem >> 3)) {      // twiddles of the first round: bp = 0 (S >= 3), every in-thread stage
            const int bp0 = (3 <= S) ? 0 : S - 3;
            ntt_prefetch_twiddles<SHOUP>(p, bp0, 0 - bp0, lo_tile * T + (tid & (T - 1)));
        }
        const fe *base = dst + ((size_t)hi << (t0 + S)) + (size_t)lo_tile * T;
        fe xs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const uint32_t e = tid + it * nthr;
            if (e < nelem) xs[it] = fe_load_global(base + ((size_t)(e >> logT) << t0) + (e & (T - 1)));
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const uint32_t e = tid + it * nthr;
            if (e < nelem) ntt_sm_store(ntt_sm, plane1, e, xs[it]);
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ rounds of up to 3 stages
    const bool pad4 = p.first && (((uint64_t)p.n_in << 2) <= ((uint64_t)1 << L));
    for (int b = 0; b < S; b += 3) {
        const int bp = (b + 3 <= S) ? b : S - 3;   // the thread's 3 index bits are [bp, bp+3)
        const int u0 = b - bp;                     // stages below u0 were done in the previous round
        for (uint32_t w = tid; w < (nelem >> 3); w += nthr) {
            uint32_t q = w & (T - 1), rest = w >> logT;
            uint32_t low = rest & ((1u << bp) - 1);
            uint32_t mid_base = low | ((rest >> bp) << (bp + 3));
            // the swizzle is XOR-linear and k << (bp + logT) touches bits the base index leaves clear, so the eight
            // shared-memory slots are the base slot XOR a per-round constant: one LOP3 per slot instead of the full swizzle
            const uint32_t slot0 = ntt_swz((mid_base << logT) | q);
            const uint32_t d1 = ntt_swz(1u << (bp + logT)), d2 = ntt_swz(2u << (bp + logT)), d4 = ntt_swz(4u << (bp + logT));
            uint32_t slot[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) slot[k] = slot0 ^ ((k & 1) ? d1 : 0u) ^ ((k & 2) ? d2 : 0u) ^ ((k & 4) ? d4 : 0u);
            fe x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = ntt_sm_load_slot(ntt_sm, plane1, slot[k]);
            // j mod 2^t for the stage at absolute bit t = t0 + bp + u:
            //   lo + ((low + (k mod 2^u) << bp) << t0)
            const uint32_t lo = p.first ? 0u : (lo_tile * T + q);
            const uint32_t jlow = lo + (low << t0);
            if (pad4 && b == 0) {
                // zero-padded input (coeff_to_extended: n_in <= N/4): after the bit reversal only every 4th
                // element is non-zero, so the first two stages just broadcast it -- no products with zero
                x[1] = x[0]; x[2] = x[0]; x[3] = x[0];
                x[5] = x[4]; x[6] = x[4]; x[7] = x[4];
            } else {
                if (u0 <= 0) {
                    int t = t0 + bp;
                    ntt_stage<0, SHOUP>(x, SHOUP ? p.tws : p.tw, jlow << (L - 1 - t), L, t == 0 ? 1 : 0);
                }
                if (u0 <= 1) {
                    int t = t0 + bp + 1;
                    ntt_stage<1, SHOUP>(x, SHOUP ? p.tws : p.tw, jlow << (L - 1 - t), L, t == 1 ? 2 : 0);
                }
            }
            {
                int t = t0 + bp + 2;
                ntt_stage<2, SHOUP>(x, SHOUP ? p.tws : p.tw, jlow << (L - 1 - t), L, t == 2 ? 2 : 0);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) ntt_sm_store_slot(ntt_sm, plane1, slot[k], x[k]);
        }
        if (!p.first && b + 3 < S && tid < (nelem >> 3)) {
            const int nb = b + 3, nbp = (nb + 3 <= S) ? nb : S - 3;
            const uint32_t q = tid & (T - 1), low = (tid >> logT) & ((1u << nbp) - 1);
            ntt_prefetch_twiddles<SHOUP>(p, nbp, nb - nbp, (lo_tile * T + q) + (low << t0));
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ store
    if (p.first) {
        // column q is the contiguous block hi_q * 2^S .. with hi_q = rev_H(tile*T + q)
        const int H = L - S;
        for (uint32_t e = tid; e < nelem; e += nthr) {
            uint32_t mid = e & ((1u << S) - 1), q = e >> S;
            uint32_t j = (bitrev(tile * T + q, H) << S) | mid;
            if (p.last && j >= p.n_out) continue;
            fe x = ntt_sm_load(ntt_sm, plane1, (mid << logT) | q);
            if (p.last) x = ntt_finish(x, p.post ? p.post + (j % p.post_mod) : nullptr, p.post_shift);
            fe_store_global(dst + j, x);
        }
    } else {
        const size_t jbase = ((size_t)hi << (t0 + S)) + (size_t)lo_tile * T;
        for (uint32_t e = tid; e < nelem; e += nthr) {
            uint32_t q = e & (T - 1), mid = e >> logT;
            size_t j = jbase + ((size_t)mid << t0) + q;
            if (p.last && j >= p.n_out) continue;
            fe x = ntt_sm_load(ntt_sm, plane1, e);
            if (p.last) x = ntt_finish(x, p.post ? p.post + (uint32_t)(j % p.post_mod) : nullptr, p.post_shift);
            fe_store_global(dst + j, x);
        }
    }
}

// N <= 4: direct DFT by one thread per column (degenerate domains; keeps the API total)
__global__ void ntt_tiny_kernel(NttPass p) {
    const uint32_t n = 1u << p.L;
    const fe *src = p.src + (size_t)blockIdx.x * p.src_stride;
    fe *dst = p.dst + (size_t)blockIdx.x * p.dst_stride;
    if (threadIdx.x != 0) return;
    fe a[4], o[4];
    for (uint32_t i = 0; i < n; ++i) {
        a[i] = fe_zero();
        if (i < p.n_in) {
            a[i] = fe_load_global(src + i);
            if (p.pre) a[i] = fe_mul<Fr>(a[i], fe_load_ro(p.pre + (i % p.pre_mod)));
        }
    }
    for (uint32_t j = 0; j < n; ++j) {
        fe acc = fe_zero();
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t e = (i * j) & (n - 1);
            fe t = a[i];
            if (n > 1) {
                uint32_t h = n >> 1;
                fe w = fe_load_ro(p.tw + (e & (h - 1)));     // w^(e mod n/2)
                t = fe_mul<Fr>(t, w);
                if (e >= h) t = fe_neg<Fr>(t);               // w^(n/2) = -1
            }
            acc = fe_add<Fr>(acc, t);
        }
        o[j] = acc;
    }
    for (uint32_t j = 0; j < n; ++j) {
        if (j >= p.n_out) continue;
        fe x = o[j];
        if (p.post) x = fe_mul<Fr>(x, fe_load_ro(p.post + (j % p.post_mod)));
        fe_store_global(dst + j, x);
    }
}

struct TwiddleParams {
    fe pows[28];     // w^(2^b), Montgomery
    uint32_t half_n; // table length
};
// tw[i] = w^i by binary exponentiation over the precomputed squarings (Montgomery); behind the table, at
// tw + half_n, the Shoup form of the same entries: pairs (canonical w^i, floor(w^i 2^256 / r)) for the NTT butterflies
__global__ void twiddle_kernel(fe *tw, TwiddleParams p) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.half_n) return;
    fe acc = fe_one<Fr>();
    for (int b = 0; b < 28; ++b) {
        if ((i >> b) == 0) break;
        if ((i >> b) & 1) acc = fe_mul<Fr>(acc, p.pows[b]);
    }
    fe_store_global(tw + i, acc);
    fe *pair = tw + p.half_n + 2 * (size_t)i;
    fe_store_global(pair, fe_from_mont<Fr>(acc));
    fe_store_global(pair + 1, fr_shoup_companion(acc));
}

// elementwise a[i] *= c[i % mod]   (standalone divide_by_vanishing_poly on a device-resident column)
__global__ void fr_scale_mod_kernel(fe *a, size_t stride, const fe *c, uint32_t mod, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe *col = a + (size_t)blockIdx.y * stride;
    fe_store_global(col + i, fe_mul<Fr>(fe_load_global(col + i), fe_load_ro(c + (i % mod))));
}

}  // namespace h2v
