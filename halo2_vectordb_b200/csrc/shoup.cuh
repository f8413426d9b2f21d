// shoup.cuh -- product of a field element with a PRECOMPUTED constant (the NTT twiddles) in fewer multiplier
// instructions than a Montgomery product.
//
// For a constant w < m (canonical) with w' = floor(w 2^256 / m) (Shoup's trick, as in Harvey's NTT), and any a < 2^256:
//     q = floor(a w' / 2^256)          -- only the HIGH half of a 512-bit product
//     t = a w - q m   (mod 2^256)      -- only the LOW halves of two 512-bit products
// gives t = a w - q m exactly, with 0 <= t < m (1 + a / 2^256) < 2m.  Unlike a Montgomery reduction, an error in q only
// moves t by a multiple of m, so the high half may be TRUNCATED: the partial products a_i w'_j with i + j <= 5 are
// dropped (their sum is below 2^227), which makes the computed q either q or q - 1 and t < 3m; one predicated
// subtraction of m, decided by the top limb alone, brings t back below 2m (fe_mul_shoup_lazy's contract, the same as
// fe_mul_lazy's).  Multiplier work: 43 wide products (high part) + 2 x (28 wide + 8 low-only) = 99 IMAD.WIDE + 16 IMAD
// against the 128 IMAD.WIDE + 8 IMAD of the CIOS product: 428 instead of 528 cycles of the multiplier pipe per warp.
// The data stays in Montgomery form (a = x R): (x R) w mod m is the Montgomery form of x w, so the twiddle tables
// hold the CANONICAL w next to w' and nothing else changes representation.
//
// Accumulation follows ff.cuh: one accumulator for the partial products that land on an even limb (E) and one for
// those on an odd limb (O), so every run of products a_i b_j, a_{i+2} b_j, ... is ONE carry chain of
// mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32(.X); E + O is formed once at the end.
#pragma once
#include "ff.cuh"

namespace h2v {

#ifdef __CUDA_ARCH__
namespace ptx {
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint32_t r;
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
}  // namespace ptx

// limbs 8..15 of sum_{i + j >= 6} a_i b_j 2^(32 (i + j)): floor(a b / 2^256) or one less
__device__ __forceinline__ void mul_high_trunc(uint32_t (&q)[8], const uint32_t *a, const uint32_t *b) {
    uint32_t E[17], O[17];      // absolute limbs; only 6..16 are touched
#pragma unroll
    for (int i = 0; i < 17; ++i) E[i] = O[i] = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int P = 0; P < 2; ++P) {
            // the products a_i b_j with i = P, P + 2, ...: limbs i + j of one parity, one carry chain
            uint32_t *acc = ((P + j) & 1) ? O : E;
            bool first = true;
            int top = 0;
#pragma unroll
            for (int i = P; i < 8; i += 2) {
                if (i + j < 6) continue;
                const int l = i + j;
                acc[l] = first ? ptx::mad_lo_cc(a[i], b[j], acc[l]) : ptx::madc_lo_cc(a[i], b[j], acc[l]);
                acc[l + 1] = ptx::madc_hi_cc(a[i], b[j], acc[l + 1]);
                first = false;
                top = l + 2;
            }
            // the limb above a chain holds at most the carry of the previous row's chain (rows are visited in increasing
            // j and each chain ends at most two limbs above the previous one of the same accumulator): no overflow
            if (!first && top <= 15) acc[top] = ptx::addc(acc[top], 0u);
        }
    }
    // q = (E + O) >> 256, with the carries out of limbs 6 and 7
    uint32_t s = ptx::add_cc(E[6], O[6]);
    s = ptx::addc_cc(E[7], O[7]);
    (void)s;
#pragma unroll
    for (int i = 8; i < 15; ++i) q[i - 8] = ptx::addc_cc(E[i], O[i]);
    q[7] = ptx::addc(E[15], O[15]);
}

// limbs 0..7 of  acc + a b  on the even / odd accumulator pair (E, O): products with i + j <= 7, those on limb 7 low-only
__device__ __forceinline__ void mul_low_acc(uint32_t (&E)[8], uint32_t (&O)[8], const uint32_t *a, const uint32_t *b) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int P = 0; P < 2; ++P) {
            uint32_t *acc = ((P + j) & 1) ? O : E;
            bool first = true;
#pragma unroll
            for (int i = P; i < 8; i += 2) {
                if (i + j > 7) continue;
                const int l = i + j;
                if (l == 7) {
                    acc[7] = first ? ptx::mad_lo(a[i], b[j], acc[7]) : ptx::madc_lo(a[i], b[j], acc[7]);
                } else {
                    acc[l] = first ? ptx::mad_lo_cc(a[i], b[j], acc[l]) : ptx::madc_lo_cc(a[i], b[j], acc[l]);
                    acc[l + 1] = (l + 1 == 7) ? ptx::madc_hi(a[i], b[j], acc[l + 1]) : ptx::madc_hi_cc(a[i], b[j], acc[l + 1]);
                }
                first = false;
            }
        }
    }
}

// t = a w (mod m), t < 2m, for any a < 2^256; w canonical (< m), wp = floor(w 2^256 / m)
template <class F> __device__ __forceinline__ fe fe_mul_shoup_lazy(const fe &a, const fe &w, const fe &wp) {
    uint32_t q[8];
    mul_high_trunc(q, a.v, wp.v);
    uint32_t E[8], O[8], nm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) E[i] = O[i] = 0;
    // -m mod 2^256 (compile-time constants): low(a w) + low(q (-m)) = a w - q m  (mod 2^256)
    {
        uint32_t bw = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint64_t d = (uint64_t)0 - F::m(i) - bw;
            nm[i] = (uint32_t)d;
            bw = (uint32_t)(d >> 63);
        }
    }
    mul_low_acc(E, O, a.v, w.v);
    mul_low_acc(E, O, q, nm);
    fe t;
    t.v[0] = ptx::add_cc(E[0], O[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) t.v[i] = ptx::addc_cc(E[i], O[i]);
    t.v[7] = ptx::addc(E[7], O[7]);
    // t < 3m (q may be one short); 2^254 lies in (m, 2m] for both BN254 fields, so: t >= 2^254 -> t - m in [0, 2m)
    if (t.v[7] >= 0x40000000u) {
        uint32_t mm[8], d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mm[i] = F::m(i);
        raw_sub(d, t.v, mm);
#pragma unroll
        for (int i = 0; i < 8; ++i) t.v[i] = d[i];
    }
    return t;
}
#else
// host instantiation (tests of the shared formulas): the same truncated quotient and correction in plain C
template <class F> inline fe fe_mul_shoup_lazy(const fe &a, const fe &w, const fe &wp) {
    uint64_t acc[17];
    for (int i = 0; i < 17; ++i) acc[i] = 0;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            if (i + j < 6) continue;
            const uint64_t p = (uint64_t)a.v[i] * wp.v[j];
            acc[i + j] += (uint32_t)p;
            acc[i + j + 1] += p >> 32;
        }
    uint32_t q[8];
    uint64_t c = 0;
    for (int i = 6; i < 16; ++i) {
        c += acc[i];
        if (i >= 8) q[i - 8] = (uint32_t)c;
        c >>= 32;
    }
    uint64_t lo[9];
    for (int i = 0; i < 9; ++i) lo[i] = 0;
    uint32_t nm[8];
    {
        uint32_t bw = 0;
        for (int i = 0; i < 8; ++i) {
            const uint64_t d = (uint64_t)0 - F::m(i) - bw;
            nm[i] = (uint32_t)d;
            bw = (uint32_t)(d >> 63);
        }
    }
    for (int i = 0; i < 8; ++i)
        for (int j = 0; i + j < 8; ++j) {
            const uint64_t p1 = (uint64_t)a.v[i] * w.v[j], p2 = (uint64_t)q[i] * nm[j];
            lo[i + j] += (uint64_t)(uint32_t)p1 + (uint32_t)p2;
            lo[i + j + 1] += (p1 >> 32) + (p2 >> 32);
        }
    fe t;
    c = 0;
    for (int i = 0; i < 8; ++i) {
        c += lo[i];
        t.v[i] = (uint32_t)c;
        c >>= 32;
    }
    if (t.v[7] >= 0x40000000u) {
        uint32_t mm[8], d[8];
        for (int i = 0; i < 8; ++i) mm[i] = F::m(i);
        raw_sub(d, t.v, mm);
        for (int i = 0; i < 8; ++i) t.v[i] = d[i];
    }
    return t;
}
#endif

// host + device, not performance critical: the Shoup companion w' = floor(w 2^256 / m) of a twiddle, from its
// MONTGOMERY form w_m = w 2^256 mod m:  w 2^256 = w' m + w_m, so w' = -w_m m^-1 (mod 2^256) = low(w_m * NINV) with
// NINV = -m^-1 mod 2^256 (whose lowest limb is the Montgomery constant F::inv()).
H2V_HD fe fe_mullo256(const fe &a, const uint32_t *b) {
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (i + j > 7) continue;
            c += (uint64_t)a.v[i] * b[j] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
    }
    fe r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = t[i];
    return r;
}
// -r^-1 mod 2^256 for BN254 Fr (checked by tests/test_host_logic.py against a big-integer computation)
H2V_HD uint32_t fr_ninv256(int i) {
    return i == 0 ? 0xefffffffu : i == 1 ? 0xc2e1f593u : i == 2 ? 0x4c6911b3u : i == 3 ? 0x6586864bu
         : i == 4 ? 0x99062391u : i == 5 ? 0xe39a9828u : i == 6 ? 0x0d8341b2u : 0x73f82f1du;
}
H2V_HD fe fr_shoup_companion(const fe &w_mont) {
    uint32_t ni[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ni[i] = fr_ninv256(i);
    return fe_mullo256(w_mont, ni);
}

}  // namespace h2v
