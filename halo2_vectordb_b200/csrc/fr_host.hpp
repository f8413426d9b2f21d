// fr_host.hpp -- BN254 Fr on the host, 4 x 64-bit Montgomery limbs (R = 2^256): the scalar arithmetic of the
// prover's host side (Fiat-Shamir transcript, challenges, SHPLONK point sets, RNG reduction).
//
// Host counterpart of halo2curves bn256::Fr; the byte layout is the same `[u64; 4]`
// little-endian Montgomery form the device code (ff.cuh, 8 x 32-bit limbs) and the C ABI use, so values move
// between the two by memcpy.  Only sequential, latency-bound work runs here (a Poseidon sponge is one dependent
// chain); everything data-parallel is a kernel.
#pragma once
#include <stdint.h>
#include <string.h>

#include "ff.cuh"

namespace h2v {

struct Fr64 {
    uint64_t l[4];
    bool operator==(const Fr64 &o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
    bool operator!=(const Fr64 &o) const { return !(*this == o); }
};

namespace frh {
typedef unsigned __int128 u128;
static const uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t INV = 0xc2e1f593efffffffull;   // -r^{-1} mod 2^64
static const Fr64 ONE = {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}};   // R mod r
static const Fr64 R2 = {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}};
static const Fr64 R3 = {{0x5e94d8e1b4bf0040ull, 0x2a489cbe1cfbb6b8ull, 0x893cc664a19fcfedull, 0x0cf8594b7fcc657cull}};

inline Fr64 zero() { return Fr64{{0, 0, 0, 0}}; }
inline bool is_zero(const Fr64 &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline bool geq_mod(const uint64_t *a) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > MOD[i]) return true;
        if (a[i] < MOD[i]) return false;
    }
    return true;
}
inline void sub_mod(uint64_t *a) {
    u128 bw = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - MOD[i] - (uint64_t)bw;
        a[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
}
// r = t - MOD if t >= MOD (or `top` set) else t; branch-free (the comparison outcome is a coin flip on random data)
inline Fr64 csub_mod(uint64_t t0, uint64_t t1, uint64_t t2, uint64_t t3, uint64_t top) {
    unsigned long long r0, r1, r2, r3;
    unsigned char b = __builtin_sub_overflow(t0, MOD[0], &r0);
    unsigned char b1 = __builtin_sub_overflow(t1, MOD[1], &r1);
    b1 |= __builtin_sub_overflow(r1, (uint64_t)b, &r1);
    unsigned char b2 = __builtin_sub_overflow(t2, MOD[2], &r2);
    b2 |= __builtin_sub_overflow(r2, (uint64_t)b1, &r2);
    unsigned char b3 = __builtin_sub_overflow(t3, MOD[3], &r3);
    b3 |= __builtin_sub_overflow(r3, (uint64_t)b2, &r3);
    const uint64_t keep = (uint64_t)0 - (uint64_t)((b3 != 0) & (top == 0));      // all ones: t < MOD, keep t
    Fr64 r;
    r.l[0] = (t0 & keep) | (r0 & ~keep);
    r.l[1] = (t1 & keep) | (r1 & ~keep);
    r.l[2] = (t2 & keep) | (r2 & ~keep);
    r.l[3] = (t3 & keep) | (r3 & ~keep);
    return r;
}
inline Fr64 add(const Fr64 &a, const Fr64 &b) {
    u128 c = (u128)a.l[0] + b.l[0];
    const uint64_t t0 = (uint64_t)c;
    c = (c >> 64) + a.l[1] + b.l[1];
    const uint64_t t1 = (uint64_t)c;
    c = (c >> 64) + a.l[2] + b.l[2];
    const uint64_t t2 = (uint64_t)c;
    c = (c >> 64) + a.l[3] + b.l[3];
    return csub_mod(t0, t1, t2, (uint64_t)c, 0);      // a + b < 2r < 2^255: no carry out
}
inline Fr64 sub(const Fr64 &a, const Fr64 &b) {
    Fr64 r;
    u128 bw = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.l[i] - b.l[i] - (uint64_t)bw;
        r.l[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    const uint64_t m = (uint64_t)0 - (uint64_t)bw;      // add MOD back when the difference went negative
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)r.l[i] + (MOD[i] & m);
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    return r;
}
inline Fr64 neg(const Fr64 &a) { return is_zero(a) ? a : sub(zero(), a); }
// Montgomery product (CIOS, 4 x 64)
inline Fr64 mul(const Fr64 &a, const Fr64 &b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * INV;
        c = (u128)q * MOD[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)q * MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    return csub_mod(t[0], t[1], t[2], t[3], 0);
}
// ---- wide (unreduced) products for dot products with ONE Montgomery reduction (the Poseidon MDS rows)
// t[0..8) += a * b as a 512-bit integer; the caller keeps the total below 2^512.  Fixed-length carry chains only.
inline void mul_acc_wide(uint64_t t[8], const Fr64 &a, const Fr64 &b) {
    uint64_t p[8];
    {
        u128 c = (u128)a.l[0] * b.l[0];
        p[0] = (uint64_t)c;
        c = (c >> 64) + (u128)a.l[0] * b.l[1];
        p[1] = (uint64_t)c;
        c = (c >> 64) + (u128)a.l[0] * b.l[2];
        p[2] = (uint64_t)c;
        c = (c >> 64) + (u128)a.l[0] * b.l[3];
        p[3] = (uint64_t)c;
        p[4] = (uint64_t)(c >> 64);
    }
    for (int i = 1; i < 4; ++i) {
        u128 c = (u128)a.l[i] * b.l[0] + p[i];
        p[i] = (uint64_t)c;
        c = (c >> 64) + (u128)a.l[i] * b.l[1] + p[i + 1];
        p[i + 1] = (uint64_t)c;
        c = (c >> 64) + (u128)a.l[i] * b.l[2] + p[i + 2];
        p[i + 2] = (uint64_t)c;
        c = (c >> 64) + (u128)a.l[i] * b.l[3] + p[i + 3];
        p[i + 3] = (uint64_t)c;
        p[i + 4] = (uint64_t)(c >> 64);
    }
    u128 c = 0;
    for (int k = 0; k < 8; ++k) {
        c += (u128)t[k] + p[k];
        t[k] = (uint64_t)c;
        c >>= 64;
    }
}
// Montgomery reduction of T < 2^511 with T / R < 2r: (T + m r) / R, one conditional subtraction
inline Fr64 redc_wide(const uint64_t t_in[8]) {
    uint64_t t[8];
    for (int i = 0; i < 8; ++i) t[i] = t_in[i];
    uint64_t top = 0;          // carries out of the running window land here
    for (int i = 0; i < 4; ++i) {
        const uint64_t q = t[i] * INV;
        u128 c = (u128)q * MOD[0] + t[i];
        c = (c >> 64) + (u128)q * MOD[1] + t[i + 1];
        t[i + 1] = (uint64_t)c;
        c = (c >> 64) + (u128)q * MOD[2] + t[i + 2];
        t[i + 2] = (uint64_t)c;
        c = (c >> 64) + (u128)q * MOD[3] + t[i + 3];
        t[i + 3] = (uint64_t)c;
        c >>= 64;
        for (int k = i + 4; k < 8; ++k) {
            c += t[k];
            t[k] = (uint64_t)c;
            c >>= 64;
        }
        top += (uint64_t)c;
    }
    return csub_mod(t[4], t[5], t[6], t[7], top);
}
// sum_k a[k] * b[k] (Montgomery), cnt <= 5: the products are summed unreduced (5 r^2 < 2^511), one reduction
inline Fr64 dot(const Fr64 *a, const Fr64 *b, int cnt) {
    uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < cnt; ++k) mul_acc_wide(t, a[k], b[k]);
    return redc_wide(t);
}
inline Fr64 sqr(const Fr64 &a) { return mul(a, a); }
inline Fr64 from_u64(uint64_t x) { return mul(Fr64{{x, 0, 0, 0}}, R2); }
inline Fr64 to_mont(const Fr64 &canon) { return mul(canon, R2); }
inline Fr64 from_mont(const Fr64 &a) { return mul(a, Fr64{{1, 0, 0, 0}}); }
inline Fr64 pow_u64(const Fr64 &a, uint64_t e) {
    Fr64 acc = ONE;
    for (int i = 63; i >= 0; --i) {
        acc = sqr(acc);
        if ((e >> i) & 1) acc = mul(acc, a);
    }
    return acc;
}
inline Fr64 pow5(const Fr64 &a) {
    Fr64 a2 = sqr(a);
    return mul(sqr(a2), a);
}
inline Fr64 inv(const Fr64 &a) {      // a^(r-2); 0 -> 0
    uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
    Fr64 acc = ONE;
    for (int i = 253; i >= 0; --i) {
        acc = sqr(acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
    }
    return acc;
}
// rotation of a point by omega^rot (EvaluationDomain::rotate_omega), omega / omega_inv given
inline Fr64 rotate(const Fr64 &x, const Fr64 &omega, const Fr64 &omega_inv, int rot) {
    return rot >= 0 ? mul(x, pow_u64(omega, (uint64_t)rot)) : mul(x, pow_u64(omega_inv, (uint64_t)(-(int64_t)rot)));
}
// numeric order of the canonical values: halo2curves `impl Ord for Fr` (compares `to_repr()` from the top byte down)
inline bool less_canonical(const Fr64 &a, const Fr64 &b) {
    Fr64 x = from_mont(a), y = from_mont(b);
    for (int i = 3; i >= 0; --i) {
        if (x.l[i] != y.l[i]) return x.l[i] < y.l[i];
    }
    return false;
}
// value mod r of a 512-bit little-endian integer (halo2curves `from_u512`: d0 * R2 + d1 * R3 in Montgomery arithmetic)
inline Fr64 from_u512(const uint64_t w[8]) {
    Fr64 d0 = {{w[0], w[1], w[2], w[3]}}, d1 = {{w[4], w[5], w[6], w[7]}};
    return add(mul(d0, R2), mul(d1, R3));
}
inline fe to_fe(const Fr64 &a) {
    fe r;
    memcpy(r.v, a.l, 32);
    return r;
}
inline Fr64 from_fe(const fe &a) {
    Fr64 r;
    memcpy(r.l, a.v, 32);
    return r;
}
inline Fr64 load(const uint64_t *p) {
    Fr64 r;
    memcpy(r.l, p, 32);
    return r;
}
inline void store(uint64_t *p, const Fr64 &a) { memcpy(p, a.l, 32); }
}  // namespace frh

}  // namespace h2v
