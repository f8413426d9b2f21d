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
inline Fr64 add(const Fr64 &a, const Fr64 &b) {
    Fr64 r;
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_mod(r.l)) sub_mod(r.l);     // a + b < 2r < 2^255: no carry out
    return r;
}
inline Fr64 sub(const Fr64 &a, const Fr64 &b) {
    Fr64 r;
    u128 bw = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.l[i] - b.l[i] - (uint64_t)bw;
        r.l[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    if (bw) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)r.l[i] + MOD[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
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
    Fr64 r = {{t[0], t[1], t[2], t[3]}};
    if (geq_mod(r.l)) sub_mod(r.l);
    return r;
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
