// g2_host.hpp -- BN254 G2 on the host, just enough for `ParamsKZG::setup`'s verifier-side outputs.
//
// [UPSTREAM] halo2-axiom poly/kzg/commitment.rs `ParamsKZG::setup`: `g2 = G2Affine::generator()`, `s_g2 = (g2 * s).into()`,
// written after the G1 bases by `ParamsKZG::write` (reached from /root/reference/src/scaffold/mod.rs:260 `gen_srs`).
// G2 is the twist y^2 = x^3 + 3/(9 + i) over Fq2 = Fq[i]/(i^2 + 1); halo2curves stores Fq2 as {c0, c1}, each a
// Montgomery Fq, and G2Affine as {x, y} (128 bytes).  One scalar multiplication per SRS: plain double-and-add in
// Jacobian coordinates over the portable host field code (ff.cuh).
#pragma once
#include "ff.cuh"

namespace h2v {
namespace g2h {

struct fq2 {
    fe c0, c1;
};
inline fq2 add(const fq2 &a, const fq2 &b) { return fq2{fe_add<FqP>(a.c0, b.c0), fe_add<FqP>(a.c1, b.c1)}; }
inline fq2 sub(const fq2 &a, const fq2 &b) { return fq2{fe_sub<FqP>(a.c0, b.c0), fe_sub<FqP>(a.c1, b.c1)}; }
inline fq2 dbl(const fq2 &a) { return add(a, a); }
inline fq2 mul(const fq2 &a, const fq2 &b) {
    fe t0 = fe_mul<FqP>(a.c0, b.c0), t1 = fe_mul<FqP>(a.c1, b.c1);
    fe s = fe_mul<FqP>(fe_add<FqP>(a.c0, a.c1), fe_add<FqP>(b.c0, b.c1));
    return fq2{fe_sub<FqP>(t0, t1), fe_sub<FqP>(fe_sub<FqP>(s, t0), t1)};
}
inline fq2 sqr(const fq2 &a) { return mul(a, a); }
inline bool is_zero(const fq2 &a) { return fe_is_zero(a.c0) && fe_is_zero(a.c1); }
inline fq2 inv(const fq2 &a) {
    fe nrm = fe_inv<FqP>(fe_add<FqP>(fe_sqr<FqP>(a.c0), fe_sqr<FqP>(a.c1)));
    return fq2{fe_mul<FqP>(a.c0, nrm), fe_neg<FqP>(fe_mul<FqP>(a.c1, nrm))};
}
inline fe fq_from_words(const uint32_t w[8]) {      // canonical little-endian words -> Montgomery
    fe c;
    for (int i = 0; i < 8; ++i) c.v[i] = w[i];
    return fe_to_mont<FqP>(c);
}
struct affine2 {
    fq2 x, y;
};     // 128 bytes, halo2curves G2Affine layout; identity = all zero
struct jac2 {
    fq2 x, y, z;
};
// EIP-197 generator (tests/golden/external_vectors.json)
inline affine2 generator() {
    static const uint32_t X0[8] = {0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu};
    static const uint32_t X1[8] = {0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u};
    static const uint32_t Y0[8] = {0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u};
    static const uint32_t Y1[8] = {0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u};
    affine2 g;
    g.x = fq2{fq_from_words(X0), fq_from_words(X1)};
    g.y = fq2{fq_from_words(Y0), fq_from_words(Y1)};
    return g;
}
inline jac2 jdouble(const jac2 &p) {
    if (is_zero(p.z)) return p;
    // dbl-2009-l, a = 0
    fq2 A = sqr(p.x), B = sqr(p.y), C = sqr(B);
    fq2 D = dbl(sub(sub(sqr(add(p.x, B)), A), C));
    fq2 E = add(dbl(A), A), F = sqr(E);
    jac2 r;
    r.x = sub(F, dbl(D));
    fq2 c8 = dbl(dbl(dbl(C)));
    r.y = sub(mul(E, sub(D, r.x)), c8);
    r.z = dbl(mul(p.y, p.z));
    return r;
}
inline jac2 jadd_mixed(const jac2 &p, const affine2 &q) {
    if (is_zero(p.z)) {
        jac2 r{q.x, q.y, fq2{fe_one<FqP>(), fe_zero()}};
        return r;
    }
    fq2 z1z1 = sqr(p.z), u2 = mul(q.x, z1z1), s2 = mul(mul(q.y, p.z), z1z1);
    if (is_zero(sub(u2, p.x))) {
        if (is_zero(sub(s2, p.y))) return jdouble(p);
        return jac2{fq2{fe_zero(), fe_zero()}, fq2{fe_one<FqP>(), fe_zero()}, fq2{fe_zero(), fe_zero()}};
    }
    fq2 h = sub(u2, p.x), hh = sqr(h), i = dbl(dbl(hh)), j = mul(h, i), rr = dbl(sub(s2, p.y)), v = mul(p.x, i);
    jac2 r;
    r.x = sub(sub(sqr(rr), j), dbl(v));
    r.y = sub(mul(rr, sub(v, r.x)), dbl(mul(p.y, j)));
    r.z = sub(sub(sqr(add(p.z, h)), z1z1), hh);
    return r;
}
inline affine2 to_affine(const jac2 &p) {
    affine2 r;
    if (is_zero(p.z)) {
        r.x = fq2{fe_zero(), fe_zero()};
        r.y = r.x;
        return r;
    }
    fq2 zi = inv(p.z), zi2 = sqr(zi);
    r.x = mul(p.x, zi2);
    r.y = mul(p.y, mul(zi2, zi));
    return r;
}
// k given as a canonical (non-Montgomery) little-endian 256-bit integer
inline affine2 mul(const affine2 &base, const fe &k_canon) {
    jac2 acc{fq2{fe_zero(), fe_zero()}, fq2{fe_one<FqP>(), fe_zero()}, fq2{fe_zero(), fe_zero()}};
    for (int bit = 255; bit >= 0; --bit) {
        acc = jdouble(acc);
        if ((k_canon.v[bit >> 5] >> (bit & 31)) & 1u) acc = jadd_mixed(acc, base);
    }
    return to_affine(acc);
}
inline bool is_on_curve(const affine2 &p) {
    // b' = 3 / (9 + i)
    fe three = fe_zero(), nine = fe_zero(), one = fe_zero();
    three.v[0] = 3; nine.v[0] = 9; one.v[0] = 1;
    fq2 b = mul(fq2{fe_to_mont<FqP>(three), fe_zero()}, inv(fq2{fe_to_mont<FqP>(nine), fe_to_mont<FqP>(one)}));
    fq2 lhs = sqr(p.y), rhs = add(mul(sqr(p.x), p.x), b);
    return is_zero(sub(lhs, rhs));
}

}  // namespace g2h
}  // namespace h2v
