// ec.cuh -- BN254 G1 (y^2 = x^3 + 3 over Fq) group law for the MSM kernels.
//
// Replaces (device side) halo2curves bn256::{G1Affine, G1} mixed add / add / double
//  as used inside halo2-axiom `multiexp_serial`
// (reached from /root/reference/src/scaffold/mod.rs:273,296).
// Affine points use the halo2curves byte layout (x, y Montgomery limbs; identity = (0,0)).
// Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2;
// identity: ZZ = 0): a mixed add is 8M + 2S, a full add 12M + 2S, with complete handling of
// P + P, P + (-P) and identities -- results are compared after affine normalisation, which is
// representative-independent.
#pragma once
#include "ff.cuh"

namespace h2v {

struct alignas(16) affine {
    fe x, y;
};
struct alignas(16) xyzz {
    fe x, y, zz, zzz;
};
struct alignas(16) jacobian {
    fe x, y, z;
};

typedef FqP Fq;

H2V_HD bool affine_is_identity(const affine &p) { return fe_is_zero(p.x) && fe_is_zero(p.y); }
H2V_HD xyzz xyzz_identity() {
    xyzz r;
    r.x = fe_zero(); r.y = fe_zero(); r.zz = fe_zero(); r.zzz = fe_zero();
    return r;
}
H2V_HD bool xyzz_is_identity(const xyzz &p) { return fe_is_zero(p.zz); }
H2V_HD xyzz xyzz_from_affine(const affine &p) {
    xyzz r;
    if (affine_is_identity(p)) return xyzz_identity();
    r.x = p.x; r.y = p.y; r.zz = fe_one<Fq>(); r.zzz = fe_one<Fq>();
    return r;
}
H2V_HD affine affine_neg(const affine &p) {
    affine r;
    r.x = p.x;
    r.y = fe_neg<Fq>(p.y);
    return r;
}

// 2 * (affine p), p != identity   (mdbl-2008-s-1, a = 0)
H2V_HD xyzz xyzz_double_affine(const affine &p) {
    xyzz r;
    fe u = fe_dbl<Fq>(p.y);
    fe v = fe_sqr<Fq>(u);
    fe w = fe_mul<Fq>(u, v);
    fe s = fe_mul<Fq>(p.x, v);
    fe xx = fe_sqr<Fq>(p.x);
    fe m = fe_add<Fq>(fe_dbl<Fq>(xx), xx);
    r.x = fe_sub<Fq>(fe_sub<Fq>(fe_sqr<Fq>(m), s), s);
    r.y = fe_sub<Fq>(fe_mul<Fq>(m, fe_sub<Fq>(s, r.x)), fe_mul<Fq>(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}
// 2 * p   (dbl-2008-s-1, a = 0)
H2V_HD xyzz xyzz_double(const xyzz &p) {
    if (xyzz_is_identity(p)) return p;
    xyzz r;
    fe u = fe_dbl<Fq>(p.y);
    fe v = fe_sqr<Fq>(u);
    fe w = fe_mul<Fq>(u, v);
    fe s = fe_mul<Fq>(p.x, v);
    fe xx = fe_sqr<Fq>(p.x);
    fe m = fe_add<Fq>(fe_dbl<Fq>(xx), xx);
    r.x = fe_sub<Fq>(fe_sub<Fq>(fe_sqr<Fq>(m), s), s);
    r.y = fe_sub<Fq>(fe_mul<Fq>(m, fe_sub<Fq>(s, r.x)), fe_mul<Fq>(w, p.y));
    r.zz = fe_mul<Fq>(v, p.zz);
    r.zzz = fe_mul<Fq>(w, p.zzz);
    return r;
}
// acc += q (affine)   (madd-2008-s)
H2V_HD void xyzz_add_mixed(xyzz &acc, const affine &q) {
    if (affine_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = xyzz_from_affine(q);
        return;
    }
    fe u2 = fe_mul<Fq>(q.x, acc.zz);
    fe s2 = fe_mul<Fq>(q.y, acc.zzz);
    fe p = fe_sub<Fq>(u2, acc.x);
    fe r = fe_sub<Fq>(s2, acc.y);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) acc = xyzz_double_affine(q);
        else acc = xyzz_identity();
        return;
    }
    fe pp = fe_sqr<Fq>(p);
    fe ppp = fe_mul<Fq>(p, pp);
    fe qq = fe_mul<Fq>(acc.x, pp);
    fe x3 = fe_sub<Fq>(fe_sub<Fq>(fe_sub<Fq>(fe_sqr<Fq>(r), ppp), qq), qq);
    fe y3 = fe_sub<Fq>(fe_mul<Fq>(r, fe_sub<Fq>(qq, x3)), fe_mul<Fq>(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fe_mul<Fq>(acc.zz, pp);
    acc.zzz = fe_mul<Fq>(acc.zzz, ppp);
}
#ifdef __CUDA_ARCH__
// The same addition for the MSM inner loop, with lazily reduced accumulator coordinates in [0, 2p): a Montgomery
// product of two values below 2p is below 4p^2/R + p = 1.76 p, so the products skip their final subtraction and the
// differences add 2p instead of p when they borrow.  q is canonical.  xyzz_canon() brings the result back to [0, p).
__device__ __forceinline__ void xyzz_add_mixed_lazy(xyzz &acc, const affine &q) {
    if (affine_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = xyzz_from_affine(q);
        return;
    }
    fe u2 = fe_mul_lazy<Fq>(q.x, acc.zz);
    fe s2 = fe_mul_lazy<Fq>(q.y, acc.zzz);
    fe p = fe_sub_lazy<Fq>(u2, acc.x);
    fe r = fe_sub_lazy<Fq>(s2, acc.y);
    if (fe_is_zero_lazy<Fq>(p)) {
        if (fe_is_zero_lazy<Fq>(r)) acc = xyzz_double_affine(q);
        else acc = xyzz_identity();
        return;
    }
    fe pp = fe_sqr_lazy<Fq>(p);
    fe ppp = fe_mul_lazy<Fq>(p, pp);
    fe qq = fe_mul_lazy<Fq>(acc.x, pp);
    fe x3 = fe_sub_lazy<Fq>(fe_sub_lazy<Fq>(fe_sub_lazy<Fq>(fe_sqr_lazy<Fq>(r), ppp), qq), qq);
    fe y3 = fe_sub_lazy<Fq>(fe_mul_lazy<Fq>(r, fe_sub_lazy<Fq>(qq, x3)), fe_mul_lazy<Fq>(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fe_mul_lazy<Fq>(acc.zz, pp);
    acc.zzz = fe_mul_lazy<Fq>(acc.zzz, ppp);
}
__device__ __forceinline__ void xyzz_canon(xyzz &a) {
    fe_reduce_once<Fq>(a.x);
    fe_reduce_once<Fq>(a.y);
    fe_reduce_once<Fq>(a.zz);
    fe_reduce_once<Fq>(a.zzz);
}
#else   // host compilation pass: the names must exist, the host never runs the lazy variant
inline void xyzz_add_mixed_lazy(xyzz &acc, const affine &q) { xyzz_add_mixed(acc, q); }
inline void xyzz_canon(xyzz &) {}
#endif
// acc += q   (add-2008-s)
H2V_HD void xyzz_add(xyzz &acc, const xyzz &q) {
    if (xyzz_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = q;
        return;
    }
    fe u1 = fe_mul<Fq>(acc.x, q.zz);
    fe u2 = fe_mul<Fq>(q.x, acc.zz);
    fe s1 = fe_mul<Fq>(acc.y, q.zzz);
    fe s2 = fe_mul<Fq>(q.y, acc.zzz);
    fe p = fe_sub<Fq>(u2, u1);
    fe r = fe_sub<Fq>(s2, s1);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) acc = xyzz_double(acc);
        else acc = xyzz_identity();
        return;
    }
    fe pp = fe_sqr<Fq>(p);
    fe ppp = fe_mul<Fq>(p, pp);
    fe qq = fe_mul<Fq>(u1, pp);
    fe x3 = fe_sub<Fq>(fe_sub<Fq>(fe_sub<Fq>(fe_sqr<Fq>(r), ppp), qq), qq);
    fe y3 = fe_sub<Fq>(fe_mul<Fq>(r, fe_sub<Fq>(qq, x3)), fe_mul<Fq>(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fe_mul<Fq>(fe_mul<Fq>(acc.zz, q.zz), pp);
    acc.zzz = fe_mul<Fq>(fe_mul<Fq>(acc.zzz, q.zzz), ppp);
}
// unique affine representative (halo2curves `to_affine`): one Fq inversion
H2V_HD affine xyzz_to_affine(const xyzz &p) {
    affine r;
    if (xyzz_is_identity(p)) {
        r.x = fe_zero(); r.y = fe_zero();
        return r;
    }
    // 1/zz = zzz^2 / zz^4 ... simpler: invert zz*zzz once:  i = 1/(zz*zzz); 1/zz = i*zzz; 1/zzz = i*zz
    fe i = fe_inv<Fq>(fe_mul<Fq>(p.zz, p.zzz));
    r.x = fe_mul<Fq>(p.x, fe_mul<Fq>(i, p.zzz));
    r.y = fe_mul<Fq>(p.y, fe_mul<Fq>(i, p.zz));
    return r;
}
// the same with the binary-Euclid inversion: for single-thread, latency-bound callers
H2V_HD affine xyzz_to_affine_fast(const xyzz &p) {
    affine r;
    if (xyzz_is_identity(p)) {
        r.x = fe_zero(); r.y = fe_zero();
        return r;
    }
    fe i = fe_inv_fast<Fq>(fe_mul<Fq>(p.zz, p.zzz));
    r.x = fe_mul<Fq>(p.x, fe_mul<Fq>(i, p.zzz));
    r.y = fe_mul<Fq>(p.y, fe_mul<Fq>(i, p.zz));
    return r;
}
// a Jacobian representative of the same point, no inversion: Z = ZZZ  =>  X' = X*ZZ^2, Y' = Y*ZZZ^2
H2V_HD jacobian xyzz_to_jacobian(const xyzz &p) {
    jacobian r;
    if (xyzz_is_identity(p)) {     // halo2curves `G1::identity()` is (0, 1, 0): keep the raw value byte-compatible
        r.x = fe_zero(); r.y = fe_one<Fq>(); r.z = fe_zero();
        return r;
    }
    r.x = fe_mul<Fq>(p.x, fe_sqr<Fq>(p.zz));
    r.y = fe_mul<Fq>(p.y, fe_sqr<Fq>(p.zzz));
    r.z = p.zzz;
    return r;
}

}  // namespace h2v
