// hostcheck.cu -- exposes the HOST instantiations of the __host__ __device__ field / curve routines
// (ff.cuh, ec.cuh) so that the CPU-only test suite can check the shared formulas (Montgomery
// constants, XYZZ group law, Jacobian conversion) against the oracle without a GPU.  The device
// instantiations of the same templates are checked on the GPU through h2v_selftest_*.
#include <string.h>
#include "ec.cuh"
using namespace h2v;
extern "C" {
void h2v_host_field(int field, int op, const uint64_t *a, const uint64_t *b, uint64_t *o) {
    fe x, y, r;
    memcpy(x.v, a, 32);
    if (b) memcpy(y.v, b, 32); else y = fe_zero();
    if (field == 0) r = op == 0 ? fe_mul<FrP>(x, y) : op == 1 ? fe_add<FrP>(x, y) : op == 2 ? fe_sub<FrP>(x, y) : op == 3 ? fe_inv<FrP>(x) : op == 4 ? fe_to_mont<FrP>(x) : op == 5 ? fe_from_mont<FrP>(x) : fe_inv_fast<FrP>(x);
    else r = op == 0 ? fe_mul<FqP>(x, y) : op == 1 ? fe_add<FqP>(x, y) : op == 2 ? fe_sub<FqP>(x, y) : op == 3 ? fe_inv<FqP>(x) : op == 4 ? fe_to_mont<FqP>(x) : op == 5 ? fe_from_mont<FqP>(x) : fe_inv_fast<FqP>(x);
    memcpy(o, r.v, 32);
}
// mode 0: mixed add p + q; 1: full add with q given a non-trivial ZZ; 2: double p; 3: p + q returned as Jacobian (12 limbs)
void h2v_host_group(int mode, const uint64_t *p, const uint64_t *q, uint64_t *o) {
    affine P, Q;
    memcpy(&P, p, 64);
    memcpy(&Q, q, 64);
    xyzz acc = xyzz_from_affine(P);
    if (mode == 0 || mode == 3) xyzz_add_mixed(acc, Q);
    else if (mode == 1) {
        xyzz t = xyzz_from_affine(Q);
        t = xyzz_double(t);
        xyzz_add_mixed(t, affine_neg(Q));
        xyzz_add(acc, t);
    } else acc = xyzz_double(acc);
    if (mode == 3) {
        // make ZZ non-trivial first so the conversion is exercised
        xyzz d = xyzz_double(acc);
        xyzz_add_mixed(d, affine_neg(xyzz_to_affine(acc)));
        jacobian j = xyzz_to_jacobian(d);
        memcpy(o, &j, 96);
    } else {
        affine r = xyzz_to_affine(acc);
        memcpy(o, &r, 64);
    }
}
}
