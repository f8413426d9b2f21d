/*
 * bn254_oracle.c -- CPU restatement of the reference's KZG-commit / EvaluationDomain hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (halo2_vectordb_b200/, include/) may
 * call, link or load this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker / CPU baseline.
 *
 * PARITY UNPINNED.  /root/reference holds no prover arithmetic and no golden vectors for this
 * boundary.  The algorithm lives in un-vendored, un-locked git dependencies:
 *   halo2-base  @ axiom-crypto/halo2-lib branch=community-edition   (/root/reference/Cargo.toml:22)
 *     -> halo2_proofs_axiom ("halo2-axiom", PSE halo2 v2023_02_02 lineage, Cargo.toml:19)
 *     -> halo2curves 0.3.x (bn256::{Fr,Fq,G1,G1Affine})
 * and there is no Rust toolchain in this image, so the reference cannot be compiled or run.
 * This file restates the *published* algorithms of those crates:
 *   arithmetic.rs      best_multiexp / multiexp_serial / best_fft / recursive_butterfly_arithmetic
 *   poly/domain.rs     EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended,
 *                      extended_to_coeff, divide_by_vanishing_poly}
 *   poly/kzg/commitment.rs  ParamsKZG::{commit, commit_lagrange}  (= best_multiexp on g / g_lagrange)
 * reached from the reference at src/scaffold/mod.rs:260 (gen_srs), :273 (create_pk),
 * :296 (gen_snark_shplonk).  It is pinned only against first-principles big-integer math
 * (oracle/pyref.py) and the constants / known answers of SURVEY.md App. B.
 *
 * Data layout = halo2curves: Fr/Fq are 4 little-endian u64 limbs in Montgomery form (R = 2^256),
 * G1Affine = {x,y} (64 B, identity = (0,0)), G1 = Jacobian {x,y,z} (96 B, identity z = 0).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

typedef struct { u64 l[4]; } fe;             /* field element, Montgomery form */
typedef struct { fe x, y; } g1a;             /* affine */
typedef struct { fe x, y, z; } g1j;          /* Jacobian */

typedef struct {
    u64 m[4];      /* modulus */
    u64 inv;       /* -m^{-1} mod 2^64 */
    u64 r1[4];     /* R mod m   (= one) */
    u64 r2[4];     /* R^2 mod m */
} field_t;

/* SURVEY.md App. B (re-derived in oracle/pyref.py and checked in tests/test_oracle.py) */
static const field_t FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x87d20782e4866389ULL,
    {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
    {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};
static const field_t FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0xc2e1f593efffffffULL,
    {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
    {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};

/* Fr::ROOT_OF_UNITY (order 2^28) and Fr::ZETA, canonical (non-Montgomery) limbs */
static const u64 FR_ROOT_OF_UNITY[4] = {0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL,
                                        0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL};
static const u64 FR_ZETA[4] = {0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL,
                               0x048b6e193fd84104ULL, 0x30644e72e131a029ULL};
#define FR_S 28

/* ------------------------------------------------------------------ field arithmetic */
static inline int ge_mod(const u64 a[4], const u64 m[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > m[i]) return 1;
        if (a[i] < m[i]) return 0;
    }
    return 1;
}
static inline void sub_mod_raw(u64 a[4], const u64 m[4]) {
    u128 b = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - m[i] - (u64)b;
        a[i] = (u64)d;
        b = (d >> 64) & 1;
    }
}
static inline void f_add(const field_t *F, fe *o, const fe *a, const fe *b) {
    u128 c = 0;
    u64 t[4];
    for (int i = 0; i < 4; ++i) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (u64)c;
        c >>= 64;
    }
    /* moduli are 254-bit, so no carry out of limb 3 */
    if (ge_mod(t, F->m)) sub_mod_raw(t, F->m);
    memcpy(o->l, t, 32);
}
static inline void f_sub(const field_t *F, fe *o, const fe *a, const fe *b) {
    u128 br = 0;
    u64 t[4];
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a->l[i] - b->l[i] - (u64)br;
        t[i] = (u64)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)t[i] + F->m[i];
            t[i] = (u64)c;
            c >>= 64;
        }
    }
    memcpy(o->l, t, 32);
}
static inline int f_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int f_eq(const fe *a, const fe *b) { return memcmp(a->l, b->l, 32) == 0; }
static inline void f_neg(const field_t *F, fe *o, const fe *a) {
    fe z = {{0, 0, 0, 0}};
    f_sub(F, o, &z, a);
}
static inline void f_dbl(const field_t *F, fe *o, const fe *a) { f_add(F, o, a, a); }

/* Montgomery multiplication, coarsely integrated operand scanning, 4x64 limbs */
static inline void f_mul(const field_t *F, fe *o, const fe *a, const fe *b) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (u64)c;
        t[5] = (u64)(c >> 64);
        u64 m = t[0] * F->inv;
        c = (u128)m * F->m[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * F->m[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (u64)c;
        t[4] = t[5] + (u64)(c >> 64);
    }
    if (t[4] || ge_mod(t, F->m)) sub_mod_raw(t, F->m);
    memcpy(o->l, t, 32);
}
static inline void f_sqr(const field_t *F, fe *o, const fe *a) { f_mul(F, o, a, a); }
static void f_to_mont(const field_t *F, fe *o, const u64 canon[4]) {
    fe a, r2;
    memcpy(a.l, canon, 32);
    memcpy(r2.l, F->r2, 32);
    f_mul(F, o, &a, &r2);
}
static void f_from_mont(const field_t *F, u64 canon[4], const fe *a) {
    fe one = {{1, 0, 0, 0}}, o;
    f_mul(F, &o, a, &one);
    memcpy(canon, o.l, 32);
}
static void f_pow(const field_t *F, fe *o, const fe *a, const u64 e[4]) {
    fe acc, base = *a;
    memcpy(acc.l, F->r1, 32);
    for (int i = 0; i < 256; ++i) {
        if ((e[i >> 6] >> (i & 63)) & 1) f_mul(F, &acc, &acc, &base);
        f_sqr(F, &base, &base);
    }
    *o = acc;
}
static void f_inv(const field_t *F, fe *o, const fe *a) {
    u64 e[4];
    memcpy(e, F->m, 32);
    e[0] -= 2; /* both moduli have low limb >= 2 */
    f_pow(F, o, a, e);
}
static void f_pow_u64(const field_t *F, fe *o, const fe *a, u64 e) {
    u64 ee[4] = {e, 0, 0, 0};
    f_pow(F, o, a, ee);
}

/* ------------------------------------------------------------------ G1: y^2 = x^3 + 3 */
static inline int j_is_identity(const g1j *p) { return f_is_zero(&p->z); }
static inline int a_is_identity(const g1a *p) { return f_is_zero(&p->x) && f_is_zero(&p->y); }
static inline void j_set_identity(g1j *p) { memset(p, 0, sizeof *p); }
static inline void j_from_affine(g1j *o, const g1a *a) {
    if (a_is_identity(a)) { j_set_identity(o); return; }
    o->x = a->x; o->y = a->y;
    memcpy(o->z.l, FQ.r1, 32);
}
static void j_double(g1j *o, const g1j *p) {       /* dbl-2009-l, a = 0 */
    if (j_is_identity(p)) { j_set_identity(o); return; }
    fe A, B, C, D, E, Fv, t, x3, y3, z3;
    f_sqr(&FQ, &A, &p->x);
    f_sqr(&FQ, &B, &p->y);
    f_sqr(&FQ, &C, &B);
    f_add(&FQ, &t, &p->x, &B);
    f_sqr(&FQ, &t, &t);
    f_sub(&FQ, &t, &t, &A);
    f_sub(&FQ, &t, &t, &C);
    f_dbl(&FQ, &D, &t);
    f_dbl(&FQ, &E, &A);
    f_add(&FQ, &E, &E, &A);
    f_sqr(&FQ, &Fv, &E);
    f_dbl(&FQ, &t, &D);
    f_sub(&FQ, &x3, &Fv, &t);
    f_sub(&FQ, &t, &D, &x3);
    f_mul(&FQ, &y3, &E, &t);
    f_dbl(&FQ, &t, &C); f_dbl(&FQ, &t, &t); f_dbl(&FQ, &t, &t);
    f_sub(&FQ, &y3, &y3, &t);
    f_mul(&FQ, &z3, &p->y, &p->z);
    f_dbl(&FQ, &z3, &z3);
    o->x = x3; o->y = y3; o->z = z3;
}
static void j_add_mixed(g1j *o, const g1j *p, const g1a *q) {  /* madd-2007-bl */
    if (a_is_identity(q)) { *o = *p; return; }
    if (j_is_identity(p)) { j_from_affine(o, q); return; }
    fe z1z1, u2, s2, h, hh, i, j, r, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z);
    f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s2, &q->y, &p->z);
    f_mul(&FQ, &s2, &s2, &z1z1);
    if (f_eq(&u2, &p->x)) {
        if (f_eq(&s2, &p->y)) { j_double(o, p); return; }
        j_set_identity(o); return;
    }
    f_sub(&FQ, &h, &u2, &p->x);
    f_sqr(&FQ, &hh, &h);
    f_dbl(&FQ, &i, &hh); f_dbl(&FQ, &i, &i);
    f_mul(&FQ, &j, &h, &i);
    f_sub(&FQ, &r, &s2, &p->y); f_dbl(&FQ, &r, &r);
    f_mul(&FQ, &v, &p->x, &i);
    f_sqr(&FQ, &x3, &r);
    f_sub(&FQ, &x3, &x3, &j);
    f_dbl(&FQ, &t, &v);
    f_sub(&FQ, &x3, &x3, &t);
    f_sub(&FQ, &t, &v, &x3);
    f_mul(&FQ, &y3, &r, &t);
    f_mul(&FQ, &t, &p->y, &j); f_dbl(&FQ, &t, &t);
    f_sub(&FQ, &y3, &y3, &t);
    f_add(&FQ, &z3, &p->z, &h);
    f_sqr(&FQ, &z3, &z3);
    f_sub(&FQ, &z3, &z3, &z1z1);
    f_sub(&FQ, &z3, &z3, &hh);
    o->x = x3; o->y = y3; o->z = z3;
}
static void j_add(g1j *o, const g1j *p, const g1j *q) {       /* add-2007-bl */
    if (j_is_identity(q)) { *o = *p; return; }
    if (j_is_identity(p)) { *o = *q; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t, x3, y3, z3;
    f_sqr(&FQ, &z1z1, &p->z);
    f_sqr(&FQ, &z2z2, &q->z);
    f_mul(&FQ, &u1, &p->x, &z2z2);
    f_mul(&FQ, &u2, &q->x, &z1z1);
    f_mul(&FQ, &s1, &p->y, &q->z); f_mul(&FQ, &s1, &s1, &z2z2);
    f_mul(&FQ, &s2, &q->y, &p->z); f_mul(&FQ, &s2, &s2, &z1z1);
    if (f_eq(&u1, &u2)) {
        if (f_eq(&s1, &s2)) { j_double(o, p); return; }
        j_set_identity(o); return;
    }
    f_sub(&FQ, &h, &u2, &u1);
    f_dbl(&FQ, &i, &h); f_sqr(&FQ, &i, &i);
    f_mul(&FQ, &j, &h, &i);
    f_sub(&FQ, &r, &s2, &s1); f_dbl(&FQ, &r, &r);
    f_mul(&FQ, &v, &u1, &i);
    f_sqr(&FQ, &x3, &r);
    f_sub(&FQ, &x3, &x3, &j);
    f_dbl(&FQ, &t, &v);
    f_sub(&FQ, &x3, &x3, &t);
    f_sub(&FQ, &t, &v, &x3);
    f_mul(&FQ, &y3, &r, &t);
    f_mul(&FQ, &t, &s1, &j); f_dbl(&FQ, &t, &t);
    f_sub(&FQ, &y3, &y3, &t);
    f_add(&FQ, &z3, &p->z, &q->z);
    f_sqr(&FQ, &z3, &z3);
    f_sub(&FQ, &z3, &z3, &z1z1);
    f_sub(&FQ, &z3, &z3, &z2z2);
    f_mul(&FQ, &z3, &z3, &h);
    o->x = x3; o->y = y3; o->z = z3;
}
static void j_to_affine(g1a *o, const g1j *p) {
    if (j_is_identity(p)) { memset(o, 0, sizeof *o); return; }
    fe zi, zi2, zi3;
    f_inv(&FQ, &zi, &p->z);
    f_sqr(&FQ, &zi2, &zi);
    f_mul(&FQ, &zi3, &zi2, &zi);
    f_mul(&FQ, &o->x, &p->x, &zi2);
    f_mul(&FQ, &o->y, &p->y, &zi3);
}
/* scalar given as canonical 4x64 limbs */
static void j_mul_canon(g1j *o, const g1a *base, const u64 k[4]) {
    g1j acc;
    j_set_identity(&acc);
    for (int i = 255; i >= 0; --i) {
        j_double(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) j_add_mixed(&acc, &acc, base);
    }
    *o = acc;
}

/* ------------------------------------------------------------------ multiexp (App. A.3) */
/* c-bit digit of a canonical little-endian 32-byte scalar at bit offset seg*c (bytes >= 32 read as 0) */
static inline u64 get_at(unsigned seg, unsigned c, const uint8_t bytes[32]) {
    unsigned skip_bits = seg * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0};
    unsigned avail = 32 - skip_bytes;
    memcpy(v, bytes + skip_bytes, avail < 8 ? avail : 8);
    u64 tmp;
    memcpy(&tmp, v, 8); /* little-endian host */
    tmp >>= (skip_bits - skip_bytes * 8);
    return tmp % (1ULL << c);
}
/* bucket states of the reference: None / Affine / Projective */
typedef struct { int st; g1a a; g1j j; } bucket_t;

static void multiexp_serial(const fe *coeffs, const g1a *bases, size_t n, g1j *acc) {
    uint8_t (*repr)[32] = malloc(n ? n * 32 : 32);
    for (size_t i = 0; i < n; ++i) {
        u64 c4[4];
        f_from_mont(&FR, c4, &coeffs[i]);      /* to_repr(): canonical LE bytes */
        memcpy(repr[i], c4, 32);
    }
    unsigned c;
    if (n < 4) c = 1;
    else if (n < 32) c = 3;
    else c = (unsigned)ceil(log((double)n));
    unsigned segments = 256 / c + 1;
    size_t nb = ((size_t)1 << c) - 1;
    bucket_t *buckets = malloc(nb * sizeof(bucket_t));
    for (int seg = (int)segments - 1; seg >= 0; --seg) {
        for (unsigned k = 0; k < c; ++k) j_double(acc, acc);
        for (size_t b = 0; b < nb; ++b) buckets[b].st = 0;
        for (size_t i = 0; i < n; ++i) {
            u64 d = get_at((unsigned)seg, c, repr[i]);
            if (d == 0) continue;
            bucket_t *bk = &buckets[d - 1];
            if (bk->st == 0) { bk->st = 1; bk->a = bases[i]; }
            else if (bk->st == 1) { j_from_affine(&bk->j, &bk->a); j_add_mixed(&bk->j, &bk->j, &bases[i]); bk->st = 2; }
            else j_add_mixed(&bk->j, &bk->j, &bases[i]);
        }
        /* running sum: sum_d d * bucket_d */
        g1j running;
        j_set_identity(&running);
        for (size_t b = nb; b-- > 0;) {
            if (buckets[b].st == 1) j_add_mixed(&running, &running, &buckets[b].a);
            else if (buckets[b].st == 2) j_add(&running, &running, &buckets[b].j);
            j_add(acc, acc, &running);
        }
    }
    free(buckets);
    free(repr);
}

typedef struct { const fe *c; const g1a *b; size_t n; g1j acc; } msm_job;
static void *msm_worker(void *p) {
    msm_job *j = p;
    j_set_identity(&j->acc);
    multiexp_serial(j->c, j->b, j->n, &j->acc);
    return NULL;
}

/* best_multiexp(coeffs, bases) with `threads` playing rayon's current_num_threads() */
void orc_best_multiexp(const u64 *coeffs, const u64 *bases, size_t n, int threads, u64 out_jac[12]) {
    const fe *c = (const fe *)coeffs;
    const g1a *b = (const g1a *)bases;
    g1j acc;
    j_set_identity(&acc);
    if (threads < 1) threads = 1;
    if (n > (size_t)threads) {
        size_t chunk = n / (size_t)threads;
        size_t nchunks = (n + chunk - 1) / chunk;
        msm_job *jobs = malloc(nchunks * sizeof(msm_job));
        pthread_t *th = malloc(nchunks * sizeof(pthread_t));
        for (size_t k = 0; k < nchunks; ++k) {
            size_t lo = k * chunk, hi = lo + chunk > n ? n : lo + chunk;
            jobs[k].c = c + lo; jobs[k].b = b + lo; jobs[k].n = hi - lo;
            if (threads == 1) msm_worker(&jobs[k]);
            else pthread_create(&th[k], NULL, msm_worker, &jobs[k]);
        }
        for (size_t k = 0; k < nchunks; ++k) {
            if (threads != 1) pthread_join(th[k], NULL);
            j_add(&acc, &acc, &jobs[k].acc);
        }
        free(jobs); free(th);
    } else {
        multiexp_serial(c, b, n, &acc);
    }
    memcpy(out_jac, &acc, 96);
}

void orc_g1_to_affine(const u64 jac[12], u64 aff[8]) { j_to_affine((g1a *)aff, (const g1j *)jac); }
void orc_g1_add(const u64 a[12], const u64 b[12], u64 o[12]) { g1j t; j_add(&t, (const g1j *)a, (const g1j *)b); memcpy(o, &t, 96); }
void orc_g1_double(const u64 a[12], u64 o[12]) { g1j t; j_double(&t, (const g1j *)a); memcpy(o, &t, 96); }
void orc_g1_add_mixed(const u64 a[12], const u64 b[8], u64 o[12]) { g1j t; j_add_mixed(&t, (const g1j *)a, (const g1a *)b); memcpy(o, &t, 96); }
/* k canonical (non-Montgomery) limbs */
void orc_g1_mul(const u64 base_aff[8], const u64 k[4], u64 out_aff[8]) {
    g1j t;
    j_mul_canon(&t, (const g1a *)base_aff, k);
    j_to_affine((g1a *)out_aff, &t);
}
void orc_g1_generator(u64 out_aff[8]) {
    g1a g;
    u64 one[4] = {1, 0, 0, 0}, two[4] = {2, 0, 0, 0};
    f_to_mont(&FQ, &g.x, one);
    f_to_mont(&FQ, &g.y, two);
    memcpy(out_aff, &g, 64);
}
int orc_g1_is_on_curve(const u64 aff[8]) {
    const g1a *p = (const g1a *)aff;
    if (a_is_identity(p)) return 1;
    fe y2, x3, three;
    u64 t[4] = {3, 0, 0, 0};
    f_to_mont(&FQ, &three, t);
    f_sqr(&FQ, &y2, &p->y);
    f_sqr(&FQ, &x3, &p->x); f_mul(&FQ, &x3, &x3, &p->x);
    f_add(&FQ, &x3, &x3, &three);
    return f_eq(&y2, &x3);
}
/* halo2curves 0.3.x to_bytes(): canonical x LE, top bit of byte 31 = lsb(canonical y); identity = zeros (recalled; >= 0.4 uses bit 6) */
void orc_g1_compress(const u64 aff[8], uint8_t out[32]) {
    const g1a *p = (const g1a *)aff;
    if (a_is_identity(p)) { memset(out, 0, 32); return; }
    u64 x[4], y[4];
    f_from_mont(&FQ, x, &p->x);
    f_from_mont(&FQ, y, &p->y);
    memcpy(out, x, 32);
    out[31] |= (uint8_t)((y[0] & 1) << 7);
}

/* ------------------------------------------------------------------ synthetic inputs */
/* bases[i] = (a*i + b) * G, affine Montgomery; generated by repeated addition + batch normalise */
typedef struct { u64 a, b; size_t lo, hi; g1a *out; } bases_job;
static void *bases_worker(void *p) {
    bases_job *j = p;
    size_t n = j->hi - j->lo;
    if (!n) return NULL;
    g1a g, step_a;
    orc_g1_generator((u64 *)&g);
    g1j step, cur;
    u64 ka[4] = {j->a, 0, 0, 0};
    j_mul_canon(&step, &g, ka);
    j_to_affine(&step_a, &step);
    /* start = (a*lo + b) G, a*lo+b as a 128-bit integer (lo < 2^32, a,b < 2^63) */
    u128 k0 = (u128)j->a * j->lo + j->b;
    u64 kk[4] = {(u64)k0, (u64)(k0 >> 64), 0, 0};
    j_mul_canon(&cur, &g, kk);
    g1j *tmp = malloc(n * sizeof(g1j));
    for (size_t i = 0; i < n; ++i) {
        tmp[i] = cur;
        j_add_mixed(&cur, &cur, &step_a);
    }
    /* batch inversion of z (Montgomery's trick) */
    fe *pref = malloc(n * sizeof(fe));
    fe acc;
    memcpy(acc.l, FQ.r1, 32);
    for (size_t i = 0; i < n; ++i) {
        pref[i] = acc;
        if (!j_is_identity(&tmp[i])) f_mul(&FQ, &acc, &acc, &tmp[i].z);
    }
    fe inv;
    f_inv(&FQ, &inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (j_is_identity(&tmp[i])) { memset(&j->out[j->lo + i], 0, 64); continue; }
        fe zi, zi2, zi3;
        f_mul(&FQ, &zi, &inv, &pref[i]);
        f_mul(&FQ, &inv, &inv, &tmp[i].z);
        f_sqr(&FQ, &zi2, &zi);
        f_mul(&FQ, &zi3, &zi2, &zi);
        f_mul(&FQ, &j->out[j->lo + i].x, &tmp[i].x, &zi2);
        f_mul(&FQ, &j->out[j->lo + i].y, &tmp[i].y, &zi3);
    }
    free(pref); free(tmp);
    return NULL;
}
void orc_gen_bases(u64 a, u64 b, size_t n, int threads, u64 *out) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    bases_job *jobs = malloc(threads * sizeof(bases_job));
    pthread_t *th = malloc(threads * sizeof(pthread_t));
    size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = (size_t)t * per, hi = lo + per > n ? n : lo + per;
        if (lo > n) lo = n;
        jobs[t] = (bases_job){a, b, lo, hi, (g1a *)out};
        pthread_create(&th[t], NULL, bases_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(jobs); free(th);
}

static inline u64 splitmix64(u64 *s) {
    u64 z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* mode 0: uniform in [0,r) (254-bit draw, one conditional subtract)
 * mode 1: witness-like: 60% zero/one, 30% < 2^lookup_bits, 10% full width (half of them r - small)
 * output Montgomery form.  Element i depends only on (seed, i), so any slice can be regenerated. */
void orc_fr_fill(u64 seed, int mode, unsigned lookup_bits, size_t n, u64 *out) {
    for (size_t i = 0; i < n; ++i) {
        u64 s = seed ^ (0xD1B54A32D192ED03ULL * (u64)(i + 1));
        u64 c[4] = {0, 0, 0, 0};
        u64 sel = splitmix64(&s) % 100;
        if (mode == 0 || sel >= 90) {
            for (int k = 0; k < 4; ++k) c[k] = splitmix64(&s);
            c[3] &= 0x3FFFFFFFFFFFFFFFULL;
            if (ge_mod(c, FR.m)) sub_mod_raw(c, FR.m);
            if (mode == 1 && sel >= 95) { /* "negative" fixed-point value: r - small */
                u64 small[4] = {splitmix64(&s) >> 16, 0, 0, 0};
                memcpy(c, FR.m, 32);
                sub_mod_raw(c, small);
            }
        } else if (sel < 60) {
            c[0] = splitmix64(&s) & 1;
        } else {
            c[0] = splitmix64(&s) & ((1ULL << lookup_bits) - 1);
        }
        f_to_mont(&FR, (fe *)(out + 4 * i), c);
    }
}

/* ------------------------------------------------------------------ Fr helpers for tests */
void orc_fr_mul(const u64 a[4], const u64 b[4], u64 o[4]) { f_mul(&FR, (fe *)o, (const fe *)a, (const fe *)b); }
void orc_fr_add(const u64 a[4], const u64 b[4], u64 o[4]) { f_add(&FR, (fe *)o, (const fe *)a, (const fe *)b); }
void orc_fr_sub(const u64 a[4], const u64 b[4], u64 o[4]) { f_sub(&FR, (fe *)o, (const fe *)a, (const fe *)b); }
void orc_fr_inv(const u64 a[4], u64 o[4]) { f_inv(&FR, (fe *)o, (const fe *)a); }
void orc_fq_mul(const u64 a[4], const u64 b[4], u64 o[4]) { f_mul(&FQ, (fe *)o, (const fe *)a, (const fe *)b); }
void orc_fq_add(const u64 a[4], const u64 b[4], u64 o[4]) { f_add(&FQ, (fe *)o, (const fe *)a, (const fe *)b); }
void orc_fq_sub(const u64 a[4], const u64 b[4], u64 o[4]) { f_sub(&FQ, (fe *)o, (const fe *)a, (const fe *)b); }
void orc_fq_inv(const u64 a[4], u64 o[4]) { f_inv(&FQ, (fe *)o, (const fe *)a); }
void orc_to_mont(int field, const u64 *canon, size_t n, u64 *out) {
    const field_t *F = field ? &FQ : &FR;
    for (size_t i = 0; i < n; ++i) f_to_mont(F, (fe *)(out + 4 * i), canon + 4 * i);
}
void orc_from_mont(int field, const u64 *mont, size_t n, u64 *out) {
    const field_t *F = field ? &FQ : &FR;
    for (size_t i = 0; i < n; ++i) f_from_mont(F, out + 4 * i, (const fe *)(mont + 4 * i));
}
/* sum_i s_i * (a*i + b) mod r  -- the closed-form MSM check of SURVEY.md 8c(3); returns canonical limbs */
void orc_fr_dot_affine_index(const u64 *scalars_mont, size_t n, u64 a, u64 b, u64 out_canon[4]) {
    fe acc = {{0, 0, 0, 0}};
    for (size_t i = 0; i < n; ++i) {
        u128 k = (u128)a * i + b;
        u64 kc[4] = {(u64)k, (u64)(k >> 64), 0, 0};
        fe km, t;
        f_to_mont(&FR, &km, kc);
        f_mul(&FR, &t, &km, (const fe *)(scalars_mont + 4 * i));
        f_add(&FR, &acc, &acc, &t);
    }
    f_from_mont(&FR, out_canon, &acc);
}
/* Horner evaluation of a Montgomery-form coefficient vector at Montgomery x */
void orc_fr_eval_poly(const u64 *a, size_t n, const u64 x[4], u64 out[4]) {
    fe acc = {{0, 0, 0, 0}};
    for (size_t i = n; i-- > 0;) {
        f_mul(&FR, &acc, &acc, (const fe *)x);
        f_add(&FR, &acc, &acc, (const fe *)(a + 4 * i));
    }
    memcpy(out, &acc, 32);
}

/* ------------------------------------------------------------------ best_fft (App. A.4) */
static inline size_t bitreverse(size_t n, unsigned l) {
    size_t r = 0;
    for (unsigned i = 0; i < l; ++i) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
static inline void butterfly_one(fe *a, fe *b) {      /* twiddle == 1 */
    fe t = *b;
    *b = *a;
    f_add(&FR, a, a, &t);
    f_sub(&FR, b, b, &t);
}
static inline void butterfly_tw(fe *a, fe *b, const fe *w) {
    fe t;
    f_mul(&FR, &t, b, w);
    *b = *a;
    f_add(&FR, a, a, &t);
    f_sub(&FR, b, b, &t);
}
typedef struct { fe *a; size_t n; size_t twiddle_chunk; const fe *tw; int depth; } rec_job;
static void recursive_butterfly(fe *a, size_t n, size_t twiddle_chunk, const fe *tw, int par_depth);
static void *rec_worker(void *p) {
    rec_job *j = p;
    recursive_butterfly(j->a, j->n, j->twiddle_chunk, j->tw, j->depth);
    return NULL;
}
/* recursive_butterfly_arithmetic: recurse on both halves (rayon::join in the reference; a thread
 * per half for the top `par_depth` levels here), then one serial combine loop over n/2 butterflies. */
static void recursive_butterfly(fe *a, size_t n, size_t twiddle_chunk, const fe *tw, int par_depth) {
    if (n == 2) { butterfly_one(&a[0], &a[1]); return; }
    fe *left = a, *right = a + n / 2;
    if (par_depth > 0) {
        pthread_t th;
        rec_job j = {left, n / 2, twiddle_chunk * 2, tw, par_depth - 1};
        pthread_create(&th, NULL, rec_worker, &j);
        recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, par_depth - 1);
        pthread_join(th, NULL);
    } else {
        recursive_butterfly(left, n / 2, twiddle_chunk * 2, tw, 0);
        recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, 0);
    }
    butterfly_one(&left[0], &right[0]);
    for (size_t i = 1; i < n / 2; ++i) butterfly_tw(&left[i], &right[i], &tw[i * twiddle_chunk]);
}
static unsigned log2_floor(unsigned v) { unsigned l = 0; while ((1u << (l + 1)) <= v) ++l; return l; }

static void best_fft_fe(fe *a, const fe *omega, unsigned log_n, int threads) {
    size_t n = (size_t)1 << log_n;
    if (threads < 1) threads = 1;
    unsigned log_threads = log2_floor((unsigned)threads);
    for (size_t k = 0; k < n; ++k) {
        size_t rk = bitreverse(k, log_n);
        if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    if (n < 2) return;
    /* twiddles rebuilt on every call, serial scan (as in the reference) */
    fe *tw = malloc((n / 2) * sizeof(fe));
    fe w;
    memcpy(w.l, FR.r1, 32);
    for (size_t i = 0; i < n / 2; ++i) { tw[i] = w; f_mul(&FR, &w, &w, omega); }
    if (log_n <= log_threads) {
        size_t chunk = 2, twiddle_chunk = n / 2;
        for (unsigned s = 0; s < log_n; ++s) {
            for (size_t base = 0; base < n; base += chunk) {
                fe *left = a + base, *right = a + base + chunk / 2;
                butterfly_one(&left[0], &right[0]);
                for (size_t i = 1; i < chunk / 2; ++i) butterfly_tw(&left[i], &right[i], &tw[i * twiddle_chunk]);
            }
            chunk *= 2;
            twiddle_chunk /= 2;
        }
    } else {
        recursive_butterfly(a, n, 1, tw, (int)log_threads);
    }
    free(tw);
}
void orc_best_fft(u64 *a, const u64 omega[4], uint32_t log_n, int threads) {
    best_fft_fe((fe *)a, (const fe *)omega, log_n, threads);
}

/* ------------------------------------------------------------------ EvaluationDomain (App. A.5) */
typedef struct {
    uint32_t k, extended_k, quotient_poly_degree;
    size_t n;
    fe omega, omega_inv, ext_omega, ext_omega_inv;
    fe g_coset, g_coset_inv;
    fe ifft_divisor, ext_ifft_divisor;
    fe t_evaluations[64];
    uint32_t n_t;
} orc_domain;

int orc_domain_new(uint32_t j, uint32_t k, orc_domain *d) {
    memset(d, 0, sizeof *d);
    d->k = k;
    d->n = (size_t)1 << k;
    d->quotient_poly_degree = j - 1;
    uint32_t ek = k;
    while (((size_t)1 << ek) < d->n * (j - 1)) ++ek;
    if (ek > FR_S || ek - k > 6) return -1;
    d->extended_k = ek;
    fe root;
    f_to_mont(&FR, &root, FR_ROOT_OF_UNITY);
    d->ext_omega = root;
    for (uint32_t i = ek; i < FR_S; ++i) f_sqr(&FR, &d->ext_omega, &d->ext_omega);
    d->omega = d->ext_omega;
    for (uint32_t i = k; i < ek; ++i) f_sqr(&FR, &d->omega, &d->omega);
    f_inv(&FR, &d->omega_inv, &d->omega);
    f_inv(&FR, &d->ext_omega_inv, &d->ext_omega);
    f_to_mont(&FR, &d->g_coset, FR_ZETA);
    f_sqr(&FR, &d->g_coset_inv, &d->g_coset);
    fe two_k, t;
    u64 c[4] = {0, 0, 0, 0};
    c[0] = 1ULL << k;
    f_to_mont(&FR, &two_k, c);
    f_inv(&FR, &d->ifft_divisor, &two_k);
    c[0] = 1ULL << ek;
    f_to_mont(&FR, &t, c);
    f_inv(&FR, &d->ext_ifft_divisor, &t);
    d->n_t = 1u << (ek - k);
    fe cur = d->g_coset, one;
    memcpy(one.l, FR.r1, 32);
    for (uint32_t i = 0; i < d->n_t; ++i) {
        fe v;
        f_pow_u64(&FR, &v, &cur, (u64)d->n);
        f_sub(&FR, &v, &v, &one);
        f_inv(&FR, &d->t_evaluations[i], &v);
        f_mul(&FR, &cur, &cur, &d->ext_omega);
    }
    return 0;
}
size_t orc_domain_sizeof(void) { return sizeof(orc_domain); }
void orc_domain_get(const orc_domain *d, int which, u64 out[4]) {
    const fe *p = which == 0 ? &d->omega : which == 1 ? &d->omega_inv : which == 2 ? &d->ext_omega
                : which == 3 ? &d->ext_omega_inv : which == 4 ? &d->g_coset : which == 5 ? &d->g_coset_inv
                : which == 6 ? &d->ifft_divisor : which == 7 ? &d->ext_ifft_divisor
                : &d->t_evaluations[which - 8];
    memcpy(out, p, 32);
}
uint32_t orc_domain_extended_k(const orc_domain *d) { return d->extended_k; }

typedef struct { fe *a; size_t lo, hi; const fe *m; } scale_job;
static void *scale_worker(void *p) {
    scale_job *j = p;
    for (size_t i = j->lo; i < j->hi; ++i) f_mul(&FR, &j->a[i], &j->a[i], j->m);
    return NULL;
}
static void par_scale(fe *a, size_t n, const fe *m, int threads) {
    if (threads <= 1 || n < 1024) { scale_job j = {a, 0, n, m}; scale_worker(&j); return; }
    pthread_t th[256]; scale_job jobs[256];
    if (threads > 256) threads = 256;
    size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = t * per, hi = lo + per > n ? n : lo + per;
        if (lo > n) lo = n;
        jobs[t] = (scale_job){a, lo, hi, m};
        pthread_create(&th[t], NULL, scale_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}
/* ifft(a, omega_inv, log_n, divisor) */
static void ifft(fe *a, const fe *omega_inv, unsigned log_n, const fe *divisor, int threads) {
    best_fft_fe(a, omega_inv, log_n, threads);
    par_scale(a, (size_t)1 << log_n, divisor, threads);
}
/* distribute_powers_zeta: a[i] *= [1, z, z^2][i % 3], z = g_coset (into_coset) or g_coset_inv */
static void distribute_powers_zeta(const orc_domain *d, fe *a, size_t n, int into_coset) {
    const fe *c1 = into_coset ? &d->g_coset : &d->g_coset_inv;
    const fe *c2 = into_coset ? &d->g_coset_inv : &d->g_coset;
    for (size_t i = 0; i < n; ++i) {
        if (i % 3 == 1) f_mul(&FR, &a[i], &a[i], c1);
        else if (i % 3 == 2) f_mul(&FR, &a[i], &a[i], c2);
    }
}
void orc_lagrange_to_coeff(const orc_domain *d, u64 *a, int threads) {
    ifft((fe *)a, &d->omega_inv, d->k, &d->ifft_divisor, threads);
}
void orc_coeff_to_lagrange(const orc_domain *d, u64 *a, int threads) {
    best_fft_fe((fe *)a, &d->omega, d->k, threads);
}
/* in: n elements, out: 2^extended_k elements */
void orc_coeff_to_extended(const orc_domain *d, const u64 *in, u64 *out, int threads) {
    size_t en = (size_t)1 << d->extended_k;
    memcpy(out, in, d->n * 32);
    distribute_powers_zeta(d, (fe *)out, d->n, 1);
    memset(out + 4 * d->n, 0, (en - d->n) * 32);
    best_fft_fe((fe *)out, &d->ext_omega, d->extended_k, threads);
}
/* in: 2^extended_k elements (clobbered), out: n*(j-1) elements */
void orc_extended_to_coeff(const orc_domain *d, u64 *in, u64 *out, int threads) {
    size_t en = (size_t)1 << d->extended_k;
    ifft((fe *)in, &d->ext_omega_inv, d->extended_k, &d->ext_ifft_divisor, threads);
    distribute_powers_zeta(d, (fe *)in, en, 0);
    memcpy(out, in, d->n * d->quotient_poly_degree * 32);
}
void orc_divide_by_vanishing_poly(const orc_domain *d, u64 *a) {
    size_t en = (size_t)1 << d->extended_k;
    fe *p = (fe *)a;
    for (size_t i = 0; i < en; ++i) f_mul(&FR, &p[i], &p[i], &d->t_evaluations[i % d->n_t]);
}

/* ------------------------------------------------------------------ "next" row 2: the callers
 * on either side of every commit, restated from halo2-axiom arithmetic.rs / plonk/permutation/prover.rs:
 *   eval_polynomial(poly, point)        Horner evaluation                          -> orc_fr_eval_poly (above)
 *   BatchInvert::batch_invert           every non-zero element replaced by its inverse, zeros untouched
 *   grand product z                      z[0] = 1, z[i+1] = z[i] * num[i] / den[i]   (the running product of
 *                                        the permutation / lookup arguments, before blinding)
 *   kate_division(a, b)                  quotient of a(X) by (X - b), remainder dropped                     */
void orc_fr_batch_invert(u64 *a, size_t n) {
    fe *p = (fe *)a;
    fe *pre = malloc((n ? n : 1) * sizeof(fe));
    fe acc;
    memcpy(acc.l, FR.r1, 32);
    for (size_t i = 0; i < n; ++i) {
        pre[i] = acc;
        if (!f_is_zero(&p[i])) f_mul(&FR, &acc, &acc, &p[i]);
    }
    fe inv;
    f_inv(&FR, &inv, &acc);
    for (size_t i = n; i-- > 0;) {
        if (f_is_zero(&p[i])) continue;
        fe t;
        f_mul(&FR, &t, &inv, &pre[i]);
        f_mul(&FR, &inv, &inv, &p[i]);
        p[i] = t;
    }
    free(pre);
}
/* out[0] = 1, out[i+1] = out[i] * num[i] / den[i], i < n - 1   (den[i] != 0) */
void orc_fr_grand_product(const u64 *num, const u64 *den, size_t n, u64 *out) {
    if (!n) return;
    fe *d = malloc(n * sizeof(fe));
    memcpy(d, den, n * 32);
    orc_fr_batch_invert((u64 *)d, n);
    fe *o = (fe *)out;
    memcpy(o[0].l, FR.r1, 32);
    for (size_t i = 0; i + 1 < n; ++i) {
        fe t;
        f_mul(&FR, &t, (const fe *)(num + 4 * i), &d[i]);
        f_mul(&FR, &o[i + 1], &o[i], &t);
    }
    free(d);
}
/* a has n coefficients (n >= 1); out has n - 1:  q[n-2] = a[n-1], q[i-1] = a[i] + b * q[i] */
void orc_fr_kate_division(const u64 *a, size_t n, const u64 b[4], u64 *out) {
    const fe *p = (const fe *)a;
    fe *q = (fe *)out;
    fe tmp = {{0, 0, 0, 0}};
    for (size_t i = n; i-- > 1;) {
        fe t;
        f_mul(&FR, &t, &tmp, (const fe *)b);
        f_add(&FR, &tmp, &p[i], &t);
        q[i - 1] = tmp;
    }
}

/* ------------------------------------------------------------------ ParamsKZG::setup
 * halo2-axiom poly/kzg/commitment.rs `ParamsKZG::setup(k, rng)`: s = Fr::random(rng) (here: supplied by the
 * caller, Montgomery form), g[i] = s^i * G, and the Lagrange basis directly from s:
 *   g_lagrange[i] = ((s^n - 1) / n) * w^i / (s - w^i) * G,   w = the 2^k-th root of unity of EvaluationDomain.
 * (g2 / s_g2 belong to the verifier and are not on the commit path.) */
typedef struct { const fe *sc; size_t lo, hi; g1a *out; } fixmul_job;
static void *fixmul_worker(void *p) {
    fixmul_job *j = p;
    g1a g;
    orc_g1_generator((u64 *)&g);
    for (size_t i = j->lo; i < j->hi; ++i) {
        u64 c[4];
        f_from_mont(&FR, c, &j->sc[i]);
        g1j t;
        j_mul_canon(&t, &g, c);
        j_to_affine(&j->out[i], &t);
    }
    return NULL;
}
static void fixed_base_mul(const fe *sc, size_t n, int threads, g1a *out) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; fixmul_job jobs[256];
    size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        size_t lo = (size_t)t * per, hi = lo + per > n ? n : lo + per;
        if (lo > n) lo = n;
        jobs[t] = (fixmul_job){sc, lo, hi, out};
        pthread_create(&th[t], NULL, fixmul_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
}
int orc_srs_setup(uint32_t k, const u64 s_mont[4], int threads, u64 *g_out, u64 *g_lagrange_out) {
    if (k > FR_S) return -1;
    size_t n = (size_t)1 << k;
    const fe *s = (const fe *)s_mont;
    fe *sc = malloc(n * sizeof(fe));
    fe one, cur;
    memcpy(one.l, FR.r1, 32);
    cur = one;
    for (size_t i = 0; i < n; ++i) { sc[i] = cur; f_mul(&FR, &cur, &cur, s); }
    fixed_base_mul(sc, n, threads, (g1a *)g_out);
    /* cur == s^n now */
    fe omega, nf, ninv, mult, w;
    f_to_mont(&FR, &omega, FR_ROOT_OF_UNITY);
    for (uint32_t i = k; i < FR_S; ++i) f_sqr(&FR, &omega, &omega);
    u64 c[4] = {(u64)n, 0, 0, 0};
    f_to_mont(&FR, &nf, c);
    f_inv(&FR, &ninv, &nf);
    f_sub(&FR, &mult, &cur, &one);
    f_mul(&FR, &mult, &mult, &ninv);
    w = one;
    for (size_t i = 0; i < n; ++i) {
        fe d, di;
        f_sub(&FR, &d, s, &w);
        f_inv(&FR, &di, &d);            /* s is not a root of unity for any honest setup */
        f_mul(&FR, &sc[i], &mult, &w);
        f_mul(&FR, &sc[i], &sc[i], &di);
        f_mul(&FR, &w, &w, &omega);
    }
    fixed_base_mul(sc, n, threads, (g1a *)g_lagrange_out);
    free(sc);
    return 0;
}

/* ---- quotient evaluation row loops ("next" row 1) ----------------------------------------------------------------
 * Restates halo2-axiom plonk/evaluation.rs Evaluator::evaluate_h [UPSTREAM, absent from /root/reference; reached from
 * src/scaffold/mod.rs:296], for the constraint system halo2-base builds.  Written from the published PLONKish argument
 * (halo2 book: permutation and lookup arguments) and the recalled structure of that function: `values[idx] = values[idx]
 * * y + term` per term, custom gates first, then the permutation argument, then each lookup; rotation r at row idx reads
 * row (idx + r * 2^(extended_k - k)) mod 2^extended_k (get_rotation_idx).  PARITY UNPINNED like the rest of this file;
 * tests additionally check the defining property (the folded expression vanishes on the 2^k domain for a satisfied
 * circuit, and h(x) * (x^n - 1) equals the expression re-evaluated at a random point with big integers). */
static inline size_t rot_idx(size_t idx, int r, size_t rot_scale, size_t n_ext) {
    return (idx + (size_t)((long)r * (long)rot_scale + (long)n_ext * 64)) & (n_ext - 1);
}
static inline void fold(fe *v, const fe *y, const fe *term) {
    f_mul(&FR, v, v, y);
    f_add(&FR, v, v, term);
}
/* custom gates of halo2-base's FlexGateConfig: for advice column j,  q_j * (a + b * c - d)  with a, b, c, d = the
 * column at rotations 0, 1, 2, 3 */
void orc_quotient_gates(const orc_domain *d, u64 *h_, const u64 y_[4], size_t n_gates, const u64 *q_, size_t q_stride,
                        const u64 *a_, size_t a_stride) {
    const size_t n_ext = (size_t)1 << d->extended_k, rs = (size_t)1 << (d->extended_k - d->k);
    fe *h = (fe *)h_;
    const fe *y = (const fe *)y_, *q = (const fe *)q_, *a = (const fe *)a_;
    for (size_t idx = 0; idx < n_ext; ++idx)
        for (size_t j = 0; j < n_gates; ++j) {
            const fe *col = a + j * a_stride;
            fe t;
            f_mul(&FR, &t, &col[rot_idx(idx, 1, rs, n_ext)], &col[rot_idx(idx, 2, rs, n_ext)]);
            f_add(&FR, &t, &col[idx], &t);
            f_sub(&FR, &t, &t, &col[rot_idx(idx, 3, rs, n_ext)]);
            f_mul(&FR, &t, &q[j * q_stride + idx], &t);
            fold(&h[idx], y, &t);
        }
}
void orc_fr_delta(u64 out[4]) {     /* Fr::DELTA = MULTIPLICATIVE_GENERATOR^(2^S) = 7^(2^28) */
    u64 seven[4] = {7, 0, 0, 0};
    fe g;
    f_to_mont(&FR, &g, seven);
    f_pow_u64(&FR, (fe *)out, &g, (u64)1 << FR_S);
}
void orc_quotient_permutation(const orc_domain *d, u64 *h_, const u64 y_[4], const u64 beta_[4], const u64 gamma_[4],
                              size_t n_cols, size_t chunk_len, const u64 *cols_, size_t cols_stride, const u64 *sigma_,
                              size_t sigma_stride, const u64 *z_, size_t z_stride, const u64 *l0_, const u64 *l_last_,
                              const u64 *l_active_, uint32_t blinding_factors) {
    if (!n_cols) return;
    const size_t n_ext = (size_t)1 << d->extended_k, rs = (size_t)1 << (d->extended_k - d->k);
    const size_t n_sets = (n_cols + chunk_len - 1) / chunk_len;
    const int last_rotation = -(int)(blinding_factors + 1);
    fe *h = (fe *)h_;
    const fe *y = (const fe *)y_, *beta = (const fe *)beta_, *gamma = (const fe *)gamma_;
    const fe *cols = (const fe *)cols_, *sigma = (const fe *)sigma_, *z = (const fe *)z_;
    const fe *l0 = (const fe *)l0_, *l_last = (const fe *)l_last_, *l_active = (const fe *)l_active_;
    fe one, delta, delta_start, beta_term;
    memcpy(one.l, FR.r1, 32);
    orc_fr_delta(delta.l);
    f_mul(&FR, &delta_start, beta, &d->g_coset);
    beta_term = one;                                    /* extended_omega^idx */
    for (size_t idx = 0; idx < n_ext; ++idx) {
        const size_t r_next = rot_idx(idx, 1, rs, n_ext), r_last = rot_idx(idx, last_rotation, rs, n_ext);
        fe t, u;
        /* l_0(X) * (1 - z_0(X)) */
        f_sub(&FR, &t, &one, &z[idx]);
        f_mul(&FR, &t, &t, &l0[idx]);
        fold(&h[idx], y, &t);
        /* l_last(X) * (z_l(X)^2 - z_l(X)) */
        const fe *zl = z + (n_sets - 1) * z_stride;
        f_sqr(&FR, &t, &zl[idx]);
        f_sub(&FR, &t, &t, &zl[idx]);
        f_mul(&FR, &t, &t, &l_last[idx]);
        fold(&h[idx], y, &t);
        /* l_0(X) * (z_i(X) - z_{i-1}(omega^last X)) */
        for (size_t s = 1; s < n_sets; ++s) {
            f_sub(&FR, &t, &z[s * z_stride + idx], &z[(s - 1) * z_stride + r_last]);
            f_mul(&FR, &t, &t, &l0[idx]);
            fold(&h[idx], y, &t);
        }
        /* l_active(X) * (z_i(omega X) prod (p + beta s_j + gamma) - z_i(X) prod (p + delta^j beta X + gamma)) */
        fe current_delta;
        f_mul(&FR, &current_delta, &delta_start, &beta_term);
        for (size_t s = 0; s < n_sets; ++s) {
            const size_t c0 = s * chunk_len, c1 = c0 + chunk_len < n_cols ? c0 + chunk_len : n_cols;
            fe left = z[s * z_stride + r_next], right = z[s * z_stride + idx];
            for (size_t j = c0; j < c1; ++j) {
                f_mul(&FR, &t, beta, &sigma[j * sigma_stride + idx]);
                f_add(&FR, &t, &cols[j * cols_stride + idx], &t);
                f_add(&FR, &t, &t, gamma);
                f_mul(&FR, &left, &left, &t);
            }
            for (size_t j = c0; j < c1; ++j) {
                f_add(&FR, &u, &cols[j * cols_stride + idx], &current_delta);
                f_add(&FR, &u, &u, gamma);
                f_mul(&FR, &right, &right, &u);
                f_mul(&FR, &current_delta, &current_delta, &delta);
            }
            f_sub(&FR, &t, &left, &right);
            f_mul(&FR, &t, &t, &l_active[idx]);
            fold(&h[idx], y, &t);
        }
        f_mul(&FR, &beta_term, &beta_term, &d->ext_omega);
    }
}
void orc_quotient_lookup(const orc_domain *d, u64 *h_, const u64 y_[4], const u64 beta_[4], const u64 gamma_[4],
                         const u64 *input_, const u64 *table_, const u64 *perm_input_, const u64 *perm_table_, const u64 *z_,
                         const u64 *l0_, const u64 *l_last_, const u64 *l_active_) {
    const size_t n_ext = (size_t)1 << d->extended_k, rs = (size_t)1 << (d->extended_k - d->k);
    fe *h = (fe *)h_;
    const fe *y = (const fe *)y_, *beta = (const fe *)beta_, *gamma = (const fe *)gamma_;
    const fe *input = (const fe *)input_, *table = (const fe *)table_, *pa = (const fe *)perm_input_, *ps = (const fe *)perm_table_;
    const fe *z = (const fe *)z_, *l0 = (const fe *)l0_, *l_last = (const fe *)l_last_, *l_active = (const fe *)l_active_;
    fe one;
    memcpy(one.l, FR.r1, 32);
    for (size_t idx = 0; idx < n_ext; ++idx) {
        const size_t r_next = rot_idx(idx, 1, rs, n_ext), r_prev = rot_idx(idx, -1, rs, n_ext);
        fe t, u, w, a_minus_s;
        f_sub(&FR, &a_minus_s, &pa[idx], &ps[idx]);
        /* l_0(X) * (1 - z(X)) */
        f_sub(&FR, &t, &one, &z[idx]);
        f_mul(&FR, &t, &t, &l0[idx]);
        fold(&h[idx], y, &t);
        /* l_last(X) * (z(X)^2 - z(X)) */
        f_sqr(&FR, &t, &z[idx]);
        f_sub(&FR, &t, &t, &z[idx]);
        f_mul(&FR, &t, &t, &l_last[idx]);
        fold(&h[idx], y, &t);
        /* l_active(X) * (z(omega X) (a'(X) + beta) (s'(X) + gamma) - z(X) (a(X) + beta) (s(X) + gamma)) */
        f_add(&FR, &t, &pa[idx], beta);
        f_mul(&FR, &t, &z[r_next], &t);
        f_add(&FR, &u, &ps[idx], gamma);
        f_mul(&FR, &t, &t, &u);
        f_add(&FR, &u, &input[idx], beta);
        f_add(&FR, &w, &table[idx], gamma);
        f_mul(&FR, &u, &u, &w);
        f_mul(&FR, &u, &z[idx], &u);
        f_sub(&FR, &t, &t, &u);
        f_mul(&FR, &t, &t, &l_active[idx]);
        fold(&h[idx], y, &t);
        /* l_0(X) * (a'(X) - s'(X)) */
        f_mul(&FR, &t, &a_minus_s, &l0[idx]);
        fold(&h[idx], y, &t);
        /* l_active(X) * (a'(X) - s'(X)) (a'(X) - a'(omega^-1 X)) */
        f_sub(&FR, &t, &pa[idx], &pa[r_prev]);
        f_mul(&FR, &t, &a_minus_s, &t);
        f_mul(&FR, &t, &t, &l_active[idx]);
        fold(&h[idx], y, &t);
    }
}

/* ---- lookup argument: permuted input / table columns ----------------------------------------------------------------
 * Restates halo2-axiom plonk/lookup/prover.rs permute_expression_pair [UPSTREAM, absent from /root/reference; create_proof
 * step 5, reached from src/scaffold/mod.rs:296], as recalled from the zcash/PSE lineage: sort the usable input rows (Ord on
 * Fr = canonical value); count the table values in an ordered map; walking the sorted input, a row that starts a new value
 * takes that value and removes one copy from the map (missing -> ConstraintSystemFailure), every other row is remembered as
 * "repeated"; finally the map's leftovers, in ascending key order, are written to the repeated rows popped from the END of
 * the list.  Returns 0, or -1 when an input value is not in the table.  PARITY UNPINNED (the pop order is the recalled one). */
static int canon_cmp(const void *a, const void *b) {
    const u64 *x = a, *y = b;
    for (int i = 3; i >= 0; --i)
        if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}
int orc_permute_expression_pair(const u64 *input, const u64 *table, size_t usable_rows, u64 *perm_input, u64 *perm_table) {
    const size_t u = usable_rows;
    if (!u) return 0;
    u64 *a = malloc(u * 32), *t = malloc(u * 32);
    size_t *repeated = malloc(u * sizeof(size_t));
    unsigned char *taken = calloc(u, 1);
    for (size_t i = 0; i < u; ++i) {
        f_from_mont(&FR, a + 4 * i, (const fe *)(input + 4 * i));
        f_from_mont(&FR, t + 4 * i, (const fe *)(table + 4 * i));
    }
    qsort(a, u, 32, canon_cmp);
    qsort(t, u, 32, canon_cmp);      /* the ordered map: sorted table values, `taken` marks removed copies */
    size_t n_rep = 0, tp = 0;
    int rc = 0;
    u64 *pa = malloc(u * 32), *ps = malloc(u * 32);
    memcpy(pa, a, u * 32);
    for (size_t row = 0; row < u && rc == 0; ++row) {
        if (row == 0 || canon_cmp(a + 4 * row, a + 4 * (row - 1)) != 0) {
            memcpy(ps + 4 * row, a + 4 * row, 32);
            while (tp < u && canon_cmp(t + 4 * tp, a + 4 * row) < 0) ++tp;      /* both sequences ascend */
            if (tp < u && canon_cmp(t + 4 * tp, a + 4 * row) == 0) taken[tp++] = 1;
            else rc = -1;
        } else {
            repeated[n_rep++] = row;
        }
    }
    if (rc == 0) {
        for (size_t p = 0; p < u; ++p) {
            if (taken[p]) continue;
            memcpy(ps + 4 * repeated[--n_rep], t + 4 * p, 32);      /* repeated_input_rows.pop() */
        }
        for (size_t i = 0; i < u; ++i) {
            f_to_mont(&FR, (fe *)(perm_input + 4 * i), pa + 4 * i);
            f_to_mont(&FR, (fe *)(perm_table + 4 * i), ps + 4 * i);
        }
    }
    free(a); free(t); free(repeated); free(taken); free(pa); free(ps);
    return rc;
}
