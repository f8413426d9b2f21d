/* Plain-C use of the ABI (include/h2v.h): build a small SRS on the device, commit one column in both bases,
 * transform it, and print the commitment in the proof wire format.  Shows that the header is C99-clean and
 * what a cgo / JNI / Rust-FFI binding has to call.
 *   gcc -std=c99 -Iinclude examples/commit_example.c -Lhalo2_vectordb_b200 -lh2v -Wl,-rpath,$PWD/halo2_vectordb_b200 -o /tmp/commit_example
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "h2v.h"

#define CHECK(x) do { if ((x) != H2V_OK) { fprintf(stderr, "%s: %s\n", #x, h2v_last_error()); return 1; } } while (0)

int main(void) {
    const uint32_t k = 10;
    const size_t n = (size_t)1 << k;
    if (h2v_device_count() <= 0) { printf("no CUDA device: %s\n", h2v_init(NULL, 0) ? h2v_last_error() : "?"); return 0; }
    CHECK(h2v_init(NULL, 0));      /* device 0; a list of devices makes the batch calls use all of them */
    /* any non-zero value below r is a valid Montgomery-form scalar */
    uint64_t s[4] = {0x0123456789abcdefULL, 0xfedcba9876543210ULL, 0x1111111111111111ULL, 0x0222222222222222ULL};
    uint64_t *g = malloc(n * 64), *gl = malloc(n * 64), *col = malloc(n * 32), *coef = malloc(n * 32);
    CHECK(h2v_srs_setup(k, s, g, gl));
    h2v_srs_t srs;
    h2v_domain_t dom;
    CHECK(h2v_srs_load(k, g, gl, &srs));
    CHECK(h2v_domain_new(4, k, &dom));
    for (size_t i = 0; i < n; ++i) { col[4 * i] = i * 0x9e3779b97f4a7c15ULL + 1; col[4 * i + 1] = i; col[4 * i + 2] = 7; col[4 * i + 3] = i & 0xff; }
    uint64_t c_lagrange[8], c_monomial[8];
    CHECK(h2v_commit(srs, H2V_BASIS_LAGRANGE, col, n, c_lagrange));
    memcpy(coef, col, n * 32);
    CHECK(h2v_lagrange_to_coeff(dom, coef));
    CHECK(h2v_commit(srs, H2V_BASIS_MONOMIAL, coef, n, c_monomial));
    uint8_t wire[32];
    CHECK(h2v_g1_to_bytes(c_lagrange, 1, wire));
    printf("commit_lagrange(evals) %s commit(coeffs); wire bytes: ", memcmp(c_lagrange, c_monomial, 64) ? "!=" : "==");
    for (int i = 0; i < 32; ++i) printf("%02x", wire[i]);
    printf("\n");
    h2v_domain_free(dom);
    h2v_srs_free(srs);
    free(g); free(gl); free(col); free(coef);
    return memcmp(c_lagrange, c_monomial, 64) != 0;
}
