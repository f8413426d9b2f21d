// poseidon.hpp -- the Fiat-Shamir transcript of the reference's prove / verify flow (host side).
//
// Restates [UPSTREAM; un-vendored dependencies of /root/reference/Cargo.toml:24,28]:
//   * Poseidon over BN254 Fr, x^5 S-box, round constants and Cauchy MDS matrix from the Grain LFSR of the Poseidon
//     paper (PSE `poseidon` crate `Spec::new`; the first generated matrix, `SECURE_MDS = 0`);
//   * snark-verifier util/hash/poseidon.rs `Poseidon<F, L, T, RATE>`: state[0] = 2^64, input absorbed RATE words at a
//     time into state[1..], 1 added to the first unused rate word, an extra permutation of the empty chunk when the
//     buffered input is a multiple of RATE, the challenge is state[1];
//   * snark-verifier system/halo2/transcript/halo2.rs `PoseidonTranscript<G1Affine, NativeLoader, W, 5, 4, 8, 60>`
//     (`new::<0>` at /root/reference/src/scaffold/mod.rs:309-310): a point is absorbed as (x mod r, y mod r) and
//     written as the 32-byte compressed `G1Affine::to_bytes()`, a scalar is absorbed as itself and written as the
//     32-byte little-endian `to_repr()`; the identity cannot be written.
// The permutation is checked against the Poseidon reference implementation's published vectors
// (tests/golden/external_vectors.json); the sponge and transcript framing are recalled.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <future>
#include <vector>

#include "ec.cuh"
#include "fr_host.hpp"

namespace h2v {

template <int T> struct PoseidonSpec {
    int r_f, r_p;
    std::vector<Fr64> rc;     // (r_f + r_p) * T round constants
    Fr64 mds[T][T];
    // Partial rounds in their sparse form (an exact rewriting of the same permutation, checked against the plain
    // rounds by tests/): with M = [[m00, v^T], [w, M^]] = diag(1, M^) * [[m00, v^T], [M^^-1 w, I]] the dense factor
    // commutes with the one-word S-box and is applied once after the last partial round (as diag(1, M^^r_p)), and only
    // the first word needs a round constant (the rest is pushed forward into the next full round's constants):
    //   round j: s0 <- (s0 + pc[j])^5;  (s0, s_rest) <- (m00 s0 + pv[j] . s_rest,  s_rest + pw[j] s0)
    std::vector<Fr64> pc;             // r_p scalars
    std::vector<Fr64> pv, pw;         // r_p x (T - 1) each
    Fr64 m00;
    Fr64 dense[T - 1][T - 1];         // M^^r_p
    Fr64 rc_after[T];                 // the constants of the first full round after the partial rounds, adjusted
};

namespace poseidon_detail {
struct Grain {
    uint8_t b[80];
    int head = 0;     // b[(head + i) % 80] is bit i
    Grain(uint32_t n_bits, uint32_t t, uint32_t r_f, uint32_t r_p) {
        int p = 0;
        auto app = [&](int nb, uint32_t v) {
            for (int i = nb - 1; i >= 0; --i) b[p++] = (v >> i) & 1u;
        };
        app(2, 1);      // prime field
        app(4, 0);      // x^alpha S-box
        app(12, n_bits);
        app(12, t);
        app(10, r_f);
        app(10, r_p);
        app(30, 0x3fffffffu);
        for (int i = 0; i < 160; ++i) new_bit();
    }
    uint8_t at(int i) const { return b[(head + i) % 80]; }
    uint8_t new_bit() {
        uint8_t nb = at(62) ^ at(51) ^ at(38) ^ at(23) ^ at(13) ^ at(0);
        b[head] = nb;               // drop bit 0, append at the end
        head = (head + 1) % 80;
        return nb;
    }
    uint8_t next() {
        uint8_t nb = new_bit();
        while (!nb) {
            new_bit();
            nb = new_bit();
        }
        return new_bit();
    }
    // 254 bits, most significant first, as a canonical 4 x 64 little-endian integer
    void bits254(uint64_t out[4]) {
        out[0] = out[1] = out[2] = out[3] = 0;
        for (int i = 253; i >= 0; --i)
            if (next()) out[i >> 6] |= (uint64_t)1 << (i & 63);
    }
};
}  // namespace poseidon_detail

namespace poseidon_detail {
// (T-1) x (T-1) inverse by Gauss-Jordan over Fr (the MDS submatrix is invertible)
template <int N> inline void mat_inverse(const Fr64 (&a)[N][N], Fr64 (&out)[N][N]) {
    Fr64 m[N][2 * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            m[i][j] = a[i][j];
            m[i][N + j] = i == j ? frh::ONE : frh::zero();
        }
    for (int c = 0; c < N; ++c) {
        int p = c;
        while (p < N && frh::is_zero(m[p][c])) ++p;
        if (p != c)
            for (int j = 0; j < 2 * N; ++j) std::swap(m[p][j], m[c][j]);
        const Fr64 iv = frh::inv(m[c][c]);
        for (int j = 0; j < 2 * N; ++j) m[c][j] = frh::mul(m[c][j], iv);
        for (int r = 0; r < N; ++r) {
            if (r == c || frh::is_zero(m[r][c])) continue;
            const Fr64 f = m[r][c];
            for (int j = 0; j < 2 * N; ++j) m[r][j] = frh::sub(m[r][j], frh::mul(f, m[c][j]));
        }
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) out[i][j] = m[i][N + j];
}
template <int T> inline void build_sparse(PoseidonSpec<T> &s) {
    constexpr int N = T - 1;
    const int half = s.r_f / 2;
    Fr64 mh[N][N], mhi[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) mh[i][j] = s.mds[i + 1][j + 1];
    mat_inverse<N>(mh, mhi);
    s.m00 = s.mds[0][0];
    Fr64 v[N], w[N];
    for (int j = 0; j < N; ++j) v[j] = s.mds[0][j + 1];
    for (int i = 0; i < N; ++i) {       // w^ = M^^-1 w
        Fr64 acc = frh::zero();
        for (int j = 0; j < N; ++j) acc = frh::add(acc, frh::mul(mhi[i][j], s.mds[j + 1][0]));
        w[i] = acc;
    }
    // constants: a_0 = c[half]; a_{j+1} = c[half + j + 1] + M (0, a_j[1..])
    Fr64 a[T];
    for (int i = 0; i < T; ++i) a[i] = s.rc[(size_t)half * T + i];
    Fr64 dense[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) dense[i][j] = i == j ? frh::ONE : frh::zero();
    s.pc.resize(s.r_p);
    s.pv.resize((size_t)s.r_p * N);
    s.pw.resize((size_t)s.r_p * N);
    for (int r = 0; r < s.r_p; ++r) {
        s.pc[r] = a[0];
        for (int j = 0; j < N; ++j) {
            s.pv[(size_t)r * N + j] = v[j];
            s.pw[(size_t)r * N + j] = w[j];
        }
        // next round's sparse factors: v^T <- v^T M^,  w^ <- M^^-1 w^
        Fr64 nv[N], nw[N];
        for (int j = 0; j < N; ++j) {
            Fr64 acc = frh::zero(), acc2 = frh::zero();
            for (int i = 0; i < N; ++i) {
                acc = frh::add(acc, frh::mul(v[i], mh[i][j]));
                acc2 = frh::add(acc2, frh::mul(mhi[j][i], w[i]));
            }
            nv[j] = acc;
            nw[j] = acc2;
        }
        for (int j = 0; j < N; ++j) { v[j] = nv[j]; w[j] = nw[j]; }
        // dense <- dense * M^
        Fr64 nd[N][N];
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                Fr64 acc = frh::zero();
                for (int k = 0; k < N; ++k) acc = frh::add(acc, frh::mul(dense[i][k], mh[k][j]));
                nd[i][j] = acc;
            }
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) dense[i][j] = nd[i][j];
        // pending constants pushed through M: a <- c[next] + M (0, a[1..])
        Fr64 na[T];
        for (int i = 0; i < T; ++i) {
            Fr64 acc = s.rc[(size_t)(half + r + 1) * T + i];
            for (int j = 1; j < T; ++j) acc = frh::add(acc, frh::mul(s.mds[i][j], a[j]));
            na[i] = acc;
        }
        for (int i = 0; i < T; ++i) a[i] = na[i];
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) s.dense[i][j] = dense[i][j];
    for (int i = 0; i < T; ++i) s.rc_after[i] = a[i];
}
}  // namespace poseidon_detail

template <int T> inline PoseidonSpec<T> poseidon_make_spec(int r_f, int r_p) {
    PoseidonSpec<T> s;
    s.r_f = r_f;
    s.r_p = r_p;
    poseidon_detail::Grain g(254, T, (uint32_t)r_f, (uint32_t)r_p);
    s.rc.reserve((size_t)(r_f + r_p) * T);
    for (int i = 0; i < (r_f + r_p) * T; ++i) {
        uint64_t v[4];
        do g.bits254(v);
        while (frh::geq_mod(v));                          // rejection sampling
        s.rc.push_back(frh::to_mont(Fr64{{v[0], v[1], v[2], v[3]}}));
    }
    Fr64 xs[T], ys[T];
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < T; ++i) {
            uint64_t v[4];
            g.bits254(v);
            if (frh::geq_mod(v)) frh::sub_mod(v);          // reduced, not rejected (2^254 < 2r)
            (pass ? ys : xs)[i] = frh::to_mont(Fr64{{v[0], v[1], v[2], v[3]}});
        }
    for (int i = 0; i < T; ++i)
        for (int j = 0; j < T; ++j) s.mds[i][j] = frh::inv(frh::add(xs[i], ys[j]));
    poseidon_detail::build_sparse<T>(s);
    return s;
}

// the permutation as the Poseidon paper writes it (every round: constants, S-box layer, dense MDS)
template <int T> inline void poseidon_permute_plain(const PoseidonSpec<T> &sp, Fr64 (&st)[T]) {
    const int half = sp.r_f / 2;
    const Fr64 *rc = sp.rc.data();
    for (int rnd = 0; rnd < sp.r_f + sp.r_p; ++rnd) {
        for (int i = 0; i < T; ++i) st[i] = frh::add(st[i], *rc++);
        if (rnd < half || rnd >= half + sp.r_p) {
            for (int i = 0; i < T; ++i) st[i] = frh::pow5(st[i]);
        } else {
            st[0] = frh::pow5(st[0]);
        }
        Fr64 nx[T];
        for (int i = 0; i < T; ++i) {
            Fr64 acc = frh::mul(sp.mds[i][0], st[0]);
            for (int j = 1; j < T; ++j) acc = frh::add(acc, frh::mul(sp.mds[i][j], st[j]));
            nx[i] = acc;
        }
        for (int i = 0; i < T; ++i) st[i] = nx[i];
    }
}
// the same permutation with the partial rounds in sparse form and one Montgomery reduction per matrix row
// (the transcript is a single dependent chain of ~10^3 permutations per proof: host latency that nothing can hide)
template <int T> inline void poseidon_permute(const PoseidonSpec<T> &sp, Fr64 (&st)[T]) {
    constexpr int N = T - 1;
    const int half = sp.r_f / 2;
    auto full_round = [&](const Fr64 *rc) {
        for (int i = 0; i < T; ++i) st[i] = frh::pow5(frh::add(st[i], rc[i]));
        Fr64 nx[T];
        for (int i = 0; i < T; ++i) nx[i] = frh::dot(sp.mds[i], st, T);
        for (int i = 0; i < T; ++i) st[i] = nx[i];
    };
    for (int r = 0; r < half; ++r) full_round(sp.rc.data() + (size_t)r * T);
    for (int r = 0; r < sp.r_p; ++r) {
        const Fr64 s0 = frh::pow5(frh::add(st[0], sp.pc[r]));
        const Fr64 *v = sp.pv.data() + (size_t)r * N, *w = sp.pw.data() + (size_t)r * N;
        uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        frh::mul_acc_wide(t, sp.m00, s0);
        for (int j = 0; j < N; ++j) frh::mul_acc_wide(t, v[j], st[j + 1]);
        for (int j = 0; j < N; ++j) st[j + 1] = frh::add(st[j + 1], frh::mul(w[j], s0));
        st[0] = frh::redc_wide(t);
    }
    {   // the dense factor collected over the partial rounds: diag(1, M^^r_p)
        Fr64 nx[N];
        for (int i = 0; i < N; ++i) nx[i] = frh::dot(sp.dense[i], st + 1, N);
        for (int i = 0; i < N; ++i) st[i + 1] = nx[i];
    }
    full_round(sp.rc_after);
    for (int r = half + sp.r_p + 1; r < sp.r_f + sp.r_p; ++r) full_round(sp.rc.data() + (size_t)r * T);
}

template <int T, int RATE> struct PoseidonSponge {
    const PoseidonSpec<T> *spec;
    Fr64 state[T];
    std::vector<Fr64> buf;
    // absorb_async(): the full RATE-word chunks buffered so far are absorbed on a worker thread (a full chunk is absorbed
    // the same way whatever follows it), so that the prover can queue kernels meanwhile; squeeze() joins it first
    std::vector<Fr64> work;
    std::future<void> job;
    explicit PoseidonSponge(const PoseidonSpec<T> *s) : spec(s) {
        for (int i = 0; i < T; ++i) state[i] = frh::zero();
        state[0] = frh::to_mont(Fr64{{0, 1, 0, 0}});      // 2^64
    }
    PoseidonSponge(const PoseidonSponge &) = delete;
    PoseidonSponge &operator=(const PoseidonSponge &) = delete;
    ~PoseidonSponge() { join(); }
    void update(const Fr64 &x) { buf.push_back(x); }
    void permute_chunk(const Fr64 *chunk, size_t len) {
        for (size_t i = 0; i < len; ++i) state[1 + i] = frh::add(state[1 + i], chunk[i]);
        if (len + 1 < (size_t)T) state[len + 1] = frh::add(state[len + 1], frh::ONE);
        poseidon_permute<T>(*spec, state);
    }
    void join() {
        if (job.valid()) job.get();
    }
    void absorb_async() {
        join();
        const size_t full = buf.size() / RATE * RATE;
        if (!full) return;
        work.assign(buf.begin(), buf.begin() + full);
        buf.erase(buf.begin(), buf.begin() + full);
        job = std::async(std::launch::async, [this] {
            for (size_t i = 0; i < work.size(); i += RATE) permute_chunk(work.data() + i, RATE);
        });
    }
    Fr64 squeeze() {
        join();
        std::vector<Fr64> b;
        b.swap(buf);
        const bool exact = b.size() % RATE == 0;
        for (size_t i = 0; i < b.size(); i += RATE) permute_chunk(b.data() + i, std::min<size_t>(RATE, b.size() - i));
        if (exact) permute_chunk(nullptr, 0);
        return state[1];
    }
};

// Proof wire format.  halo2curves 0.3.x `G1Affine::to_bytes()`: canonical x little-endian with the parity of the
// canonical y in the top bit of the last byte; `Fr::to_repr()`: canonical little-endian.  (RECALLED: halo2curves >= 0.4
// moved the sign to bit 6; H2V_G1_SIGN_BIT is the one place to change.)
#define H2V_G1_SIGN_BIT 7
inline void g1_affine_to_bytes(const affine &p, uint8_t out[32]) {
    if (affine_is_identity(p)) {
        memset(out, 0, 32);
        return;
    }
    fe x = fe_from_mont<FqP>(p.x), y = fe_from_mont<FqP>(p.y);
    memcpy(out, x.v, 32);
    out[31] |= (uint8_t)((y.v[0] & 1u) << H2V_G1_SIGN_BIT);
}
// canonical Fq value mod r as a Montgomery Fr (snark-verifier `fe_to_fe`): p < 2r, one conditional subtraction
inline Fr64 fq_canonical_to_fr(const fe &canon) {
    uint64_t v[4];
    memcpy(v, canon.v, 32);
    if (frh::geq_mod(v)) frh::sub_mod(v);
    return frh::to_mont(Fr64{{v[0], v[1], v[2], v[3]}});
}

const PoseidonSpec<5> &poseidon_transcript_spec();      // T = 5, R_F = 8, R_P = 60 (built once; prover.cu)

struct PoseidonTranscript {
    PoseidonSponge<5, 4> sponge;
    std::vector<uint8_t> out;
    PoseidonTranscript() : sponge(&poseidon_transcript_spec()) {}
    Fr64 squeeze_challenge() { return sponge.squeeze(); }
    void absorb_async() { sponge.absorb_async(); }      // hash what has been written so far while the caller goes on
    // false: the identity cannot be absorbed (upstream: Error::Transcript "Cannot write points at infinity to the transcript")
    bool common_point(const affine &p) {
        if (affine_is_identity(p)) return false;
        sponge.update(fq_canonical_to_fr(fe_from_mont<FqP>(p.x)));
        sponge.update(fq_canonical_to_fr(fe_from_mont<FqP>(p.y)));
        return true;
    }
    void common_scalar(const Fr64 &s) { sponge.update(s); }
    bool write_point(const affine &p) {
        if (!common_point(p)) return false;
        uint8_t b[32];
        g1_affine_to_bytes(p, b);
        out.insert(out.end(), b, b + 32);
        return true;
    }
    void write_scalar(const Fr64 &s) {
        common_scalar(s);
        Fr64 c = frh::from_mont(s);
        const uint8_t *b = reinterpret_cast<const uint8_t *>(c.l);
        out.insert(out.end(), b, b + 32);
    }
};

}  // namespace h2v
