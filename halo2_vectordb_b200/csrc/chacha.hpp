// chacha.hpp -- the prover's seeded random number generator (host side).
//
// Restates rand_chacha `ChaCha20Rng::from_seed(seed)` [UPSTREAM; halo2-base `gen_srs` seeds it with [0u8; 32],
// reached from /root/reference/src/scaffold/mod.rs:260; the "seeded RNG" of BASELINE.json's configs is the same
// generator handed to create_proof]: the ChaCha20 key stream (RFC 7539 block function, 64-bit block counter in state
// words 12-13, stream id 0) consumed as consecutive little-endian words; `next_u64` = two consecutive words, low
// word first.  `Fr::random(rng)` (halo2curves derive/field.rs) reads eight `next_u64` values as a 512-bit
// little-endian integer and reduces it mod r.
#pragma once
#include <stdint.h>
#include <string.h>

#include "fr_host.hpp"

namespace h2v {

struct ChaCha20Rng {
    uint32_t key[8];
    uint64_t counter = 0;
    uint32_t buf[16];
    int pos = 16;

    explicit ChaCha20Rng(const uint8_t seed[32]) { memcpy(key, seed, 32); }   // little-endian host

    static inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
    static inline void qr(uint32_t *s, int a, int b, int c, int d) {
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 16);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 12);
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 8);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 7);
    }
    static void block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
        uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                             key[4], key[5], key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
        uint32_t s[16];
        memcpy(s, init, sizeof s);
        for (int i = 0; i < 10; ++i) {
            qr(s, 0, 4, 8, 12); qr(s, 1, 5, 9, 13); qr(s, 2, 6, 10, 14); qr(s, 3, 7, 11, 15);
            qr(s, 0, 5, 10, 15); qr(s, 1, 6, 11, 12); qr(s, 2, 7, 8, 13); qr(s, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; ++i) out[i] = s[i] + init[i];
    }
    uint32_t next_u32() {
        if (pos == 16) {
            block(key, counter++, buf);
            pos = 0;
        }
        return buf[pos++];
    }
    uint64_t next_u64() {
        uint64_t lo = next_u32();
        return lo | ((uint64_t)next_u32() << 32);
    }
    Fr64 fr_random() {
        uint64_t w[8];
        for (int i = 0; i < 8; ++i) w[i] = next_u64();
        return frh::from_u512(w);
    }
    // `Fr::random` consumes exactly one 64-byte block, so as long as nothing else is drawn the i-th draw is block i of the
    // key stream: a run of draws can be produced out of line (chacha_fr_kernel) and skipped here
    bool block_aligned() const { return pos == 16; }
    uint64_t next_block() const { return counter; }
    void skip_fr(uint64_t n) { counter += n; }
};

}  // namespace h2v
