"""Fiat-Shamir transcript, proof wire format and seeded RNG of the reference's prove / verify flow, restated in
plain Python integers (TEST INFRASTRUCTURE ONLY -- see bn254_oracle.c header; nothing on the product path imports it).

What is restated, and from where (all [UPSTREAM]: un-vendored git dependencies of /root/reference/Cargo.toml:19-28):

* `ChaCha20Rng` (rand_chacha; `gen_srs` seeds it with zeros, halo2-base utils/fs.rs, reached from
  /root/reference/src/scaffold/mod.rs:260) and `Fr::random(rng)` = `from_u512` of eight `next_u64` words
  (halo2curves derive/field.rs).  Pinned by the RFC 7539 A.1 key-stream vectors (tests/golden/external_vectors.json).
* Poseidon over BN254 Fr with the Grain-LFSR round constants and Cauchy MDS matrix of the Poseidon paper
  (PSE `poseidon` crate `Spec::new(R_F, R_P)`), pinned by the reference implementation's published test vectors
  `poseidonperm_x5_254_3` / `poseidonperm_x5_254_5` (hadeshash, tests/golden/external_vectors.json).
* snark-verifier `util/hash/poseidon.rs` sponge (state[0] = 2^64, RATE-sized chunks, padding by adding 1 to the
  first unused rate word, one extra permutation when the buffer is a multiple of RATE) and
  `system/halo2/transcript/halo2.rs` `PoseidonTranscript<G1Affine, NativeLoader, _, T = 5, RATE = 4, R_F = 8, R_P = 60>`
  created with `new::<0>` at /root/reference/src/scaffold/mod.rs:309-310: points are absorbed as (x mod r, y mod r),
  scalars as themselves; `write_point` appends the 32-byte compressed point, `write_scalar` the 32-byte little-endian
  canonical scalar.  RECALLED, not verifiable here (DESIGN.md "recalled conventions").
"""
import struct

from . import pyref as P

R = P.R
FQ = P.P

# ------------------------------------------------------------------------------------------- ChaCha20Rng


def _rotl(v, n):
    return ((v << n) & 0xFFFFFFFF) | (v >> (32 - n))


def _qr(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & 0xFFFFFFFF
    s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & 0xFFFFFFFF
    s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & 0xFFFFFFFF
    s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & 0xFFFFFFFF
    s[b] = _rotl(s[b] ^ s[c], 7)


def chacha20_block(key_words, counter, stream=0):
    """One 64-byte block as 16 little-endian u32 words; 64-bit block counter in words 12-13, 64-bit stream id in 14-15
    (rand_chacha's layout; identical to RFC 7539 for counter < 2^32 and a zero nonce)."""
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, stream & 0xFFFFFFFF, (stream >> 32) & 0xFFFFFFFF]
    s = list(init)
    for _ in range(10):
        _qr(s, 0, 4, 8, 12); _qr(s, 1, 5, 9, 13); _qr(s, 2, 6, 10, 14); _qr(s, 3, 7, 11, 15)
        _qr(s, 0, 5, 10, 15); _qr(s, 1, 6, 11, 12); _qr(s, 2, 7, 8, 13); _qr(s, 3, 4, 9, 14)
    return [(a + b) & 0xFFFFFFFF for a, b in zip(s, init)]


class ChaCha20Rng:
    """rand_chacha `ChaCha20Rng::from_seed(seed)`: the key stream of ChaCha20 (key = seed, counter 0, stream 0) read
    as consecutive little-endian words.  Only `next_u64` is used on this path (`Fr::random`)."""

    def __init__(self, seed=bytes(32)):
        assert len(seed) == 32
        self.key = struct.unpack("<8I", seed)
        self.counter = 0
        self.buf = []

    def next_u32(self):
        if not self.buf:
            self.buf = chacha20_block(self.key, self.counter)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self):
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)

    def fr_random(self):
        """halo2curves `Fr::random(rng)` = `from_u512([rng.next_u64(); 8])`: the 512-bit little-endian integer mod r"""
        v = 0
        for i in range(8):
            v |= self.next_u64() << (64 * i)
        return v % R


# ------------------------------------------------------------------------------------------- Poseidon
class _Grain:
    """Grain LFSR of the Poseidon paper (supplementary material F), as in `generate_parameters_grain.sage`"""

    def __init__(self, n_bits, t, r_f, r_p):
        bits = []

        def app(nb, v):
            for i in range(nb - 1, -1, -1):
                bits.append((v >> i) & 1)

        app(2, 1)          # prime field
        app(4, 0)          # x^alpha S-box
        app(12, n_bits)
        app(12, t)
        app(10, r_f)
        app(10, r_p)
        app(30, (1 << 30) - 1)
        self.b = bits
        for _ in range(160):
            self._new_bit()

    def _new_bit(self):
        b = self.b
        nb = b[62] ^ b[51] ^ b[38] ^ b[23] ^ b[13] ^ b[0]
        b.pop(0)
        b.append(nb)
        return nb

    def _next(self):
        nb = self._new_bit()
        while not nb:
            self._new_bit()
            nb = self._new_bit()
        return self._new_bit()

    def bits(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | self._next()
        return v


_SPECS = {}


def poseidon_spec(t, r_f, r_p):
    """(round constants [(r_f + r_p) * t], MDS matrix t x t) for BN254 Fr, alpha = 5"""
    key = (t, r_f, r_p)
    if key not in _SPECS:
        g = _Grain(254, t, r_f, r_p)
        rc = []
        for _ in range((r_f + r_p) * t):
            v = g.bits(254)
            while v >= R:
                v = g.bits(254)
            rc.append(v)
        xs = [g.bits(254) % R for _ in range(t)]
        ys = [g.bits(254) % R for _ in range(t)]
        mds = [[pow(xs[i] + ys[j], -1, R) for j in range(t)] for i in range(t)]
        _SPECS[key] = (rc, mds)
    return _SPECS[key]


def poseidon_permutation(state, r_f, r_p):
    t = len(state)
    rc, mds = poseidon_spec(t, r_f, r_p)
    s = list(state)
    c = 0
    for rnd in range(r_f + r_p):
        for i in range(t):
            s[i] = (s[i] + rc[c]) % R
            c += 1
        if rnd < r_f // 2 or rnd >= r_f // 2 + r_p:
            s = [pow(x, 5, R) for x in s]
        else:
            s[0] = pow(s[0], 5, R)
        s = [sum(mds[i][j] * s[j] for j in range(t)) % R for i in range(t)]
    return s


class PoseidonSponge:
    """snark-verifier util/hash/poseidon.rs `Poseidon<F, L, T, RATE>`"""

    def __init__(self, t=5, rate=4, r_f=8, r_p=60):
        self.t, self.rate, self.r_f, self.r_p = t, rate, r_f, r_p
        self.state = [1 << 64] + [0] * (t - 1)
        self.buf = []

    def update(self, elems):
        self.buf.extend(int(e) % R for e in elems)

    def _permute(self, chunk):
        s = self.state
        for i, v in enumerate(chunk):
            s[1 + i] = (s[1 + i] + v) % R
        if len(chunk) + 1 < self.t:
            s[len(chunk) + 1] = (s[len(chunk) + 1] + 1) % R
        self.state = poseidon_permutation(s, self.r_f, self.r_p)

    def squeeze(self):
        buf, self.buf = self.buf, []
        exact = len(buf) % self.rate == 0
        for i in range(0, len(buf), self.rate):
            self._permute(buf[i:i + self.rate])
        if exact:
            self._permute([])
        return self.state[1]


# ------------------------------------------------------------------------------------------- wire format
# halo2curves 0.3.x `G1Affine::to_bytes()` (derive/curve.rs `new_curve_impl!`): canonical x little-endian, the parity of
# canonical y in the top bit of the last byte (`(y.to_bytes()[0] & 1) << 7`), identity = 32 zero bytes.  RECALLED.
SIGN_BIT = 7


def g1_to_bytes(pt):
    if pt is None:
        return bytes(32)
    b = bytearray(pt[0].to_bytes(32, "little"))
    b[31] |= (pt[1] & 1) << SIGN_BIT
    return bytes(b)


def g1_from_bytes(b):
    b = bytearray(b)
    sign = (b[31] >> SIGN_BIT) & 1
    b[31] &= ~(1 << SIGN_BIT) & 0xFF
    x = int.from_bytes(b, "little")
    if x == 0 and sign == 0:
        return None
    if x >= FQ:
        raise ValueError("point encoding: x not reduced")
    y2 = (x * x * x + 3) % FQ
    y = pow(y2, (FQ + 1) // 4, FQ)
    if y * y % FQ != y2:
        raise ValueError("point encoding: x is not on the curve")
    if (y & 1) != sign:
        y = FQ - y
    return (x, y)


def fr_to_repr(v):
    return (int(v) % R).to_bytes(32, "little")


def fr_from_repr(b):
    v = int.from_bytes(b, "little")
    if v >= R:
        raise ValueError("scalar encoding: not reduced")
    return v


class PoseidonTranscript:
    """`PoseidonTranscript<G1Affine, NativeLoader, W/R, 5, 4, 8, 60>::new::<0>`; writer when `proof` is None, else reader"""

    def __init__(self, proof=None):
        self.sponge = PoseidonSponge()
        self.out = bytearray()
        self.inp = None if proof is None else memoryview(bytes(proof))
        self.pos = 0

    # Transcript
    def squeeze_challenge(self):
        return self.sponge.squeeze()

    def common_point(self, pt):
        if pt is None:
            raise ValueError("Cannot write points at infinity to the transcript")
        self.sponge.update([pt[0] % R, pt[1] % R])

    def common_scalar(self, s):
        self.sponge.update([s])

    # TranscriptWrite
    def write_point(self, pt):
        self.common_point(pt)
        self.out += g1_to_bytes(pt)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.out += fr_to_repr(s)

    # TranscriptRead
    def _take(self):
        if self.pos + 32 > len(self.inp):
            raise ValueError("proof too short")
        b = bytes(self.inp[self.pos:self.pos + 32])
        self.pos += 32
        return b

    def read_point(self):
        pt = g1_from_bytes(self._take())
        self.common_point(pt)
        return pt

    def read_scalar(self):
        s = fr_from_repr(self._take())
        self.common_scalar(s)
        return s

    def finalize(self):
        return bytes(self.out)
