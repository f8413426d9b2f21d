"""create_proof: the product's C++/CUDA prover against the CPU restatement, byte for byte,
and against the verifier.

CPU: the oracle prover / verifier pair on satisfied and unsatisfied toy circuits (they only share the transcript and
point-set helpers), and the product's host-side transcript against the oracle's.
GPU: `h2v_create_proof` through the C ABI -- proof bytes identical to the oracle prover's for the same circuit, witness
and RNG seed at k = 6 ... 16; the verifier accepts; a flipped witness cell / public input / proof byte is rejected."""
import numpy as np
import pytest

from common import fr_arr
from oracle import oracle as O
from oracle import plonk as PL
from oracle import pyref as P
from oracle import transcript as T
from toy_circuit import Toy

SEED = bytes(range(32))
SECRET = 0x1CE1CEBABE5EED0123456789ABCDEF0FEDCBA9876543210


def _oracle_proof(params, t, seed=SEED):
    return PL.create_proof(params, t.cs, t.fixed, t.sigma, t.vk_repr, t.advice, t.instances, seed)


@pytest.fixture(scope="module")
def params6():
    return PL.Params.setup(6, SECRET)


@pytest.mark.parametrize("gate_cols,lookup_cols,degree", [(2, 1, 4), (3, 2, 4), (1, 1, 5), (4, 0, 4), (2, 1, 6)])
def test_oracle_prover_verifier_accepts(params6, gate_cols, lookup_cols, degree):
    t = Toy(6, seed=gate_cols * 10 + lookup_cols, n_gate_cols=gate_cols, n_lookup_cols=lookup_cols, degree=degree)
    proof = _oracle_proof(params6, t)
    vk = PL.keygen_vk(params6, t.cs, t.fixed, t.sigma)
    assert PL.verify_proof(params6, t.cs, vk, t.vk_repr, t.instances, proof)
    # deterministic in the seed, and the seed matters
    assert proof == _oracle_proof(params6, t)
    assert proof != _oracle_proof(params6, t, bytes(32))
    # any flipped byte is rejected (a commitment, an evaluation, the last opening proof)
    for pos in (5, len(proof) // 2, len(proof) - 7):
        bad = bytearray(proof)
        bad[pos] ^= 4
        assert not PL.verify_proof(params6, t.cs, vk, t.vk_repr, t.instances, bytes(bad))
    assert not PL.verify_proof(params6, t.cs, vk, (t.vk_repr + 1) % P.R, t.instances, proof)
    assert not PL.verify_proof(params6, t.cs, vk, t.vk_repr, [[(t.instances[0][0] + 1) % P.R] + t.instances[0][1:]], proof)


@pytest.mark.parametrize("how", ["break_gate", "break_copy", "break_public_input"])
def test_oracle_verifier_rejects_false_statements(params6, how):
    t = Toy(6, seed=3, n_gate_cols=3, n_lookup_cols=1)
    vk = PL.keygen_vk(params6, t.cs, t.fixed, t.sigma)
    getattr(t, how)()
    assert not PL.verify_proof(params6, t.cs, vk, t.vk_repr, t.instances, _oracle_proof(params6, t))


def test_oracle_prover_rejects_value_outside_table(params6):
    t = Toy(6, seed=4)
    t.advice[t.G][2] = t.tsize + 3
    with pytest.raises(ValueError):
        _oracle_proof(params6, t)


def test_host_transcript_matches_oracle():
    import halo2_vectordb_b200 as h

    rnd = np.random.default_rng(5)
    tr, ot = h.PoseidonTranscript(), T.PoseidonTranscript()
    g = O.g1_generator()
    for step in range(40):
        kind = int(rnd.integers(0, 4))
        if kind == 0:
            v = int(rnd.integers(0, 1 << 62)) * int(rnd.integers(1, 1 << 62)) % P.R
            tr.write_scalar(fr_arr([v])[0])
            ot.write_scalar(v)
        elif kind == 1:
            pt = O.g1_mul(g, int(rnd.integers(1, 1 << 62)))
            tr.write_point(pt)
            ot.write_point(O.g1_affine_to_ints(pt))
        elif kind == 2:
            v = int(rnd.integers(0, 1 << 62))
            tr.common_scalar(fr_arr([v])[0])
            ot.common_scalar(v)
        else:
            assert O.fr_to_ints(tr.squeeze_challenge())[0] == ot.squeeze_challenge()
    assert O.fr_to_ints(tr.squeeze_challenge())[0] == ot.squeeze_challenge()
    assert tr.finalize() == ot.finalize()
    with pytest.raises(ValueError):
        tr.write_point(np.zeros(8, dtype=np.uint64))        # upstream: "Cannot write points at infinity to the transcript"
    # the reader side of the oracle parses what the writer produced
    rd = T.PoseidonTranscript(ot.finalize())
    assert rd.pos == 0 and len(rd.inp) == len(tr.finalize())


def test_point_encoding_round_trip():
    import halo2_vectordb_b200 as h

    g = O.g1_generator()
    pts = [O.g1_mul(g, k) for k in (1, 2, 3, 0xDEADBEEF, P.R - 1)]
    enc = h.g1_to_bytes(np.stack(pts))
    for p_, e in zip(pts, enc):
        ints = O.g1_affine_to_ints(p_)
        assert e == T.g1_to_bytes(ints) and T.g1_from_bytes(e) == ints
    assert h.g1_to_bytes(np.zeros((1, 8), dtype=np.uint64))[0] == bytes(32)


# ----------------------------------------------------------------------------------------------- GPU
def _gpu_setup(h2v, k, t, params):
    srs = h2v.ParamsKZG(k, params.g, params.g_lagrange)
    pk = h2v.ProvingKey(srs, t.cs, [fr_arr(c) for c in t.fixed], [fr_arr(c) for c in t.sigma], fr_arr([t.vk_repr])[0])
    return srs, pk


def _gpu_proof(pk, t, seed=SEED):
    return pk.create_proof([fr_arr(c) for c in t.advice], [fr_arr(c) for c in t.instances], seed)


@pytest.mark.gpu
@pytest.mark.parametrize("k,gate_cols,lookup_cols,degree", [(6, 2, 1, 4), (6, 3, 2, 4), (7, 1, 1, 5), (6, 4, 0, 4), (8, 5, 3, 4), (10, 6, 2, 4)])
def test_gpu_proof_bytes_equal_oracle(h2v, k, gate_cols, lookup_cols, degree):
    t = Toy(k, seed=k * 100 + gate_cols, n_gate_cols=gate_cols, n_lookup_cols=lookup_cols, degree=degree)
    params = PL.Params.setup(k, SECRET)
    srs, pk = _gpu_setup(h2v, k, t, params)
    got = _gpu_proof(pk, t)
    assert len(got) == pk.proof_size()
    want = _oracle_proof(params, t)
    assert got == want
    vk = PL.keygen_vk(params, t.cs, t.fixed, t.sigma)
    assert PL.verify_proof(params, t.cs, vk, t.vk_repr, t.instances, got)
    # a second proof on the same key (workspace reuse), other seed
    got2 = _gpu_proof(pk, t, bytes(32))
    assert got2 != got and PL.verify_proof(params, t.cs, vk, t.vk_repr, t.instances, got2)
    pk.close()
    srs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k,gate_cols,lookup_cols,degree,scr,reverse_gates",
                         [(8, 5, 3, 4, 6, False), (7, 3, 2, 5, 8, False), (10, 6, 2, 4, 7, False), (6, 4, 0, 4, 64, False), (8, 5, 3, 4, 6, True), (9, 7, 1, 4, 9, True)])
def test_gpu_proof_streamed_extended_columns(h2v, monkeypatch, k, gate_cols, lookup_cols, degree, scr, reverse_gates):
    """The k = 20 path: the key keeps no extended-coset columns and evaluate_h rebuilds them in slices of `scr` columns
    (forced here at small k); the proof must be the same bytes as with everything resident, and as the CPU restatement's.
    With the gates in the order of their columns, one pass over the permutation slices folds gates and permutation together
    (each advice column extended once); `reverse_gates` lists the gates backwards, which takes the two separate loops."""
    t = Toy(k, seed=k * 100 + gate_cols, n_gate_cols=gate_cols, n_lookup_cols=lookup_cols, degree=degree)
    if reverse_gates:
        t.cs["gates"] = list(reversed(t.cs["gates"]))
    params = PL.Params.setup(k, SECRET)
    srs, pk = _gpu_setup(h2v, k, t, params)
    resident = _gpu_proof(pk, t)
    pk.close()
    monkeypatch.setenv("H2V_STREAM_EXT", "1")
    monkeypatch.setenv("H2V_STREAM_COLS", str(scr))
    if k != 8:      # how many extended sigma columns the key caches once the first proof's buffers exist: some / none / all that fit
        monkeypatch.setenv("H2V_EXT_CACHE_COLS", "3" if k == 10 else "0")
    pk = h2v.ProvingKey(srs, t.cs, [fr_arr(c) for c in t.fixed], [fr_arr(c) for c in t.sigma], fr_arr([t.vk_repr])[0])
    streamed = _gpu_proof(pk, t)
    assert streamed == resident == _oracle_proof(params, t)
    assert _gpu_proof(pk, t) == streamed          # workspace reuse
    pk.close()
    srs.close()


@pytest.mark.gpu
def test_gpu_proof_k16(h2v):
    """BASELINE configs[2]'s size: k = 16, LOOKUP_BITS = 15 (few columns, so that the CPU restatement finishes)"""
    k = 16
    t = Toy(k, seed=16, n_gate_cols=4, n_lookup_cols=1, lookup_bits=15)
    s = fr_arr([SECRET])[0]
    g, gl = h2v.srs_setup(k, s)              # bases from the device setup (checked against the oracle's in test_gpu_parity)
    params = PL.Params(k, g, gl, SECRET)
    srs, pk = _gpu_setup(h2v, k, t, params)
    got = _gpu_proof(pk, t)
    assert got == _oracle_proof(params, t)
    vk = PL.keygen_vk(params, t.cs, t.fixed, t.sigma)
    assert PL.verify_proof(params, t.cs, vk, t.vk_repr, t.instances, got)
    pk.close()
    srs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("how", ["break_gate", "break_copy", "break_public_input"])
def test_gpu_proof_of_false_statement_is_rejected(h2v, how):
    k = 7
    t = Toy(k, seed=77, n_gate_cols=3, n_lookup_cols=1)
    params = PL.Params.setup(k, SECRET)
    vk = PL.keygen_vk(params, t.cs, t.fixed, t.sigma)
    getattr(t, how)()
    srs, pk = _gpu_setup(h2v, k, t, params)
    proof = _gpu_proof(pk, t)
    assert proof == _oracle_proof(params, t)              # both provers follow the same flow on a bad witness too
    assert not PL.verify_proof(params, t.cs, vk, t.vk_repr, t.instances, proof)
    pk.close()
    srs.close()


@pytest.mark.gpu
def test_gpu_prover_errors(h2v):
    k = 6
    t = Toy(k, seed=9)
    params = PL.Params.setup(k, SECRET)
    srs, pk = _gpu_setup(h2v, k, t, params)
    bad = [list(c) for c in t.advice]
    bad[t.G][2] = t.tsize + 3                              # not in the table: Error::ConstraintSystemFailure upstream
    with pytest.raises(ValueError):
        pk.create_proof([fr_arr(c) for c in bad], [fr_arr(c) for c in t.instances], SEED)
    with pytest.raises(ValueError):                        # InstanceTooLarge
        pk.create_proof([fr_arr(c) for c in t.advice], [fr_arr([1] * (1 << k))], SEED)
    with pytest.raises(ValueError):
        pk.create_proof([fr_arr(c) for c in t.advice[:-1]], [fr_arr(c) for c in t.instances], SEED)
    # the key still works afterwards
    assert _gpu_proof(pk, t) == _oracle_proof(params, t)
    pk.close()
    srs.close()


@pytest.mark.gpu
def test_gpu_proof_of_generated_circuit_verifies(h2v):
    """the vectorised circuit generator bench.py uses for the full-size shapes (halo2_vectordb_b200.synthetic): its
    circuits are satisfied -- the oracle verifier accepts the device proof -- and the oracle prover produces the same bytes"""
    from halo2_vectordb_b200.synthetic import synthetic_circuit

    k = 9
    c = synthetic_circuit(k, n_gate_cols=5, n_lookup_cols=2, lookup_bits=6, seed=3)
    params = PL.Params.setup(k, SECRET)
    srs = h2v.ParamsKZG(k, params.g, params.g_lagrange)
    pk = h2v.ProvingKey(srs, c["cs"], c["fixed"], c["sigma"], c["vk_repr"])
    proof = pk.create_proof(c["advice"], c["instances"], SEED)
    ints = lambda cols: [O.fr_to_ints(col) for col in cols]
    fixed, sigma, advice, inst = ints(c["fixed"]), ints(c["sigma"]), ints(c["advice"]), ints(c["instances"])
    vk_repr = O.fr_to_ints(c["vk_repr"].reshape(1, 4))[0]
    vk = PL.keygen_vk(params, c["cs"], fixed, sigma)
    assert PL.verify_proof(params, c["cs"], vk, vk_repr, inst, proof)
    assert proof == PL.create_proof(params, c["cs"], fixed, sigma, vk_repr, advice, inst, SEED)
    pk.close()
    srs.close()
