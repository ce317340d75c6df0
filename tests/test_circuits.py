"""The MASP circuits restated in the oracle against the reference's own pins
(the only golden values the reference holds near this path):

  Convert  47 358 constraints, 4 inputs, hash f74b47ef...f814   masp_proofs/src/circuit/convert.rs:218-224
  Output   31 205 constraints, 6 inputs, hash 93e445d7...276a   sapling.rs:1024-1045, docs/protocol.tex:3157
  Spend   100 637 constraints, 8 inputs, hash 34e4a634...eb3d   sapling.rs:730-741

and, end to end, a proof of the real Output circuit under a trusted setup with
a known trapdoor, checked with the pairing as the reference does after proving
(masp_proofs/src/sapling/prover.rs:148).
"""
import random

import pytest

from masp_b200 import synthetic as syn
from oracle.py.bls12_381 import R
from oracle.py.r1cs_gadgets import AllocatedBit, Boolean, ConstraintSystem, blake2s
from oracle.py import groth16 as g
from oracle.py import masp_circuits as mc

RND = random.Random(20261017)
PATH = [(RND.randrange(R), bool(RND.getrandbits(1))) for _ in range(32)]
VALUE, RCV, RCM, AR, ESK, NSK = (RND.getrandbits(64), RND.randrange(mc.JUBJUB_ORDER), RND.randrange(mc.JUBJUB_ORDER),
                                 RND.randrange(mc.JUBJUB_ORDER), RND.randrange(mc.JUBJUB_ORDER),
                                 RND.randrange(mc.JUBJUB_ORDER))
G_D = mc.jj_mul(mc.PROOF_GENERATION_KEY_GENERATOR, 777)
AK = mc.jj_mul(mc.SPENDING_KEY_GENERATOR, 4242)


def test_jubjub_constants():
    pts = mc.PEDERSEN_HASH_GENERATORS + [mc.VALUE_COMMITMENT_RANDOMNESS_GENERATOR, mc.SPENDING_KEY_GENERATOR,
                                         mc.PROOF_GENERATION_KEY_GENERATOR, mc.NOTE_COMMITMENT_RANDOMNESS_GENERATOR,
                                         mc.NULLIFIER_POSITION_GENERATOR]
    assert all(mc.jj_on_curve(p) for p in pts)
    assert all(mc.jj_mul(p, mc.JUBJUB_ORDER) == mc.JJ_IDENTITY for p in pts)
    # masp_proofs/src/constants.rs:189-205: d = -(10240/10241), scale = sqrt(4 / (a - d)) with a = -1
    assert mc.EDWARDS_D == (-10240 * pow(10241, R - 2, R)) % R
    assert mc.MONTGOMERY_SCALE * mc.MONTGOMERY_SCALE % R * ((-1 - mc.EDWARDS_D) % R) % R == 4


def test_blake2s_gadget_matches_hashlib():
    import hashlib
    for n in (32, 64):
        cs = ConstraintSystem()
        msg = bytes(RND.getrandbits(8) for _ in range(n))
        bits = [Boolean.from_bit(AllocatedBit.alloc(cs, b)) for b in mc.bytes_to_bits_le(msg)]
        out = blake2s(cs, bits, b"MASP__nf")
        got = mc.bits_to_bytes_le([b.value for b in out])
        assert got == hashlib.blake2s(msg, digest_size=32, person=b"MASP__nf").digest()
        assert cs.is_satisfied()
        assert cs.num_constraints() - n * 8 == 21006  # 21 006 constraints per call


@pytest.mark.slow
def test_convert_circuit_pins():
    ag = mc.jj_mul(mc.PROOF_GENERATION_KEY_GENERATOR, 12345)
    anchor = mc.convert_native_anchor(ag, PATH)
    cs = ConstraintSystem()
    mc.convert_circuit(cs, ag, VALUE, RCV, PATH, anchor)
    assert cs.is_satisfied()
    assert cs.num_constraints() == 47358
    assert cs.hash() == "f74b47ef6e59081548f81f5806bd15b1f4a65d2e57681e6db2b8db7eef2ff814"
    assert cs.num_inputs() == 4 and len(cs.aux) == 47322
    cv = mc.jj_add(mc.jj_mul(mc.jj_mul(ag, 8), VALUE), mc.jj_mul(mc.VALUE_COMMITMENT_RANDOMNESS_GENERATOR, RCV))
    assert cs.inputs == [1, cv[0], cv[1], anchor]
    dens = cs.proving_assignment()[3]
    sh = syn.CONVERT  # the shape the benchmark uses is this circuit's
    assert (sum(dens[0]), sum(dens[1]), sum(dens[2])) == (sh.a_dense, 1, sh.b_dense)
    bad = ConstraintSystem()
    mc.convert_circuit(bad, ag, VALUE, RCV, PATH, (anchor + 1) % R)
    assert not bad.is_satisfied()


@pytest.mark.slow
def test_output_circuit_pins():
    ident, ag = mc.find_asset()
    pk_d = mc.jj_mul(G_D, 999)
    cs = ConstraintSystem()
    mc.output_circuit(cs, mc.bytes_to_bits_le(ident), ag, VALUE, RCV, G_D, pk_d, RCM, ESK)
    assert cs.is_satisfied()
    assert cs.num_constraints() == 31205
    assert cs.hash() == "93e445d7858e98c7138558df341f020aedfe75893535025587d64731e244276a"
    assert cs.num_inputs() == 6 and len(cs.aux) == 30896
    nat = mc.output_native(ag, VALUE, RCV, G_D, pk_d, RCM, ESK)
    assert cs.inputs == [1, nat["cv"][0], nat["cv"][1], nat["epk"][0], nat["epk"][1], nat["cmu"]]
    dens = cs.proving_assignment()[3]
    assert (sum(dens[0]), sum(dens[1]), sum(dens[2])) == (syn.OUTPUT.a_dense, 1, syn.OUTPUT.b_dense)


@pytest.mark.slow
def test_spend_circuit_pins():
    _, ag = mc.find_asset()
    nat = mc.spend_native(AK, NSK, G_D, ag, VALUE, RCV, RCM, AR, PATH)
    cs = ConstraintSystem()
    mc.spend_circuit(cs, AK, NSK, G_D, ag, VALUE, RCV, RCM, AR, PATH, nat["anchor"])
    assert cs.is_satisfied()
    assert cs.num_constraints() == 100637
    assert cs.hash() == "34e4a634c80e4e4c6250e63b7855532e60b36d1371d4d7b1163218b69f09eb3d"
    assert cs.num_inputs() == 8 and len(cs.aux) == 100497
    want = [1, nat["rk"][0], nat["rk"][1], nat["cv"][0], nat["cv"][1], nat["anchor"]] + \
        mc.multipack_inputs(mc.bytes_to_bits_le(nat["nf"]))
    assert cs.inputs == want
    dens = cs.proving_assignment()[3]
    assert (sum(dens[0]), sum(dens[1]), sum(dens[2])) == (syn.SPEND.a_dense, 1, syn.SPEND.b_dense)


def real_instance(name):
    """(cs, key bytes, densities, assignment byte strings) for a real MASP circuit
    with a native witness and a trusted setup whose trapdoor is known."""
    from oracle.setup import generate_parameters_bytes
    ident, ag = mc.find_asset()
    cs = ConstraintSystem()
    if name == "output":
        mc.output_circuit(cs, mc.bytes_to_bits_le(ident), ag, VALUE, RCV, G_D, mc.jj_mul(G_D, 999), RCM, ESK)
    elif name == "convert":
        mc.convert_circuit(cs, ag, VALUE, RCV, PATH, mc.convert_native_anchor(ag, PATH))
    else:
        nat = mc.spend_native(AK, NSK, G_D, ag, VALUE, RCV, RCM, AR, PATH)
        mc.spend_circuit(cs, AK, NSK, G_D, ag, VALUE, RCV, RCM, AR, PATH, nat["anchor"])
    td = g.Trapdoor(0x1111111111222233334444, 0x5555AAAA, 0x7777BBBBCC, 0x99990000111, 0x1234567890ABCDEF)
    key, dens = generate_parameters_bytes(cs, td)
    a, b, c, d2 = cs.proving_assignment()
    assert dens == (syn.pack_bits(d2[0]), syn.pack_bits(d2[1]), syn.pack_bits(d2[2]))
    ib = syn.ints_to_bytes
    w = {"a": ib(a), "b": ib(b), "c": ib(c), "inputs": ib(cs.inputs), "aux": ib(cs.aux),
         "r": ib([0xDEADBEEFCAFEBABE1234]), "s": ib([0xFEEDFACE5678])}
    return cs, key, dens, w


def real_output_instance():
    return real_instance("output")


def verify_with_pairing(key, proof_bytes, public_inputs):
    from oracle.py.bls12_381 import G1, G2
    vk = g.VerifyingKey(*[G1.decode_uncompressed(key[o:o + 96]) if n == 96 else G2.decode_uncompressed(key[o:o + 192])
                          for o, n in ((0, 96), (96, 96), (192, 192), (384, 192), (576, 96), (672, 192))], ic=None)
    n_ic = int.from_bytes(key[864:868], "big")
    vk.ic = [G1.decode_uncompressed(key[868 + 96 * i:868 + 96 * (i + 1)]) for i in range(n_ic)]
    return g.verify_proof(vk, g.proof_read(proof_bytes), public_inputs)


@pytest.mark.slow
def test_real_output_circuit_proof_verifies_cpu(oracle):
    cs, key, dens, w = real_output_instance()
    assert len(key) == syn.OUTPUT.params_file_bytes()  # the real key's size minus the MPC tail
    P = oracle.Params(key, len(cs.aux), *dens)
    proof = P.prove(len(cs.constraints) + len(cs.inputs), w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["r"], w["s"])
    assert verify_with_pairing(key, proof, cs.inputs[1:])
    wrong = list(cs.inputs[1:])
    wrong[4] = (wrong[4] + 1) % R
    assert not verify_with_pairing(key, proof, wrong)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["output", "convert", "spend"])
def test_real_circuit_proofs_verify_gpu(gpu, oracle, name):
    """The CUDA path on the real MASP circuits (real witnesses, real density
    positions): byte-identical to the CPU oracle and accepted by the Groth16
    verification equation, as the reference checks after proving
    (masp_proofs/src/sapling/prover.rs:148, :266)."""
    cs, key, dens, w = real_instance(name)
    sh = syn.SHAPES[name]
    assert len(key) == sh.params_file_bytes()
    P = gpu.Parameters.read(key, dens)
    assert (P.n_inputs, P.n_aux, P.a_len, P.b_len, P.m) == (sh.n_inputs, sh.n_aux, sh.a_len, sh.b_len, sh.m)
    asg = gpu.ProvingAssignment(w["a"], w["b"], w["c"], w["inputs"], w["aux"])
    proof = gpu.create_proof(asg, P, w["r"], w["s"])
    ref = oracle.Params(key, len(cs.aux), *dens)
    assert proof == ref.prove(asg.rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["r"], w["s"])
    assert verify_with_pairing(key, proof, cs.inputs[1:])
    wrong = list(cs.inputs[1:])
    wrong[-1] = (wrong[-1] + 1) % R
    assert not verify_with_pairing(key, proof, wrong)
