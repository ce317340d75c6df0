"""The device verifier (csrc/pairing.cuh) against the Python oracle's pairing
check and the golden Groth16 instance: groth16::verify_proof as the reference
runs it right after proving (masp_proofs/src/sapling/prover.rs:148, :266).

The golden instance (tests/golden/vectors.json "r1cs") is a satisfied R1CS with
a trusted setup whose trapdoor is known, so its proof verifies; it is the same
vector the C++ oracle is pinned on (tests/test_oracle_c.py).
"""
import json
import os

import pytest

from oracle.py import groth16 as g
from oracle.py.bls12_381 import G1, G2, R

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))["r1cs"]
H = bytes.fromhex


def uncompressed(proof_bytes):
    a, b, c = g.proof_read(proof_bytes)
    return G1.encode_uncompressed(a) + G2.encode_uncompressed(b) + G1.encode_uncompressed(c)


def asg_for_batch(pv, v):
    return pv.ProvingAssignment(H(v["a"]), H(v["b"]), H(v["c"]), H(v["inputs"]), H(v["aux"]))


def _check(pv):
    v = GOLD
    dens = (H(v["a_aux_density"]), H(v["b_input_density"]), H(v["b_aux_density"]))
    P = pv.Parameters.read(H(v["params"]), dens)
    inputs = [int.from_bytes(H(v["inputs"])[32 * i:32 * i + 32], "little") for i in range(P.n_inputs)]
    proof = H(v["proof"])
    # the oracle's verdicts first: this proof verifies, a wrong input does not
    params, _ = g.Parameters.read(H(v["params"]))
    assert g.verify_proof(params.vk, g.proof_read(proof), inputs[1:])
    wrong = [inputs[1], (inputs[2] + 1) % R]
    other = uncompressed(proof)
    swapped = other[288:] + other[96:288] + other[:96]  # A and C exchanged: well-formed points, wrong proof
    got = pv.verify_batch(P, [other, other, swapped], [inputs[1:], wrong, inputs[1:]])
    assert got == [True, False, False]
    assert g.verify_proof(params.vk, g.proof_read(proof), wrong) is False
    # prove-then-verify: the self-check passes for an honest witness ...
    asg = pv.ProvingAssignment(H(v["a"]), H(v["b"]), H(v["c"]), H(v["inputs"]), H(v["aux"]))
    pv.set_option("verify", 1)
    try:
        assert pv.create_proof(asg, P, H(v["r"]), H(v["s"])) == proof
        # ... and fails (the reference's Err(()) at prover.rs:148) when the public inputs
        # handed to the prover are not the ones the witness satisfies
        bad_inputs = bytearray(H(v["inputs"]))
        bad_inputs[32:64] = ((inputs[1] + 1) % R).to_bytes(32, "little")
        bad = pv.ProvingAssignment(H(v["a"]), H(v["b"]), H(v["c"]), bytes(bad_inputs), H(v["aux"]))
        with pytest.raises(pv.Mb200Error) as e:
            pv.create_proof(bad, P, H(v["r"]), H(v["s"]))
        assert e.value.code == -8
    finally:
        pv.set_option("verify", 0)
    # bilinearity, as a property of the device pairing: (k A, k^-1 B, C) satisfies the same equation
    a_pt, b_pt, c_pt = g.proof_read(proof)
    k = 0x1D2C3B4A5968778695A4B3C2D1E0F
    ka = G1.to_affine(G1.mul(G1.from_affine(a_pt), k))
    kb = G2.to_affine(G2.mul(G2.from_affine(b_pt), pow(k, R - 2, R)))
    kb_bad = G2.to_affine(G2.mul(G2.from_affine(b_pt), pow(k + 1, R - 2, R)))
    enc = lambda A, B, Cc: G1.encode_uncompressed(A) + G2.encode_uncompressed(B) + G1.encode_uncompressed(Cc)
    assert pv.verify_batch(P, [enc(ka, kb, c_pt), enc(ka, kb_bad, c_pt)], [inputs[1:]] * 2) == [True, False]
    # proofs in wire form: Proof::read (decompression, subgroup checks) happens on the device
    flipped = bytearray(proof)
    flipped[0] ^= 0x20                          # the other square root for A: a valid point, wrong proof
    off_curve = bytearray(proof)
    off_curve[47] ^= 0x01                       # x of A changed: almost surely no point / not in the subgroup
    no_flag = bytes([proof[0] & 0x7F]) + proof[1:]   # compression flag cleared: does not read
    inf_a = bytes([0xC0]) + bytes(47) + proof[48:]   # A = identity: reads, cannot verify
    got = pv.verify_proofs(P, [proof, bytes(flipped), bytes(off_curve), no_flag, inf_a, proof],
                           [inputs[1:]] * 5 + [wrong])
    assert got == [True, False, False, False, False, False]
    # the randomised batch check: all-or-nothing, one final exponentiation for the batch
    second = pv.create_proof(asg_for_batch(pv, v), P, (5).to_bytes(32, "little"), (6).to_bytes(32, "little"))
    assert second != proof
    assert pv.verify_proofs_batch(P, [proof, second, proof], [inputs[1:]] * 3) is True
    assert pv.verify_proofs_batch(P, [proof, second, bytes(flipped)], [inputs[1:]] * 3) is False
    assert pv.verify_proofs_batch(P, [proof, second], [inputs[1:], wrong]) is False
    assert pv.verify_proofs_batch(P, [no_flag, proof], [inputs[1:]] * 2) is False
    assert pv.verify_proofs_batch(P, [second], [inputs[1:]], rng=lambda k: b"\x01" + bytes(k - 1)) is True
    # malformed proof encodings are an error, not a verdict
    with pytest.raises(pv.Mb200Error):
        pv.verify_batch(P, [b"\xff" * 384], [inputs[1:]])


@pytest.mark.slow
def test_verifier_emulated(emu):
    _check(emu)


@pytest.mark.gpu
def test_verifier_gpu(gpu):
    _check(gpu)


@pytest.mark.gpu
def test_self_check_on_real_output_circuit_gpu(gpu):
    """Prove-then-verify on the real Output circuit, end to end on the device."""
    from masp_b200 import circuits as C
    from test_circuits import real_instance
    cs, key, dens, w = real_instance("output")
    c = C.Circuit(C.OUTPUT)
    P = gpu.Parameters.read(key, dens).bind_circuit(c)
    gpu.set_option("verify", 1)
    try:
        proofs = gpu.create_proof_batch_from_witness(P, w["inputs"] * 3, w["aux"] * 3, [1, 2, 3], [4, 5, 6])
        assert len(set(proofs)) == 3
        bad_aux = bytearray(w["aux"])
        bad_aux[32 * 500:32 * 501] = (12345).to_bytes(32, "little")  # an unsatisfied witness
        with pytest.raises(gpu.Mb200Error) as e:
            gpu.create_proof_batch_from_witness(P, w["inputs"], bytes(bad_aux), [1], [4])
        assert e.value.code == -8
    finally:
        gpu.set_option("verify", 0)
    ok = gpu.verify_batch(P, [uncompressed(p) for p in proofs], [cs.inputs[1:]] * 3)
    assert ok == [True, True, True]
