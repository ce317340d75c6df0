"""The product-side circuit layer (csrc/host/*.hpp behind mb200_circuit_*)
against the reference's own pins and against the Python oracle.

  * structure: TestConstraintSystem::hash, constraint and input counts of the
    recorded Spend / Output / Convert circuits equal the strings pinned at
    masp_proofs/src/circuit/sapling.rs:730-741, 1024-1045 and convert.rs:218-224;
    the density counts equal the real keys' query lengths;
  * witnesses: inputs and aux byte-identical to the oracle's synthesis of the
    same instance; the public inputs are the natively computed cv / rk / nf /
    cmu / epk / anchor (the checks of sapling.rs:743-759, 1045-1065);
  * rows: a = A z, b = B z, c = C z from the r1cs_eval kernel equal the
    oracle's ProvingAssignment (host-emulated here, on the B200 under -m gpu);
  * end to end on the GPU: witness-only proofs are byte-identical to the oracle
    prover and accepted by the pairing check.
"""
import random

import pytest

from masp_b200 import circuits as C
from masp_b200 import synthetic as syn
from oracle.py.bls12_381 import R
from oracle.py.r1cs_gadgets import ConstraintSystem
from oracle.py import masp_circuits as mc

RND = random.Random(77)
IDENT, AG = mc.find_asset()
G_D = mc.jj_mul(mc.PROOF_GENERATION_KEY_GENERATOR, 777)
AK = mc.jj_mul(mc.SPENDING_KEY_GENERATOR, 4242)
ib = syn.ints_to_bytes


def jscalar():
    return RND.randrange(mc.JUBJUB_ORDER)


def make_instance(kind, depth):
    """(reference-style circuit instance, oracle ConstraintSystem synthesised from the same witness)."""
    path = [(RND.randrange(R), bool(RND.getrandbits(1))) for _ in range(depth)]
    value, rcv, rcm = RND.getrandbits(64), jscalar(), jscalar()
    vc = C.ValueCommitmentOpening(AG, value, rcv)
    cs = ConstraintSystem()
    if kind == C.CONVERT:
        anchor = mc.convert_native_anchor(AG, path)
        mc.convert_circuit(cs, AG, value, rcv, path, anchor)
        return C.Convert(vc, path, anchor), cs
    if kind == C.OUTPUT:
        pk_d, esk = mc.jj_mul(G_D, RND.randrange(1, 1 << 60)), jscalar()
        mc.output_circuit(cs, mc.bytes_to_bits_le(IDENT), AG, value, rcv, G_D, pk_d, rcm, esk)
        return C.Output(vc, IDENT, G_D, pk_d, rcm, esk), cs
    nsk, ar = jscalar(), jscalar()
    nat = mc.spend_native(AK, nsk, G_D, AG, value, rcv, rcm, ar, path)
    mc.spend_circuit(cs, AK, nsk, G_D, AG, value, rcv, rcm, ar, path, nat["anchor"])
    return C.Spend(vc, AK, nsk, G_D, rcm, ar, path, nat["anchor"]), cs


def test_pedersen_hash_reference_vectors():
    """The reference's own golden vectors for the Pedersen hash (37 vectors,
    masp_primitives/src/test_vectors/pedersen_hash_vectors.rs, checked there by
    sapling/pedersen_hash.rs:131-152; fixture made by tests/golden/make_pedersen_vectors.py):
    the product's window tables, Montgomery chains with shared inversions and Edwards maps, and
    the oracle's independent native implementation, both reproduce every one of them."""
    import json
    import os
    vecs = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pedersen_hash_vectors.json")))
    assert len(vecs) == 37
    for v in vecs:
        bits = [ch == "1" for ch in v["input_bits"]]
        want = (int(v["hash_u"], 16), int(v["hash_v"], 16))
        assert C.pedersen_hash(bits[:6], bits[6:]) == want, v["personalization"]
        assert mc.pedersen_hash_native(bits[:6], bits[6:]) == want, v["personalization"]


@pytest.mark.parametrize("kind,shape", [(C.SPEND, syn.SPEND), (C.OUTPUT, syn.OUTPUT), (C.CONVERT, syn.CONVERT)])
def test_recorded_circuits_match_reference_pins(kind, shape):
    c = C.Circuit(kind)
    n_cons, n_in, h = C.PINS[kind]
    assert (c.n_constraints, c.n_inputs, c.hash()) == (n_cons, n_in, h)
    assert (c.n_aux, c.a_dense, c.b_input_dense, c.b_dense) == (shape.n_aux, shape.a_dense, 1, shape.b_dense)
    a_d, bi_d, ba_d = c.densities()
    assert (sum(bin(x).count("1") for x in a_d), sum(bin(x).count("1") for x in ba_d)) == (shape.a_dense, shape.b_dense)
    assert bi_d[0] == 1  # only ONE appears in a B row


@pytest.mark.parametrize("kind,depth", [(C.CONVERT, 2), (C.OUTPUT, 0), (C.SPEND, 1)])
def test_witness_matches_oracle(kind, depth):
    c = C.Circuit(kind, depth)
    insts = [make_instance(kind, depth) for _ in range(2)]
    inputs, aux = c.synthesize([i for i, _ in insts], threads=2)
    for k, (_, cs) in enumerate(insts):
        assert cs.is_satisfied()
        assert (c.n_constraints, c.n_inputs, c.n_aux) == (cs.num_constraints(), cs.num_inputs(), len(cs.aux))
        assert c.hash() == cs.hash()
        assert inputs[32 * c.n_inputs * k:32 * c.n_inputs * (k + 1)] == ib(cs.inputs)
        assert aux[32 * c.n_aux * k:32 * c.n_aux * (k + 1)] == ib(cs.aux)
    # the recorded matrices are the oracle's constraints
    cs = insts[0][1]
    dens = cs.proving_assignment()[3]
    assert c.densities() == (syn.pack_bits(dens[0]), syn.pack_bits(dens[1]), syn.pack_bits(dens[2]))
    rp, col, val = c.matrix(1)
    row = len(cs.constraints) // 2
    want = {(0 if k == "I" else 0x80000000) | i: cf for (k, i), cf in cs._canonical(cs.constraints[row][1])}
    assert {col[e]: val[e] for e in range(rp[row], rp[row + 1])} == want


@pytest.mark.parametrize("value,rights", [(0, 0b00), (2 ** 64 - 1, 0b11), (1, 0b01)])
def test_edge_values_match_oracle(value, rights):
    """Extremes the reference's own circuit tests walk (value 0 / u64::MAX, every
    position bit pattern: circuit/sapling.rs:640-700 loops over tree positions)."""
    depth = 2
    c = C.Circuit(C.CONVERT, depth)
    path = [(RND.randrange(R), bool((rights >> i) & 1)) for i in range(depth)]
    rcv = mc.JUBJUB_ORDER - 1  # largest scalar
    anchor = mc.convert_native_anchor(AG, path)
    cs = ConstraintSystem()
    mc.convert_circuit(cs, AG, value, rcv, path, anchor)
    assert cs.is_satisfied()
    inst = C.Convert(C.ValueCommitmentOpening(AG, value, rcv), path, anchor)
    assert c.root(inst) == anchor
    inputs, aux = c.synthesize([inst])
    assert inputs == ib(cs.inputs) and aux == ib(cs.aux)
    if value == 0:
        # value 0: the anchor check `(cur - rt) * value = 0` holds for ANY anchor (convert.rs:113-121)
        free = C.Convert(C.ValueCommitmentOpening(AG, 0, rcv), path, (anchor + 5) % R)
        i2, a2 = c.synthesize([free])
        cs2 = ConstraintSystem()
        mc.convert_circuit(cs2, AG, 0, rcv, path, (anchor + 5) % R)
        assert cs2.is_satisfied() and i2 == ib(cs2.inputs) and a2 == ib(cs2.aux)


def test_empty_batch_and_threads():
    c = C.Circuit(C.OUTPUT)
    assert c.synthesize([]) == (b"", b"")
    inst, cs = make_instance(C.OUTPUT, 0)
    one = c.synthesize([inst], threads=1)
    many = c.synthesize([inst] * 5, threads=3)
    assert many[0] == one[0] * 5 and many[1] == one[1] * 5


def test_bad_witnesses_are_rejected():
    c = C.Circuit(C.CONVERT, 1)
    inst, _ = make_instance(C.CONVERT, 1)
    from masp_b200._lib import Mb200Error
    w = bytearray(inst.pack())
    w[0:32] = R.to_bytes(32, "little")        # non-canonical field element
    with pytest.raises(Mb200Error) as e:
        c.synthesize([bytes(w)])
    assert e.value.code == -6
    # a small-order asset generator: bellman's assert_nonzero returns DivisionByZero,
    # the reference then panics at sapling/prover.rs:252; here it is an error code
    inst.value_commitment.asset_generator = (0, 1)
    with pytest.raises(Mb200Error) as e:
        c.synthesize([inst])
    assert e.value.code == -7
    with pytest.raises(ValueError):
        c.synthesize([inst.pack()[:-1]])



def _cheap_instance(kind, depth, rnd):
    """A witness without the Python oracle (its anchor is arbitrary: witness generation does not need a satisfied circuit)."""
    js = lambda: rnd.randrange(mc.JUBJUB_ORDER)
    path = [(rnd.randrange(R), bool(rnd.getrandbits(1))) for _ in range(depth)]
    vc = C.ValueCommitmentOpening(AG, rnd.getrandbits(64), js())
    if kind == C.CONVERT:
        return C.Convert(vc, path, rnd.randrange(R))
    if kind == C.OUTPUT:
        return C.Output(vc, IDENT, G_D, mc.jj_mul(G_D, rnd.randrange(1, 1 << 60)), js(), js())
    return C.Spend(vc, AK, js(), G_D, js(), js(), path, rnd.randrange(R))


@pytest.mark.parametrize("kind,depth,n", [(C.SPEND, 32, 11), (C.OUTPUT, 0, 9), (C.CONVERT, 32, 17), (C.CONVERT, 3, 2)])
def test_simd_witnesses_equal_scalar_witnesses(kind, depth, n):
    """The eight-lane AVX-512 IFMA generator (csrc/circuits_simd.cpp: batches of two or more) against the
    one-witness scalar generator (a batch of one always takes it): same inputs and aux, byte for byte,
    including the padded last group and lanes holding the extreme values."""
    from masp_b200._lib import lib
    if not lib().mb200_circuit_simd():
        pytest.skip("this CPU has no AVX-512 IFMA: the scalar generator is the only one")
    rnd = random.Random(1000 * kind + n)
    c = C.Circuit(kind, depth)
    insts = [_cheap_instance(kind, depth, rnd) for _ in range(n)]
    insts[1].value_commitment.value = 0
    insts[n - 1].value_commitment.value = 2 ** 64 - 1
    insts[0].value_commitment.randomness = mc.JUBJUB_ORDER - 1
    inputs, aux = c.synthesize(insts, threads=2)
    for k, inst in enumerate(insts):
        i1, a1 = c.synthesize([inst])
        assert inputs[32 * c.n_inputs * k:32 * c.n_inputs * (k + 1)] == i1, k
        assert aux[32 * c.n_aux * k:32 * c.n_aux * (k + 1)] == a1, k


def test_simd_lane_errors_are_reported():
    """One bad lane fails the call with the scalar path's code, wherever in the group it sits."""
    from masp_b200._lib import lib, Mb200Error
    if not lib().mb200_circuit_simd():
        pytest.skip("this CPU has no AVX-512 IFMA")
    rnd = random.Random(31)
    c = C.Circuit(C.CONVERT, 2)
    for lane in (0, 5, 8):
        packed = [_cheap_instance(C.CONVERT, 2, rnd).pack() for _ in range(10)]
        w = bytearray(packed[lane])
        w[0:32] = R.to_bytes(32, "little")
        packed[lane] = bytes(w)
        with pytest.raises(Mb200Error) as e:
            c.synthesize(packed)
        assert e.value.code == -6
        insts = [_cheap_instance(C.CONVERT, 2, rnd) for _ in range(10)]
        insts[lane].value_commitment.asset_generator = (0, 1)   # small order: assert_nonzero fails
        with pytest.raises(Mb200Error) as e:
            c.synthesize(insts)
        assert e.value.code == -7


def _rows_check(circ_mod, kind, depth):
    c = circ_mod.Circuit(kind, depth)
    insts = [make_instance(kind, depth) for _ in range(2)]
    inputs, aux = c.synthesize([i for i, _ in insts])
    a, b, cc = c.rows_on_device(inputs, aux, len(insts))
    for k, (_, cs) in enumerate(insts):
        wa, wb, wc, _ = cs.proving_assignment()
        n = 32 * c.rows
        assert a[n * k:n * (k + 1)] == ib(wa)
        assert b[n * k:n * (k + 1)] == ib(wb)
        assert cc[n * k:n * (k + 1)] == ib(wc)


def test_rows_kernel_emulated(emu):
    _rows_check(emu.circuits, C.CONVERT, 1)
    _rows_check(emu.circuits, C.OUTPUT, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,depth", [(C.CONVERT, 3), (C.OUTPUT, 0), (C.SPEND, 2)])
def test_rows_kernel_gpu(gpu, kind, depth):
    _rows_check(C, kind, depth)


@pytest.mark.gpu
@pytest.mark.parametrize("name,kind", [("output", C.OUTPUT), ("convert", C.CONVERT), ("spend", C.SPEND)])
def test_witness_only_proofs_gpu(gpu, oracle, name, kind):
    """The whole drop-in path on the real circuits at full size: product-side
    witness generation (host) -> rows on the device -> CUDA prover, against the
    oracle prover fed with the oracle's own synthesis, and the pairing check the
    reference runs after proving (masp_proofs/src/sapling/prover.rs:148, :266)."""
    from test_circuits import real_instance, verify_with_pairing, VALUE, RCV, RCM, AR, ESK, NSK, PATH
    from test_circuits import G_D as GD, AK as AK_
    cs, key, dens, w = real_instance(name)
    c = C.Circuit(kind)
    assert c.densities() == dens
    ident, ag = mc.find_asset()
    vc = C.ValueCommitmentOpening(ag, VALUE, RCV)
    if name == "output":
        inst = C.Output(vc, ident, GD, mc.jj_mul(GD, 999), RCM, ESK)
    elif name == "convert":
        inst = C.Convert(vc, PATH, mc.convert_native_anchor(ag, PATH))
    else:
        nat = mc.spend_native(AK_, NSK, GD, ag, VALUE, RCV, RCM, AR, PATH)
        inst = C.Spend(vc, AK_, NSK, GD, RCM, AR, PATH, nat["anchor"])
    inputs, aux = c.synthesize([inst, inst])
    assert inputs[:32 * c.n_inputs] == w["inputs"] and aux[:32 * c.n_aux] == w["aux"]
    P = gpu.Parameters.read(key, c.densities()).bind_circuit(c)
    r, s = int.from_bytes(w["r"], "little"), int.from_bytes(w["s"], "little")
    proofs = gpu.create_proof_batch_from_witness(P, inputs, aux, [r, s], [s, r])
    ref = oracle.Params(key, len(cs.aux), *dens)
    rows = len(cs.constraints) + len(cs.inputs)
    assert proofs[0] == ref.prove(rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["r"], w["s"])
    assert proofs[1] == ref.prove(rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["s"], w["r"])
    assert all(verify_with_pairing(key, p, cs.inputs[1:]) for p in proofs)
    wrong = list(cs.inputs[1:])
    wrong[0] = (wrong[0] + 1) % R
    assert not verify_with_pairing(key, proofs[0], wrong)


@pytest.mark.gpu
def test_local_tx_prover_with_circuit_instances_gpu(gpu, oracle):
    """LocalTxProver::from_bytes(spend, output, convert) with the reference's signature (no density
    argument: the library's recorded circuits supply them) and *_proof called with circuit
    instances, Output and Convert under keys with a known trapdoor so that the proofs verify."""
    from test_circuits import real_instance, verify_with_pairing, VALUE, RCV, RCM, ESK, PATH
    from test_circuits import G_D as GD
    cs_o, key_o, _, _ = real_instance("output")
    cs_c, key_c, _, _ = real_instance("convert")
    spend_key = gpu.params_synthesize(syn.SPEND)  # a key of the right shape; Spend is not proved here
    prover = gpu.LocalTxProver.from_bytes(spend_key, key_o, key_c, verify_hashes=False)
    assert prover.output_params.circuit.hash() == C.PINS[C.OUTPUT][2]
    ident, ag = mc.find_asset()
    vc = C.ValueCommitmentOpening(ag, VALUE, RCV)
    out_inst = C.Output(vc, ident, GD, mc.jj_mul(GD, 999), RCM, ESK)
    conv_inst = C.Convert(vc, PATH, mc.convert_native_anchor(ag, PATH))
    p_out = prover.output_proof(out_inst)
    p_conv = prover.convert_proof(conv_inst)       # runs verify_proof on the device, like the reference
    assert verify_with_pairing(key_o, p_out, cs_o.inputs[1:])
    assert verify_with_pairing(key_c, p_conv, cs_c.inputs[1:])
    assert gpu.verify_proofs(prover.convert_params, [p_conv, p_out], [cs_c.inputs[1:]] * 2) == [True, False]
    # a convert instance whose anchor is not the path's root cannot be proved: Err(()) in the reference
    bad = C.Convert(vc, PATH, (conv_inst.anchor + 1) % R)
    with pytest.raises(gpu.Mb200Error) as e:
        prover.convert_proof(bad)
    assert e.value.code == -8
    _, conv, outs = prover.prove_bundle(converts=[conv_inst, conv_inst], outputs=[out_inst])
    assert all(verify_with_pairing(key_c, p, cs_c.inputs[1:]) for p in conv) and len(set(conv)) == 2
    assert verify_with_pairing(key_o, outs[0], cs_o.inputs[1:])
