"""SaplingProvingContext / TxProver (masp_b200/sapling.py), the callers of the proving path.

Pinned on values the reference itself holds (tests/golden/sapling_vectors.json, extracted by
tests/golden/make_sapling_vectors.py):

  * the 11 fixed generators of masp_primitives/src/constants.rs, re-derived with find_group_hash
    exactly as the reference's tests do (constants.rs:323-375);
  * HEX_EMPTY_ROOTS, the 33 empty roots of the commitment tree (merkle_tree.rs:912-946), and the 16
    commitments / roots of test_sapling_tree (merkle_tree.rs:1091-1135);
  * the note commitments of the note-encryption vectors (sapling/note_encryption.rs:1357-1361);
  * the ZIP 32 key vectors (zip32/sapling.rs:1372-2135): ask, nsk -> ak, nk -> ivk, found diversifiers.

and cross-checked against the product's own circuits: the public inputs this module computes
natively (as sapling/prover.rs:121-145 does) equal the input assignment the Spend / Output / Convert
witness generators produce for the same instance.  Under -m gpu the whole TxProver surface runs on
the device: spend_proof / output_proof / convert_proof / binding_sig, serial and batched.
"""
import json
import os
import random

import pytest

from masp_b200 import circuits as C
from masp_b200 import sapling as S

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sapling_vectors.json")))
RND = random.Random(2024)


def pt(pair):
    return (int(pair[0], 16), int(pair[1], 16))


def test_fixed_generators_rederived_like_the_reference():
    g = GOLD["generators"]
    assert S.find_group_hash(b"", S.PROOF_GENERATION_KEY_BASE_GENERATOR_PERSONALIZATION) == \
        pt(g["proof_generation_key_generator"]) == S.PROOF_GENERATION_KEY_GENERATOR
    assert S.find_group_hash(b"r", S.PEDERSEN_HASH_GENERATORS_PERSONALIZATION) == \
        pt(g["note_commitment_randomness_generator"]) == S.NOTE_COMMITMENT_RANDOMNESS_GENERATOR
    assert S.find_group_hash(b"", S.NULLIFIER_POSITION_IN_TREE_GENERATOR_PERSONALIZATION) == \
        pt(g["nullifier_position_generator"]) == S.NULLIFIER_POSITION_GENERATOR
    assert S.find_group_hash(b"r", S.VALUE_COMMITMENT_RANDOMNESS_PERSONALIZATION) == \
        pt(g["value_commitment_randomness_generator"]) == S.VALUE_COMMITMENT_RANDOMNESS_GENERATOR
    assert S.find_group_hash(b"", S.SPENDING_KEY_GENERATOR_PERSONALIZATION) == \
        pt(g["spending_key_generator"]) == S.SPENDING_KEY_GENERATOR
    for m, want in enumerate(g["pedersen_hash_generators"]):
        assert S.find_group_hash(m.to_bytes(4, "little"), S.PEDERSEN_HASH_GENERATORS_PERSONALIZATION) == pt(want)


def test_point_encoding_round_trip_and_rejections():
    for k in (1, 2, 12345, S.JUBJUB_ORDER - 1):
        p = S.jj_mul(S.SPENDING_KEY_GENERATOR, k)
        assert S.jj_on_curve(p) and S.jj_from_bytes(S.jj_to_bytes(p)) == p
        assert S.jj_add(p, S.jj_neg(p)) == S.IDENTITY
    assert S.jj_mul(S.SPENDING_KEY_GENERATOR, S.JUBJUB_ORDER) == S.IDENTITY
    assert S.jj_from_bytes(S.jj_to_bytes(S.IDENTITY)) == S.IDENTITY
    ident_sign = bytearray(S.jj_to_bytes(S.IDENTITY))
    ident_sign[31] |= 0x80                       # u = 0 with the sign bit set: rejected under ZIP 216
    assert S.jj_from_bytes(bytes(ident_sign)) is None
    assert S.jj_from_bytes((S.Q).to_bytes(32, "little")) is None      # non-canonical v
    assert sum(S.jj_from_bytes(i.to_bytes(32, "little")) is None for i in range(2, 40)) > 5


def test_empty_roots_match_the_reference():
    roots = [int.from_bytes(bytes.fromhex(h), "little") for h in GOLD["empty_roots"]]
    assert roots[0] == S.UNCOMMITTED
    cur = roots[0]
    for d in range(32):
        cur = S.merkle_hash(d, cur, cur)
        assert cur == roots[d + 1], "empty root %d" % (d + 1)
    assert S.empty_root(32) == roots[32]
    # a path of empty siblings from an empty leaf is the empty tree
    path = S.MerklePath.from_position(roots[:32], 0)
    assert path.root(S.UNCOMMITTED) == roots[32]


def test_commitment_tree_roots_match_the_reference():
    """test_sapling_tree (merkle_tree.rs:1091-1135, depth-4 test tree): the root after each of 16 appends, and
    MerklePath::root from every leaf of the full tree."""
    cm = [int.from_bytes(bytes.fromhex(h), "little") for h in GOLD["tree_commitments"]]
    want = [int.from_bytes(bytes.fromhex(h), "little") for h in GOLD["tree_roots"]]
    empty = [S.empty_root(d) for d in range(5)]

    def levels(leaves):
        out, level = [], list(leaves)
        for d in range(4):
            if len(level) % 2:
                level.append(empty[d])
            out.append(level)
            level = [S.merkle_hash(d, level[i], level[i + 1]) for i in range(0, len(level), 2)]
        return out, level[0]

    for n in range(1, 17):
        assert levels(cm[:n])[1] == want[n - 1], "root after %d appends" % n
    lv, root = levels(cm)
    for pos in range(16):
        path = S.MerklePath.from_position([lv[d][(pos >> d) ^ 1] for d in range(4)], pos)
        assert path.root(cm[pos]) == root == want[15]


def test_note_commitments_match_the_reference_vectors():
    asset = S.AssetType.from_identifier(bytes.fromhex(GOLD["asset_identifier"]))
    assert asset is not None
    for tv in GOLD["note_commitments"]:
        pk_d = S.jj_from_bytes(bytes.fromhex(tv["default_pk_d"]))
        to = S.PaymentAddress.from_parts(S.Diversifier(bytes.fromhex(tv["default_d"])), pk_d)
        rcm = int.from_bytes(bytes.fromhex(tv["rcm"]), "little")
        note = to.create_note(asset, tv["v"], S.Rseed.before_zip212(rcm))
        assert note is not None
        assert note.cmu().to_bytes(32, "little").hex() == tv["cmu"]


def test_zip32_key_vectors_match_the_reference():
    """zip32/sapling.rs:2074-2106: ak = ask * G_spend, nk = nsk * G_proof, ivk = CRH^ivk(ak, nk), for the external
    and the internal branch, and every diversifier the reference lists as found is valid under group_hash."""
    le = lambda h: int.from_bytes(bytes.fromhex(h), "little")
    for tv in GOLD["zip32_keys"]:
        ak, nk = S.jj_from_bytes(bytes.fromhex(tv["ak"])), S.jj_from_bytes(bytes.fromhex(tv["nk"]))
        assert ak is not None and nk is not None
        if tv["ask"]:
            assert S.jj_to_bytes(S.jj_mul(S.SPENDING_KEY_GENERATOR, le(tv["ask"]))).hex() == tv["ak"]
        if tv["nsk"]:
            vk = S.ProofGenerationKey(ak, le(tv["nsk"])).to_viewing_key()
            assert S.jj_to_bytes(vk.nk).hex() == tv["nk"]
        assert S.ViewingKey(ak, nk).ivk() == le(tv["ivk"])
        nk_int = S.jj_from_bytes(bytes.fromhex(tv["internal_nk"]))
        if tv["internal_nsk"]:
            assert S.jj_mul(S.PROOF_GENERATION_KEY_GENERATOR, le(tv["internal_nsk"])) == nk_int
        assert S.ViewingKey(ak, nk_int).ivk() == le(tv["internal_ivk"])
        for d in ("d0", "d1", "d2", "dmax"):
            if tv[d]:
                g_d = S.Diversifier(bytes.fromhex(tv[d])).g_d()
                assert g_d is not None and S.jj_mul(g_d, S.JUBJUB_ORDER) == S.IDENTITY
                addr = S.ViewingKey(ak, nk).to_payment_address(S.Diversifier(bytes.fromhex(tv[d])))
                assert addr is not None and addr.pk_d == S.jj_mul(g_d, le(tv["ivk"]))


def test_asset_type_rules():
    a = S.AssetType.new(b"BTC")
    assert a.nonce is not None and S.AssetType.new_with_nonce(b"BTC", a.nonce) == a
    assert all(S.AssetType.new_with_nonce(b"BTC", n) is None for n in range(a.nonce))
    assert S.AssetType.from_identifier(a.get_identifier()) == a
    g = a.asset_generator()
    assert S.jj_on_curve(g) and a.value_commitment_generator() == S.jj_mul(g, 8) != S.IDENTITY
    bad = next(i for i in range(256) if S.AssetType.from_identifier(bytes([i]) * 32) is None)
    assert S.AssetType.from_identifier(bytes([bad]) * 32) is None
    assert len(a.identifier_bits()) == 256
    # homomorphism of conversions (convert.rs:253-265)
    b, c = S.AssetType.new(b"ZEC"), S.AssetType.new(b"XAN")
    x = S.AllowedConversion({a: 5, b: 6, c: 7})
    y = S.AllowedConversion({a: 2, c: 10})
    z = S.AllowedConversion({a: 7, b: 6, c: 17})
    assert S.jj_add(x.generator, y.generator) == z.generator
    assert S.AllowedConversion({a: -3}).generator == S.jj_neg(S.AllowedConversion({a: 3}).generator)


def make_wallet(seed=5):
    r = random.Random(seed)
    ask, nsk = r.randrange(1, S.JUBJUB_ORDER), r.randrange(1, S.JUBJUB_ORDER)
    pgk = S.ProofGenerationKey(S.jj_mul(S.SPENDING_KEY_GENERATOR, ask), nsk)
    while True:
        d = S.Diversifier(bytes(r.getrandbits(8) for _ in range(11)))
        if d.g_d() is not None:
            return pgk, d


def spend_args(asset, value, seed=9, depth=32):
    r = random.Random(seed)
    pgk, d = make_wallet(seed)
    rcm, ar, rcv = (r.randrange(S.JUBJUB_ORDER) for _ in range(3))
    position = r.getrandbits(depth)
    path = S.MerklePath.from_position([r.randrange(S.Q) for _ in range(depth)], position)
    vk = pgk.to_viewing_key()
    note = vk.to_payment_address(d).create_note(asset, value, S.Rseed.before_zip212(rcm))
    anchor = path.root(note.cmu())
    return dict(proof_generation_key=pgk, diversifier=d, rseed=note.rseed, ar=ar, asset_type=asset, value=value,
                anchor=anchor, merkle_path=path, rcv=rcv), note


def test_native_public_inputs_equal_the_circuits_inputs():
    """What sapling/prover.rs computes natively for verify_proof equals what the product's witness
    generators assign to the input variables of the same instance (the reference's circuit tests
    check the same equalities: circuit/sapling.rs:743-759, 1045-1065, convert.rs:224-234)."""
    asset = S.AssetType.new(b"NAM")
    ints = lambda b: [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]
    # Spend
    ctx = S.SaplingProvingContext()
    args, note = spend_args(asset, 1234567)
    inst, pub, cv, rk = ctx._prepare_spend(**args)
    circ = C.Circuit(C.SPEND)
    assert circ.root(inst) == args["anchor"]          # the circuit's Merkle path agrees with MerklePath.root
    inputs, _ = circ.synthesize([inst])
    assert ints(inputs) == [1] + pub
    assert ctx.bsk == args["rcv"] and S.jj_on_curve(cv) and S.jj_on_curve(rk)
    # ZIP 212 rseed: rcm derived by PRF^expand
    z = S.Note(asset, 5, note.g_d, note.pk_d, S.Rseed.after_zip212(b"\x07" * 32))
    assert 0 <= z.rcm() < S.JUBJUB_ORDER and z.rcm() != S.Note(asset, 5, note.g_d, note.pk_d,
                                                                S.Rseed.after_zip212(b"\x08" * 32)).rcm()
    # Output
    esk, rcm, rcv = (RND.randrange(S.JUBJUB_ORDER) for _ in range(3))
    pgk, d = make_wallet(11)
    to = pgk.to_viewing_key().to_payment_address(d)
    inst, cv = ctx._prepare_output(esk, to, rcm, asset, 77, rcv)
    out_note = to.create_note(asset, 77, S.Rseed.before_zip212(rcm))
    inputs, _ = C.Circuit(C.OUTPUT).synthesize([inst])
    assert ints(inputs) == [1] + S.output_public_inputs(cv, S.jj_mul(to.g_d(), esk), out_note.cmu())
    assert ctx.bsk == (args["rcv"] - rcv) % S.JUBJUB_ORDER
    # Convert
    conv = S.AllowedConversion({asset: -1, S.AssetType.new(b"NAM/epoch1"): 1})
    path = S.MerklePath.from_position([RND.randrange(S.Q) for _ in range(32)], RND.getrandbits(32))
    anchor = path.root(conv.cmu())
    inst, pub, cv = ctx._prepare_convert(conv, 31337, anchor, path, 5)
    circ = C.Circuit(C.CONVERT)
    assert circ.root(inst) == anchor
    inputs, _ = circ.synthesize([inst])
    assert ints(inputs) == [1] + pub


def test_invalid_diversifier_is_err():
    asset = S.AssetType.new(b"NAM")
    args, _ = spend_args(asset, 1)
    bad = next(S.Diversifier(bytes([i]) * 11) for i in range(256) if S.Diversifier(bytes([i]) * 11).g_d() is None)
    args["diversifier"] = bad
    ctx = S.SaplingProvingContext()
    with pytest.raises(S.SaplingError):
        ctx._prepare_spend(**args)
    assert ctx.bsk == args["rcv"]      # the reference has already accumulated rcv when it returns Err(()) (:69-84)


def test_binding_sig_bookkeeping_and_redjubjub():
    """bsk / cv_sum as sapling/prover.rs:69-75, 154, 177-183, 205 keep them, and the consistency check and
    signature of binding_sig (:279-326), without any proof: BatchingTxProver only enqueues."""
    a, b = S.AssetType.new(b"NAM"), S.AssetType.new(b"ETH")
    tx = S.BatchingTxProver(local=None)
    ctx = tx.new_sapling_proving_context()
    args, _ = spend_args(a, 100)
    pend, cv_s, rk = tx.spend_proof(ctx, **args)
    pgk, d = make_wallet(3)
    to = pgk.to_viewing_key().to_payment_address(d)
    _, cv_o = tx.output_proof(ctx, 11, to, 22, a, 60, 33)
    _, cv_o2 = tx.output_proof(ctx, 12, to, 23, b, 5, 34)
    assert ctx.cv_sum == S.jj_add(cv_s, S.jj_neg(S.jj_add(cv_o, cv_o2)))
    assert ctx.bsk == (args["rcv"] - 33 - 34) % S.JUBJUB_ORDER
    assert len(ctx._pending) == 3 and pend._slot.proof is None
    sighash = bytes(range(32))
    sig = ctx.binding_sig({a: 40, b: -5}, sighash)        # value balance = spends - outputs, per asset
    bvk = S.jj_mul(S.VALUE_COMMITMENT_RANDOMNESS_GENERATOR, ctx.bsk)
    msg = S.jj_to_bytes(bvk) + sighash
    assert len(sig) == 64 and S.redjubjub_verify(bvk, msg, sig, S.VALUE_COMMITMENT_RANDOMNESS_GENERATOR)
    assert not S.redjubjub_verify(bvk, msg[:-1] + b"\x00", sig, S.VALUE_COMMITMENT_RANDOMNESS_GENERATOR)
    assert not S.redjubjub_verify(S.jj_add(bvk, bvk), msg, sig, S.VALUE_COMMITMENT_RANDOMNESS_GENERATOR)
    tampered = sig[:32] + ((int.from_bytes(sig[32:], "little") + 1) % S.JUBJUB_ORDER).to_bytes(32, "little")
    assert not S.redjubjub_verify(bvk, msg, tampered, S.VALUE_COMMITMENT_RANDOMNESS_GENERATOR)
    with pytest.raises(S.SaplingError):
        ctx.binding_sig({a: 41, b: -5}, sighash)          # wrong value balance: Err(())
    with pytest.raises(S.SaplingError):
        ctx.binding_sig({a: 40, b: -5, S.AssetType.new(b"X"): -(1 << 127)}, sighash)
    # deterministic rng -> deterministic signature
    fixed = lambda n: b"\x42" * n
    assert ctx.binding_sig({a: 40, b: -5}, sighash, rng=fixed) == ctx.binding_sig({a: 40, b: -5}, sighash, rng=fixed)


def test_multipacking_matches_bellman():
    nf = bytes(range(1, 33))
    bits = S.bytes_to_bits_le(nf)
    packed = S.compute_multipacking(bits)
    assert len(packed) == 2
    assert packed[0] + (packed[1] << 254) == int.from_bytes(nf, "little")
    assert packed[1] < 4


def test_v5_sapling_bundle_layout_round_trip():
    """Transaction::write_v5_sapling / read_v5_sapling (transaction.rs:612-720, 746-806): where the 192-byte
    proofs, cv, rk and the binding signature this path returns end up on the wire."""
    a, b = S.AssetType.new(b"NAM"), S.AssetType.new(b"ETH")
    pt = lambda k: S.jj_mul(S.SPENDING_KEY_GENERATOR, k)
    bundle = S.SaplingBundle(
        [S.SpendDescription(pt(3), 77, b"\x01" * 32, pt(4), b"\x02" * 192, b"\x03" * 64),
         S.SpendDescription(pt(5), 77, b"\x04" * 32, pt(6), b"\x05" * 192, b"\x06" * 64)],
        [S.ConvertDescription(pt(7), 99, b"\x07" * 192)],
        [S.OutputDescription(pt(8), 1234, b"\x08" * 32, b"\x09" * 612, b"\x0a" * 80, b"\x0b" * 192)],
        {a: -5, b: 1 << 100}, b"\x0c" * 64)
    w = S.write_v5_sapling(bundle)
    # 3 counts, 2 x (cv nf rk), 1 x cv, 1 x (cv cmu epk enc out), value balance (count + 2 x 48),
    # two anchors, 4 proofs, 2 spend auth sigs, binding sig
    assert len(w) == 3 + 2 * 96 + 32 + (96 + 612 + 80) + (1 + 2 * 48) + 64 + 4 * 192 + 2 * 64 + 64
    got, off = S.read_v5_sapling(w + b"tail")
    assert off == len(w) and got == bundle
    assert w[0] == 2 and w[1:33] == S.jj_to_bytes(pt(3)) and w[33:65] == b"\x01" * 32
    assert w.endswith(b"\x0c" * 64) and w[-64 - 192:-64] == b"\x0b" * 192
    # value balance is written in identifier order (a BTreeMap in the reference)
    vb = w[3 + 192 + 32 + 788:][:97]
    ids = [vb[1:33], vb[49:81]]
    assert vb[0] == 2 and ids == sorted(ids)
    assert S.write_v5_sapling(None) == b"\x00\x00\x00" and S.read_v5_sapling(b"\x00\x00\x00") == (None, 3)
    with pytest.raises(ValueError):
        S.read_v5_sapling(w[:-1])                        # truncated
    bad = bytearray(w)
    bad[1:33] = (S.Q).to_bytes(32, "little")             # cv with a non-canonical v coordinate
    with pytest.raises(ValueError):
        S.read_v5_sapling(bytes(bad))
    with pytest.raises(ValueError):                      # one anchor per bundle in v5
        S.write_v5_sapling(S.SaplingBundle([bundle.shielded_spends[0],
                                            S.SpendDescription(pt(5), 78, b"\x04" * 32, pt(6), b"\x05" * 192,
                                                               b"\x06" * 64)], binding_sig=b"\x00" * 64))
    assert [S._read_compact_size(S._compact_size(n), 0)[0] for n in (0, 252, 253, 65535, 65536, 1 << 33)] == \
        [0, 252, 253, 65535, 65536, 1 << 33]
    with pytest.raises(ValueError):
        S._read_compact_size(b"\xfd\x10\x00", 0)         # non-canonical CompactSize


# ---------------------------------------------------------------------------
# the whole surface on the device
# ---------------------------------------------------------------------------
def _real_key(name):
    """Key of a real MASP circuit under a known trapdoor (test_circuits.real_instance: the oracle's setup over
    the oracle's recording of the circuit).  Generating the three takes about a minute of CPU, so the bytes are
    cached under tests/_cache (git-ignored; it travels to the GPU box with the snapshot when present)."""
    cache = os.path.join(os.path.dirname(__file__), "_cache")
    path = os.path.join(cache, "key_%s.bin" % name)
    if os.path.exists(path):
        return open(path, "rb").read()
    from test_circuits import real_instance
    key = real_instance(name)[1]
    try:
        os.makedirs(cache, exist_ok=True)
        with open(path + ".tmp", "wb") as f:
            f.write(key)
        os.replace(path + ".tmp", path)
    except OSError:
        pass
    return key


@pytest.mark.gpu
def test_tx_prover_end_to_end_gpu(gpu):
    """One shielded transaction through the TxProver trait, serial and batched: every proof is accepted by
    verify_proof under NATIVELY computed public inputs (the device pairing check inside spend_proof /
    convert_proof, and again here including the Output proofs the reference does not self-check), a
    tampered public input is rejected, and the binding signature verifies against the accumulated bvk."""
    keys = {n: _real_key(n) for n in ("spend", "output", "convert")}   # trusted setups with known trapdoors
    local = gpu.LocalTxProver.from_bytes(keys["spend"], keys["output"], keys["convert"], verify_hashes=False)
    nam, epoch1 = S.AssetType.new(b"NAM"), S.AssetType.new(b"NAM/epoch1")
    conv = S.AllowedConversion({nam: -1, epoch1: 1})
    cpath = S.MerklePath.from_position([RND.randrange(S.Q) for _ in range(32)], 3)
    canchor = cpath.root(conv.cmu())
    pgk, d = make_wallet(21)
    to = pgk.to_viewing_key().to_payment_address(d)
    sighash = b"\x5a" * 32

    def run(tx):
        ctx = tx.new_sapling_proving_context()
        a1, n1 = spend_args(nam, 100, seed=31)
        a2, n2 = spend_args(nam, 50, seed=32)
        s1 = tx.spend_proof(ctx, **a1)
        s2 = tx.spend_proof(ctx, **a2)
        c1 = tx.convert_proof(ctx, conv, 150, canchor, cpath, 4242)
        o1 = tx.output_proof(ctx, 777, to, 888, epoch1, 120, 999)
        o2 = tx.output_proof(ctx, 778, to, 889, epoch1, 30, 1000)
        sig = tx.binding_sig(ctx, {}, sighash)      # 150 NAM in, converted 1:1, 150 epoch1 out: balance zero
        bvk = S.jj_mul(S.VALUE_COMMITMENT_RANDOMNESS_GENERATOR, ctx.bsk)
        assert S.redjubjub_verify(bvk, S.jj_to_bytes(bvk) + sighash, sig, S.VALUE_COMMITMENT_RANDOMNESS_GENERATOR)
        get = lambda p: p.proof if isinstance(p, S.PendingProof) else p
        proofs = [get(x[0]) for x in (s1, s2, c1, o1, o2)]
        assert all(len(p) == 192 for p in proofs) and len(set(proofs)) == 5
        # verifier-side public inputs (masp_proofs/src/sapling/verifier.rs:70-98, 120-150, 170-190)
        vk1, vk2 = a1["proof_generation_key"].to_viewing_key(), a2["proof_generation_key"].to_viewing_key()
        pub_s = [S.spend_public_inputs(s1[2], s1[1], a1["anchor"], n1.nf(vk1.nk, a1["merkle_path"].position)),
                 S.spend_public_inputs(s2[2], s2[1], a2["anchor"], n2.nf(vk2.nk, a2["merkle_path"].position))]
        assert gpu.verify_proofs(local.spend_params, proofs[:2], pub_s) == [True, True]
        assert gpu.verify_proofs(local.spend_params, proofs[:2], pub_s[::-1]) == [False, False]
        assert gpu.verify_proofs(local.convert_params, [proofs[2]], [S.convert_public_inputs(c1[1], canchor)]) == [True]
        pub_o = [S.output_public_inputs(o[1], S.jj_mul(to.g_d(), esk),
                                        to.create_note(epoch1, v, S.Rseed.before_zip212(rcm)).cmu())
                 for o, esk, rcm, v in ((o1, 777, 888, 120), (o2, 778, 889, 30))]
        assert gpu.verify_proofs(local.output_params, proofs[3:], pub_o) == [True, True]
        assert gpu.verify_proofs_batch(local.output_params, proofs[3:], pub_o)
        return ctx

    ctx_serial = run(S.TxProver(local))
    ctx_batched = run(S.BatchingTxProver(local))
    assert (ctx_serial.bsk, ctx_serial.cv_sum) == (ctx_batched.bsk, ctx_batched.cv_sum)

    # a Spend whose anchor is not the root of its path: the witness is computable but the anchor constraint
    # is not satisfied, the proof fails verify_proof and the reference returns Err(()) (sapling/prover.rs:148)
    tx = S.TxProver(local)
    bad, _ = spend_args(nam, 1, seed=33)
    bad["anchor"] = (bad["anchor"] + 1) % S.Q
    ctx = tx.new_sapling_proving_context()
    with pytest.raises(S.SaplingError):
        tx.spend_proof(ctx, **bad)
    assert ctx.cv_sum == S.IDENTITY and ctx.bsk == bad["rcv"]   # cv_sum is only touched after the check (:154)
    btx = S.BatchingTxProver(local)
    ctx = btx.new_sapling_proving_context()
    pend = btx.spend_proof(ctx, **bad)[0]
    with pytest.raises(S.SaplingError):
        pend.proof
