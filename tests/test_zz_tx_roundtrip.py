"""A shielded transaction from prover to verifier: TxProver -> SaplingBundle -> v5 bytes ->
SaplingVerificationContext / BatchValidator (masp_proofs/src/sapling/verifier.rs, verifier/single.rs,
verifier/batch.rs), the callers on both sides of the proving path.

On CPU the Groth16 calls are replaced by stand-ins that check the PUBLIC INPUTS the verifier side derives
from the wire bytes against the ones the prover side computed natively (the consensus logic, the
signatures and the cv_sum / bvk bookkeeping are real).  Under -m gpu the same transaction is proved,
serialised, parsed and verified on the device, singly and by the randomised batch check.
(The file sorts last on purpose: it is the newest GPU test of the round.)
"""
import pytest

from masp_b200 import sapling as S
from test_sapling import make_wallet, spend_args, RND

SIGHASH = b"\xa7" * 32


class _FakeLocal:
    """Stands in for LocalTxProver on CPU: 'proofs' are digests of the public inputs they would attest."""
    spend_params, output_params, convert_params = "spend", "output", "convert"

    @staticmethod
    def tag(kind, public_input):
        import hashlib
        return hashlib.blake2b(repr((kind, public_input)).encode(), digest_size=64).digest() * 3


def build_transaction(tx, local, asks):
    """2 Spends + 1 Convert + 2 Outputs with zero value balance; returns (bundle, expected public inputs)."""
    nam, epoch1 = S.AssetType.new(b"NAM"), S.AssetType.new(b"NAM/epoch1")
    conv = S.AllowedConversion({nam: -1, epoch1: 1})
    cpath = S.MerklePath.from_position([RND.randrange(S.Q) for _ in range(32)], 3)
    canchor = cpath.root(conv.cmu())
    pgk, d = make_wallet(21)
    to = pgk.to_viewing_key().to_payment_address(d)
    ctx = tx.new_sapling_proving_context()
    spends = []
    for seed, value in ((41, 100), (42, 50)):
        a, note = spend_args(nam, value, seed=seed)
        # (each test note sits in a tree of its own; v5 carries one anchor per bundle, so the wire part
        # below serialises the transaction as two bundles)
        proof, cv, rk = tx.spend_proof(ctx, **a)
        vk = a["proof_generation_key"].to_viewing_key()
        nf = note.nf(vk.nk, a["merkle_path"].position)
        spends.append((a, proof, cv, rk, nf))
    c_proof, c_cv = tx.convert_proof(ctx, conv, 150, canchor, cpath, 4242)
    outs = []
    for esk, rcm, value, rcv in ((777, 888, 120, 999), (778, 889, 30, 1000)):
        proof, cv = tx.output_proof(ctx, esk, to, rcm, epoch1, value, rcv)
        note = to.create_note(epoch1, value, S.Rseed.before_zip212(rcm))
        outs.append((proof, cv, note.cmu(), S.jj_mul(to.g_d(), esk)))
    binding_sig = tx.binding_sig(ctx, {}, SIGHASH)
    get = lambda p: p.proof if isinstance(p, S.PendingProof) else p
    return dict(ctx=ctx, spends=[(a, get(p), cv, rk, nf) for a, p, cv, rk, nf in spends],
                convert=(get(c_proof), c_cv, canchor), outputs=[(get(p), cv, cmu, epk) for p, cv, cmu, epk in outs],
                binding_sig=binding_sig, asks=asks)


def check_with_contexts(t, spend_vk, convert_vk, output_vk, single_verifier=None, batch_verifier=None):
    """The verifier side, description by description and batched; returns the accumulated cv_sum."""
    kw = {} if single_verifier is None else {"proof_verifier": single_verifier}
    v = S.SaplingVerificationContext(**kw)
    for a, proof, cv, rk, nf in t["spends"]:
        ask = t["asks"][a["_seed"]]
        sig = S.spend_sig(ask, a["ar"], SIGHASH)
        assert v.check_spend(cv, a["anchor"], nf, rk, SIGHASH, sig, proof, spend_vk)
        # a signature under the un-randomised key does not verify against rk
        w = S.SaplingVerificationContext(**kw)
        assert not w.check_spend(cv, a["anchor"], nf, rk, SIGHASH, S.spend_sig(ask, 0, SIGHASH), proof, spend_vk)
    c_proof, c_cv, canchor = t["convert"]
    assert v.check_convert(c_cv, canchor, c_proof, convert_vk)
    for proof, cv, cmu, epk in t["outputs"]:
        assert v.check_output(cv, cmu, epk, proof, output_vk)
    assert v.cv_sum == t["ctx"].cv_sum
    assert v.final_check({}, SIGHASH, t["binding_sig"])
    assert not v.final_check({S.AssetType.new(b"NAM"): 1}, SIGHASH, t["binding_sig"])
    assert not v.final_check({}, b"\x00" * 32, t["binding_sig"])
    # small-order points are refused before anything else (verifier.rs:48-50, 112-114, 143-145)
    w = S.SaplingVerificationContext(**kw)
    assert not w.check_convert(S.IDENTITY, canchor, c_proof, convert_vk)
    assert not w.check_output(t["outputs"][0][1], t["outputs"][0][2], (0, S.Q - 1), t["outputs"][0][0], output_vk)
    assert w.cv_sum == S.IDENTITY

    # the wire: one v5 bundle per spend anchor (the two test notes sit in different trees)
    ok_all = True
    bv = S.BatchValidator(**({} if batch_verifier is None else {"batch_verifier": batch_verifier}))
    a0, p0, cv0, rk0, nf0 = t["spends"][0]
    sig0 = S.spend_sig(t["asks"][a0["_seed"]], a0["ar"], SIGHASH)
    # a bundle of its own needs its own binding signature: re-derive bsk / cv_sum for the sub-transaction
    sub = S.SaplingProvingContext()
    sub.bsk, sub.cv_sum = a0["rcv"], cv0
    sub_sig = sub.binding_sig({a0["asset_type"]: a0["value"]}, SIGHASH)
    b1 = S.SaplingBundle([S.SpendDescription(cv0, a0["anchor"], nf0, rk0, p0, sig0)], [], [],
                         {a0["asset_type"]: a0["value"]}, sub_sig)
    wire = S.write_v5_sapling(b1)
    parsed, off = S.read_v5_sapling(wire)
    assert off == len(wire) and parsed == b1
    ok_all &= bv.check_bundle(parsed, SIGHASH)
    # the rest of the transaction as a second bundle: spend 2, the convert, both outputs
    a1, p1, cv1, rk1, nf1 = t["spends"][1]
    sig1 = S.spend_sig(t["asks"][a1["_seed"]], a1["ar"], SIGHASH)
    rest = S.SaplingProvingContext()
    rest.bsk = (t["ctx"].bsk - a0["rcv"]) % S.JUBJUB_ORDER
    rest.cv_sum = S.jj_add(t["ctx"].cv_sum, S.jj_neg(cv0))
    rest_sig = rest.binding_sig({a0["asset_type"]: -a0["value"]}, SIGHASH)
    b2 = S.SaplingBundle(
        [S.SpendDescription(cv1, a1["anchor"], nf1, rk1, p1, sig1)],
        [S.ConvertDescription(t["convert"][1], t["convert"][2], t["convert"][0])],
        [S.OutputDescription(cv, cmu, S.jj_to_bytes(epk), b"\x11" * S.ENC_CIPHERTEXT_SIZE, b"\x22" * S.OUT_CIPHERTEXT_SIZE,
                             proof) for proof, cv, cmu, epk in t["outputs"]],
        {a0["asset_type"]: -a0["value"]}, rest_sig)
    parsed2, _ = S.read_v5_sapling(S.write_v5_sapling(b2))
    assert parsed2 == b2
    ok_all &= bv.check_bundle(parsed2, SIGHASH)
    assert ok_all and len(bv.signatures) == 4 and (len(bv.spend_proofs), len(bv.convert_proofs),
                                                    len(bv.output_proofs)) == (2, 1, 2)
    assert bv.validate(spend_vk, convert_vk, output_vk)
    assert S.BatchValidator().validate(spend_vk, convert_vk, output_vk)          # nothing queued: true
    # one wrong nullifier anywhere fails the whole batch
    bad = S.BatchValidator(**({} if batch_verifier is None else {"batch_verifier": batch_verifier}))
    tampered = S.SaplingBundle([S.SpendDescription(cv0, a0["anchor"], bytes(32), rk0, p0, sig0)], [], [],
                               b1.value_balance, sub_sig)
    assert bad.check_bundle(tampered, SIGHASH) and bad.check_bundle(parsed2, SIGHASH)
    assert not bad.validate(spend_vk, convert_vk, output_vk)
    # a wrong binding signature is caught by the signature pass
    bad2 = S.BatchValidator(**({} if batch_verifier is None else {"batch_verifier": batch_verifier}))
    assert bad2.check_bundle(S.SaplingBundle(b1.shielded_spends, [], [], b1.value_balance, rest_sig), SIGHASH)
    assert not bad2.validate(spend_vk, convert_vk, output_vk)


def _patch_spend_args():
    """spend_args with the seed recorded, and the wallets' spend authorising keys by seed."""
    import random
    import test_sapling as T
    asks = {}
    orig = T.spend_args

    def wrapped(asset, value, seed=9, depth=32):
        a, note = orig(asset, value, seed=seed, depth=depth)
        asks[seed] = random.Random(seed).randrange(1, S.JUBJUB_ORDER)   # first draw of make_wallet(seed)
        a = dict(a)
        return _Seeded(a, seed), note
    return wrapped, asks


class _Seeded(dict):
    """spend_proof(**args) must not see the bookkeeping key."""

    def __init__(self, d, seed):
        super().__init__(d)
        self.seed = seed

    def __getitem__(self, k):
        return self.seed if k == "_seed" else super().__getitem__(k)


def test_transaction_round_trip_consensus_logic_cpu(monkeypatch):
    wrapped, asks = _patch_spend_args()
    monkeypatch.setattr("test_zz_tx_roundtrip.spend_args", wrapped)
    local = _FakeLocal()

    class FakeTx(S.BatchingTxProver):
        def flush(self, ctx, rng=None):
            for s in ctx._pending:
                s.proof = _FakeLocal.tag(s.kind, s.public_input)[:192]
            ctx._pending = []
    t = build_transaction(FakeTx(local), local, asks)
    seen = []

    def single(vk, proof, public_input):
        seen.append((vk, len(public_input)))
        return vk == "output" or proof == _FakeLocal.tag(vk, public_input)[:192]

    def batch(vk, proofs, inputs, rng):
        return all(single(vk, p, x) for p, x in zip(proofs, inputs))
    check_with_contexts(t, "spend", "convert", "output", single, batch)
    assert ("spend", 7) in seen and ("convert", 3) in seen and ("output", 5) in seen


@pytest.mark.gpu
def test_transaction_round_trip_gpu(gpu, monkeypatch):
    from test_sapling import _real_key
    wrapped, asks = _patch_spend_args()
    monkeypatch.setattr("test_zz_tx_roundtrip.spend_args", wrapped)
    local = gpu.LocalTxProver.from_bytes(_real_key("spend"), _real_key("output"), _real_key("convert"),
                                         verify_hashes=False)
    t = build_transaction(local.tx_prover(batching=True), local, asks)
    check_with_contexts(t, local.spend_params, local.convert_params, local.output_params)
