"""The device code compiled for the host (tests/emu/libmasp_b200_emu.so) against
the oracle: launch orchestration, bucket indexing, scan / scatter, reduction
levels, Stockham passes and chunking, all on CPU.  A test aid only: the product
library has no CPU path (tests/test_abi.py)."""
import pytest

from masp_b200 import synthetic as syn
from util import ib, rand_scalars, assignment, oracle_proofs


def test_selftest(emu):
    assert emu.selftest() == 0


def test_ntt_and_h(emu, oracle):
    for log_n in (1, 2, 3, 4, 5, 7, 10):
        v = ib(rand_scalars(1 << log_n, log_n, "uniform"))
        for inv in (False, True):
            for cos in (False, True):
                assert emu.ntt(v, log_n, inv, cos) == oracle.ntt(v, log_n, inv, cos), (log_n, inv, cos)
    for rows in (2, 9, 100, 300):
        a, b = ib(rand_scalars(rows, 1, "uniform")), ib(rand_scalars(rows, 2, "uniform"))
        c = oracle.fr_mul(a, b, rows)
        assert emu.fr_mul(a, b, rows) == c
        assert emu.h_coeffs(a, b, c, rows) == oracle.h_coeffs(a, b, c, rows)


def test_msm(emu, oracle):
    logs = syn.fr_uniform(syn.MASTER_SEED, 12, 300)
    bases = oracle.g1_gen_mul(syn.limbs_to_bytes(logs), 300)
    assert emu.synth_points(12, 0, 20, 1) == bases[:96 * 20]
    for n in (0, 1, 5, 40, 300):
        sc = ib(rand_scalars(n, n))
        assert emu.msm_g1(bases[:96 * n], sc, n) == oracle.msm_g1(bases[:96 * n], sc, n), n
    # heavy buckets: every scalar equal -> one bucket per window holds all 300 entries (3 segments)
    for val in (7, syn.R_INT - 1, 1, (1 << 200) + 5):
        sc = ib([val] * 300)
        assert emu.msm_g1(bases, sc, 300) == oracle.msm_g1(bases, sc, 300), val
    b2 = oracle.g2_gen_mul(syn.limbs_to_bytes(logs[:12]), 12)
    sc = ib(rand_scalars(12, 77))
    assert emu.msm_g2(b2, sc, 12) == oracle.msm_g2(b2, sc, 12)
    # a very heavy bucket: 9 000 equal scalars put 9 000 entries = 71 segments into one bucket per window, which the
    # tree combine folds in three rounds of fan-in 8 (71 -> 9 -> 2 -> 1); 520 is the two-round case (5 segments)
    big_logs = syn.fr_uniform(syn.MASTER_SEED, 12, 9000)
    big = oracle.g1_gen_mul(syn.limbs_to_bytes(big_logs), 9000)
    for n, val in ((9000, 5), (9000, syn.R_INT - 2), (520, 3)):
        sc = ib([val] * n)
        assert emu.msm_g1(big[:96 * n], sc, n) == oracle.msm_g1(big[:96 * n], sc, n), (n, val)
    parts = [emu.G1Bases(bases[96 * lo:96 * hi], hi - lo).msm_partial(ib(rand_scalars(300, 5))[32 * lo:32 * hi])
             for lo, hi in ((0, 100), (100, 300))]
    assert emu.g1_sum_partials(parts) == oracle.msm_g1(bases, ib(rand_scalars(300, 5)), 300)
    # slabs: a standalone MSM uploads / sorts / accumulates its scalars in slabs (two buffers in turn) and
    # adds the slab partials; 300 scalars in slabs of <= 64 is five slabs, 7 a single short one
    emu.set_option("msm_slab", 64)
    try:
        for n in (300, 65, 64, 7):
            sc = ib(rand_scalars(n, n + 1))
            assert emu.msm_g1(bases[:96 * n], sc, n) == oracle.msm_g1(bases[:96 * n], sc, n), n
        sc = ib(rand_scalars(12, 78))
        emu.set_option("msm_slab", 5)
        assert emu.msm_g2(b2, sc, 12) == oracle.msm_g2(b2, sc, 12)
    finally:
        emu.set_option("msm_slab", 1 << 22)


@pytest.mark.slow
def test_prove_batch_chunked(emu, oracle):
    sh = syn.micro_shape()
    kb = emu.params_synthesize(sh)
    assert kb == oracle.params_from_logs(syn.key_logs(sh))
    dens = sh.densities()
    P = emu.Parameters.read(kb, dens)
    assert (P.n_inputs, P.n_aux, P.h_len, P.a_len, P.b_len, P.m) == (sh.n_inputs, sh.n_aux, sh.h_len, sh.a_len,
                                                                     sh.b_len, sh.m)
    ws = [syn.witness(sh, i, oracle.fr_mul) for i in range(5)]
    emu.set_option("chunk", 2)
    got = emu.create_proof_batch([assignment(emu, w) for w in ws], P, [w["r"] for w in ws], [w["s"] for w in ws])
    assert got == oracle_proofs(oracle, kb, sh, dens, ws)
    bad = bytearray(ws[0]["aux"])
    bad[0:32] = syn.R_INT.to_bytes(32, "little")
    w = dict(ws[0], aux=bytes(bad))
    with pytest.raises(emu.Mb200Error) as e:
        emu.create_proof(assignment(emu, w), P, w["r"], w["s"])
    assert e.value.code == -6


@pytest.mark.slow
def test_local_tx_prover_surface(emu, oracle):
    """LocalTxProver::from_bytes / prove_bundle, and a parameter stream followed
    by transcript bytes as in the real .params files (masp_proofs/src/lib.rs:343-388)."""
    sh = syn.micro_shape()
    kb = oracle.params_from_logs(syn.key_logs(sh))
    tail = bytes(range(256)) * 5
    dens = {"spend": sh.densities(), "output": sh.densities(), "convert": sh.densities()}
    with pytest.raises(emu.ParameterError):  # BLAKE2b of a synthetic file cannot match the pinned hash
        emu.LocalTxProver.from_bytes(kb + tail, kb + tail, kb + tail, dens)
    prover = emu.LocalTxProver.from_bytes(kb + tail, kb, kb + tail, dens, verify_hashes=False)
    assert prover.spend_params.consumed == len(kb)
    ws = [syn.witness(sh, i, oracle.fr_mul) for i in range(3)]
    fixed = iter([int.from_bytes(w[k], "little") for w in ws for k in ("r", "s")])
    # rng is called r, s per proof in order within each circuit group
    spends, converts, outputs = [assignment(emu, ws[0])], [assignment(emu, ws[1])], [assignment(emu, ws[2])]
    rs = {0: (ws[0]["r"], ws[0]["s"]), 1: (ws[1]["r"], ws[1]["s"]), 2: (ws[2]["r"], ws[2]["s"])}
    seq = [rs[0][0], rs[0][1], rs[1][0], rs[1][1], rs[2][0], rs[2][1]]
    it = iter(seq)
    got = prover.prove_bundle(spends, converts, outputs, rng=lambda: next(it))
    want = oracle_proofs(oracle, kb, sh, sh.densities(), ws)
    assert [got[0][0], got[1][0], got[2][0]] == want
    # create_random_proof draws r, s itself: two calls must differ, both must be 192 bytes
    p1, p2 = prover.spend_proof(spends[0], self_check=False), prover.spend_proof(spends[0], self_check=False)
    assert len(p1) == len(p2) == 192 and p1 != p2
    # with the self-check on (the reference's behaviour) an unsatisfied witness is Err(()), not a proof
    with pytest.raises(emu.Mb200Error):
        prover.spend_proof(spends[0])
    # load_parameters semantics behind the C ABI (masp_proofs/src/lib.rs:278-325, 343-388): size first, then
    # Parameters::read, then BLAKE2b-512 over the whole stream, transcript included
    import hashlib
    import os
    import tempfile
    blob = kb + tail
    digest = hashlib.blake2b(blob, digest_size=64).hexdigest()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "masp-test.params")
        open(path, "wb").write(blob)
        P = emu.Parameters.read_file(path, len(blob), digest, sh.densities())
        assert P.consumed == len(kb) and (P.n_inputs, P.n_aux) == (sh.n_inputs, sh.n_aux)
        for size, dg, code in ((len(blob) + 1, digest, -9), (len(blob), digest[:-1] + ("0" if digest[-1] != "0" else "1"), -9),
                               (0, "", None)):
            if code is None:
                emu.Parameters.read_file(path, size, dg, sh.densities())
                continue
            with pytest.raises(emu.Mb200Error) as e:
                emu.Parameters.read_file(path, size, dg, sh.densities())
            assert e.value.code == code
        with pytest.raises(emu.Mb200Error) as e:
            emu.Parameters.read_file(os.path.join(d, "absent.params"), 0, "", sh.densities())
        assert e.value.code == -10
    assert emu.Parameters.read_verified(blob, len(blob), digest.upper(), sh.densities()).consumed == len(kb)
    with pytest.raises(emu.Mb200Error):
        emu.Parameters.read_verified(blob[:-1], len(blob) - 1, digest, sh.densities())
    # the library's BLAKE2b against hashlib on awkward lengths (block boundaries)
    import ctypes
    for n in (0, 1, 127, 128, 129, 255, 256, 257, 1000):
        out = ctypes.create_string_buffer(64)
        assert emu._lib.lib().mb200_blake2b512(blob[:n], n, ctypes.cast(out, ctypes.c_void_p)) == 0
        assert out.raw == hashlib.blake2b(blob[:n], digest_size=64).digest(), n


@pytest.mark.parametrize("c", [5, 13, 15, 16, 17])
def test_msm_window_sizes_and_the_scalar_fold(emu, oracle, c, monkeypatch):
    """Every window size the library may pick, including 15 and 17 (where folding s > (r - 1) / 2 to r - s saves
    a window: 254 bits in 17 / 15 windows): scalars on both sides of the fold, at the window boundaries, with the
    largest top digit, and the values whose folded image lands in the "ones" buckets."""
    monkeypatch.setenv("MB200_C_MSM", str(c))
    R = syn.R_INT
    H = (R - 1) // 2
    scalars = [H, H + 1, H - 1, H + 2, R - 1, R - 2, 2, 1, 0, (1 << 254) - 1, 1 << 254, R - (1 << 200), (1 << c) - 1,
               1 << (c - 1), (1 << (c - 1)) + 1, H - (1 << (c - 1)), H >> 3, R - ((1 << c) - 1), (1 << (2 * c)) - 1]
    scalars += [int(x) for x in rand_scalars(13, 1000 + c)]
    n = len(scalars)
    logs = syn.fr_uniform(syn.MASTER_SEED, 12, n)
    bases = oracle.g1_gen_mul(syn.limbs_to_bytes(logs), n)
    assert emu.msm_g1(bases, ib(scalars), n) == oracle.msm_g1(bases, ib(scalars), n)


def test_msm_special_bases(emu, oracle):
    """Repeated bases (P + P), negated pairs (P + (-P)), identities: the XYZZ accumulator
    must take the tangent, drop the cancelling pair and pass infinities through."""
    P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    logs = syn.fr_uniform(syn.MASTER_SEED, 12, 8)
    bases = oracle.g1_gen_mul(syn.limbs_to_bytes(logs), 8)
    pts = [bases[96 * i:96 * (i + 1)] for i in range(8)]
    neg = lambda p: p[:48] + ((P_MOD - int.from_bytes(p[48:], "big")) % P_MOD).to_bytes(48, "big")
    inf = bytes([0x40]) + bytes(95)
    seq = [pts[0], pts[0], pts[0], neg(pts[0]), pts[1], neg(pts[1]), inf, pts[2], pts[2], inf, pts[3], pts[3], pts[3], pts[3]]
    b = b"".join(seq)
    for scalars in ([5] * len(seq), [syn.R_INT - 1] * len(seq), [1] * len(seq), list(range(len(seq))),
                    [7, 7, 9, 7, 3, 3, 5, 1, 1, 0, 2, 2, 2, 2],
                    # either side of the fold at (r - 1) / 2 (the digit walk decomposes r - s above it) and the
                    # values whose folded image is 1 or 2 (the "ones" buckets with a negated entry)
                    [(syn.R_INT - 1) // 2, (syn.R_INT + 1) // 2, (syn.R_INT - 1) // 2 - 1, (syn.R_INT + 1) // 2 + 1,
                     syn.R_INT - 1, syn.R_INT - 2, 2, 1, syn.R_INT - 1, 0, 1 << 254, (1 << 254) - 1, (1 << 253) + 1,
                     syn.R_INT - (1 << 200)]):
        assert emu.msm_g1(b, ib(scalars), len(seq)) == oracle.msm_g1(b, ib(scalars), len(seq)), scalars
    b2 = oracle.g2_gen_mul(syn.limbs_to_bytes(logs[:3]), 3)
    q = [b2[192 * i:192 * (i + 1)] for i in range(3)]
    seq2 = [q[0], q[0], q[1], q[1], q[1], q[2], q[0]]
    for scalars in ([3] * 7, [1] * 7, [9, 9, 4, 4, 4, 1, 9]):
        assert emu.msm_g2(b"".join(seq2), ib(scalars), 7) == oracle.msm_g2(b"".join(seq2), ib(scalars), 7), scalars


@pytest.mark.slow
def test_ntt_shared_memory_form(emu, oracle):
    """The two-kernel shared-memory NTT (csrc/ntt_smem.cuh: every size from 2^12 to 2^18) and the
    pass-per-launch kernels (sizes outside that range) against the oracle: every mode (forward /
    inverse, coset or not) at sizes on both sides of the boundary, and through the six-transform H
    pipeline at 2^12 and the Output size 2^15 -- for satisfied rows AND for unsatisfied ones, where
    the division by Z is not exact (the six-transform identity must hold coefficient by coefficient
    for any rows) -- and it really is two launches where the pass form needs four to six."""
    for log_n in (11, 12, 16, 17):
        v = ib(rand_scalars(1 << log_n, log_n, "uniform"))
        emu.ntt(v, log_n)     # builds the domain tables of this size (their launches are not the transform's)
        for inv in (False, True):
            for cos in (False, True):
                l0 = emu.get_counter("launches")
                assert emu.ntt(v, log_n, inv, cos) == oracle.ntt(v, log_n, inv, cos), (log_n, inv, cos)
                launches = int(emu.get_counter("launches") - l0) - 1      # minus the scalar range check
                assert launches == (2 if log_n >= 12 else (log_n + 2) // 3), (log_n, launches)
    for rows in (2049, 31211):
        a, b = ib(rand_scalars(rows, 1, "uniform")), ib(rand_scalars(rows, 2, "uniform"))
        c = emu.fr_mul(a, b, rows)
        assert emu.h_coeffs(a, b, c, rows) == oracle.h_coeffs(a, b, c, rows), rows
        c = ib(rand_scalars(rows, 3))
        assert emu.h_coeffs(a, b, c, rows) == oracle.h_coeffs(a, b, c, rows), -rows


@pytest.mark.slow
def test_extreme_r_s(emu, oracle):
    """r and s at the ends of the range: the GLV split of the two scalar multiplications in C
    (k = k1 + k2 lambda, csrc/ec.cuh) must hold for 0, 1, lambda +- 1 and r - 1."""
    sh = syn.micro_shape()
    kb = oracle.params_from_logs(syn.key_logs(sh))
    dens = sh.densities()
    P = emu.Parameters.read(kb, dens)
    w = syn.witness(sh, 0, oracle.fr_mul)
    lam = 0xAC45A4010001A40200000000FFFFFFFF
    ref = oracle.Params(kb, sh.n_aux, *dens)
    for r_, s_ in ((0, 0), (1, syn.R_INT - 1), (syn.R_INT - 1, 1), (lam, lam - 1), (lam + 1, lam * lam % syn.R_INT),
                   (1 << 254, (1 << 128) - 1)):
        rb, sb = ib([r_]), ib([s_])
        got = emu.create_proof(assignment(emu, w), P, rb, sb)
        assert got == ref.prove(sh.rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], rb, sb), (hex(r_), hex(s_))
