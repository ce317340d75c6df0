"""Parity of the CUDA path with the oracle, through the C ABI (-m gpu).

Bit-exact bar: every byte string must equal the oracle's (integer / byte work).
Nothing here reads /root/reference.
"""
import random

import pytest

from masp_b200 import synthetic as syn
from util import R, ib, rand_scalars, assignment, oracle_proofs

pytestmark = pytest.mark.gpu


def test_selftest_device_arithmetic(gpu):
    assert gpu.selftest() == 0


def test_fr_mul(gpu, oracle):
    n = 5000
    a, b = rand_scalars(n, 1), rand_scalars(n, 2)
    assert gpu.fr_mul(ib(a), ib(b), n) == oracle.fr_mul(ib(a), ib(b), n)


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 7, 10, 13, 15, 16, 17])
def test_ntt_all_modes(gpu, oracle, log_n):
    v = ib(rand_scalars(1 << log_n, log_n, "uniform"))
    for inverse in (False, True):
        for coset in (False, True):
            assert gpu.ntt(v, log_n, inverse, coset) == oracle.ntt(v, log_n, inverse, coset), (inverse, coset)


def test_ntt_roundtrip_large(gpu):
    log_n = 20
    v = syn.limbs_to_bytes(syn.fr_uniform(7, 1, 1 << log_n))
    assert gpu.ntt(gpu.ntt(v, log_n, False, True), log_n, True, True) == v


@pytest.mark.parametrize("rows", [2, 9, 100, 1000, 31211, 47362, 100645])
def test_h_coefficients(gpu, oracle, rows):
    a = syn.limbs_to_bytes(syn.fr_uniform(11, rows, rows))
    b = syn.limbs_to_bytes(syn.fr_uniform(12, rows, rows))
    c = oracle.fr_mul(a, b, rows)
    assert gpu.h_coeffs(a, b, c, rows) == oracle.h_coeffs(a, b, c, rows)


def test_h_coefficients_unsatisfied_rows(gpu, oracle):
    # c != a*b: the division is inexact, bellman still truncates the top coefficient
    rows = 300
    a, b, c = (ib(rand_scalars(rows, s, "uniform")) for s in (1, 2, 3))
    assert gpu.h_coeffs(a, b, c, rows) == oracle.h_coeffs(a, b, c, rows)


def test_non_canonical_scalar_is_rejected(gpu):
    bad = (R).to_bytes(32, "little") * 4
    with pytest.raises(gpu.Mb200Error) as e:
        gpu.ntt(bad, 2)
    assert e.value.code == -6


def _g1_bases(oracle, n, stream=12):
    logs = syn.fr_uniform(syn.MASTER_SEED, stream, n)
    return logs, oracle.g1_gen_mul(syn.limbs_to_bytes(logs), n)


def test_synth_points_match_oracle(gpu, oracle):
    logs, bases = _g1_bases(oracle, 300)
    assert gpu.synth_points(12, 0, 300, 1) == bases
    l2 = syn.fr_uniform(syn.MASTER_SEED, 15, 40)
    assert gpu.synth_points(15, 0, 40, 2) == oracle.g2_gen_mul(syn.limbs_to_bytes(l2), 40)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 1000, 1 << 12, 1 << 14])
def test_msm_g1(gpu, oracle, n):
    logs, bases = _g1_bases(oracle, max(n, 1))
    sc = ib(rand_scalars(n, n))
    assert gpu.msm_g1(bases[:96 * n], sc, n) == oracle.msm_g1(bases[:96 * n], sc, n)


def test_msm_g1_special_bases(gpu, oracle):
    # repeated bases (P + P inside a bucket), negated pairs (P + (-P)), identity
    logs, bases = _g1_bases(oracle, 8)
    pts = [bases[96 * i:96 * (i + 1)] for i in range(8)]
    neg = lambda p: p[:48] + ((syn_p() - int.from_bytes(p[48:], "big")) % syn_p()).to_bytes(48, "big")
    inf = bytes([0x40]) + bytes(95)
    seq = [pts[0], pts[0], pts[0], neg(pts[0]), pts[1], neg(pts[1]), inf, pts[2], pts[2], inf]
    for scalars in ([5] * len(seq), [R - 1] * len(seq), [1] * len(seq), list(range(len(seq))), [7, 7, 9, 7, 3, 3, 5, 1, 1, 0],
                    # either side of the fold at (r - 1) / 2 in the digit walk
                    [(R - 1) // 2, (R + 1) // 2, (R - 1) // 2 - 1, (R + 1) // 2 + 1, R - 1, R - 2, 2, 1 << 254, (1 << 254) - 1, R - (1 << 200)]):
        b = b"".join(seq)
        assert gpu.msm_g1(b, ib(scalars), len(seq)) == oracle.msm_g1(b, ib(scalars), len(seq)), scalars


def test_msm_heavy_buckets(gpu, oracle):
    # every scalar equal: one bucket per window receives all n entries and is cut into segments
    n = 3000
    logs, bases = _g1_bases(oracle, n)
    for val in (7, R - 1, 1, (1 << 200) + 5):
        sc = ib([val] * n)
        assert gpu.msm_g1(bases, sc, n) == oracle.msm_g1(bases, sc, n), val
    l2 = syn.fr_uniform(syn.MASTER_SEED, 15, 400)
    b2 = oracle.g2_gen_mul(syn.limbs_to_bytes(l2), 400)
    sc = ib([12345] * 400)
    assert gpu.msm_g2(b2, sc, 400) == oracle.msm_g2(b2, sc, 400)


def syn_p():
    return 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def test_msm_g1_large_closed_form(gpu, oracle):
    # 2^16 bases with known discrete logs: result == (sum s_i k_i) * G  (BASELINE config 3)
    n = 1 << 16
    logs = syn.fr_uniform(syn.MASTER_SEED, syn.STREAM_MSM_BASE, n)
    bases = gpu.synth_points(syn.STREAM_MSM_BASE, 0, n, 1)
    for kind in ("U", "W"):
        sc = syn.msm_scalars(n, kind)
        dot = oracle.fr_dot(syn.limbs_to_bytes(sc), syn.limbs_to_bytes(logs), n)
        want = oracle.g1_gen_mul(dot.to_bytes(32, "little"), 1)
        assert gpu.msm_g1(bases, syn.limbs_to_bytes(sc), n) == want, kind


@pytest.mark.parametrize("n", [0, 1, 33, 700])
def test_msm_g2(gpu, oracle, n):
    logs = syn.fr_uniform(syn.MASTER_SEED, 15, max(n, 1))
    bases = oracle.g2_gen_mul(syn.limbs_to_bytes(logs), max(n, 1))[:192 * n]
    sc = ib(rand_scalars(n, n + 5))
    assert gpu.msm_g2(bases, sc, n) == oracle.msm_g2(bases, sc, n)


def test_split_msm_partials(gpu, oracle):
    # bases range-split in 4 shards, projective partials added (the 1/2/4/8-GPU MSM path on one device)
    n = 4096
    logs, bases = _g1_bases(oracle, n)
    sc = ib(rand_scalars(n, 99))
    parts = []
    for k in range(4):
        lo, hi = k * n // 4, (k + 1) * n // 4
        gb = gpu.G1Bases(bases[96 * lo:96 * hi], hi - lo)
        parts.append(gb.msm_partial(sc[32 * lo:32 * hi]))
    assert gpu.g1_sum_partials(parts) == oracle.msm_g1(bases, sc, n)


def test_malformed_parameters_are_rejected(gpu):
    sh = syn.tiny_shape()
    kb = bytearray(gpu.params_synthesize(sh))
    with pytest.raises(gpu.Mb200Error):
        gpu.Parameters.read(bytes(kb[:1000]), sh.densities())
    kb2 = bytearray(kb)
    kb2[0] |= 0x80  # compression flag on an uncompressed point
    with pytest.raises(gpu.Mb200Error):
        gpu.Parameters.read(bytes(kb2), sh.densities())
    with pytest.raises(gpu.Mb200Error):  # densities that do not match the query lengths
        gpu.Parameters.read(bytes(kb), None)


def _prove_and_compare(gpu, oracle, shape, n_proofs, chunk=None, n_check=None):
    kb = gpu.params_synthesize(shape)
    dens = shape.densities()
    P = gpu.Parameters.read(kb, dens)
    assert (P.n_inputs, P.n_aux, P.h_len, P.a_len, P.b_len, P.m) == (
        shape.n_inputs, shape.n_aux, shape.h_len, shape.a_len, shape.b_len, shape.m)
    assert P.consumed == len(kb) == shape.params_file_bytes()
    ws = [syn.witness(shape, i, gpu.fr_mul) for i in range(n_proofs)]
    if chunk:
        gpu.set_option("chunk", chunk)
    proofs = gpu.create_proof_batch([assignment(gpu, w) for w in ws], P, [w["r"] for w in ws], [w["s"] for w in ws])
    idx = list(range(n_proofs)) if n_check is None else sorted(random.Random(1).sample(range(n_proofs), n_check))
    want = oracle_proofs(oracle, kb, shape, dens, [ws[i] for i in idx])
    for i, wbytes in zip(idx, want):
        assert proofs[i] == wbytes, "proof %d differs" % i
    return kb, P, ws, proofs


def test_prove_tiny_batch(gpu, oracle):
    _prove_and_compare(gpu, oracle, syn.tiny_shape(), 7, chunk=3)


def test_synth_key_matches_oracle_key(gpu, oracle):
    sh = syn.tiny_shape()
    assert gpu.params_synthesize(sh) == oracle.params_from_logs(syn.key_logs(sh))


def test_prove_scaled_shapes(gpu, oracle):
    for name, frac in (("spend", 0.02), ("output", 0.05), ("convert", 0.03)):
        base = syn.SHAPES[name]
        sh = base.scaled(name + "_small", int(base.n_constraints * frac), frac)
        _prove_and_compare(gpu, oracle, sh, 3)


def test_prove_output_shape(gpu, oracle):
    # BASELINE config 0 shape on the GPU: Output, m = 2^15
    _prove_and_compare(gpu, oracle, syn.OUTPUT, 4, chunk=16)


def test_prove_convert_shape(gpu, oracle):
    _prove_and_compare(gpu, oracle, syn.CONVERT, 3, chunk=16)


def test_prove_spend_shape_batch(gpu, oracle):
    # BASELINE config 1 shape: Spend, m = 2^17.  A batch larger than one chunk,
    # a sample compared with the oracle, and chunking must not change any byte.
    kb, P, ws, proofs = _prove_and_compare(gpu, oracle, syn.SPEND, 20, chunk=8, n_check=3)
    gpu.set_option("chunk", 5)
    again = gpu.create_proof_batch([assignment(gpu, w) for w in ws], P, [w["r"] for w in ws], [w["s"] for w in ws])
    assert again == proofs
    assert len(set(proofs)) == len(proofs)
    gpu.set_option("chunk", 64)


def test_prove_all_zero_and_all_one_witness(gpu, oracle):
    sh = syn.tiny_shape()
    kb = gpu.params_synthesize(sh)
    dens = sh.densities()
    P = gpu.Parameters.read(kb, dens)
    one = (1).to_bytes(32, "little")
    zero = bytes(32)
    for aux_val in (zero, one):
        w = {"a": zero * sh.rows, "b": zero * sh.rows, "c": zero * sh.rows, "inputs": one * sh.n_inputs,
             "aux": aux_val * sh.n_aux, "r": zero, "s": zero}
        got = gpu.create_proof(assignment(gpu, w), P, w["r"], w["s"])
        assert got == oracle_proofs(oracle, kb, sh, dens, [w])[0]


def test_streamed_batches_overlap_and_match(gpu, oracle):
    """mb200_prove_submit / mb200_prove_wait: three batches in flight, waited out of order."""
    import ctypes
    sh = syn.tiny_shape()
    kb = gpu.params_synthesize(sh)
    dens = sh.densities()
    P = gpu.Parameters.read(kb, dens)
    batches, outs, tickets = [], [], []
    for k in range(3):
        ws = [syn.witness(sh, 10 * k + i, gpu.fr_mul) for i in range(4)]
        cat = lambda key: b"".join(w[key] for w in ws)
        bufs = [cat(key) for key in ("a", "b", "c", "inputs", "aux", "r", "s")]
        out = ctypes.create_string_buffer(192 * len(ws))
        tickets.append(gpu.prove_submit(P, len(ws), sh.rows, *bufs, ctypes.addressof(out)))
        batches.append((ws, bufs))
        outs.append(out)
    for k in (1, 0, 2):
        gpu.prove_wait(tickets[k])
    for (ws, _), out in zip(batches, outs):
        want = oracle_proofs(oracle, kb, sh, dens, ws)
        assert [out.raw[192 * i:192 * (i + 1)] for i in range(len(ws))] == want
    with pytest.raises(gpu.Mb200Error):
        gpu.prove_wait(tickets[0])  # a ticket can be waited on once
