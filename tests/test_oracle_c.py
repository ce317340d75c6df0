"""The C++ oracle (the timed CPU baseline) against the committed golden
vectors of the Python oracle, byte for byte."""
import hashlib
import json
import os

import pytest

from masp_b200 import synthetic as syn
from util import ib, rand_scalars

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
H = bytes.fromhex


def test_r1cs_proof_and_h(oracle):
    v = GOLD["r1cs"]
    P = oracle.Params(H(v["params"]), v["n_aux"], H(v["a_aux_density"]), H(v["b_input_density"]), H(v["b_aux_density"]))
    rows = len(H(v["a"])) // 32
    assert P.consumed == len(H(v["params"]))
    got = P.prove(rows, H(v["a"]), H(v["b"]), H(v["c"]), H(v["inputs"]), H(v["aux"]), H(v["r"]), H(v["s"]))
    assert got.hex() == v["proof"]
    assert oracle.h_coeffs(H(v["a"]), H(v["b"]), H(v["c"]), rows).hex() == v["h"]


def test_ntt_vectors(oracle):
    for log_n, v in GOLD["ntt"].items():
        for name, inv, cos in (("fft", 0, 0), ("ifft", 1, 0), ("coset_fft", 0, 1), ("icoset_fft", 1, 1)):
            assert oracle.ntt(H(v["in"]), int(log_n), inv, cos).hex() == v[name], (log_n, name)


def test_msm_vectors(oracle):
    v = GOLD["msm"]
    n = len(H(v["scalars"])) // 32
    assert oracle.g1_gen_mul(H(v["logs"]), n).hex() == v["bases_g1"]
    assert oracle.g2_gen_mul(H(v["logs"])[:32 * 16], 16).hex() == v["bases_g2"]
    assert oracle.msm_g1(H(v["bases_g1"]), H(v["scalars"]), n).hex() == v["result_g1"]
    assert oracle.msm_g2(H(v["bases_g2"]), H(v["scalars"])[:32 * 16], 16).hex() == v["result_g2"]


def test_tiny_shape_proofs(oracle):
    sh = syn.tiny_shape()
    kb = oracle.params_from_logs(syn.key_logs(sh))
    assert hashlib.sha256(kb).hexdigest() == GOLD["tiny_key_sha256"]
    assert len(kb) == sh.params_file_bytes()
    P = oracle.Params(kb, sh.n_aux, *sh.densities())
    for i, want in enumerate(GOLD["tiny_proofs"]):
        w = syn.witness(sh, i, oracle.fr_mul)
        assert P.prove(sh.rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["r"], w["s"]).hex() == want


def test_threads_do_not_change_bytes(oracle):
    sh = syn.tiny_shape()
    kb = oracle.params_from_logs(syn.key_logs(sh))
    P = oracle.Params(kb, sh.n_aux, *sh.densities())
    w = syn.witness(sh, 1, oracle.fr_mul)
    outs = []
    for t in (1, 3, 0):
        oracle.set_threads(t)
        outs.append(P.prove(sh.rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["r"], w["s"]))
    oracle.set_threads(0)
    assert outs[0] == outs[1] == outs[2] == H(GOLD["tiny_proofs"][1])


def test_rejects_malformed(oracle):
    sh = syn.tiny_shape()
    kb = oracle.params_from_logs(syn.key_logs(sh))
    with pytest.raises(ValueError):
        oracle.Params(kb[:5000], sh.n_aux, *sh.densities())
    with pytest.raises(ValueError):
        oracle.Params(kb, sh.n_aux, None, None, None)
    with pytest.raises(ValueError):
        oracle.ntt((syn.R_INT).to_bytes(32, "little") * 2, 1)


@pytest.mark.slow
def test_config0_output_shape_on_cpu(oracle):
    """BASELINE config 0: one Output-shaped proof on CPU, C++ path == Python oracle."""
    if "output_shape_proof" not in GOLD:
        pytest.skip("golden generated without --full")
    sh = syn.OUTPUT
    kb = oracle.params_from_logs(syn.key_logs(sh))
    assert hashlib.sha256(kb).hexdigest() == GOLD["output_shape_key_sha256"]
    assert len(kb) == 15032568  # key bytes in file for Output
    P = oracle.Params(kb, sh.n_aux, *sh.densities())
    w = syn.witness(sh, 0, oracle.fr_mul)
    got = P.prove(sh.rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["r"], w["s"])
    assert got.hex() == GOLD["output_shape_proof"]
