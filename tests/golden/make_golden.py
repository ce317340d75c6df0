"""Generates tests/golden/vectors.json with the Python-integer oracle (oracle/py).

Run from the repo root:  python tests/golden/make_golden.py [--full]
--full also proves one Output-shaped instance (BASELINE config 0) in Python
(about two minutes).  The Python oracle is the root of trust here: its proofs
pass the Groth16 pairing check and the trapdoor closed form (see
tests/test_oracle_py.py); everything faster is compared with these bytes.
"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.py.bls12_381 import G1, G2, R  # noqa: E402
from oracle.py import groth16 as g  # noqa: E402
from oracle import c_oracle as co  # noqa: E402  (only to materialise synthetic keys quickly)
from masp_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json")
ib = syn.ints_to_bytes
tolist = lambda b: [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]
unbits = lambda bm, n: [(bm[i >> 3] >> (i & 7)) & 1 for i in range(n)]


def py_prove_synth(shape, key_bytes, w):
    pp, used = g.Parameters.read(key_bytes)
    assert used == len(key_bytes)
    da, dbi, dba = shape.densities()
    dens = (unbits(da, shape.n_aux), unbits(dbi, shape.n_inputs), unbits(dba, shape.n_aux))
    proof = g.create_proof(pp, tolist(w["a"]), tolist(w["b"]), tolist(w["c"]), tolist(w["inputs"]), tolist(w["aux"]),
                           dens, tolist(w["r"])[0], tolist(w["s"])[0])
    return g.proof_write(proof)


def main(full):
    v = json.load(open(OUT)) if os.path.exists(OUT) else {}
    rnd = random.Random(20261017)
    v["g1_generator_compressed"] = G1.encode_compressed(G1.gen).hex()
    v["g2_generator_compressed"] = G2.encode_compressed(G2.gen).hex()

    # --- a real R1CS with a known trapdoor -------------------------------
    r1cs, wit = g.synthetic_r1cs(20, 10)
    inputs, aux = wit(0x2F5)
    td = g.Trapdoor(0x1111111111222233334444, 0x5555AAAA, 0x7777BBBBCC, 0x99990000111, 0x1234567890ABCDEF)
    params = g.generate_parameters(r1cs, td)
    a, b, c = r1cs.evaluate(inputs, aux)
    dens = r1cs.densities()
    r_, s_ = 0xDEADBEEFCAFEBABE1234, 0xFEEDFACE5678
    proof = g.create_proof(params, a, b, c, inputs, aux, dens, r_, s_)
    assert proof == g.expected_proof_from_trapdoor(r1cs, td, inputs, aux, r_, s_)
    assert g.verify_proof(params.vk, proof, inputs[1:])
    v["r1cs"] = {
        "params": params.write().hex(), "n_aux": r1cs.n_aux,
        "a_aux_density": syn.pack_bits(dens[0]).hex(), "b_input_density": syn.pack_bits(dens[1]).hex(),
        "b_aux_density": syn.pack_bits(dens[2]).hex(),
        "a": ib(a).hex(), "b": ib(b).hex(), "c": ib(c).hex(), "inputs": ib(inputs).hex(), "aux": ib(aux).hex(),
        "r": ib([r_]).hex(), "s": ib([s_]).hex(), "proof": g.proof_write(proof).hex(),
        "h": ib(g.h_coefficients(a, b, c)).hex()}

    # --- NTT ----------------------------------------------------------------
    ntt = {}
    for log_n in (1, 4, 6):
        n = 1 << log_n
        x = [rnd.randrange(R) for _ in range(n)]
        d = g.Domain(n)
        outs = {}
        for name, fn in (("fft", d.fft), ("ifft", d.ifft), ("coset_fft", d.coset_fft), ("icoset_fft", d.icoset_fft)):
            y = list(x)
            fn(y)
            outs[name] = ib(y).hex()
        ntt[str(log_n)] = {"in": ib(x).hex(), **outs}
    v["ntt"] = ntt

    # --- MSM ----------------------------------------------------------------
    n = 48
    ks = [rnd.randrange(R) for _ in range(n)]
    bases1 = G1.gen_mul_many(ks)
    bases2 = G2.gen_mul_many(ks[:16])
    sc = [rnd.choice([0, 1, rnd.randrange(R), rnd.randrange(R), R - 1, rnd.randrange(1 << 20)]) for _ in range(n)]
    r1 = G1.to_affine(g.multiexp(G1, bases1, sc))
    assert r1 == G1.to_affine(g.multiexp_naive(G1, bases1, sc))
    r2 = G2.to_affine(g.multiexp(G2, bases2, sc[:16]))
    v["msm"] = {"logs": ib(ks).hex(), "scalars": ib(sc).hex(),
                "bases_g1": b"".join(G1.encode_uncompressed(p) for p in bases1).hex(),
                "bases_g2": b"".join(G2.encode_uncompressed(p) for p in bases2).hex(),
                "result_g1": G1.encode_uncompressed(r1).hex(), "result_g2": G2.encode_uncompressed(r2).hex()}

    # --- tiny synthetic shape ------------------------------------------------
    sh = syn.tiny_shape()
    kb = co.params_from_logs(syn.key_logs(sh))
    # the key itself is checked against Python on a sample of points
    pp, _ = g.Parameters.read(kb)
    logs = syn.key_logs(sh)
    for q, pts in (("h", pp.h), ("l", pp.l), ("a", pp.a)):
        k0 = syn.limbs_to_ints(logs[q][:2])
        assert pts[:2] == G1.gen_mul_many(k0)
    assert pp.b_g2[:2] == G2.gen_mul_many(syn.limbs_to_ints(logs["b"][:2]))
    tiny = []
    for i in range(3):
        w = syn.witness(sh, i, co.fr_mul)
        tiny.append(py_prove_synth(sh, kb, w).hex())
    v["tiny_proofs"] = tiny
    import hashlib
    v["tiny_key_sha256"] = hashlib.sha256(kb).hexdigest()

    if full:
        t0 = time.time()
        sh = syn.OUTPUT
        kb = co.params_from_logs(syn.key_logs(sh))
        w = syn.witness(sh, 0, co.fr_mul)
        v["output_shape_proof"] = py_prove_synth(sh, kb, w).hex()
        v["output_shape_key_sha256"] = hashlib.sha256(kb).hexdigest()
        print("output-shape proof in %.0f s" % (time.time() - t0))
    json.dump(v, open(OUT, "w"), indent=0, sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main("--full" in sys.argv)
