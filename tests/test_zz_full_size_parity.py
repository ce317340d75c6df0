"""Parity at BASELINE.json's full sizes (-m gpu), the sample sizes SURVEY.md §8(d) asks for.

  * configs[1]: ALL 256 Spend-shaped proofs of the batch equal the CPU oracle byte for byte
    (~100 s of oracle time on the GPU box's host cores; the per-proof verdicts go to
    gpurun_out/full_parity_spend256.log so the run can be committed under profiles/);
  * configs[2]: a 64-proof sample of a 128-proof Convert batch (one GPU's share of the 1024);
  * Output: a 64-proof sample of a 256-proof batch;
  * configs[3]: the closed form (sum s_i k_i) G at 2^18 and 2^20 bases, uniform and witness-like scalars;
  * a full-size parameter file (the real files' sizes: key + a 1 366 052-byte transcript tail) through
    mb200_params_load_file with the reference's size / BLAKE2b-512 semantics
    (masp_proofs/src/lib.rs:278-325, 343-388).
Nothing here reads /root/reference.
"""
import hashlib
import os
import time

import numpy as np
import pytest

from masp_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(gpu, shape, n):
    """n synthetic witnesses as contiguous numpy buffers (the bench's distribution and streams)."""
    rows, n_aux, n_in = shape.rows, shape.n_aux, shape.n_inputs
    buf = {k: np.empty(n * per * 4, dtype="<u8") for k, per in
           (("a", rows), ("b", rows), ("c", rows), ("aux", n_aux), ("inputs", n_in), ("r", 1), ("s", 1))}
    is_bool = (shape.aux_classes() & 4) != 0
    for i in range(n):
        base = syn.STREAM_WIT_A + 8 * i
        a = syn.fr_uniform(syn.MASTER_SEED, base + 0, rows)
        b = syn.fr_uniform(syn.MASTER_SEED, base + 1, rows)
        inputs = syn.fr_uniform(syn.MASTER_SEED, base + 2, n_in)
        inputs[0] = (1, 0, 0, 0)
        a[shape.n_constraints:] = inputs
        b[shape.n_constraints:] = 0
        aux = syn.fr_uniform(syn.MASTER_SEED, base + 3, n_aux)
        bits = syn.fr_bits(syn.MASTER_SEED, base + 4, n_aux)
        aux[is_bool] = bits[is_bool]
        rs = syn.fr_uniform(syn.MASTER_SEED, base + 5, 2)
        for k, per, val in (("a", rows, a), ("b", rows, b), ("aux", n_aux, aux), ("inputs", n_in, inputs),
                            ("r", 1, rs[0:1]), ("s", 1, rs[1:2])):
            buf[k][i * per * 4:(i + 1) * per * 4] = val.reshape(-1)
    slab = 32
    for lo in range(0, n, slab):      # c = a * b through the library (device kernel)
        hi = min(n, lo + slab)
        sl = slice(lo * rows * 4, hi * rows * 4)
        c = gpu.fr_mul(buf["a"][sl].tobytes(), buf["b"][sl].tobytes(), (hi - lo) * rows)
        buf["c"][sl] = np.frombuffer(c, dtype="<u8")
    return buf


def _prove_and_sample(gpu, oracle, shape, n, sample, log_name=None):
    key = gpu.params_synthesize(shape)
    dens = shape.densities()
    P = gpu.Parameters.read(key, dens)
    buf = _batch(gpu, shape, n)
    t0 = time.perf_counter()
    out = gpu.prove_batch_raw(P, n, shape.rows, buf["a"], buf["b"], buf["c"], buf["inputs"], buf["aux"], buf["r"], buf["s"])
    t_gpu = time.perf_counter() - t0
    ref = oracle.Params(key, shape.n_aux, *dens)
    idx = list(range(n)) if sample >= n else [int(i) for i in np.linspace(0, n - 1, num=sample).astype(int)]
    g = lambda k, per, i: buf[k][i * per * 4:(i + 1) * per * 4].tobytes()
    lines, bad, t0 = [], [], time.perf_counter()
    for i in idx:
        want = ref.prove(shape.rows, g("a", shape.rows, i), g("b", shape.rows, i), g("c", shape.rows, i),
                         g("inputs", shape.n_inputs, i), g("aux", shape.n_aux, i), g("r", 1, i), g("s", 1, i))
        same = want == out[192 * i:192 * (i + 1)]
        lines.append("%s proof %4d %s %s" % (shape.name, i, hashlib.sha256(want).hexdigest()[:16], "==" if same else "!="))
        if not same:
            bad.append(i)
    t_cpu = time.perf_counter() - t0
    if log_name:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", log_name), "w") as f:
            f.write("# %s shape, batch %d through mb200_prove_batch (%.2f s incl. copies), %d proofs re-proved by oracle/c "
                    "on %d host threads (%.1f s): %d byte-identical, %d different\n"
                    % (shape.name, n, t_gpu, len(idx), oracle.get_threads(), t_cpu, len(idx) - len(bad), len(bad)))
            f.write("\n".join(lines) + "\n")
    assert not bad, "proofs differ from the oracle: %s" % bad[:10]
    assert len(set(out[192 * i:192 * (i + 1)] for i in range(n))) == n


def test_spend_batch_256_every_proof_equals_the_oracle(gpu, oracle):
    _prove_and_sample(gpu, oracle, syn.SPEND, 256, 256, "full_parity_spend256.log")


def test_convert_batch_sample_64(gpu, oracle):
    _prove_and_sample(gpu, oracle, syn.CONVERT, 128, 64, "full_parity_convert128.log")


def test_output_batch_sample_64(gpu, oracle):
    _prove_and_sample(gpu, oracle, syn.OUTPUT, 256, 64, "full_parity_output256.log")


@pytest.mark.parametrize("log_n", [18, 20])
def test_msm_g1_closed_form_large(gpu, oracle, log_n):
    n = 1 << log_n
    logs = syn.limbs_to_bytes(syn.fr_uniform(syn.MASTER_SEED, syn.STREAM_MSM_BASE, n))
    bases = gpu.synth_points(syn.STREAM_MSM_BASE, 0, n, 1)
    for kind in ("U", "W"):
        sc = syn.limbs_to_bytes(syn.msm_scalars(n, kind))
        dot = oracle.fr_dot(sc, logs, n)
        assert gpu.msm_g1(bases, sc, n) == oracle.g1_gen_mul(dot.to_bytes(32, "little"), 1), (log_n, kind)


def test_full_size_parameter_file_through_load_file(gpu, tmp_path):
    """A file of exactly MASP_SPEND_BYTES: the Spend-shaped key (48 482 520 bytes) followed by a
    1 366 052-byte transcript tail, loaded by path with the reference's checks."""
    from masp_b200.prover import MASP_SPEND_BYTES
    sh = syn.SPEND
    key = gpu.params_synthesize(sh)
    tail = hashlib.shake_128(b"transcript").digest(MASP_SPEND_BYTES - len(key))
    assert len(tail) == 1366052
    blob = key + tail
    digest = hashlib.blake2b(blob, digest_size=64).hexdigest()
    path = tmp_path / "masp-spend.params"
    path.write_bytes(blob)
    P = gpu.Parameters.read_file(str(path), MASP_SPEND_BYTES, digest, sh.densities())
    assert P.consumed == len(key) and (P.n_inputs, P.n_aux, P.h_len, P.a_len, P.b_len) == (
        sh.n_inputs, sh.n_aux, sh.h_len, sh.a_len, sh.b_len)
    # one flipped bit in the TAIL is caught (the hash covers the transcript), as is a wrong size
    flipped = bytearray(blob)
    flipped[-1] ^= 1
    path.write_bytes(bytes(flipped))
    with pytest.raises(gpu.Mb200Error) as e:
        gpu.Parameters.read_file(str(path), MASP_SPEND_BYTES, digest, sh.densities())
    assert e.value.code == -9
    path.write_bytes(blob[:-1])
    with pytest.raises(gpu.Mb200Error) as e:
        gpu.Parameters.read_file(str(path), MASP_SPEND_BYTES, digest, sh.densities())
    assert e.value.code == -9
    # the reference's own digest cannot match a synthetic file
    spec_bytes, spec_hash = MASP_SPEND_BYTES, gpu.MASP_SPEND_HASH
    path.write_bytes(blob)
    with pytest.raises(gpu.Mb200Error):
        gpu.Parameters.read_file(str(path), spec_bytes, spec_hash, sh.densities())
    # and the proofs from the file-loaded key equal those from the in-memory key
    Q = gpu.Parameters.read(key, sh.densities())
    w = syn.witness(sh, 0, gpu.fr_mul)
    from util import assignment
    assert gpu.create_proof(assignment(gpu, w), P, w["r"], w["s"]) == gpu.create_proof(assignment(gpu, w), Q, w["r"], w["s"])
