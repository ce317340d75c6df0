"""Groth16 parameter generation for a recorded circuit with a supplied trapdoor
(bellman generate_parameters, SURVEY.md Appendix A 'Generator'), fast enough
for the real MASP circuits: the QAP is evaluated at tau on Python integers and
the ~10^5 scalar multiplications of the generators are done by the C++ oracle.

TEST INFRASTRUCTURE ONLY.
"""
import struct

from .py.bls12_381 import R
from .py.groth16 import Domain
from . import c_oracle as co


def _batch_inv(xs):
    pref, acc = [], 1
    for x in xs:
        pref.append(acc)
        acc = acc * x % R
    inv = pow(acc, R - 2, R)
    out = [0] * len(xs)
    for i in range(len(xs) - 1, -1, -1):
        out[i] = inv * pref[i] % R
        inv = inv * xs[i] % R
    return out


def generate_parameters_bytes(cs, td):
    """cs: oracle.py.r1cs_gadgets.ConstraintSystem after synthesis; td: groth16.Trapdoor.
    Returns (Parameters bytes, (a_aux_density, b_input_density, b_aux_density) bitmaps)."""
    from masp_b200.synthetic import pack_bits, ints_to_bytes
    n_in, n_aux = len(cs.inputs), len(cs.aux)
    rows = len(cs.constraints) + n_in
    d = Domain(rows)
    m = d.m
    # L_j(tau) = z(tau)/m * omega^j / (tau - omega^j)
    zt = d.z(td.tau)
    w, ws = 1, []
    for _ in range(m):
        ws.append(w)
        w = w * d.omega % R
    dinv = _batch_inv([(td.tau - x) % R for x in ws])
    lag = [zt * d.minv % R * x % R * y % R for x, y in zip(ws, dinv)]
    At = {"I": [0] * n_in, "A": [0] * n_aux}
    Bt = {"I": [0] * n_in, "A": [0] * n_aux}
    Ct = {"I": [0] * n_in, "A": [0] * n_aux}
    for j, (A, B, C) in enumerate(cs.constraints):
        lj = lag[j]
        for T, lc in ((At, A), (Bt, B), (Ct, C)):
            for (kind, i), coeff in lc:
                T[kind][i] = (T[kind][i] + coeff * lj) % R
    for i in range(n_in):  # the prover's extra rows: input_i * 0 = 0
        At["I"][i] = (At["I"][i] + lag[len(cs.constraints) + i]) % R
    deltainv, gammainv = pow(td.delta, R - 2, R), pow(td.gamma, R - 2, R)
    ext = lambda k, i: (td.beta * At[k][i] + td.alpha * Bt[k][i] + Ct[k][i]) % R
    h_s, t = [], zt * deltainv % R
    for _ in range(m - 1):
        h_s.append(t)
        t = t * td.tau % R
    ic_s = [ext("I", i) * gammainv % R for i in range(n_in)]
    l_s = [ext("A", i) * deltainv % R for i in range(n_aux)]
    a_all, b_all = At["I"] + At["A"], Bt["I"] + Bt["A"]
    a_s, b_s = [x for x in a_all if x], [x for x in b_all if x]
    g1 = lambda s: co.g1_gen_mul(ints_to_bytes(s), len(s))
    g2 = lambda s: co.g2_gen_mul(ints_to_bytes(s), len(s))
    out = [g1([td.alpha]), g1([td.beta]), g2([td.beta]), g2([td.gamma]), g1([td.delta]), g2([td.delta]),
           struct.pack(">I", n_in), g1(ic_s)]
    for q in (h_s, l_s, a_s, b_s):
        out += [struct.pack(">I", len(q)), g1(q)]
    out += [struct.pack(">I", len(b_s)), g2(b_s)]
    a_aux_d = [x != 0 for x in At["A"]]
    b_in_d = [x != 0 for x in Bt["I"]]
    b_aux_d = [x != 0 for x in Bt["A"]]
    return b"".join(out), (pack_bits(a_aux_d), pack_bits(b_in_d), pack_bits(b_aux_d))
