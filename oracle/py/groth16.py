"""Groth16 prover as bellperson executes it, on Python integers.

TEST INFRASTRUCTURE ONLY (oracle).  Nothing under masp_b200/ may import this.

PARITY UNPINNED (SURVEY.md finding 3 and §8c): the reference calls
bellman::groth16::create_random_proof at masp_proofs/src/sapling/prover.rs:116-117,
201-202, 251-252, and that function lives in nam-bellperson 0.26.6-nam.1
(reference Cargo.lock:1355-1358), with the multiexp / FFT engines in
nam-ec-gpu-gen 0.7.2-nam.0 (Cargo.lock:1416-1419); neither is vendored.  No
reference test pins proof bytes.  What follows restates the published
algorithm and is anchored by

  * the Groth16 verification equation, run through the pairing in
    bls12_381.py exactly as the reference does after proving
    (masp_proofs/src/sapling/prover.rs:148, :266);
  * the closed-form "trapdoor" check on keys generated from known
    tau, alpha, beta, gamma, delta;
  * the Parameters byte layout, which reconciles the three real parameter
    file sizes masp_proofs/src/lib.rs:74-76 to the byte (SURVEY finding 6).
"""
import math
import struct

from .bls12_381 import (R, G1, G2, FR_GENERATOR, FR_ROOT_OF_UNITY, FR_S,
                        fr_inv, fr_to_bytes, fr_from_bytes, multi_pairing_is_one)


# ----------------------------------------------------------------------------
# EvaluationDomain (bellperson domain.rs)
# ----------------------------------------------------------------------------
class Domain:
    def __init__(self, n_rows):
        m, exp = 1, 0
        while m < n_rows:
            m *= 2
            exp += 1
        if exp >= FR_S:
            raise ValueError("polynomial degree too large")
        self.m = m
        self.exp = exp
        self.omega = pow(FR_ROOT_OF_UNITY, 1 << (FR_S - exp), R)
        self.omegainv = fr_inv(self.omega)
        self.geninv = fr_inv(FR_GENERATOR)
        self.minv = fr_inv(m)

    def _fft(self, a, omega):
        """In-place radix-2 Cooley-Tukey: bit-reverse, then log m passes."""
        n, log_n = self.m, self.exp
        for k in range(n):
            rk = int(bin(k)[2:].zfill(log_n)[::-1], 2) if log_n else 0
            if k < rk:
                a[k], a[rk] = a[rk], a[k]
        mm = 1
        for _ in range(log_n):
            w_m = pow(omega, n // (2 * mm), R)
            for k in range(0, n, 2 * mm):
                w = 1
                for j in range(mm):
                    t = a[k + j + mm] * w % R
                    a[k + j + mm] = (a[k + j] - t) % R
                    a[k + j] = (a[k + j] + t) % R
                    w = w * w_m % R
            mm *= 2

    def fft(self, a):
        self._fft(a, self.omega)

    def ifft(self, a):
        self._fft(a, self.omegainv)
        for i in range(self.m):
            a[i] = a[i] * self.minv % R

    @staticmethod
    def distribute_powers(a, g):
        u = 1
        for i in range(len(a)):
            a[i] = a[i] * u % R
            u = u * g % R

    def coset_fft(self, a):
        self.distribute_powers(a, FR_GENERATOR)
        self.fft(a)

    def icoset_fft(self, a):
        self.ifft(a)
        self.distribute_powers(a, self.geninv)

    def z(self, tau):
        return (pow(tau, self.m, R) - 1) % R

    def divide_by_z_on_coset(self, a):
        i = fr_inv(self.z(FR_GENERATOR))
        for k in range(len(a)):
            a[k] = a[k] * i % R


def h_coefficients(a, b, c):
    """The H-polynomial scalars: 3 ifft, 3 coset_fft, a*b-c, /Z, icoset_fft,
    drop the last coefficient (Appendix A 'H')."""
    d = Domain(len(a))
    pad = lambda v: [x % R for x in v] + [0] * (d.m - len(v))
    a, b, c = pad(a), pad(b), pad(c)
    for v in (a, b, c):
        d.ifft(v)
        d.coset_fft(v)
    for i in range(d.m):
        a[i] = (a[i] * b[i] - c[i]) % R
    d.divide_by_z_on_coset(a)
    d.icoset_fft(a)
    return a[:d.m - 1]


# ----------------------------------------------------------------------------
# multiexp (ec-gpu-gen multiexp_cpu): window-parallel Pippenger, unsigned
# digits, zero skip, one -> direct add in the first window.
# ----------------------------------------------------------------------------
def multiexp(curve, bases, scalars):
    n = len(scalars)
    assert len(bases) >= n
    c = 3 if n < 32 else int(math.ceil(math.log(n)))
    parts = []
    skip = 0
    while skip < 255 + 1:  # Fr::NUM_BITS = 255; windows while skip < NUM_BITS
        if skip >= 255:
            break
        acc = curve.identity
        buckets = [curve.identity] * ((1 << c) - 1)
        mask = (1 << c) - 1
        for s, base in zip(scalars, bases):
            if s == 0:
                continue
            if s == 1:
                if skip == 0:
                    acc = curve.add_mixed(acc, base)
                continue
            d = (s >> skip) & mask
            if d:
                buckets[d - 1] = curve.add_mixed(buckets[d - 1], base)
        running = curve.identity
        for bkt in reversed(buckets):
            running = curve.add(running, bkt)
            acc = curve.add(acc, running)
        parts.append(acc)
        skip += c
    total = curve.identity
    for part in reversed(parts):
        for _ in range(c):
            total = curve.double(total)
        total = curve.add(total, part)
    return total


def multiexp_naive(curve, bases, scalars):
    acc = curve.identity
    for s, b in zip(scalars, bases):
        if s:
            acc = curve.add(acc, curve.mul(curve.from_affine(b), s % R))
    return acc


# ----------------------------------------------------------------------------
# Parameters
# ----------------------------------------------------------------------------
class VerifyingKey:
    def __init__(self, alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1, delta_g2, ic):
        self.alpha_g1 = alpha_g1; self.beta_g1 = beta_g1; self.beta_g2 = beta_g2
        self.gamma_g2 = gamma_g2; self.delta_g1 = delta_g1; self.delta_g2 = delta_g2
        self.ic = ic

    def write(self):
        out = [G1.encode_uncompressed(self.alpha_g1), G1.encode_uncompressed(self.beta_g1),
               G2.encode_uncompressed(self.beta_g2), G2.encode_uncompressed(self.gamma_g2),
               G1.encode_uncompressed(self.delta_g1), G2.encode_uncompressed(self.delta_g2),
               struct.pack(">I", len(self.ic))]
        out += [G1.encode_uncompressed(p) for p in self.ic]
        return b"".join(out)


class Parameters:
    def __init__(self, vk, h, l, a, b_g1, b_g2):
        self.vk = vk; self.h = h; self.l = l; self.a = a; self.b_g1 = b_g1; self.b_g2 = b_g2

    def write(self):
        out = [self.vk.write()]
        for q in (self.h, self.l, self.a, self.b_g1):
            out.append(struct.pack(">I", len(q)))
            out += [G1.encode_uncompressed(p) for p in q]
        out.append(struct.pack(">I", len(self.b_g2)))
        out += [G2.encode_uncompressed(p) for p in self.b_g2]
        return b"".join(out)

    @staticmethod
    def read(buf, checked=False):
        """Parameters::read(reader, checked).  The reference passes false
        (masp_proofs/src/lib.rs:336-341): no curve / subgroup checks.  Returns
        (params, bytes_consumed); trailing bytes (the MPC transcript) are the
        caller's to hash (lib.rs:343-388)."""
        pos = 0
        def g1():
            nonlocal pos
            pt = G1.decode_uncompressed(buf[pos:pos + 96], check=checked); pos += 96
            return pt
        def g2():
            nonlocal pos
            pt = G2.decode_uncompressed(buf[pos:pos + 192], check=checked); pos += 192
            return pt
        def u32():
            nonlocal pos
            v = struct.unpack(">I", buf[pos:pos + 4])[0]; pos += 4
            return v
        alpha_g1 = g1(); beta_g1 = g1(); beta_g2 = g2(); gamma_g2 = g2()
        delta_g1 = g1(); delta_g2 = g2()
        ic = [g1() for _ in range(u32())]
        vk = VerifyingKey(alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1, delta_g2, ic)
        h = [g1() for _ in range(u32())]
        l = [g1() for _ in range(u32())]
        a = [g1() for _ in range(u32())]
        b_g1 = [g1() for _ in range(u32())]
        b_g2 = [g2() for _ in range(u32())]
        return Parameters(vk, h, l, a, b_g1, b_g2), pos


# ----------------------------------------------------------------------------
# R1CS + ProvingAssignment semantics (Appendix A 'Synthesis')
# ----------------------------------------------------------------------------
class R1CS:
    """Variables: ('I', i) with ('I', 0) == ONE, and ('A', i).  A linear
    combination is a list of (var, coeff)."""
    def __init__(self, n_inputs, n_aux, constraints):
        self.n_inputs = n_inputs      # including ONE
        self.n_aux = n_aux
        self.constraints = constraints

    def rows(self):
        """Constraint rows followed by the prover's extra `input_i * 0 = 0` rows."""
        extra = [([(("I", i), 1)], [], []) for i in range(self.n_inputs)]
        return list(self.constraints) + extra

    def densities(self):
        a_aux = [False] * self.n_aux
        b_inp = [False] * self.n_inputs
        b_aux = [False] * self.n_aux
        for (A, B, C) in self.constraints:
            for (kind, i), coeff in A:
                if coeff % R and kind == "A":
                    a_aux[i] = True
            for (kind, i), coeff in B:
                if coeff % R:
                    if kind == "A":
                        b_aux[i] = True
                    else:
                        b_inp[i] = True
        return a_aux, b_inp, b_aux

    @staticmethod
    def eval_lc(lc, inputs, aux):
        acc = 0
        for (kind, i), coeff in lc:
            acc += coeff * (inputs[i] if kind == "I" else aux[i])
        return acc % R

    def evaluate(self, inputs, aux):
        a, b, c = [], [], []
        for (A, B, C) in self.rows():
            a.append(self.eval_lc(A, inputs, aux))
            b.append(self.eval_lc(B, inputs, aux))
            c.append(self.eval_lc(C, inputs, aux))
        return a, b, c

    def is_satisfied(self, inputs, aux):
        a, b, c = self.evaluate(inputs, aux)
        return all((x * y - z) % R == 0 for x, y, z in zip(a, b, c))


def synthetic_r1cs(n_squarings, n_bits, n_pub=2):
    """A small satisfiable system: a chain of squarings x_{k+1} = x_k^2 + ONE*k
    seeded by a public input, a few booleanity rows (bit*(ONE-bit)=0) whose
    bits are also packed into a public input, mirroring the reference
    circuits' mix of full-width and boolean aux variables.

    Returns (r1cs, witness_fn) with witness_fn(seed_int) -> (inputs, aux)."""
    ONE = ("I", 0)
    cons = []
    n_inputs = 1 + n_pub
    # aux layout: bits [0, n_bits), chain [n_bits, n_bits + n_squarings]
    x0 = n_bits
    for k in range(n_bits):
        cons.append(([(("A", k), 1)], [(ONE, 1), (("A", k), R - 1)], []))
    # pack bits == input 2
    pack = [(("A", k), pow(2, k, R)) for k in range(n_bits)]
    cons.append((pack, [(ONE, 1)], [(("I", 2), 1)]))
    # x0 * ONE = input 1
    cons.append(([(("A", x0), 1)], [(ONE, 1)], [(("I", 1), 1)]))
    for k in range(n_squarings):
        cons.append(([(("A", x0 + k), 1)], [(("A", x0 + k), 1)],
                     [(("A", x0 + k + 1), 1), (ONE, (R - (k + 1)) % R)]))
    n_aux = n_bits + n_squarings + 1
    r1cs = R1CS(n_inputs, n_aux, cons)

    def witness(seed):
        bits = [(seed >> k) & 1 for k in range(n_bits)]
        x = [pow(seed + 3, 5, R)]
        for k in range(n_squarings):
            x.append((x[-1] * x[-1] + (k + 1)) % R)
        inputs = [1, x[0], sum(b << k for k, b in enumerate(bits)) % R]
        return inputs, bits + x
    return r1cs, witness


# ----------------------------------------------------------------------------
# generator (Appendix A 'Generator') with an explicit trapdoor
# ----------------------------------------------------------------------------
class Trapdoor:
    def __init__(self, tau, alpha, beta, gamma, delta):
        self.tau = tau % R; self.alpha = alpha % R; self.beta = beta % R
        self.gamma = gamma % R; self.delta = delta % R


def _lagrange_at_tau(d, tau):
    """L_j(tau) for j in 0..m over {omega^j}."""
    m = d.m
    zt = d.z(tau)
    out = []
    wj = 1
    for _ in range(m):
        # L_j(tau) = z(tau)/m * omega^j / (tau - omega^j)
        out.append(zt * d.minv % R * wj % R * fr_inv((tau - wj) % R) % R)
        wj = wj * d.omega % R
    return out


def qap_at_tau(r1cs, td):
    rows = r1cs.rows()
    d = Domain(len(rows))
    lag = _lagrange_at_tau(d, td.tau)
    At = {"I": [0] * r1cs.n_inputs, "A": [0] * r1cs.n_aux}
    Bt = {"I": [0] * r1cs.n_inputs, "A": [0] * r1cs.n_aux}
    Ct = {"I": [0] * r1cs.n_inputs, "A": [0] * r1cs.n_aux}
    for j, (A, B, C) in enumerate(rows):
        for (kind, i), coeff in A:
            At[kind][i] = (At[kind][i] + coeff * lag[j]) % R
        for (kind, i), coeff in B:
            Bt[kind][i] = (Bt[kind][i] + coeff * lag[j]) % R
        for (kind, i), coeff in C:
            Ct[kind][i] = (Ct[kind][i] + coeff * lag[j]) % R
    return d, At, Bt, Ct


def generate_parameters(r1cs, td):
    """bellperson generate_parameters with the toxic waste supplied."""
    d, At, Bt, Ct = qap_at_tau(r1cs, td)
    dinv = fr_inv(td.delta); ginv = fr_inv(td.gamma)
    zt = d.z(td.tau)
    h_s = [pow(td.tau, i, R) * zt % R * dinv % R for i in range(d.m - 1)]
    ext = lambda kind, i: (td.beta * At[kind][i] + td.alpha * Bt[kind][i] + Ct[kind][i]) % R
    ic_s = [ext("I", i) * ginv % R for i in range(r1cs.n_inputs)]
    l_s = [ext("A", i) * dinv % R for i in range(r1cs.n_aux)]
    a_s = [x for x in At["I"] + At["A"] if x]
    b_s = [x for x in Bt["I"] + Bt["A"] if x]
    vk = VerifyingKey(G1.to_affine(G1.gen_mul(td.alpha)), G1.to_affine(G1.gen_mul(td.beta)),
                      G2.to_affine(G2.gen_mul(td.beta)), G2.to_affine(G2.gen_mul(td.gamma)),
                      G1.to_affine(G1.gen_mul(td.delta)), G2.to_affine(G2.gen_mul(td.delta)),
                      G1.gen_mul_many(ic_s))
    return Parameters(vk, G1.gen_mul_many(h_s), G1.gen_mul_many(l_s), G1.gen_mul_many(a_s),
                      G1.gen_mul_many(b_s), G2.gen_mul_many(b_s))


# ----------------------------------------------------------------------------
# "Structureless" synthetic keys: every query point is a PRNG scalar times the
# generator.  Same prover code path and cost as a real key of that shape; the
# discrete logs are known so the proof has a closed form in Fr.
# ----------------------------------------------------------------------------
class KeyLogs:
    """Discrete logs of a synthetic key (what mb200_params_synthesize draws)."""
    def __init__(self, alpha, beta, gamma, delta, ic, h, l, a, b):
        self.alpha = alpha; self.beta = beta; self.gamma = gamma; self.delta = delta
        self.ic = ic; self.h = h; self.l = l; self.a = a; self.b = b


def parameters_from_logs(kl):
    vk = VerifyingKey(G1.to_affine(G1.gen_mul(kl.alpha)), G1.to_affine(G1.gen_mul(kl.beta)),
                      G2.to_affine(G2.gen_mul(kl.beta)), G2.to_affine(G2.gen_mul(kl.gamma)),
                      G1.to_affine(G1.gen_mul(kl.delta)), G2.to_affine(G2.gen_mul(kl.delta)),
                      G1.gen_mul_many(kl.ic))
    return Parameters(vk, G1.gen_mul_many(kl.h), G1.gen_mul_many(kl.l), G1.gen_mul_many(kl.a),
                      G1.gen_mul_many(kl.b), G2.gen_mul_many(kl.b))


# ----------------------------------------------------------------------------
# create_proof (Appendix A 'MSMs' and 'Assembly')
# ----------------------------------------------------------------------------
def dense_select(values, density):
    return [v for v, d in zip(values, density) if d]


def msm_scalars(inputs, aux, a_aux_density, b_input_density, b_aux_density):
    """Scalar vectors in base order for the A, B queries."""
    a_sc = list(inputs) + dense_select(aux, a_aux_density)
    b_sc = dense_select(inputs, b_input_density) + dense_select(aux, b_aux_density)
    return a_sc, b_sc


def create_proof(params, a, b, c, inputs, aux, densities, r, s, msm=multiexp):
    """bellperson create_proof(circuit, params, r, s) after synthesis.

    a, b, c: per-row evaluations (len == n_constraints + n_inputs);
    densities: (a_aux, b_input, b_aux) boolean lists.  Returns the three
    affine proof points (A in G1, B in G2, C in G1)."""
    a_aux_d, b_in_d, b_aux_d = densities
    hs = h_coefficients(a, b, c)
    vk = params.vk
    n_in = len(inputs)
    h = msm(G1, params.h, hs)
    l = msm(G1, params.l, aux)
    a_in = msm(G1, params.a[:n_in], inputs)
    a_aux = msm(G1, params.a[n_in:], dense_select(aux, a_aux_d))
    nb_in = sum(b_in_d)
    b_in_sc = dense_select(inputs, b_in_d)
    b_aux_sc = dense_select(aux, b_aux_d)
    b1_in = msm(G1, params.b_g1[:nb_in], b_in_sc)
    b1_aux = msm(G1, params.b_g1[nb_in:], b_aux_sc)
    b2_in = msm(G2, params.b_g2[:nb_in], b_in_sc)
    b2_aux = msm(G2, params.b_g2[nb_in:], b_aux_sc)

    g_a = G1.mul(G1.from_affine(vk.delta_g1), r)
    g_a = G1.add_mixed(g_a, vk.alpha_g1)
    g_b = G2.mul(G2.from_affine(vk.delta_g2), s)
    g_b = G2.add_mixed(g_b, vk.beta_g2)
    g_c = G1.mul(G1.from_affine(vk.delta_g1), r * s % R)
    g_c = G1.add(g_c, G1.mul(G1.from_affine(vk.alpha_g1), s))
    g_c = G1.add(g_c, G1.mul(G1.from_affine(vk.beta_g1), r))
    a_answer = G1.add(a_in, a_aux)
    g_a = G1.add(g_a, a_answer)
    g_c = G1.add(g_c, G1.mul(a_answer, s))
    b1_answer = G1.add(b1_in, b1_aux)
    b2_answer = G2.add(b2_in, b2_aux)
    g_b = G2.add(g_b, b2_answer)
    g_c = G1.add(g_c, G1.mul(b1_answer, r))
    g_c = G1.add(g_c, h)
    g_c = G1.add(g_c, l)
    return G1.to_affine(g_a), G2.to_affine(g_b), G1.to_affine(g_c)


def proof_write(proof):
    """Proof::write: A (G1 compressed) || B (G2 compressed) || C (G1 compressed);
    consumer masp_proofs/src/prover.rs:190-193 ([u8; 192])."""
    A, B, C = proof
    return G1.encode_compressed(A) + G2.encode_compressed(B) + G1.encode_compressed(C)


def proof_read(buf):
    return (G1.decode_compressed(buf[:48]), G2.decode_compressed(buf[48:144]),
            G1.decode_compressed(buf[144:192]))


# ----------------------------------------------------------------------------
# checks
# ----------------------------------------------------------------------------
def verify_proof(vk, proof, public_inputs):
    """e(A,B) = e(alpha,beta) e(sum x_i IC_i, gamma) e(C, delta); public_inputs
    excludes ONE (as verify_proof's `public_inputs` slice does)."""
    A, B, C = proof
    acc = G1.from_affine(vk.ic[0])
    for x, icp in zip(public_inputs, vk.ic[1:]):
        acc = G1.add(acc, G1.mul(G1.from_affine(icp), x % R))
    acc = G1.to_affine(acc)
    return multi_pairing_is_one([
        (A, B), (G1.neg_affine(vk.alpha_g1), vk.beta_g2),
        (G1.neg_affine(acc), vk.gamma_g2), (G1.neg_affine(C), vk.delta_g2)])


def expected_proof_from_logs(kl, hs, inputs, aux, densities, r, s):
    """Closed form of the proof under a KeyLogs key; hs = H scalars."""
    a_aux_d, b_in_d, b_aux_d = densities
    a_sc, b_sc = msm_scalars(inputs, aux, a_aux_d, b_in_d, b_aux_d)
    dot = lambda xs, ys: sum(x * y for x, y in zip(xs, ys)) % R
    assert len(a_sc) == len(kl.a) and len(b_sc) == len(kl.b)
    la = (kl.alpha + dot(a_sc, kl.a) + r * kl.delta) % R
    lb = (kl.beta + dot(b_sc, kl.b) + s * kl.delta) % R
    lb1 = (kl.beta + dot(b_sc, kl.b)) % R
    lc = (s * la + r * lb1 + dot(hs, kl.h) + dot(aux, kl.l)) % R
    return (G1.to_affine(G1.gen_mul(la)), G2.to_affine(G2.gen_mul(lb)), G1.to_affine(G1.gen_mul(lc)))


def expected_proof_from_trapdoor(r1cs, td, inputs, aux, r, s):
    """Appendix A 'Trapdoor check'."""
    d, At, Bt, Ct = qap_at_tau(r1cs, td)
    z = {"I": inputs, "A": aux}
    sa = sum(z[k][i] * At[k][i] for k in ("I", "A") for i in range(len(z[k]))) % R
    sb = sum(z[k][i] * Bt[k][i] for k in ("I", "A") for i in range(len(z[k]))) % R
    sc_aux = sum(aux[i] * (td.beta * At["A"][i] + td.alpha * Bt["A"][i] + Ct["A"][i])
                 for i in range(len(aux))) % R
    a, b, c = r1cs.evaluate(inputs, aux)
    hs = h_coefficients(a, b, c)
    htau = sum(hv * pow(td.tau, i, R) for i, hv in enumerate(hs)) % R
    la = (td.alpha + sa + r * td.delta) % R
    lb = (td.beta + sb + s * td.delta) % R
    lc = ((sc_aux + htau * d.z(td.tau)) * fr_inv(td.delta) + s * la + r * lb - r * s * td.delta) % R
    return (G1.to_affine(G1.gen_mul(la)), G2.to_affine(G2.gen_mul(lb)), G1.to_affine(G1.gen_mul(lc)))
