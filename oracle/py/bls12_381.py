"""BLS12-381 field, curve and encoding arithmetic on Python integers.

TEST INFRASTRUCTURE ONLY (oracle).  Nothing under masp_b200/ may import this.

PARITY UNPINNED: the reference's own tests hold no golden vector for this
layer (SURVEY.md finding 3).  The arithmetic lives in crates that are not
vendored under /root/reference: nam-blstrs 0.7.1-nam.0 over nam-blst
0.3.15-nam.0 (reference Cargo.lock:1385-1411), selected at
masp_proofs/Cargo.toml:22.  This file restates the published mathematics of
BLS12-381 (constants: SURVEY.md Appendix C; encodings: Appendix D) and is
anchored on the reference call sites masp_proofs/src/prover.rs:190-193
(Proof::write -> 192 bytes) and masp_proofs/src/lib.rs:336-341
(Parameters::read(.., false)).  Self-consistency pins live in
tests/test_oracle_py.py: curve membership, subgroup order, bilinearity of the
pairing, the public compressed encodings of the two generators.
"""

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BLS_X = 0xD201000000010000  # |x|; the curve parameter is -x

G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G2_X = (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E)
G2_Y = (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE)

# Fr: multiplicative generator and 2^32-th root of unity (ff::PrimeField for
# bls12_381::Scalar: MULTIPLICATIVE_GENERATOR = 7, S = 32).
FR_GENERATOR = 7
FR_S = 32
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R - 1) >> FR_S, R)


# ----------------------------------------------------------------------------
# Field "ops" objects: the curve code below is generic over these.
# ----------------------------------------------------------------------------
class FpOps:
    zero = 0
    one = 1
    nbytes = 48

    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def mul(a, b): return (a * b) % P
    @staticmethod
    def sqr(a): return (a * a) % P
    @staticmethod
    def neg(a): return (-a) % P
    @staticmethod
    def inv(a): return pow(a, P - 2, P)
    @staticmethod
    def is_zero(a): return a == 0
    @staticmethod
    def muli(a, k): return (a * k) % P

    @staticmethod
    def lex_larger(y):
        """True iff y is lexicographically larger than -y."""
        return y > (P - y) % P

    @staticmethod
    def to_bytes(a): return a.to_bytes(48, "big")
    @staticmethod
    def from_bytes(b):
        v = int.from_bytes(b, "big")
        if v >= P:
            raise ValueError("Fp element not canonical")
        return v

    @staticmethod
    def sqrt(a):
        # p = 3 mod 4
        s = pow(a, (P + 1) // 4, P)
        return s if (s * s) % P == a % P else None


class Fp2Ops:
    zero = (0, 0)
    one = (1, 0)
    nbytes = 96

    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
    @staticmethod
    def mul(a, b):
        a0, a1 = a; b0, b1 = b
        return ((a0 * b0 - a1 * b1) % P, (a0 * b1 + a1 * b0) % P)
    @staticmethod
    def sqr(a):
        a0, a1 = a
        return ((a0 + a1) * (a0 - a1) % P, (2 * a0 * a1) % P)
    @staticmethod
    def neg(a): return ((-a[0]) % P, (-a[1]) % P)
    @staticmethod
    def inv(a):
        a0, a1 = a
        t = pow((a0 * a0 + a1 * a1) % P, P - 2, P)
        return ((a0 * t) % P, (-a1 * t) % P)
    @staticmethod
    def is_zero(a): return a[0] == 0 and a[1] == 0
    @staticmethod
    def muli(a, k): return ((a[0] * k) % P, (a[1] * k) % P)

    @staticmethod
    def lex_larger(y):
        """Compare c1 first, then c0."""
        n = ((-y[0]) % P, (-y[1]) % P)
        return (y[1], y[0]) > (n[1], n[0])

    @staticmethod
    def to_bytes(a):  # c1 || c0, big-endian
        return a[1].to_bytes(48, "big") + a[0].to_bytes(48, "big")
    @staticmethod
    def from_bytes(b):
        c1 = int.from_bytes(b[:48], "big"); c0 = int.from_bytes(b[48:96], "big")
        if c0 >= P or c1 >= P:
            raise ValueError("Fp2 element not canonical")
        return (c0, c1)

    @staticmethod
    def sqrt(a):
        # Generic: a^((p^2+7)/16) candidates (p^2 = 9 mod 16); small helper
        # used only by tests (decompression is not on the proving path).
        def fpow(x, e):
            r = (1, 0)
            while e:
                if e & 1:
                    r = Fp2Ops.mul(r, x)
                x = Fp2Ops.sqr(x); e >>= 1
            return r
        if Fp2Ops.is_zero(a):
            return (0, 0)
        c = fpow(a, (P * P + 7) // 16)
        # multiply by 8th roots of unity until it squares to a
        # an element of order 8 in Fp2*: take a non-residue power
        g = fpow((1, 1), (P * P - 1) // 8)
        for _ in range(8):
            if Fp2Ops.sqr(c) == (a[0] % P, a[1] % P):
                return c
            c = Fp2Ops.mul(c, g)
        return None


# ----------------------------------------------------------------------------
# Short Weierstrass y^2 = x^3 + b, a = 0.  Jacobian (X, Y, Z); Z == zero is
# the identity.  Affine points are (x, y) or None for the identity.
# ----------------------------------------------------------------------------
class Curve:
    def __init__(self, F, b, gen, name):
        self.F = F
        self.b = b
        self.gen = gen
        self.name = name
        self.identity = (F.one, F.one, F.zero)

    def is_on_curve(self, pt):
        if pt is None:
            return True
        F = self.F
        x, y = pt
        return F.sqr(y) == F.add(F.mul(F.sqr(x), x), self.b)

    def from_affine(self, pt):
        if pt is None:
            return self.identity
        return (pt[0], pt[1], self.F.one)

    def to_affine(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def batch_to_affine(self, Js):
        """Montgomery's simultaneous inversion over the Z coordinates."""
        F = self.F
        acc = F.one
        pref = []
        for (_, _, Z) in Js:
            pref.append(acc)
            if not F.is_zero(Z):
                acc = F.mul(acc, Z)
        inv = F.inv(acc)
        out = [None] * len(Js)
        for i in range(len(Js) - 1, -1, -1):
            X, Y, Z = Js[i]
            if F.is_zero(Z):
                continue
            zi = F.mul(inv, pref[i])
            inv = F.mul(inv, Z)
            zi2 = F.sqr(zi)
            out[i] = (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))
        return out

    def double(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return J
        A = F.sqr(X); B = F.sqr(Y); C = F.sqr(B)
        D = F.muli(F.sub(F.sub(F.sqr(F.add(X, B)), A), C), 2)
        E = F.muli(A, 3)
        Fq = F.sqr(E)
        X3 = F.sub(Fq, F.muli(D, 2))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.muli(C, 8))
        Z3 = F.muli(F.mul(Y, Z), 2)
        return (X3, Y3, Z3)

    def add(self, J1, J2):
        F = self.F
        X1, Y1, Z1 = J1
        X2, Y2, Z2 = J2
        if F.is_zero(Z1):
            return J2
        if F.is_zero(Z2):
            return J1
        Z1Z1 = F.sqr(Z1); Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2); U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(Y1, F.mul(Z2, Z2Z2)); S2 = F.mul(Y2, F.mul(Z1, Z1Z1))
        if U1 == U2:
            if S1 == S2:
                return self.double(J1)
            return self.identity
        H = F.sub(U2, U1); Rr = F.sub(S2, S1)
        HH = F.sqr(H); HHH = F.mul(H, HH); V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.muli(V, 2))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def add_mixed(self, J1, pt):
        F = self.F
        if pt is None:
            return J1
        X1, Y1, Z1 = J1
        if F.is_zero(Z1):
            return (pt[0], pt[1], F.one)
        x2, y2 = pt
        Z1Z1 = F.sqr(Z1)
        U2 = F.mul(x2, Z1Z1); S2 = F.mul(y2, F.mul(Z1, Z1Z1))
        if X1 == U2:
            if Y1 == S2:
                return self.double(J1)
            return self.identity
        H = F.sub(U2, X1); Rr = F.sub(S2, Y1)
        HH = F.sqr(H); HHH = F.mul(H, HH); V = F.mul(X1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.muli(V, 2))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(Y1, HHH))
        Z3 = F.mul(Z1, H)
        return (X3, Y3, Z3)

    def neg_affine(self, pt):
        if pt is None:
            return None
        return (pt[0], self.F.neg(pt[1]))

    def mul(self, J, k):
        """Left-to-right double and add; k is reduced mod r by the caller."""
        acc = self.identity
        if k == 0:
            return acc
        for bit in bin(k)[2:]:
            acc = self.double(acc)
            if bit == "1":
                acc = self.add(acc, J)
        return acc

    def mul_affine(self, pt, k):
        return self.to_affine(self.mul(self.from_affine(pt), k))

    def eq(self, J1, J2):
        return self.to_affine(J1) == self.to_affine(J2)

    # -- fixed-base table for the generator: makes synthetic keys affordable
    def gen_table(self, w=8):
        key = ("_tab", w)
        tab = getattr(self, "_gen_tab", {}).get(key)
        if tab is not None:
            return tab
        nwin = (255 + w - 1) // w
        rows = []
        base = self.from_affine(self.gen)
        for _ in range(nwin):
            row = [self.identity]
            cur = self.identity
            for _ in range((1 << w) - 1):
                cur = self.add(cur, base)
                row.append(cur)
            rows.append(self.batch_to_affine(row))
            for _ in range(w):
                base = self.double(base)
        if not hasattr(self, "_gen_tab"):
            self._gen_tab = {}
        self._gen_tab[key] = rows
        return rows

    def gen_mul(self, k, w=8):
        """k * generator as a Jacobian point, via the fixed-base table."""
        k %= R
        tab = self.gen_table(w)
        acc = self.identity
        i = 0
        mask = (1 << w) - 1
        while k:
            d = k & mask
            if d:
                acc = self.add_mixed(acc, tab[i][d])
            k >>= w
            i += 1
        return acc

    def gen_mul_many(self, ks, w=8):
        return self.batch_to_affine([self.gen_mul(k, w) for k in ks])

    # -- encodings -------------------------------------
    def encode_uncompressed(self, pt):
        n = self.F.nbytes
        if pt is None:
            return bytes([0x40]) + bytes(2 * n - 1)
        return self.F.to_bytes(pt[0]) + self.F.to_bytes(pt[1])

    def decode_uncompressed(self, b, check=True):
        n = self.F.nbytes
        if len(b) != 2 * n:
            raise ValueError("bad length")
        flags = b[0] >> 5
        if flags & 0b100:
            raise ValueError("compression flag set on uncompressed encoding")
        if flags & 0b010:
            if any(b[1:]) or (b[0] & 0x3F):
                raise ValueError("non-zero bytes in identity encoding")
            return None
        if flags & 0b001:
            raise ValueError("sort flag set on uncompressed encoding")
        x = self.F.from_bytes(b[:n])
        y = self.F.from_bytes(b[n:])
        pt = (x, y)
        if check and not self.is_on_curve(pt):
            raise ValueError("point not on curve")
        return pt

    def encode_compressed(self, pt):
        n = self.F.nbytes
        if pt is None:
            return bytes([0xC0]) + bytes(n - 1)
        out = bytearray(self.F.to_bytes(pt[0]))
        out[0] |= 0x80
        if self.F.lex_larger(pt[1]):
            out[0] |= 0x20
        return bytes(out)

    def decode_compressed(self, b):
        n = self.F.nbytes
        if len(b) != n:
            raise ValueError("bad length")
        if not (b[0] & 0x80):
            raise ValueError("compression flag not set")
        if b[0] & 0x40:
            return None
        sort = bool(b[0] & 0x20)
        xb = bytes([b[0] & 0x1F]) + b[1:]
        x = self.F.from_bytes(xb)
        y = self.F.sqrt(self.F.add(self.F.mul(self.F.sqr(x), x), self.b))
        if y is None:
            raise ValueError("not on curve")
        if self.F.lex_larger(y) != sort:
            y = self.F.neg(y)
        return (x, y)


G1 = Curve(FpOps, 4, (G1_X, G1_Y), "G1")
G2 = Curve(Fp2Ops, (4, 4), (G2_X, G2_Y), "G2")


# ----------------------------------------------------------------------------
# Fr helpers
# ----------------------------------------------------------------------------
def fr_to_bytes(a):
    """PrimeField::to_repr for bls12_381::Scalar: 32 bytes little-endian."""
    return (a % R).to_bytes(32, "little")


def fr_from_bytes(b):
    v = int.from_bytes(b, "little")
    if v >= R:
        raise ValueError("scalar not canonical")
    return v


def fr_inv(a):
    return pow(a, R - 2, R)


# ----------------------------------------------------------------------------
# Optimal-ate pairing (test-side only: used to run the Groth16 verification
# equation the reference runs after proving, masp_proofs/src/sapling/prover.rs:148).
# Straightforward tower Fp2 -> Fp6 -> Fp12, affine Miller loop, no tricks.
# ----------------------------------------------------------------------------
_XI = (1, 1)  # Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v)


def _f2_mul_xi(a):
    return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)


def _f6_add(a, b): return tuple(Fp2Ops.add(x, y) for x, y in zip(a, b))
def _f6_sub(a, b): return tuple(Fp2Ops.sub(x, y) for x, y in zip(a, b))
def _f6_neg(a): return tuple(Fp2Ops.neg(x) for x in a)


def _f6_mul(a, b):
    a0, a1, a2 = a; b0, b1, b2 = b
    m = Fp2Ops.mul
    t0 = m(a0, b0); t1 = m(a1, b1); t2 = m(a2, b2)
    c0 = Fp2Ops.add(t0, _f2_mul_xi(Fp2Ops.add(m(a1, b2), m(a2, b1))))
    c1 = Fp2Ops.add(Fp2Ops.add(m(a0, b1), m(a1, b0)), _f2_mul_xi(t2))
    c2 = Fp2Ops.add(Fp2Ops.add(m(a0, b2), m(a2, b0)), t1)
    return (c0, c1, c2)


def _f6_mul_v(a):
    return (_f2_mul_xi(a[2]), a[0], a[1])


def _f6_inv(a):
    a0, a1, a2 = a
    m = Fp2Ops.mul; s = Fp2Ops.sqr
    c0 = Fp2Ops.sub(s(a0), _f2_mul_xi(m(a1, a2)))
    c1 = Fp2Ops.sub(_f2_mul_xi(s(a2)), m(a0, a1))
    c2 = Fp2Ops.sub(s(a1), m(a0, a2))
    t = Fp2Ops.add(m(a0, c0), _f2_mul_xi(Fp2Ops.add(m(a2, c1), m(a1, c2))))
    ti = Fp2Ops.inv(t)
    return (m(c0, ti), m(c1, ti), m(c2, ti))


_F6_ZERO = ((0, 0), (0, 0), (0, 0))
_F6_ONE = ((1, 0), (0, 0), (0, 0))
F12_ONE = (_F6_ONE, _F6_ZERO)


def f12_mul(a, b):
    a0, a1 = a; b0, b1 = b
    t0 = _f6_mul(a0, b0); t1 = _f6_mul(a1, b1)
    c0 = _f6_add(t0, _f6_mul_v(t1))
    c1 = _f6_sub(_f6_sub(_f6_mul(_f6_add(a0, a1), _f6_add(b0, b1)), t0), t1)
    return (c0, c1)


def f12_sqr(a): return f12_mul(a, a)


def f12_inv(a):
    a0, a1 = a
    t = _f6_sub(_f6_mul(a0, a0), _f6_mul_v(_f6_mul(a1, a1)))
    ti = _f6_inv(t)
    return (_f6_mul(a0, ti), _f6_neg(_f6_mul(a1, ti)))


def f12_conj(a): return (a[0], _f6_neg(a[1]))


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_sqr(a); e >>= 1
    return r


def _line_eval(lam, xt, yt, Pp):
    """Sparse Fp12 value of the line through the (untwisted) G2 point
    (xt, yt) with slope lam, evaluated at the G1 point Pp = (xp, yp).

    With the M-twist psi(x, y) = (x / w^2, y / w^3) and after clearing the
    w^3 denominator (a factor in a proper subfield is killed by the final
    exponentiation) the line is
        yp * w^3 - lam * xp * w^2 + (lam * xt - yt)
    with w^2 = v, w^3 = v * w.
    """
    xp, yp = Pp
    c = Fp2Ops.sub(Fp2Ops.mul(lam, xt), yt)
    c0 = (c, Fp2Ops.muli(Fp2Ops.neg(lam), xp), (0, 0))       # 1, v, v^2
    c1 = ((0, 0), (yp % P, 0), (0, 0))                       # w, v w, v^2 w
    return (c0, c1)


def miller_loop(Pp, Q):
    """f_{|x|,Q}(P), conjugated because the BLS parameter is negative."""
    if Pp is None or Q is None:
        return F12_ONE
    F = Fp2Ops
    f = F12_ONE
    T = Q
    for bit in bin(BLS_X)[3:]:
        lam = F.mul(F.muli(F.sqr(T[0]), 3), F.inv(F.muli(T[1], 2)))
        f = f12_mul(f12_sqr(f), _line_eval(lam, T[0], T[1], Pp))
        x3 = F.sub(F.sqr(lam), F.muli(T[0], 2))
        y3 = F.sub(F.mul(lam, F.sub(T[0], x3)), T[1])
        T = (x3, y3)
        if bit == "1":
            lam = F.mul(F.sub(T[1], Q[1]), F.inv(F.sub(T[0], Q[0])))
            f = f12_mul(f, _line_eval(lam, T[0], T[1], Pp))
            x3 = F.sub(F.sub(F.sqr(lam), T[0]), Q[0])
            y3 = F.sub(F.mul(lam, F.sub(T[0], x3)), T[1])
            T = (x3, y3)
    return f12_conj(f)


def final_exponentiation(f):
    return f12_pow(f, (P ** 12 - 1) // R)


def pairing(Pp, Q):
    return final_exponentiation(miller_loop(Pp, Q))


def multi_pairing_is_one(pairs):
    f = F12_ONE
    for Pp, Q in pairs:
        f = f12_mul(f, miller_loop(Pp, Q))
    return final_exponentiation(f) == F12_ONE
