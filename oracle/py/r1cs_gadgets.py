"""bellman's constraint-system recorder and gadgets, restated on Python ints.

TEST INFRASTRUCTURE ONLY (oracle).  Nothing under masp_b200/ may import this.

The gadget library is bellman's (crate `nam-bellperson` / `bellpepper-core`,
reference Cargo.lock:154-155, 1355-1358; imported by the reference at
masp_proofs/src/circuit/sapling.rs:19, ecc.rs:9-13, pedersen_hash.rs:4-6); it
is not vendored, so this follows SURVEY.md Appendix B.  Unlike the proving
arithmetic this layer IS pinned by the reference's own tests: the structural
hash of a synthesised circuit must equal the strings at
masp_proofs/src/circuit/convert.rs:218-224 and sapling.rs:730-741, 1024-1045
(tests/test_circuits.py).

A linear combination is a list of (variable, coefficient); a variable is
('I', i) or ('A', i); ('I', 0) is ONE.
"""
import hashlib

from .bls12_381 import R

ONE = ("I", 0)


def lc_scale(lc, k):
    return [(v, c * k % R) for v, c in lc]


def lc_add(*lcs):
    out = []
    for lc in lcs:
        out += lc
    return out


def lc_neg(lc):
    return [(v, (-c) % R) for v, c in lc]


class ConstraintSystem:
    """TestConstraintSystem + ProvingAssignment in one: records constraints,
    assignments and the three density trackers."""

    def __init__(self):
        self.inputs = [1]
        self.aux = []
        self.constraints = []

    def alloc(self, value):
        self.aux.append(value % R)
        return ("A", len(self.aux) - 1)

    def alloc_input(self, value):
        self.inputs.append(value % R)
        return ("I", len(self.inputs) - 1)

    def enforce(self, a, b, c):
        self.constraints.append((a, b, c))

    def value(self, var):
        return self.inputs[var[1]] if var[0] == "I" else self.aux[var[1]]

    def eval(self, lc):
        return sum(c * self.value(v) for v, c in lc) % R

    # -- TestConstraintSystem ------------------------------------------------
    def num_constraints(self):
        return len(self.constraints)

    def num_inputs(self):
        return len(self.inputs)

    def which_is_unsatisfied(self):
        for i, (a, b, c) in enumerate(self.constraints):
            if (self.eval(a) * self.eval(b) - self.eval(c)) % R:
                return i
        return None

    def is_satisfied(self):
        return self.which_is_unsatisfied() is None

    @staticmethod
    def _canonical(lc):
        acc = {}
        for v, c in lc:
            acc[v] = (acc.get(v, 0) + c) % R
        terms = [(v, c) for v, c in acc.items() if c]
        terms.sort(key=lambda t: (0 if t[0][0] == "I" else 1, t[0][1]))
        return terms

    def hash(self):
        """TestConstraintSystem::hash: BLAKE2s-256 over the
        counts and every constraint's canonical A, B, C."""
        h = hashlib.blake2s(digest_size=32)
        for n in (len(self.inputs), len(self.aux), len(self.constraints)):
            h.update(n.to_bytes(8, "big"))
        for abc in self.constraints:
            for lc in abc:
                terms = self._canonical(lc)
                h.update(len(terms).to_bytes(8, "big"))
                for (kind, idx), coeff in terms:
                    h.update(kind.encode())
                    h.update(idx.to_bytes(8, "big"))
                    h.update(coeff.to_bytes(32, "big"))
        return h.hexdigest()

    # -- ProvingAssignment ------------------------------------------------------
    def proving_assignment(self):
        """(a, b, c evaluations incl. the extra input rows, densities) exactly as
        bellman's prover records them (SURVEY Appendix A 'Synthesis')."""
        a_aux = [False] * len(self.aux)
        b_inp = [False] * len(self.inputs)
        b_aux = [False] * len(self.aux)
        a, b, c = [], [], []
        for (A, B, C) in self.constraints:
            for (kind, i), coeff in A:
                if coeff % R and kind == "A":
                    a_aux[i] = True
            for (kind, i), coeff in B:
                if coeff % R:
                    (b_aux if kind == "A" else b_inp)[i] = True
            a.append(self.eval(A)); b.append(self.eval(B)); c.append(self.eval(C))
        for i in range(len(self.inputs)):
            a.append(self.inputs[i]); b.append(0); c.append(0)
        return a, b, c, (a_aux, b_inp, b_aux)

    def to_r1cs(self):
        from .groth16 import R1CS
        return R1CS(len(self.inputs), len(self.aux), list(self.constraints))


# ----------------------------------------------------------------------------
# booleans
# ----------------------------------------------------------------------------
class AllocatedBit:
    def __init__(self, var, value):
        self.var = var
        self.value = value

    @staticmethod
    def alloc(cs, value):
        v = cs.alloc(1 if value else 0)
        cs.enforce([(ONE, 1), (v, R - 1)], [(v, 1)], [])
        return AllocatedBit(v, bool(value))

    @staticmethod
    def alloc_conditionally(cs, value, must_be_false):
        v = cs.alloc(1 if value else 0)
        cs.enforce([(ONE, 1), (must_be_false.var, R - 1), (v, R - 1)], [(v, 1)], [])
        return AllocatedBit(v, bool(value))

    @staticmethod
    def and_(cs, a, b):
        v = cs.alloc(1 if (a.value and b.value) else 0)
        cs.enforce([(a.var, 1)], [(b.var, 1)], [(v, 1)])
        return AllocatedBit(v, a.value and b.value)

    @staticmethod
    def and_not(cs, a, b):
        val = a.value and not b.value
        v = cs.alloc(1 if val else 0)
        cs.enforce([(a.var, 1)], [(ONE, 1), (b.var, R - 1)], [(v, 1)])
        return AllocatedBit(v, val)

    @staticmethod
    def nor(cs, a, b):
        val = (not a.value) and (not b.value)
        v = cs.alloc(1 if val else 0)
        cs.enforce([(ONE, 1), (a.var, R - 1)], [(ONE, 1), (b.var, R - 1)], [(v, 1)])
        return AllocatedBit(v, val)

    @staticmethod
    def xor(cs, a, b):
        val = a.value != b.value
        v = cs.alloc(1 if val else 0)
        cs.enforce([(a.var, 1), (a.var, 1)], [(b.var, 1)], [(a.var, 1), (b.var, 1), (v, R - 1)])
        return AllocatedBit(v, val)


class Boolean:
    """kind: 'is' | 'not' | 'const'."""

    def __init__(self, kind, bit=None, const=None):
        self.kind = kind
        self.bit = bit
        self.const = const

    @staticmethod
    def constant(b):
        return Boolean("const", const=bool(b))

    @staticmethod
    def from_bit(bit):
        return Boolean("is", bit=bit)

    @property
    def value(self):
        if self.kind == "const":
            return self.const
        return self.bit.value if self.kind == "is" else not self.bit.value

    def not_(self):
        if self.kind == "const":
            return Boolean.constant(not self.const)
        return Boolean("not" if self.kind == "is" else "is", bit=self.bit)

    def lc(self, k):
        k %= R
        if self.kind == "const":
            return [(ONE, k)] if self.const else []
        if self.kind == "is":
            return [(self.bit.var, k)]
        return [(ONE, k), (self.bit.var, (-k) % R)]

    @staticmethod
    def and_(cs, a, b):
        if a.kind == "const":
            return b if a.const else Boolean.constant(False)
        if b.kind == "const":
            return a if b.const else Boolean.constant(False)
        if a.kind == "is" and b.kind == "is":
            return Boolean.from_bit(AllocatedBit.and_(cs, a.bit, b.bit))
        if a.kind == "is" and b.kind == "not":
            return Boolean.from_bit(AllocatedBit.and_not(cs, a.bit, b.bit))
        if a.kind == "not" and b.kind == "is":
            return Boolean.from_bit(AllocatedBit.and_not(cs, b.bit, a.bit))
        return Boolean.from_bit(AllocatedBit.nor(cs, a.bit, b.bit))

    @staticmethod
    def xor(cs, a, b):
        if a.kind == "const":
            return b.not_() if a.const else b
        if b.kind == "const":
            return a.not_() if b.const else a
        if a.kind == "is" and b.kind == "not":
            return Boolean.from_bit(AllocatedBit.xor(cs, a.bit, b.bit)).not_()
        if a.kind == "not" and b.kind == "is":
            return Boolean.from_bit(AllocatedBit.xor(cs, b.bit, a.bit)).not_()
        return Boolean.from_bit(AllocatedBit.xor(cs, a.bit, b.bit))

    @staticmethod
    def enforce_equal(cs, a, b):
        if a.kind == "const" and b.kind == "const":
            assert a.const == b.const
            return
        cs.enforce([], [], lc_add(a.lc(1), lc_neg(b.lc(1))))


def u64_into_boolean_vec_le(cs, value):
    return [Boolean.from_bit(AllocatedBit.alloc(cs, (value >> i) & 1)) for i in range(64)]


def field_into_boolean_vec_le(cs, value, num_bits):
    """field_into_allocated_bits_le / the in-tree copy
    masp_proofs/src/circuit/gadgets.rs:6-50: NUM_BITS allocations, LSB first."""
    return [Boolean.from_bit(AllocatedBit.alloc(cs, (value >> i) & 1)) for i in range(num_bits)]


# ----------------------------------------------------------------------------
# numbers
# ----------------------------------------------------------------------------
class AllocatedNum:
    def __init__(self, var, value):
        self.var = var
        self.value = value % R

    @staticmethod
    def alloc(cs, value):
        return AllocatedNum(cs.alloc(value), value)

    def mul(self, cs, other):
        out = AllocatedNum.alloc(cs, self.value * other.value)
        cs.enforce([(self.var, 1)], [(other.var, 1)], [(out.var, 1)])
        return out

    def square(self, cs):
        out = AllocatedNum.alloc(cs, self.value * self.value)
        cs.enforce([(self.var, 1)], [(self.var, 1)], [(out.var, 1)])
        return out

    def assert_nonzero(self, cs):
        assert self.value != 0
        inv = cs.alloc(pow(self.value, R - 2, R))
        cs.enforce([(self.var, 1)], [(inv, 1)], [(ONE, 1)])

    def inputize(self, cs):
        inp = cs.alloc_input(self.value)
        cs.enforce([(inp, 1)], [(ONE, 1)], [(self.var, 1)])

    @staticmethod
    def conditionally_reverse(cs, a, b, cond):
        c = AllocatedNum.alloc(cs, b.value if cond.value else a.value)
        cs.enforce([(a.var, 1), (b.var, R - 1)], cond.lc(1), [(a.var, 1), (c.var, R - 1)])
        d = AllocatedNum.alloc(cs, a.value if cond.value else b.value)
        cs.enforce([(b.var, 1), (a.var, R - 1)], cond.lc(1), [(b.var, 1), (d.var, R - 1)])
        return c, d

    def to_bits_le(self, cs):
        bits = [AllocatedBit.alloc(cs, (self.value >> i) & 1) for i in range(255)]
        lc = [(b.var, pow(2, i, R)) for i, b in enumerate(bits)] + [(self.var, R - 1)]
        cs.enforce([], [], lc)
        return [Boolean.from_bit(b) for b in bits]

    def to_bits_le_strict(self, cs):
        """Bits of self with the proof that they encode a value <= r - 1."""
        bound = R - 1
        result = []  # big-endian
        last_run = None
        current_run = []
        for i in range(254, -1, -1):  # 255-bit modulus: the top (256th) bit of r - 1 is unset and skipped
            a_bit = (self.value >> i) & 1
            if (bound >> i) & 1:
                bit = AllocatedBit.alloc(cs, a_bit)
                current_run.append(bit)
                result.append(bit)
            else:
                if current_run:
                    if last_run is not None:
                        current_run.append(last_run)
                    cur = current_run[0]
                    for nxt in current_run[1:]:
                        cur = AllocatedBit.and_(cs, cur, nxt)
                    last_run = cur
                    current_run = []
                result.append(AllocatedBit.alloc_conditionally(cs, a_bit, last_run))
        assert not current_run
        lc = [(b.var, pow(2, i, R)) for i, b in enumerate(reversed(result))] + [(self.var, R - 1)]
        cs.enforce([], [], lc)
        return [Boolean.from_bit(b) for b in reversed(result)]


class Num:
    """A bare linear combination with its value."""

    def __init__(self, lc=None, value=0):
        self.lc_terms = lc or []
        self.value = value % R

    @staticmethod
    def from_allocated(n):
        return Num([(n.var, 1)], n.value)

    def add_bool_with_coeff(self, bit, coeff):
        return Num(self.lc_terms + bit.lc(coeff), self.value + (coeff if bit.value else 0))

    def lc(self, k=1):
        return lc_scale(self.lc_terms, k)


# ----------------------------------------------------------------------------
# lookups
# ----------------------------------------------------------------------------
def _synth(window, consts):
    n = 1 << window
    a = [0] * n
    for i, c in enumerate(consts):
        cur = (c - a[i]) % R
        a[i] = cur
        for j in range(i + 1, n):
            if j & i == i:
                a[j] = (a[j] + cur) % R
    return a


def lookup3_xy(cs, bits, coords):
    """3-bit window lookup of an (x, y) pair out of 8 (bellman lookup3_xy)."""
    i = (1 if bits[0].value else 0) | (2 if bits[1].value else 0) | (4 if bits[2].value else 0)
    res_x = AllocatedNum.alloc(cs, coords[i][0])
    res_y = AllocatedNum.alloc(cs, coords[i][1])
    xc = _synth(3, [c[0] for c in coords])
    yc = _synth(3, [c[1] for c in coords])
    precomp = Boolean.and_(cs, bits[1], bits[2])
    for co, res in ((xc, res_x), (yc, res_y)):
        a = lc_add([(ONE, co[1])], bits[1].lc(co[3]), bits[2].lc(co[5]), precomp.lc(co[7]))
        c = lc_add([(res.var, 1), (ONE, (-co[0]) % R)], lc_neg(bits[1].lc(co[2])), lc_neg(bits[2].lc(co[4])),
                   lc_neg(precomp.lc(co[6])))
        cs.enforce(a, bits[0].lc(1), c)
    return res_x, res_y


def lookup3_xy_with_conditional_negation(cs, bits, coords):
    """2-bit lookup of (x, y) out of 4 with y negated when bits[2] is set."""
    i = (1 if bits[0].value else 0) | (2 if bits[1].value else 0)
    yv = coords[i][1]
    if bits[2].value:
        yv = (-yv) % R
    y = AllocatedNum.alloc(cs, yv)
    xc = _synth(2, [c[0] for c in coords])
    yc = _synth(2, [c[1] for c in coords])
    precomp = Boolean.and_(cs, bits[0], bits[1])
    x = (Num().add_bool_with_coeff(Boolean.constant(True), xc[0]).add_bool_with_coeff(bits[0], xc[1])
         .add_bool_with_coeff(bits[1], xc[2]).add_bool_with_coeff(precomp, xc[3]))
    y_lc = lc_add(precomp.lc(yc[3]), bits[1].lc(yc[2]), bits[0].lc(yc[1]), [(ONE, yc[0])])
    cs.enforce(lc_add(y_lc, y_lc), bits[2].lc(1), lc_add(y_lc, [(y.var, R - 1)]))
    return x, Num.from_allocated(y)


def pack_into_inputs(cs, bits):
    """multipack::pack_into_inputs: 254 bits per public input."""
    for k in range(0, len(bits), 254):
        num = Num()
        coeff = 1
        for bit in bits[k:k + 254]:
            num = num.add_bool_with_coeff(bit, coeff)
            coeff = coeff * 2 % R
        inp = cs.alloc_input(num.value)
        cs.enforce(num.lc(1), [(ONE, 1)], [(inp, 1)])


# ----------------------------------------------------------------------------
# UInt32, MultiEq, BLAKE2s (bellman gadgets::{uint32, multieq, blake2s})
# ----------------------------------------------------------------------------
class MultiEq:
    """Packs several small equalities into one constraint (254-bit capacity)."""

    def __init__(self, cs):
        self.cs = cs
        self.bits_used = 0
        self.lhs = []
        self.rhs = []

    def accumulate(self):
        self.cs.enforce(self.lhs, [(ONE, 1)], self.rhs)
        self.lhs, self.rhs, self.bits_used = [], [], 0

    def enforce_equal(self, num_bits, lhs, rhs):
        if 254 <= self.bits_used + num_bits:
            self.accumulate()
        coeff = pow(2, self.bits_used, R)
        self.lhs = self.lhs + lc_scale(lhs, coeff)
        self.rhs = self.rhs + lc_scale(rhs, coeff)
        self.bits_used += num_bits

    def close(self):
        if self.bits_used > 0:
            self.accumulate()


class UInt32:
    def __init__(self, bits):
        assert len(bits) == 32
        self.bits = bits  # LSB first

    @staticmethod
    def constant(v):
        return UInt32([Boolean.constant((v >> i) & 1) for i in range(32)])

    @property
    def value(self):
        return sum((1 << i) for i, b in enumerate(self.bits) if b.value)

    def rotr(self, k):
        return UInt32([self.bits[(i + k) % 32] for i in range(32)])

    def xor(self, cs, other):
        return UInt32([Boolean.xor(cs, a, b) for a, b in zip(self.bits, other.bits)])

    @staticmethod
    def addmany(meq, operands):
        cs = meq.cs
        max_value = len(operands) * 0xFFFFFFFF
        total = sum(op.value for op in operands)
        lc = []
        all_constants = True
        for op in operands:
            for i, bit in enumerate(op.bits):
                lc = lc + bit.lc(1 << i)
                all_constants &= bit.kind == "const"
        if all_constants:
            return UInt32.constant(total & 0xFFFFFFFF)
        result_bits = []
        result_lc = []
        i = 0
        while max_value:
            b = AllocatedBit.alloc(cs, (total >> i) & 1)
            result_lc.append((b.var, pow(2, i, R)))
            result_bits.append(Boolean.from_bit(b))
            max_value >>= 1
            i += 1
        meq.enforce_equal(i, lc, result_lc)
        return UInt32(result_bits[:32])


_BLAKE2S_IV = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]
_SIGMA = [
    [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
    [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
    [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
    [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
    [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0]]


def _mixing_g(meq, v, a, b, c, d, x, y):
    cs = meq.cs
    v[a] = UInt32.addmany(meq, [v[a], v[b], x])
    v[d] = v[d].xor(cs, v[a]).rotr(16)
    v[c] = UInt32.addmany(meq, [v[c], v[d]])
    v[b] = v[b].xor(cs, v[c]).rotr(12)
    v[a] = UInt32.addmany(meq, [v[a], v[b], y])
    v[d] = v[d].xor(cs, v[a]).rotr(8)
    v[c] = UInt32.addmany(meq, [v[c], v[d]])
    v[b] = v[b].xor(cs, v[c]).rotr(7)


def _blake2s_compression(cs, h, m, t, final):
    v = list(h) + [UInt32.constant(x) for x in _BLAKE2S_IV]
    v[12] = v[12].xor(cs, UInt32.constant(t & 0xFFFFFFFF))
    v[13] = v[13].xor(cs, UInt32.constant((t >> 32) & 0xFFFFFFFF))
    if final:
        v[14] = v[14].xor(cs, UInt32.constant(0xFFFFFFFF))
    meq = MultiEq(cs)
    for i in range(10):
        s = _SIGMA[i % 10]
        _mixing_g(meq, v, 0, 4, 8, 12, m[s[0]], m[s[1]])
        _mixing_g(meq, v, 1, 5, 9, 13, m[s[2]], m[s[3]])
        _mixing_g(meq, v, 2, 6, 10, 14, m[s[4]], m[s[5]])
        _mixing_g(meq, v, 3, 7, 11, 15, m[s[6]], m[s[7]])
        _mixing_g(meq, v, 0, 5, 10, 15, m[s[8]], m[s[9]])
        _mixing_g(meq, v, 1, 6, 11, 12, m[s[10]], m[s[11]])
        _mixing_g(meq, v, 2, 7, 8, 13, m[s[12]], m[s[13]])
        _mixing_g(meq, v, 3, 4, 9, 14, m[s[14]], m[s[15]])
    meq.close()
    for i in range(8):
        h[i] = h[i].xor(cs, v[i]).xor(cs, v[i + 8])


def blake2s(cs, input_bits, personalization):
    """BLAKE2s-256 of a bit string (bytes little-endian bit order), 8-byte
    personalization, no key.  Returns 256 output Booleans."""
    assert len(personalization) == 8 and len(input_bits) % 8 == 0
    h = [UInt32.constant(x) for x in _BLAKE2S_IV]
    h[0] = UInt32.constant(_BLAKE2S_IV[0] ^ 0x01010000 ^ 32)
    h[6] = UInt32.constant(_BLAKE2S_IV[6] ^ int.from_bytes(personalization[0:4], "little"))
    h[7] = UInt32.constant(_BLAKE2S_IV[7] ^ int.from_bytes(personalization[4:8], "little"))
    blocks = []
    for k in range(0, len(input_bits), 512):
        chunk = input_bits[k:k + 512]
        words = []
        for w in range(0, len(chunk), 32):
            wb = chunk[w:w + 32]
            wb = wb + [Boolean.constant(False)] * (32 - len(wb))
            words.append(UInt32(wb))
        while len(words) < 16:
            words.append(UInt32.constant(0))
        blocks.append(words)
    if not blocks:
        blocks.append([UInt32.constant(0)] * 16)
    for i, block in enumerate(blocks[:-1]):
        _blake2s_compression(cs, h, block, (i + 1) * 64, False)
    _blake2s_compression(cs, h, blocks[-1], len(input_bits) // 8, True)
    out = []
    for word in h:
        out += word.bits
    return out
