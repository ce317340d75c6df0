"""The MASP circuits (Convert first) and their Jubjub gadgets on the recorder
of r1cs_gadgets.py, with native Jubjub arithmetic for the witness values.

TEST INFRASTRUCTURE ONLY (oracle).  Restates, gadget by gadget and in the same
emission order:
  masp_proofs/src/circuit/ecc.rs        (witness :130-143, interpret :250-276,
      double :278-371, add :374-471, conditionally_select :147-198, mul :203-248,
      repr :112-126, fixed_base_multiplication :27-73, Montgomery add :543-617,
      into_edwards :483-531)
  masp_proofs/src/circuit/pedersen_hash.rs:19-103
  masp_proofs/src/circuit/sapling.rs:71-137   (expose_value_commitment)
  masp_proofs/src/circuit/convert.rs:29-128   (Convert::synthesize)
  masp_proofs/src/constants.rs:10-38, 76-94, 100-173 (curve constants, window tables)
Generator coordinates are data taken from masp_primitives/src/constants.rs:50-251.

Pinned by the reference's own test masp_proofs/src/circuit/convert.rs:218-224:
47 358 constraints, 4 inputs, cs.hash() = f74b47ef...f814.
"""
from .bls12_381 import R
from .r1cs_gadgets import (ONE, AllocatedBit, AllocatedNum, Boolean, Num, lc_add, lc_neg, lookup3_xy,
                           lookup3_xy_with_conditional_negation, u64_into_boolean_vec_le,
                           field_into_boolean_vec_le)

# masp_proofs/src/constants.rs:10-38
EDWARDS_D = 0x2A9318E74BFA2B48F5FD9207E6BD7FD4292D7F6D37579D2601065FD6D6343EB1
MONTGOMERY_A = 0xA002
MONTGOMERY_SCALE = 0x2762DE61E862645E31DE341E77D764E5CE4069703DA88ABD8F4535F7CF82B8D9
JUBJUB_ORDER = 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7  # prime subgroup order
JUBJUB_FR_BITS = 252

# masp_primitives/src/constants.rs:50-251 (u, v)
PROOF_GENERATION_KEY_GENERATOR = (
    0x4CAAEACAAF28ED4B4BA1F065E719FD031E24F83267F15ABD5F3C723AA2531B66,
    0x00930D67D6906365C654DFDD36004DE936B49C71A2AF0708FE6F96BEC575BFF8)
NOTE_COMMITMENT_RANDOMNESS_GENERATOR = (
    0x434C9BE15267B091C6DE7556ABB84082CD80EDF5FE44C7BFFC033FA2BF88CB2E,
    0x29E2926993D3BC736D277197E97AF8F0690B295C66B85C64C6B8DAA0EE22AEED)
NULLIFIER_POSITION_GENERATOR = (
    0x50DE6D98FEE5282F84678DC2D85293DF1E09674F28A4B844AAFEE844265FC1E7,
    0x03260F0BF1244050F3F70DC31AFE799D226945AEE96DFE0AED034E3EE13A1EB3)
VALUE_COMMITMENT_RANDOMNESS_GENERATOR = (
    0x1C6DA0CE9A5E5FDBCFA86026B8D99BE991CC3E3835675450DD93D364CB8CEC7E,
    0x555F11F9B720D50BBC900CD4B8AE1150F94C2DAA360302FE28E5FCE99CE692D0)
SPENDING_KEY_GENERATOR = (
    0x5B389522A9E81532F831C2B19FEC602639F5B03380AF6020EC75293D81248452,
    0x0CBC5F9F1E52E0AB75DEFECFF1F49EF22012D031F624FD5214B62623A186B4B1)
PEDERSEN_HASH_GENERATORS = [
    (0x113DE62BE6E0D32398BA470B0D28801B5C22A82A281C91811010503570C3EBF6,
     0x5059678472ABB6AE15CEA14BC9F6B04B2BA3032D7064D633F031EDFF274EFB14),
    (0x08C02A4C57F7F2CFFC7CBEA3C311F67F0A0DF10182A290FDB9EFA2CB80331936,
     0x2E560A50271FD3FC4DC07857131F22A0EC376560C925452DDAF19AC3AB182662),
    (0x210F22D61B65767D413BC3C44E7AABE0DF0694E57C6CBC03C93573B98709291E,
     0x3F46B3371CFF7474FB33884C42727482C6262ED4231796594781E2656B1DDAAD),
    (0x274E99B16D4AF911A02F0D3F7AAD771D2BCC52DBBA0EBF3ACF0BC7224A63D094,
     0x31F5E34F0804A8746B15EC6E59478694FD0153CFE15EC653E82E9061620A1DF4),
    (0x3CA8B98873E5D19E50AA77AD2F57D2F77058160B9AFAAFAFC64E25CA51961B53,
     0x10609CE821A5A292238AF7C9376608D65EB152C4606BEB7E9DAB539B32327842),
    (0x1AB3FE2AC6B3FF8ADB3FF866EAF1BC855BDD5C30D83781F0F0EF2A816469118E,
     0x2031E442C4AF8277D5681F2F5C740D19A6B5863148627619E7C079B4E48233F5),
]
PEDERSEN_HASH_CHUNKS_PER_GENERATOR = 63
FIXED_BASE_CHUNKS_PER_GENERATOR = 84


def inv(x):
    return pow(x % R, R - 2, R)


# ----------------------------------------------------------------------------
# native Jubjub: -u^2 + v^2 = 1 + d u^2 v^2 over Fr (affine, complete formulas)
# ----------------------------------------------------------------------------
JJ_IDENTITY = (0, 1)


def jj_on_curve(p):
    u, v = p
    return (-u * u + v * v) % R == (1 + EDWARDS_D * u * u % R * v * v) % R


def jj_add(p, q):
    u1, v1 = p
    u2, v2 = q
    t = EDWARDS_D * u1 % R * u2 % R * v1 % R * v2 % R
    return ((u1 * v2 + v1 * u2) * inv(1 + t) % R, (v1 * v2 + u1 * u2) * inv(1 - t) % R)


def jj_double(p):
    return jj_add(p, p)


def jj_mul(p, k):
    acc = JJ_IDENTITY
    for bit in bin(k)[2:] if k else "":
        acc = jj_double(acc)
        if bit == "1":
            acc = jj_add(acc, p)
    return acc


def jj_to_montgomery(p):
    """masp_proofs/src/constants.rs:100-141."""
    u, v = p
    if v == 1:
        return None
    if u == 0:
        return (0, 0)
    x = (1 + v) * inv(1 - v) % R
    y = x * inv(u) % R
    return (x, y * MONTGOMERY_SCALE % R)


def pedersen_hash_native(personalization_bits, bits):
    """masp_primitives/src/sapling/pedersen_hash.rs:31-117 on booleans."""
    allbits = list(personalization_bits) + list(bits)
    result = JJ_IDENTITY
    pos = 0
    seg = 0
    while pos < len(allbits):
        acc = 0
        cur = 1
        chunks = 0
        while pos < len(allbits) and chunks < PEDERSEN_HASH_CHUNKS_PER_GENERATOR:
            a = allbits[pos]
            b = allbits[pos + 1] if pos + 1 < len(allbits) else False
            c = allbits[pos + 2] if pos + 2 < len(allbits) else False
            pos += 3
            tmp = cur
            if a:
                tmp += cur
            if b:
                tmp += 2 * cur
            if c:
                tmp = -tmp
            acc += tmp
            cur *= 16
            chunks += 1
        result = jj_add(result, jj_mul(PEDERSEN_HASH_GENERATORS[seg], acc % JUBJUB_ORDER))
        seg += 1
    return result


def merkle_personalization(depth):
    return [bool((depth >> i) & 1) for i in range(6)]


NOTE_COMMITMENT_PERSONALIZATION = [True] * 6

_cache = {}


def fixed_generator_table(gen):
    """3-bit window tables [0, g, ..., 7g] for 84 windows (constants.rs:76-94)."""
    if gen in _cache:
        return _cache[gen]
    windows = []
    g0 = gen
    for _ in range(FIXED_BASE_CHUNKS_PER_GENERATOR):
        coeffs = [(0, 1)]
        g = g0
        for _ in range(7):
            coeffs.append(g)
            g = jj_add(g, g0)
        windows.append(coeffs)
        g0 = g  # 8 * g0
    _cache[gen] = windows
    return windows


def pedersen_circuit_generators():
    """2-bit window tables [g, 2g, 3g, 4g] in Montgomery coordinates, 63 windows
    16x apart, for each of the six segment generators (constants.rs:143-173)."""
    if "pedersen" in _cache:
        return _cache["pedersen"]
    out = []
    for gen in PEDERSEN_HASH_GENERATORS:
        windows = []
        g0 = gen
        for _ in range(PEDERSEN_HASH_CHUNKS_PER_GENERATOR):
            coeffs = []
            g = g0
            for _ in range(4):
                coeffs.append(jj_to_montgomery(g))
                g = jj_add(g, g0)
            windows.append(coeffs)
            for _ in range(4):
                g0 = jj_double(g0)
        out.append(windows)
    _cache["pedersen"] = out
    return out


# ----------------------------------------------------------------------------
# ecc gadgets
# ----------------------------------------------------------------------------
class EdwardsPoint:
    def __init__(self, u, v):
        self.u = u
        self.v = v

    @staticmethod
    def interpret(cs, u, v):
        u2 = u.square(cs)
        v2 = v.square(cs)
        u2v2 = u2.mul(cs, v2)
        cs.enforce([(u2.var, R - 1), (v2.var, 1)], [(ONE, 1)], [(ONE, 1), (u2v2.var, EDWARDS_D)])
        return EdwardsPoint(u, v)

    @staticmethod
    def witness(cs, p):
        u = AllocatedNum.alloc(cs, p[0])
        v = AllocatedNum.alloc(cs, p[1])
        return EdwardsPoint.interpret(cs, u, v)

    def inputize(self, cs):
        self.u.inputize(cs)
        self.v.inputize(cs)

    def repr(self, cs):
        u = self.u.to_bits_le_strict(cs)
        v = self.v.to_bits_le_strict(cs)
        return v + [u[0]]

    def double(self, cs):
        u, v = self.u, self.v
        t = AllocatedNum.alloc(cs, (u.value + v.value) * (u.value + v.value))
        cs.enforce([(u.var, 1), (v.var, 1)], [(u.var, 1), (v.var, 1)], [(t.var, 1)])
        a = u.mul(cs, v)
        c = AllocatedNum.alloc(cs, a.value * a.value % R * EDWARDS_D)
        cs.enforce([(a.var, EDWARDS_D)], [(a.var, 1)], [(c.var, 1)])
        u3 = AllocatedNum.alloc(cs, 2 * a.value * inv(1 + c.value))
        cs.enforce([(ONE, 1), (c.var, 1)], [(u3.var, 1)], [(a.var, 1), (a.var, 1)])
        v3 = AllocatedNum.alloc(cs, (t.value - 2 * a.value) * inv(1 - c.value))
        cs.enforce([(ONE, 1), (c.var, R - 1)], [(v3.var, 1)], [(t.var, 1), (a.var, R - 1), (a.var, R - 1)])
        return EdwardsPoint(u3, v3)

    def add(self, cs, other):
        big_u = AllocatedNum.alloc(cs, (self.u.value + self.v.value) * (other.u.value + other.v.value))
        cs.enforce([(self.u.var, 1), (self.v.var, 1)], [(other.u.var, 1), (other.v.var, 1)], [(big_u.var, 1)])
        a = other.v.mul(cs, self.u)
        b = other.u.mul(cs, self.v)
        c = AllocatedNum.alloc(cs, a.value * b.value % R * EDWARDS_D)
        cs.enforce([(a.var, EDWARDS_D)], [(b.var, 1)], [(c.var, 1)])
        u3 = AllocatedNum.alloc(cs, (a.value + b.value) * inv(1 + c.value))
        cs.enforce([(ONE, 1), (c.var, 1)], [(u3.var, 1)], [(a.var, 1), (b.var, 1)])
        v3 = AllocatedNum.alloc(cs, (big_u.value - a.value - b.value) * inv(1 - c.value))
        cs.enforce([(ONE, 1), (c.var, R - 1)], [(v3.var, 1)], [(big_u.var, 1), (a.var, R - 1), (b.var, R - 1)])
        return EdwardsPoint(u3, v3)

    def conditionally_select(self, cs, condition):
        u_prime = AllocatedNum.alloc(cs, self.u.value if condition.value else 0)
        cs.enforce([(self.u.var, 1)], condition.lc(1), [(u_prime.var, 1)])
        v_prime = AllocatedNum.alloc(cs, self.v.value if condition.value else 1)
        cs.enforce([(self.v.var, 1)], condition.lc(1), lc_add([(v_prime.var, 1)], lc_neg(condition.not_().lc(1))))
        return EdwardsPoint(u_prime, v_prime)

    def mul(self, cs, by):
        curbase = None
        result = None
        for bit in by:
            curbase = self if curbase is None else curbase.double(cs)
            thisbase = curbase.conditionally_select(cs, bit)
            result = thisbase if result is None else result.add(cs, thisbase)
        return result


def fixed_base_multiplication(cs, gen, by):
    table = fixed_generator_table(gen)
    result = None
    false = Boolean.constant(False)
    for i in range(0, len(by), 3):
        window = table[i // 3]
        chunk = by[i:i + 3] + [false] * (3 - len(by[i:i + 3]))
        u, v = lookup3_xy(cs, chunk, window)
        p = EdwardsPoint(u, v)
        result = p if result is None else result.add(cs, p)
    return result


class MontgomeryPoint:
    def __init__(self, x, y):
        self.x = x  # Num
        self.y = y

    def into_edwards(self, cs):
        u = AllocatedNum.alloc(cs, self.x.value * MONTGOMERY_SCALE % R * inv(self.y.value))
        cs.enforce(self.y.lc(1), [(u.var, 1)], self.x.lc(MONTGOMERY_SCALE))
        v = AllocatedNum.alloc(cs, (self.x.value - 1) * inv(self.x.value + 1))
        cs.enforce(lc_add(self.x.lc(1), [(ONE, 1)]), [(v.var, 1)], lc_add(self.x.lc(1), [(ONE, R - 1)]))
        return EdwardsPoint(u, v)

    def add(self, cs, other):
        lam = AllocatedNum.alloc(cs, (other.y.value - self.y.value) * inv(other.x.value - self.x.value))
        cs.enforce(lc_add(other.x.lc(1), lc_neg(self.x.lc(1))), [(lam.var, 1)],
                   lc_add(other.y.lc(1), lc_neg(self.y.lc(1))))
        xprime = AllocatedNum.alloc(cs, lam.value * lam.value - MONTGOMERY_A - self.x.value - other.x.value)
        cs.enforce([(lam.var, 1)], [(lam.var, 1)],
                   lc_add([(ONE, MONTGOMERY_A)], self.x.lc(1), other.x.lc(1), [(xprime.var, 1)]))
        yprime = AllocatedNum.alloc(cs, -((xprime.value - self.x.value) * lam.value + self.y.value))
        cs.enforce(lc_add(self.x.lc(1), [(xprime.var, R - 1)]), [(lam.var, 1)],
                   lc_add([(yprime.var, 1)], self.y.lc(1)))
        return MontgomeryPoint(Num.from_allocated(xprime), Num.from_allocated(yprime))


def pedersen_hash(cs, personalization_bits, bits):
    allbits = [Boolean.constant(b) for b in personalization_bits] + list(bits)
    assert len(personalization_bits) == 6
    generators = pedersen_circuit_generators()
    false = Boolean.constant(False)
    edwards_result = None
    pos = 0
    seg = 0
    while pos < len(allbits):
        segment_result = None
        windows = generators[seg]
        w = 0
        while pos < len(allbits):
            a = allbits[pos]
            b = allbits[pos + 1] if pos + 1 < len(allbits) else false
            c = allbits[pos + 2] if pos + 2 < len(allbits) else false
            pos += 3
            x, y = lookup3_xy_with_conditional_negation(cs, [a, b, c], windows[w])
            tmp = MontgomeryPoint(x, y)
            segment_result = tmp if segment_result is None else tmp.add(cs, segment_result)
            w += 1
            if w == len(windows):
                break
        seg_edwards = segment_result.into_edwards(cs)
        edwards_result = seg_edwards if edwards_result is None else seg_edwards.add(cs, edwards_result)
        seg += 1
    return edwards_result


# ----------------------------------------------------------------------------
# circuits
# ----------------------------------------------------------------------------
def expose_value_commitment(cs, asset_generator, value, randomness):
    """sapling.rs:71-137.  asset_generator: affine Jubjub point; value: u64;
    randomness: jubjub::Fr as an integer.  Returns (asset_generator_bits, value_bits)."""
    ag = EdwardsPoint.witness(cs, asset_generator)
    asset_generator_bits = ag.repr(cs)
    ag = ag.double(cs)
    ag = ag.double(cs)
    ag = ag.double(cs)
    ag.u.assert_nonzero(cs)
    value_bits = u64_into_boolean_vec_le(cs, value)
    val = ag.mul(cs, value_bits)
    rcv_bits = field_into_boolean_vec_le(cs, randomness, JUBJUB_FR_BITS)
    rcv = fixed_base_multiplication(cs, VALUE_COMMITMENT_RANDOMNESS_GENERATOR, rcv_bits)
    cv = val.add(cs, rcv)
    cv.inputize(cs)
    return asset_generator_bits, value_bits


def convert_circuit(cs, asset_generator, value, randomness, auth_path, anchor):
    """Convert::synthesize (convert.rs:29-128).  auth_path: list of
    (sibling scalar, is_right bool)."""
    asset_generator_bits, value_bits = expose_value_commitment(cs, asset_generator, value, randomness)
    value_num = Num()
    coeff = 1
    for bit in value_bits:
        value_num = value_num.add_bool_with_coeff(bit, coeff)
        coeff = coeff * 2 % R
    assert len(asset_generator_bits) == 256
    cm = pedersen_hash(cs, NOTE_COMMITMENT_PERSONALIZATION, asset_generator_bits)
    cur = cm.u
    for i, (sibling, is_right) in enumerate(auth_path):
        cur_is_right = Boolean.from_bit(AllocatedBit.alloc(cs, is_right))
        path_element = AllocatedNum.alloc(cs, sibling)
        ul, ur = AllocatedNum.conditionally_reverse(cs, cur, path_element, cur_is_right)
        preimage = ul.to_bits_le(cs) + ur.to_bits_le(cs)
        cur = pedersen_hash(cs, merkle_personalization(i), preimage).u
    rt = AllocatedNum.alloc(cs, anchor)
    cs.enforce([(cur.var, 1), (rt.var, R - 1)], value_num.lc(1), [])
    rt.inputize(cs)
    return cur


def convert_native_anchor(asset_generator, auth_path):
    """The root the circuit computes, natively (convert.rs test :170-199)."""
    def bits_le(x, n):
        return [bool((x >> i) & 1) for i in range(n)]
    u, v = asset_generator
    repr_bits = bits_le(v, 255) + [bool(u & 1)]
    cur = pedersen_hash_native(NOTE_COMMITMENT_PERSONALIZATION, repr_bits)[0]
    for i, (sibling, is_right) in enumerate(auth_path):
        lhs, rhs = (sibling, cur) if is_right else (cur, sibling)
        cur = pedersen_hash_native(merkle_personalization(i), bits_le(lhs, 255) + bits_le(rhs, 255))[0]
    return cur


# ----------------------------------------------------------------------------
# Spend and Output (masp_proofs/src/circuit/sapling.rs:139-417, 419-596)
# ----------------------------------------------------------------------------
from .r1cs_gadgets import blake2s, pack_into_inputs  # noqa: E402

CRH_IVK_PERSONALIZATION = b"MASP_ivk"   # masp_primitives/src/constants.rs:17
PRF_NF_PERSONALIZATION = b"MASP__nf"    # :20
VALUE_COMMITMENT_GENERATOR_PERSONALIZATION = b"MASP__v_"  # :36
JUBJUB_FR_CAPACITY = 251
SCALAR_BITS = 255


def assert_not_small_order(cs, point):
    t = point.double(cs)
    t = t.double(cs)
    t = t.double(cs)
    t.u.assert_nonzero(cs)


def _merkle_and_anchor(cs, cur, auth_path, anchor, value_num):
    position_bits = []
    for i, (sibling, is_right) in enumerate(auth_path):
        cur_is_right = Boolean.from_bit(AllocatedBit.alloc(cs, is_right))
        position_bits.append(cur_is_right)
        path_element = AllocatedNum.alloc(cs, sibling)
        ul, ur = AllocatedNum.conditionally_reverse(cs, cur, path_element, cur_is_right)
        preimage = ul.to_bits_le(cs) + ur.to_bits_le(cs)
        cur = pedersen_hash(cs, merkle_personalization(i), preimage).u
    rt = AllocatedNum.alloc(cs, anchor)
    cs.enforce([(cur.var, 1), (rt.var, R - 1)], value_num.lc(1), [])
    rt.inputize(cs)
    return position_bits


def spend_circuit(cs, ak, nsk, g_d, asset_generator, value, rcv, rcm, ar, auth_path, anchor):
    """Spend::synthesize.  ak, g_d, asset_generator: affine Jubjub points;
    nsk, rcv, rcm, ar: jubjub::Fr integers.  Returns the nullifier bits' values."""
    ak_p = EdwardsPoint.witness(cs, ak)
    assert_not_small_order(cs, ak_p)
    ar_bits = field_into_boolean_vec_le(cs, ar, JUBJUB_FR_BITS)
    ar_p = fixed_base_multiplication(cs, SPENDING_KEY_GENERATOR, ar_bits)
    rk = ak_p.add(cs, ar_p)
    rk.inputize(cs)
    nsk_bits = field_into_boolean_vec_le(cs, nsk, JUBJUB_FR_BITS)
    nk = fixed_base_multiplication(cs, PROOF_GENERATION_KEY_GENERATOR, nsk_bits)
    ivk_preimage = ak_p.repr(cs)
    repr_nk = nk.repr(cs)
    ivk_preimage = ivk_preimage + repr_nk
    nf_preimage = list(repr_nk)
    assert len(ivk_preimage) == 512 and len(nf_preimage) == 256
    ivk = blake2s(cs, ivk_preimage, CRH_IVK_PERSONALIZATION)[:JUBJUB_FR_CAPACITY]
    g_d_p = EdwardsPoint.witness(cs, g_d)
    assert_not_small_order(cs, g_d_p)
    pk_d = g_d_p.mul(cs, ivk)
    asset_generator_bits, value_bits = expose_value_commitment(cs, asset_generator, value, rcv)
    value_num = Num()
    coeff = 1
    for bit in value_bits:
        value_num = value_num.add_bool_with_coeff(bit, coeff)
        coeff = coeff * 2 % R
    note_contents = asset_generator_bits + value_bits + g_d_p.repr(cs) + pk_d.repr(cs)
    assert len(note_contents) == 256 + 64 + 256 + 256
    cm = pedersen_hash(cs, NOTE_COMMITMENT_PERSONALIZATION, note_contents)
    rcm_bits = field_into_boolean_vec_le(cs, rcm, JUBJUB_FR_BITS)
    rcm_p = fixed_base_multiplication(cs, NOTE_COMMITMENT_RANDOMNESS_GENERATOR, rcm_bits)
    cm = cm.add(cs, rcm_p)
    position_bits = _merkle_and_anchor(cs, cm.u, auth_path, anchor, value_num)
    position = fixed_base_multiplication(cs, NULLIFIER_POSITION_GENERATOR, position_bits)
    rho = cm.add(cs, position)
    nf_preimage = nf_preimage + rho.repr(cs)
    assert len(nf_preimage) == 512
    nf = blake2s(cs, nf_preimage, PRF_NF_PERSONALIZATION)
    pack_into_inputs(cs, nf)
    return [b.value for b in nf], (rk.u.value, rk.v.value)


def output_circuit(cs, asset_identifier_bits, asset_generator, value, rcv, g_d, pk_d, rcm, esk):
    """Output::synthesize.  asset_identifier_bits: 256 booleans whose BLAKE2s image
    (personalised MASP__v_) is the encoding of asset_generator."""
    assert len(asset_identifier_bits) == 256
    preimage = [Boolean.from_bit(AllocatedBit.alloc(cs, b)) for b in asset_identifier_bits]
    image = blake2s(cs, preimage, VALUE_COMMITMENT_GENERATOR_PERSONALIZATION)
    asset_generator_bits, value_bits = expose_value_commitment(cs, asset_generator, value, rcv)
    assert len(asset_generator_bits) == 256 and len(image) == 256
    for a, b in zip(asset_generator_bits, image):
        Boolean.enforce_equal(cs, a, b)
    note_contents = asset_generator_bits + value_bits
    g_d_p = EdwardsPoint.witness(cs, g_d)
    assert_not_small_order(cs, g_d_p)
    note_contents = note_contents + g_d_p.repr(cs)
    esk_bits = field_into_boolean_vec_le(cs, esk, JUBJUB_FR_BITS)
    epk = g_d_p.mul(cs, esk_bits)
    epk.inputize(cs)
    v_contents = field_into_boolean_vec_le(cs, pk_d[1], SCALAR_BITS)
    sign_bit = Boolean.from_bit(AllocatedBit.alloc(cs, pk_d[0] & 1))
    note_contents = note_contents + v_contents + [sign_bit]
    assert len(note_contents) == 256 + 64 + 256 + 256
    cm = pedersen_hash(cs, NOTE_COMMITMENT_PERSONALIZATION, note_contents)
    rcm_bits = field_into_boolean_vec_le(cs, rcm, JUBJUB_FR_BITS)
    rcm_p = fixed_base_multiplication(cs, NOTE_COMMITMENT_RANDOMNESS_GENERATOR, rcm_bits)
    cm = cm.add(cs, rcm_p)
    cm.u.inputize(cs)
    return cm.u.value


# -- native helpers for witnesses ---------------------------------------------
def fr_sqrt(a):
    """Tonelli-Shanks in the BLS12-381 scalar field (r - 1 = 2^32 * odd)."""
    a %= R
    if a == 0:
        return 0
    if pow(a, (R - 1) // 2, R) != 1:
        return None
    q, s = R - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 7  # a generator, hence a non-residue
    m, c, t, r_ = s, pow(z, q, R), pow(a, q, R), pow(a, (q + 1) // 2, R)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % R
            i += 1
        b = pow(c, 1 << (m - i - 1), R)
        m, c = i, b * b % R
        t, r_ = t * c % R, r_ * b % R
    return r_


def jj_decompress(b32):
    """jubjub AffinePoint::from_bytes: v little-endian in 255 bits, top bit = sign of u."""
    enc = int.from_bytes(b32, "little")
    sign = enc >> 255
    v = enc & ((1 << 255) - 1)
    if v >= R:
        return None
    u2 = (v * v - 1) * inv(1 + EDWARDS_D * v * v) % R
    u = fr_sqrt(u2)
    if u is None:
        return None
    if (u & 1) != sign:
        u = (-u) % R
    if u == 0 and sign:
        return None
    return (u, v)


def bytes_to_bits_le(b):
    return [bool((byte >> i) & 1) for byte in b for i in range(8)]


def find_asset(seed=0):
    """An asset identifier whose BLAKE2s image decodes to a prime-order-free
    Jubjub point, like AssetType::new's nonce search
    (masp_primitives/src/asset_type.rs:20-153) in spirit: identifier bytes
    are drawn until the image is a valid, non-small-order point."""
    import hashlib
    k = seed
    while True:
        ident = hashlib.blake2s(b"masp_b200 test asset %d" % k, digest_size=32).digest()
        image = hashlib.blake2s(ident, digest_size=32, person=VALUE_COMMITMENT_GENERATOR_PERSONALIZATION).digest()
        p = jj_decompress(image)
        if p is not None and jj_mul(p, 8)[0] != 0:
            return ident, p
        k += 1


def jj_repr_bits(p):
    """EdwardsPoint::repr as booleans: 255 bits of v (LSB first) then u's parity."""
    u, v = p
    return [bool((v >> i) & 1) for i in range(255)] + [bool(u & 1)]


def bits_to_bytes_le(bits):
    out = bytearray(len(bits) // 8)
    for i, b in enumerate(bits):
        if b:
            out[i // 8] |= 1 << (i % 8)
    return bytes(out)


def spend_native(ak, nsk, g_d, asset_generator, value, rcv, rcm, ar, auth_path):
    """What the Spend circuit must expose, computed natively (the checks of
    masp_proofs/src/circuit/sapling.rs:743-759): rk, cv, anchor, nf."""
    import hashlib
    nk = jj_mul(PROOF_GENERATION_KEY_GENERATOR, nsk)
    rk = jj_add(ak, jj_mul(SPENDING_KEY_GENERATOR, ar))
    ivk_bytes = hashlib.blake2s(bits_to_bytes_le(jj_repr_bits(ak) + jj_repr_bits(nk)), digest_size=32,
                                person=CRH_IVK_PERSONALIZATION).digest()
    ivk = int.from_bytes(ivk_bytes, "little") & ((1 << JUBJUB_FR_CAPACITY) - 1)
    pk_d = jj_mul(g_d, ivk)
    cv = jj_add(jj_mul(jj_mul(asset_generator, 8), value), jj_mul(VALUE_COMMITMENT_RANDOMNESS_GENERATOR, rcv))
    value_bits = [bool((value >> i) & 1) for i in range(64)]
    contents = jj_repr_bits(asset_generator) + value_bits + jj_repr_bits(g_d) + jj_repr_bits(pk_d)
    cm = jj_add(pedersen_hash_native(NOTE_COMMITMENT_PERSONALIZATION, contents),
                jj_mul(NOTE_COMMITMENT_RANDOMNESS_GENERATOR, rcm))
    cur = cm[0]
    position = 0
    for i, (sibling, is_right) in enumerate(auth_path):
        position |= int(is_right) << i
        lhs, rhs = (sibling, cur) if is_right else (cur, sibling)
        bits = [bool((lhs >> k) & 1) for k in range(255)] + [bool((rhs >> k) & 1) for k in range(255)]
        cur = pedersen_hash_native(merkle_personalization(i), bits)[0]
    rho = jj_add(cm, jj_mul(NULLIFIER_POSITION_GENERATOR, position))
    nf = hashlib.blake2s(bits_to_bytes_le(jj_repr_bits(nk) + jj_repr_bits(rho)), digest_size=32,
                         person=PRF_NF_PERSONALIZATION).digest()
    return {"rk": rk, "cv": cv, "anchor": cur, "nf": nf}


def multipack_inputs(bits):
    """multipack::compute_multipacking: 254 bits per scalar, LSB first."""
    return [sum((1 << i) for i, b in enumerate(bits[k:k + 254]) if b) % R for k in range(0, len(bits), 254)]


def output_native(asset_generator, value, rcv, g_d, pk_d, rcm, esk):
    """cv, epk, cmu as the Output circuit exposes them (sapling.rs:1045-1065)."""
    cv = jj_add(jj_mul(jj_mul(asset_generator, 8), value), jj_mul(VALUE_COMMITMENT_RANDOMNESS_GENERATOR, rcv))
    epk = jj_mul(g_d, esk)
    value_bits = [bool((value >> i) & 1) for i in range(64)]
    contents = jj_repr_bits(asset_generator) + value_bits + jj_repr_bits(g_d) + jj_repr_bits(pk_d)
    cm = jj_add(pedersen_hash_native(NOTE_COMMITMENT_PERSONALIZATION, contents),
                jj_mul(NOTE_COMMITMENT_RANDOMNESS_GENERATOR, rcm))
    return {"cv": cv, "epk": epk, "cmu": cm[0]}
