"""Synthetic shapes, witnesses and key scalars for the MASP circuits.

The reference benches build witnesses for the real circuits
(masp_proofs/benches/sapling.rs:38-69, benches/convert.rs:31-54) and random
parameters via generate_random_parameters (sapling.rs:24-36).  Witness
synthesis is outside this path (NEXT-1), so the proving
path is driven by synthetic evaluation vectors of exactly the reference
circuits' shapes (SURVEY.md §8 shape and scalar make-up tables).

Everything here is a pure function of (seed, stream, index) through a
SplitMix64-style counter PRNG; the same derivation is implemented on the
device in csrc/synth.cuh (mb200_params_synthesize) so keys can be compared
byte for byte.  Master seed: the first 8 bytes of the reference bench seed
(masp_proofs/benches/sapling.rs:19-22).
"""
from dataclasses import dataclass

import numpy as np

MASTER_SEED = 0x5962BE3D763D318D
R_INT = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_R_LIMBS = np.array([(R_INT >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_STREAM_MUL = np.uint64(0xD1342543DE82EF95)

# PRNG stream ids
STREAM_VK = 10       # alpha, beta, gamma, delta at index 0..3
STREAM_IC = 11
STREAM_H = 12
STREAM_L = 13
STREAM_A = 14
STREAM_B = 15
STREAM_DENSITY = 5
STREAM_WIT_A = 100   # + 8 * proof index ...
STREAM_MSM_BASE = 7
STREAM_MSM_SCALAR_U = 8
STREAM_MSM_SCALAR_W = 9


def mix64(z):
    """SplitMix64 output function on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(z, dtype=np.uint64) + _GOLD)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def stream_key(seed, stream):
    with np.errstate(over="ignore"):
        return mix64(np.uint64(seed) + np.uint64(stream) * _STREAM_MUL)


def raw_u64(seed, stream, idx):
    """One uint64 per counter value idx (array)."""
    with np.errstate(over="ignore"):
        return mix64(stream_key(seed, stream) + np.asarray(idx, dtype=np.uint64) * _GOLD)


def _ge_r(l):
    """l: (n,4) uint64 little-endian limbs; returns l >= r."""
    ge = np.ones(l.shape[0], dtype=bool)
    decided = np.zeros(l.shape[0], dtype=bool)
    for i in (3, 2, 1, 0):
        gt = l[:, i] > _R_LIMBS[i]
        lt = l[:, i] < _R_LIMBS[i]
        ge = np.where(~decided & lt, False, ge)
        decided |= gt | lt
    return ge


def _sub_r_where(l, mask):
    out = l.copy()
    borrow = np.zeros(l.shape[0], dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(4):
            a = l[:, i]
            t = a - _R_LIMBS[i]
            b1 = (a < _R_LIMBS[i]).astype(np.uint64)
            t2 = t - borrow
            b2 = (t < borrow).astype(np.uint64)
            out[:, i] = np.where(mask, t2, a)
            borrow = b1 | b2
    return out


def fr_uniform(seed, stream, n, start=0):
    """n scalars in [0, r) as an (n, 4) uint64 array of little-endian limbs:
    255 random bits, minus r if that is >= r."""
    idx = (np.arange(start, start + n, dtype=np.uint64)[:, None] * np.uint64(4)
           + np.arange(4, dtype=np.uint64)[None, :])
    l = raw_u64(seed, stream, idx)
    l[:, 3] &= np.uint64(0x7FFFFFFFFFFFFFFF)
    return _sub_r_where(l, _ge_r(l))


def fr_bits(seed, stream, n, start=0):
    """n scalars uniform in {0, 1} as (n, 4) uint64 limbs."""
    out = np.zeros((n, 4), dtype=np.uint64)
    out[:, 0] = raw_u64(seed, stream ^ 0x8000, np.arange(start, start + n, dtype=np.uint64)) >> np.uint64(63)
    return out


def limbs_to_bytes(l):
    return np.ascontiguousarray(l.astype("<u8")).tobytes()


def limbs_to_ints(l):
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in l]


def ints_to_bytes(vals):
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def pack_bits(bits):
    """Boolean array -> bitmap bytes, LSB-first (bit i of byte i//8)."""
    return np.packbits(np.asarray(bits, dtype=np.uint8), bitorder="little").tobytes()


# class codes of an aux variable: bit0 = in A density, bit1 = in B density, bit2 = boolean
@dataclass(frozen=True)
class Shape:
    """Exact circuit shape: counts of aux variables by
    (boolean?, dense in A?, dense in B?)."""
    name: str
    n_constraints: int
    n_inputs: int           # including ONE
    bool_ab: int
    bool_a: int
    bool_b: int
    bool_none: int
    full_ab: int
    full_a: int
    full_b: int
    full_none: int

    @property
    def n_aux(self):
        return (self.bool_ab + self.bool_a + self.bool_b + self.bool_none
                + self.full_ab + self.full_a + self.full_b + self.full_none)

    @property
    def rows(self):
        return self.n_constraints + self.n_inputs

    @property
    def log_m(self):
        e = 0
        while (1 << e) < self.rows:
            e += 1
        return e

    @property
    def m(self):
        return 1 << self.log_m

    @property
    def h_len(self):
        return self.m - 1

    @property
    def a_dense(self):
        return self.bool_ab + self.bool_a + self.full_ab + self.full_a

    @property
    def b_dense(self):
        return self.bool_ab + self.bool_b + self.full_ab + self.full_b

    @property
    def a_len(self):
        return self.n_inputs + self.a_dense

    @property
    def b_len(self):
        return 1 + self.b_dense       # b_input_density = {ONE}

    @property
    def n_bool(self):
        return self.bool_ab + self.bool_a + self.bool_b + self.bool_none

    def key_bytes(self):
        return 96 * (self.h_len + self.n_aux + self.a_len + self.b_len) + 192 * self.b_len

    def scalar_bytes(self):
        return 32 * (self.h_len + self.n_aux + (self.a_len - self.n_inputs) + 2 * (self.b_len - 1))

    def algorithmic_bytes(self):
        """key + scalars + 7 NTTs (64 m each) + pointwise (128 m)."""
        return self.key_bytes() + self.scalar_bytes() + 7 * 64 * self.m + 128 * self.m

    def params_file_bytes(self):
        vk = 3 * 96 + 3 * 192 + 4 + 96 * self.n_inputs
        return vk + 4 * 5 + 96 * (self.h_len + self.n_aux + self.a_len + self.b_len) + 192 * self.b_len

    def scaled(self, name, n_constraints, frac):
        """A smaller shape with the same class proportions (tests)."""
        f = lambda x: max(1, int(round(x * frac))) if x else 0
        return Shape(name, n_constraints, self.n_inputs, f(self.bool_ab), f(self.bool_a), f(self.bool_b),
                     f(self.bool_none), f(self.full_ab), f(self.full_a), f(self.full_b), f(self.full_none))

    # ---- deterministic layout of the classes over the aux vector ----------
    def aux_classes(self, seed=MASTER_SEED):
        counts = [(0b111, self.bool_ab), (0b101, self.bool_a), (0b110, self.bool_b), (0b100, self.bool_none),
                  (0b011, self.full_ab), (0b001, self.full_a), (0b010, self.full_b), (0b000, self.full_none)]
        cls = np.concatenate([np.full(n, code, dtype=np.uint8) for code, n in counts])
        order = np.argsort(raw_u64(seed, STREAM_DENSITY, np.arange(self.n_aux, dtype=np.uint64)), kind="stable")
        out = np.empty_like(cls)
        out[order] = cls
        return out

    def densities(self, seed=MASTER_SEED):
        """(a_aux_density, b_input_density, b_aux_density) bitmaps, LSB-first."""
        cls = self.aux_classes(seed)
        b_in = np.zeros(self.n_inputs, dtype=np.uint8)
        b_in[0] = 1
        return pack_bits(cls & 1), pack_bits(b_in), pack_bits((cls >> 1) & 1)


# SURVEY.md §8 "Scalar make-up" table, decomposed (boolean variables that are
# in neither query are taken as zero; the totals reproduce every column).
SPEND = Shape("spend", 100637, 8, 51483, 18250, 318, 0, 9599, 7591, 769, 12487)
OUTPUT = Shape("output", 31205, 6, 17414, 6581, 290, 0, 3440, 1318, 239, 1614)
CONVERT = Shape("convert", 47358, 4, 17264, 5787, 0, 0, 6576, 6116, 248, 11331)
SHAPES = {"spend": SPEND, "output": OUTPUT, "convert": CONVERT}


def tiny_shape(name="tiny", n_constraints=200, n_inputs=4):
    """A few hundred rows with every class populated (CPU-side tests)."""
    return Shape(name, n_constraints, n_inputs, 40, 15, 3, 2, 30, 20, 6, 25)


def micro_shape(name="micro", n_constraints=56, n_inputs=3):
    """The smallest shape with every class populated (host-emulated device tests)."""
    return Shape(name, n_constraints, n_inputs, 9, 4, 2, 1, 7, 5, 2, 6)


def witness(shape, proof_index, fr_mul, seed=MASTER_SEED):
    """Synthetic evaluation vectors for one proof:
    a_i, b_i uniform, c_i = a_i * b_i for i < rows; aux boolean/full-width by
    class; inputs full-width with inputs[0] = 1; r, s uniform.

    fr_mul(a_bytes, b_bytes, n) -> bytes computes the elementwise product
    (the device kernel behind mb200_fr_mul in production; tests may pass the
    oracle's).  Returns a dict of byte strings."""
    base = STREAM_WIT_A + 8 * proof_index
    rows = shape.rows
    a = fr_uniform(seed, base + 0, rows)
    b = fr_uniform(seed, base + 1, rows)
    # the prover's extra rows: input_i * 0 = 0  ->  b = c = 0 there, a = input_i
    inputs = fr_uniform(seed, base + 2, shape.n_inputs)
    inputs[0] = np.array([1, 0, 0, 0], dtype=np.uint64)
    a[shape.n_constraints:] = inputs
    b[shape.n_constraints:] = 0
    a_b, b_b = limbs_to_bytes(a), limbs_to_bytes(b)
    c_b = fr_mul(a_b, b_b, rows)
    cls = shape.aux_classes(seed)
    aux = fr_uniform(seed, base + 3, shape.n_aux)
    bits = fr_bits(seed, base + 4, shape.n_aux)
    is_bool = (cls & 4) != 0
    aux[is_bool] = bits[is_bool]
    rs = fr_uniform(seed, base + 5, 2)
    return {"a": a_b, "b": b_b, "c": c_b, "inputs": limbs_to_bytes(inputs), "aux": limbs_to_bytes(aux),
            "r": limbs_to_bytes(rs[0:1]), "s": limbs_to_bytes(rs[1:2])}


def key_logs(shape, seed=MASTER_SEED):
    """Discrete logs of the synthetic key of this shape: dict of (n,4) limb arrays."""
    return {"vk": fr_uniform(seed, STREAM_VK, 4), "ic": fr_uniform(seed, STREAM_IC, shape.n_inputs),
            "h": fr_uniform(seed, STREAM_H, shape.h_len), "l": fr_uniform(seed, STREAM_L, shape.n_aux),
            "a": fr_uniform(seed, STREAM_A, shape.a_len), "b": fr_uniform(seed, STREAM_B, shape.b_len)}


def msm_scalars(n, kind="U", seed=MASTER_SEED):
    """BASELINE config 3 scalar sets: 'U' uniform; 'W' witness-like (75% in {0,1})."""
    if kind == "U":
        return fr_uniform(seed, STREAM_MSM_SCALAR_U, n)
    full = fr_uniform(seed, STREAM_MSM_SCALAR_W, n)
    bits = fr_bits(seed, STREAM_MSM_SCALAR_W, n)
    sel = (raw_u64(seed, STREAM_MSM_SCALAR_W ^ 0x4000, np.arange(n, dtype=np.uint64)) & np.uint64(3)) != 0
    full[sel] = bits[sel]
    return full
