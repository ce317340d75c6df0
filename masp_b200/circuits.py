"""The MASP circuits as the reference hands them to bellman, over the C ABI.

Mirrors the `Circuit` instances of the reference (field names as there):

  * circuit::sapling::Spend    masp_proofs/src/circuit/sapling.rs:40-68   (synthesize :139-417)
  * circuit::sapling::Output   masp_proofs/src/circuit/sapling.rs:419-450 (synthesize :452-596)
  * circuit::convert::Convert  masp_proofs/src/circuit/convert.rs:16-27   (synthesize :29-128)

A circuit is recorded once (`Circuit(kind)`: matrices, densities, structural
hash) and bound to its key (`Parameters.bind_circuit`); after that a proof
needs only the witness: `Circuit.synthesize` runs the witness half of
`Circuit::synthesize` on the host cores and the device evaluates the rows.

Jubjub points are affine `(u, v)` integer pairs, Jubjub scalars integers.
"""
import ctypes
from dataclasses import dataclass
from typing import List, Tuple

from . import _lib
from ._lib import check

SPEND, OUTPUT, CONVERT = 0, 1, 2
TREE_DEPTH = 32  # masp_primitives/src/sapling.rs SAPLING_COMMITMENT_TREE_DEPTH

# reference pins: (constraints, inputs, TestConstraintSystem::hash)
PINS = {
    SPEND: (100637, 8, "34e4a634c80e4e4c6250e63b7855532e60b36d1371d4d7b1163218b69f09eb3d"),    # sapling.rs:730-741
    OUTPUT: (31205, 6, "93e445d7858e98c7138558df341f020aedfe75893535025587d64731e244276a"),    # sapling.rs:1024-1045
    CONVERT: (47358, 4, "f74b47ef6e59081548f81f5806bd15b1f4a65d2e57681e6db2b8db7eef2ff814"),   # convert.rs:218-224
}

Point = Tuple[int, int]

# fixed generators, affine (u, v): masp_primitives/src/constants.rs:50-251
PROOF_GENERATION_KEY_GENERATOR = (0x4CAAEACAAF28ED4B4BA1F065E719FD031E24F83267F15ABD5F3C723AA2531B66,
                                  0x00930D67D6906365C654DFDD36004DE936B49C71A2AF0708FE6F96BEC575BFF8)
SPENDING_KEY_GENERATOR = (0x5B389522A9E81532F831C2B19FEC602639F5B03380AF6020EC75293D81248452,
                          0x0CBC5F9F1E52E0AB75DEFECFF1F49EF22012D031F624FD5214B62623A186B4B1)
VALUE_COMMITMENT_RANDOMNESS_GENERATOR = (0x1C6DA0CE9A5E5FDBCFA86026B8D99BE991CC3E3835675450DD93D364CB8CEC7E,
                                         0x555F11F9B720D50BBC900CD4B8AE1150F94C2DAA360302FE28E5FCE99CE692D0)
JUBJUB_ORDER = 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7


def _f(x):
    return int(x).to_bytes(32, "little")


def _pt(p):
    return _f(p[0]) + _f(p[1])


def _path(path):
    return b"".join(_f(sib) + _f(1 if right else 0) for sib, right in path)


@dataclass
class ValueCommitmentOpening:  # masp_primitives ValueCommitment: asset_generator, value, randomness
    asset_generator: Point
    value: int
    randomness: int


@dataclass
class Spend:
    value_commitment: ValueCommitmentOpening
    ak: Point                    # proof_generation_key.ak
    nsk: int                     # proof_generation_key.nsk
    g_d: Point                   # payment_address.g_d()
    commitment_randomness: int   # rcm
    ar: int
    auth_path: List[Tuple[int, bool]]
    anchor: int
    kind = SPEND

    def pack(self):
        vc = self.value_commitment
        return (_pt(self.ak) + _f(self.nsk) + _pt(self.g_d) + _pt(vc.asset_generator) + _f(vc.value) +
                _f(vc.randomness) + _f(self.commitment_randomness) + _f(self.ar) + _f(self.anchor) +
                _path(self.auth_path))


@dataclass
class Output:
    value_commitment: ValueCommitmentOpening
    asset_identifier: bytes      # 32 bytes whose BLAKE2s image (MASP__v_) encodes asset_generator
    g_d: Point
    pk_d: Point
    commitment_randomness: int
    esk: int
    kind = OUTPUT

    def pack(self):
        vc = self.value_commitment
        assert len(self.asset_identifier) == 32
        return (bytes(self.asset_identifier) + _pt(vc.asset_generator) + _f(vc.value) + _f(vc.randomness) +
                _pt(self.g_d) + _pt(self.pk_d) + _f(self.commitment_randomness) + _f(self.esk))


@dataclass
class Convert:
    value_commitment: ValueCommitmentOpening
    auth_path: List[Tuple[int, bool]]
    anchor: int
    kind = CONVERT

    def pack(self):
        vc = self.value_commitment
        return _pt(vc.asset_generator) + _f(vc.value) + _f(vc.randomness) + _f(self.anchor) + _path(self.auth_path)


def pedersen_hash(personalization_bits, bits):
    """masp_primitives::sapling::pedersen_hash as the circuits compute it: six personalization
    bits, then the message bits; returns the affine (u, v) of the hash point."""
    allbits = bytes(1 if b else 0 for b in list(personalization_bits) + list(bits))
    u, v = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
    check(_lib.lib().mb200_pedersen_hash(allbits, len(allbits), u, v))
    return int.from_bytes(u.raw, "little"), int.from_bytes(v.raw, "little")


class Circuit:
    """One recorded circuit (mb200_circuit).  Host only: needs no device."""

    def __init__(self, kind, depth=TREE_DEPTH):
        self.kind = kind
        self.depth = 0 if kind == OUTPUT else depth
        self._h = ctypes.c_void_p()
        check(_lib.lib().mb200_circuit_new(kind, self.depth, ctypes.byref(self._h)))
        info = (ctypes.c_uint64 * 10)()
        check(_lib.lib().mb200_circuit_info(self._h, info))
        (self.n_inputs, self.n_aux, self.n_constraints, self.nnz_a, self.nnz_b, self.nnz_c, self.witness_bytes,
         self.a_dense, self.b_input_dense, self.b_dense) = [int(x) for x in info]

    @property
    def rows(self):
        return self.n_constraints + self.n_inputs

    def hash(self):
        buf = ctypes.create_string_buffer(65)
        check(_lib.lib().mb200_circuit_hash(self._h, buf))
        return buf.value.decode()

    def densities(self):
        """(a_aux_density, b_input_density, b_aux_density) bitmaps for Parameters.read."""
        a = ctypes.create_string_buffer((self.n_aux + 7) // 8)
        bi = ctypes.create_string_buffer((self.n_inputs + 7) // 8)
        ba = ctypes.create_string_buffer((self.n_aux + 7) // 8)
        check(_lib.lib().mb200_circuit_densities(self._h, a, bi, ba))
        return a.raw, bi.raw, ba.raw

    def matrix(self, which):
        """CSR of A (0), B (1) or C (2): (rowptr list, col list, coefficient ints)."""
        nnz = (self.nnz_a, self.nnz_b, self.nnz_c)[which]
        rp = (ctypes.c_uint32 * (self.n_constraints + 1))()
        col = (ctypes.c_uint32 * max(nnz, 1))()
        coef = ctypes.create_string_buffer(32 * max(nnz, 1))
        check(_lib.lib().mb200_circuit_matrix(self._h, which, rp, col, coef))
        vals = [int.from_bytes(coef.raw[32 * i:32 * i + 32], "little") for i in range(nnz)]
        return list(rp), list(col)[:nnz], vals

    def synthesize(self, instances, threads=0, numpy=False, out=None):
        """Witness half of Circuit::synthesize for a list of instances (or packed
        witness bytes): returns (inputs, aux), n x n_inputs / n x n_aux scalars, as
        bytes -- or, with numpy=True, as uint8 arrays written in place (no copies:
        what a caller feeding the prover in a loop wants)."""
        import numpy as np
        ws = [w if isinstance(w, (bytes, bytearray)) else w.pack() for w in instances]
        n = len(ws)
        for w in ws:
            if len(w) != self.witness_bytes:
                raise ValueError("witness has %d bytes, this circuit takes %d" % (len(w), self.witness_bytes))
        if out is not None:  # caller-owned uint8 arrays, e.g. views of pinned memory
            inp, aux = out
            if inp.nbytes < n * self.n_inputs * 32 or aux.nbytes < n * self.n_aux * 32:
                raise ValueError("output buffers too small")
            numpy = True
        else:
            inp = np.empty(max(1, n * self.n_inputs * 32), dtype=np.uint8)
            aux = np.empty(max(1, n * self.n_aux * 32), dtype=np.uint8)
        check(_lib.lib().mb200_circuit_synthesize(self._h, n, b"".join(ws), inp.ctypes.data_as(ctypes.c_char_p),
                                                  aux.ctypes.data_as(ctypes.c_char_p), threads))
        inp, aux = inp[:n * self.n_inputs * 32], aux[:n * self.n_aux * 32]
        return (inp, aux) if numpy else (inp.tobytes(), aux.tobytes())

    def root(self, instance):
        """The Merkle root this Spend / Convert witness leads to (its anchor field is ignored)."""
        w = instance if isinstance(instance, (bytes, bytearray)) else instance.pack()
        out = ctypes.create_string_buffer(32)
        check(_lib.lib().mb200_circuit_root(self._h, bytes(w), out))
        return int.from_bytes(out.raw, "little")

    def rows_on_device(self, inputs, aux, n):
        """(a, b, c) row evaluations of n witnesses, computed by the r1cs_eval kernel:
        what bellman's ProvingAssignment holds after synthesize (incl. the input rows)."""
        sz = max(1, n * self.rows * 32)
        a, b, c = (ctypes.create_string_buffer(sz) for _ in range(3))
        check(_lib.lib().mb200_circuit_rows(self._h, n, inputs, aux, a, b, c))
        k = n * self.rows * 32
        return a.raw[:k], b.raw[:k], c.raw[:k]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().mb200_circuit_free(h)
            except Exception:
                pass
            self._h = None
