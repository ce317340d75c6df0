"""The callers of the proving path: SaplingProvingContext and the TxProver trait.

Mirrors, with the reference's names, argument order and error behaviour:

  * SaplingProvingContext::{new, spend_proof, output_proof, convert_proof, binding_sig}
                                                   masp_proofs/src/sapling/prover.rs:26-327
  * trait TxProver as LocalTxProver implements it  masp_primitives/src/sapling/prover.rs:17-83,
                                                   masp_proofs/src/prover.rs:156-261
  * the key / note types those signatures take, restricted to what the prover touches:
    group_hash (sapling/group_hash.rs:15-43), AssetType (asset_type.rs:28-154),
    ProofGenerationKey / ViewingKey / Diversifier / PaymentAddress / Note / ValueCommitment
    (sapling.rs:196-225, 333-360, 453-480, 503-565, 796-863), AllowedConversion
    (convert.rs:23-120), RedJubjub sign / verify (sapling/redjubjub.rs:133-260)
  * SaplingVerificationContext / BatchValidator      masp_proofs/src/sapling/verifier.rs:19-215,
    (the verifier-side callers of verify_proof)      verifier/single.rs, verifier/batch.rs:60-243
  * write_v5_sapling / read_v5_sapling               masp_primitives/src/transaction.rs:612-720, 746-806
  * BatchingTxProver: SURVEY.md §8(f)-2 -- `*_proof` calls update the context and enqueue,
    every proof of the transaction is produced by one launch per circuit when the first
    result is needed (binding_sig at the latest).

What runs where.  The Groth16 work (witness generation, row evaluation, the
prover, verify_proof) is the library's: C++ host threads and CUDA kernels
behind the C ABI.  This module is the bookkeeping the reference keeps outside
create_random_proof -- `bsk`, `cv_sum`, the value commitment and `rk` returned
to the caller, and the public inputs for the self-check -- a few Jubjub
operations per description on Python integers.  The public inputs are computed
natively here (as the reference does at sapling/prover.rs:121-145, 256-263),
NOT taken from the circuit's witness, so the self-check stays an independent one.

Jubjub points are affine (u, v) integer pairs, scalars are integers.
"""
import hashlib
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from . import circuits as C
from . import prover as P

Q = P.R_ORDER                       # the BLS12-381 scalar field = Jubjub's base field
JUBJUB_ORDER = C.JUBJUB_ORDER       # prime subgroup order (jubjub::Fr)
EDWARDS_D = (-10240 * pow(10241, -1, Q)) % Q   # -u^2 + v^2 = 1 + d u^2 v^2
IDENTITY = (0, 1)
Point = Tuple[int, int]

# masp_primitives/src/constants.rs:12-46
GH_FIRST_BLOCK = b"096b36a5804bfacef1691e173c366a47ff5ba84a44f26ddd7e8d9f79d5b42df0"
CRH_IVK_PERSONALIZATION = b"MASP_ivk"
PRF_NF_PERSONALIZATION = b"MASP__nf"
PEDERSEN_HASH_GENERATORS_PERSONALIZATION = b"MASP__PH"
KEY_DIVERSIFICATION_PERSONALIZATION = b"MASP__gd"
SPENDING_KEY_GENERATOR_PERSONALIZATION = b"MASP__G_"
PROOF_GENERATION_KEY_BASE_GENERATOR_PERSONALIZATION = b"MASP__H_"
VALUE_COMMITMENT_GENERATOR_PERSONALIZATION = b"MASP__v_"
VALUE_COMMITMENT_RANDOMNESS_PERSONALIZATION = b"MASP__r_"
NULLIFIER_POSITION_IN_TREE_GENERATOR_PERSONALIZATION = b"MASP__J_"
ASSET_IDENTIFIER_PERSONALIZATION = b"MASP__t_"
PRF_EXPAND_PERSONALIZATION = b"MASP__ExpandSeed"   # keys.rs:5
REDJUBJUB_H_PERSONALIZATION = b"MASP__RedJubjubH"  # sapling/redjubjub.rs:37-39


class SaplingError(Exception):
    """The reference's `Err(())`: an invalid diversifier (sapling/prover.rs:84), a proof that does
    not pass verify_proof (:148, :266), a value balance that does not match (:306-310)."""


# ---------------------------------------------------------------------------
# Jubjub on integers (bookkeeping only: a handful of operations per description)
# ---------------------------------------------------------------------------
def _inv(x):
    return pow(x, -1, Q)


def jj_add(p: Point, q: Point) -> Point:
    (u1, v1), (u2, v2) = p, q
    t = EDWARDS_D * u1 * u2 % Q * v1 % Q * v2 % Q
    return ((u1 * v2 + v1 * u2) * _inv(1 + t) % Q, (v1 * v2 + u1 * u2) * _inv(1 - t) % Q)


def jj_neg(p: Point) -> Point:
    return ((-p[0]) % Q, p[1])


def jj_mul(p: Point, k: int) -> Point:
    # extended coordinates (U, V, Z, T) with a = -1; one inversion at the end
    def ext_add(a, b):
        U1, V1, Z1, T1 = a
        U2, V2, Z2, T2 = b
        A = (V1 - U1) * (V2 - U2) % Q
        B = (V1 + U1) * (V2 + U2) % Q
        Cc = 2 * EDWARDS_D * T1 % Q * T2 % Q
        D = 2 * Z1 * Z2 % Q
        E, F, G, H = B - A, D - Cc, D + Cc, B + A
        return (E * F % Q, G * H % Q, F * G % Q, E * H % Q)

    acc = (0, 1, 1, 0)
    base = (p[0], p[1], 1, p[0] * p[1] % Q)
    k = int(k)
    if k < 0:
        raise ValueError("negative scalar")
    while k:
        if k & 1:
            acc = ext_add(acc, base)
        base = ext_add(base, base)
        k >>= 1
    zi = _inv(acc[2])
    return (acc[0] * zi % Q, acc[1] * zi % Q)


def jj_clear_cofactor(p: Point) -> Point:
    return jj_mul(p, 8)


def jj_on_curve(p: Point) -> bool:
    u2, v2 = p[0] * p[0] % Q, p[1] * p[1] % Q
    return (v2 - u2 - 1 - EDWARDS_D * u2 % Q * v2) % Q == 0


def _sqrt(a):
    """A square root in F_q (q = 1 mod 2^32: Tonelli-Shanks), or None."""
    a %= Q
    if a == 0:
        return 0
    if pow(a, (Q - 1) // 2, Q) != 1:
        return None
    s, t = 32, (Q - 1) >> 32
    z = pow(7, t, Q)            # 7 generates the 2^32 torsion (ROOT_OF_UNITY = 7^t)
    x, b, m = pow(a, (t + 1) // 2, Q), pow(a, t, Q), s
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2, i = b2 * b2 % Q, i + 1
        w = pow(z, 1 << (m - i - 1), Q)
        x, z = x * w % Q, w * w % Q
        b, m = b * z % Q, i
    return x


def jj_to_bytes(p: Point) -> bytes:
    """jubjub::AffinePoint::to_bytes: v little-endian, the sign (lowest bit) of u in the top bit."""
    b = bytearray(p[1].to_bytes(32, "little"))
    b[31] |= (p[0] & 1) << 7
    return bytes(b)


def jj_from_bytes(b: bytes) -> Optional[Point]:
    """jubjub::AffinePoint::from_bytes (ZIP 216 rules: canonical v, and u = 0 only with sign 0)."""
    if len(b) != 32:
        return None
    sign = b[31] >> 7
    v = int.from_bytes(b, "little") & ((1 << 255) - 1)
    if v >= Q:
        return None
    v2 = v * v % Q
    u = _sqrt((v2 - 1) * _inv(1 + EDWARDS_D * v2) % Q)
    if u is None:
        return None
    if (u & 1) != sign:
        u = (-u) % Q
    if u == 0 and sign:
        return None
    return (u, v)


def _blake2s(person, *parts):
    h = hashlib.blake2s(digest_size=32, person=person)
    for p in parts:
        h.update(p)
    return h.digest()


def group_hash(tag: bytes, personalization: bytes) -> Optional[Point]:
    """sapling/group_hash.rs:15-43: a prime-order, non-identity point or None."""
    assert len(personalization) == 8
    p = jj_from_bytes(_blake2s(personalization, GH_FIRST_BLOCK, tag))
    if p is None:
        return None
    p = jj_clear_cofactor(p)
    return None if p == IDENTITY else p


def find_group_hash(m: bytes, personalization: bytes) -> Point:
    """constants.rs:305-322 (how every fixed generator of the protocol is derived)."""
    for i in range(255):
        gh = group_hash(m + bytes([i]), personalization)
        if gh is not None:
            return gh
    raise AssertionError("no group hash below 255")


PROOF_GENERATION_KEY_GENERATOR = C.PROOF_GENERATION_KEY_GENERATOR
SPENDING_KEY_GENERATOR = C.SPENDING_KEY_GENERATOR
VALUE_COMMITMENT_RANDOMNESS_GENERATOR = C.VALUE_COMMITMENT_RANDOMNESS_GENERATOR
# constants.rs:70-110
NOTE_COMMITMENT_RANDOMNESS_GENERATOR = (0x434C9BE15267B091C6DE7556ABB84082CD80EDF5FE44C7BFFC033FA2BF88CB2E,
                                        0x29E2926993D3BC736D277197E97AF8F0690B295C66B85C64C6B8DAA0EE22AEED)
NULLIFIER_POSITION_GENERATOR = (0x50DE6D98FEE5282F84678DC2D85293DF1E09674F28A4B844AAFEE844265FC1E7,
                                0x03260F0BF1244050F3F70DC31AFE799D226945AEE96DFE0AED034E3EE13A1EB3)


def bytes_to_bits_le(b: bytes) -> List[int]:
    return [(x >> i) & 1 for x in b for i in range(8)]


def compute_multipacking(bits: Sequence[int]) -> List[int]:
    """bellman multipack::compute_multipacking: 254-bit little-endian chunks."""
    return [sum(bit << i for i, bit in enumerate(bits[k:k + 254])) for k in range(0, len(bits), 254)]


def jubjub_fr_from_bytes_wide(b: bytes) -> int:
    assert len(b) == 64
    return int.from_bytes(b, "little") % JUBJUB_ORDER


def prf_expand(sk: bytes, t: bytes) -> bytes:
    return hashlib.blake2b(sk + t, digest_size=64, person=PRF_EXPAND_PERSONALIZATION).digest()


# ---------------------------------------------------------------------------
# key and note types (only what the prover's signatures need)
# ---------------------------------------------------------------------------
@dataclass(frozen=True)
class AssetType:
    """asset_type.rs:20-154.  `identifier` is the 32-byte preimage whose BLAKE2s image is the generator."""
    identifier: bytes
    nonce: Optional[int] = None

    @staticmethod
    def _hash_to_point(identifier: bytes) -> Optional[Point]:
        p = jj_from_bytes(_blake2s(VALUE_COMMITMENT_GENERATOR_PERSONALIZATION, identifier))
        if p is None or jj_clear_cofactor(p) == IDENTITY:
            return None
        return p  # cofactor NOT cleared (asset_type.rs:86-88)

    @classmethod
    def new(cls, name: bytes) -> "AssetType":
        for nonce in range(256):
            a = cls.new_with_nonce(name, nonce)
            if a is not None:
                return a
        raise SaplingError("no nonce gives a valid asset identifier")

    @classmethod
    def new_with_nonce(cls, name: bytes, nonce: int) -> Optional["AssetType"]:
        ident = _blake2s(ASSET_IDENTIFIER_PERSONALIZATION, GH_FIRST_BLOCK, name, bytes([nonce]))
        return cls(ident, nonce) if cls._hash_to_point(ident) is not None else None

    @classmethod
    def from_identifier(cls, identifier: bytes) -> Optional["AssetType"]:
        identifier = bytes(identifier)
        if len(identifier) != 32 or cls._hash_to_point(identifier) is None:
            return None
        return cls(identifier, None)

    def get_identifier(self) -> bytes:
        return self.identifier

    def asset_generator(self) -> Point:
        p = self._hash_to_point(self.identifier)
        assert p is not None, "AssetType internal identifier state inconsistent"
        return p

    def value_commitment_generator(self) -> Point:
        return jj_clear_cofactor(self.asset_generator())

    def identifier_bits(self) -> List[int]:
        return bytes_to_bits_le(self.identifier)

    def value_commitment(self, value: int, randomness: int) -> "ValueCommitment":
        return ValueCommitment(self.asset_generator(), value, randomness)

    def __eq__(self, other):
        return isinstance(other, AssetType) and self.identifier == other.identifier

    def __hash__(self):
        return hash(self.identifier)


@dataclass
class ValueCommitment:
    """sapling.rs:196-209."""
    asset_generator: Point
    value: int
    randomness: int

    def commitment(self) -> Point:
        return jj_add(jj_mul(jj_clear_cofactor(self.asset_generator), self.value),
                      jj_mul(VALUE_COMMITMENT_RANDOMNESS_GENERATOR, self.randomness))

    def opening(self) -> C.ValueCommitmentOpening:
        return C.ValueCommitmentOpening(self.asset_generator, self.value, self.randomness)


@dataclass(frozen=True)
class Diversifier:
    d: bytes  # 11 bytes

    def g_d(self) -> Optional[Point]:
        return group_hash(self.d, KEY_DIVERSIFICATION_PERSONALIZATION)


@dataclass(frozen=True)
class PaymentAddress:
    diversifier: Diversifier
    pk_d: Point

    @classmethod
    def from_parts(cls, diversifier, pk_d) -> Optional["PaymentAddress"]:
        return None if pk_d == IDENTITY else cls(diversifier, pk_d)

    def g_d(self) -> Optional[Point]:
        return self.diversifier.g_d()

    def create_note(self, asset_type, value, rseed) -> Optional["Note"]:
        g_d = self.g_d()
        return None if g_d is None else Note(asset_type, value, g_d, self.pk_d, rseed)


@dataclass(frozen=True)
class ViewingKey:
    ak: Point
    nk: Point

    def rk(self, ar: int) -> Point:
        return jj_add(self.ak, jj_mul(SPENDING_KEY_GENERATOR, ar))

    def ivk(self) -> int:
        h = bytearray(_blake2s(CRH_IVK_PERSONALIZATION, jj_to_bytes(self.ak), jj_to_bytes(self.nk)))
        h[31] &= 0b0000_0111
        return int.from_bytes(h, "little")

    def to_payment_address(self, diversifier: Diversifier) -> Optional[PaymentAddress]:
        g_d = diversifier.g_d()
        if g_d is None:
            return None
        return PaymentAddress.from_parts(diversifier, jj_mul(g_d, self.ivk()))


@dataclass(frozen=True)
class ProofGenerationKey:
    ak: Point
    nsk: int

    def to_viewing_key(self) -> ViewingKey:
        return ViewingKey(self.ak, jj_mul(PROOF_GENERATION_KEY_GENERATOR, self.nsk))


@dataclass(frozen=True)
class Rseed:
    """Rseed::BeforeZip212(rcm) or Rseed::AfterZip212(bytes)."""
    rcm: Optional[int] = None
    rseed: Optional[bytes] = None

    @classmethod
    def before_zip212(cls, rcm: int):
        return cls(rcm=int(rcm))

    @classmethod
    def after_zip212(cls, rseed: bytes):
        return cls(rseed=bytes(rseed))


NOTE_COMMITMENT_PERSONALIZATION = [1, 1, 1, 1, 1, 1]  # Personalization::NoteCommitment (pedersen_hash.rs:17-30)


@dataclass
class Note:
    """sapling.rs:762-893."""
    asset_type: AssetType
    value: int
    g_d: Point
    pk_d: Point
    rseed: Rseed

    def rcm(self) -> int:
        if self.rseed.rseed is None:
            return self.rseed.rcm
        return jubjub_fr_from_bytes_wide(prf_expand(self.rseed.rseed, b"\x04"))

    def cm_full_point(self) -> Point:
        contents = (jj_to_bytes(self.asset_type.asset_generator()) + int(self.value).to_bytes(8, "little") +
                    jj_to_bytes(self.g_d) + jj_to_bytes(self.pk_d))
        # the Pedersen hash is the library's (mb200_pedersen_hash, pinned on the reference's vectors)
        h = C.pedersen_hash(NOTE_COMMITMENT_PERSONALIZATION, bytes_to_bits_le(contents))
        return jj_add(jj_mul(NOTE_COMMITMENT_RANDOMNESS_GENERATOR, self.rcm()), h)

    def cmu(self) -> int:
        return self.cm_full_point()[0]

    def nf(self, nk: Point, position: int) -> bytes:
        rho = jj_add(self.cm_full_point(), jj_mul(NULLIFIER_POSITION_GENERATOR, position))
        return _blake2s(PRF_NF_PERSONALIZATION, jj_to_bytes(nk), jj_to_bytes(rho))


SAPLING_COMMITMENT_TREE_DEPTH = C.TREE_DEPTH
UNCOMMITTED = 1   # Note::uncommitted(): the smallest u that is not on the curve (sapling.rs:789-793)


def merkle_hash(depth: int, lhs: int, rhs: int) -> int:
    """sapling.rs:54-80 on scalars: Pedersen hash, personalised by the level, of the low 255 bits of each child."""
    pers = [(depth >> i) & 1 for i in range(6)]   # Personalization::MerkleTree(depth).get_bits()
    bits = [(lhs >> i) & 1 for i in range(255)] + [(rhs >> i) & 1 for i in range(255)]
    return C.pedersen_hash(pers, bits)[0]


def empty_root(level: int) -> int:
    """Node::empty_root (sapling.rs:131-133 over merkle_tree.rs EMPTY_ROOTS)."""
    r = UNCOMMITTED
    for d in range(level):
        r = merkle_hash(d, r, r)
    return r


@dataclass
class MerklePath:
    """merkle_tree.rs MerklePath<Node>: (sibling, is-right-child) pairs from the leaf up, and the leaf position."""
    auth_path: List[Tuple[object, bool]]
    position: int

    @classmethod
    def from_position(cls, siblings, position: int):
        """The path whose left/right flags are the bits of `position`, as CommitmentTree::path builds it."""
        return cls([(s, bool((position >> i) & 1)) for i, s in enumerate(siblings)], position)

    def root(self, leaf: int) -> int:
        """MerklePath::root (merkle_tree.rs): fold merkle_hash from the leaf up."""
        cur = leaf
        for i, (sib, right) in enumerate(self.scalars()):
            cur = merkle_hash(i, sib, cur) if right else merkle_hash(i, cur, sib)
        return cur

    def scalars(self) -> List[Tuple[int, bool]]:
        out = []
        for node, b in self.auth_path:
            v = node if isinstance(node, int) else int.from_bytes(bytes(node), "little")
            if v >= Q:
                raise ValueError("Merkle node is not a canonical scalar")
            out.append((v, bool(b)))
        return out


class AllowedConversion:
    """convert.rs:23-120: a sum of (asset, signed amount) whose generator is sum amount_i * asset_generator_i."""

    def __init__(self, assets: Dict[AssetType, int]):
        self.assets = dict(assets)
        g = IDENTITY
        for asset, value in self.assets.items():
            if not -(1 << 127) < value < (1 << 127):
                raise ValueError("invalid conversion")
            t = jj_mul(asset.asset_generator(), abs(value) % (1 << 64))  # `a as u64` (convert.rs:100-103)
            g = jj_add(g, jj_neg(t) if value < 0 else t)
        self.generator = g

    def cm_full_point(self) -> Point:
        return C.pedersen_hash(NOTE_COMMITMENT_PERSONALIZATION, bytes_to_bits_le(jj_to_bytes(self.generator)))

    def cmu(self) -> int:
        return self.cm_full_point()[0]

    def value_commitment(self, value: int, randomness: int) -> ValueCommitment:
        return ValueCommitment(self.generator, value, randomness)


# ---------------------------------------------------------------------------
# RedJubjub (binding signature)
# ---------------------------------------------------------------------------
def _h_star(a: bytes, b: bytes) -> int:
    return jubjub_fr_from_bytes_wide(hashlib.blake2b(a + b, digest_size=64, person=REDJUBJUB_H_PERSONALIZATION).digest())


def redjubjub_sign(sk: int, msg: bytes, p_g: Point, rng=os.urandom) -> bytes:
    """PrivateKey::sign (redjubjub.rs:133-157): rbar || sbar."""
    t = rng(80)
    r = _h_star(t, msg)
    rbar = jj_to_bytes(jj_mul(p_g, r))
    s = (_h_star(rbar, msg) * sk + r) % JUBJUB_ORDER
    return rbar + s.to_bytes(32, "little")


def redjubjub_verify(vk: Point, msg: bytes, sig: bytes, p_g: Point) -> bool:
    """PublicKey::verify (redjubjub.rs:190-260): [8](-[c]vk + R - [S]P_G)... == identity, with ZIP 216."""
    if len(sig) != 64:
        return False
    c = _h_star(sig[:32], msg)
    r = jj_from_bytes(sig[:32])
    s = int.from_bytes(sig[32:], "little")
    if r is None or s >= JUBJUB_ORDER:
        return False
    lhs = jj_add(jj_add(jj_mul(vk, c), r), jj_neg(jj_mul(p_g, s)))
    return jj_clear_cofactor(lhs) == IDENTITY


def masp_compute_value_balance(asset_type: AssetType, value: int) -> Optional[Point]:
    """masp_proofs/src/sapling/mod.rs:14-38."""
    if not -(1 << 127) < value < (1 << 127):
        return None
    t = jj_mul(asset_type.value_commitment_generator(), abs(value))
    return jj_neg(t) if value < 0 else t


# ---------------------------------------------------------------------------
# SaplingProvingContext
# ---------------------------------------------------------------------------
def spend_public_inputs(rk: Point, cv: Point, anchor: int, nullifier: bytes) -> List[int]:
    """sapling/prover.rs:121-145 (= verifier.rs:70-98): rk, cv, anchor, the nullifier multipacked."""
    nf = compute_multipacking(bytes_to_bits_le(nullifier))
    assert len(nf) == 2
    return [rk[0], rk[1], cv[0], cv[1], anchor % Q, nf[0], nf[1]]


def convert_public_inputs(cv: Point, anchor: int) -> List[int]:
    """sapling/prover.rs:256-263."""
    return [cv[0], cv[1], anchor % Q]


def output_public_inputs(cv: Point, epk: Point, cmu: int) -> List[int]:
    """sapling/verifier.rs check_output: cv, epk, cmu."""
    return [cv[0], cv[1], epk[0], epk[1], cmu % Q]


@dataclass
class _Pending:
    kind: str
    instance: object
    public_input: Optional[List[int]]   # None: the reference does not self-check outputs
    proof: Optional[bytes] = None


class SaplingProvingContext:
    """masp_proofs::sapling::SaplingProvingContext.  `bsk` and `cv_sum` are updated exactly where the
    reference updates them (sapling/prover.rs:69-75, 154, 177-183, 205, 228-234, 272); they do not depend
    on the proof bytes, which is what lets BatchingTxProver defer the proofs."""

    def __init__(self):
        self.bsk = 0
        self.cv_sum = IDENTITY
        self._pending: List[_Pending] = []

    @classmethod
    def new(cls):
        return cls()

    # -- instance preparation (everything except create_random_proof / verify_proof) ------------------
    def _prepare_spend(self, proof_generation_key, diversifier, rseed, ar, asset_type, value, anchor, merkle_path,
                       rcv):
        self.bsk = (self.bsk + rcv) % JUBJUB_ORDER
        value_commitment = asset_type.value_commitment(value, rcv)
        viewing_key = proof_generation_key.to_viewing_key()
        payment_address = viewing_key.to_payment_address(diversifier)
        if payment_address is None:
            raise SaplingError("invalid diversifier")
        rk = viewing_key.rk(ar)   # PublicKey(ak).randomize(ar, SPENDING_KEY_GENERATOR)
        note = Note(asset_type, value, payment_address.g_d(), payment_address.pk_d, rseed)
        nullifier = note.nf(viewing_key.nk, merkle_path.position)
        instance = C.Spend(value_commitment.opening(), proof_generation_key.ak, proof_generation_key.nsk, note.g_d,
                           note.rcm(), ar, merkle_path.scalars(), anchor)
        cv = value_commitment.commitment()
        return instance, spend_public_inputs(rk, cv, anchor, nullifier), cv, rk

    def _prepare_output(self, esk, payment_address, rcm, asset_type, value, rcv):
        self.bsk = (self.bsk - rcv) % JUBJUB_ORDER      # outputs subtract from the total
        value_commitment = asset_type.value_commitment(value, rcv)
        g_d = payment_address.g_d()
        if g_d is None:                                  # the circuit's `ok_or(AssignmentMissing)` -> the expect panics
            raise SaplingError("proving should not fail: payment address has no g_d")
        instance = C.Output(value_commitment.opening(), asset_type.get_identifier(), g_d, payment_address.pk_d, rcm,
                            esk)
        return instance, value_commitment.commitment()

    def _prepare_convert(self, allowed_conversion, value, anchor, merkle_path, rcv):
        self.bsk = (self.bsk + rcv) % JUBJUB_ORDER
        value_commitment = allowed_conversion.value_commitment(value, rcv)
        instance = C.Convert(value_commitment.opening(), merkle_path.scalars(), anchor)
        cv = value_commitment.commitment()
        return instance, convert_public_inputs(cv, anchor), cv

    @staticmethod
    def _prove_one(instance, params, rng):
        circ = getattr(params, "circuit", None)
        if circ is None:
            raise ValueError("these parameters have no circuit bound (load them with densities=None)")
        inputs, aux = circ.synthesize([instance])
        return P.create_proof_batch_from_witness(params, inputs, aux, [rng()], [rng()])[0]

    # -- the reference's methods --------------------------------------------------------------------
    def spend_proof(self, proof_generation_key, diversifier, rseed, ar, asset_type, value, anchor, merkle_path,
                    proving_key, rcv, rng=P._os_rng_scalar):
        """-> (proof, cv, rk).  sapling/prover.rs:51-160; `verifying_key` is part of `proving_key` here."""
        instance, public_input, cv, rk = self._prepare_spend(proof_generation_key, diversifier, rseed, ar, asset_type,
                                                             value, anchor, merkle_path, rcv)
        proof = self._prove_one(instance, proving_key, rng)
        if not P.verify_proofs(proving_key, [proof], [public_input])[0]:
            raise SaplingError("spend proof does not verify")
        self.cv_sum = jj_add(self.cv_sum, cv)
        return proof, cv, rk

    def output_proof(self, esk, payment_address, rcm, asset_type, value, proving_key, rcv, rng=P._os_rng_scalar):
        """-> (proof, cv).  sapling/prover.rs:163-208."""
        instance, cv = self._prepare_output(esk, payment_address, rcm, asset_type, value, rcv)
        proof = self._prove_one(instance, proving_key, rng)
        self.cv_sum = jj_add(self.cv_sum, jj_neg(cv))
        return proof, cv

    def convert_proof(self, allowed_conversion, value, anchor, merkle_path, proving_key, rcv,
                      rng=P._os_rng_scalar):
        """-> (proof, cv).  sapling/prover.rs:214-275."""
        instance, public_input, cv = self._prepare_convert(allowed_conversion, value, anchor, merkle_path, rcv)
        proof = self._prove_one(instance, proving_key, rng)
        if not P.verify_proofs(proving_key, [proof], [public_input])[0]:
            raise SaplingError("convert proof does not verify")
        self.cv_sum = jj_add(self.cv_sum, cv)
        return proof, cv

    def binding_sig(self, assets_and_values: Dict[AssetType, int], sighash: bytes, rng=os.urandom) -> bytes:
        """-> 64-byte RedJubjub signature.  sapling/prover.rs:279-326."""
        if len(sighash) != 32:
            raise ValueError("sighash is 32 bytes")
        bvk = jj_mul(VALUE_COMMITMENT_RANDOMNESS_GENERATOR, self.bsk)
        final_bvk = self.cv_sum
        for asset_type, value_balance in assets_and_values.items():
            vb = masp_compute_value_balance(asset_type, value_balance)
            if vb is None:
                raise SaplingError("bad value balance")
            final_bvk = jj_add(final_bvk, jj_neg(vb))
        if bvk != final_bvk:
            raise SaplingError("value balance does not match the accumulated value commitments")
        return redjubjub_sign(self.bsk, jj_to_bytes(bvk) + bytes(sighash), VALUE_COMMITMENT_RANDOMNESS_GENERATOR, rng)


class PendingProof:
    """Handle returned by BatchingTxProver: `.proof` resolves the whole transaction's batch on first use."""

    def __init__(self, owner, ctx, slot):
        self._owner, self._ctx, self._slot = owner, ctx, slot

    @property
    def proof(self) -> bytes:
        if self._slot.proof is None:
            self._owner.flush(self._ctx)
        return self._slot.proof


class TxProver:
    """trait TxProver (masp_primitives/src/sapling/prover.rs:17-83) as LocalTxProver implements it
    (masp_proofs/src/prover.rs:156-261): wraps a masp_b200.prover.LocalTxProver whose keys have the
    library's circuits bound."""

    def __init__(self, local: "P.LocalTxProver"):
        self.local = local

    def new_sapling_proving_context(self) -> SaplingProvingContext:
        return SaplingProvingContext()

    def spend_proof(self, ctx, proof_generation_key, diversifier, rseed, ar, asset_type, value, anchor, merkle_path,
                    rcv, rng=P._os_rng_scalar):
        return ctx.spend_proof(proof_generation_key, diversifier, rseed, ar, asset_type, value, anchor, merkle_path,
                               self.local.spend_params, rcv, rng)

    def output_proof(self, ctx, esk, payment_address, rcm, asset_type, value, rcv, rng=P._os_rng_scalar):
        return ctx.output_proof(esk, payment_address, rcm, asset_type, value, self.local.output_params, rcv, rng)

    def convert_proof(self, ctx, allowed_conversion, value, anchor, merkle_path, rcv, rng=P._os_rng_scalar):
        return ctx.convert_proof(allowed_conversion, value, anchor, merkle_path, self.local.convert_params, rcv, rng)

    def binding_sig(self, ctx, assets_and_values, sighash, rng=os.urandom):
        return ctx.binding_sig(assets_and_values, sighash, rng)


class BatchingTxProver(TxProver):
    """SURVEY.md §8(f)-2.  Same calls; each returns a PendingProof in place of the proof bytes (cv and rk are
    returned at once, the builder needs them for the descriptions).  The serial `.map()` over descriptions in
    components/sapling/builder.rs:941-1140 then costs one witness pass on the host cores and one launch per
    circuit for the whole transaction.  `binding_sig` flushes, so a transaction cannot be signed with a
    proof still unresolved; a Spend or Convert proof that fails verify_proof raises SaplingError there."""

    def _enqueue(self, ctx, kind, instance, public_input):
        slot = _Pending(kind, instance, public_input)
        ctx._pending.append(slot)
        return PendingProof(self, ctx, slot)

    def spend_proof(self, ctx, proof_generation_key, diversifier, rseed, ar, asset_type, value, anchor, merkle_path,
                    rcv, rng=None):
        instance, public_input, cv, rk = ctx._prepare_spend(proof_generation_key, diversifier, rseed, ar, asset_type,
                                                            value, anchor, merkle_path, rcv)
        ctx.cv_sum = jj_add(ctx.cv_sum, cv)
        return self._enqueue(ctx, "spend", instance, public_input), cv, rk

    def output_proof(self, ctx, esk, payment_address, rcm, asset_type, value, rcv, rng=None):
        instance, cv = ctx._prepare_output(esk, payment_address, rcm, asset_type, value, rcv)
        ctx.cv_sum = jj_add(ctx.cv_sum, jj_neg(cv))
        return self._enqueue(ctx, "output", instance, None), cv

    def convert_proof(self, ctx, allowed_conversion, value, anchor, merkle_path, rcv, rng=None):
        instance, public_input, cv = ctx._prepare_convert(allowed_conversion, value, anchor, merkle_path, rcv)
        ctx.cv_sum = jj_add(ctx.cv_sum, cv)
        return self._enqueue(ctx, "convert", instance, public_input), cv

    def flush(self, ctx, rng=P._os_rng_scalar):
        todo = [s for s in ctx._pending if s.proof is None]
        by = {k: [s for s in todo if s.kind == k] for k in ("spend", "convert", "output")}
        spends, converts, outputs = self.local.prove_bundle(spends=[s.instance for s in by["spend"]],
                                                            converts=[s.instance for s in by["convert"]],
                                                            outputs=[s.instance for s in by["output"]], rng=rng)
        bad = []
        for kind, params, proofs in (("spend", self.local.spend_params, spends),
                                     ("convert", self.local.convert_params, converts)):
            if proofs:
                ok = P.verify_proofs(params, proofs, [s.public_input for s in by[kind]])
                bad += [kind for good in ok if not good]
        if bad:
            raise SaplingError("%d proof(s) of the transaction do not verify (%s)" % (len(bad), ", ".join(sorted(set(bad)))))
        for kind, proofs in (("spend", spends), ("convert", converts), ("output", outputs)):
            for s, p in zip(by[kind], proofs):
                s.proof = p
        ctx._pending = []

    def binding_sig(self, ctx, assets_and_values, sighash, rng=os.urandom):
        self.flush(ctx)
        return ctx.binding_sig(assets_and_values, sighash, rng)


# ---------------------------------------------------------------------------
# where the proofs go: the Sapling part of a v5 transaction
# ---------------------------------------------------------------------------
ENC_CIPHERTEXT_SIZE = 580 + 32   # transaction/components/sapling.rs (note plaintext with the asset identifier)
OUT_CIPHERTEXT_SIZE = 80


def _compact_size(n: int) -> bytes:
    """zcash_encoding::CompactSize::write."""
    if n < 253:
        return bytes([n])
    if n <= 0xFFFF:
        return b"\xfd" + n.to_bytes(2, "little")
    if n <= 0xFFFFFFFF:
        return b"\xfe" + n.to_bytes(4, "little")
    return b"\xff" + n.to_bytes(8, "little")


def _read_compact_size(buf: bytes, off: int):
    flag = buf[off]
    if flag < 253:
        return flag, off + 1
    width = {253: 2, 254: 4, 255: 8}[flag]
    n = int.from_bytes(buf[off + 1:off + 1 + width], "little")
    if n < {2: 253, 4: 0x10000, 8: 0x100000000}[width]:
        raise ValueError("non-canonical CompactSize")
    return n, off + 1 + width


@dataclass
class SpendDescription:
    """transaction/components/sapling.rs SpendDescription<Authorized>."""
    cv: Point
    anchor: int
    nullifier: bytes
    rk: Point
    zkproof: bytes
    spend_auth_sig: bytes


@dataclass
class ConvertDescription:
    cv: Point
    anchor: int
    zkproof: bytes


@dataclass
class OutputDescription:
    cv: Point
    cmu: int
    ephemeral_key: bytes
    enc_ciphertext: bytes
    out_ciphertext: bytes
    zkproof: bytes


@dataclass
class SaplingBundle:
    """transaction/components/sapling.rs Bundle<Authorized>."""
    shielded_spends: List[SpendDescription] = field(default_factory=list)
    shielded_converts: List[ConvertDescription] = field(default_factory=list)
    shielded_outputs: List[OutputDescription] = field(default_factory=list)
    value_balance: Dict[AssetType, int] = field(default_factory=dict)   # I128Sum
    binding_sig: bytes = b""

    def is_empty(self):
        return not (self.shielded_spends or self.shielded_converts or self.shielded_outputs)


def write_v5_sapling(bundle: Optional[SaplingBundle]) -> bytes:
    """Transaction::write_v5_sapling (transaction.rs:746-806): the descriptions without their proofs, the
    value balance, the shared anchors, then the proofs as arrays (spend proofs, spend auth signatures,
    convert proofs, output proofs) and the binding signature."""
    if bundle is None:
        return _compact_size(0) * 3
    w = bytearray()
    if any(len(s.zkproof) != P.GROTH_PROOF_SIZE for s in
           bundle.shielded_spends + bundle.shielded_converts + bundle.shielded_outputs):
        raise ValueError("a zkproof is %d bytes" % P.GROTH_PROOF_SIZE)
    w += _compact_size(len(bundle.shielded_spends))
    for s in bundle.shielded_spends:       # write_v5_without_witness_data (sapling.rs:242-246)
        if len(s.nullifier) != 32 or len(s.spend_auth_sig) != 64:
            raise ValueError("nullifier is 32 bytes, spend_auth_sig 64")
        w += jj_to_bytes(s.cv) + s.nullifier + jj_to_bytes(s.rk)
    w += _compact_size(len(bundle.shielded_converts))
    for c in bundle.shielded_converts:     # :564-566
        w += jj_to_bytes(c.cv)
    w += _compact_size(len(bundle.shielded_outputs))
    for o in bundle.shielded_outputs:      # write_v5_without_proof (:361-367)
        if (len(o.ephemeral_key), len(o.enc_ciphertext), len(o.out_ciphertext)) != (32, ENC_CIPHERTEXT_SIZE,
                                                                                   OUT_CIPHERTEXT_SIZE):
            raise ValueError("output description field sizes")
        w += jj_to_bytes(o.cv) + (o.cmu % Q).to_bytes(32, "little") + o.ephemeral_key + o.enc_ciphertext + \
            o.out_ciphertext
    if not bundle.is_empty():              # I128Sum::write (amount.rs:361-368): BTreeMap order = identifier order
        items = sorted(((a.get_identifier(), v) for a, v in bundle.value_balance.items() if v != 0))
        w += _compact_size(len(items))
        for ident, v in items:
            w += ident + int(v).to_bytes(16, "little", signed=True)
    if bundle.shielded_spends:
        if any(s.anchor != bundle.shielded_spends[0].anchor for s in bundle.shielded_spends):
            raise ValueError("v5 carries one anchor for all spends")
        w += (bundle.shielded_spends[0].anchor % Q).to_bytes(32, "little")
    if bundle.shielded_converts:
        if any(c.anchor != bundle.shielded_converts[0].anchor for c in bundle.shielded_converts):
            raise ValueError("v5 carries one anchor for all converts")
        w += (bundle.shielded_converts[0].anchor % Q).to_bytes(32, "little")
    w += b"".join(s.zkproof for s in bundle.shielded_spends)
    w += b"".join(s.spend_auth_sig for s in bundle.shielded_spends)
    w += b"".join(c.zkproof for c in bundle.shielded_converts)
    w += b"".join(o.zkproof for o in bundle.shielded_outputs)
    if not bundle.is_empty():
        if len(bundle.binding_sig) != 64:
            raise ValueError("binding_sig is 64 bytes")
        w += bundle.binding_sig
    return bytes(w)


def read_v5_sapling(buf: bytes, off: int = 0):
    """Transaction::read_v5_sapling (transaction.rs:612-720) -> (SaplingBundle or None, next offset).
    Canonical encodings are enforced where the reference enforces them (points, base-field scalars,
    asset identifiers)."""
    def take(n):
        nonlocal off
        if off + n > len(buf):
            raise ValueError("truncated")
        off += n
        return bytes(buf[off - n:off])

    def point(what):
        p = jj_from_bytes(take(32))
        if p is None:
            raise ValueError(what + " not in field / on curve")
        return p

    def base(what):
        v = int.from_bytes(take(32), "little")
        if v >= Q:
            raise ValueError(what + " not in field")
        return v

    n_spends, off = _read_compact_size(buf, off)
    sd = [(point("cv"), take(32), point("rk")) for _ in range(n_spends)]
    n_converts, off = _read_compact_size(buf, off)
    cd = [point("cv") for _ in range(n_converts)]
    n_outputs, off = _read_compact_size(buf, off)
    od = [(point("cv"), base("cmu"), take(32), take(ENC_CIPHERTEXT_SIZE), take(OUT_CIPHERTEXT_SIZE))
          for _ in range(n_outputs)]
    if not (n_spends or n_converts or n_outputs):
        return None, off
    n_assets, off = _read_compact_size(buf, off)
    vb: Dict[AssetType, int] = {}
    for _ in range(n_assets):
        a = AssetType.from_identifier(take(32))
        if a is None:
            raise ValueError("invalid asset type")
        vb[a] = vb.get(a, 0) + int.from_bytes(take(16), "little", signed=True)
    spend_anchor = base("spend anchor") if n_spends else None
    convert_anchor = base("convert anchor") if n_converts else None
    sp = [take(P.GROTH_PROOF_SIZE) for _ in range(n_spends)]
    sigs = [take(64) for _ in range(n_spends)]
    cp = [take(P.GROTH_PROOF_SIZE) for _ in range(n_converts)]
    op = [take(P.GROTH_PROOF_SIZE) for _ in range(n_outputs)]
    binding = take(64)
    bundle = SaplingBundle(
        [SpendDescription(cv, spend_anchor, nf, rk, z, sig) for (cv, nf, rk), z, sig in zip(sd, sp, sigs)],
        [ConvertDescription(cv, convert_anchor, z) for cv, z in zip(cd, cp)],
        [OutputDescription(cv, cmu, epk, enc, out, z) for (cv, cmu, epk, enc, out), z in zip(od, op)],
        {a: v for a, v in vb.items() if v != 0}, binding)
    return bundle, off


# ---------------------------------------------------------------------------
# the verifier side of the same proofs
# ---------------------------------------------------------------------------
def spend_sig(ask: int, ar: int, sighash: bytes, rng=os.urandom) -> bytes:
    """sapling.rs:166-195: the spendAuthSig under rsk = ask + ar, over rk || sighash."""
    rsk = (ask + ar) % JUBJUB_ORDER
    rk = jj_mul(SPENDING_KEY_GENERATOR, rsk)
    return redjubjub_sign(rsk, jj_to_bytes(rk) + bytes(sighash), SPENDING_KEY_GENERATOR, rng)


def jj_is_small_order(p: Point) -> bool:
    return jj_clear_cofactor(p) == IDENTITY


def _device_verify(params, zkproof, public_input):
    return P.verify_proofs(params, [zkproof], [public_input])[0]


class SaplingVerificationContext:
    """masp_proofs::sapling::SaplingVerificationContext (sapling/verifier.rs:19-215 with the closures of
    verifier/single.rs): consensus checks per description, cv_sum bookkeeping, final_check of the
    binding signature.  `verifying_key` is a masp_b200 Parameters object (its vk part is used);
    verify_proof runs on the device (mb200_verify_proofs: Proof::read + the pairing check)."""

    def __init__(self, zip216_enabled=True, proof_verifier=_device_verify):
        self.cv_sum = IDENTITY
        self.zip216_enabled = zip216_enabled
        self._verify = proof_verifier

    def check_spend(self, cv, anchor, nullifier, rk, sighash_value, spend_auth_sig, zkproof, verifying_key) -> bool:
        if jj_is_small_order(cv) or jj_is_small_order(rk):
            return False
        self.cv_sum = jj_add(self.cv_sum, cv)
        msg = jj_to_bytes(rk) + bytes(sighash_value)
        if not redjubjub_verify(rk, msg, spend_auth_sig, SPENDING_KEY_GENERATOR):
            return False
        return bool(self._verify(verifying_key, zkproof, spend_public_inputs(rk, cv, anchor, nullifier)))

    def check_convert(self, cv, anchor, zkproof, verifying_key) -> bool:
        if jj_is_small_order(cv):
            return False
        self.cv_sum = jj_add(self.cv_sum, cv)
        return bool(self._verify(verifying_key, zkproof, convert_public_inputs(cv, anchor)))

    def check_output(self, cv, cmu, epk, zkproof, verifying_key) -> bool:
        if jj_is_small_order(cv) or jj_is_small_order(epk):
            return False
        self.cv_sum = jj_add(self.cv_sum, jj_neg(cv))
        return bool(self._verify(verifying_key, zkproof, output_public_inputs(cv, epk, cmu)))

    def _bvk(self, value_balance: Dict[AssetType, int]) -> Optional[Point]:
        bvk = self.cv_sum
        for asset_type, v in value_balance.items():
            vb = masp_compute_value_balance(asset_type, v)
            if vb is None:
                return None
            bvk = jj_add(bvk, jj_neg(vb))
        return bvk

    def final_check(self, value_balance: Dict[AssetType, int], sighash_value: bytes, binding_sig: bytes) -> bool:
        bvk = self._bvk(value_balance)
        if bvk is None:
            return False
        return redjubjub_verify(bvk, jj_to_bytes(bvk) + bytes(sighash_value), binding_sig,
                                VALUE_COMMITMENT_RANDOMNESS_GENERATOR)


class BatchValidator:
    """masp_proofs::sapling::BatchValidator (sapling/verifier/batch.rs:60-243): `check_bundle` runs the
    per-description consensus checks and queues signatures and proofs, `validate` verifies every queued
    signature and then each circuit's proofs with ONE randomised batch check on the device
    (mb200_verify_proofs_batch: n + 3 Miller loops, one final exponentiation).  All-or-nothing like the
    reference.  One difference in timing, none in outcome: a zkproof that does not even parse
    (Proof::read) makes the reference's check_bundle return false at once; here proofs travel in wire
    form and are read on the device, so such a bundle is rejected by `validate`."""

    def __init__(self, batch_verifier=None):
        self.bundles_added = False
        self.spend_proofs: List[Tuple[bytes, List[int]]] = []
        self.convert_proofs: List[Tuple[bytes, List[int]]] = []
        self.output_proofs: List[Tuple[bytes, List[int]]] = []
        self.signatures: List[Tuple[Point, bytes, bytes, Point]] = []   # (vk, msg, sig, generator)
        self._batch_verify = batch_verifier or P.verify_proofs_batch

    def check_bundle(self, bundle: SaplingBundle, sighash: bytes) -> bool:
        self.bundles_added = True
        ctx = SaplingVerificationContext(proof_verifier=lambda vk, proof, inputs: True)
        for s in bundle.shielded_spends:
            if jj_is_small_order(s.cv) or jj_is_small_order(s.rk):
                return False
            ctx.cv_sum = jj_add(ctx.cv_sum, s.cv)
            self.signatures.append((s.rk, jj_to_bytes(s.rk) + bytes(sighash), s.spend_auth_sig,
                                    SPENDING_KEY_GENERATOR))
            self.spend_proofs.append((s.zkproof, spend_public_inputs(s.rk, s.cv, s.anchor, s.nullifier)))
        for c in bundle.shielded_converts:
            if not ctx.check_convert(c.cv, c.anchor, c.zkproof, None):
                return False
            self.convert_proofs.append((c.zkproof, convert_public_inputs(c.cv, c.anchor)))
        for o in bundle.shielded_outputs:
            epk = jj_from_bytes(o.ephemeral_key)
            if epk is None or not ctx.check_output(o.cv, o.cmu, epk, o.zkproof, None):
                return False
            self.output_proofs.append((o.zkproof, output_public_inputs(o.cv, epk, o.cmu)))
        bvk = ctx._bvk(bundle.value_balance)
        if bvk is None:
            return False
        self.signatures.append((bvk, jj_to_bytes(bvk) + bytes(sighash), bundle.binding_sig,
                                VALUE_COMMITMENT_RANDOMNESS_GENERATOR))
        return True

    def validate(self, spend_vk, convert_vk, output_vk, rng=os.urandom) -> bool:
        if not self.bundles_added:
            return True
        for vk, msg, sig, gen in self.signatures:
            if not redjubjub_verify(vk, msg, sig, gen):
                return False
        for batch, vk in ((self.spend_proofs, spend_vk), (self.convert_proofs, convert_vk),
                          (self.output_proofs, output_vk)):
            if batch and not self._batch_verify(vk, [p for p, _ in batch], [x for _, x in batch], rng):
                return False
        return True
