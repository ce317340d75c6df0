"""Host-side mirror of the reference's proving surface over the C ABI.

Mirrors, for the Groth16 path only:

  * groth16::Parameters::read(reader, false)       masp_proofs/src/lib.rs:336-341
  * load_parameters / parse_parameters             masp_proofs/src/lib.rs:278-403
  * create_random_proof / create_proof             call sites masp_proofs/src/sapling/prover.rs:116-117, 201-202, 251-252
  * LocalTxProver::{new, from_bytes, with_default_location, spend_proof,
    output_proof, convert_proof}                   masp_proofs/src/prover.rs:55-136, 156-261

Witness synthesis (Circuit::synthesize into bellman's ProvingAssignment) is
above this path (NEXT-1): the *_proof methods take the
ProvingAssignment a Rust caller would hand to the FFI -- the per-row
evaluations a, b, c, the input and aux assignments -- plus, at key load, the
three density bitmaps -- or, for keys loaded with the library's own circuits
bound, the circuit instance itself.  The callers' side of those methods
(SaplingProvingContext: bsk, cv_sum, binding_sig, and the TxProver trait with
the wallet-level argument lists) is masp_b200/sapling.py.
"""
import ctypes
import hashlib
import os
from dataclasses import dataclass

from . import _lib
from ._lib import Mb200Error, check

GROTH_PROOF_SIZE = 192  # masp_primitives/src/transaction/components.rs:14-15

# masp_proofs/src/lib.rs:61-76
MASP_SPEND_NAME = "masp-spend.params"
MASP_OUTPUT_NAME = "masp-output.params"
MASP_CONVERT_NAME = "masp-convert.params"
MASP_SPEND_HASH = "196e7c717f25e16653431559ce2c8816e750a4490f98696e3c031efca37e25e0647182b7b013660806db11eb2b1e365fb2d6a0f24dbbd9a4a8314fef10a7cba2"
MASP_OUTPUT_HASH = "eafc3b1746cccc8b9eed2b69395692c5892f6aca83552a07dceb2dcbaa64dcd0e22434260b3aa3b049b633a08b008988cbe0d31effc77e2bc09bfab690a23724"
MASP_CONVERT_HASH = "dc4aaf3c3ce056ab448b6c4a7f43c1d68502c2902ea89ab8769b1524a2e8ace9a5369621a73ee1daa52aec826907a19974a37874391cf8f11bbe0b0420de1ab7"
MASP_SPEND_BYTES = 49848572
MASP_CONVERT_BYTES = 22570940
MASP_OUTPUT_BYTES = 16398620

R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

_initialised = False


def init(device=None):
    """Open the GPU(s) this process drives: None = the current CUDA device, an int = that
    device, a list = those devices (the first is the primary one), "all" = every visible device.
    With several devices a key is replicated to each and every batch is cut into per-device
    slices inside the library (one prover object for the whole process, like the reference's
    LocalTxProver behind `&self`, masp_proofs/src/prover.rs:27-33, 156-261)."""
    global _initialised
    L = _lib.lib()
    if device is None:
        check(L.mb200_init(None, 0))
    elif device == "all":
        check(L.mb200_init(None, -1))
    elif isinstance(device, (list, tuple)):
        ids = (ctypes.c_int * len(device))(*[int(d) for d in device])
        check(L.mb200_init(ids, len(device)))
    else:
        ids = (ctypes.c_int * 1)(int(device))
        check(L.mb200_init(ids, 1))
    _initialised = True


def shutdown():
    global _initialised
    check(_lib.lib().mb200_shutdown())
    _initialised = False


def device_count():
    return int(_lib.lib().mb200_device_count())


def _ensure_init():
    if not _initialised:
        init()


def set_option(name, value):
    _ensure_init()
    check(_lib.lib().mb200_set_option(name.encode(), int(value)))


def get_counter(name):
    v = ctypes.c_double()
    check(_lib.lib().mb200_get_counter(name.encode(), ctypes.byref(v)))
    return v.value


def _ptr(x):
    """Host bytes-like or an integer device / pinned-host address -> c_void_p."""
    if x is None:
        return ctypes.c_void_p(None)
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(x)) if isinstance(x, bytearray) else ctypes.c_char_p(x), ctypes.c_void_p)
    if isinstance(x, ctypes.Array):
        return ctypes.cast(x, ctypes.c_void_p)
    if hasattr(x, "ctypes"):  # numpy array
        return ctypes.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # torch tensor
        return ctypes.c_void_p(x.data_ptr())
    raise TypeError("unsupported buffer type %r" % type(x))


class Parameters:
    """groth16::Parameters<Bls12>, device resident.

    `densities` = (a_aux_density, b_input_density, b_aux_density) bitmaps
    (LSB-first) as bellman's ProvingAssignment records them for the circuit;
    None means every variable is dense in that query."""

    def __init__(self, handle, info):
        self._h = handle
        (self.n_inputs, self.n_aux, self.h_len, self.a_len, self.b_len, self.m, self.consumed,
         self.table_bytes, self.window_hl, self.window_a) = info

    @classmethod
    def read(cls, buf, densities=None, checked=False):
        """Parameters::read(reader, checked); the reference passes checked = false
        (masp_proofs/src/lib.rs:336-341) and so does this path: encodings must be
        well formed, curve / subgroup membership is not tested."""
        if checked:
            raise NotImplementedError("checked = true is not on the reference's path (lib.rs:336-341)")
        _ensure_init()
        a_d, bi_d, ba_d = densities if densities is not None else (None, None, None)
        h = ctypes.c_void_p()
        buf = bytes(buf) if not isinstance(buf, bytes) else buf
        check(_lib.lib().mb200_params_load(buf, len(buf), a_d, bi_d, ba_d, ctypes.byref(h)))
        info = (ctypes.c_uint64 * 10)()
        check(_lib.lib().mb200_params_info(h, info))
        return cls(h, [int(x) for x in info])

    @classmethod
    def read_verified(cls, buf, expected_bytes, expected_hash, densities=None):
        """One stream of parse_parameters (masp_proofs/src/lib.rs:343-388) behind the C ABI:
        Parameters::read(.., false), then BLAKE2b-512 over the whole stream (transcript included)."""
        _ensure_init()
        a_d, bi_d, ba_d = densities if densities is not None else (None, None, None)
        h = ctypes.c_void_p()
        buf = bytes(buf) if not isinstance(buf, bytes) else buf
        check(_lib.lib().mb200_params_load_verified(buf, len(buf), int(expected_bytes or 0),
                                                    (expected_hash or "").encode(), a_d, bi_d, ba_d, ctypes.byref(h)))
        info = (ctypes.c_uint64 * 10)()
        check(_lib.lib().mb200_params_info(h, info))
        return cls(h, [int(x) for x in info])

    @classmethod
    def read_file(cls, path, expected_bytes, expected_hash, densities=None):
        """One file of load_parameters (lib.rs:278-325): size check before anything is read,
        then as read_verified.  The file never passes through Python."""
        _ensure_init()
        a_d, bi_d, ba_d = densities if densities is not None else (None, None, None)
        h = ctypes.c_void_p()
        check(_lib.lib().mb200_params_load_file(os.fsencode(path), int(expected_bytes or 0), (expected_hash or "").encode(),
                                                a_d, bi_d, ba_d, ctypes.byref(h)))
        info = (ctypes.c_uint64 * 10)()
        check(_lib.lib().mb200_params_info(h, info))
        return cls(h, [int(x) for x in info])

    def bind_circuit(self, circuit):
        """Attach a recorded circuit (masp_b200.circuits.Circuit): its matrices go to
        the device once, after which create_proof_batch_from_witness needs only
        the witnesses.  The key and the circuit must agree (counts, densities)."""
        check(_lib.lib().mb200_params_bind_circuit(self._h, circuit._h))
        self.circuit = circuit
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().mb200_params_free(h)
            except Exception:
                pass
            self._h = None


@dataclass
class ProvingAssignment:
    """What bellman's ProvingAssignment holds after Circuit::synthesize plus
    the prover's extra `input_i * 0 = 0` rows: rows = n_constraints + n_inputs.
    All fields are concatenated 32-byte little-endian canonical scalars."""
    a: bytes
    b: bytes
    c: bytes
    input_assignment: bytes
    aux_assignment: bytes

    @property
    def rows(self):
        return len(self.a) // 32


def create_proof_batch(assignments, params, r_s, s_s):
    """bellperson create_proof_batch(circuits, params, r_s, s_s) after synthesis.
    Returns a list of 192-byte proofs (Proof::write)."""
    n = len(assignments)
    if n == 0:
        return []
    if len(r_s) != n or len(s_s) != n:
        raise ValueError("r_s / s_s length mismatch")
    rows = assignments[0].rows
    for x in assignments:
        if (x.rows != rows or len(x.b) != len(x.a) or len(x.c) != len(x.a)
                or len(x.input_assignment) != 32 * params.n_inputs or len(x.aux_assignment) != 32 * params.n_aux):
            raise ValueError("assignment shape does not match the parameters")
    cat = lambda f: b"".join(f(x) for x in assignments)
    to32 = lambda v: v if isinstance(v, (bytes, bytearray)) else int(v).to_bytes(32, "little")
    out = prove_batch_raw(params, n, rows, cat(lambda x: x.a), cat(lambda x: x.b), cat(lambda x: x.c),
                          cat(lambda x: x.input_assignment), cat(lambda x: x.aux_assignment),
                          b"".join(to32(v) for v in r_s), b"".join(to32(v) for v in s_s))
    return [out[192 * i:192 * (i + 1)] for i in range(n)]


def prove_batch_raw(params, n_proofs, rows, a, b, c, inputs, aux, r, s, device=False, out=None):
    """Thin call of mb200_prove_batch / mb200_prove_batch_device.  Buffers are
    bytes, numpy arrays, torch tensors or raw addresses; with device=True they
    are device pointers.  Returns n_proofs * 192 bytes."""
    _ensure_init()
    L = _lib.lib()
    buf = out if out is not None else ctypes.create_string_buffer(192 * n_proofs)
    outp = _ptr(out) if out is not None else ctypes.cast(buf, ctypes.c_void_p)
    fn = L.mb200_prove_batch_device if device else L.mb200_prove_batch
    keep = [a, b, c, inputs, aux, r, s]
    check(fn(params._h, n_proofs, rows, *[_ptr(x) for x in keep], outp))
    return buf.raw if out is None else out


def prove_submit(params, n_proofs, rows, a, b, c, inputs, aux, r, s, out, device=False):
    """mb200_prove_submit: enqueue a batch, return a ticket.  `out` must be a
    writable buffer of n_proofs * 192 bytes that outlives prove_wait."""
    _ensure_init()
    t = ctypes.c_uint64()
    check(_lib.lib().mb200_prove_submit(params._h, n_proofs, rows, *[_ptr(x) for x in (a, b, c, inputs, aux, r, s)],
                                        int(device), _ptr(out), ctypes.byref(t)))
    return t.value


def prove_submit_witness(params, n_proofs, inputs, aux, r, s, out):
    """Streaming form of create_proof_batch_from_witness: enqueue, return a ticket
    (mb200_prove_submit with no row evaluations: the bound circuit supplies them)."""
    rows = params.circuit.rows
    return prove_submit(params, n_proofs, rows, None, None, None, inputs, aux, r, s, out)


def prove_wait(ticket):
    check(_lib.lib().mb200_prove_wait(ticket))


def create_proof_batch_from_witness(params, inputs, aux, r_s, s_s):
    """create_proof_batch for a key with a bound circuit, from witnesses alone
    (`inputs`, `aux` as Circuit.synthesize returns them): the row evaluations
    bellman computes during synthesis are done on the device."""
    _ensure_init()
    n = len(r_s)
    if n == 0:
        return []
    nbytes = lambda x: x.nbytes if hasattr(x, "nbytes") else len(x)
    if len(s_s) != n or nbytes(inputs) != 32 * n * params.n_inputs or nbytes(aux) != 32 * n * params.n_aux:
        raise ValueError("witness shape does not match the parameters")
    to32 = lambda v: v if isinstance(v, (bytes, bytearray)) else int(v).to_bytes(32, "little")
    out = ctypes.create_string_buffer(192 * n)
    check(_lib.lib().mb200_prove_batch_witness(params._h, n, _ptr(inputs), _ptr(aux),
                                               _ptr(b"".join(to32(v) for v in r_s)),
                                               _ptr(b"".join(to32(v) for v in s_s)),
                                               ctypes.cast(out, ctypes.c_void_p)))
    return [out.raw[192 * i:192 * (i + 1)] for i in range(n)]


def verify_batch(params, proofs_uncompressed, public_inputs):
    """groth16::verify_proof on the device for a batch (mb200_verify_batch).
    proofs_uncompressed: list of 384-byte A | B | C encodings (zkcrypto
    uncompressed); public_inputs: per proof the list of input scalars WITHOUT
    the leading ONE, as verify_proof's `public_inputs` slice.  Returns a list of bools."""
    _ensure_init()
    n = len(proofs_uncompressed)
    if n == 0:
        return []
    inp = b"".join((1).to_bytes(32, "little") + b"".join(int(x).to_bytes(32, "little") for x in xs)
                   for xs in public_inputs)
    if len(inp) != 32 * n * params.n_inputs:
        raise ValueError("public input count does not match the verifying key")
    ok = ctypes.create_string_buffer(n)
    check(_lib.lib().mb200_verify_batch(params._h, n, _ptr(b"".join(proofs_uncompressed)), _ptr(inp),
                                        ctypes.cast(ok, ctypes.c_void_p)))
    return [b != 0 for b in ok.raw]


def verify_proofs(params, proofs, public_inputs):
    """groth16::verify_proof for proofs in wire form (192 bytes each, Proof::read applied on the
    device: decompression + subgroup checks).  public_inputs as for verify_batch.  A proof that
    does not even read is reported False, like any other invalid proof."""
    _ensure_init()
    n = len(proofs)
    if n == 0:
        return []
    if any(len(p) != GROTH_PROOF_SIZE for p in proofs):
        raise ValueError("a proof is %d bytes" % GROTH_PROOF_SIZE)
    inp = b"".join((1).to_bytes(32, "little") + b"".join(int(x).to_bytes(32, "little") for x in xs)
                   for xs in public_inputs)
    if len(inp) != 32 * n * params.n_inputs:
        raise ValueError("public input count does not match the verifying key")
    ok = ctypes.create_string_buffer(n)
    check(_lib.lib().mb200_verify_proofs(params._h, n, _ptr(b"".join(proofs)), _ptr(inp), ctypes.cast(ok, ctypes.c_void_p)))
    return [b != 0 for b in ok.raw]


def verify_proofs_batch(params, proofs, public_inputs, rng=os.urandom):
    """One verdict for a batch of wire-form proofs (bellman's batch::Verifier / the reference's
    BatchValidator): random 128-bit coefficients from `rng(16)` per proof, n + 3 Miller loops
    and one final exponentiation on the device."""
    _ensure_init()
    n = len(proofs)
    if n == 0:
        return True
    if any(len(p) != GROTH_PROOF_SIZE for p in proofs):
        raise ValueError("a proof is %d bytes" % GROTH_PROOF_SIZE)
    inp = b"".join((1).to_bytes(32, "little") + b"".join(int(x).to_bytes(32, "little") for x in xs)
                   for xs in public_inputs)
    if len(inp) != 32 * n * params.n_inputs:
        raise ValueError("public input count does not match the verifying key")
    z = b"".join(rng(16) for _ in range(n))
    ok = ctypes.c_int(0)
    check(_lib.lib().mb200_verify_proofs_batch(params._h, n, _ptr(b"".join(proofs)), _ptr(inp), _ptr(z), ctypes.byref(ok)))
    return ok.value != 0


def create_proof(assignment, params, r, s):
    """bellman create_proof(circuit, params, r, s) after synthesis -> 192 bytes."""
    return create_proof_batch([assignment], params, [r], [s])[0]


def _os_rng_scalar():
    # OsRng + Fr::random: uniform below r by rejection on 255-bit draws
    while True:
        v = int.from_bytes(os.urandom(32), "little") & ((1 << 255) - 1)
        if v < R_ORDER:
            return v


def create_random_proof(assignment, params, rng=_os_rng_scalar):
    """create_random_proof(circuit, params, &mut OsRng): r, s drawn here, as the
    reference does inside spend_proof / output_proof / convert_proof
    (masp_proofs/src/sapling/prover.rs:66, 174, 225)."""
    return create_proof(assignment, params, rng(), rng())


# ---------------------------------------------------------------------------
# load_parameters / LocalTxProver
# ---------------------------------------------------------------------------
class ParameterError(Exception):
    """The reference panics here (masp_proofs/src/lib.rs:290-293, 316-318, 359-388)."""


def masp_circuits():
    """The three recorded circuits (masp_b200.circuits.Circuit), built once per process."""
    global _circuits
    if _circuits is None:
        from . import circuits as C
        _circuits = {"spend": C.Circuit(C.SPEND), "output": C.Circuit(C.OUTPUT), "convert": C.Circuit(C.CONVERT)}
    return _circuits


_circuits = None
_SPEC = {"spend": (MASP_SPEND_BYTES, MASP_SPEND_HASH), "output": (MASP_OUTPUT_BYTES, MASP_OUTPUT_HASH),
         "convert": (MASP_CONVERT_BYTES, MASP_CONVERT_HASH)}


def _load_three(loader, sources, densities, verify):
    """Shared tail of load_parameters / parse_parameters: the size and BLAKE2b-512 semantics live in
    the library (mb200_params_load_file / _verified); a failed check is the reference's panic."""
    bind = densities is None
    if bind:
        circs = masp_circuits()
        densities = {k: c.densities() for k, c in circs.items()}
    dens = densities or {}
    out = {}
    for name in ("spend", "output", "convert"):
        size, digest = _SPEC[name] if verify else (0, "")
        try:
            out[name] = loader(sources[name], size, digest, dens.get(name))
        except Mb200Error as e:
            if e.code in (-9, -10):
                raise ParameterError("MASP %s parameter file is not correct: %s" % (name, e)) from e
            raise
        if bind:
            out[name].bind_circuit(circs[name])
    return out


def parse_parameters(spend_bytes, output_bytes, convert_bytes, densities=None, verify_hashes=True):
    """parse_parameters (masp_proofs/src/lib.rs:330-403): Parameters::read on each
    stream, then the whole stream -- including the MPC transcript after the
    key -- is BLAKE2b-512 hashed and compared with the pinned constants.

    densities=None (the reference's signature has no such argument): the
    bitmaps come from the library's own recorded circuits, which are then bound
    to the keys so that the *_proof methods accept circuit instances."""
    src = {"spend": spend_bytes, "output": output_bytes, "convert": convert_bytes}
    # from bytes the reference checks the digest only (no file size to look at)
    loader = lambda buf, size, digest, d: Parameters.read_verified(buf, 0, digest, d)
    return _load_three(loader, src, densities, verify_hashes)


def load_parameters(spend_path, output_path, convert_path, densities=None, verify=True):
    """load_parameters (masp_proofs/src/lib.rs:278-325): file sizes first, then parse_parameters."""
    src = {"spend": spend_path, "output": output_path, "convert": convert_path}
    if verify:   # all three sizes are checked before any file is opened (lib.rs:284-311)
        for name in ("spend", "output", "convert"):
            size = os.path.getsize(src[name])
            if size != _SPEC[name][0]:
                raise ParameterError("masp %s parameter file size is not correct: %d, expected %d" % (name, size, _SPEC[name][0]))
    return _load_three(Parameters.read_file, src, densities, verify)


def default_params_folder():
    """masp_proofs/src/lib.rs:100-108."""
    return os.path.join(os.path.expanduser("~"), ".masp-params")


class LocalTxProver:
    """masp_proofs::prover::LocalTxProver restricted to its Groth16 work."""

    def __init__(self, spend_params, output_params, convert_params):
        self.spend_params = spend_params
        self.output_params = output_params
        self.convert_params = convert_params

    @classmethod
    def new(cls, spend_path, output_path, convert_path, densities=None, verify=True):
        p = load_parameters(spend_path, output_path, convert_path, densities, verify)
        return cls(p["spend"], p["output"], p["convert"])

    @classmethod
    def from_bytes(cls, spend_param_bytes, output_param_bytes, convert_param_bytes, densities=None,
                   verify_hashes=True):
        p = parse_parameters(spend_param_bytes, output_param_bytes, convert_param_bytes, densities, verify_hashes)
        return cls(p["spend"], p["output"], p["convert"])

    @classmethod
    def with_default_location(cls, densities=None):
        d = default_params_folder()
        paths = [os.path.join(d, n) for n in (MASP_SPEND_NAME, MASP_OUTPUT_NAME, MASP_CONVERT_NAME)]
        if not all(os.path.exists(p) for p in paths):
            return None  # the reference returns None (prover.rs:120-136)
        return cls.new(*paths, densities=densities)

    # Each returns the [u8; 192] zkproof of the description.  `instance` is either the
    # ProvingAssignment a caller synthesised itself, or the circuit instance the reference hands
    # to create_random_proof (masp_b200.circuits.Spend / Output / Convert) when the keys were
    # loaded with the library's own circuits bound (densities=None).
    @staticmethod
    def _prove(instance, params, rng, check):
        if isinstance(instance, ProvingAssignment):
            proof = create_random_proof(instance, params, rng)
            inputs = instance.input_assignment
        else:
            circ = getattr(params, "circuit", None)
            if circ is None:
                raise ValueError("these parameters have no circuit bound; pass a ProvingAssignment")
            inputs, aux = circ.synthesize([instance])
            inputs = bytes(inputs)
            proof = create_proof_batch_from_witness(params, inputs, aux, [rng()], [rng()])[0]
        if check:
            # verify_proof right after proving, as sapling/prover.rs:148 and :266 do -- an explicit call on
            # the returned bytes, not the library-wide "verify" option (which belongs to whoever set it and
            # would race with batches other threads have in flight); failure is the reference's Err(())
            xs = [int.from_bytes(inputs[32 * i:32 * (i + 1)], "little") for i in range(1, params.n_inputs)]
            if not verify_proofs(params, [proof], [xs])[0]:
                raise Mb200Error(-8, "proof does not satisfy the verification equation")
        return proof

    # self_check: the reference always verifies Spend and Convert proofs before returning them
    # (sapling/prover.rs:148, :266); False is for keys that are not a valid CRS (benchmarks on
    # synthetic parameters, like the reference's generate_random_parameters benches)
    def spend_proof(self, instance, rng=_os_rng_scalar, self_check=True):
        return self._prove(instance, self.spend_params, rng, self_check)

    def output_proof(self, instance, rng=_os_rng_scalar):
        return self._prove(instance, self.output_params, rng, False)   # the reference does not self-check outputs

    def convert_proof(self, instance, rng=_os_rng_scalar, self_check=True):
        return self._prove(instance, self.convert_params, rng, self_check)

    def prove_bundle(self, spends=(), converts=(), outputs=(), rng=_os_rng_scalar):
        """All descriptions of a transaction in one launch per circuit (the
        batching TxProver of SURVEY.md §8f-2): returns three lists of proofs."""
        res = []
        for params, group in ((self.spend_params, spends), (self.convert_params, converts),
                              (self.output_params, outputs)):
            group = list(group)
            pairs = [(rng(), rng()) for _ in group]  # r then s for each description, in order
            if group and not isinstance(group[0], ProvingAssignment):
                # circuit instances: one witness pass on the host cores, rows on the device
                inputs, aux = params.circuit.synthesize(group)
                res.append(create_proof_batch_from_witness(params, inputs, aux, [p[0] for p in pairs],
                                                           [p[1] for p in pairs]))
            else:
                res.append(create_proof_batch(group, params, [p[0] for p in pairs], [p[1] for p in pairs]))
        return tuple(res)

    def tx_prover(self, batching=False):
        """`impl TxProver for LocalTxProver` (masp_proofs/src/prover.rs:156-261): the trait object over
        these keys; batching=True defers every proof of a transaction to one launch per circuit."""
        from . import sapling
        return (sapling.BatchingTxProver if batching else sapling.TxProver)(self)


# ---------------------------------------------------------------------------
# standalone pieces of the path
# ---------------------------------------------------------------------------
def msm_g1(bases, scalars, n):
    _ensure_init()
    out = ctypes.create_string_buffer(96)
    check(_lib.lib().mb200_msm_g1(_ptr(bases), _ptr(scalars), n, ctypes.cast(out, ctypes.c_void_p)))
    return out.raw


def msm_g2(bases, scalars, n):
    _ensure_init()
    out = ctypes.create_string_buffer(192)
    check(_lib.lib().mb200_msm_g2(_ptr(bases), _ptr(scalars), n, ctypes.cast(out, ctypes.c_void_p)))
    return out.raw


class G1Bases:
    """Device-resident decoded bases for repeated / split MSMs."""

    def __init__(self, bases, n):
        _ensure_init()
        self.n = n
        self._p = ctypes.c_void_p()
        check(_lib.lib().mb200_g1_bases_upload(_ptr(bases), n, ctypes.byref(self._p)))

    def msm_partial(self, scalars, n=None):
        out = ctypes.create_string_buffer(192)
        check(_lib.lib().mb200_msm_g1_partial(self._p, _ptr(scalars), self.n if n is None else n,
                                              ctypes.cast(out, ctypes.c_void_p)))
        return out.raw

    def msm_partial_into(self, scalars, dev_partial, n=None):
        """Same, the 192-byte XYZZ partial written to device memory (an address or a CUDA
        tensor): the buffer an all-gather sends, no host hop."""
        check(_lib.lib().mb200_msm_g1_partial_device(self._p, _ptr(scalars), self.n if n is None else n,
                                                     _ptr(dev_partial)))

    def __del__(self):
        p = getattr(self, "_p", None)
        if p:
            try:
                _lib.lib().mb200_dev_free(p)
            except Exception:
                pass
            self._p = None


class SplitG1Bases:
    """Bases of one large MSM range-split over every device this process opened
    (BASELINE config 3 in one process): msm() reduces each range on its GPU, gathers the
    192-byte partials GPU -> GPU on the primary device and adds them there."""

    def __init__(self, bases, n):
        _ensure_init()
        self.n = n
        self._h = ctypes.c_void_p()
        check(_lib.lib().mb200_g1_bases_new(_ptr(bases), n, ctypes.byref(self._h)))

    def msm(self, scalars):
        out = ctypes.create_string_buffer(96)
        check(_lib.lib().mb200_msm_g1_bases(self._h, _ptr(scalars), self.n, ctypes.cast(out, ctypes.c_void_p)))
        return out.raw

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().mb200_g1_bases_free(h)
            except Exception:
                pass
            self._h = None


def g1_sum_partials_device(dev_partials, count):
    """Adds `count` XYZZ partials that sit in device memory (e.g. the output of an
    NCCL all-gather) on the device and returns the uncompressed sum."""
    _ensure_init()
    out = ctypes.create_string_buffer(96)
    check(_lib.lib().mb200_g1_sum_partials_device(_ptr(dev_partials), count, ctypes.cast(out, ctypes.c_void_p)))
    return out.raw


def g1_sum_partials(partials):
    _ensure_init()
    buf = b"".join(partials)
    out = ctypes.create_string_buffer(96)
    check(_lib.lib().mb200_g1_sum_partials(_ptr(buf), len(partials), ctypes.cast(out, ctypes.c_void_p)))
    return out.raw


def ntt(data, log_n, inverse=False, coset=False):
    _ensure_init()
    buf = ctypes.create_string_buffer(bytes(data), len(data))
    check(_lib.lib().mb200_ntt(ctypes.cast(buf, ctypes.c_void_p), log_n, int(inverse), int(coset)))
    return buf.raw


def h_coeffs(a, b, c, rows):
    _ensure_init()
    m = 1
    while m < rows:
        m *= 2
    out = ctypes.create_string_buffer(32 * (m - 1))
    check(_lib.lib().mb200_h_coeffs(_ptr(a), _ptr(b), _ptr(c), rows, ctypes.cast(out, ctypes.c_void_p)))
    return out.raw


def fr_mul(a, b, n):
    _ensure_init()
    out = ctypes.create_string_buffer(32 * n)
    check(_lib.lib().mb200_fr_mul(_ptr(a), _ptr(b), n, ctypes.cast(out, ctypes.c_void_p)))
    return out.raw


def fr_mul_device(a_ptr, b_ptr, n, out_ptr):
    _ensure_init()
    check(_lib.lib().mb200_fr_mul_device(_ptr(a_ptr), _ptr(b_ptr), n, _ptr(out_ptr)))


def params_synthesize(shape, seed=None):
    """generate_random_parameters stand-in: Parameters bytes of `shape`
    (masp_b200.synthetic.Shape) with known discrete logs."""
    from .synthetic import MASTER_SEED
    _ensure_init()
    seed = MASTER_SEED if seed is None else seed
    L = _lib.lib()
    n = L.mb200_params_synth_size(shape.n_inputs, shape.h_len, shape.n_aux, shape.a_len, shape.b_len)
    out = ctypes.create_string_buffer(n)
    check(L.mb200_params_synthesize(seed, shape.n_inputs, shape.h_len, shape.n_aux, shape.a_len, shape.b_len,
                                    ctypes.cast(out, ctypes.c_void_p), n))
    return out.raw


def synth_points(stream, start, n, group=1, seed=None):
    from .synthetic import MASTER_SEED
    _ensure_init()
    seed = MASTER_SEED if seed is None else seed
    out = ctypes.create_string_buffer(n * (96 if group == 1 else 192))
    check(_lib.lib().mb200_synth_points(seed, stream, start, n, group, ctypes.cast(out, ctypes.c_void_p)))
    return out.raw


def selftest():
    _ensure_init()
    return _lib.lib().mb200_selftest()


def bench_fpmul():
    _ensure_init()
    v = ctypes.c_double()
    check(_lib.lib().mb200_bench_fpmul(ctypes.byref(v)))
    return v.value


def bench_latency(mode):
    _ensure_init()
    v = ctypes.c_double()
    check(_lib.lib().mb200_bench_latency(mode, ctypes.byref(v)))
    return v.value
