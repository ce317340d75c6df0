"""ctypes binding of libmasp_b200.so (include/masp_b200.h).

There is no CPU implementation behind this module: loading fails loudly when
the library has not been built, and mb200_init fails loudly when there is no
sm_100 device.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmasp_b200.so")

SYMBOLS = [
    "mb200_init", "mb200_shutdown", "mb200_device_count", "mb200_params_load", "mb200_params_load_file", "mb200_params_load_verified", "mb200_masp_params_spec", "mb200_blake2b512", "mb200_params_info", "mb200_params_free",
    "mb200_params_synth_size", "mb200_params_synthesize", "mb200_synth_points", "mb200_prove_batch",
    "mb200_prove_batch_device", "mb200_prove_submit", "mb200_prove_wait", "mb200_msm_g1", "mb200_msm_g2", "mb200_g1_bases_upload", "mb200_dev_free",
    "mb200_msm_g1_partial", "mb200_g1_sum_partials", "mb200_msm_g1_partial_device", "mb200_g1_sum_partials_device",
    "mb200_g1_bases_new", "mb200_g1_bases_free", "mb200_msm_g1_bases", "mb200_ntt", "mb200_h_coeffs", "mb200_fr_mul",
    "mb200_fr_mul_device", "mb200_set_option", "mb200_get_counter", "mb200_selftest", "mb200_bench_fpmul", "mb200_bench_latency",
    "mb200_strerror", "mb200_last_error",
    "mb200_circuit_new", "mb200_circuit_free", "mb200_circuit_info", "mb200_circuit_hash", "mb200_circuit_densities",
    "mb200_circuit_matrix", "mb200_circuit_synthesize", "mb200_circuit_simd", "mb200_circuit_rows", "mb200_circuit_root", "mb200_pedersen_hash", "mb200_params_bind_circuit", "mb200_prove_batch_witness", "mb200_verify_batch", "mb200_verify_proofs", "mb200_verify_proofs_batch",
]


class Mb200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("masp_b200 error %d: %s" % (code, msg))
        self.code = code


def bind(path):
    """Load a build of the library and declare every prototype."""
    if not os.path.exists(path):
        raise ImportError(
            "%s is missing: build it with `python -m masp_b200.build` (nvcc, sm_100a). "
            "masp_b200 has no CPU fallback." % path)
    L = ctypes.CDLL(path)
    c = ctypes
    u8p, vp, sz, u32, u64 = c.c_char_p, c.c_void_p, c.c_size_t, c.c_uint32, c.c_uint64
    L.mb200_init.argtypes = [c.POINTER(c.c_int), c.c_int]
    L.mb200_shutdown.argtypes = []
    L.mb200_device_count.argtypes = []
    L.mb200_params_load.argtypes = [u8p, sz, u8p, u8p, u8p, c.POINTER(vp)]
    L.mb200_params_load_file.argtypes = [u8p, u64, u8p, u8p, u8p, u8p, c.POINTER(vp)]
    L.mb200_params_load_verified.argtypes = [u8p, sz, u64, u8p, u8p, u8p, u8p, c.POINTER(vp)]
    L.mb200_masp_params_spec.argtypes = [c.c_int, c.POINTER(u64), vp, c.POINTER(c.c_char_p)]
    L.mb200_blake2b512.argtypes = [vp, sz, vp]
    L.mb200_params_info.argtypes = [vp, c.POINTER(u64)]
    L.mb200_params_free.argtypes = [vp]
    L.mb200_params_free.restype = None
    L.mb200_params_synth_size.argtypes = [u32] * 5
    L.mb200_params_synth_size.restype = sz
    L.mb200_params_synthesize.argtypes = [u64, u32, u32, u32, u32, u32, vp, sz]
    L.mb200_synth_points.argtypes = [u64, u32, u64, sz, c.c_int, vp]
    L.mb200_prove_batch.argtypes = [vp, sz, sz] + [vp] * 7 + [vp]
    L.mb200_prove_batch_device.argtypes = [vp, sz, sz] + [vp] * 7 + [vp]
    L.mb200_prove_submit.argtypes = [vp, sz, sz] + [vp] * 7 + [c.c_int, vp, c.POINTER(u64)]
    L.mb200_prove_wait.argtypes = [u64]
    L.mb200_msm_g1.argtypes = [vp, vp, sz, vp]
    L.mb200_msm_g2.argtypes = [vp, vp, sz, vp]
    L.mb200_g1_bases_upload.argtypes = [vp, sz, c.POINTER(vp)]
    L.mb200_dev_free.argtypes = [vp]
    L.mb200_msm_g1_partial.argtypes = [vp, vp, sz, vp]
    L.mb200_g1_sum_partials.argtypes = [vp, sz, vp]
    L.mb200_msm_g1_partial_device.argtypes = [vp, vp, sz, vp]
    L.mb200_g1_sum_partials_device.argtypes = [vp, sz, vp]
    L.mb200_g1_bases_new.argtypes = [vp, sz, c.POINTER(vp)]
    L.mb200_g1_bases_free.argtypes = [vp]
    L.mb200_g1_bases_free.restype = None
    L.mb200_msm_g1_bases.argtypes = [vp, vp, sz, vp]
    L.mb200_ntt.argtypes = [vp, c.c_uint, c.c_int, c.c_int]
    L.mb200_h_coeffs.argtypes = [vp, vp, vp, sz, vp]
    L.mb200_fr_mul.argtypes = [vp, vp, sz, vp]
    L.mb200_fr_mul_device.argtypes = [vp, vp, sz, vp]
    L.mb200_set_option.argtypes = [u8p, c.c_long]
    L.mb200_get_counter.argtypes = [u8p, c.POINTER(c.c_double)]
    L.mb200_selftest.argtypes = []
    L.mb200_bench_fpmul.argtypes = [c.POINTER(c.c_double)]
    L.mb200_bench_latency.argtypes = [c.c_int, c.POINTER(c.c_double)]
    L.mb200_circuit_new.argtypes = [c.c_int, u32, c.POINTER(vp)]
    L.mb200_circuit_free.argtypes = [vp]
    L.mb200_circuit_free.restype = None
    L.mb200_circuit_info.argtypes = [vp, c.POINTER(u64)]
    L.mb200_circuit_hash.argtypes = [vp, u8p]
    L.mb200_circuit_densities.argtypes = [vp, u8p, u8p, u8p]
    L.mb200_circuit_matrix.argtypes = [vp, c.c_int, vp, vp, vp]
    L.mb200_circuit_synthesize.argtypes = [vp, sz, u8p, u8p, u8p, c.c_int]
    L.mb200_circuit_simd.argtypes = []
    L.mb200_circuit_root.argtypes = [vp, u8p, u8p]
    L.mb200_pedersen_hash.argtypes = [u8p, sz, u8p, u8p]
    L.mb200_circuit_rows.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    L.mb200_params_bind_circuit.argtypes = [vp, vp]
    L.mb200_verify_batch.argtypes = [vp, sz, vp, vp, vp]
    L.mb200_verify_proofs.argtypes = [vp, sz, vp, vp, vp]
    L.mb200_verify_proofs_batch.argtypes = [vp, sz, vp, vp, vp, c.POINTER(c.c_int)]
    L.mb200_prove_batch_witness.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    L.mb200_strerror.argtypes = [c.c_int]
    L.mb200_strerror.restype = u8p
    L.mb200_last_error.argtypes = []
    L.mb200_last_error.restype = u8p
    return L


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = bind(LIB_PATH)
    return _lib


def use_library(handle):
    """Point the package at an already-bound library (tests/emu only)."""
    global _lib
    _lib = handle


def check(rc):
    if rc != 0:
        L = lib()
        detail = L.mb200_last_error().decode(errors="replace")
        raise Mb200Error(rc, detail or L.mb200_strerror(rc).decode())
