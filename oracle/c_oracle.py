"""ctypes binding of oracle/c/liboracle.so (the C++ restatement of the
reference's CPU prover; see oracle/c/oracle.cpp for the citations).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by masp_b200/.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "liboracle.so")
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, "c", f) for f in ("oracle.cpp", "field.hpp")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "c"), "-B", "liboracle.so"],
                          stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        u8p, sz, u32, vp = ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_void_p
        L.orc_params_load.restype = vp
        L.orc_params_load.argtypes = [u8p, sz, u32, u8p, u8p, u8p]
        L.orc_params_free.argtypes = [vp]
        L.orc_params_consumed.restype = sz
        L.orc_params_consumed.argtypes = [vp]
        L.orc_prove.argtypes = [vp, sz, u8p, u8p, u8p, u8p, u8p, u8p, u8p, u8p, ctypes.POINTER(ctypes.c_double)]
        L.orc_h_coeffs.argtypes = [u8p, u8p, u8p, sz, u8p]
        L.orc_ntt.argtypes = [u8p, ctypes.c_uint, ctypes.c_int, ctypes.c_int]
        L.orc_msm_g1.argtypes = [u8p, u8p, sz, u8p]
        L.orc_msm_g2.argtypes = [u8p, u8p, sz, u8p]
        L.orc_g1_gen_mul.argtypes = [u8p, sz, u8p]
        L.orc_g2_gen_mul.argtypes = [u8p, sz, u8p]
        L.orc_fr_mul.argtypes = [u8p, u8p, sz, u8p]
        L.orc_fr_dot.argtypes = [u8p, u8p, sz, u8p]
        L.orc_g1_compress.argtypes = [u8p, u8p]
        L.orc_g2_compress.argtypes = [u8p, u8p]
        L.orc_set_threads.argtypes = [ctypes.c_int]
        L.orc_get_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def set_threads(n):
    lib().orc_set_threads(n)


def get_threads():
    return lib().orc_get_threads()


def _check(rc, what):
    if rc != 0:
        raise ValueError("%s failed: %d" % (what, rc))


class Params:
    def __init__(self, buf, n_aux, a_aux_density=None, b_input_density=None, b_aux_density=None):
        self._h = lib().orc_params_load(buf, len(buf), n_aux, a_aux_density, b_input_density, b_aux_density)
        if not self._h:
            raise ValueError("malformed parameters or densities do not match query lengths")
        self.consumed = lib().orc_params_consumed(self._h)

    def prove(self, rows, a, b, c, inputs, aux, r, s, timings=False):
        out = ctypes.create_string_buffer(192)
        t = (ctypes.c_double * 8)()
        _check(lib().orc_prove(self._h, rows, a, b, c, inputs, aux, r, s, out, t), "orc_prove")
        return (out.raw, list(t)) if timings else out.raw

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_params_free(self._h)
            self._h = None


def h_coeffs(a, b, c, rows):
    m = 1
    while m < rows:
        m *= 2
    out = ctypes.create_string_buffer(32 * (m - 1))
    _check(lib().orc_h_coeffs(a, b, c, rows, out), "orc_h_coeffs")
    return out.raw


def ntt(data, log_n, inverse=False, coset=False):
    buf = ctypes.create_string_buffer(bytes(data), len(data))
    _check(lib().orc_ntt(buf, log_n, int(inverse), int(coset)), "orc_ntt")
    return buf.raw


def msm_g1(bases, scalars, n):
    out = ctypes.create_string_buffer(96)
    _check(lib().orc_msm_g1(bases, scalars, n, out), "orc_msm_g1")
    return out.raw


def msm_g2(bases, scalars, n):
    out = ctypes.create_string_buffer(192)
    _check(lib().orc_msm_g2(bases, scalars, n, out), "orc_msm_g2")
    return out.raw


def g1_gen_mul(scalars, n):
    out = ctypes.create_string_buffer(96 * n)
    _check(lib().orc_g1_gen_mul(scalars, n, out), "orc_g1_gen_mul")
    return out.raw


def g2_gen_mul(scalars, n):
    out = ctypes.create_string_buffer(192 * n)
    _check(lib().orc_g2_gen_mul(scalars, n, out), "orc_g2_gen_mul")
    return out.raw


def fr_mul(a, b, n):
    out = ctypes.create_string_buffer(32 * n)
    _check(lib().orc_fr_mul(a, b, n, out), "orc_fr_mul")
    return out.raw


def fr_dot(a, b, n):
    out = ctypes.create_string_buffer(32)
    _check(lib().orc_fr_dot(a, b, n, out), "orc_fr_dot")
    return int.from_bytes(out.raw, "little")


def g1_compress(p96):
    out = ctypes.create_string_buffer(48)
    _check(lib().orc_g1_compress(p96, out), "orc_g1_compress")
    return out.raw


def g2_compress(p192):
    out = ctypes.create_string_buffer(96)
    _check(lib().orc_g2_compress(p192, out), "orc_g2_compress")
    return out.raw


# ---------------------------------------------------------------------------
# synthetic ("structureless") keys: every query point is a known scalar times
# the generator.  logs: dict of (n,4) uint64 limb arrays from
# masp_b200.synthetic.key_logs (the same derivation the device synthesiser
# uses, so the two byte strings must be equal).
# ---------------------------------------------------------------------------
def params_from_logs(logs):
    import struct
    import numpy as np
    tb = lambda l: np.ascontiguousarray(l.astype("<u8")).tobytes()
    vk = logs["vk"]
    g1 = lambda l: g1_gen_mul(tb(l), len(l))
    g2 = lambda l: g2_gen_mul(tb(l), len(l))
    alpha, beta, gamma, delta = vk[0:1], vk[1:2], vk[2:3], vk[3:4]
    out = [g1(alpha), g1(beta), g2(beta), g2(gamma), g1(delta), g2(delta),
           struct.pack(">I", len(logs["ic"])), g1(logs["ic"])]
    for q in ("h", "l", "a", "b"):
        out += [struct.pack(">I", len(logs[q])), g1(logs[q])]
    out += [struct.pack(">I", len(logs["b"])), g2(logs["b"])]
    return b"".join(out)
