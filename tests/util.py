import random

from masp_b200 import synthetic as syn

R = syn.R_INT
ib = syn.ints_to_bytes


def rand_scalars(n, seed, kinds="mixed"):
    rnd = random.Random(seed)
    if kinds == "uniform":
        return [rnd.randrange(R) for _ in range(n)]
    pool = [lambda: 0, lambda: 1, lambda: rnd.randrange(R), lambda: rnd.randrange(R), lambda: R - 1,
            lambda: rnd.randrange(1 << 16), lambda: 1 << rnd.randrange(255), lambda: 2]
    return [rnd.choice(pool)() for _ in range(n)]


def assignment(pv, w):
    return pv.ProvingAssignment(w["a"], w["b"], w["c"], w["inputs"], w["aux"])


def oracle_proofs(co, key_bytes, shape, dens, ws):
    P = co.Params(key_bytes, shape.n_aux, *dens)
    return [P.prove(shape.rows, w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["r"], w["s"]) for w in ws]


def load_emu():
    """masp_b200.prover bound to the host-compiled device code (tests/emu)."""
    import importlib.util
    from masp_b200 import _lib
    from masp_b200.build import build_emu
    handle = _lib.bind(build_emu())
    spec = importlib.util.find_spec("masp_b200.prover")
    mod = importlib.util.module_from_spec(spec)

    class _EmuLib:
        Mb200Error = _lib.Mb200Error

        @staticmethod
        def lib():
            return handle

        @staticmethod
        def check(rc):
            if rc != 0:
                raise _lib.Mb200Error(rc, handle.mb200_last_error().decode(errors="replace"))
    spec.loader.exec_module(mod)
    mod._lib = _EmuLib
    mod.check = _EmuLib.check
    mod.init()
    # the circuit layer bound to the same build, so handles never cross libraries
    cspec = importlib.util.find_spec("masp_b200.circuits")
    cmod = importlib.util.module_from_spec(cspec)
    cspec.loader.exec_module(cmod)
    cmod._lib = _EmuLib
    cmod.check = _EmuLib.check
    mod.circuits = cmod
    return mod
