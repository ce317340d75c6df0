#!/usr/bin/env python
"""Latency of one proof at a time (how the reference's SaplingBuilder calls the prover)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import masp_b200.prover as pv  # noqa: E402
from masp_b200 import synthetic as syn  # noqa: E402

pv.init(0)
for name in sys.argv[1:] or ["output", "convert", "spend"]:
    sh = syn.SHAPES[name]
    P = pv.Parameters.read(pv.params_synthesize(sh), sh.densities())
    w = syn.witness(sh, 0, pv.fr_mul)
    a = pv.ProvingAssignment(w["a"], w["b"], w["c"], w["inputs"], w["aux"])
    wall, dev = [], []
    for i in range(12):  # the first calls pay buffer growth and clock ramp-up; report the settled figures
        t0 = time.perf_counter()
        pv.create_proof(a, P, w["r"], w["s"])
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append(pv.get_counter("last_batch_us") / 1e3)
    wall, dev = sorted(wall[2:]), sorted(dev[2:])
    print(name, "wall min %.1f median %.1f ms" % (wall[0], wall[len(wall) // 2]),
          "device min %.1f median %.1f ms" % (dev[0], dev[len(dev) // 2]), flush=True)
    del P
