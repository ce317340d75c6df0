import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from masp_b200 import synthetic as syn
import masp_b200.prover as pv
pv.init(0)
n = 1 << int(sys.argv[1])
mode, kind = sys.argv[2], sys.argv[3]
slab = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bases = pv.synth_points(syn.STREAM_MSM_BASE, 0, n, 1)
gb = pv.G1Bases(bases, n)
sc_all = syn.msm_scalars(n, kind)
raw = np.ascontiguousarray(sc_all).view(np.uint8).reshape(-1)
sc = torch.from_numpy(raw).pin_memory() if mode == "pinned" else raw.tobytes()
mine = torch.zeros(192, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
gb.msm_partial_into(sc, mine)
one = pv.g1_sum_partials_device(mine, 1)
if slab:
    pv.set_option("msm_slab", 1 << slab)
    gb.msm_partial_into(sc, mine)
    many = pv.g1_sum_partials_device(mine, 1)
    print("n 2^%s %s %s slabs of 2^%d: %s" % (sys.argv[1], mode, kind, slab, "SAME" if many == one else "DIFFERENT"), flush=True)
else:
    print("ok", one.hex()[:16])
