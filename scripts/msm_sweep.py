#!/usr/bin/env python
"""BASELINE config 3: standalone BLS12-381 G1 MSM sweep, bases range-split over
the ranks, 192-byte projective partials all-gathered and added.

  python scripts/msm_sweep.py --sizes 16 18 20              # one GPU
  torchrun --nproc-per-node 8 scripts/msm_sweep.py ...      # bases split 8 ways

Bases are PRNG scalars times the generator (known discrete logs), so every
result is checked in closed form against (sum s_i k_i) * G when --check is
given (oracle, rank 0).  Reports per size: wall time of the whole MSM call
with device-resident bases (scalars uploaded inside), the device time of the
bucket-accumulation kernels, and the achieved algorithmic GB/s (128 N bytes).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from masp_b200 import sharding, synthetic as syn  # noqa: E402
import masp_b200.prover as pv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[16, 18, 20, 22])
    ap.add_argument("--kinds", nargs="+", default=["U", "W"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pv.init(local)
    peak = 6547.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for log_n in args.sizes:
        n = 1 << log_n
        lo, hi = sharding.shard_range(n, rank, world)
        bases = pv.synth_points(syn.STREAM_MSM_BASE, lo, hi - lo, 1)   # this rank's range only
        gb = pv.G1Bases(bases, hi - lo)
        del bases
        for kind in args.kinds:
            sc_all = syn.msm_scalars(n, kind)
            sc = syn.limbs_to_bytes(sc_all[lo:hi])
            times, acc_us = [], []
            result = None
            for rep in range(args.reps + 1):
                pv.set_option("profile", 1)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                partial = gb.msm_partial(sc)
                if world > 1:
                    parts = sharding._gather_bytes(partial, torch.device("cuda", local))
                else:
                    parts = [partial]
                result = pv.g1_sum_partials(parts)
                dt = time.perf_counter() - t0
                if world > 1:
                    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                if rep:  # first repetition is warm-up
                    times.append(dt)
                    acc_us.append(pv.get_counter("acc_us"))
                pv.set_option("profile", 0)
            if rank == 0:
                ok = None
                if args.check:
                    from oracle import c_oracle as co
                    logs = syn.fr_uniform(syn.MASTER_SEED, syn.STREAM_MSM_BASE, n)
                    dot = co.fr_dot(syn.limbs_to_bytes(sc_all), syn.limbs_to_bytes(logs), n)
                    ok = result == co.g1_gen_mul(dot.to_bytes(32, "little"), 1)
                best, acc = min(times), min(acc_us) * 1e-6
                print(json.dumps({"config": "msm_g1_sweep", "log_n": log_n, "scalars": kind, "n_gpus": world,
                                  "ms_total": 1e3 * best, "ms_accumulate_kernel": 1e3 * acc,
                                  "gbs_total": 128.0 * n / best / 1e9,
                                  "gbs_accumulate_kernel_per_gpu": 128.0 * (hi - lo) / acc / 1e9 if acc else None,
                                  "frac_of_hbm_peak_kernel": 128.0 * (hi - lo) / acc / 1e9 / peak if acc else None,
                                  "closed_form_ok": ok}), flush=True)
        del gb
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
