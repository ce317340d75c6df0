#!/usr/bin/env python
"""BASELINE config 3: standalone BLS12-381 G1 MSM sweep, bases range-split over the GPUs, the 192-byte
projective partials gathered on a device and added there (no host hop).

  python scripts/msm_sweep.py --sizes 16 18 20 22 24                 # one GPU
  torchrun --nproc-per-node 8 scripts/msm_sweep.py ...                # one rank per GPU, NCCL all-gather of device partials
  python scripts/msm_sweep.py --devices 8 ...                         # ONE process, 8 GPUs behind the C ABI (peer copies)

Bases are PRNG scalars times the generator (known discrete logs), so every result is checked in closed
form against (sum s_i k_i) * G (oracle, rank 0) unless --no-check.  Scalars sit in pinned host memory and
are uploaded inside the timed call.  One JSON line per (size, scalar kind): wall time of the whole call,
device time of the bucket-accumulation kernels, achieved algorithmic GB/s (128 N bytes).
The measurement itself is bench.py's msm_sweep (the same function fills `configs.msm_sweep` of the bench line).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import masp_b200.prover as pv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[16, 18, 20, 22, 24])
    ap.add_argument("--kinds", nargs="+", default=["U", "W"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--devices", type=int, default=1, help="GPUs driven by this one process")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pv.init(list(range(args.devices)) if args.devices > 1 else local)
    peak, _ = bench.measured_peak()
    rows = bench.msm_sweep(pv, torch, dist, rank, world, args.devices, args.sizes, args.kinds, args.reps,
                           check=not args.no_check, emit=lambda m: print(m, file=sys.stderr, flush=True))
    if rank == 0:
        for r in rows:
            r["config"] = "msm_g1_sweep"
            r["limit_ms_1.3x_acc_plus_5"] = 1.3 * r["ms_accumulate_kernel"] + 5.0
            r["within_limit"] = r["ms_total"] <= r["limit_ms_1.3x_acc_plus_5"]
            if r.get("gbs_accumulate_kernel_per_gpu"):
                r["frac_of_hbm_peak_kernel"] = r["gbs_accumulate_kernel_per_gpu"] / peak
            print(json.dumps(r), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
