"""Opt-in kernel variants that have not been measured on a B200 yet.

They are switched by environment variables read once per process, so each runs the GPU parity
tests again in a child process.  NOT part of the default `-m gpu` run (set MB200_TEST_VARIANTS=1):
a variant that has never executed on the hardware must not be able to stop the suite the
shipped path is judged by.  scripts/gpu_r2_variants.sh runs these and the A/B bench.

  * MB200_ACC_LOCKSTEP=1|2  csrc/msm.cuh: one block per SM, a barrier per bucket-addition iteration
                            (instruction-cache locality), G1 or G1 + G2; 3 = G1 with an L2 prefetch of the next table point;
                            4 = G1 with 512 threads per SM and the accumulators in shared memory.
  * MB200_NTT_SMEM=1|2      csrc/ntt_smem.cuh: the Stockham transform as two shared-memory kernels
                            (2 global passes instead of 6); =2 stores kernel 1's contiguous runs with
                            TMA bulk copies (cp.async.bulk.global.shared::cta); =5 / =7 run the H pipeline with six
                            transforms instead of seven.
  * MB200_ACC_G1_SMEM=1     csrc/msm.cuh: msm_accumulate_g1 with its accumulator in shared memory: 128 registers
                            with 4 bytes of spill, four blocks (16 warps) per SM instead of three.
  * MB200_ACC_G2_SMEM=1     csrc/msm.cuh: msm_accumulate_g2 with its XYZZ<Fp2> accumulator in shared memory
                            (168 registers and 12 warps per SM instead of 255 and 8).
  * MB200_H_SIX=1           csrc/ntt.cuh: the H pipeline with six transforms on the default pass-per-launch NTT.
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
enabled = pytest.mark.skipif(os.environ.get("MB200_TEST_VARIANTS") != "1",
                             reason="opt-in variants: set MB200_TEST_VARIANTS=1")


def _rerun(env_extra, select, timeout=300):
    env = dict(os.environ, **env_extra)
    env.pop("MB200_TEST_VARIANTS", None)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_parity.py"), "-x", "-q", "-m",
                        "gpu", "-k", select], env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout


@enabled
@pytest.mark.gpu
@pytest.mark.parametrize("level", ["1", "2", "3", "4"])
def test_accumulate_lockstep_variant_gpu(level):
    _rerun({"MB200_ACC_LOCKSTEP": level}, "msm or prove or proof")


@enabled
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["1", "2", "5", "7"])
def test_ntt_shared_memory_variant_gpu(mode):
    _rerun({"MB200_NTT_SMEM": mode}, "ntt or h_coefficients or prove")


@enabled
@pytest.mark.gpu
def test_h_six_transforms_variant_gpu():
    _rerun({"MB200_H_SIX": "1"}, "h_coefficients or prove")


@enabled
@pytest.mark.gpu
def test_g2_shared_accumulator_variant_gpu():
    _rerun({"MB200_ACC_G2_SMEM": "1"}, "msm_g2 or prove")


@enabled
@pytest.mark.gpu
def test_g1_shared_accumulator_variant_gpu():
    _rerun({"MB200_ACC_G1_SMEM": "1"}, "msm or prove or proof")
