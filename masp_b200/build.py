"""Builds libmasp_b200.so (nvcc, sm_100a only) in-tree.

    python -m masp_b200.build          # product library
    python -m masp_b200.build --emu    # tests/emu/libmasp_b200_emu.so (CPU test aid, never a product path)

Each kernel group is its own translation unit (csrc/k_*.cu) so the units
compile in parallel; objects go to masp_b200/_build/.
"""
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libmasp_b200.so")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libmasp_b200_emu.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]
UNIT_TIMEOUT = int(os.environ.get("MB200_BUILD_TIMEOUT", "1500"))
# host-only units with their own instruction-set flags: the eight-lane witness generator (AVX-512 IFMA);
# csrc/circuits.cu checks the CPU at run time before calling into it
SIMD_FLAGS = ["-mavx512f", "-mavx512ifma", "-mavx512vl", "-mavx512dq", "-mavx512bw"]
CPP_UNITS = {"circuits_simd.cpp": SIMD_FLAGS}


def _units():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    host = os.path.join(CSRC, "host")
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
           [os.path.join(host, f) for f in os.listdir(host) if f.endswith(".hpp")] + \
           [os.path.join(ROOT, "include", "masp_b200.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(cmd, log):
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=UNIT_TIMEOUT)
    with open(log, "wb") as f:
        f.write(p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout.decode(errors="replace")[-6000:])
        raise RuntimeError("build failed: " + " ".join(cmd))
    return time.time() - t0


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdrs = _headers()
    jobs = []
    for u in _units():
        src = os.path.join(CSRC, u)
        obj = os.path.join(OBJ, u[:-3] + ".o")
        if force or _newer(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v", "-c", "-o", obj, src]
            jobs.append((u, cmd, os.path.join(OBJ, u[:-3] + ".log")))
    cxx = os.environ.get("CXX", "g++")
    for u, flags in CPP_UNITS.items():
        src = os.path.join(CSRC, u)
        obj = os.path.join(OBJ, u[:-4] + ".o")
        if force or _newer(obj, [src] + hdrs):
            cmd = [cxx, "-O3", "-std=c++17", "-fPIC", "-pthread"] + flags + ["-c", "-o", obj, src]
            jobs.append((u, cmd, os.path.join(OBJ, u[:-4] + ".log")))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = {u: ex.submit(_compile, cmd, log) for u, cmd, log in jobs}
            for u, f in futs.items():
                dt = f.result()
                if verbose:
                    print("  %-14s %.1fs" % (u, dt))
    objs = [os.path.join(OBJ, u[:-3] + ".o") for u in _units()] + [os.path.join(OBJ, u[:-4] + ".o") for u in CPP_UNITS]
    if jobs or _newer(LIB, objs):
        subprocess.check_call([nvcc] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs)
    return LIB


def build_emu(force=False):
    srcs = [os.path.join(CSRC, u) for u in _units()] + [os.path.join(CSRC, u) for u in CPP_UNITS]
    if not force and not _newer(EMU_LIB, srcs + _headers()):
        return EMU_LIB
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    objdir = os.path.join(OBJ, "emu")
    os.makedirs(objdir, exist_ok=True)
    base = ["g++", "-x", "c++", "-DMB200_EMU", "-O2", "-std=c++17", "-fPIC", "-pthread", "-Wno-unused-function"]

    def one(src):
        name = os.path.basename(src)
        obj = os.path.join(objdir, name.rsplit(".", 1)[0] + ".o")
        subprocess.check_call(base + CPP_UNITS.get(name, []) + ["-c", "-o", obj, src])
        return obj
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, srcs))
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", EMU_LIB] + objs)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force=True))
    else:
        print(build(force="-f" in sys.argv, verbose=True))
