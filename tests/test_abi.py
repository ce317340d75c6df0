"""The C-ABI library: loads, exports every symbol the header declares, refuses
to work without a GPU, and the product package never touches the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from masp_b200.build import build
    return build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "masp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from masp_b200 import _lib
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    for name in header_symbols():
        assert hasattr(L, name), name


def test_library_contains_sm100a_code_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import masp_b200.prover as pv\n"
            "try:\n    pv.init()\n    print('INIT-OK')\n"
            "except pv.Mb200Error as e:\n    print('ERR', e.code)\n"
            "try:\n    pv.fr_mul(bytes(32), bytes(32), 1)\n    print('COMPUTED')\n"
            "except pv.Mb200Error as e:\n    print('ERR', e.code)\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "INIT-OK" not in out and "COMPUTED" not in out
    assert out.count("ERR -3") == 2, out


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "masp_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "oracle/c" not in text, f
                assert "libmasp_b200_emu" not in text or f == "build.py", f


def test_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line
    with the contract's keys; one bounded step of one Spend-shaped proof on this box's cores."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-sample", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-1000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "spend_proofs_per_sec" and line["unit"] == "proofs/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
