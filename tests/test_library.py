"""The C-ABI library: loads, exports every symbol include/nhans_b200.h declares, and fails loudly (no CPU
fallback) when no B200 is present.  No compute calls here."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from nhans_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nhans_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(nhans_[a-z0-9_]+)\s*\(", hdr)))
    assert declared and set(declared) == set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.SO_PATH)
    for s in declared:
        assert hasattr(lib, s), s


def test_sass_is_blackwell_native():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP"):       # tcgen05.mma, TMA tensor load, tcgen05.ld, 1-D bulk copy
        assert mnemonic in sass, mnemonic
    # per kernel: both tensor-core kernels issue tcgen05.mma and read TMEM; the STFT kernel stages its samples with a bulk copy
    funcs = re.split(r"Function : ", sass)
    def has(kernel, mnemonic):
        return any(kernel in f.split("\n", 1)[0] and mnemonic in f for f in funcs)
    assert has("gemm_shift_kernel", "UTCHMMA") and has("gemm_shift_kernel", ".2CTA") and has("gemm_shift_kernel", "LDTM")
    assert has("conv64_walk_kernel", "UTCHMMA") and has("conv64_walk_kernel", "UTMALDG") and has("conv64_walk_kernel", "LDTM")
    assert has("stft_kernel", "UBLKCP")
    # the committed table is generated from the same disassembly (scripts/sass_table.py)
    assert os.path.exists(os.path.join(ROOT, "profiles", "r02_sass_table.txt"))


def test_no_cpu_fallback():
    """Without a GPU the engine must raise, not compute on the CPU."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from nhans_b200.engine import Engine, NhansError\n"
            "try:\n    Engine(0, 0)\n    print('CREATED')\nexcept NhansError as e:\n    print('RAISED', e)\n" % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env).stdout
    assert "RAISED" in out and "no CPU fallback" in out


def test_product_path_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under nhans_b200/ may import, include or link it."""
    pat_py = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    pat_c = re.compile(r"#include\s+[\"<][^\">]*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nhans_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                assert not pat_py.search(open(path).read()), path
            elif f.endswith((".cu", ".cc", ".h", ".cuh")):
                assert not pat_c.search(open(path).read()), path
            elif f == "Makefile":
                assert "oracle" not in open(path).read(), path
