"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/xee_b200.h declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "xee_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xee_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from xlab_ee_fortran_b200 import _lib
    _lib.build()
    L = ctypes.CDLL(_lib.SO)
    syms = _declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_library_is_sm100a_and_has_no_oracle_dependency():
    from xlab_ee_fortran_b200 import _lib
    _lib.build()
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", _lib.SO], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", _lib.SO], capture_output=True, text=True).stdout
    assert "oracle" not in ldd
    # product sources never reference the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xlab_ee_fortran_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, fn)


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import xlab_ee_fortran_b200 as X
    with pytest.raises(RuntimeError, match="no CUDA device"):
        X.Plan(16, 16)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        X.cal_coe(np.ones((2, 3)), np.ones((3, 3)), np.ones((3, 2)), np.zeros((4, 4, 9)), 1.0, 1.0, 4, 4)


def test_both_criteria_nonpositive_stops_like_fortran():
    """elliptic_tools.f90:126-129: message + STOP.  The Python mirror raises SystemExit(0) after the same text."""
    code = ("import numpy as np, xlab_ee_fortran_b200 as X\n"
            "X._lib.require_gpu = lambda: None\n"
            "z=np.zeros((4,4)); X.solve_elliptic(10,1,1,1,0.0,-1.0,1.0,z,np.zeros((4,4,9)),z,z.copy(),4,4)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0
    assert "ERROR: [check_abs_err] and [check_rel_err] cannot both be non-positive." in r.stdout
