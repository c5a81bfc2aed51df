"""Builds (nvcc, sm_100a, in-tree) and loads libxee_b200.so, the C-ABI library of include/xee_b200.h.

There is no CPU fallback anywhere in this package: if the shared library is missing or no CUDA
device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIBDIR = os.path.join(_PKG, "lib")
SO = os.environ.get("XEE_SO") or os.path.join(LIBDIR, "libxee_b200.so")     # XEE_SO: load a variant build (kernel experiments)
HEADER = os.path.join(_ROOT, "include", "xee_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


DIAGNOSE = os.path.join(_PKG, "bin", "xee_diagnose")
OLD_DIAGNOSE = os.path.join(_PKG, "bin", "xee_old_diagnose")


def _build_host_program(src_name: str, exe: str, force: bool) -> str:
    src = os.path.join(CSRC, src_name)
    if force or not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe) or os.path.getmtime(SO) > os.path.getmtime(exe):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-O2", "-std=c++17", "-ffp-contract=off", src, "-L" + LIBDIR, "-lxee_b200", "-Wl,-rpath," + LIBDIR,
               "-Wl,-rpath,$ORIGIN/../lib", "-o", exe]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return exe


def build_diagnose(force: bool = False) -> str:
    """Host-only C++ re-host of the reference driver (csrc/xee_diagnose.cpp), linked against libxee_b200.so."""
    return _build_host_program("xee_diagnose.cpp", DIAGNOSE, force)


def build_old_diagnose(force: bool = False) -> str:
    """C++ re-host of the reference's legacy driver (csrc/xee_old_diagnose.cpp: TENDENCY efficiency decomposition)."""
    return _build_host_program("xee_old_diagnose.cpp", OLD_DIAGNOSE, force)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> xlab_ee_fortran_b200/lib/libxee_b200.so"""
    if force or _stale():
        os.makedirs(LIBDIR, exist_ok=True)
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc, *NVCC_FLAGS, "-o", SO, *sources()]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return SO


_lib = None
_build_error = None


def lib() -> C.CDLL:
    """The loaded C-ABI library.  Raises if it cannot be built/loaded: never a silent fallback."""
    global _lib, _build_error
    if _build_error is not None:          # a failed build stays failed for this process: no minutes-long retry per call
        raise _build_error
    if _lib is None:
        if not os.path.exists(SO):
            if os.environ.get("XEE_NO_BUILD"):
                _build_error = RuntimeError(f"xee_b200: {SO} is missing and XEE_NO_BUILD is set")
                raise _build_error
            try:
                build()
            except Exception as e:
                _build_error = e
                raise
        _lib = C.CDLL(SO)
        _lib.xee_last_error.restype = C.c_char_p
        _lib.xee_build_info.restype = C.c_char_p
        _lib.xee_launch_count.restype = C.c_longlong
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"xee_b200: {what} failed: {lib().xee_last_error().decode()}")


def require_gpu() -> None:
    if lib().xee_device_count() < 1:
        raise RuntimeError("xee_b200: no CUDA device available - this package has no CPU fallback")
