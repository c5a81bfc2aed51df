"""Host-side mirror of the reference's `module elliptic_tools` (xtt-lib-fortran/elliptic_tools.f90)
over the C-ABI library: same names, argument order and meaning, error behaviour and printed text.

Array convention: the Fortran field f(nx, ny) (i fastest) is a C-order numpy array of shape
(ny, nx); coe(9, nx, ny) is a C-order array of shape (ny, nx, 9) - i.e. exactly the bytes Fortran
would hand to the bind(C) interface.  dtype float32 mirrors the reference's real(4); float64 is the
promoted build.  All compute runs on the GPU (CUDA, sm_100a); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib

# elliptic_tools.f90:3-4
err_over_max_iteration = 1
err_explode = 2


def _sfx(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return "f32", C.c_float
    if dt == np.float64:
        return "f64", C.c_double
    raise TypeError("elliptic_tools: real(4) (float32) or real(8) (float64) arrays required")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fn(name):
    f = getattr(_lib.lib(), name)
    f.restype = None
    return f


def cal_coe(a, b, c, workspace, dx, dy, nx, ny):
    """cal_coe(a,b,c,workspace,dx,dy,nx,ny,err) - elliptic_tools.f90:8-60.

    a (ny-2,nx-1), b (ny-1,nx-1), c (ny-1,nx-2); workspace (ny,nx,9) is updated IN PLACE on its
    interior entries only.  Returns err (0).
    """
    _lib.require_gpu()
    s, ct = _sfx(workspace.dtype)
    dt = workspace.dtype
    a = np.ascontiguousarray(a, dt); b = np.ascontiguousarray(b, dt); c = np.ascontiguousarray(c, dt)
    assert a.shape == (ny - 2, nx - 1) and b.shape == (ny - 1, nx - 1) and c.shape == (ny - 1, nx - 2)
    assert workspace.shape == (ny, nx, 9) and workspace.flags.c_contiguous
    err = C.c_int(0)
    _fn(f"xee_cal_coe_{s}")(_p(a), _p(b), _p(c), _p(workspace), C.byref(ct(dx)), C.byref(ct(dy)),
                            C.byref(C.c_int(nx)), C.byref(C.c_int(ny)), C.byref(err))
    return err.value


def do_elliptic(psi, coe, outdat, nx, ny):
    """do_elliptic(psi,coe,outdat,nx,ny,err) - elliptic_tools.f90:64-90.  outdat interior updated in place."""
    _lib.require_gpu()
    s, _ = _sfx(psi.dtype)
    dt = psi.dtype
    psi = np.ascontiguousarray(psi); coe = np.ascontiguousarray(coe, dt)
    assert outdat.dtype == dt and outdat.flags.c_contiguous and outdat.shape == (ny, nx) == psi.shape
    err = C.c_int(0)
    _fn(f"xee_do_elliptic_{s}")(_p(psi), _p(coe), _p(outdat), C.byref(C.c_int(nx)), C.byref(C.c_int(ny)), C.byref(err))
    return err.value


def solve_elliptic(max_iter, check_step, converge_time, lost_rate, strategy_r1, strategy_r2, alpha, dat, coe, f,
                   workspace, nx, ny, debug=0):
    """solve_elliptic(...) - elliptic_tools.f90:93-265.

    dat (boundary + first guess in, solution out) and workspace are updated IN PLACE.  The Fortran
    intent(inout) scalars come back as a tuple: (max_iter, strategy_r1, strategy_r2, err).
    Both criteria non-positive prints the reference's message and stops (SystemExit, like STOP).
    """
    _lib.require_gpu()
    s, ct = _sfx(dat.dtype)
    dt = dat.dtype
    assert dat.flags.c_contiguous and workspace.flags.c_contiguous and workspace.dtype == dt
    assert dat.shape == (ny, nx) == workspace.shape
    coe = np.ascontiguousarray(coe, dt); f = np.ascontiguousarray(f, dt)
    if not (strategy_r1 > 0) and not (strategy_r2 > 0):      # elliptic_tools.f90:126-129
        print(" ERROR: [check_abs_err] and [check_rel_err] cannot both be non-positive.")
        sys.stdout.flush()
        raise SystemExit(0)
    mi = C.c_int(max_iter); r1 = ct(strategy_r1); r2 = ct(strategy_r2); err = C.c_int(0)
    sys.stdout.flush()
    _fn(f"xee_solve_elliptic_{s}")(C.byref(mi), C.byref(C.c_int(check_step)), C.byref(C.c_int(converge_time)),
                                   C.byref(C.c_int(lost_rate)), C.byref(r1), C.byref(r2), C.byref(ct(alpha)),
                                   _p(dat), _p(coe), _p(f), _p(workspace), C.byref(C.c_int(nx)),
                                   C.byref(C.c_int(ny)), C.byref(err), C.byref(C.c_int(debug)))
    return mi.value, r1.value, r2.value, err.value


def judge_error(err):
    """judge_error(err) - elliptic_tools.f90:333-358."""
    _fn("xee_judge_error")(C.byref(C.c_int(err)))
