"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end to oracle/_build/libxee_oracle.so, the literal CPU restatement of
the reference hot path (oracle/xee_oracle.hpp cites the reference file:line of each
routine).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product package
(xlab_ee_fortran_b200) never does.

Array convention: a Fortran field f(nx, ny) (i fastest) is a C-order numpy array of
shape (ny, nx); coe(9, nx, ny) is a C-order array of shape (ny, nx, 9).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

DT = {"f32": (np.float32, C.c_float), "f64": (np.float64, C.c_double)}


def build(force: bool = False) -> None:
    """Compile the oracle with its Makefile (g++ only; seconds)."""
    so = os.path.join(_HERE, "_build", "libxee_oracle.so")
    src = [os.path.join(_HERE, n) for n in ("xee_oracle.cpp", "xee_oracle.hpp", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib(variant: str = "O3") -> C.CDLL:
    name = "libxee_oracle.so" if variant == "O3" else "libxee_oracle_O0.so"
    if name not in _LIBS:
        build()
        _LIBS[name] = C.CDLL(os.path.join(_HERE, "_build", name))
    return _LIBS[name]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def sfx(dtype) -> str:
    return "f32" if np.dtype(dtype) == np.float32 else "f64"


def max_threads() -> int:
    return int(lib().xee_oracle_max_threads())


# --------------------------------------------------------------------------- elliptic_tools
def cal_coe(a, b, c, dx, dy, nx, ny, variant="O3"):
    """elliptic_tools.f90:8-60.  a (ny-2,nx-1), b (ny-1,nx-1), c (ny-1,nx-2) -> coe (ny,nx,9)."""
    s = sfx(a.dtype)
    npdt, ct = DT[s]
    a = np.ascontiguousarray(a, npdt); b = np.ascontiguousarray(b, npdt); c = np.ascontiguousarray(c, npdt)
    assert a.shape == (ny - 2, nx - 1) and b.shape == (ny - 1, nx - 1) and c.shape == (ny - 1, nx - 2)
    coe = np.zeros((ny, nx, 9), npdt)
    err = C.c_int(0)
    fn = getattr(lib(variant), f"xee_oracle_cal_coe_{s}")
    fn.restype = None
    fn(_p(a), _p(b), _p(c), _p(coe), ct(dx), ct(dy), C.c_int(nx), C.c_int(ny), C.byref(err))
    return coe, err.value


def do_elliptic(psi, coe, variant="O3"):
    """elliptic_tools.f90:64-90.  Interior only; boundary of the result is 0."""
    s = sfx(psi.dtype)
    npdt, _ = DT[s]
    ny, nx = psi.shape
    psi = np.ascontiguousarray(psi, npdt); coe = np.ascontiguousarray(coe, npdt)
    out = np.zeros((ny, nx), npdt)
    err = C.c_int(0)
    fn = getattr(lib(variant), f"xee_oracle_do_elliptic_{s}")
    fn.restype = None
    fn(_p(psi), _p(coe), _p(out), C.c_int(nx), C.c_int(ny), C.byref(err))
    return out


def solve_elliptic(max_iter, check_step, converge_time, lost_rate, r1, r2, alpha, dat, coe, f,
                   debug=0, quiet=True, trace_cap=0, variant="O3"):
    """elliptic_tools.f90:93-265.  Returns dict(dat, workspace, max_iter, r1, r2, err, rc, trace)."""
    s = sfx(dat.dtype)
    npdt, ct = DT[s]
    ny, nx = dat.shape
    dat = np.array(dat, npdt, order="C", copy=True)
    coe = np.ascontiguousarray(coe, npdt); f = np.ascontiguousarray(f, npdt)
    wk = np.zeros_like(dat)
    mi = C.c_int(max_iter); cr1 = ct(r1); cr2 = ct(r2); err = C.c_int(0); tn = C.c_int(0)
    ti = np.zeros(max(trace_cap, 1), np.int32); te = np.zeros(max(trace_cap, 1)); tr = np.zeros(max(trace_cap, 1))
    fn = getattr(lib(variant), f"xee_oracle_solve_elliptic_{s}")
    fn.restype = C.c_int
    rc = fn(C.byref(mi), C.c_int(check_step), C.c_int(converge_time), C.c_int(lost_rate), C.byref(cr1),
            C.byref(cr2), ct(alpha), _p(dat), _p(coe), _p(f), _p(wk), C.c_int(nx), C.c_int(ny), C.byref(err),
            C.c_int(debug), C.c_int(1 if quiet else 0), C.c_int(trace_cap), C.byref(tn), _p(ti), _p(te), _p(tr))
    n = tn.value
    return dict(dat=dat, workspace=wk, max_iter=mi.value, r1=cr1.value, r2=cr2.value, err=err.value, rc=rc,
                trace=dict(iter=ti[:n].copy(), err_now=te[:n].copy(), ratio=tr[:n].copy()))


def solve_batch(max_iter, check_step, converge_time, lost_rate, r1, r2, alpha, dat, coe, f, threads=None,
                variant="O3"):
    """n independent literal solves, one per host thread.  dat,f (n,ny,nx); coe (ny,nx,9) shared or (n,ny,nx,9)."""
    s = sfx(dat.dtype)
    npdt, ct = DT[s]
    n, ny, nx = dat.shape
    dat = np.array(dat, npdt, order="C", copy=True)
    coe = np.ascontiguousarray(coe, npdt); f = np.ascontiguousarray(f, npdt)
    stride = 0 if coe.ndim == 3 else ny * nx * 9
    mi = np.full(n, max_iter, np.int32)
    a1 = np.full(n, r1, npdt); a2 = np.full(n, r2, npdt); err = np.zeros(n, np.int32)
    threads = threads or max_threads()
    fn = getattr(lib(variant), f"xee_oracle_solve_batch_{s}")
    fn.restype = C.c_double
    sec = fn(C.c_int(n), _p(mi), C.c_int(check_step), C.c_int(converge_time), C.c_int(lost_rate), _p(a1), _p(a2),
             ct(alpha), _p(dat), _p(coe), C.c_longlong(stride), _p(f), C.c_int(nx), C.c_int(ny), _p(err),
             C.c_int(threads))
    return dict(dat=dat, max_iter=mi, r1=a1, r2=a2, err=err, seconds=sec, threads=threads)


def fair_batch(x, coe_planar, f, alpha, sweeps, want_rms=False, threads=None):
    """cpu_fair: fused contiguous sweeps.  x,f (n,ny,nx); coe_planar (9,ny,nx) shared or (n,9,ny,nx)."""
    s = sfx(x.dtype)
    npdt, ct = DT[s]
    n, ny, nx = x.shape
    x = np.array(x, npdt, order="C", copy=True)
    coe_planar = np.ascontiguousarray(coe_planar, npdt); f = np.ascontiguousarray(f, npdt)
    stride = 0 if coe_planar.ndim == 3 else 9 * ny * nx
    rms = np.zeros(n)
    threads = threads or max_threads()
    fn = getattr(lib(), f"xee_oracle_fair_batch_{s}")
    fn.restype = C.c_double
    sec = fn(C.c_int(n), _p(x), _p(coe_planar), C.c_longlong(stride), _p(f), ct(alpha), C.c_int(nx), C.c_int(ny),
             C.c_int(sweeps), _p(rms) if want_rms else None, C.c_int(threads))
    return dict(x=x, rms=rms, seconds=sec, threads=threads)


def judge_error(err: int) -> None:
    lib().xee_oracle_judge_error(C.c_int(err))


# --------------------------------------------------------------------------- driver pieces
class Domain:
    """Domain + modes of src/diagnose/read-input.f90:55-78 (values as the driver holds them)."""

    def __init__(self, Lr, Lz, nr, nz, density_mode=0, geometry=0, planet_radius=0.0):
        self.Lr, self.Lz, self.nr, self.nz = tuple(Lr), tuple(Lz), int(nr), int(nz)
        self.density_mode, self.geometry, self.planet_radius = int(density_mode), int(geometry), float(planet_radius)

    @property
    def dom(self):
        return (C.c_double * 5)(self.Lr[0], self.Lr[1], self.Lz[0], self.Lz[1], self.planet_radius)

    def tail(self):
        return (self.dom, C.c_int(self.nr), C.c_int(self.nz), C.c_int(self.density_mode), C.c_int(self.geometry))


def geometry(d: Domain, dtype):
    """initialize-variables.f90:45-67 -> dict(dr,dz,ra,za,exner,rho,rcuva)."""
    s = sfx(dtype); npdt, ct = DT[s]
    dr = ct(0); dz = ct(0)
    ra = np.zeros(d.nr, npdt); za = np.zeros(d.nz, npdt); ex = np.zeros(d.nz, npdt); rho = np.zeros(d.nz, npdt)
    rc = np.zeros(d.nr, npdt)
    fn = getattr(lib(), f"xee_oracle_geometry_{s}"); fn.restype = None
    fn(*d.tail(), C.byref(dr), C.byref(dz), _p(ra), _p(za), _p(ex), _p(rho), _p(rc))
    return dict(dr=dr.value, dz=dz.value, ra=ra, za=za, exner=ex, rho=rho, rcuva=rc)


def build_abc(A, B, Cc, d: Domain):
    """initialize-variables.f90:72-95.  A,B,C (nz,nr) -> a (nz-2,nr-1), b (nz-1,nr-1), c (nz-1,nr-2)."""
    s = sfx(A.dtype); npdt, _ = DT[s]
    A = np.ascontiguousarray(A, npdt); B = np.ascontiguousarray(B, npdt); Cc = np.ascontiguousarray(Cc, npdt)
    a = np.zeros((d.nz - 2, d.nr - 1), npdt); b = np.zeros((d.nz - 1, d.nr - 1), npdt)
    c = np.zeros((d.nz - 1, d.nr - 2), npdt)
    fn = getattr(lib(), f"xee_oracle_build_abc_{s}"); fn.restype = None
    fn(_p(A), _p(B), _p(Cc), *d.tail(), _p(a), _p(b), _p(c))
    return a, b, c


def cal_eta(rchi, d: Domain):
    """quick-tools1.f90:1-13.  (nz,nr) -> (nz,nr-1)."""
    s = sfx(rchi.dtype); npdt, _ = DT[s]
    rchi = np.ascontiguousarray(rchi, npdt)
    eta = np.zeros((d.nz, d.nr - 1), npdt)
    fn = getattr(lib(), f"xee_oracle_cal_eta_{s}"); fn.restype = None
    fn(_p(rchi), _p(eta), *d.tail())
    return eta


def cal_uw(rpsi, d: Domain):
    """quick-tools1.f90:15-41.  -> u (nz-1,nr), w (nz,nr-1)."""
    s = sfx(rpsi.dtype); npdt, _ = DT[s]
    rpsi = np.ascontiguousarray(rpsi, npdt)
    u = np.zeros((d.nz - 1, d.nr), npdt); w = np.zeros((d.nz, d.nr - 1), npdt)
    fn = getattr(lib(), f"xee_oracle_cal_uw_{s}"); fn.restype = None
    fn(_p(rpsi), _p(u), _p(w), *d.tail())
    return u, w


def integrate_weight_B(wB, d: Domain):
    """old-diagnose/diagnose.f90:1029-1048."""
    s = sfx(wB.dtype); npdt, ct = DT[s]
    wB = np.ascontiguousarray(wB, npdt)
    fn = getattr(lib(), f"xee_oracle_integrate_weight_B_{s}"); fn.restype = ct
    return float(fn(_p(wB), *d.tail()))


def cal_sum_Qeta(Q, eta, d: Domain):
    """old-diagnose/diagnose.f90:1073-1092."""
    s = sfx(Q.dtype); npdt, ct = DT[s]
    Q = np.ascontiguousarray(Q, npdt); eta = np.ascontiguousarray(eta, npdt)
    fn = getattr(lib(), f"xee_oracle_cal_sum_Qeta_{s}"); fn.restype = ct
    return float(fn(_p(Q), _p(eta), *d.tail()))


def cal_wtheta(w_A, theta_B, d: Domain):
    """old-diagnose/diagnose.f90:1117-1127."""
    s = sfx(w_A.dtype); npdt, _ = DT[s]
    w_A = np.ascontiguousarray(w_A, npdt); theta_B = np.ascontiguousarray(theta_B, npdt)
    out = np.zeros((d.nz - 1, d.nr - 1), npdt)
    fn = getattr(lib(), f"xee_oracle_cal_wtheta_{s}"); fn.restype = None
    fn(_p(w_A), _p(theta_B), _p(out), *d.tail())
    return out


def rhs_thermal(Q_B, d: Domain):
    """old-diagnose/diagnose.f90:383-387,396-406 -> (J_B, rhs_O)."""
    s = sfx(Q_B.dtype); npdt, _ = DT[s]
    Q_B = np.ascontiguousarray(Q_B, npdt)
    J = np.zeros((d.nz - 1, d.nr - 1), npdt); rhs = np.zeros((d.nz, d.nr), npdt)
    fn = getattr(lib(), f"xee_oracle_rhs_thermal_{s}"); fn.restype = None
    fn(_p(Q_B), _p(J), _p(rhs), *d.tail())
    return J, rhs


def rhs_momentum(m2_B, F_B, d: Domain):
    """old-diagnose/diagnose.f90:412-436."""
    s = sfx(m2_B.dtype); npdt, _ = DT[s]
    m2_B = np.ascontiguousarray(m2_B, npdt); F_B = np.ascontiguousarray(F_B, npdt)
    rhs = np.zeros((d.nz, d.nr), npdt)
    fn = getattr(lib(), f"xee_oracle_rhs_momentum_{s}"); fn.restype = None
    fn(_p(m2_B), _p(F_B), _p(rhs), *d.tail())
    return rhs


def angular_momentum_sq(rhoC_C, d: Domain):
    """old-diagnose/diagnose.f90:359-367 (intended maths)."""
    s = sfx(rhoC_C.dtype); npdt, _ = DT[s]
    rhoC_C = np.ascontiguousarray(rhoC_C, npdt)
    m2 = np.zeros((d.nz - 1, d.nr - 1), npdt)
    fn = getattr(lib(), f"xee_oracle_angular_momentum_sq_{s}"); fn.restype = None
    fn(_p(rhoC_C), _p(m2), *d.tail())
    return m2


def rhs_from_B(b_B, d: Domain):
    """old-diagnose/diagnose.f90:524-530."""
    s = sfx(b_B.dtype); npdt, _ = DT[s]
    b_B = np.ascontiguousarray(b_B, npdt)
    f = np.zeros((d.nz, d.nr), npdt)
    fn = getattr(lib(), f"xee_oracle_rhs_from_B_{s}"); fn.restype = None
    fn(_p(b_B), _p(f), *d.tail())
    return f


def stagger_averages(A, B, Cc, d: Domain):
    """initialize-variables.f90:100-125 -> rhoA_A (nz,nr-1), rhoB_C (nz-1,nr), rhoB_B (nz-1,nr-1), rhoC_C (nz-1,nr)."""
    s = sfx(A.dtype); npdt, _ = DT[s]
    A = np.ascontiguousarray(A, npdt); B = np.ascontiguousarray(B, npdt); Cc = np.ascontiguousarray(Cc, npdt)
    rA = np.zeros((d.nz, d.nr - 1), npdt); rBC = np.zeros((d.nz - 1, d.nr), npdt)
    rBB = np.zeros((d.nz - 1, d.nr - 1), npdt); rCC = np.zeros((d.nz - 1, d.nr), npdt)
    fn = getattr(lib(), f"xee_oracle_stagger_averages_{s}"); fn.restype = None
    fn(_p(A), _p(B), _p(Cc), _p(rA), _p(rBC), _p(rBB), _p(rCC), *d.tail())
    return rA, rBC, rBB, rCC


def relative_theta(dtheta_dz_A, dtheta_dr_C, d: Domain):
    """old-diagnose/diagnose.f90:893-912."""
    s = sfx(dtheta_dz_A.dtype); npdt, _ = DT[s]
    dz_ = np.ascontiguousarray(dtheta_dz_A, npdt); dr_ = np.ascontiguousarray(dtheta_dr_C, npdt)
    th = np.zeros((d.nz - 1, d.nr - 1), npdt)
    fn = getattr(lib(), f"xee_oracle_relative_theta_{s}"); fn.restype = None
    fn(_p(th), _p(dz_), _p(dr_), *d.tail())
    return th
