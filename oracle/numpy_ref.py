"""ORACLE — TEST INFRASTRUCTURE ONLY.

Independent, vectorised numpy restatement of the reference hot path.  It exists to
pin oracle/xee_oracle.hpp (scalar C++ loops) with a second, differently-structured
implementation: numpy ufuncs round every multiply and add separately (no FMA), and
the expressions below keep the reference's left-to-right operation order, so the
iterates agree BIT FOR BIT with the C++ oracle.  tests/golden/make_golden.py runs
this module to produce the committed golden vectors.

Array convention: Fortran f(i, j), i fastest  <->  numpy f[j-1, i-1], shape (ny, nx).
"""
from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.pow.restype = ctypes.c_double; _libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
_libm.powf.restype = ctypes.c_float; _libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]


def _pow(x, y):
    """Elementwise libm pow/powf (what gfortran's ** and std::pow call), so rho is bit-identical to the C++ oracle."""
    fn = _libm.powf if x.dtype == np.float32 else _libm.pow
    return np.array([fn(float(v), float(y)) for v in x], x.dtype)


# xtt-lib-fortran/constants.f90:4-5, evaluated in the working precision
def constants(dt):
    dt = np.dtype(dt).type
    g0 = dt(9.8); theta0 = dt(298.0); Rd = dt(287.0)
    Cv = dt(5.0) / dt(2.0) * Rd
    Cp = Cv + Rd
    kappa = Rd / Cp
    h0 = Cp * theta0 / g0
    p0 = dt(101300.0)
    return dict(g0=g0, theta0=theta0, Rd=Rd, Cv=Cv, Cp=Cp, kappa=kappa, h0=h0, p0=p0)


def cal_coe(a, b, c, dx, dy):
    """elliptic_tools.f90:8-60.  Returns coe of shape (ny, nx, 9); boundary left 0."""
    dt = a.dtype.type
    ny, nx = b.shape[0] + 1, b.shape[1] + 1
    dx = dt(dx); dy = dt(dy)
    PP = dx * dx; QQ = dy * dy; PQ4 = dt(4) * dx * dy
    two_pq4 = dt(2.0) * PQ4
    Ap = a[0:ny - 2, 1:nx - 1] / PP
    Am = a[0:ny - 2, 0:nx - 2] / PP
    Cp = c[1:ny - 1, 0:nx - 2] / QQ
    Cm = c[0:ny - 2, 0:nx - 2] / QQ
    b_ij = b[1:ny - 1, 1:nx - 1]; b_ijm = b[0:ny - 2, 1:nx - 1]
    b_imj = b[1:ny - 1, 0:nx - 2]; b_imjm = b[0:ny - 2, 0:nx - 2]
    BXp = (b_ij + b_ijm) / two_pq4
    BXm = (b_imj + b_imjm) / two_pq4
    BYp = (b_imj + b_ij) / two_pq4
    BYm = (b_imjm + b_ijm) / two_pq4
    coe = np.zeros((ny, nx, 9), a.dtype)
    I = (slice(1, ny - 1), slice(1, nx - 1))
    coe[I + (0,)] = -(BXm + BYp)
    coe[I + (1,)] = Cp + (BXp - BXm)
    coe[I + (2,)] = BXp + BYp
    coe[I + (3,)] = Am - (BYp - BYm)
    coe[I + (4,)] = -(Am + Ap + Cm + Cp)
    coe[I + (5,)] = Ap + (BYp - BYm)
    coe[I + (6,)] = BXm + BYm
    coe[I + (7,)] = Cm - (BXp - BXm)
    coe[I + (8,)] = -(BXp + BYm)
    return coe


def apply_interior(psi, coe):
    """elliptic_tools.f90:75-88 on the interior; slots 1..3 at j+1, 4..6 at j, 7..9 at j-1."""
    ny, nx = psi.shape
    jp, j0, jm = slice(2, ny), slice(1, ny - 1), slice(0, ny - 2)
    im, i0, ip = slice(0, nx - 2), slice(1, nx - 1), slice(2, nx)
    k = coe[1:ny - 1, 1:nx - 1]
    s = k[..., 0] * psi[jp, im]
    s = s + k[..., 1] * psi[jp, i0]
    s = s + k[..., 2] * psi[jp, ip]
    s = s + k[..., 3] * psi[j0, im]
    s = s + k[..., 4] * psi[j0, i0]
    s = s + k[..., 5] * psi[j0, ip]
    s = s + k[..., 6] * psi[jm, im]
    s = s + k[..., 7] * psi[jm, i0]
    s = s + k[..., 8] * psi[jm, ip]
    return s


def do_elliptic(psi, coe):
    out = np.zeros_like(psi)
    out[1:-1, 1:-1] = apply_interior(psi, coe)
    return out


def rms_residual_sequential(r):
    """elliptic_tools.f90:193-199: sequential sum, i outer / j inner, in the working precision."""
    dt = r.dtype.type
    ny, nx = r.shape[0] + 2, r.shape[1] + 2
    sq = (r * r).T.ravel()                       # element order: i outer, j inner
    tot = np.add.accumulate(sq, dtype=r.dtype)[-1]   # accumulate is strictly sequential
    return np.sqrt(tot / dt((nx - 2) * (ny - 2)))


def solve_elliptic(max_iter, check_step, converge_time, lost_rate, r1, r2, alpha, dat, coe, f,
                   snapshots=()):
    """elliptic_tools.f90:93-265.  Returns dict(dat, max_iter, r1, r2, err, trace, snaps)."""
    dt = dat.dtype.type
    r1 = dt(r1); r2 = dt(r2); alpha = dt(alpha)
    huge = np.finfo(dat.dtype).max
    check_abs = r1 > 0
    if not check_abs: r1 = huge
    check_rel = r2 > 0
    if not check_rel: r2 = huge
    if not check_abs and not check_rel:
        raise SystemExit(" ERROR: [check_abs_err] and [check_rel_err] cannot both be non-positive.")
    cs = check_step if check_step > 0 else 100
    ct = converge_time if converge_time > 0 else 10
    lr = lost_rate if lost_rate > 0 else 5
    converge_cnt = 0; lose = 0
    err_before = huge; err = 0
    err_now = dt(0); ratio = dt(0)
    fr = dat.copy(); to = dat.copy()      # workspace = dat; both buffers hold boundary + guess
    negc5 = -coe[1:-1, 1:-1, 4]
    fint = f[1:-1, 1:-1]
    trace = []; snaps = {}
    used = max_iter
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for cnt in range(1, max_iter + 1):
            fr, to = to, fr
            res = apply_interior(fr, coe) - fint
            stop = False
            if cnt % cs == 0:
                err_now = rms_residual_sequential(res)
                ratio = (err_before - err_now) / err_before
                trace.append((cnt, float(err_now), float(ratio)))
                ratio = abs(ratio)
                if err_before == 0:
                    stop = True
                elif err_now < r1 and ratio < r2:
                    converge_cnt += 1; lose = 0
                    if converge_cnt >= ct: stop = True
                elif converge_cnt > 0:
                    lose += 1
                    if lose >= lr:
                        converge_cnt -= 1; lose = 0
                err_before = err_now
            to[1:-1, 1:-1] = fr[1:-1, 1:-1] + alpha * res / negc5
            if cnt in snapshots:
                snaps[cnt] = to.copy()
            if cnt == max_iter:
                stop = True; err |= 1
            if stop:
                used = cnt; r1 = err_now; r2 = ratio
                break
    return dict(dat=to if max_iter > 0 else dat.copy(), max_iter=used, r1=float(r1), r2=float(r2), err=err,
                trace=trace, snaps=snaps)


# ---------------------------------------------------------------- driver (src/diagnose)
def geometry(Lr, Lz, nr, nz, dt, density_mode=0):
    """initialize-variables.f90:45-60 (cylindrical)."""
    dt = np.dtype(dt).type
    k = constants(dt)
    dr = (dt(Lr[1]) - dt(Lr[0])) / dt(nr - 1)
    dz = (dt(Lz[1]) - dt(Lz[0])) / dt(nz - 1)
    ra = dt(Lr[0]) + np.arange(nr).astype(dt) * dr
    za = dt(Lz[0]) + np.arange(nz).astype(dt) * dz
    if density_mode == 0:
        exner = dt(1.0) - za / k["h0"]
        rho = k["p0"] / (k["theta0"] * k["Rd"]) * _pow(exner.astype(dt), dt(1.0) / k["kappa"] - dt(1.0))
    else:
        exner = np.ones(nz, dt); rho = np.ones(nz, dt)
    return dict(dr=dr, dz=dz, ra=ra, za=za, exner=exner.astype(dt), rho=rho.astype(dt), rcuva=ra.copy())


def build_abc(A, B, C, g):
    """initialize-variables.f90:72-95."""
    rc, rho = g["rcuva"], g["rho"]
    rs = (rc[:-1] + rc[1:])[None, :]                   # rcuva(i)+rcuva(i+1), i=1..nr-1
    a = (A[1:-1, :-1] + A[1:-1, 1:]) / rs / rho[1:-1, None]
    rr = (rho[:-1] + rho[1:])[:, None]                 # rho(j)+rho(j+1), j=1..nz-1
    b = (B[:-1, :-1] + B[:-1, 1:] + B[1:, :-1] + B[1:, 1:]) / rs / rr
    c = (C[:-1, 1:-1] + C[1:, 1:-1]) / rc[None, 1:-1] / rr
    return a, b, c


def cal_eta(rchi, g):
    """quick-tools1.f90:1-13 with quick-tools2.f90:59-85."""
    dt = rchi.dtype.type
    k = constants(dt)
    ra, rc, rho, ex = g["ra"], g["rcuva"], g["rho"], g["exner"]
    d = (rchi[:, 1:] - rchi[:, :-1]) / (ra[1:] - ra[:-1])[None, :]
    d = d / ((rc[:-1] + rc[1:]) / dt(2.0))[None, :]
    return d * k["g0"] / (rho * k["Cp"] * ex * k["theta0"])[:, None]


def cal_uw(rpsi, g):
    """quick-tools1.f90:15-41."""
    dt = rpsi.dtype.type
    ra, rc, rho, za = g["ra"], g["rcuva"], g["rho"], g["za"]
    w = (rpsi[:, 1:] - rpsi[:, :-1]) / (ra[1:] - ra[:-1])[None, :]
    w = w / ((rc[:-1] + rc[1:]) / dt(2.0))[None, :]
    w = w / rho[:, None]
    u = -((rpsi[1:, :] - rpsi[:-1, :]) / (za[1:] - za[:-1])[:, None])
    den = rc[None, :] * (rho[:-1] + rho[1:])[:, None] / dt(2.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.where(ra[None, :] != 0, u / den, dt(0.0)).astype(rpsi.dtype)
    return u, w


def integrate_weight_B(w, g):
    """old-diagnose/diagnose.f90:1029-1048, sequential i-outer/j-inner sum."""
    dt = w.dtype.type
    ra, rc, rho, za = g["ra"], g["rcuva"], g["rho"], g["za"]
    rcuv = ((rc[:-1] + rc[1:]) / dt(2.0))[None, :]
    dr = (ra[1:] - ra[:-1])[None, :]
    dz = (za[1:] - za[:-1])[:, None]
    rho_ = ((rho[1:] + rho[:-1]) / dt(2.0))[:, None]
    t = w * rho_ * rcuv * dr * dz
    return np.add.accumulate(t.T.ravel(), dtype=w.dtype)[-1]
