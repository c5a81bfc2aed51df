"""Oracle-side composition of the efficiency-map chain from the C++ oracle's routines (CPU, small grids).
Follows src/old-diagnose/diagnose.f90 with testing_dt = 0 (see xlab_ee_fortran_b200/csrc/xee_map.cu header)."""
import numpy as np

from oracle import numpy_ref as N
from oracle import oracle as O


def heat_field(row, g, dt):
    rc, zc, sr, sz, q0 = row
    rm = ((g["ra"][:-1] + g["ra"][1:]) / dt(2)).astype(np.float64)
    zm = ((g["za"][:-1] + g["za"][1:]) / dt(2)).astype(np.float64)
    R, Z = np.meshgrid(rm, zm)
    return (q0 * np.exp(-((R - rc) / sr) ** 2 - ((Z - zc) / sz) ** 2)).astype(dt)


def background_theta(A, B, C, d, dt):
    k = N.constants(dt)
    rA, rBC, rBB, rCC = O.stagger_averages(A, B, C, d)
    rBC = rBC.copy()
    rBC[:, 1:-1] = (rBB[:, :-1] + rBB[:, 1:]) / dt(2)          # old-diagnose/diagnose.f90:503-508
    return O.relative_theta(rA * (k["theta0"] / k["g0"]), rBC * (-(k["theta0"] / k["g0"])), d)


def efficiency_rows(A32, B32, C32, Lr, Lz, heat, dt, solve_kw, adjoint=True, density_mode=0):
    """Returns (table [n,8], psi [n,nz,nr], f, theta, eta)."""
    nz, nr = A32.shape
    d = O.Domain(Lr, Lz, nr, nz, density_mode, 0)
    g = O.geometry(d, dt)
    k = N.constants(dt)
    A, B, C = A32.astype(dt), B32.astype(dt), C32.astype(dt)
    a, b, c = O.build_abc(A, B, C, d)
    coe, _ = O.cal_coe(a, b, c, g["dr"], g["dz"], nr, nz)
    theta = background_theta(A, B, C, d, dt)
    eta = None
    if adjoint:
        _, _, rBB, _ = O.stagger_averages(A, B, C, d)
        fchi = O.rhs_from_B(rBB, d)
        rms = float(np.sqrt((fchi[1:-1, 1:-1].astype(np.float64) ** 2).mean()))
        rc = O.solve_elliptic(solve_kw["max_iter"], solve_kw["check_step"], solve_kw["converge_time"], 5,
                              solve_kw["r1_rel"] * rms, 0.0, 1.0, np.zeros((nz, nr), dt), coe, fchi)
        eta = O.cal_eta(rc["dat"], d)
    rows = []; psis = []; fs = []
    for row in heat:
        Q = heat_field(row, g, dt)
        _, f = O.rhs_thermal(Q, d)
        rms = float(np.sqrt((f[1:-1, 1:-1].astype(np.float64) ** 2).mean()))
        r = O.solve_elliptic(solve_kw["max_iter"], solve_kw["check_step"], solve_kw["converge_time"], 5,
                             dt(solve_kw["r1_rel"] * rms), 0.0, 1.0, np.zeros((nz, nr), dt), coe, f)
        u, w = O.cal_uw(r["dat"], d)
        wth = O.cal_wtheta(w, theta, d)
        sum_q = O.integrate_weight_B(Q, d)
        ke = O.integrate_weight_B(wth, d) * float(k["g0"]) / float(k["theta0"])
        sqe = O.cal_sum_Qeta(Q, eta, d) if adjoint else 0.0
        rows.append([r["max_iter"], r["r1"], r["err"], sum_q, ke, ke / sum_q, sqe, sqe / sum_q])
        psis.append(r["dat"]); fs.append(f)
    return np.array(rows), np.stack(psis), np.stack(fs), theta, eta
