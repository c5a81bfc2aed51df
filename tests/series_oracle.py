"""Oracle-side composition of the time-series chain (BASELINE config 5) from the C++ oracle's routines."""
import numpy as np

from oracle import numpy_ref as N
from oracle import oracle as O
from tests.map_oracle import background_theta, heat_field
from xlab_ee_fortran_b200 import workloads as W


def series_rows(params, nr, nz, Lr, Lz, dt, solve_kw):
    """Returns dict(table [n,8], psi, f, u, w, A, B, C, m2, theta)."""
    d = O.Domain(Lr, Lz, nr, nz, 0, 0)
    g = O.geometry(d, dt)
    k = N.constants(dt)
    out = {q: [] for q in ("table", "psi", "f", "u", "w", "A", "B", "C", "m2", "theta")}
    for row in params:
        A32, B32, C32, bottom, F = W.series_fields_host(row, nr, nz, Lr, Lz)
        A, B, C = A32.astype(dt), B32.astype(dt), C32.astype(dt)
        a, b, c = O.build_abc(A, B, C, d)
        coe, _ = O.cal_coe(a, b, c, g["dr"], g["dz"], nr, nz)
        heat_row = row[14:19]
        Q = heat_field(heat_row, g, dt)
        _, f_thm = O.rhs_thermal(Q, d)
        _, _, _, rhoC_C = O.stagger_averages(A, B, C, d)
        m2 = O.angular_momentum_sq(rhoC_C, d)
        f_mom = O.rhs_momentum(m2, F.astype(dt), d)
        f = f_thm + f_mom
        psi0 = np.zeros((nz, nr), dt); psi0[0, :] = bottom.astype(dt)
        r0 = O.do_elliptic(psi0, coe) - f                                   # tolerance relative to the initial residual
        rms = float(np.sqrt((r0[1:-1, 1:-1].astype(np.float64) ** 2).mean()))
        r = O.solve_elliptic(solve_kw["max_iter"], solve_kw["check_step"], solve_kw["converge_time"], 5,
                             dt(solve_kw["r1_rel"] * rms), 0.0, 1.0, psi0, coe, f)
        u, w = O.cal_uw(r["dat"], d)
        theta = background_theta(A, B, C, d, dt)
        sum_q = O.integrate_weight_B(Q, d)
        ke = O.integrate_weight_B(O.cal_wtheta(w, theta, d), d) * float(k["g0"]) / float(k["theta0"])
        wfin = np.abs(w[np.isfinite(w)]).max(); ufin = np.abs(u[np.isfinite(u)]).max()
        out["table"].append([r["max_iter"], r["r1"], r["err"], sum_q, ke, ke / sum_q, wfin, ufin])
        for q, v in (("psi", r["dat"]), ("f", f), ("u", u), ("w", w), ("A", A), ("B", B), ("C", C), ("m2", m2), ("theta", theta)):
            out[q].append(v)
    return {q: np.array(v) for q, v in out.items()}
