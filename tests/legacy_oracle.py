"""Oracle-side composition of the legacy TENDENCY decomposition (src/old-diagnose/diagnose.f90:238-841) in numpy
(float64) on top of the C++ oracle's solver, with the same stated deviations [D1]-[D6] as
xlab_ee_fortran_b200/csrc/xee_old_diagnose.cpp.  Arrays are [j][i] (Fortran f(i,j) -> f[j-1, i-1])."""
import numpy as np

from oracle import numpy_ref as N
from oracle import oracle as O


def old_solve(strategy, strategy_r, max_iter, alpha, dat, coe, f):
    """Legacy 12-argument solve_elliptic (old-diagnose/xtt-lib/elliptic_tools.f90:93-300), strategies 1 and 2."""
    eff = (max_iter // 100) * 100
    if strategy == 1:
        r = O.solve_elliptic(eff, 100, 1, 5, strategy_r, 0.0, alpha, dat, coe, f)
    else:
        r = O.solve_elliptic(eff, 100, 10, 5, 0.0, strategy_r, alpha, dat, coe, f)
    return r["dat"], r["max_iter"], r["r1"]


def old_solve_loop(strategy, strategy_r, max_iter, alpha, dat, coe, f):
    """The legacy loop itself (old-diagnose/xtt-lib/elliptic_tools.f90:168-300) for strategies 1..4, in the working precision
    of `dat`: strategies 3 / 4 measure err_now = maxval(abs(to_dat)) (:203-204) over the WHOLE array - residual in the interior,
    Dirichlet values on the rim.  Returns dict(dat, strategy (sweeps used, or the input when the loop falls through),
    strategy_r, err)."""
    dt = dat.dtype.type
    sr = dt(strategy_r); alpha = dt(alpha)
    huge = np.finfo(dat.dtype).max
    err_before = huge; err_now = dt(0); err = 0
    conv = 0; lose = 0
    fr = dat.copy(); to = dat.copy()                                    # :160-165  workspace = dat
    negc5 = -coe[1:-1, 1:-1, 4]; fint = f[1:-1, 1:-1]
    out_strategy, out_r = strategy, sr
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for cnt in range(1, max_iter + 1):
            flag = cnt % 100 == 0                                       # :170-174
            fr, to = to, fr
            res = N.apply_interior(fr, coe) - fint                      # :180-181
            stop = False
            if flag:
                if strategy in (1, 2):
                    err_now = N.rms_residual_sequential(res)            # :193-201
                else:
                    whole = to.copy(); whole[1:-1, 1:-1] = res
                    err_now = np.abs(whole).max()                       # :203-204
                ratio = (err_before - err_now) / err_before             # :208
            to[1:-1, 1:-1] = fr[1:-1, 1:-1] + alpha * res / negc5       # :241
            if flag:
                if strategy in (1, 3):
                    if err_now < sr: stop = True                        # :246-249
                else:
                    if err_before == 0: stop = True
                    elif abs(ratio) < sr:
                        conv += 1; lose = 0
                        if conv >= 10: stop = True
                    elif conv > 0:
                        lose += 1
                        if lose >= 5: conv -= 1; lose = 0
                    err_before = err_now                                # :279
                if cnt == max_iter: stop = True; err |= 1               # :281-287
                if stop:
                    out_strategy, out_r = cnt, err_now                  # :292
                    break
    return dict(dat=to, strategy=out_strategy, strategy_r=float(out_r), err=err)


def exchange_conversion(rpsi, rchi, rhoC, g):
    """cal_exchange_conversion (old-diagnose/diagnose.f90:1143-1174): the top/bottom boundary exchange term
    (rhoC/rho) (psi d(chi)/dz - chi d(psi)/dz) / r^2 at the mid-points of the first and last row, and its integral
    sum = -sum_i (top - bottom) r dr.  Deviation [D5]: r, dr, dz are REAL here (the legacy code declares them INTEGER,
    :1146, which truncates r to whole metres and makes dz = za(2)-za(1) an integer).  Returns (bndconv [2][nr-1], sum)."""
    ra, za, rho = g["ra"], g["za"], g["rho"]
    dz = za[1] - za[0]; dr = ra[1] - ra[0]
    r = (ra[:-1] + ra[1:]) / 2.0
    pair = lambda x, j: x[j, :-1] + x[j, 1:]
    def row(jb, jin, sign):     # jb = boundary row, jin = its inner neighbour; sign: +1 bottom (inner - boundary), top (boundary - inner)
        dchi = sign * (pair(rchi, jin) - pair(rchi, jb)) / (2.0 * dz)
        dpsi = sign * (pair(rpsi, jin) - pair(rpsi, jb)) / (2.0 * dz)
        return (pair(rhoC, jb) / (2.0 * rho[jb])) * ((pair(rpsi, jb) / 2.0) * dchi - (pair(rchi, jb) / 2.0) * dpsi) / r ** 2.0
    nz = rpsi.shape[0]
    bottom = row(0, 1, 1.0); top = row(nz - 1, nz - 2, -1.0)
    total = 0.0
    for i in range(len(r)):                 # the reference accumulates sequentially (:1172)
        total = total - (top[i] - bottom[i]) * r[i] * dr
    return np.stack([bottom, top]), total


def decompose(A, B, C, Q, F, Lr, Lz, testing_dt, rpsi_set, rchi_set, baro=2, rchi_bc=None, rpsi_bc=None, tendency=True):
    """rpsi_set / rchi_set = (strategy, strategy_r, max_iter, alpha).  Returns dict of sums and fields."""
    dt = np.float64
    nz, nr = A.shape
    d = O.Domain(Lr, Lz, nr, nz, 0, 0)
    g = O.geometry(d, dt)
    k = N.constants(dt)
    g0, th0, Cp = k["g0"], k["theta0"], k["Cp"]
    ra, za, rho, ex = g["ra"], g["za"], g["rho"], g["exner"]
    A, B, C, Q, F = (np.asarray(x, dt) for x in (A, B, C, Q, F))
    intB = lambda w: O.integrate_weight_B(np.ascontiguousarray(w), d)
    sum_Q = intB(Q)
    a, b_basic_s, c = O.build_abc(A, B, C, d)
    rA, rBC, rBB, rCC = O.stagger_averages(A, B, C, d)
    b_basic = rBB.copy()
    m2 = O.angular_momentum_sq(rCC, d)
    J, rhs_thm = O.rhs_thermal(Q, d)
    rhs_mom = O.rhs_momentum(m2, F, d)
    out = dict(sum_Q=sum_Q)
    coe_of = lambda bb: O.cal_coe(a, bb, c, g["dr"], g["dz"], nr, nz)[0]
    b_anom = np.zeros_like(rBB); sb_anom = np.zeros_like(rBB)
    if tendency:
        psi0 = np.zeros((nz, nr)) if rpsi_bc is None else rpsi_bc.astype(dt)
        rpsi, it, res = old_solve(*rpsi_set[:2], rpsi_set[2], rpsi_set[3], psi0, coe_of(b_basic_s), rhs_thm + rhs_mom)
        out["rpsi_before"] = rpsi
        u, w = O.cal_uw(rpsi, d)
        th = J - th0 / g0 * (rA[:-1] * w[:-1] + rA[1:] * w[1:]) / 2.0 + th0 / g0 * (rBC[:, :-1] * u[:, :-1] + rBC[:, 1:] * u[:, 1:]) / 2.0
        out["dtheta_dt"] = th.copy(); out["sum_dtheta_dt"] = intB(th)
        th = th * testing_dt
        dB = np.zeros_like(th)                                       # d_dr_B2B
        dB[:, 1:-1] = (th[:, :-2] - th[:, 2:]) / (ra[:-3] - ra[2:-1])[None, :]
        dB[:, 0] = (th[:, 0] - th[:, 1]) / (ra[0] - ra[1]); dB[:, -1] = (th[:, -2] - th[:, -1]) / (ra[-3] - ra[-2])
        b_anom = -g0 / th0 * dB
        rBB = rBB + b_anom
        dA = np.zeros((nz, nr - 1))                                   # d_dz_B2A on rows 2..nz-2, [D3] 0 elsewhere
        dA[1:nz - 2] = (th[1:nz - 2] - th[0:nz - 3]) / ((za[2:nz - 1] - za[0:nz - 3]) / 2.0)[:, None]
        rA = rA.copy(); rA[1:nz - 1] = rA[1:nz - 1] + g0 / th0 * dA[1:nz - 1]
    rBC = rBC.copy(); rBC[:, 1:-1] = (rBB[:, :-1] + rBB[:, 1:]) / 2.0
    theta = O.relative_theta(rA * (th0 / g0), rBC * (-th0 / g0), d)
    out["theta_after"] = theta
    rs = ((ra[:-1] + ra[1:]) / 2.0)[None, :]; rr = ((rho[:-1] + rho[1:]) / 2.0)[:, None]
    sb_anom = b_anom / rs / rr
    f_basic = O.rhs_from_B(b_basic, d); f_anom = O.rhs_from_B(b_anom, d)
    ops = {}
    if baro in (0, 2): ops["0"] = coe_of(np.zeros_like(b_basic_s))
    if baro in (1, 2): ops["B0dB"] = coe_of(b_basic_s + sb_anom)
    def eta_sum(chi):
        eta = O.cal_eta(chi, d)
        return O.cal_sum_Qeta(Q, eta, d)
    order = []
    rchi = np.zeros((nz, nr))
    if rchi_bc is not None:
        rchi = rchi_bc.astype(dt)
        for tag in ("0", "B0dB"):
            if tag in ops: order.append((tag, None, f"{tag}_0"))
    seq2 = []
    for rhs, name in ((f_anom, "dB"), (f_basic, "B0")):
        for tag in ("0", "B0dB"):
            if tag in ops: seq2.append((tag, rhs, f"{tag}_{name}"))
    for tag, rhs, name in order:
        rchi, it, res = old_solve(*rchi_set[:2], rchi_set[2], rchi_set[3], rchi, ops[tag], np.zeros((nz, nr)))
        out[f"rchi_{name}"] = rchi; out[f"sum_Qeta_{name}"] = eta_sum(rchi)
    rchi = np.zeros((nz, nr))
    for tag, rhs, name in seq2:
        rchi, it, res = old_solve(*rchi_set[:2], rchi_set[2], rchi_set[3], rchi, ops[tag], rhs)
        out[f"rchi_{name}"] = rchi; out[f"sum_Qeta_{name}"] = eta_sum(rchi)
    rpsi = np.zeros((nz, nr)) if rpsi_bc is None else rpsi_bc.astype(dt)
    for tag in ("0", "B0dB"):
        if tag not in ops: continue
        rpsi, it, res = old_solve(*rpsi_set[:2], rpsi_set[2], rpsi_set[3], rpsi, ops[tag], rhs_thm + rhs_mom)
        u, w = O.cal_uw(rpsi, d)
        out[f"rpsi_after_{tag}"] = rpsi
        out[f"sum_wtheta_{tag}"] = intB(O.cal_wtheta(w, theta, d)) * g0 / th0
    if rchi_bc is not None:
        # exchange conversion :730-772: the driver re-reads the float32 files it wrote, sums the chi fields in working precision
        f32 = lambda x: np.asarray(x, np.float32).astype(dt)
        for tag in ops:
            psi_ = f32(out[f"rpsi_after_{tag}"])
            chi3 = f32(out[f"rchi_{tag}_0"]) + f32(out[f"rchi_{tag}_dB"]) + f32(out[f"rchi_{tag}_B0"])
            chi2 = f32(out[f"rchi_{tag}_dB"]) + f32(out[f"rchi_{tag}_B0"])
            out[f"bndconv_{tag}"], out[f"sum_bndconv_{tag}"] = exchange_conversion(psi_, chi3, C, g)      # method 1
            out[f"bndconv2_{tag}"], out[f"sum_bndconv2_{tag}"] = exchange_conversion(psi_, chi2, C, g)    # method 2
    return out
