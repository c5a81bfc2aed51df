"""CPU restatement (numpy, test infrastructure only) of the segment-line relaxation of csrc/xee_sweep_line.cuh.

The reference has no such method: the residual r = L psi - f is do_elliptic's nine-term sum
(xtt-lib-fortran/elliptic_tools.f90:77-85) minus f, as in solve_elliptic (:189-190); the correction solves
coe4 z(i-1) + coe5 z(i) + coe6 z(i+1) = r(i) on blocks of SEG radial points (global index aligned) instead of the
reference's point-wise r / (-coe5) (:238).  Used to check the CUDA kernel sweep by sweep; the converged solution is
checked against the reference algorithm itself in the tests.
"""
import numpy as np

SEG = 8      # points per thread of the CUDA kernel
BLK = 4      # threads per block: the relaxation solves blocks of SEG * BLK = 32 radial points
_OFFS = [(-1, 1), (0, 1), (1, 1), (-1, 0), (0, 0), (1, 0), (-1, -1), (0, -1), (1, -1)]


def residual(x, coe, f):
    """r = L x - f on the interior (0 on the boundary); coe is the reference's (ny, nx, 9) array."""
    ny, nx = x.shape
    r = np.zeros_like(x)
    acc = None
    for k, (di, dj) in enumerate(_OFFS):
        t = coe[1:-1, 1:-1, k] * x[1 + dj:ny - 1 + dj, 1 + di:nx - 1 + di]
        acc = t if acc is None else acc + t
    r[1:-1, 1:-1] = acc - f[1:-1, 1:-1]
    return r


def _seg_factors(coe):
    """Thomas factors m, u of the 8-point segments (0 on boundary points)."""
    ny, nx = coe.shape[:2]
    interior = np.zeros((ny, nx), bool); interior[1:-1, 1:-1] = True
    m = np.zeros((ny, nx)); u = np.zeros((ny, nx))
    for i in range(nx):
        e = i % SEG
        lo = coe[:, i, 3] if e != 0 else 0.0
        up = coe[:, i, 5] if e != SEG - 1 else 0.0
        uprev = u[:, i - 1] if e != 0 else 0.0
        with np.errstate(divide="ignore", invalid="ignore"):
            mi = 1.0 / (coe[:, i, 4] - lo * uprev)
        m[:, i] = np.where(interior[:, i], mi, 0.0)
        u[:, i] = np.where(interior[:, i], up * m[:, i], 0.0)
    return m, u


def _seg_solve(rhs, coe, m, u):
    """Thomas solve of every 8-point segment: forward y(i) = (r(i) - coe4(i) y(i-1)) m(i), back z(i) = y(i) - u(i) z(i+1)."""
    nx = rhs.shape[1]
    y = np.zeros_like(rhs)
    for i in range(nx):
        prev = y[:, i - 1] if (i % SEG) != 0 else 0.0
        y[:, i] = (rhs[:, i] - coe[:, i, 3] * prev) * m[:, i]
    z = np.zeros_like(rhs)
    for i in range(nx - 1, -1, -1):
        nxt = z[:, i + 1] if ((i % SEG) != SEG - 1 and i < nx - 1) else 0.0
        z[:, i] = y[:, i] - u[:, i] * nxt
    return z


def factors(coe):
    """What line_factor_kernel precomputes: segment factors m, u; spikes v, w; and, per segment t of a block, the rows of
    the inverse reduced system that give b(t-1) and a(t+1).  v, w and those rows are rounded to float32 like the kernel's."""
    coe = coe.astype(np.float64)
    ny, nx = coe.shape[:2]
    m, u = _seg_factors(coe)
    nseg = (nx + SEG - 1) // SEG
    interior = np.zeros((ny, nx), bool); interior[1:-1, 1:-1] = True
    ev = np.zeros((ny, nx)); ew = np.zeros((ny, nx))
    for sgm in range(nseg):
        t = sgm % BLK
        i0, i1 = sgm * SEG, sgm * SEG + SEG - 1
        if t > 0 and i0 < nx:
            ev[:, i0] = np.where(interior[:, i0], coe[:, i0, 3], 0.0)
        if t < BLK - 1 and i1 < nx:
            ew[:, i1] = np.where(interior[:, i1], coe[:, i1, 5], 0.0)
    v = _seg_solve(ev, coe, m, u); w = _seg_solve(ew, coe, m, u)
    v = v.astype(np.float32).astype(np.float64); w = w.astype(np.float32).astype(np.float64)   # kernel keeps them in float
    # NOTE: the reduced system is built from the UNROUNDED spikes in the kernel (rounding happens at the store)
    vv = _seg_solve(ev, coe, m, u); ww = _seg_solve(ew, coe, m, u)
    nblk = (nx + SEG * BLK - 1) // (SEG * BLK)
    cB = np.zeros((ny, nseg, 2 * BLK)); cA = np.zeros((ny, nseg, 2 * BLK))     # natural column order (a_0, b_0, a_1, ...)
    get = lambda arr, i: arr[:, i] if i < nx else np.zeros(ny)
    for b in range(nblk):
        R = np.zeros((ny, 2 * BLK, 2 * BLK)); R[:] = np.eye(2 * BLK)
        for t in range(BLK):
            i0 = (b * BLK + t) * SEG
            if t > 0:
                R[:, 2 * t, 2 * (t - 1) + 1] = get(vv, i0); R[:, 2 * t + 1, 2 * (t - 1) + 1] = get(vv, i0 + SEG - 1)
            if t < BLK - 1:
                R[:, 2 * t, 2 * (t + 1)] = get(ww, i0); R[:, 2 * t + 1, 2 * (t + 1)] = get(ww, i0 + SEG - 1)
        Ri = np.linalg.inv(R)
        for t in range(BLK):
            sgm = b * BLK + t
            if sgm >= nseg:
                continue
            if t > 0:
                cB[:, sgm, :] = Ri[:, 2 * (t - 1) + 1, :]
            if t < BLK - 1:
                cA[:, sgm, :] = Ri[:, 2 * (t + 1), :]
    cB = cB.astype(np.float32).astype(np.float64); cA = cA.astype(np.float32).astype(np.float64)
    return dict(m=m, u=u, v=v, w=w, cB=cB, cA=cA)


def correction(r, coe, fac):
    """z = (approximate) solution of the block systems, in the kernel's partitioned form: local segment solves, true end
    values of the neighbouring segments from the reduced system, minus the spikes."""
    ny, nx = r.shape
    z0 = _seg_solve(r, coe.astype(np.float64), fac["m"], fac["u"])
    nseg = (nx + SEG - 1) // SEG
    z = z0.copy()
    zero = np.zeros(ny)
    for sgm in range(nseg):
        b, t = divmod(sgm, BLK)
        g = []
        for s2 in range(BLK):
            i0 = (b * BLK + s2) * SEG
            g.append(z0[:, i0] if i0 < nx else zero)
            g.append(z0[:, i0 + SEG - 1] if i0 + SEG - 1 < nx else zero)
        g = np.stack(g, axis=1)
        bl = (fac["cB"][:, sgm, :] * g).sum(axis=1); ar = (fac["cA"][:, sgm, :] * g).sum(axis=1)
        i0 = sgm * SEG; i1 = min(i0 + SEG, nx)
        z[:, i0:i1] = z0[:, i0:i1] - fac["v"][:, i0:i1] * bl[:, None] - fac["w"][:, i0:i1] * ar[:, None]
    return z


def exact_block_correction(r, coe):
    """The exact solve of the 32-point block systems (what the partitioned form equals up to the float rounding of its
    coupling data)."""
    ny, nx = r.shape
    blk = SEG * BLK
    z = np.zeros_like(r, dtype=np.float64)
    for j in range(1, ny - 1):
        for i0 in range(0, nx, blk):
            idx = [i for i in range(i0, min(i0 + blk, nx)) if 0 < i < nx - 1]
            if not idx:
                continue
            n = len(idx)
            M = np.zeros((n, n))
            for k, i in enumerate(idx):
                M[k, k] = coe[j, i, 4]
                if k > 0: M[k, k - 1] = coe[j, i, 3]
                if k < n - 1: M[k, k + 1] = coe[j, i, 5]
            z[j, idx] = np.linalg.solve(M, r[j, idx])
    return z


def line_jacobi(x0, coe, f, alpha, sweeps):
    """`sweeps` sweeps of psi <- psi - alpha z.  Returns (psi, rms residual seen by the last sweep)."""
    fac = factors(coe)
    x = x0.copy(); rms = 0.0
    for _ in range(sweeps):
        r = residual(x, coe, f)
        rms = float(np.sqrt((r[1:-1, 1:-1].astype(np.float64) ** 2).mean()))
        x = x - alpha * correction(r, coe, fac)
    return x, rms
