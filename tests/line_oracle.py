"""CPU restatement (numpy, test infrastructure only) of the segment-line relaxation of csrc/xee_sweep_line.cuh.

The reference has no such method: the residual r = L psi - f is do_elliptic's nine-term sum
(xtt-lib-fortran/elliptic_tools.f90:77-85) minus f, as in solve_elliptic (:189-190); the correction solves
coe4 z(i-1) + coe5 z(i) + coe6 z(i+1) = r(i) on blocks of SEG radial points (global index aligned) instead of the
reference's point-wise r / (-coe5) (:238).  Used to check the CUDA kernel sweep by sweep; the converged solution is
checked against the reference algorithm itself in the tests.
"""
import numpy as np

SEG = 32     # block length of the relaxation (4 threads x 8 points in the CUDA kernel)
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


def factors(coe):
    ny, nx = coe.shape[:2]
    interior = np.zeros((ny, nx), bool); interior[1:-1, 1:-1] = True
    m = np.zeros((ny, nx), coe.dtype); u = np.zeros((ny, nx), coe.dtype)
    for i in range(nx):
        e = i % SEG
        lo = coe[:, i, 3] if e != 0 else 0.0
        up = coe[:, i, 5] if e != SEG - 1 else 0.0
        uprev = u[:, i - 1] if (i > 0 and e != 0) else 0.0
        with np.errstate(divide="ignore", invalid="ignore"):
            mi = 1.0 / (coe[:, i, 4] - lo * uprev)
        m[:, i] = np.where(interior[:, i], mi, 0.0)
        u[:, i] = np.where(interior[:, i], up * m[:, i], 0.0)
    return m, u


def correction(r, coe, m, u):
    ny, nx = r.shape
    y = np.zeros_like(r)
    for i in range(nx):
        prev = y[:, i - 1] if (i % SEG) != 0 else 0.0
        y[:, i] = (r[:, i] - coe[:, i, 3] * prev) * m[:, i]
    z = np.zeros_like(r)
    for i in range(nx - 1, -1, -1):
        nxt = z[:, i + 1] if ((i % SEG) != SEG - 1 and i < nx - 1) else 0.0
        z[:, i] = y[:, i] - u[:, i] * nxt
    return z


def line_jacobi(x0, coe, f, alpha, sweeps):
    """`sweeps` sweeps of psi <- psi - alpha z.  Returns (psi, rms residual seen by the last sweep)."""
    m, u = factors(coe)
    x = x0.copy(); rms = 0.0
    for _ in range(sweeps):
        r = residual(x, coe, f)
        rms = float(np.sqrt((r[1:-1, 1:-1].astype(np.float64) ** 2).mean()))
        x = x - alpha * correction(r, coe, m, u)
    return x, rms
