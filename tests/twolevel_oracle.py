"""CPU restatement (numpy, test infrastructure only) of the planned TWO-LEVEL extension of the block-line relaxation
(DESIGN.md section 10): the approximate inverse applied to the residual r = L psi - f (do_elliptic's nine-term sum,
xtt-lib-fortran/elliptic_tools.f90:77-85, minus f) becomes

    z = M^-1 r  +  P Ac^-1 P^T r,        psi' = psi - alpha z      (then Chebyshev acceleration)

with M the 32-point radial block systems of tests/line_oracle.py, P bilinear interpolation from a coarse grid with
nodes every (AX, AZ) grid points (Dirichlet boundary: no nodes on it) and Ac = P^T L P the Galerkin coarse operator.
Not in the reference.  The CUDA library implements it as XEE_METHOD_LINE2_* (csrc/xee_twolevel.cuh, nodes every 16 x 16 grid
points); tests/test_gpu_twolevel.py compares the kernels with this restatement sweep by sweep.
"""
import numpy as np

from tests import line_oracle as LO


def hat(n, step):
    """1-D linear interpolation from the coarse nodes at global index step, 2 step, ... (< n-1) to all n points."""
    nodes = [c for c in range(step, n - 1, step)]
    P = np.zeros((n, len(nodes)))
    for k, xc in enumerate(nodes):
        for i in range(max(1, xc - step + 1), min(n - 2, xc + step - 1) + 1):
            P[i, k] = 1.0 - abs(i - xc) / step
    return P


class TwoLevel:
    def __init__(self, coe, ax=16, az=16):
        self.coe = coe.astype(np.float64)
        ny, nx = coe.shape[:2]
        self.Px, self.Pz = hat(nx, ax), hat(ny, az)            # P = Pz (x) Px, applied as Pz @ C @ Px^T
        self.fac = LO.factors(self.coe)
        ncx, ncz = self.Px.shape[1], self.Pz.shape[1]
        Ac = np.zeros((ncz * ncx, ncz * ncx))
        zero = np.zeros((ny, nx))
        for kz in range(ncz):
            for kx in range(ncx):
                basis = np.outer(self.Pz[:, kz], self.Px[:, kx])
                Lb = LO.residual(basis, self.coe, zero)          # L basis (interior), 0 on the boundary
                Ac[:, kz * ncx + kx] = self.restrict(Lb).ravel()
        self.Aci = np.linalg.inv(Ac)
        self.shape = (ncz, ncx)

    def restrict(self, r):
        return self.Pz.T @ r @ self.Px

    def prolong(self, c):
        return self.Pz @ c @ self.Px.T

    def correction(self, r):
        zc = (self.Aci @ self.restrict(r).ravel()).reshape(self.shape)
        return LO.correction(r, self.coe, self.fac) + self.prolong(zc)

    def jacobi(self, x0, f, alpha, sweeps):
        x = x0.copy(); rms = 0.0
        for _ in range(sweeps):
            r = LO.residual(x, self.coe, f)
            rms = float(np.sqrt((r[1:-1, 1:-1] ** 2).mean()))
            x = x - alpha * self.correction(r)
        return x, rms

    def chebyshev(self, x0, f, alpha, rho, tol_rms, max_sweeps=100000):
        """psi+ = omega_k ((psi - alpha z) - psi-) + psi-, omega_k the Chebyshev weights for spectral radius rho of
        I - alpha (M^-1 + P Ac^-1 P^T) L.  Returns (psi, sweeps)."""
        x = x0.copy(); xm = x0.copy()
        sg = 1.0 / rho; q = sg - np.sqrt(sg * sg - 1.0)
        for k in range(1, max_sweeps + 1):
            r = LO.residual(x, self.coe, f)
            if np.sqrt((r[1:-1, 1:-1] ** 2).mean()) < tol_rms:
                return x, k - 1
            om = 1.0 if k == 1 else (2.0 / rho) * q * (1.0 + q ** (2 * (k - 1))) / (1.0 + q ** (2 * (k - 1)) * q * q)
            xj = x - alpha * self.correction(r)
            x, xm = om * (xj - xm) + xm, x
        return x, max_sweeps

    def chebyshev_sweeps(self, x0, f, gamma, rho, sweeps):
        """Exactly `sweeps` Chebyshev sweeps with step gamma and spectral radius rho (the values the library reports through
        xee_plan_cheb_params); the weights in the closed form of cheb_omega (csrc/xee_kernels.cuh)."""
        x = x0.copy(); xm = x0.copy()
        sg = 1.0 / rho; q = sg - np.sqrt(sg * sg - 1.0)
        for k in range(1, sweeps + 1):
            om = 1.0 if k == 1 else (2.0 / rho) * q * (1.0 + q ** (2 * (k - 1))) / (1.0 + q ** (2 * (k - 1)) * q * q)
            r = LO.residual(x, self.coe, f)
            xj = x - gamma * self.correction(r)
            x, xm = om * (xj - xm) + xm, x
        return x

    def spectral_radius(self, alpha, iters=400):
        ny, nx = self.coe.shape[:2]
        e = np.zeros((ny, nx))
        e[1:-1, 1:-1] = np.outer(np.sin(np.pi * np.arange(1, ny - 1) / (ny - 1)), np.sin(np.pi * np.arange(1, nx - 1) / (nx - 1)))
        zero = np.zeros((ny, nx)); lam = 0.0
        for _ in range(iters):
            e2 = e - alpha * self.correction(LO.residual(e, self.coe, zero))
            lam = np.linalg.norm(e2) / np.linalg.norm(e); e = e2 / np.linalg.norm(e2)
        return lam
