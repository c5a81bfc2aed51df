"""CPU checks of the numpy restatement of the planned two-level method (tests/twolevel_oracle.py): Galerkin property,
convergence to the discrete solution, and the gain in Chebyshev sweeps over the block-line method alone on an
anisotropic operator of the secondary-circulation kind (radial coupling dominant)."""
import numpy as np

from oracle import oracle as O
from tests import line_oracle as LO
from tests.twolevel_oracle import TwoLevel


def _operator(nx, ny, seed=3):
    rng = np.random.default_rng(seed)
    a = 40.0 * (1.0 + 0.3 * rng.random((ny - 2, nx - 1))); c = 1.0 + 0.3 * rng.random((ny - 1, nx - 2))
    b = 0.1 * rng.standard_normal((ny - 1, nx - 1))
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    return coe


def test_coarse_correction_is_a_galerkin_projection():
    nx, ny = 96, 48
    coe = _operator(nx, ny)
    tl = TwoLevel(coe, ax=16, az=8)
    rng = np.random.default_rng(0)
    r = np.zeros((ny, nx)); r[1:-1, 1:-1] = rng.standard_normal((ny - 2, nx - 2))
    zc = tl.prolong((tl.Aci @ tl.restrict(r).ravel()).reshape(tl.shape))
    r2 = r - LO.residual(zc, coe, np.zeros((ny, nx)))          # residual after the coarse correction alone
    assert np.abs(tl.restrict(r2)).max() <= 1e-10 * np.abs(tl.restrict(r)).max()
    assert np.all(zc[0] == 0) and np.all(zc[-1] == 0) and np.all(zc[:, 0] == 0) and np.all(zc[:, -1] == 0)


def test_two_level_chebyshev_needs_fewer_sweeps_and_converges_to_the_same_solution():
    nx, ny = 128, 64
    coe = _operator(nx, ny)
    yy, xx = np.mgrid[0:ny, 0:nx]
    f = np.exp(-((xx - 40.0) / 9.0) ** 2 - ((yy - 30.0) / 6.0) ** 2)
    tol = 1e-11 * float(np.sqrt((f[1:-1, 1:-1] ** 2).mean()))
    x0 = np.zeros((ny, nx))
    tl = TwoLevel(coe, ax=32, az=8)
    alpha = 0.8
    rho2 = tl.spectral_radius(alpha)
    x2, k2 = tl.chebyshev(x0, f, alpha, 1.0 - 0.9 * (1.0 - rho2), tol)
    # block-line alone (alpha = 1): the method the CUDA library runs today
    class BlockOnly(TwoLevel):
        def correction(self, r):
            return LO.correction(r, self.coe, self.fac)
    bl = BlockOnly(coe, ax=32, az=8)
    rho1 = bl.spectral_radius(1.0)
    x1, k1 = bl.chebyshev(x0, f, 1.0, 1.0 - 0.9 * (1.0 - rho1), tol)
    assert k2 * 1.5 < k1, (k1, k2)
    assert np.linalg.norm(x2 - x1) <= 1e-8 * np.linalg.norm(x1)
    r = LO.residual(x2, coe, f)
    assert np.sqrt((r[1:-1, 1:-1] ** 2).mean()) < tol
