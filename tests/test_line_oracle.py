"""CPU checks of the numpy restatement of the block-line relaxation (tests/line_oracle.py): the partitioned form the
CUDA kernel uses (local 8-point Thomas solves + reduced system + spikes, coupling data rounded to float32) equals the
exact solve of the 32-point block systems up to that rounding, and the iteration converges to the solution of L psi = f."""
import numpy as np

from oracle import oracle as O
from tests import line_oracle as LO
from tests.test_gpu_parity import _rand_case


def test_partitioned_form_equals_exact_block_solve():
    for nx, ny, seed in ((140, 20, 1), (64, 9, 2), (37, 6, 3), (8, 4, 4)):
        a, b, c, f, x0 = _rand_case(nx, ny, np.float64, seed=seed)
        coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
        r = LO.residual(x0, coe, f)
        z = LO.correction(r, coe, LO.factors(coe)); ze = LO.exact_block_correction(r, coe)
        assert np.linalg.norm(z - ze) <= 1e-7 * np.linalg.norm(ze)
        assert np.all(z[0] == 0) and np.all(z[-1] == 0) and np.all(z[:, 0] == 0) and np.all(z[:, -1] == 0)


def test_line_jacobi_converges_to_the_discrete_solution():
    nx, ny = 72, 24
    rng = np.random.default_rng(5)
    a = 30.0 * (1.0 + 0.3 * rng.random((ny - 2, nx - 1))); c = 1.0 + 0.3 * rng.random((ny - 1, nx - 2))
    b = 0.1 * rng.standard_normal((ny - 1, nx - 1))
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    f = rng.standard_normal((ny, nx)); x0 = np.zeros((ny, nx))
    x, rms = LO.line_jacobi(x0, coe, f, 1.0, 1200)
    assert rms < 1e-9 * np.sqrt((f[1:-1, 1:-1] ** 2).mean())
