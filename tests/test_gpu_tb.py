"""GPU parity tests of the temporally blocked sweep kernel (v4, kernel=4, csrc/xee_sweep_tb.cuh).

Several sweeps of solve_elliptic (xtt-lib-fortran/elliptic_tools.f90:175-248) are done per pass over HBM on
overlapping tiles; the arithmetic per point is unchanged, so STRICT iterates must be BIT-IDENTICAL to the oracle
and FAST iterates bit-identical to the one-sweep-per-launch kernels, for every sweep count, tile-edge position,
blocking depth and stop pattern.
"""
import os

import numpy as np
import pytest

from tests.test_gpu_parity import DTS, RES_TOL, _mods, _rand_case

pytestmark = pytest.mark.gpu


class _env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _batch(nx, ny, nb, dt, seed):
    a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=seed)
    F = np.stack([f * dt(k + 1) for k in range(nb)]); P = np.stack([x0 * dt(1 + 0.25 * k) for k in range(nb)])
    return a, b, c, F, P


@pytest.mark.parametrize("name", ["f32", "f64"])
@pytest.mark.parametrize("shape,nb", [((512, 256), 5), ((140, 70), 20), ((260, 13), 37), ((64, 32), 9), ((68, 40), 3),
                                      ((200, 200), 2), ((8, 4), 3), ((120, 57), 70)])
def test_tb_solve_matches_oracle_bitwise(name, shape, nb):
    """STRICT Jacobi solve through the stop-rule loop: checks every 10 sweeps = passes of 4+4+2 sweeps, max_iter on a
    non-check sweep (57), tiles cut by every domain edge, single-tile grids, more solves than one chunk."""
    torch, X, O = _mods()
    dt = DTS[name]; nx, ny = shape
    a, b, c, F, P = _batch(nx, ny, nb, dt, seed=nx + nb)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    plan = X.Plan(nx, ny, nbatch=nb, dtype=name, shared_coe=True, arith="strict", kernel=4)
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
    out = plan.solve(psi, ft, X.SolveParams(max_iter=57, check_step=10, converge_time=10, r1=1e-30, r2=1.0, alpha=0.9))
    plan.close()
    rb = O.solve_batch(57, 10, 10, 5, 1e-30, 1.0, 0.9, P, coe, F, threads=4)
    assert np.array_equal(psi.cpu().numpy(), rb["dat"]) and list(out["iters"]) == list(rb["max_iter"])
    assert list(out["err"]) == list(rb["err"])
    assert np.allclose(out["r1"], rb["r1"], rtol=RES_TOL[name])


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("sweeps", [1, 2, 3, 4, 5, 7, 8, 9, 16, 23])
def test_tb_fixed_sweeps_every_depth_bitwise(depth, sweeps):
    """Exactly `sweeps` sweeps (no stop rule) for blocking depths 1..8, including short last passes; the RMS residual of
    the last sweep against the oracle's."""
    torch, X, O = _mods()
    nx, ny, nb = 192, 100, 6
    a, b, c, F, P = _batch(nx, ny, nb, np.float64, seed=depth * 100 + sweeps)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    with _env(XEE_TB=depth):
        plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=True, arith="strict", kernel=4)
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
    rms = plan.sweeps(psi, ft, 0.8, sweeps, want_rms=True)
    plan.close()
    got = psi.cpu().numpy()
    for k in range(nb):
        ref = O.solve_elliptic(sweeps, sweeps, 10, 5, 1e-30, 1.0, 0.8, P[k], coe, F[k])
        assert np.array_equal(got[k], ref["dat"]), (k, depth, sweeps)
        assert abs(rms[k] - ref["r1"]) <= 1e-12 * abs(ref["r1"])


@pytest.mark.parametrize("name", ["f32", "f64"])
@pytest.mark.parametrize("shape,nb", [((512, 256), 40), ((256, 128), 33), ((140, 70), 20)])
def test_tb_chebyshev_fast_matches_tma_and_direct_kernels_bitwise(name, shape, nb):
    """FAST + Chebyshev, the bench configuration: v4 runs the same FMA sequence as v1/v2 -> identical bits, identical
    stop sweeps (solves finish at different checks, so finished solves are skipped while others go on), and the same
    `workspace` leftovers."""
    torch, X, O = _mods()
    dt = DTS[name]; nx, ny = shape
    a, b, c, F, P = _batch(nx, ny, nb, dt, seed=7 * nx + nb)
    sc = np.array([10.0 ** (-(k % 4)) for k in range(nb)], dt)[:, None, None]       # different stop sweeps
    F = F * sc; P = P * sc
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    res = {}
    for kern in (1, 2, 4):
        plan = X.Plan(nx, ny, nbatch=nb, dtype=name, shared_coe=True, arith="fast", method="chebyshev", kernel=kern)
        plan.set_coe_aos(coe)
        psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
        prm = X.SolveParams(max_iter=2000, check_step=10, converge_time=2, r1=1e-3 if name == "f64" else 1e-1, r2=0.0, rho_jacobi=0.97)
        out = plan.solve(psi, ft, prm)
        res[kern] = (psi.cpu().numpy(), out)
        plan.close()
    assert len(set(res[4][1]["iters"])) > 1
    for kern in (1, 2):
        assert list(res[kern][1]["iters"]) == list(res[4][1]["iters"])
        assert np.array_equal(res[kern][0], res[4][0])
        assert np.allclose(res[kern][1]["r1"], res[4][1]["r1"], rtol=RES_TOL[name])


@pytest.mark.parametrize("sweeps", [30, 57, 64])
def test_tb_through_the_fortran_facing_entry(sweeps, monkeypatch):
    """solve_elliptic (the drop-in signature) forced onto v4: dat AND workspace as the reference leaves them
    (elliptic_tools.f90:259-264) for even and odd sweep counts."""
    torch, X, O = _mods()
    nx, ny = 136, 90
    a, b, c, f, x0 = _rand_case(nx, ny, np.float64, seed=sweeps)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    monkeypatch.setenv("XEE_KERNEL", "4")
    dat = x0.copy(); wk = np.zeros_like(dat)
    it, r1, r2, err = X.solve_elliptic(sweeps, 10, 10, 5, 1e-30, 1.0, 0.9, dat, coe, f, wk, nx, ny)
    ref = O.solve_elliptic(sweeps, 10, 10, 5, 1e-30, 1.0, 0.9, x0, coe, f)
    assert (it, err) == (ref["max_iter"], ref["err"])
    assert np.array_equal(dat, ref["dat"])
    assert np.array_equal(wk, ref["workspace"])
    assert abs(r1 - ref["r1"]) <= 1e-12 * abs(ref["r1"])


def test_tb_is_the_default_for_large_fast_batches():
    """auto selection: a large shared-operator FAST batch runs on v4 (pass launches < sweeps)."""
    torch, X, O = _mods()
    nx, ny, nb = 256, 128, 128
    a, b, c, F, P = _batch(nx, ny, nb, np.float64, seed=3)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=True, arith="fast")
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
    X.plan.launch_count(reset=True)
    plan.sweeps(psi, ft, 1.0, 40)
    n = X.plan.launch_count(reset=True)
    plan.close()
    assert n == 10, n
