"""GPU tests of the two-level block-line methods (XEE_METHOD_LINE2_*, csrc/xee_twolevel.cuh + the TWO variant of
csrc/xee_sweep_line.cuh): block-line relaxation plus a Galerkin coarse-grid correction on a 16 x 16-spaced bilinear space.

Not the reference's iteration, so no bit-identity: (1) sweep by sweep against the numpy restatement of the same iteration
(tests/twolevel_oracle.py) to rounding; (2) the CONVERGED solution against the reference algorithm (solve_elliptic,
xtt-lib-fortran/elliptic_tools.f90:93-265, STRICT arithmetic = bit-identical to the oracle) within the north_star tolerance
of 1e-8 relative L2, plus an independent residual check through do_elliptic; (3) the efficiency map against the one-level
method within 1e-6 relative.
"""
import numpy as np
import pytest

from tests.test_gpu_line import _batch
from tests.test_gpu_parity import _mods
from tests.twolevel_oracle import TwoLevel
from tests.util import rel_l2

pytestmark = pytest.mark.gpu

SHAPES = [((512, 256), 2), ((200, 200), 2), ((140, 70), 5), ((72, 40), 9), ((64, 48), 3), ((34, 34), 2), ((128, 40), 33)]


def _aniso(nx, ny, seed=11):
    """Radial coupling dominant, like the secondary-circulation operator."""
    rng = np.random.default_rng(seed)
    a = 40.0 * (1.0 + 0.3 * rng.random((ny - 2, nx - 1))); c = 1.0 + 0.3 * rng.random((ny - 1, nx - 2))
    b = 0.2 * rng.standard_normal((ny - 1, nx - 1))
    return a, b, c


@pytest.mark.parametrize("shape,nb", SHAPES)
@pytest.mark.parametrize("sweeps", [1, 3])
def test_line2_jacobi_sweeps_match_numpy_restatement(shape, nb, sweeps):
    torch, X, O = _mods()
    nx, ny = shape
    a, b, c, F, P = _batch(nx, ny, nb, np.float64, seed=nx + sweeps)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=True, arith="fast", method="line2_jacobi")
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
    rms = plan.sweeps(psi, ft, 0.45, sweeps, want_rms=True)
    assert plan.kernel_info()[0] == 5
    plan.close()
    got = psi.cpu().numpy()
    tl = TwoLevel(coe)
    for k in range(min(nb, 3)):
        ref, rref = tl.jacobi(P[k], F[k], 0.45, sweeps)
        assert rel_l2(got[k], ref) < 1e-11, (k, rel_l2(got[k], ref))
        assert np.array_equal(got[k][0], P[k][0]) and np.array_equal(got[k][:, -1], P[k][:, -1])   # Dirichlet values untouched
        assert np.array_equal(got[k][-1], P[k][-1]) and np.array_equal(got[k][:, 0], P[k][:, 0])
        assert abs(rms[k] - rref) <= 1e-10 * rref


@pytest.mark.parametrize("shape,nb", [((256, 128), 3), ((200, 200), 2)])
def test_line2_chebyshev_sweeps_match_numpy_restatement(shape, nb):
    """Chebyshev-accelerated two-level sweeps with the step gamma and spectral radius rho the library estimated."""
    torch, X, O = _mods()
    nx, ny = shape
    a, b, c = _aniso(nx, ny)
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    rng = np.random.default_rng(3)
    F = rng.standard_normal((nb, ny, nx)); P = np.zeros((nb, ny, nx))
    plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=True, arith="fast", method="line2_chebyshev")
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
    plan.sweeps(psi, ft, 1.0, 6)
    rho, gamma = plan.cheb_params()
    plan.close()
    assert 0.5 < rho < 1.0 and 0.3 < gamma < 1.2, (rho, gamma)
    got = psi.cpu().numpy()
    tl = TwoLevel(coe)
    for k in range(nb):
        ref = tl.chebyshev_sweeps(P[k], F[k], gamma, rho, 6)
        assert rel_l2(got[k], ref) < 1e-11, (k, rel_l2(got[k], ref))


@pytest.mark.parametrize("shape", [(256, 128), (200, 200)])
def test_line2_chebyshev_converges_to_the_reference_solution(shape):
    """The two-level method reaches the tolerance in fewer sweeps than the one-level block-line method and both agree with
    the reference's own Jacobi iteration (STRICT, bit-identical to the oracle) to < 1e-8 relative L2."""
    torch, X, O = _mods()
    nx, ny = shape; nb = 6
    a, b, c = _aniso(nx, ny)
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    yy, xx = np.mgrid[0:ny, 0:nx]
    F = np.stack([np.exp(-((xx - nx * (0.2 + 0.1 * k)) / 9.0) ** 2 - ((yy - ny * 0.5) / 7.0) ** 2) for k in range(nb)])
    rms = np.sqrt((F[:, 1:-1, 1:-1] ** 2).mean(axis=(1, 2)))
    out = {}
    for method, arith in (("line2_chebyshev", "fast"), ("line_chebyshev", "fast"), ("jacobi", "strict")):
        plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=True, arith=arith, method=method)
        plan.set_coe_aos(coe)
        psi = torch.zeros((nb, ny, nx), dtype=torch.float64, device="cuda"); ft = torch.from_numpy(F).cuda()
        r1 = torch.from_numpy(1e-11 * rms).cuda()
        cs = 10 if method == "line2_chebyshev" else 50
        res = plan.solve(psi, ft, X.SolveParams(max_iter=2000000, check_step=cs, converge_time=1, r1=1.0, r2=0.0, r1_per_solve=r1, sync_every=3))
        assert np.all(res["err"] == 0), (method, res)
        out[method] = (psi.cpu().numpy(), res["iters"])
        if method == "line2_chebyshev":      # independent residual through do_elliptic (APPLY mode of a separate strict plan)
            chk = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=True, arith="strict")
            chk.set_coe_aos(coe)
            lpsi = chk.apply(psi).cpu().numpy(); chk.close()
            resid = np.sqrt(((lpsi - F)[:, 1:-1, 1:-1] ** 2).mean(axis=(1, 2)))
            assert np.all(resid <= 1.01e-11 * rms), resid / rms
            print("two-level: rho, gamma =", plan.cheb_params(), "sweeps", res["iters"])
        plan.close()
    for k in range(nb):
        assert rel_l2(out["line2_chebyshev"][0][k], out["jacobi"][0][k]) < 1e-8
    print("sweeps: two-level", out["line2_chebyshev"][1], "one-level", out["line_chebyshev"][1], "jacobi", out["jacobi"][1])
    assert out["line2_chebyshev"][1].max() * 1.5 <= out["line_chebyshev"][1].min(), (out["line2_chebyshev"][1], out["line_chebyshev"][1])


def test_line2_rejects_what_it_cannot_do():
    torch, X, O = _mods()
    with pytest.raises(RuntimeError, match="two-level"):
        X.Plan(16, 16, nbatch=2, dtype="f64", shared_coe=True, arith="fast", method="line2_chebyshev")     # no coarse node
    with pytest.raises(RuntimeError, match="two-level"):
        X.Plan(64, 48, nbatch=2, dtype="f32", shared_coe=True, arith="fast", method="line2_chebyshev")     # fp64 fields only


@pytest.mark.parametrize("shape,nb", [((128, 72), 3), ((200, 56), 5)])
def test_line2_one_operator_per_solve_matches_numpy_restatement(shape, nb):
    """Time-series layout: every solve has its own operator, hence its own Galerkin coarse operator and inverse."""
    torch, X, O = _mods()
    nx, ny = shape
    coes, Fs, Ps = [], [], []
    for k in range(nb):
        a, b, c, F, P = _batch(nx, ny, 1, np.float64, seed=100 + 7 * k)
        coes.append(O.cal_coe(a * (1.0 + 0.2 * k), b, c, 1.0, 0.5, nx, ny)[0]); Fs.append(F[0]); Ps.append(P[0])
    coe = np.stack(coes); F = np.stack(Fs); P = np.stack(Ps)
    plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=False, arith="fast", method="line2_jacobi")
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
    plan.sweeps(psi, ft, 0.45, 3)
    plan.close()
    got = psi.cpu().numpy()
    for k in range(nb):
        ref, _ = TwoLevel(coe[k]).jacobi(P[k], F[k], 0.45, 3)
        assert rel_l2(got[k], ref) < 1e-11, (k, rel_l2(got[k], ref))


def test_series_with_the_two_level_method():
    """BASELINE config 5 chain (one operator per snapshot) with the two-level method against the one-level block-line method:
    same tables to the tolerances of the solve, far fewer sweeps."""
    import xlab_ee_fortran_b200 as X
    from xlab_ee_fortran_b200 import workloads as W
    from xlab_ee_fortran_b200.time_series import TimeSeries
    nr, nz, ns = 256, 128, 6
    Lr, Lz = (0.0, 1.0e6), (0.0, 1.5e4)
    params = W.series_params(ns, total=64, first=5)
    tabs = {}; psis = {}
    for method, cs in (("line_chebyshev", 25), ("line2_chebyshev", 10)):
        ts = TimeSeries(nr, nz, Lr, Lz, ns, "f64", arith="fast", method=method, r1_rel=1e-11)
        tabs[method] = ts.run(params, X.SolveParams(max_iter=400000, check_step=cs, converge_time=2, r1=1.0, r2=0.0, stall_checks=20))
        psis[method] = ts.field("psi")
        ts.close()
    t0, t1 = tabs["line_chebyshev"], tabs["line2_chebyshev"]
    print("sweeps one-level", t0[:, 0], "two-level", t1[:, 0])
    assert np.all(t1[:, 2] == 0) and np.all(t0[:, 2] == 0)
    for k in range(ns):
        assert rel_l2(psis["line2_chebyshev"][k], psis["line_chebyshev"][k]) < 1e-8
    assert np.allclose(t1[:, 5], t0[:, 5], rtol=1e-6) and np.allclose(t1[:, 6:], t0[:, 6:], rtol=1e-6)
    assert t1[:, 0].max() * 1.5 <= t0[:, 0].min()


def test_efficiency_map_with_the_two_level_method():
    """BASELINE config 4 geometry (512 x 256), 24 heating locations: efficiencies of the two-level and the one-level method
    within 1e-6 relative (north_star), streamfunctions within 1e-8 relative L2, and at most 400 sweeps per solve."""
    torch, X, O = _mods()
    from xlab_ee_fortran_b200 import workloads as W
    from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap
    nr, nz = 512, 256
    LR, LZ = (0.0, 1.0e6), (0.0, 1.5e4)
    A, B, C = W.vortex_fields(nr, nz, LR, LZ)[:3]
    lat = W.heating_lattice(64, 64, LR, LZ, 2 * LR[1] / (nr - 1), 2 * LZ[1] / (nz - 1))
    heat = lat[(np.arange(24) * 171) % len(lat)]
    tabs = {}; psis = {}
    for method, cs in (("line_chebyshev", 25), ("line2_chebyshev", 10)):
        prm = X.SolveParams(max_iter=400000, check_step=cs, converge_time=2, r1=1.0, r2=0.0, sync_every=2, stall_checks=20)
        m = EfficiencyMap(A, B, C, LR, LZ, len(heat), "f64", arith="fast", method=method, r1_rel=1e-12)
        tabs[method] = m.run(heat, prm)
        psis[method] = m.field("psi")
        m.close()
    t0, t1 = tabs["line_chebyshev"], tabs["line2_chebyshev"]
    print("sweeps one-level", t0[:, 0].min(), t0[:, 0].max(), "two-level", t1[:, 0].min(), t1[:, 0].max())
    assert np.all((t1[:, 2] == 0) | (t1[:, 2] == 4)) and np.all((t0[:, 2] == 0) | (t0[:, 2] == 4))
    assert np.all(np.abs(t1[:, 5] - t0[:, 5]) <= 1e-6 * np.abs(t0[:, 5]).max())
    for k in range(len(heat)):
        assert rel_l2(psis["line2_chebyshev"][k], psis["line_chebyshev"][k]) < 1e-8
    assert t1[:, 0].max() <= 400


def test_drop_in_entry_with_the_two_level_method(monkeypatch):
    """solve_elliptic (the Fortran-facing signature) with XEE_METHOD=line2_chebyshev XEE_ARITH=fast: reference case test1
    (200x200, fp64) converged to a tight r1 agrees with the reference iteration run to the same tolerance."""
    torch, X, O = _mods()
    from tests.util import ref_test1_inputs
    A, B, C, bc = ref_test1_inputs()
    d = O.Domain((0.0, 1.0), (0.0, 1.0), 200, 200, 0, 0)
    a, b, c = O.build_abc(A.astype(np.float64), B.astype(np.float64), C.astype(np.float64), d)
    g = O.geometry(d, np.float64)
    coe, _ = O.cal_coe(a, b, c, g["dr"], g["dz"], 200, 200)
    f = -B.astype(np.float64)                       # DYNAMIC_EFFICIENCY: f = -B_in (initialize-variables.f90:38-42)
    r1 = 1e-10 * float(np.sqrt((f[1:-1, 1:-1] ** 2).mean()))
    ref = O.solve_elliptic(2000000, 100, 2, 5, r1, 0.0, 1.0, bc.astype(np.float64), coe, f)
    assert ref["err"] == 0
    monkeypatch.setenv("XEE_METHOD", "line2_chebyshev"); monkeypatch.setenv("XEE_ARITH", "fast")
    dat = bc.astype(np.float64).copy(); wk = np.zeros_like(dat)
    it, r1o, r2o, err = X.solve_elliptic(2000000, 10, 2, 5, r1, 0.0, 1.0, dat, coe, f, wk, 200, 200)
    assert err == 0 and r1o < r1 and it * 50 < ref["max_iter"], (it, ref["max_iter"])
    assert rel_l2(dat, ref["dat"]) < 1e-8
