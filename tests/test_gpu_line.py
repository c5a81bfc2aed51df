"""GPU tests of the segment-line relaxation methods (XEE_METHOD_LINE_*, csrc/xee_sweep_line.cuh).

Not the reference's iteration, so no bit-identity: (1) sweep by sweep against the numpy restatement of the same
iteration (tests/line_oracle.py) to rounding; (2) the CONVERGED solution against the reference algorithm
(solve_elliptic, xtt-lib-fortran/elliptic_tools.f90:93-265, STRICT arithmetic = bit-identical to the oracle) within the
north_star tolerance of 1e-8 relative L2, plus an independent residual check through do_elliptic.
"""
import numpy as np
import pytest

from tests import line_oracle as LO
from tests.test_gpu_parity import DTS, _mods, _rand_case
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _batch(nx, ny, nb, dt, seed):
    a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=seed)
    F = np.stack([f * dt(k + 1) for k in range(nb)]); P = np.stack([x0 * dt(1 + 0.25 * k) for k in range(nb)])
    return a, b, c, F, P


@pytest.mark.parametrize("shape,nb", [((512, 256), 3), ((200, 200), 2), ((140, 70), 5), ((72, 40), 40), ((64, 32), 7),
                                      ((8, 4), 3), ((260, 13), 9), ((128, 40), 5), ((192, 23), 33)])   # last two: TMA-store path, ragged rows
@pytest.mark.parametrize("sweeps", [1, 2, 7])
def test_line_jacobi_sweeps_match_numpy_restatement(shape, nb, sweeps):
    torch, X, O = _mods()
    nx, ny = shape
    a, b, c, F, P = _batch(nx, ny, nb, np.float64, seed=nx + sweeps)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=True, arith="fast", method="line_jacobi")
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
    rms = plan.sweeps(psi, ft, 0.9, sweeps, want_rms=True)
    assert plan.kernel_info()[0] == 5
    plan.close()
    got = psi.cpu().numpy()
    for k in range(nb):
        ref, rref = LO.line_jacobi(P[k], coe, F[k], 0.9, sweeps)
        assert rel_l2(got[k], ref) < 1e-13, (k, rel_l2(got[k], ref))
        assert np.array_equal(got[k][0], P[k][0]) and np.array_equal(got[k][:, -1], P[k][:, -1])   # Dirichlet values
        assert abs(rms[k] - rref) <= 1e-11 * rref


def test_line_jacobi_f32():
    torch, X, O = _mods()
    nx, ny, nb = 200, 120, 4
    a, b, c, F, P = _batch(nx, ny, nb, np.float32, seed=5)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    plan = X.Plan(nx, ny, nbatch=nb, dtype="f32", shared_coe=True, arith="fast", method="line_jacobi")
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
    plan.sweeps(psi, ft, 1.0, 5)
    plan.close()
    got = psi.cpu().numpy()
    for k in range(nb):
        ref, _ = LO.line_jacobi(P[k].astype(np.float64), coe.astype(np.float64), F[k].astype(np.float64), 1.0, 5)
        assert rel_l2(got[k], ref) < 2e-5


@pytest.mark.parametrize("shape", [(256, 128), (200, 200)])
def test_line_chebyshev_converges_to_the_reference_solution(shape):
    """Anisotropic operator (radial coupling dominant, like the secondary-circulation operator): the line method reaches
    the tolerance in far fewer sweeps than Chebyshev-accelerated point Jacobi and both agree with the reference's own
    Jacobi iteration (STRICT, bit-identical to the oracle) to < 1e-8 relative L2."""
    torch, X, O = _mods()
    nx, ny = shape; nb = 6
    rng = np.random.default_rng(11)
    a = (40.0 * (1.0 + 0.3 * rng.random((ny - 2, nx - 1)))); c = (1.0 + 0.3 * rng.random((ny - 1, nx - 2)))
    b = 0.2 * rng.standard_normal((ny - 1, nx - 1))
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    yy, xx = np.mgrid[0:ny, 0:nx]
    F = np.stack([np.exp(-((xx - nx * (0.2 + 0.1 * k)) / 9.0) ** 2 - ((yy - ny * 0.5) / 7.0) ** 2) for k in range(nb)])
    rms = np.sqrt((F[:, 1:-1, 1:-1] ** 2).mean(axis=(1, 2)))
    out = {}
    for method, arith in (("line_chebyshev", "fast"), ("chebyshev", "fast"), ("jacobi", "strict")):
        plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=True, arith=arith, method=method)
        plan.set_coe_aos(coe)
        psi = torch.zeros((nb, ny, nx), dtype=torch.float64, device="cuda"); ft = torch.from_numpy(F).cuda()
        r1 = torch.from_numpy(1e-11 * rms).cuda()
        res = plan.solve(psi, ft, X.SolveParams(max_iter=2000000, check_step=50, converge_time=1, r1=1.0, r2=0.0, r1_per_solve=r1, sync_every=3))
        assert np.all(res["err"] == 0), (method, res)
        lpsi = plan.apply(psi).cpu().numpy()
        resid = np.sqrt(((lpsi - F)[:, 1:-1, 1:-1] ** 2).mean(axis=(1, 2)))
        assert np.all(resid <= 1.01e-11 * rms), (method, resid / rms)   # the solver's own sum order differs in the last digits
        out[method] = (psi.cpu().numpy(), res["iters"])
        plan.close()
    for k in range(nb):
        assert rel_l2(out["line_chebyshev"][0][k], out["jacobi"][0][k]) < 1e-8
        assert rel_l2(out["chebyshev"][0][k], out["jacobi"][0][k]) < 1e-8
    assert out["line_chebyshev"][1].max() * 2 <= out["chebyshev"][1].min(), (out["line_chebyshev"][1], out["chebyshev"][1])


def test_line_method_rejects_what_it_cannot_do():
    torch, X, O = _mods()
    with pytest.raises(RuntimeError, match="line-relaxation"):
        X.Plan(70, 32, nbatch=2, dtype="f32", shared_coe=True, arith="fast", method="line_chebyshev")   # rows not 16-byte multiples
    with pytest.raises(RuntimeError, match="line-relaxation"):
        X.Plan(64, 32, nbatch=2, dtype="f64", shared_coe=True, arith="strict", method="line_chebyshev")


def test_efficiency_map_with_the_line_method_matches_point_chebyshev():
    """BASELINE config 3 geometry, 24 heating locations: efficiencies of the two accelerated methods within 1e-6
    relative (north_star), streamfunctions within 1e-8 relative L2."""
    torch, X, O = _mods()
    from xlab_ee_fortran_b200 import workloads as W
    from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap
    nr, nz = 256, 128
    LR, LZ = (0.0, 1.0e6), (0.0, 1.5e4)
    A, B, C = W.vortex_fields(nr, nz, LR, LZ)[:3]
    lat = W.heating_lattice(64, 32, LR, LZ, 2 * LR[1] / (nr - 1), 2 * LZ[1] / (nz - 1))
    heat = lat[(np.arange(24) * 85) % len(lat)]
    prm = X.SolveParams(max_iter=400000, check_step=100, converge_time=1, r1=1.0, r2=0.0, sync_every=3)
    tabs = {}; psis = {}
    for method in ("chebyshev", "line_chebyshev"):
        m = EfficiencyMap(A, B, C, LR, LZ, len(heat), "f64", arith="fast", method=method, r1_rel=1e-12)
        tabs[method] = m.run(heat, prm)
        psis[method] = m.field("psi")
        m.close()
    t0, t1 = tabs["chebyshev"], tabs["line_chebyshev"]
    assert np.all(t1[:, 2] == 0) and np.all(t0[:, 2] == 0)
    assert np.all(np.abs(t1[:, 5] - t0[:, 5]) <= 1e-6 * np.abs(t0[:, 5]).max())
    for k in range(len(heat)):
        assert rel_l2(psis["line_chebyshev"][k], psis["chebyshev"][k]) < 1e-8
    assert t1[:, 0].max() * 2 <= t0[:, 0].min()


def test_drop_in_entry_with_the_line_method(monkeypatch):
    """solve_elliptic (the Fortran-facing signature) with XEE_METHOD=line_chebyshev XEE_ARITH=fast: reference case test1
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
    monkeypatch.setenv("XEE_METHOD", "line_chebyshev"); monkeypatch.setenv("XEE_ARITH", "fast")
    dat = bc.astype(np.float64).copy(); wk = np.zeros_like(dat)
    it, r1o, r2o, err = X.solve_elliptic(2000000, 25, 2, 5, r1, 0.0, 1.0, dat, coe, f, wk, 200, 200)
    assert err == 0 and r1o < r1 and it * 20 < ref["max_iter"], (it, ref["max_iter"])
    assert rel_l2(dat, ref["dat"]) < 1e-8


@pytest.mark.parametrize("shape,nb", [((136, 70), 5), ((512, 256), 3)])
def test_line_jacobi_with_one_operator_per_solve(shape, nb):
    """Time-series layout: every solve has its own operator (and its own factors, repacked per solve)."""
    torch, X, O = _mods()
    nx, ny = shape
    coes = []; Fs = []; Ps = []
    for k in range(nb):
        a, b, c, f, x0 = _rand_case(nx, ny, np.float64, seed=100 + k, bscale=0.02 * (k + 1))
        coe, _ = O.cal_coe(a * (1.0 + k), b, c, 1.0, 0.5, nx, ny)
        coes.append(coe); Fs.append(f); Ps.append(x0)
    plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=False, arith="fast", method="line_jacobi")
    plan.set_coe_aos(np.stack(coes))
    psi = torch.from_numpy(np.stack(Ps)).cuda(); ft = torch.from_numpy(np.stack(Fs)).cuda()
    rms = plan.sweeps(psi, ft, 1.0, 5, want_rms=True)
    plan.close()
    got = psi.cpu().numpy()
    for k in range(nb):
        ref, rref = LO.line_jacobi(Ps[k], coes[k], Fs[k], 1.0, 5)
        assert rel_l2(got[k], ref) < 1e-13, (k, rel_l2(got[k], ref))
        assert abs(rms[k] - rref) <= 1e-11 * rref


def test_time_series_with_the_line_method_matches_point_chebyshev():
    """BASELINE config 5 shape on a small grid: per-snapshot operators, pumping boundary condition, thermal + dynamical
    source; block-line Chebyshev against point Chebyshev: efficiencies within 1e-6 relative, fields within 1e-8."""
    torch, X, O = _mods()
    from xlab_ee_fortran_b200 import workloads as W
    from xlab_ee_fortran_b200.time_series import TimeSeries
    nr, nz, ns = 256, 128, 6
    LR, LZ = (0.0, 1.0e6), (0.0, 1.5e4)
    params = W.series_params(ns, total=1024, first=500)
    prm = X.SolveParams(max_iter=400000, check_step=50, converge_time=2, r1=1.0, r2=0.0, sync_every=3, stall_checks=40)
    tabs = {}; psis = {}
    for method in ("chebyshev", "line_chebyshev"):
        ts = TimeSeries(nr, nz, LR, LZ, ns, "f64", arith="fast", method=method, r1_rel=1e-10)
        tabs[method] = ts.run(params, prm)
        psis[method] = ts.field("psi")
        ts.close()
    t0, t1 = tabs["chebyshev"], tabs["line_chebyshev"]
    assert np.all(t0[:, 2] == 0) and np.all(t1[:, 2] == 0), (t0[:, 2], t1[:, 2])
    assert np.all(np.abs(t1[:, 5] - t0[:, 5]) <= 1e-6 * np.abs(t0[:, 5]))
    for k in range(ns):
        assert rel_l2(psis["line_chebyshev"][k], psis["chebyshev"][k]) < 1e-8
    assert t1[:, 0].max() * 2 <= t0[:, 0].min(), (t1[:, 0], t0[:, 0])


def test_rehosted_driver_on_test1_with_the_line_method(tmp_path, monkeypatch):
    """BASELINE config 1 end to end (the reference's diag.txt, real(4) .bin files in and out) with
    XEE_METHOD=line_chebyshev XEE_ARITH=fast: same stop rule (r1 = r2 = 5e-3), same converged field and eta as the
    reference iteration within the real(4) round-off floor the reference itself stops on."""
    from tests.test_gpu_parity import _run_diagnose
    from tests.util import golden_json
    gj = golden_json()
    monkeypatch.setenv("XEE_METHOD", "line_chebyshev"); monkeypatch.setenv("XEE_ARITH", "fast")
    monkeypatch.setenv("XEE_STALL_CHECKS", "10")   # the ratio criterion is noise on the real(4) floor of an accelerated method
    out = _run_diagnose(tmp_path, gj["reference_test1_diag_txt"])
    assert " Elliptic Tools: Iteration success." in out
    rchi = np.fromfile(tmp_path / "rchi-[BAROTROPIC]-O.bin", np.float32).reshape(200, 200)
    gold = gj["test1"]["f32"]
    assert np.sqrt((rchi.astype(np.float64) ** 2).sum()) == pytest.approx(gold["psi_stop_l2"], rel=2e-4)
    eta = np.fromfile(tmp_path / "eta-[BAROTROPIC]-A.bin", np.float32)
    assert float(eta.max()) == pytest.approx(gold["eta_stop_max"], rel=2e-3)
