"""GPU parity of the time-series chain (BASELINE config 5: one operator per snapshot, thermal + dynamical source
term, pumping boundary condition, device-side vortex builder) against the oracle composition."""
import numpy as np
import pytest

from tests.series_oracle import series_rows
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _f32_close(a, b):
    """Device-built fields go through float32 like the reference's files: allow a last-bit difference."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= 1.3e-7 * np.maximum(np.abs(a), np.abs(b)))


def test_series_matches_oracle():
    import xlab_ee_fortran_b200 as X
    from xlab_ee_fortran_b200 import workloads as W
    from xlab_ee_fortran_b200.time_series import TimeSeries
    nr, nz, ns = 64, 40, 4
    Lr, Lz = (0.0, 6.0e5), (0.0, 1.5e4)
    params = W.series_params(ns)
    kw = dict(max_iter=2000000, check_step=100, converge_time=2, r1_rel=1e-11)
    ref = series_rows(params, nr, nz, Lr, Lz, np.float64, kw)
    ts = TimeSeries(nr, nz, Lr, Lz, ns, "f64", arith="fast", method="chebyshev", r1_rel=1e-11)
    tab = ts.run(params, X.SolveParams(max_iter=400000, check_step=50, converge_time=2, r1=1.0, r2=0.0))
    for name in ("A", "B", "C"):
        assert _f32_close(ts.field(name), ref[name]), name            # device vortex builder
    # feed the oracle's float32 fields? no: compare the downstream chain at tolerance (inputs may differ in 1 f32 ulp)
    assert rel_l2(ts.field("m2"), ref["m2"]) < 1e-6
    assert rel_l2(ts.field("f"), ref["f"]) < 1e-6
    assert np.all(tab[:, 2] == 0) and np.all(tab[:, 0] * 4 < ref["table"][:, 0])
    psi = ts.field("psi")
    # pumping boundary row: a quartic with cancellation (pow vs x*x*x*x): absolute tolerance 1e-14 of its scale
    assert np.allclose(psi[:, 0, :], ref["psi"][:, 0, :], rtol=1e-12, atol=1e-14 * 7e7)
    assert np.all(psi[:, -1, :] == 0) and np.all(psi[:, :, 0] == 0) and np.all(psi[:, :, -1] == 0)
    for n in range(ns):
        assert rel_l2(psi[n], ref["psi"][n]) < 2e-6                      # limited by the 1-ulp(f32) input differences
    assert np.allclose(tab[:, 3], ref["table"][:, 3], rtol=1e-12)
    assert np.allclose(tab[:, 5], ref["table"][:, 5], rtol=1e-5)
    assert np.allclose(tab[:, 6], ref["table"][:, 6], rtol=1e-5) and np.allclose(tab[:, 7], ref["table"][:, 7], rtol=1e-5)
    print("snapshots: sweeps", tab[:, 0], "efficiency", tab[:, 5], "max|w|", tab[:, 6])


def test_series_exact_chain_on_identical_inputs():
    """Same chain with the oracle fed the DEVICE-built A,B,C (so inputs are identical): north_star tolerances."""
    import xlab_ee_fortran_b200 as X
    from oracle import numpy_ref as N
    from oracle import oracle as O
    from tests.map_oracle import background_theta, heat_field
    from xlab_ee_fortran_b200 import workloads as W
    from xlab_ee_fortran_b200.time_series import TimeSeries
    nr, nz, ns = 48, 36, 2
    Lr, Lz = (0.0, 5.0e5), (0.0, 1.4e4)
    dt = np.float64
    params = W.series_params(ns, total=9, first=3)
    ts = TimeSeries(nr, nz, Lr, Lz, ns, "f64", arith="strict", method="jacobi", r1_rel=1e-10)
    tab = ts.run(params, X.SolveParams(max_iter=2000000, check_step=100, converge_time=2, r1=1.0, r2=0.0))
    A, B, C = ts.field("A"), ts.field("B"), ts.field("C")
    d = O.Domain(Lr, Lz, nr, nz, 0, 0); g = O.geometry(d, dt); k = N.constants(dt)
    for n in range(ns):
        a, b, c = O.build_abc(A[n], B[n], C[n], d)
        coe, _ = O.cal_coe(a, b, c, g["dr"], g["dz"], nr, nz)
        _, _, _, rhoC_C = O.stagger_averages(A[n], B[n], C[n], d)
        m2 = O.angular_momentum_sq(rhoC_C, d)
        assert rel_l2(ts.field("m2")[n], m2) < 1e-14
        _, _, _, bottom, F = W.series_fields_host(params[n], nr, nz, Lr, Lz)
        f = O.rhs_thermal(heat_field(params[n][14:19], g, dt), d)[1] + O.rhs_momentum(m2, F, d)
        assert rel_l2(ts.field("f")[n], f) < 1e-12
        fdev = ts.field("f")[n]
        assert np.allclose(ts.field("psi")[n][0], bottom, rtol=1e-12, atol=1e-14 * 7e7)
        psi0 = np.zeros((nz, nr)); psi0[0] = ts.field("psi")[n][0]        # identical boundary data on both sides
        rms = float(np.sqrt(((O.do_elliptic(psi0, coe) - fdev)[1:-1, 1:-1] ** 2).mean()))
        r = O.solve_elliptic(2000000, 100, 2, 5, 1e-10 * rms, 0.0, 1.0, psi0, coe, fdev)
        assert abs(r["max_iter"] - tab[n, 0]) <= 100 and r["err"] == 0     # strict Jacobi on identical inputs (r1 differs in the last bits)
        if r["max_iter"] == tab[n, 0]:
            assert np.array_equal(ts.field("psi")[n], r["dat"])          # ... and then a bit-identical field
        assert rel_l2(ts.field("psi")[n], r["dat"]) < 1e-8
        u, w = O.cal_uw(ts.field("psi")[n], d)
        assert np.array_equal(ts.field("u")[n], u) and np.array_equal(ts.field("w")[n], w)
        assert np.array_equal(ts.field("theta")[n], background_theta(A[n], B[n], C[n], d, dt))
        ke = O.integrate_weight_B(O.cal_wtheta(w, ts.field("theta")[n], d), d) * float(k["g0"]) / float(k["theta0"])
        assert tab[n, 4] == pytest.approx(ke, rel=1e-9)


def test_series_sharded_equals_unsharded():
    """time_series.run_sharded (BASELINE config 5 over several GPUs): the concatenated per-rank tables are the table of
    the whole series run in one piece (independent solves: identical sweep counts and fields' scalars)."""
    import xlab_ee_fortran_b200 as X
    from xlab_ee_fortran_b200 import workloads as W
    from xlab_ee_fortran_b200.time_series import TimeSeries, run_sharded
    nr, nz, total = 64, 48, 7
    Lr, Lz = (0.0, 6.0e5), (0.0, 1.5e4)
    prm = X.SolveParams(max_iter=400000, check_step=25, converge_time=2, r1=1.0, r2=0.0, stall_checks=20)
    kw = dict(dtype="f64", arith="fast", method="line_chebyshev", r1_rel=1e-10)
    ts = TimeSeries(nr, nz, Lr, Lz, total, **kw)
    whole = ts.run(W.series_params(total), prm)
    assert ts.kernel_info()[0] == 5 and ts.probe_ms() > 0.0
    ts.close()
    parts = []
    for rank in range(3):
        a, b, tab = run_sharded(nr, nz, Lr, Lz, total, prm, world=3, rank=rank, **kw)
        assert tab.shape == (b - a, 8)
        parts.append(tab)
    got = np.concatenate(parts)
    assert np.all(got[:, 2] == 0) and np.array_equal(got[:, 0], whole[:, 0])
    assert np.allclose(got[:, 3:], whole[:, 3:], rtol=1e-9, atol=0)


def test_line_methods_converge_on_the_nonsymmetric_late_snapshots():
    """Regression for the first 8-GPU run of config 5: on the later snapshots of the series the operator is strongly
    non-symmetric (sharp B next to the vortex ring) and the iteration matrix of the line splittings has a complex eigenvalue
    pair; with Chebyshev weights for a real interval the one-level method diverged and the two-level one crawled.  With the
    measured imaginary semi-axis (stage C of the spectral estimate, elliptic parameters) both converge and agree with the
    point method, whose spectrum is real."""
    import xlab_ee_fortran_b200 as X
    from xlab_ee_fortran_b200 import workloads as W
    from xlab_ee_fortran_b200.time_series import TimeSeries
    nr, nz, ns = 512, 256, 2
    Lr, Lz = (0.0, 1.0e6), (0.0, 1.5e4)
    params = W.series_params(ns, total=1024, first=700)
    res = {}
    for method, cs in (("chebyshev", 100), ("line_chebyshev", 25), ("line2_chebyshev", 10)):
        ts = TimeSeries(nr, nz, Lr, Lz, ns, "f64", arith="fast", method=method, r1_rel=1e-11)
        tab = ts.run(params, X.SolveParams(max_iter=400000, check_step=cs, converge_time=2, r1=1.0, r2=0.0, stall_checks=40))
        res[method] = (tab, ts.field("psi"))
        ts.close()
        # (the point method stops on its round-off floor, a few 1e-20, just above this tolerance: err 4; the line methods reach it)
        assert np.all(tab[:, 2] == 0) or (method == "chebyshev" and np.all((tab[:, 2] == 0) | (tab[:, 2] == 4))), (method, tab[:, :3])
    print("sweeps:", {m: res[m][0][:, 0].tolist() for m in res})
    for m in ("line_chebyshev", "line2_chebyshev"):
        for k in range(ns):
            assert rel_l2(res[m][1][k], res["chebyshev"][1][k]) < 1e-8, (m, k)
    assert res["line2_chebyshev"][0][:, 0].max() < 2500


def test_subsampled_spectral_probes_match_full_probes(monkeypatch):
    """Smooth series of 40 snapshots: probing every 8th operator and interpolating gives the same solutions and about the
    same sweep counts as probing every operator."""
    import xlab_ee_fortran_b200 as X
    from xlab_ee_fortran_b200 import workloads as W
    from xlab_ee_fortran_b200.time_series import TimeSeries
    nr, nz, ns = 128, 64, 40
    Lr, Lz = (0.0, 1.0e6), (0.0, 1.5e4)
    params = W.series_params(ns, total=1024, first=100)
    out = {}
    for sub in ("0", "8"):
        monkeypatch.setenv("XEE_RHO_SUBSAMPLE", sub)
        ts = TimeSeries(nr, nz, Lr, Lz, ns, "f64", arith="fast", method="line2_chebyshev", r1_rel=1e-11)
        tab = ts.run(params, X.SolveParams(max_iter=200000, check_step=10, converge_time=2, r1=1.0, r2=0.0, stall_checks=20))
        out[sub] = (tab, ts.field("psi"), ts.probe_ms())
        ts.close()
        assert np.all(tab[:, 2] == 0)
    for k in range(ns):
        assert rel_l2(out["8"][1][k], out["0"][1][k]) < 1e-8
    assert out["8"][0][:, 0].max() <= 1.3 * out["0"][0][:, 0].max()
    print("sweeps full", out["0"][0][:, 0].min(), out["0"][0][:, 0].max(), "subsampled", out["8"][0][:, 0].min(), out["8"][0][:, 0].max(),
          "probe ms", out["0"][2], out["8"][2])
