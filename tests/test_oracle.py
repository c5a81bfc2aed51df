"""CPU tests pinning the oracle (oracle/xee_oracle.hpp, the C++ restatement of the reference).

Pins, in the order SURVEY section 8(c) lists them:
  1. the reference's own fixtures: test/test1 inputs regenerate byte-identically (sha256 recorded
     from /root/reference by tests/golden/make_golden.py);
  2. golden vectors from the independent numpy restatement (oracle/numpy_ref.py) - bit for bit;
  3. the surveyor's known-answer values;
  4. structural invariants of the operator.
The reference holds no expected OUTPUTS, so parity stays "unpinned by the reference's tests".
"""
import os

import numpy as np
import pytest

from oracle import numpy_ref as N
from oracle import oracle as O
from tests.util import GOLDEN, golden_json, rel_l2, sha, ref_test1_inputs

DTS = {"f32": np.float32, "f64": np.float64}


def _test1_setup(dt):
    A, B, C, bc = ref_test1_inputs()
    d = O.Domain((0.0, 1.0), (0.0, 1.0), 200, 200, 0, 0)
    g = O.geometry(d, dt)
    a, b, c = O.build_abc(A.astype(dt), B.astype(dt), C.astype(dt), d)
    coe, err = O.cal_coe(a, np.zeros_like(b), c, g["dr"], g["dz"], 200, 200)   # BAROTROPIC: diagnose.f90:6
    assert err == 0
    return d, g, coe, -(B.astype(dt)), bc.astype(dt)


def test_test1_inputs_match_reference_fixtures():
    gj = golden_json()
    A, B, C, bc = ref_test1_inputs()
    ref = gj["reference_test1_sha256"]
    assert sha(A) == ref["A.bin"] and sha(B) == ref["B.bin"] and sha(C) == ref["C.bin"] and sha(bc) == ref["bc_init.bin"]
    assert gj["reference_test1_diag_txt"].splitlines()[0].startswith("DYNAMIC_EFFICIENCY-CYLINDRICAL-DENSITY_NORMAL-BAROTROPIC")


@pytest.mark.parametrize("name", ["f32", "f64"])
def test_test1_1000_sweeps_bitwise_vs_golden(name):
    dt = DTS[name]
    gold = golden_json()["test1"][name]
    d, g, coe, f, bc = _test1_setup(dt)
    assert sha(coe) == gold["coe_sha256"]
    r = O.solve_elliptic(1000, 100, 10, 5, 5e-3, 5e-3, 1.0, bc, coe, f, trace_cap=16)
    assert r["max_iter"] == 1000 and r["err"] == 1          # max_iter hit: err bit 0 (elliptic_tools.f90:242-244)
    assert sha(r["dat"]) == gold["psi1000_sha256"]
    tr = gold["trace_1000"]
    assert [t[0] for t in tr] == list(r["trace"]["iter"])
    assert [t[1] for t in tr] == list(r["trace"]["err_now"])
    assert [t[2] for t in tr] == list(r["trace"]["ratio"])
    r100 = O.solve_elliptic(100, 100, 10, 5, 5e-3, 5e-3, 1.0, bc, coe, f)
    assert sha(r100["dat"]) == gold["psi100_sha256"]


def test_survey_known_answers_fp64():
    """SURVEY.md section 6 (surveyor's transliteration, rho via numpy pow: agree to ~1e-9, not bitwise)."""
    d, g, coe, f, bc = _test1_setup(np.float64)
    r = O.solve_elliptic(1000, 100, 10, 5, 5e-3, 5e-3, 1.0, bc, coe, f, trace_cap=16)
    e = r["trace"]["err_now"]
    assert e[0] == pytest.approx(4.610952877e-3, rel=1e-9)
    assert e[1] == pytest.approx(4.227257955e-3, rel=1e-9)
    assert e[9] == pytest.approx(2.113276796e-3, rel=1e-9)
    psi = r["dat"]
    assert psi.min() == pytest.approx(-3.894914300e-5, rel=1e-9)
    assert psi.max() == pytest.approx(3.893837563e-5, rel=1e-9)
    assert np.sqrt((psi ** 2).sum()) == pytest.approx(2.856300598e-3, rel=1e-9)
    assert psi[99, 99] == pytest.approx(5.033140907e-7, rel=1e-8)


def test_survey_known_answers_fp32():
    d, g, coe, f, bc = _test1_setup(np.float32)
    r = O.solve_elliptic(1000, 100, 10, 5, 5e-3, 5e-3, 1.0, bc, coe, f, trace_cap=16)
    e = r["trace"]["err_now"]
    assert e[0] == pytest.approx(4.610931501e-3, rel=2e-6)
    assert e[1] == pytest.approx(4.227243830e-3, rel=2e-6)
    assert e[9] == pytest.approx(2.113273833e-3, rel=1e-5)   # 1-ulp rho (powf) differences accumulate in real(4)
    assert np.sqrt((r["dat"].astype(np.float64) ** 2).sum()) == pytest.approx(2.856300e-3, rel=1e-6)


@pytest.mark.parametrize("name", ["f32", "f64"])
def test_test1_full_run_to_stop_rule(name):
    """BASELINE config 1 end to end on the oracle: stop sweep, residuals, rchi and eta vs golden."""
    dt = DTS[name]
    gold = golden_json()["test1"][name]
    d, g, coe, f, bc = _test1_setup(dt)
    r = O.solve_elliptic(100000, 100, 10, 5, 5e-3, 5e-3, 1.0, bc, coe, f, trace_cap=1024)
    assert r["max_iter"] == gold["stop_sweeps"] and r["err"] == gold["stop_err"] == 0
    assert r["r1"] == gold["stop_r1"] and r["r2"] == gold["stop_r2"]
    assert sha(r["dat"]) == gold["psi_stop_sha256"]
    eta = O.cal_eta(r["dat"], d)
    assert sha(eta) == gold["eta_stop_sha256"]
    ds = np.load(os.path.join(GOLDEN, f"test1_psi_stop_{name}_ds.npy"))
    assert np.array_equal(ds, r["dat"][::8, ::8])
    # SURVEY section 6 sanity band for the converged field (rounding decides the stop sweep, not the field)
    assert np.sqrt((r["dat"].astype(np.float64) ** 2).sum()) == pytest.approx(5.107995457e-3, rel=2e-5)
    assert float(eta.max()) == pytest.approx(2.853969e-8, rel=2e-5)


@pytest.mark.parametrize("name", ["f32", "f64"])
def test_small_baroclinic_case_bitwise(name):
    dt = DTS[name]
    z = np.load(os.path.join(GOLDEN, f"small_{name}.npz"))
    nr, nz = z["A"].shape[1], z["A"].shape[0]
    d = O.Domain(tuple(z["Lr"]), tuple(z["Lz"]), nr, nz, 0, 0)
    g = O.geometry(d, dt)
    for k in ("ra", "za", "rho", "exner"):
        assert np.array_equal(g[k], z[k]), k
    a, b, c = O.build_abc(z["A"].astype(dt), z["B"].astype(dt), z["C"].astype(dt), d)
    assert np.array_equal(a, z["a"]) and np.array_equal(b, z["b"]) and np.array_equal(c, z["c"])
    coe, _ = O.cal_coe(a, b, c, g["dr"], g["dz"], nr, nz)
    assert np.array_equal(coe, z["coe"])
    assert np.array_equal(O.do_elliptic(z["bc"].astype(dt), coe), z["Lpsi"])
    f = z["f"].astype(dt)
    r = O.solve_elliptic(50, 10, 3, 2, 1e-30, 1.0, 0.8, z["bc"].astype(dt), coe, f, trace_cap=8)
    assert np.array_equal(r["dat"], z["psi50"])
    assert np.array_equal(np.c_[r["trace"]["iter"], r["trace"]["err_now"], r["trace"]["ratio"]], z["trace50"])
    rs = O.solve_elliptic(200000, 50, 4, 3, float(z["r1_in"]), 2.0, 0.8, z["bc"].astype(dt), coe, f, trace_cap=64)
    assert rs["max_iter"] == int(z["stop_sweeps"]) and rs["err"] == int(z["stop_err"]) == 0
    assert rs["r1"] == float(z["stop_r1"]) and rs["r2"] == float(z["stop_r2"])
    assert np.array_equal(rs["dat"], z["psi_stop"])
    assert np.array_equal(O.cal_eta(rs["dat"], d), z["eta"], equal_nan=True)
    u, w = O.cal_uw(rs["dat"], d)
    assert np.array_equal(u, z["u"], equal_nan=True) and np.array_equal(w, z["w"], equal_nan=True)


def test_O0_build_matches_O3_bitwise():
    """The -O0 library (what make-diagnosis.sh:10-11 builds) and the -O3 parity build agree bit for bit."""
    d, g, coe, f, bc = _test1_setup(np.float32)
    r3 = O.solve_elliptic(300, 100, 10, 5, 5e-3, 5e-3, 1.0, bc, coe, f)
    r0 = O.solve_elliptic(300, 100, 10, 5, 5e-3, 5e-3, 1.0, bc, coe, f, variant="O0")
    assert np.array_equal(r3["dat"], r0["dat"]) and r3["r1"] == r0["r1"]


# ------------------------------------------------------------------ structural invariants
def _random_abc(nx, ny, rng, bscale):
    a = (1.0 + rng.random((ny - 2, nx - 1))); c = (1.0 + rng.random((ny - 1, nx - 2)))
    b = bscale * rng.standard_normal((ny - 1, nx - 1))
    return a, b, c


def test_coefficients_sum_to_zero_and_corners_vanish_for_B0():
    rng = np.random.default_rng(1)
    nx, ny = 31, 23
    a, b, c = _random_abc(nx, ny, rng, 0.3)
    coe, _ = O.cal_coe(a, b, c, 0.7, 1.3, nx, ny)
    assert np.abs(coe[1:-1, 1:-1].sum(-1)).max() < 1e-13 * np.abs(coe).max()
    coe0, _ = O.cal_coe(a, np.zeros_like(b), c, 0.7, 1.3, nx, ny)
    assert np.all(coe0[1:-1, 1:-1][..., [0, 2, 6, 8]] == 0)
    assert np.all(coe[0] == 0) and np.all(coe[:, 0] == 0)      # boundary entries never written


def test_stencil_equals_flux_form_divergence():
    """L psi = d_r(a d_r psi + b d_z psi) + d_z(b d_r psi + c d_z psi) with 4-point-averaged cross terms."""
    rng = np.random.default_rng(2)
    nx, ny = 27, 19
    dx, dy = 0.6, 1.1
    a, b, c = _random_abc(nx, ny, rng, 0.2)
    psi = rng.standard_normal((ny, nx))
    coe, _ = O.cal_coe(a, b, c, dx, dy, nx, ny)
    L = O.do_elliptic(psi, coe)[1:-1, 1:-1]
    J, I = np.meshgrid(np.arange(1, ny - 1), np.arange(1, nx - 1), indexing="ij")   # 0-based interior
    P = lambda dj, di: psi[J + dj, I + di]
    # a(i,j-1) lives at (i+1/2, j): a[J-1, I]; c(i-1,j) at (i, j+1/2): c[J, I-1]; b(i,j) at (i+1/2, j+1/2): b[J, I]
    Fr_p = a[J - 1, I] * (P(0, 1) - P(0, 0)) / dx; Fr_m = a[J - 1, I - 1] * (P(0, 0) - P(0, -1)) / dx
    Fz_p = c[J, I - 1] * (P(1, 0) - P(0, 0)) / dy; Fz_m = c[J - 1, I - 1] * (P(0, 0) - P(-1, 0)) / dy
    bxp = 0.5 * (b[J, I] + b[J - 1, I]); bxm = 0.5 * (b[J, I - 1] + b[J - 1, I - 1])
    byp = 0.5 * (b[J, I - 1] + b[J, I]); bym = 0.5 * (b[J - 1, I - 1] + b[J - 1, I])
    dz_at_xp = (P(1, 0) + P(1, 1) - P(-1, 0) - P(-1, 1)) / (4 * dy)
    dz_at_xm = (P(1, -1) + P(1, 0) - P(-1, -1) - P(-1, 0)) / (4 * dy)
    dr_at_yp = (P(0, 1) + P(1, 1) - P(0, -1) - P(1, -1)) / (4 * dx)
    dr_at_ym = (P(-1, 1) + P(0, 1) - P(-1, -1) - P(0, -1)) / (4 * dx)
    flux = (Fr_p - Fr_m) / dx + (Fz_p - Fz_m) / dy + (bxp * dz_at_xp - bxm * dz_at_xm) / dx + (byp * dr_at_yp - bym * dr_at_ym) / dy
    assert np.abs(L - flux).max() < 5e-13 * np.abs(L).max()


def test_operator_symmetric_for_constant_B():
    rng = np.random.default_rng(3)
    nx, ny = 12, 10
    a, _, c = _random_abc(nx, ny, rng, 0.0)
    b = np.full((ny - 1, nx - 1), 0.17)
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    n = (nx - 2) * (ny - 2)
    M = np.zeros((n, n))
    for q in range(n):
        e = np.zeros((ny, nx)); e[1 + q // (nx - 2), 1 + q % (nx - 2)] = 1.0
        M[:, q] = O.do_elliptic(e, coe)[1:-1, 1:-1].ravel()
    assert np.abs(M - M.T).max() < 1e-13 * np.abs(M).max()


# ------------------------------------------------------------------ solver control flow
def _tiny(dt=np.float64):
    rng = np.random.default_rng(5)
    nx, ny = 18, 14
    a, b, c = _random_abc(nx, ny, rng, 0.05)
    coe, _ = O.cal_coe(a.astype(dt), b.astype(dt), c.astype(dt), 1.0, 1.0, nx, ny)
    f = rng.standard_normal((ny, nx)).astype(dt); x0 = rng.standard_normal((ny, nx)).astype(dt)
    return coe, f, x0


def test_both_criteria_disabled_is_an_error():
    coe, f, x0 = _tiny()
    r = O.solve_elliptic(10, 5, 1, 1, 0.0, -1.0, 1.0, x0, coe, f)
    assert r["rc"] == -1 and np.array_equal(r["dat"], x0)


def test_defaults_when_nonpositive_control_args():
    coe, f, x0 = _tiny()
    r = O.solve_elliptic(1000, 0, 0, 0, 1e30, 1e30, 1.0, x0, coe, f, trace_cap=64)
    # check_step->100, converge_time->10: first 10 checks all pass => stops at sweep 1000
    assert list(r["trace"]["iter"]) == list(range(100, 1001, 100)) and r["max_iter"] == 1000
    assert r["err"] == 1          # cnt == max_iter coincides with the converged stop: bit 0 still set (:242-244)


def test_r1_only_and_boundary_untouched_and_fixed_point():
    coe, f, x0 = _tiny()
    rms_f = np.sqrt((f[1:-1, 1:-1] ** 2).mean())
    r = O.solve_elliptic(100000, 10, 2, 5, 1e-11 * rms_f, 0.0, 1.0, x0, coe, f, trace_cap=4096)
    assert r["err"] == 0 and r["r1"] < 1e-11 * rms_f
    x = r["dat"]
    assert np.array_equal(x[0], x0[0]) and np.array_equal(x[-1], x0[-1])
    assert np.array_equal(x[:, 0], x0[:, 0]) and np.array_equal(x[:, -1], x0[:, -1])
    res = O.do_elliptic(x, coe)[1:-1, 1:-1] - f[1:-1, 1:-1]
    assert np.sqrt((res ** 2).mean()) < 2e-11 * rms_f
    # linearity of the fixed point: solve(2f, 2bc) == 2 solve(f, bc)
    r2 = O.solve_elliptic(100000, 10, 2, 5, 2e-11 * rms_f, 0.0, 1.0, 2 * x0, coe, 2 * f)
    assert rel_l2(r2["dat"], 2 * x) < 1e-9


def test_zero_max_iter_and_parity_of_buffers():
    coe, f, x0 = _tiny()
    r0 = O.solve_elliptic(0, 10, 1, 1, 1.0, 1.0, 1.0, x0, coe, f)
    assert np.array_equal(r0["dat"], x0) and r0["max_iter"] == 0 and r0["err"] == 0
    r3 = O.solve_elliptic(3, 10, 1, 1, 1.0, 1.0, 1.0, x0, coe, f)
    r4 = O.solve_elliptic(4, 10, 1, 1, 1.0, 1.0, 1.0, x0, coe, f)
    rn3 = N.solve_elliptic(3, 10, 1, 1, 1.0, 1.0, 1.0, x0, coe, f)
    rn4 = N.solve_elliptic(4, 10, 1, 1, 1.0, 1.0, 1.0, x0, coe, f)
    assert np.array_equal(r3["dat"], rn3["dat"]) and np.array_equal(r4["dat"], rn4["dat"])
    assert not np.array_equal(r3["dat"], r4["dat"])


def test_fair_fused_sweeps_are_bitwise_the_same_iteration():
    coe, f, x0 = _tiny()
    r = O.solve_elliptic(7, 100, 1, 1, 1.0, 1.0, 0.9, x0, coe, f)
    planar = np.ascontiguousarray(np.moveaxis(coe, -1, 0))
    fb = O.fair_batch(x0[None], planar, f[None], 0.9, 7, want_rms=True, threads=1)
    assert np.array_equal(fb["x"][0], r["dat"])


def test_batch_runs_independent_solves():
    coe, f, x0 = _tiny()
    F = np.stack([f, 2 * f, -f]); X = np.stack([x0, x0, 0 * x0])
    rb = O.solve_batch(40, 10, 1, 1, 1e-30, 1.0, 1.0, X, coe, F, threads=3)
    for k in range(3):
        r = O.solve_elliptic(40, 10, 1, 1, 1e-30, 1.0, 1.0, X[k], coe, F[k])
        assert np.array_equal(rb["dat"][k], r["dat"]) and rb["r1"][k] == r["r1"]


def test_energy_integrals_and_rhs_chain():
    """Legacy integral kernels: consistency of the restated maths (old-diagnose/diagnose.f90:1029-1127)."""
    dt = np.float64
    d = O.Domain((0.0, 3.0e5), (0.0, 1.0e4), 36, 28, 0, 0)
    g = O.geometry(d, dt)
    gN = N.geometry(d.Lr, d.Lz, d.nr, d.nz, dt)
    rng = np.random.default_rng(7)
    Q = rng.random((d.nz - 1, d.nr - 1))
    assert O.integrate_weight_B(Q, d) == float(N.integrate_weight_B(Q, gN))
    # integral of a constant = sum of cell masses
    one = np.ones_like(Q)
    mass = (((g["rho"][1:] + g["rho"][:-1]) / 2)[:, None] * ((g["ra"][1:] + g["ra"][:-1]) / 2)[None, :] * g["dr"] * g["dz"]).sum()
    assert O.integrate_weight_B(one, d) == pytest.approx(mass, rel=1e-12)
    eta = rng.random((d.nz, d.nr - 1))
    assert O.cal_sum_Qeta(Q, eta, d) == pytest.approx(O.integrate_weight_B(Q * (eta[:-1] + eta[1:]) / 2, d), rel=1e-12)
    w = rng.random((d.nz, d.nr - 1)); th = rng.random((d.nz - 1, d.nr - 1))
    assert np.allclose(O.cal_wtheta(w, th, d), (w[:-1] + w[1:]) / 2 * th, rtol=1e-15)
    J, rhs = O.rhs_thermal(Q, d)
    k = N.constants(dt)
    assert np.allclose(J, Q / (k["Cp"] * g["exner"][:-1, None]), rtol=1e-15)
    dJ = (J[:, 1:] - J[:, :-1]) / g["dr"]                # (ra(i+1)-ra(i-1))/2 == dr
    exp = np.zeros((d.nz, d.nr)); exp[1:-1, 1:-1] = (dJ[1:, :] + dJ[:-1, :]) / 2 * k["g0"] / k["theta0"]
    assert np.allclose(rhs, exp, rtol=1e-10, atol=1e-30)
    assert np.all(rhs[0] == 0) and np.all(rhs[:, 0] == 0)


def test_integral_check_identity_of_the_legacy_driver():
    """old-diagnose/diagnose.f90:677-725: sum(Q eta) ~ (g0/theta0) sum(w theta) when A and B derive from one theta
    field.  Validates the restated heating -> efficiency chain end to end on the oracle (first-order agreement)."""
    from tests.map_oracle import efficiency_rows
    from xlab_ee_fortran_b200 import workloads as W
    nr, nz = 64, 48
    Lr, Lz = (0.0, 6.0e5), (0.0, 1.5e4)
    A, B, C = W.vortex_fields(nr, nz, Lr, Lz)
    heat = W.heating_lattice(2, 2, Lr, Lz, 3 * Lr[1] / (nr - 1), 3 * Lz[1] / (nz - 1), r_frac=(0.02, 0.4), z_frac=(0.15, 0.75))
    tab, *_ = efficiency_rows(A, B, C, Lr, Lz, heat, np.float64, dict(max_iter=400000, check_step=100, converge_time=2, r1_rel=1e-10))
    assert np.all(tab[:, 2] == 0)
    assert np.allclose(tab[:, 7], tab[:, 5], rtol=2e-2)
    assert np.all(tab[:2, 5] > 1e-3)            # heating inside the vortex core is the efficient one


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("strategy,sr", [(1, 1e-3), (2, 0.3)])
def test_legacy_loop_restatement_agrees_with_the_mapped_current_solver(dt, strategy, sr):
    """Two independent restatements of the legacy stop rules (src/old-diagnose/xtt-lib/elliptic_tools.f90:168-300): the literal
    loop in tests/legacy_oracle.py and the mapping onto the current solver's criteria (strategy 1 = r1 only, converge_time 1;
    strategy 2 = r2 only, converge_time 10, lost_rate 5) that the drop-in uses.  Same sweeps, same field bit for bit."""
    from tests import legacy_oracle as L
    rng = np.random.default_rng(31)
    nx, ny = 48, 36
    a = (1.0 + rng.random((ny - 2, nx - 1))).astype(dt); c = (1.0 + rng.random((ny - 1, nx - 2))).astype(dt)
    b = (0.05 * rng.standard_normal((ny - 1, nx - 1))).astype(dt)
    f = rng.standard_normal((ny, nx)).astype(dt); x0 = rng.standard_normal((ny, nx)).astype(dt)
    x0[0] = 0; x0[-1] = 0; x0[:, 0] = 0; x0[:, -1] = 0
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    loop = L.old_solve_loop(strategy, sr, 3000, 0.9, x0, coe, f)
    dat, used, r = L.old_solve(strategy, sr, 3000, 0.9, x0, coe, f)
    assert 0 < loop["strategy"] < 3000 and loop["strategy"] == used and loop["err"] == 0
    assert np.array_equal(loop["dat"], dat)
    assert loop["strategy_r"] == pytest.approx(r, rel=1e-12 if dt is np.float64 else 1e-6)


def test_legacy_loop_max_norm_sees_the_dirichlet_rim():
    """Strategies 3 / 4 take maxval(abs(to_dat)) over the whole array (:203-204): with a non-zero boundary the norm never drops
    below the largest boundary value, so strategy 3 runs to max_iter and strategy 4 stops after its ten constant checks."""
    from tests import legacy_oracle as L
    rng = np.random.default_rng(5)
    nx, ny = 24, 20
    a = (1.0 + rng.random((ny - 2, nx - 1))); c = (1.0 + rng.random((ny - 1, nx - 2))); b = np.zeros((ny - 1, nx - 1))
    f = rng.standard_normal((ny, nx)); x0 = rng.standard_normal((ny, nx))
    rim = max(np.abs(x0[0]).max(), np.abs(x0[-1]).max(), np.abs(x0[:, 0]).max(), np.abs(x0[:, -1]).max())
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    r3 = L.old_solve_loop(3, 1e-3, 2000, 1.0, x0, coe, f)
    assert r3["strategy"] == 2000 and r3["err"] == 1 and r3["strategy_r"] == rim
    r4 = L.old_solve_loop(4, 0.05, 3000, 1.0, x0, coe, f)
    assert r4["err"] == 0 and r4["strategy_r"] == rim and r4["strategy"] % 100 == 0
