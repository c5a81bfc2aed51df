"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle on the same inputs.

Tolerances (north_star): streamfunction within 1e-8 relative L2, efficiency within 1e-6 relative.
STRICT arithmetic is held to a tighter bar: iterates BIT-IDENTICAL to the oracle at fixed sweep
counts; only the residual norm (a parallel instead of sequential sum) is compared to a tolerance.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.util import GOLDEN, golden_json, ref_test1_inputs, rel_l2, sha

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DTS = {"f32": np.float32, "f64": np.float64}
RES_TOL = {"f32": 2e-4, "f64": 1e-12}     # residual norm: tree sum (GPU) vs sequential sum (reference)


def _mods():
    import torch
    import xlab_ee_fortran_b200 as X
    from oracle import oracle as O
    return torch, X, O


def _rand_case(nx, ny, dt, seed=0, bscale=0.05):
    rng = np.random.default_rng(seed)
    a = (1.0 + rng.random((ny - 2, nx - 1))).astype(dt); c = (1.0 + rng.random((ny - 1, nx - 2))).astype(dt)
    b = (bscale * rng.standard_normal((ny - 1, nx - 1))).astype(dt)
    f = rng.standard_normal((ny, nx)).astype(dt); x0 = rng.standard_normal((ny, nx)).astype(dt)
    return a, b, c, f, x0


@pytest.mark.parametrize("name", ["f32", "f64"])
@pytest.mark.parametrize("shape", [(3, 3), (5, 4), (67, 35), (200, 200), (512, 256)])
def test_cal_coe_and_do_elliptic_bitwise(name, shape):
    torch, X, O = _mods()
    dt = DTS[name]; nx, ny = shape
    a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=nx)
    coe = np.full((ny, nx, 9), 7.0, dt)                    # sentinel: boundary entries must stay untouched
    assert X.cal_coe(a, b, c, coe, 0.7, 1.3, nx, ny) == 0
    ref, _ = O.cal_coe(a, b, c, 0.7, 1.3, nx, ny)
    assert np.array_equal(coe[1:-1, 1:-1], ref[1:-1, 1:-1])
    assert np.all(coe[0] == 7) and np.all(coe[-1] == 7) and np.all(coe[:, 0] == 7) and np.all(coe[:, -1] == 7)
    out = np.full((ny, nx), 3.0, dt)
    assert X.do_elliptic(x0, ref, out, nx, ny) == 1        # err stays 1 (elliptic_tools.f90:73)
    oref = O.do_elliptic(x0, ref)
    assert np.array_equal(out[1:-1, 1:-1], oref[1:-1, 1:-1])
    assert np.all(out[0] == 3) and np.all(out[:, 0] == 3)


@pytest.mark.parametrize("name", ["f32", "f64"])
@pytest.mark.parametrize("shape,sweeps,alpha", [((3, 3), 5, 1.0), ((34, 19), 37, 0.8), ((130, 70), 200, 1.0), ((512, 256), 101, 0.95)])
def test_fixed_sweeps_bitwise(name, shape, sweeps, alpha):
    torch, X, O = _mods()
    dt = DTS[name]; nx, ny = shape
    a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=sweeps)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    dat = x0.copy(); wk = np.zeros_like(dat)
    it, r1, r2, err = X.solve_elliptic(sweeps, 10, 10, 5, 1e-30, 1.0, alpha, dat, coe, f, wk, nx, ny)
    ref = O.solve_elliptic(sweeps, 10, 10, 5, 1e-30, 1.0, alpha, x0, coe, f)
    assert (it, err) == (ref["max_iter"], ref["err"]) == (sweeps, 1)
    assert np.array_equal(dat, ref["dat"])
    assert np.array_equal(wk, ref["workspace"])            # the other ping-pong buffer, as the reference leaves it
    assert r1 == pytest.approx(ref["r1"], rel=RES_TOL[name])
    assert r2 == pytest.approx(ref["r2"], rel=1e-3, abs=1e-6)


@pytest.mark.parametrize("name", ["f32", "f64"])
def test_stop_rule_r1_decides(name):
    """Parity protocol (b): cases where r1 alone decides - sweep count, err and field must match exactly."""
    torch, X, O = _mods()
    dt = DTS[name]
    z = np.load(os.path.join(GOLDEN, f"small_{name}.npz"))
    ny, nx = z["A"].shape
    coe = z["coe"]; f = z["f"].astype(dt); bc = z["bc"].astype(dt)
    dat = bc.copy(); wk = np.zeros_like(dat)
    it, r1, r2, err = X.solve_elliptic(200000, 50, 4, 3, float(z["r1_in"]), 2.0, 0.8, dat, coe, f, wk, nx, ny)
    assert (it, err) == (int(z["stop_sweeps"]), 0)
    assert np.array_equal(dat, z["psi_stop"])              # golden from the numpy restatement, bit for bit
    assert r1 == pytest.approx(float(z["stop_r1"]), rel=RES_TOL[name])
    # 50 sweeps, golden field
    dat = bc.copy()
    it, r1, r2, err = X.solve_elliptic(50, 10, 3, 2, 1e-30, 1.0, 0.8, dat, coe, f, wk, nx, ny)
    assert np.array_equal(dat, z["psi50"]) and (it, err) == (50, 1)


@pytest.mark.parametrize("name", ["f32", "f64"])
def test_test1_reference_case_through_dropin(name):
    """BASELINE config 1: test/test1 (200x200, BAROTROPIC, r1=r2=5e-3) through cal_coe + solve_elliptic + cal_eta."""
    torch, X, O = _mods()
    dt = DTS[name]
    gold = golden_json()["test1"][name]
    A, B, C, bc = ref_test1_inputs()
    d = O.Domain((0.0, 1.0), (0.0, 1.0), 200, 200, 0, 0)
    g = O.geometry(d, dt)
    a, b, c = O.build_abc(A.astype(dt), B.astype(dt), C.astype(dt), d)
    coe = np.zeros((200, 200, 9), dt)
    X.cal_coe(a, np.zeros_like(b), c, coe, g["dr"], g["dz"], 200, 200)
    assert sha(coe) == gold["coe_sha256"]
    f = -(B.astype(dt))
    for sweeps, key in ((100, "psi100_sha256"), (1000, "psi1000_sha256")):
        dat = bc.astype(dt).copy(); wk = np.zeros_like(dat)
        X.solve_elliptic(sweeps, 100, 10, 5, 5e-3, 5e-3, 1.0, dat, coe, f, wk, 200, 200)
        assert sha(dat) == gold[key]
    # run to the stop rule: the stop sweep sits on the round-off floor, so report counts
    # side by side and assert the converged FIELD and eta, not the sweep count.
    dat = bc.astype(dt).copy(); wk = np.zeros_like(dat)
    it, r1, r2, err = X.solve_elliptic(100000, 100, 10, 5, 5e-3, 5e-3, 1.0, dat, coe, f, wk, 200, 200)
    ref = O.solve_elliptic(100000, 100, 10, 5, 5e-3, 5e-3, 1.0, bc.astype(dt), coe, f)
    print(f"test1 {name}: GPU stop sweep {it}, oracle {ref['max_iter']}, r1 {r1:.3e} vs {ref['r1']:.3e}")
    assert err == 0 and ref["err"] == 0
    tol = 1e-8 if name == "f64" else 2e-4
    assert rel_l2(dat, ref["dat"]) < tol
    eta = np.zeros((200, 199), dt)
    import ctypes as C
    X._lib.lib().xee_cal_eta_f64 if name == "f64" else None
    fn = getattr(X._lib.lib(), f"xee_cal_eta_{name}"); fn.restype = None
    p = lambda arr: arr.ctypes.data_as(C.c_void_p)
    fn(p(dat), p(eta), p(g["ra"]), p(g["rcuva"]), p(g["rho"]), p(g["exner"]), C.byref(C.c_int(200)), C.byref(C.c_int(200)))
    assert np.array_equal(eta, O.cal_eta(dat, d))          # K5 bitwise on the same input
    assert float(eta.max()) == pytest.approx(gold["eta_stop_max"], rel=1e-6 if name == "f64" else 1e-3)


@pytest.mark.parametrize("name", ["f32", "f64"])
@pytest.mark.parametrize("shared", [True, False])
def test_batched_plan_matches_oracle_bitwise(name, shared):
    torch, X, O = _mods()
    dt = DTS[name]; nx, ny, nb = 70, 45, 11
    a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=3)
    rng = np.random.default_rng(9)
    F = np.stack([f * dt(k + 1) for k in range(nb)]); P = np.stack([x0 * dt(0.5 * k) for k in range(nb)])
    if shared:
        coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    else:
        coe = np.stack([O.cal_coe(a * dt(1 + 0.1 * k), b, c, 1.0, 0.5, nx, ny)[0] for k in range(nb)])
    plan = X.Plan(nx, ny, nbatch=nb, dtype=name, shared_coe=shared, arith="strict")
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
    # different RHS scales finish at different checks: exercises per-solve done flags
    rms_f = float(np.sqrt((F[0][1:-1, 1:-1].astype(np.float64) ** 2).mean()))
    prm = X.SolveParams(max_iter=4000, check_step=20, converge_time=2, lost_rate=5, r1=2e-3 * rms_f if name == "f64" else 2e-2 * rms_f, r2=0.0, alpha=0.9)
    out = plan.solve(psi, ft, prm)
    rb = O.solve_batch(4000, 20, 2, 5, prm.r1, 0.0, 0.9, P, coe, F, threads=4)
    assert list(out["iters"]) == list(rb["max_iter"]) and list(out["err"]) == list(rb["err"])
    assert len(set(out["iters"])) > 1
    assert np.array_equal(psi.cpu().numpy(), rb["dat"])
    assert np.allclose(out["r1"], rb["r1"], rtol=RES_TOL[name])
    # apply == do_elliptic for the batch
    Lp = plan.apply(torch.from_numpy(P).cuda()).cpu().numpy()
    for k in range(nb):
        assert np.array_equal(Lp[k], O.do_elliptic(P[k], coe if shared else coe[k]))


def test_set_abc_on_device_matches_cal_coe():
    torch, X, O = _mods()
    nx, ny = 90, 50
    a, b, c, f, x0 = _rand_case(nx, ny, np.float64, seed=5)
    coe, _ = O.cal_coe(a, b, c, 0.3, 0.9, nx, ny)
    p1 = X.Plan(nx, ny, 1, "f64"); p1.set_coe_aos(coe)
    p2 = X.Plan(nx, ny, 1, "f64"); p2.set_abc(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda(), 0.3, 0.9)
    x = torch.from_numpy(x0[None]).cuda()
    assert torch.equal(p1.apply(x), p2.apply(x))


@pytest.mark.parametrize("name", ["f32", "f64"])
def test_fast_arithmetic_within_tolerance(name):
    torch, X, O = _mods()
    dt = DTS[name]; nx, ny = 128, 96
    a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=11)
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    plan = X.Plan(nx, ny, 1, name, arith="fast"); plan.set_coe_aos(coe)
    psi = torch.from_numpy(x0[None].copy()).cuda(); ft = torch.from_numpy(f[None].copy()).cuda()
    rms = plan.sweeps(psi, ft, 1.0, 500, want_rms=True)
    ref = O.solve_elliptic(500, 500, 10, 5, 1e-30, 1.0, 1.0, x0, coe, f)
    assert rel_l2(psi.cpu().numpy()[0], ref["dat"]) < (1e-12 if name == "f64" else 1e-4)
    assert rms[0] == pytest.approx(ref["r1"], rel=1e-9 if name == "f64" else 1e-3)


def test_chebyshev_reaches_the_same_solution():
    """Accelerated mode: same discretisation, same residual definition, iterated to the same tolerance."""
    torch, X, O = _mods()
    nx, ny = 160, 96
    a, b, c, f, x0 = _rand_case(nx, ny, np.float64, seed=13, bscale=0.02)
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    rms_f = float(np.sqrt((f[1:-1, 1:-1] ** 2).mean()))
    r1 = 1e-12 * rms_f
    ref = O.solve_elliptic(400000, 100, 3, 5, r1, 0.0, 1.0, x0, coe, f)
    assert ref["err"] == 0
    plan = X.Plan(nx, ny, 1, "f64", arith="fast", method="chebyshev"); plan.set_coe_aos(coe)
    psi = torch.from_numpy(x0[None].copy()).cuda(); ft = torch.from_numpy(f[None].copy()).cuda()
    out = plan.solve(psi, ft, X.SolveParams(max_iter=400000, check_step=100, converge_time=3, r1=r1, r2=0.0, alpha=1.0))
    print("chebyshev sweeps", out["iters"][0], "jacobi sweeps", ref["max_iter"])
    assert out["err"][0] == 0 and out["r1"][0] < r1
    assert out["iters"][0] * 5 < ref["max_iter"]
    assert rel_l2(psi.cpu().numpy()[0], ref["dat"]) < 1e-8


def test_uw_kernels_bitwise():
    torch, X, O = _mods()
    import ctypes as C
    for name, dt in DTS.items():
        d = O.Domain((0.0, 3.0e5), (0.0, 1.0e4), 75, 41, 0, 0)
        g = O.geometry(d, dt)
        rpsi = np.random.default_rng(2).standard_normal((41, 75)).astype(dt)
        u = np.zeros((40, 75), dt); w = np.zeros((41, 74), dt)
        fn = getattr(X._lib.lib(), f"xee_cal_uw_{name}"); fn.restype = None
        p = lambda arr: arr.ctypes.data_as(C.c_void_p)
        fn(p(rpsi), p(u), p(w), p(g["ra"]), p(g["rcuva"]), p(g["za"]), p(g["rho"]), C.byref(C.c_int(75)), C.byref(C.c_int(41)))
        uo, wo = O.cal_uw(rpsi, d)
        assert np.array_equal(u, uo) and np.array_equal(w, wo)
        assert np.all(u[:, 0] == 0)                        # ra(1) == 0 -> u = 0 (quick-tools1.f90:33-37)
        A = np.random.default_rng(3).random((41, 75)).astype(dt)
        a = np.zeros((39, 74), dt); b = np.zeros((40, 74), dt); c = np.zeros((40, 73), dt)
        fn = getattr(X._lib.lib(), f"xee_build_abc_{name}"); fn.restype = None
        fn(p(A), p(A * 2), p(A + 1), p(g["rcuva"]), p(g["rho"]), p(a), p(b), p(c), C.byref(C.c_int(75)), C.byref(C.c_int(41)))
        ao, bo, co = O.build_abc(A, A * 2, A + 1, d)
        assert np.array_equal(a, ao) and np.array_equal(b, bo) and np.array_equal(c, co)


def test_fortran_stop_behaviour_in_the_c_abi():
    """Both criteria non-positive inside the C entry point: message on stdout, process exits 0 (Fortran STOP)."""
    code = ("import ctypes as C, numpy as np\n"
            "from xlab_ee_fortran_b200 import _lib\n"
            "L=_lib.lib(); z=np.zeros((4,4)); c=np.zeros((4,4,9)); p=lambda a:a.ctypes.data_as(C.c_void_p)\n"
            "i=lambda v:C.byref(C.c_int(v)); d=lambda v:C.byref(C.c_double(v))\n"
            "L.xee_solve_elliptic_f64(i(10),i(1),i(1),i(1),d(0.0),d(-1.0),d(1.0),p(z),p(c),p(z),p(z.copy()),i(4),i(4),i(0),i(0))\n"
            "print('NOT REACHED')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "cannot both be non-positive" in r.stdout and "NOT REACHED" not in r.stdout


def test_debug_prints_match_reference_format(capfd):
    torch, X, O = _mods()
    nx, ny = 20, 16
    a, b, c, f, x0 = _rand_case(nx, ny, np.float64, seed=1)
    coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
    dat = x0.copy(); wk = np.zeros_like(dat)
    X.solve_elliptic(30, 10, 10, 5, 1e-30, 1.0, 1.0, dat, coe, f, wk, nx, ny, debug=2)
    out = capfd.readouterr().out
    assert " ----- Solve Elliptic Inputs -----" in out
    assert out.count("Iter: ") == 3 and "Iter:       10, err_now: " in out
    assert " Elliptic Tools: [Error] Max iteration reached." in out


@pytest.mark.parametrize("name", ["f32", "f64"])
@pytest.mark.parametrize("shape,nb", [((140, 70), 20), ((512, 256), 5), ((260, 13), 37), ((8, 4), 3)])
def test_tma_kernel_matches_direct_kernel_and_oracle_bitwise(name, shape, nb):
    """v2 (TMA pipeline, kernel=2) against v1 (direct, kernel=1) and the oracle: STRICT iterates bit-identical,
    partial tiles on every edge, more solves than one work-unit chunk, residual partials per TMA tile."""
    torch, X, O = _mods()
    dt = DTS[name]; nx, ny = shape
    a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=nx + nb)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    F = np.stack([f * dt(k + 1) for k in range(nb)]); P = np.stack([x0 * dt(1 + 0.25 * k) for k in range(nb)])
    res = {}
    for kern in (1, 2):
        plan = X.Plan(nx, ny, nbatch=nb, dtype=name, shared_coe=True, arith="strict", kernel=kern)
        plan.set_coe_aos(coe)
        psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
        out = plan.solve(psi, ft, X.SolveParams(max_iter=57, check_step=10, converge_time=10, r1=1e-30, r2=1.0, alpha=0.9))
        res[kern] = (psi.cpu().numpy(), out)
        plan.close()
    assert np.array_equal(res[1][0], res[2][0])
    assert np.allclose(res[1][1]["r1"], res[2][1]["r1"], rtol=1e-12)
    rb = O.solve_batch(57, 10, 10, 5, 1e-30, 1.0, 0.9, P, coe, F, threads=4)
    assert np.array_equal(res[2][0], rb["dat"]) and list(res[2][1]["iters"]) == list(rb["max_iter"])
    assert np.allclose(res[2][1]["r1"], rb["r1"], rtol=RES_TOL[name])
    # FAST + Chebyshev: the two kernels run the same FMA sequence -> identical bits as well
    for kern in (1, 2):
        plan = X.Plan(nx, ny, nbatch=nb, dtype=name, shared_coe=True, arith="fast", method="chebyshev", kernel=kern)
        plan.set_coe_aos(coe)
        psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
        prm = X.SolveParams(max_iter=2000, check_step=10, converge_time=2, r1=1e-3 if name == "f64" else 1e-1, r2=0.0, rho_jacobi=0.97)
        out = plan.solve(psi, ft, prm)
        res[kern] = (psi.cpu().numpy(), out)
        plan.close()
    assert np.array_equal(res[1][0], res[2][0]) and list(res[1][1]["iters"]) == list(res[2][1]["iters"])


def _run_diagnose(tmp_path, diag_txt, extra_files=(), r8=False, flag_file=None):
    from xlab_ee_fortran_b200 import _lib
    exe = _lib.build_diagnose()
    A, B, C, bc = ref_test1_inputs()
    for n, arr in (("A.bin", A), ("B.bin", B), ("C.bin", C), ("bc_init.bin", bc)) + tuple(extra_files):
        arr.astype(np.float32).tofile(tmp_path / n)
    if flag_file:
        (tmp_path / flag_file).write_text("")
    r = subprocess.run([exe] + (["--r8"] if r8 else []), input=diag_txt, capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_driver_rehost_runs_test1_end_to_end(tmp_path):
    """BASELINE config 1 through the re-hosted driver: the reference's own diag.txt (CRLF and all) on stdin, raw
    float32 .bin files in, the reference's output names/shapes out."""
    torch, X, O = _mods()
    gj = golden_json()
    diag = gj["reference_test1_diag_txt"]
    # (a) deterministic variant: max_iter = 1000 -> bit-identical to the golden 1000-sweep field
    d1000 = diag.replace("100000", "1000")
    out = _run_diagnose(tmp_path, d1000, flag_file="debug_mode_2")
    assert "Iter:      100, err_now: " in out and " Elliptic Tools: [Error] Max iteration reached." in out
    rchi = np.fromfile(tmp_path / "rchi-[BAROTROPIC]-O.bin", np.float32)
    assert rchi.size == 200 * 200 and sha(rchi) == gj["test1"]["f32"]["psi1000_sha256"]
    eta = np.fromfile(tmp_path / "eta-[BAROTROPIC]-A.bin", np.float32)
    assert eta.size == 199 * 200
    d = O.Domain((0.0, 1.0), (0.0, 1.0), 200, 200, 0, 0)
    assert np.array_equal(eta.reshape(200, 199), O.cal_eta(rchi.reshape(200, 200), d))
    A, B, C, bc = ref_test1_inputs()
    a, b, c = O.build_abc(A, B, C, d)
    assert np.array_equal(np.fromfile(tmp_path / "solver_a-sA.bin", np.float32).reshape(198, 199), a)
    assert np.array_equal(np.fromfile(tmp_path / "solver_b-B.bin", np.float32).reshape(199, 199), b)
    assert np.array_equal(np.fromfile(tmp_path / "solver_c-sC.bin", np.float32).reshape(199, 198), c)
    assert (tmp_path / "result.txt").read_text().startswith(" Time elapsed (sec) : ")
    # (b) the reference's own settings, to the stop rule (real(4): the stop sweep sits on the round-off floor)
    (tmp_path / "debug_mode_2").unlink()
    out = _run_diagnose(tmp_path, diag)
    assert " Elliptic Tools: Iteration success." in out
    rchi = np.fromfile(tmp_path / "rchi-[BAROTROPIC]-O.bin", np.float32).reshape(200, 200)
    gold = gj["test1"]["f32"]
    assert np.sqrt((rchi.astype(np.float64) ** 2).sum()) == pytest.approx(gold["psi_stop_l2"], rel=2e-4)
    eta = np.fromfile(tmp_path / "eta-[BAROTROPIC]-A.bin", np.float32)
    assert float(eta.max()) == pytest.approx(gold["eta_stop_max"], rel=2e-3)


def test_driver_rehost_secondary_circulation_baro_all_r8(tmp_path):
    """SECONDARY_CIRCULATION + BARO_ALL (+ forcing file), promoted build: both passes, u/w/rpsi outputs."""
    torch, X, O = _mods()
    A, B, C, bc = ref_test1_inputs()
    forcing = (1e-3 * np.cos(np.linspace(0, 3, 200))[None, :] * np.sin(np.linspace(0, 2, 200))[:, None]).astype(np.float32)
    diag = ("SECONDARY_CIRCULATION-CYLINDRICAL-DENSITY_BOUSSINESQ-BARO_ALL   // mode\n"
            "0.0 2.0 0.0 1.0 // domain\n\n200 200 // grid\n. // in\n. // out\nA.bin // A\nB.bin // B\nC.bin // C\n"
            "forcing.bin // forcing\nbc_init.bin // bc\n1e-30 1.0 300 0.9 // criteria\n")
    out = _run_diagnose(tmp_path, diag, extra_files=(("forcing.bin", forcing),), r8=True)
    assert out.count("Relaxation uses") == 2
    d = O.Domain((0.0, 2.0), (0.0, 1.0), 200, 200, 1, 0)
    g = O.geometry(d, np.float64)
    a, b, c = O.build_abc(A.astype(np.float64), B.astype(np.float64), C.astype(np.float64), d)
    for tag, bb in (("[BAROTROPIC]", np.zeros_like(b)), ("[BAROCLINIC]", b)):
        coe, _ = O.cal_coe(a, bb, c, g["dr"], g["dz"], 200, 200)
        ref = O.solve_elliptic(300, 100, 10, 5, 1e-30, 1.0, 0.9, bc.astype(np.float64), coe, forcing.astype(np.float64))
        u, w = O.cal_uw(ref["dat"], d)
        assert np.array_equal(np.fromfile(tmp_path / f"rpsi-{tag}-O.bin", np.float32).reshape(200, 200), ref["dat"].astype(np.float32))
        assert np.array_equal(np.fromfile(tmp_path / f"w-{tag}-A.bin", np.float32).reshape(200, 199), w.astype(np.float32))
        assert np.array_equal(np.fromfile(tmp_path / f"u-{tag}-C.bin", np.float32).reshape(199, 200), u.astype(np.float32))


def test_chebyshev_with_one_operator_per_solve():
    """Time-series shape (BASELINE config 5): every solve has its own operator, hence its own Jacobi spectral
    radius; the Chebyshev weights are computed per solve inside the kernel."""
    torch, X, O = _mods()
    nx, ny, nb = 96, 64, 5
    rng = np.random.default_rng(21)
    coes, F, P = [], [], []
    for k in range(nb):
        a, b, c, f, x0 = _rand_case(nx, ny, np.float64, seed=100 + k, bscale=0.02)
        coes.append(O.cal_coe(a * (1 + 0.5 * k), b, c * (1 + 2.0 * (nb - k)), 1.0, 0.7, nx, ny)[0]); F.append(f); P.append(x0)
    coe = np.stack(coes); F = np.stack(F); P = np.stack(P)
    rms_f = np.sqrt((F[:, 1:-1, 1:-1] ** 2).mean(axis=(1, 2)))
    plan = X.Plan(nx, ny, nbatch=nb, dtype="f64", shared_coe=False, arith="fast", method="chebyshev")
    plan.set_coe_aos(coe)
    psi = torch.from_numpy(P.copy()).cuda(); ft = torch.from_numpy(F).cuda()
    r1 = torch.from_numpy(1e-11 * rms_f).cuda()
    out = plan.solve(psi, ft, X.SolveParams(max_iter=100000, check_step=50, converge_time=2, r1=1.0, r2=0.0, r1_per_solve=r1))
    assert np.all(out["err"] == 0)
    got = psi.cpu().numpy()
    for k in range(nb):
        ref = O.solve_elliptic(400000, 100, 2, 5, 1e-11 * rms_f[k], 0.0, 1.0, P[k], coe[k], F[k])
        assert ref["err"] == 0 and rel_l2(got[k], ref["dat"]) < 1e-8
        assert out["iters"][k] * 4 < ref["max_iter"]
    print("per-solve chebyshev sweeps", out["iters"])


@pytest.mark.parametrize("name", ["f32", "f64"])
@pytest.mark.parametrize("shape,nb", [((200, 200), 1), ((512, 256), 1), ((140, 70), 2), ((67, 35), 1), ((3, 3), 1), ((700, 5), 1)])
def test_resident_solver_matches_direct_kernel_bitwise(name, shape, nb):
    """v3 (whole solve in one cooperative launch, kernel=3) against v1 (kernel=1): identical sweep counts, residuals
    and bit-identical fields, for a stop decided by r1 and for max_iter, STRICT Jacobi and FAST Chebyshev."""
    torch, X, O = _mods()
    dt = DTS[name]; nx, ny = shape
    a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=7 * nx + nb)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    F = np.stack([f * dt(k + 1) for k in range(nb)]); P = np.stack([x0 * dt(1 + 0.25 * k) for k in range(nb)])
    rms_f = float(np.sqrt((F[0].astype(np.float64) ** 2).mean()))
    cases = [("strict", "jacobi", X.SolveParams(max_iter=237, check_step=10, converge_time=3, r1=1e-30, r2=1.0, alpha=0.9)),
             ("strict", "jacobi", X.SolveParams(max_iter=100000, check_step=20, converge_time=2, lost_rate=3, r1=(3e-2 if name == "f32" else 1e-3) * rms_f, r2=0.0, alpha=1.0)),
             ("fast", "chebyshev", X.SolveParams(max_iter=5000, check_step=10, converge_time=2, r1=(1e-1 if name == "f32" else 1e-4) * rms_f, r2=0.0, rho_jacobi=0.98))]
    for arith, method, prm in cases:
        res = {}
        for kern in (1, 3):
            plan = X.Plan(nx, ny, nbatch=nb, dtype=name, shared_coe=True, arith=arith, method=method, kernel=kern)
            plan.set_coe_aos(coe)
            psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
            out = plan.solve(psi, ft, prm)
            res[kern] = (psi.cpu().numpy(), out)
            plan.close()
        assert list(res[1][1]["iters"]) == list(res[3][1]["iters"]), (arith, method)
        assert list(res[1][1]["err"]) == list(res[3][1]["err"])
        assert np.array_equal(res[1][0], res[3][0]), (arith, method)
        assert np.allclose(res[1][1]["r1"], res[3][1]["r1"], rtol=1e-12 if name == "f64" else 1e-5)
    # the first case also against the oracle, bit for bit
    rb = O.solve_batch(237, 10, 3, 5, 1e-30, 1.0, 0.9, P, coe, F, threads=2)
    plan = X.Plan(nx, ny, nbatch=nb, dtype=name, shared_coe=True, arith="strict", kernel=3); plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); plan.solve(psi, torch.from_numpy(F).cuda(), cases[0][2])
    assert np.array_equal(psi.cpu().numpy(), rb["dat"])


@pytest.mark.parametrize("name", ["f32", "f64"])
@pytest.mark.parametrize("shape,nb", [((140, 70), 7), ((512, 256), 3), ((260, 13), 9)])
def test_tma_kernel_with_one_operator_per_solve(name, shape, nb):
    """v2 with the operator planes travelling through the TMA stage (time-series shape) against v1: bit-identical."""
    torch, X, O = _mods()
    dt = DTS[name]; nx, ny = shape
    coes, F, P = [], [], []
    for k in range(nb):
        a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=1000 + 13 * k + nx)
        coes.append(O.cal_coe(a * dt(1 + 0.3 * k), b, c * dt(1 + 0.2 * k), 1.0, 0.5, nx, ny)[0]); F.append(f); P.append(x0)
    coe = np.stack(coes); F = np.stack(F); P = np.stack(P)
    for arith, method, prm in (("strict", "jacobi", X.SolveParams(max_iter=43, check_step=10, converge_time=10, r1=1e-30, r2=1.0, alpha=0.9)),
                               ("fast", "chebyshev", X.SolveParams(max_iter=60, check_step=10, converge_time=10, r1=1e-30, r2=1.0, rho_jacobi=0.97))):
        res = {}
        for kern in (1, 2):
            plan = X.Plan(nx, ny, nbatch=nb, dtype=name, shared_coe=False, arith=arith, method=method, kernel=kern)
            plan.set_coe_aos(coe)
            psi = torch.from_numpy(P).cuda(); ft = torch.from_numpy(F).cuda()
            out = plan.solve(psi, ft, prm)
            res[kern] = (psi.cpu().numpy(), out)
            plan.close()
        assert np.array_equal(res[1][0], res[2][0]), (arith, method)
        assert np.allclose(res[1][1]["r1"], res[2][1]["r1"], rtol=1e-12 if name == "f64" else 1e-5)
    rb = O.solve_batch(43, 10, 10, 5, 1e-30, 1.0, 0.9, P, coe, F, threads=4)
    plan = X.Plan(nx, ny, nbatch=nb, dtype=name, shared_coe=False, arith="strict", kernel=2); plan.set_coe_aos(coe)
    psi = torch.from_numpy(P).cuda(); plan.solve(psi, torch.from_numpy(F).cuda(), X.SolveParams(max_iter=43, check_step=10, converge_time=10, r1=1e-30, r2=1.0, alpha=0.9))
    assert np.array_equal(psi.cpu().numpy(), rb["dat"])


def test_driver_rehost_spherical_geometry_r8(tmp_path):
    """SPHERICAL geometry (initialize-variables.f90:59-67) through the re-hosted driver, AS THE REFERENCE WRITES IT:
    rcuva(i) = planet_radius * cos(Lat(1) + (i-1) dlat) with the latitudes in DEGREES handed to cos() (the upstream bug is
    kept), Lr = Lat * DEG2RAD * planet_radius (read-input.f90:66-70).  The resulting operator is
    not elliptic, so this is a fixed-sweep parity case: a/b/c, the iterate after 100 sweeps and eta are bit-identical to
    the oracle's restatement (make_geometry(..., geometry = 1))."""
    import math
    torch, X, O = _mods()
    A, B, C, bc = ref_test1_inputs()
    pr = 2.0
    diag = ("DYNAMIC_EFFICIENCY-SPHERICAL-DENSITY_NORMAL-BAROCLINIC   // mode\n"
            f"{pr} 0.0 1.0 // planet radius, z domain\n200 200 // grid\n. // in\n. // out\nA.bin // A\nB.bin // B\nC.bin // C\n"
            "bc_init.bin // bc\n1e-30 1.0 100 0.5 // criteria\n")
    out = _run_diagnose(tmp_path, diag, r8=True)
    assert "Using spherical mode, domain is forced to be global." in out and out.count("Relaxation uses") == 1
    deg2rad = math.acos(-1.0) / 180.0
    d = O.Domain((-90.0 * deg2rad * pr, 90.0 * deg2rad * pr), (0.0, 1.0), 200, 200, 0, 1, pr)
    g = O.geometry(d, np.float64)
    assert g["rcuva"].min() < 0 < g["rcuva"].max()          # cos() of degrees: the curvature radius changes sign (kept as written)
    a, b, c = O.build_abc(A.astype(np.float64), B.astype(np.float64), C.astype(np.float64), d)
    assert np.array_equal(np.fromfile(tmp_path / "solver_a-sA.bin", np.float32).reshape(198, 199), a.astype(np.float32))
    assert np.array_equal(np.fromfile(tmp_path / "solver_b-B.bin", np.float32).reshape(199, 199), b.astype(np.float32))
    assert np.array_equal(np.fromfile(tmp_path / "solver_c-sC.bin", np.float32).reshape(199, 198), c.astype(np.float32))
    coe, _ = O.cal_coe(a, b, c, g["dr"], g["dz"], 200, 200)
    ref = O.solve_elliptic(100, 100, 10, 5, 1e-30, 1.0, 0.5, bc.astype(np.float64), coe, -B.astype(np.float64))
    assert np.isfinite(ref["dat"]).all()
    rchi = np.fromfile(tmp_path / "rchi-[BAROCLINIC]-O.bin", np.float32).reshape(200, 200)
    assert np.array_equal(rchi, ref["dat"].astype(np.float32))
    eta = np.fromfile(tmp_path / "eta-[BAROCLINIC]-A.bin", np.float32).reshape(200, 199)
    assert np.array_equal(eta, O.cal_eta(ref["dat"], d).astype(np.float32))


@pytest.mark.parametrize("max_iter,expect_err,expect_done", [(250, 0, False), (50, 0, False), (300, 1, True)])
def test_legacy_solve_elliptic_runs_all_max_iter_sweeps(max_iter, expect_err, expect_done):
    """Legacy 12-argument solve_elliptic (src/old-diagnose/xtt-lib/elliptic_tools.f90:93-300): `do cnt = 1, max_iter` runs ALL
    sweeps, the stop tests and `cnt == max_iter` are evaluated on check sweeps (every 100th) only.  A run that is not
    converged and whose max_iter is not a multiple of 100 falls out of the loop: max_iter sweeps done, err = 0, strategy and
    strategy_r untouched; with a multiple of 100 the last check sets err_over_max_iteration and writes both."""
    import ctypes as C
    torch, X, O = _mods()
    nx, ny = 48, 36
    a, b, c, f, x0 = _rand_case(nx, ny, np.float64, seed=31)
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    dat = x0.copy(); wk = np.zeros_like(dat)
    strategy = C.c_int(1); sr = C.c_double(1e-30); err = C.c_int(-7)
    p = lambda arr: arr.ctypes.data_as(C.c_void_p)
    X._lib.lib().xee_solve_elliptic_old_f64(C.byref(C.c_int(max_iter)), C.byref(strategy), C.byref(sr), C.byref(C.c_double(0.9)),
                                            p(dat), p(coe), p(f), p(wk), C.byref(C.c_int(nx)), C.byref(C.c_int(ny)), C.byref(err),
                                            C.byref(C.c_int(0)))
    ref = O.solve_elliptic(max_iter, 100, 1, 5, 1e-30, 0.0, 0.9, x0, coe, f)      # the same max_iter sweeps of the same iteration
    assert ref["max_iter"] == max_iter
    assert np.array_equal(dat, ref["dat"])
    assert err.value == expect_err
    if expect_done:
        assert strategy.value == max_iter and sr.value == pytest.approx(ref["r1"], rel=1e-12)
    else:
        assert strategy.value == 1 and sr.value == 1e-30


def test_pin_checker_accepts_the_rehosted_driver_output(tmp_path):
    """scripts/pin_check.py is what will judge the REAL reference's output the day a Fortran compiler exists
    (scripts/pin_against_reference.sh).  Until then it is exercised on the re-hosted driver (same stdin, same files, same
    debug lines, every solve on the GPU): the checker must parse that output and find every pin green."""
    import subprocess
    import sys
    gj = golden_json()
    diag = gj["reference_test1_diag_txt"]
    for case, txt in (("sweeps1000", diag.replace("100000", "1000")), ("stop", diag)):
        dd = tmp_path / case; dd.mkdir()
        out = _run_diagnose(dd, txt, flag_file="debug_mode_2")
        (dd / "stdout.txt").write_text(out)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "pin_check.py"), str(tmp_path / "sweeps1000"), str(tmp_path / "stop")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "PINNED" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("strategy,sr,zero_rim", [(3, 1e-2, True), (3, 1e-3, True), (4, 0.3, True), (4, 0.05, False), (3, 1e-3, False),
                                                 (1, 1e-3, False), (2, 0.3, True)])
def test_legacy_solve_elliptic_strategies(dt, strategy, sr, zero_rim):
    """Legacy strategies 1..4 (src/old-diagnose/xtt-lib/elliptic_tools.f90:190-204, :246-279) through the drop-in against a
    restatement of the legacy loop: 3 / 4 measure maxval(abs(to_dat)) - the largest interior residual joined with the largest
    Dirichlet value on the rim (so with a non-zero rim strategy 3 never converges and strategy 4 sees a constant).  Sweeps used,
    returned residual, error code and field bit for bit."""
    import ctypes as C
    from tests import legacy_oracle as L
    torch, X, O = _mods()
    nx, ny = 48, 36
    a, b, c, f, x0 = _rand_case(nx, ny, dt, seed=31)
    if zero_rim:
        x0[0] = 0; x0[-1] = 0; x0[:, 0] = 0; x0[:, -1] = 0
    coe, _ = O.cal_coe(a, b, c, 1.0, 0.5, nx, ny)
    ref = L.old_solve_loop(strategy, sr, 3000, 0.9, x0, coe, f)
    dat = x0.copy(); wk = np.zeros_like(dat)
    ct = C.c_double if dt is np.float64 else C.c_float
    st = C.c_int(strategy); srv = ct(sr); err = C.c_int(-7)
    p = lambda arr: arr.ctypes.data_as(C.c_void_p)
    fn = X._lib.lib().xee_solve_elliptic_old_f64 if dt is np.float64 else X._lib.lib().xee_solve_elliptic_old_f32
    fn(C.byref(C.c_int(3000)), C.byref(st), C.byref(srv), C.byref(ct(0.9)), p(dat), p(coe), p(f), p(wk),
       C.byref(C.c_int(nx)), C.byref(C.c_int(ny)), C.byref(err), C.byref(C.c_int(0)))
    assert st.value == ref["strategy"] and err.value == ref["err"]
    if strategy >= 3:      # a maximum is exact in any order
        assert dt(srv.value) == dt(ref["strategy_r"])
    else:                  # the RMS is a sum: sequential in the reference, a fixed tree on the device
        assert srv.value == pytest.approx(ref["strategy_r"], rel=1e-12 if dt is np.float64 else 2e-5)
    assert np.array_equal(dat, ref["dat"])


def test_legacy_solve_elliptic_rejects_unknown_strategy():
    import ctypes as C
    torch, X, O = _mods()
    nx, ny = 16, 12
    z = np.zeros((ny, nx)); coe = np.zeros((ny, nx, 9)); wk = np.zeros_like(z)
    st = C.c_int(5); sr = C.c_double(1e-3); err = C.c_int(0)
    p = lambda arr: arr.ctypes.data_as(C.c_void_p)
    X._lib.lib().xee_solve_elliptic_old_f64(C.byref(C.c_int(100)), C.byref(st), C.byref(sr), C.byref(C.c_double(1.0)), p(z), p(coe), p(z), p(wk),
                                            C.byref(C.c_int(nx)), C.byref(C.c_int(ny)), C.byref(err), C.byref(C.c_int(0)))
    assert err.value == 1 << 8 and st.value == 5
