"""Runs BASELINE configs 2, 3 and 5 at full size on one GPU, with size-independent property checks
(independent residual through apply(), linearity, boundary rows) and timings.  python scripts/run_configs.py [2 3 5]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import xlab_ee_fortran_b200 as X
from xlab_ee_fortran_b200 import workloads as W
from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap
from xlab_ee_fortran_b200.time_series import TimeSeries

ACC = "chebyshev"      # accelerated method of configs 2 and 3: --method line_chebyshev selects the block-line kernel
argv = sys.argv[1:]
if "--method" in argv:
    k = argv.index("--method"); ACC = argv[k + 1]; del argv[k:k + 2]
CS = 10 if ACC.startswith("line2") else 25 if ACC.startswith("line") else 100
which = [int(a) for a in argv] or [2, 3, 5]
out = {}
Lr, Lz = (0.0, 1.0e6), (0.0, 1.5e4)

def sync(): torch.cuda.synchronize()

if 2 in which:   # single large solve: 512x256, one heating source, fp64
    nr, nz = 512, 256
    A, B, C = W.vortex_fields(nr, nz, Lr, Lz)
    heat = np.array([[4.0e4, 5.0e3, 1.0e4, 2.0e3, 3.5 * 287.0 * 10.0 / 86400.0]])
    res = {}
    for method, arith in ((ACC, "fast"), ("jacobi", "strict")):
        m = EfficiencyMap(A, B, C, Lr, Lz, 1, "f64", arith=arith, method=method, r1_rel=1e-12, adjoint_check=False)
        prm = X.SolveParams(max_iter=5000000, check_step=CS if method == ACC else 100, converge_time=2, r1=1.0, r2=0.0, sync_every=3, stall_checks=20)
        m.run(heat, prm)
        t = time.time(); tab = m.run(heat, prm); dt = time.time() - t
        res[method] = dict(seconds=dt, sweeps=int(tab[0, 0]), err=int(tab[0, 2]), efficiency=float(tab[0, 5]), us_per_sweep=dt / tab[0, 0] * 1e6)
        res[method + "_psi"] = m.field("psi")[0]; f = m.field("f")[0]
        m.close()
    rel = float(np.linalg.norm(res[ACC + "_psi"] - res["jacobi_psi"]) / np.linalg.norm(res["jacobi_psi"]))
    out["config2"] = {k: v for k, v in res.items() if not k.endswith("_psi")}
    out["config2"]["rel_l2_accelerated_vs_jacobi"] = rel
    out["config2"]["rel_efficiency_accelerated_vs_jacobi"] = abs(res[ACC]["efficiency"] - res["jacobi"]["efficiency"]) / abs(res["jacobi"]["efficiency"])
    print("config 2:", json.dumps(out["config2"]), flush=True)

if 3 in which:   # map: 64x32 heating-location sweep on 256x128, one GPU
    nr, nz = 256, 128
    A, B, C = W.vortex_fields(nr, nz, Lr, Lz)
    dr, dz = Lr[1] / (nr - 1), Lz[1] / (nz - 1)
    heat = W.heating_lattice(64, 32, Lr, Lz, 2 * dr, 2 * dz)
    m = EfficiencyMap(A, B, C, Lr, Lz, len(heat), "f64", arith="fast", method=ACC, r1_rel=1e-12, adjoint_check=True)
    prm = X.SolveParams(max_iter=2000000, check_step=CS, converge_time=2, r1=1.0, r2=0.0, sync_every=2, stall_checks=20)
    m.run(heat, prm)
    t = time.time(); tab = m.run(heat, prm); dt = time.time() - t
    psi = m.field("psi"); f = m.field("f")
    # independent residual check on a few solves through apply() (do_elliptic): ||L psi - f|| <= tolerance
    plan = X.Plan(nr, nz, 8, "f64", shared_coe=True, arith="strict")
    a, b, c = None, None, None
    from oracle import oracle as O   # scripts/ may use the checker; the product path above does not
    d = O.Domain(Lr, Lz, nr, nz); g = O.geometry(d, np.float64)
    a, b, c = O.build_abc(A.astype(np.float64), B.astype(np.float64), C.astype(np.float64), d)
    coe, _ = O.cal_coe(a, b, c, g["dr"], g["dz"], nr, nz)
    plan.set_coe_aos(coe)
    idx = np.linspace(0, len(heat) - 1, 8).astype(int)
    Lp = plan.apply(torch.from_numpy(psi[idx]).cuda()).cpu().numpy()
    res_rms = np.sqrt(((Lp - f[idx])[:, 1:-1, 1:-1] ** 2).mean(axis=(1, 2))); f_rms = np.sqrt((f[idx][:, 1:-1, 1:-1] ** 2).mean(axis=(1, 2)))
    out["config3"] = dict(seconds=dt, solves=len(heat), solves_per_s=len(heat) / dt, sweeps=[float(tab[:, 0].min()), float(tab[:, 0].max())],
                          err_max=int(tab[:, 2].max()), independent_residual_over_rms_f=float((res_rms / f_rms).max()),
                          efficiency_range=[float(tab[:, 5].min()), float(tab[:, 5].max())],
                          adjoint_vs_direct_max_abs_diff=float(np.abs(tab[:, 7] - tab[:, 5]).max()))
    print("config 3:", json.dumps(out["config3"]), flush=True)
    m.close(); plan.close()

if 5 in which:   # time series: 128 of the 1024 snapshots (one GPU's share of 8) on 512x256, one operator per snapshot
    nr, nz, ns = 512, 256, 128
    params = W.series_params(ns, total=1024, first=0)
    ts = TimeSeries(nr, nz, Lr, Lz, ns, "f64", arith="fast", method=ACC, r1_rel=1e-12)
    prm = X.SolveParams(max_iter=2000000, check_step=CS, converge_time=2, r1=1.0, r2=0.0, sync_every=2, stall_checks=20)
    ts.run(params, prm); ts.sweep_kernel_stats(reset=True)
    t = time.time(); tab = ts.run(params, prm); dt = time.time() - t
    ms, n = ts.sweep_kernel_stats()
    pts = (nr - 2) * (nz - 2)
    # psi r/w, psi_{k-1}, f, 9 coefficients (+ 2 factors in working precision and 4 float planes for the block-line method)
    alg = float(tab[:, 0].sum()) * pts * 8 * (17 if ACC.startswith("line") else 13)
    out["config5"] = dict(seconds=dt, snapshots=ns, solves_per_s=ns / dt, sweeps=[float(tab[:, 0].min()), float(tab[:, 0].max())],
                          err_max=int(tab[:, 2].max()), sweep_kernel_ms=ms, algorithmic_GBps=alg / (ms * 1e-3) / 1e9,
                          w_absmax_range=[float(tab[:, 6].min()), float(tab[:, 6].max())], efficiency_range=[float(tab[:, 5].min()), float(tab[:, 5].max())])
    print("config 5:", json.dumps(out["config5"]), flush=True)
    ts.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
