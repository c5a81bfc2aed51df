"""Measures the constants bench.py needs and cannot measure inside its own time budget, on the GPU:
the number of sweeps the REFERENCE algorithm (weighted Jacobi, alpha = 1, STRICT arithmetic = bit-identical
iterates) needs to reach the bench tolerance r1 = 1e-12 * rms(f) on the bench workload (512x256 vortex).
Writes profiles/workload_constants.json.   python scripts/measure_constants.py [n_locations]
                                          python scripts/measure_constants.py series [n_snapshots]   (BASELINE config 5)
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import xlab_ee_fortran_b200 as X
from xlab_ee_fortran_b200 import workloads as W
from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap
import bench

path = os.path.join(ROOT, "profiles", "workload_constants.json")
if len(sys.argv) > 1 and sys.argv[1] == "series":
    # one operator per solve: a sample of the 1024-snapshot series, spread over its whole length
    from xlab_ee_fortran_b200.time_series import TimeSeries
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    prm = np.concatenate([W.series_params(1, total=1024, first=int(q)) for q in np.linspace(0, 1023, n).astype(int)])
    ts = TimeSeries(bench.NR, bench.NZ, bench.LR, bench.LZ, n, "f64", arith="strict", method="jacobi", r1_rel=bench.R1_REL)
    t = time.time()
    tab = ts.run(prm, X.SolveParams(max_iter=20000000, check_step=100, converge_time=2, r1=1.0, r2=0.0, sync_every=3, stall_checks=50))
    assert np.all((tab[:, 2] == 0) | (tab[:, 2] == 4)), tab[:, 2]
    print("series jacobi sweeps", tab[:, 0], "seconds", time.time() - t, flush=True)
    ts.close()
    k = json.load(open(path)) if os.path.exists(path) else {}
    k.update({"jacobi_sweeps_to_tol_series": float(tab[:, 0].mean()), "jacobi_sweeps_range_series": [float(tab[:, 0].min()), float(tab[:, 0].max())],
              "n_snapshots_sampled": n,
              "how_series": "scripts/measure_constants.py series on a B200: strict-arithmetic GPU Jacobi (iterates bit-identical to the "
                            "reference) to r1=1e-12*rms(initial residual), check_step 100, converge_time 2, snapshots spread over the 1024-long series"})
    json.dump(k, open(path, "w"), indent=1)
    print("wrote", path)
    sys.exit(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A, B, C = W.vortex_fields(bench.NR, bench.NZ, bench.LR, bench.LZ)
rows = bench.heat_rows(4096)
pick = rows[:n]   # heat_rows() is already a scrambled, uniform sample of the 64x64 lattice
out = {}
for method in ("jacobi", "chebyshev"):
    m = EfficiencyMap(A, B, C, bench.LR, bench.LZ, n, "f64", arith="strict" if method == "jacobi" else "fast", method=method, r1_rel=bench.R1_REL)
    t = time.time()
    tab = m.run(pick, X.SolveParams(max_iter=20000000, check_step=100, converge_time=2, r1=1.0, r2=0.0, sync_every=3, stall_checks=50))
    assert np.all((tab[:, 2] == 0) | (tab[:, 2] == 4))
    out[method] = dict(mean=float(tab[:, 0].mean()), min=float(tab[:, 0].min()), max=float(tab[:, 0].max()), seconds=time.time() - t)
    print(method, out[method], flush=True)
    m.close()
k = json.load(open(path)) if os.path.exists(path) else {}
k.update({"jacobi_sweeps_to_tol": out["jacobi"]["mean"], "jacobi_sweeps_range": [out["jacobi"]["min"], out["jacobi"]["max"]],
          "chebyshev_sweeps_to_tol": out["chebyshev"]["mean"], "n_locations_sampled": n,
          "how": "scripts/measure_constants.py on a B200: strict-arithmetic GPU Jacobi (iterates bit-identical to the reference) "
                 "to r1=1e-12*rms(f), check_step 100, converge_time 2, sample of the 4096-location lattice"})
os.makedirs(os.path.dirname(path), exist_ok=True)
json.dump(k, open(path, "w"), indent=1)
print("wrote", path)
