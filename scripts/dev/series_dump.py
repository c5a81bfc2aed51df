import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import xlab_ee_fortran_b200 as X
from xlab_ee_fortran_b200 import workloads as W
from xlab_ee_fortran_b200.time_series import TimeSeries
import bench
out = {}
for snap in (100, 688, 700):
    prm = W.series_params(1, total=1024, first=snap)
    ts = TimeSeries(bench.NR, bench.NZ, bench.LR, bench.LZ, 1, "f64", arith="fast", method="line_chebyshev", r1_rel=1e-12)
    for mi in (25, 400, 1600):
        tab = ts.run(prm, X.SolveParams(max_iter=mi, check_step=25, converge_time=2, r1=1.0, r2=0.0))
        print(snap, "one-level max_iter", mi, "res %.3e" % tab[0, 1], flush=True)
    A, B, C = ts.field("A")[0], ts.field("B")[0], ts.field("C")[0]
    hA, hB, hC, bottom, F = W.series_fields_host(prm[0], bench.NR, bench.NZ, bench.LR, bench.LZ)
    for nm, dv, hv in (("A", A, hA), ("B", B, hB), ("C", C, hC)):
        diff = np.abs(dv - hv.astype(np.float64)); rel = diff / np.maximum(np.abs(hv), 1e-300)
        print("   field", nm, "max rel diff device vs host %.3e" % rel.max(), "at", np.unravel_index(rel.argmax(), rel.shape))
    out["A%d" % snap] = A.astype(np.float32); out["B%d" % snap] = B.astype(np.float32); out["C%d" % snap] = C.astype(np.float32)
    out["f%d" % snap] = ts.field("f")[0]; out["psi0_%d" % snap] = ts.field("psi")[0][0]
    ts.close()
np.savez_compressed("gpurun_out/series_fields.npz", **out)
