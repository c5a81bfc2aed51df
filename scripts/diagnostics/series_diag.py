"""Residual against sweeps for a few snapshots of the series and three methods (fixed max_iter runs).
   python scripts/diagnostics/series_diag.py <first snapshot> <count>"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import xlab_ee_fortran_b200 as X
from xlab_ee_fortran_b200 import workloads as W
from xlab_ee_fortran_b200.time_series import TimeSeries
import bench
first = int(sys.argv[1]); n = int(sys.argv[2])
prm = W.series_params(n, total=1024, first=first)
for method, cs in (("line2_chebyshev", 10), ("line_chebyshev", 25), ("chebyshev", 100)):
    ts = TimeSeries(bench.NR, bench.NZ, bench.LR, bench.LZ, n, "f64", arith="fast", method=method, r1_rel=1e-12)
    for mi in (cs, 4 * cs, 16 * cs, 64 * cs, 256 * cs):
        p = X.SolveParams(max_iter=mi, check_step=cs, converge_time=2, r1=1.0, r2=0.0, sync_every=2)
        tab = ts.run(prm, p)
        print(method, "max_iter", mi, "iters", tab[:, 0].astype(int).tolist(), "err", tab[:, 2].astype(int).tolist(), "res", ["%.2e" % v for v in tab[:, 1]], flush=True)
    ts.close()
