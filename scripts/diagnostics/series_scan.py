"""Scans shards of the 1024-snapshot series (BASELINE config 5) on one GPU: sweeps, error flags and time per shard and method.
   python scripts/diagnostics/series_scan.py 0 640 896      (first snapshot of every 128-snapshot shard to scan)
This is how the non-converging shard of the first 8-GPU run was found (stage C)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import xlab_ee_fortran_b200 as X
from xlab_ee_fortran_b200 import workloads as W
from xlab_ee_fortran_b200.time_series import TimeSeries
import bench
for first in [int(a) for a in sys.argv[1:]]:
    prm = W.series_params(128, total=1024, first=first)
    for method, cs in (("line2_chebyshev", 10), ("line_chebyshev", 25)):
        ts = TimeSeries(bench.NR, bench.NZ, bench.LR, bench.LZ, 128, "f64", arith="fast", method=method, r1_rel=1e-12)
        p = X.SolveParams(max_iter=200000, check_step=cs, converge_time=2, r1=1.0, r2=0.0, sync_every=2, stall_checks=20)
        t = time.time(); tab = ts.run(prm, p); dt = time.time() - t
        it = tab[:, 0]
        print(first, method, "sec %.3f" % dt, "sweeps", it.min(), np.median(it), it.max(), "err", sorted(set(tab[:, 2].astype(int))), "argmax", int(it.argmax()), "r1/min", tab[:, 1].max(), flush=True)
        if it.max() > 2000:
            print("   slow solves:", np.nonzero(it > 2000)[0].tolist(), it[it > 2000].tolist())
        ts.close()
