import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xlab_ee_fortran_b200 as X
from xlab_ee_fortran_b200 import workloads as W
from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap
nr, nz, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
Lr, Lz = (0.0, 1.0e6), (0.0, 1.5e4)
A, B, C = W.vortex_fields(nr, nz, Lr, Lz)
heat = W.heating_lattice(nb, 1, Lr, Lz, 2 * Lr[1] / (nr - 1), 2 * Lz[1] / (nz - 1))
for kern in ("1", "2"):
    os.environ["XEE_KERNEL"] = kern
    m = EfficiencyMap(A, B, C, Lr, Lz, nb, "f64", arith="fast", method="chebyshev", r1_rel=1e-30)
    prm = X.SolveParams(max_iter=300, check_step=100, converge_time=2, r1=1.0, r2=0.0)
    m.run(heat, prm); m.sweep_kernel_stats(reset=True)
    t = time.time(); m.run(heat, prm); wall = time.time() - t
    ms, n = m.sweep_kernel_stats()
    pts = (nr - 2) * (nz - 2) * nb
    print(f"{nr}x{nz} nb={nb} kernel={kern}: {ms/n*1e3:8.1f} us/sweep  {pts*32/(ms/n*1e-3)/1e9:8.1f} GB/s alg  wall {wall:.3f}s for {n} sweeps", flush=True)
    m.close()
