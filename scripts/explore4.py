import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import xlab_ee_fortran_b200 as X
from xlab_ee_fortran_b200 import workloads as W
from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap
nr, nz = 512, 256
Lr, Lz = (0.0, 1.0e6), (0.0, 1.5e4)
A, B, C = W.vortex_fields(nr, nz, Lr, Lz)
allrows = W.heating_lattice(64, 64, Lr, Lz, 2 * Lr[1] / (nr - 1), 2 * Lz[1] / (nz - 1))
for rel in (1e-12, 3e-12, 1e-11):
    m = EfficiencyMap(A, B, C, Lr, Lz, 4096, "f64", arith="fast", method="chebyshev", r1_rel=rel)
    tab = m.run(allrows, X.SolveParams(max_iter=2000000, check_step=100, converge_time=2, r1=1.0, r2=0.0, sync_every=2, stall_checks=10))
    st = np.where(tab[:, 2] != 0)[0]
    print(f"r1_rel={rel:g}: stalled {len(st)} of 4096; sweeps min/mean/max {tab[:,0].min():.0f}/{tab[:,0].mean():.0f}/{tab[:,0].max():.0f}; stalled r1/r1_target:",
          [(int(i), f"{tab[i,1]:.2e}") for i in st[:6]], flush=True)
    m.close()
