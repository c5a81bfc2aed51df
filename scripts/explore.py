"""GPU exploration: sweep counts and kernel bandwidth on the 512x256 efficiency-map workload."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xlab_ee_fortran_b200 as X
from xlab_ee_fortran_b200 import workloads as W
from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap

nr, nz = 512, 256
Lr, Lz = (0.0, 1.0e6), (0.0, 1.5e4)
A, B, C = W.vortex_fields(nr, nz, Lr, Lz)
dr, dz = Lr[1] / (nr - 1), Lz[1] / (nz - 1)
what = sys.argv[1:] or ["bw", "cheb", "jacobi"]

if "bw" in what:
    for nb in (64, 512):
        for arith in ("fast", "strict"):
            for method in ("jacobi", "chebyshev"):
                heat = W.heating_lattice(nb, 1, Lr, Lz, 2 * dr, 2 * dz)
                m = EfficiencyMap(A, B, C, Lr, Lz, nb, "f64", arith=arith, method=method, r1_rel=1e-30)
                prm = X.SolveParams(max_iter=300, check_step=100, converge_time=2, r1=1.0, r2=0.0)
                m.run(heat, prm); m.sweep_kernel_stats(reset=True)
                t = time.time(); m.run(heat, prm); wall = time.time() - t
                ms, n = m.sweep_kernel_stats()
                pts = (nr - 2) * (nz - 2) * nb
                byt = 24 if method == "jacobi" else 32
                print(f"nb={nb} {arith:6s} {method:9s}: {ms/n*1e3:8.1f} us/sweep  {pts*byt/(ms/n*1e-3)/1e9:8.1f} GB/s alg ({byt} B/pt)  wall {wall:.3f}s for {n} sweeps", flush=True)
                m.close()

if "cheb" in what:
    heat = W.heating_lattice(8, 8, Lr, Lz, 2 * dr, 2 * dz)
    m = EfficiencyMap(A, B, C, Lr, Lz, 64, "f64", arith="fast", method="chebyshev", r1_rel=1e-12)
    for rho in (0.0,):
        t = time.time()
        tab = m.run(heat, X.SolveParams(max_iter=400000, check_step=100, converge_time=2, r1=1.0, r2=0.0, rho_jacobi=rho))
        print(f"chebyshev rho={rho}: sweeps min/max {tab[:,0].min():.0f}/{tab[:,0].max():.0f} err {tab[:,2].max():.0f} time {time.time()-t:.2f}s eff range {tab[:,5].min():.3e} {tab[:,5].max():.3e}", flush=True)
    psi_c = m.field("psi")[:4].copy()
    m.close()

if "jacobi" in what:
    heat4 = W.heating_lattice(8, 8, Lr, Lz, 2 * dr, 2 * dz)[:4]
    m = EfficiencyMap(A, B, C, Lr, Lz, 4, "f64", arith="fast", method="jacobi", r1_rel=1e-12)
    t = time.time()
    tab = m.run(heat4, X.SolveParams(max_iter=20000000, check_step=1000, converge_time=2, r1=1.0, r2=0.0, sync_every=3))
    print(f"jacobi: sweeps {tab[:,0]} err {tab[:,2]} time {time.time()-t:.2f}s", flush=True)
    if "cheb" in what:
        psi_j = m.field("psi")
        for n in range(4):
            print("rel l2 cheb vs jacobi", np.linalg.norm(psi_c[n] - psi_j[n]) / np.linalg.norm(psi_j[n]))
