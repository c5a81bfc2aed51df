"""GPU parity of the efficiency-map chain (K1, K2, K7, solve, K5, K6) against the oracle composition.
Tolerances: streamfunction 1e-8 relative L2, efficiency 1e-6 relative (north_star)."""
import numpy as np
import pytest

from tests.map_oracle import efficiency_rows
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _setup(nr, nz):
    from xlab_ee_fortran_b200 import workloads as W
    Lr, Lz = (0.0, 6.0e5), (0.0, 1.5e4)
    A, B, C = W.vortex_fields(nr, nz, Lr, Lz)
    dr, dz = Lr[1] / (nr - 1), Lz[1] / (nz - 1)
    heat = W.heating_lattice(3, 2, Lr, Lz, 3 * dr, 3 * dz, r_frac=(0.02, 0.5), z_frac=(0.15, 0.75))
    return A, B, C, Lr, Lz, heat


def test_map_jacobi_strict_matches_oracle():
    import xlab_ee_fortran_b200 as X
    from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap
    nr, nz = 48, 40
    A, B, C, Lr, Lz, heat = _setup(nr, nz)
    kw = dict(max_iter=400000, check_step=100, converge_time=2, r1_rel=1e-11)
    ref, psi_ref, f_ref, th_ref, eta_ref = efficiency_rows(A, B, C, Lr, Lz, heat, np.float64, kw)
    m = EfficiencyMap(A, B, C, Lr, Lz, len(heat), "f64", arith="strict", method="jacobi", adjoint_check=True, r1_rel=1e-11)
    tab = m.run(heat, X.SolveParams(max_iter=400000, check_step=100, converge_time=2, r1=1.0, r2=0.0, alpha=1.0))
    f = m.field("f"); psi = m.field("psi")
    assert np.array_equal(m.field("theta"), th_ref)                      # background theta: bitwise
    assert rel_l2(f, f_ref) < 1e-13                                     # K7 (exp differs by an ulp)
    assert list(tab[:, 0]) == list(ref[:, 0]) and np.all(tab[:, 2] == 0)  # same sweep counts, err = 0
    for n in range(len(heat)):
        assert rel_l2(psi[n], psi_ref[n]) < 1e-8
    assert np.allclose(tab[:, 3], ref[:, 3], rtol=1e-12)                 # sum_Q
    assert np.allclose(tab[:, 5], ref[:, 5], rtol=1e-6)                  # efficiency (wtheta route)
    assert np.allclose(tab[:, 7], ref[:, 7], rtol=1e-6)                  # efficiency (eta route)
    assert rel_l2(m.field("eta"), eta_ref) < 1e-8
    print("efficiency (w theta):", tab[:, 5], "\nefficiency (eta)    :", tab[:, 7])


def test_map_chebyshev_fast_matches_oracle_and_is_much_cheaper():
    import xlab_ee_fortran_b200 as X
    from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap
    nr, nz = 64, 48
    A, B, C, Lr, Lz, heat = _setup(nr, nz)
    kw = dict(max_iter=2000000, check_step=100, converge_time=2, r1_rel=1e-12)
    ref, psi_ref, _, _, _ = efficiency_rows(A, B, C, Lr, Lz, heat, np.float64, kw, adjoint=False)
    m = EfficiencyMap(A, B, C, Lr, Lz, len(heat), "f64", arith="fast", method="chebyshev", r1_rel=1e-12)
    tab = m.run(heat, X.SolveParams(max_iter=200000, check_step=50, converge_time=2, r1=1.0, r2=0.0, alpha=1.0))
    psi = m.field("psi")
    print("sweeps chebyshev", tab[:, 0], "jacobi", ref[:, 0])
    assert np.all(tab[:, 2] == 0) and np.all(tab[:, 0] * 4 < ref[:, 0])
    for n in range(len(heat)):
        assert rel_l2(psi[n], psi_ref[n]) < 1e-8
    assert np.allclose(tab[:, 5], ref[:, 5], rtol=1e-6)


def test_map_device_resident_entry_equals_host_entry():
    import torch
    import xlab_ee_fortran_b200 as X
    from xlab_ee_fortran_b200.efficiency_map import EfficiencyMap
    nr, nz = 40, 32
    A, B, C, Lr, Lz, heat = _setup(nr, nz)
    prm = X.SolveParams(max_iter=50000, check_step=50, converge_time=2, r1=1.0, r2=0.0)
    m = EfficiencyMap(A, B, C, Lr, Lz, len(heat), "f64", r1_rel=1e-10)
    t_host = m.run(heat, prm)
    td = torch.zeros((len(heat), 8), dtype=torch.float64, device="cuda")
    m.run_dev(torch.from_numpy(heat).cuda(), td, prm)
    assert np.array_equal(td.cpu().numpy(), t_host)
