import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xlab_ee_fortran_b200 as X
from oracle import oracle as O
nx, ny = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
a = 1.0 + rng.random((ny - 2, nx - 1)); c = 1.0 + rng.random((ny - 1, nx - 2)); b = 0.02 * rng.standard_normal((ny - 1, nx - 1))
coe, _ = O.cal_coe(a, b, c, 1.0, 1.0, nx, ny)
f = rng.standard_normal((1, ny, nx)); x0 = rng.standard_normal((1, ny, nx))
for kern in (1, 2, 3):
    for arith in ("strict", "fast"):
        try:
            plan = X.Plan(nx, ny, 1, "f64", shared_coe=True, arith=arith, method="jacobi", kernel=kern); plan.set_coe_aos(coe)
        except Exception as e:
            print(kern, arith, "n/a", str(e)[:80]); continue
        psi = torch.from_numpy(x0.copy()).cuda(); ft = torch.from_numpy(f).cuda()
        prm = X.SolveParams(max_iter=20000, check_step=100, r1=1e-300, r2=0.0, sync_every=3)
        plan.solve(psi, ft, prm)
        psi = torch.from_numpy(x0.copy()).cuda()
        torch.cuda.synchronize(); t = time.time(); out = plan.solve(psi, ft, prm); torch.cuda.synchronize(); dt = time.time() - t
        print(f"{nx}x{ny} kernel={kern} {arith:6s}: {dt/out['iters'][0]*1e6:7.2f} us/sweep ({out['iters'][0]} sweeps, {dt:.3f}s) checksum {float(psi.sum()):.12e}", flush=True)
        plan.close()
