"""Prototype (CPU, scipy) of the MULTIPLICATIVE two-grid cycle on the bench operator:

    cycle:  r = f - L x;  x += P Ac^-1 P^T r          (bilinear coarse space, Galerkin Ac)
            nu Chebyshev-accelerated block-line sweeps on the interval [lmax/eta, lmax] of M^-1 L (restarted every cycle)

and the number of cycles / fine-grid passes to rms(r) < 1e-12 rms(f) for heating right-hand sides.
    python scripts/prototypes/twogrid_cycle.py 512 256
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

sys.path.insert(0, '/root/repo')
from oracle import oracle as O
from xlab_ee_fortran_b200 import workloads as W
from tests import map_oracle as MO

nr, nz = int(sys.argv[1]), int(sys.argv[2])
LR, LZ = (0.0, 1.0e6), (0.0, 1.5e4)
A, B, C = W.vortex_fields(nr, nz, LR, LZ)[:3]
d = O.Domain(LR, LZ, nr, nz, 0, 0); g = O.geometry(d, np.float64)
a, b, c = O.build_abc(A.astype(np.float64), B.astype(np.float64), C.astype(np.float64), d)
coe, _ = O.cal_coe(a, b, c, g["dr"], g["dz"], nr, nz)
ni, nj = nr - 2, nz - 2
idx = lambda i, j: j * ni + i
rows = []; cols = []; vals = []
offs = [(-1, 1), (0, 1), (1, 1), (-1, 0), (0, 0), (1, 0), (-1, -1), (0, -1), (1, -1)]
I, J = np.meshgrid(np.arange(ni), np.arange(nj))
for k, (di, dj) in enumerate(offs):
    ii, jj = I + di, J + dj
    m = (ii >= 0) & (ii < ni) & (jj >= 0) & (jj < nj)
    rows.append(idx(I[m], J[m])); cols.append(idx(ii[m], jj[m])); vals.append(coe[1:-1, 1:-1, k][m])
L = -sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(ni * nj, ni * nj))
Lc = L.tocoo()
gi = Lc.row % ni + 1; gic = Lc.col % ni + 1; jr = Lc.row // ni; jc = Lc.col // ni
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 32
k = (jr == jc) & ((gi // blk) == (gic // blk))
M = sp.csr_matrix((Lc.data[k], (Lc.row[k], Lc.col[k])), shape=L.shape).tocsc()
Mlu = spl.splu(M)
n = L.shape[0]
prec = lambda r: Mlu.solve(r)

# largest eigenvalue of M^-1 L by power iteration (what the device would do)
x = np.random.default_rng(0).standard_normal(n)
for it in range(60):
    y = prec(L @ x); lam = np.linalg.norm(y) / np.linalg.norm(x); x = y / np.linalg.norm(y)
print("lmax(M^-1 L) ~ %.4f (power iteration, 60 its)" % lam)
lmax = 1.02 * lam


def hat_matrix(npts, step):
    nodes = np.arange(step, npts + 1, step); nodes = nodes[nodes <= npts]
    r_ = []; c_ = []; v_ = []
    for cc, xc in enumerate(nodes):
        for gi_ in range(max(1, xc - step + 1), min(npts, xc + step - 1) + 1):
            w = 1.0 - abs(gi_ - xc) / step
            if w > 0: r_.append(gi_ - 1); c_.append(cc); v_.append(w)
    return sp.csr_matrix((v_, (r_, c_)), shape=(npts, len(nodes)))


dr, dz = LR[1] / (nr - 1), LZ[1] / (nz - 1)
lat = W.heating_lattice(64, 64, LR, LZ, 2 * dr, 2 * dz)
rhs = []
for loc in (2080, 5, 4000, 130):
    Q = MO.heat_field(lat[loc], g, np.float64); _, f = O.rhs_thermal(Q, d)
    rhs.append(-f[1:-1, 1:-1].ravel())


def cheb_smooth(x, fv, nu, lo, hi):
    th = (hi + lo) / 2; de = (hi - lo) / 2; sig = th / de
    r = fv - L @ x
    rho_k = 1 / sig; dvec = prec(r) / th
    for kk in range(1, nu + 1):
        x = x + dvec
        if kk == nu: break
        r = fv - L @ x
        rho_n = 1 / (2 * sig - rho_k)
        dvec = rho_n * rho_k * dvec + 2 * rho_n / de * prec(r); rho_k = rho_n
    return x


def run(P, Aci, fv, nu, eta, tol=1e-12, maxcyc=400):
    x = np.zeros(n); rms0 = np.sqrt((fv ** 2).mean())
    hist = []
    for cyc in range(maxcyc):
        r = fv - L @ x
        rel = np.sqrt((r ** 2).mean()) / rms0
        hist.append(rel)
        if rel < tol: return cyc, hist
        if P is not None:
            x = x + P @ (Aci @ (P.T @ r))
        x = cheb_smooth(x, fv, nu, lmax / eta, lmax)
    return maxcyc, hist


for (ax, az) in ((32, 16), (16, 16), (32, 8), (16, 8), (8, 8)):
    Px = hat_matrix(ni, ax); Pz = hat_matrix(nj, az)
    P = sp.kron(Pz, Px).tocsr()
    t0 = time.time()
    Ac = (P.T @ L @ P).toarray(); Aci = np.linalg.inv(Ac)
    print("bilinear spacing %dx%d: coarse %d x %d = %d unknowns (setup %.1fs)" % (ax, az, Px.shape[1], Pz.shape[1], P.shape[1], time.time() - t0))
    for nu in (8, 12, 16, 24):
        for eta in (10, 20, 40, 80):
            res = [run(P, Aci, fv, nu, eta) for fv in rhs[:2]]
            cyc = [r_[0] for r_ in res]
            fac = [(h[-1] / h[1]) ** (1.0 / max(len(h) - 2, 1)) for _, h in res]
            print("   nu=%2d eta=%3d: cycles %s  passes(nu+1 per cycle) %s  conv/cycle %s" % (nu, eta, cyc, [c_ * (nu + 1) for c_ in cyc], ["%.3f" % f_ for f_ in fac]), flush=True)
