"""Prototype (CPU, scipy): spectral gap of point Jacobi vs radial segment/line Jacobi on the bench operator.
python line_jacobi_spectrum.py 256 128"""
import numpy as np, sys, scipy.sparse as sp, scipy.sparse.linalg as spl
sys.path.insert(0,'/root/repo')
from oracle import oracle as O
from xlab_ee_fortran_b200 import workloads as W
nr,nz=int(sys.argv[1]),int(sys.argv[2])
LR,LZ=(0.0,1.0e6),(0.0,1.5e4)
A,B,C=W.vortex_fields(nr,nz,LR,LZ)[:3]
d=O.Domain(LR,LZ,nr,nz,0,0); g=O.geometry(d,np.float64)
a,b,c=O.build_abc(A.astype(np.float64),B.astype(np.float64),C.astype(np.float64),d)
coe,_=O.cal_coe(a,b,c,g["dr"],g["dz"],nr,nz)
ni,nj=nr-2,nz-2
idx=lambda i,j: j*ni+i
rows=[];cols=[];vals=[]
offs=[(-1,1),(0,1),(1,1),(-1,0),(0,0),(1,0),(-1,-1),(0,-1),(1,-1)]
for k,(di,dj) in enumerate(offs):
    I,J=np.meshgrid(np.arange(ni),np.arange(nj))
    ii,jj=I+di,J+dj
    m=(ii>=0)&(ii<ni)&(jj>=0)&(jj<nj)
    rows.append(idx(I[m],J[m])); cols.append(idx(ii[m],jj[m])); vals.append(coe[1:-1,1:-1,k][m])
L=sp.csr_matrix((np.concatenate(vals),(np.concatenate(rows),np.concatenate(cols))),shape=(ni*nj,ni*nj))
def rho_of(M):  # M = splitting matrix (sparse); iteration G = I - M^-1 L
    lu=spl.splu(M.tocsc())
    n=L.shape[0]
    op=spl.LinearOperator((n,n),matvec=lambda x: x-lu.solve(L@x))
    ev=spl.eigs(op,k=2,which='LM',return_eigenvectors=False,tol=1e-7,maxiter=20000)
    return np.abs(ev).max()
Dg=sp.diags(L.diagonal())
r=rho_of(Dg); print("point Jacobi: 1-rho = %.3e"%(1-r))
Lc=L.tocoo()
def seg_mask(m):
    i_r=Lc.row%ni; i_c=Lc.col%ni; j_r=Lc.row//ni; j_c=Lc.col//ni
    return (j_r==j_c)&((i_r//m)==(i_c//m))
for m in (8,16,32,64,ni):
    k=seg_mask(m)
    M=sp.csr_matrix((Lc.data[k],(Lc.row[k],Lc.col[k])),shape=L.shape)
    r=rho_of(M); print("x-segment line Jacobi m=%d: 1-rho = %.3e"%(m,1-r))
