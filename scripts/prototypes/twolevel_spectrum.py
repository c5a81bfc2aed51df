"""Prototype (CPU, scipy): additive two-level preconditioner = 32-point radial block solves + Galerkin coarse solve on the
bench operator; extreme eigenvalues of the preconditioned operator and Chebyshev sweep counts.  python twolevel_spectrum.py 256 128"""
import numpy as np, sys, scipy.sparse as sp, scipy.sparse.linalg as spl, time
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
I,J=np.meshgrid(np.arange(ni),np.arange(nj))
for k,(di,dj) in enumerate(offs):
    ii,jj=I+di,J+dj
    m=(ii>=0)&(ii<ni)&(jj>=0)&(jj<nj)
    rows.append(idx(I[m],J[m])); cols.append(idx(ii[m],jj[m])); vals.append(coe[1:-1,1:-1,k][m])
L=sp.csr_matrix((np.concatenate(vals),(np.concatenate(rows),np.concatenate(cols))),shape=(ni*nj,ni*nj))
L=-L   # positive definite-ish
Lc=L.tocoo()
gi=Lc.row%ni+1; gic=Lc.col%ni+1; jr=Lc.row//ni; jc=Lc.col//ni   # global i = interior+1
blk=32
k=(jr==jc)&((gi//blk)==(gic//blk))
M=sp.csr_matrix((Lc.data[k],(Lc.row[k],Lc.col[k])),shape=L.shape).tocsc()
Mlu=spl.splu(M)
n=L.shape[0]
def extremes(apply_prec):
    op=spl.LinearOperator((n,n),matvec=lambda x: apply_prec(L@x))
    lmax=spl.eigs(op,k=1,which='LM',return_eigenvectors=False,tol=1e-6,maxiter=5000)[0].real
    # smallest: shift
    op2=spl.LinearOperator((n,n),matvec=lambda x: lmax*x-apply_prec(L@x))
    l2=spl.eigs(op2,k=1,which='LM',return_eigenvectors=False,tol=1e-8,maxiter=20000)[0].real
    return lmax-l2, lmax
lmin,lmax=extremes(lambda r: Mlu.solve(r))
print("block-line only: lambda in [%.3e, %.3f], kappa=%.0f, sqrt=%.1f"%(lmin,lmax,lmax/lmin,np.sqrt(lmax/lmin)))
for (ax,az) in ((32,8),(16,8),(32,4),(16,4),(8,4),(64,16)):
    # piecewise-constant aggregates over global indices
    agx=(I+1)//ax; agz=(J+1)//az
    ncx=agx.max()+1; ncz=agz.max()+1
    agg=(agz*ncx+agx).ravel()
    P=sp.csr_matrix((np.ones(n),(np.arange(n),agg)),shape=(n,ncx*ncz))
    Ac=(P.T@L@P).toarray()
    Aci=np.linalg.inv(Ac)
    prec=lambda r: Mlu.solve(r)+P@(Aci@(P.T@r))
    lmin,lmax=extremes(prec)
    print("agg %dx%d (coarse %d): lambda in [%.3e, %.3f], kappa=%.0f, sqrt=%.1f"%(ax,az,ncx*ncz,lmin,lmax,lmax/lmin,np.sqrt(lmax/lmin)))
print("--- bilinear coarse spaces")
def hat_matrix(npts, step):
    # 1-D linear interpolation from coarse nodes at global index 0, step, 2*step, ... (Dirichlet ends dropped) to interior points 1..npts
    nodes=np.arange(step, npts+1, step)
    nodes=nodes[nodes<=npts]
    rows=[];cols=[];vals=[]
    for c,xc in enumerate(nodes):
        for gi in range(max(1,xc-step+1), min(npts, xc+step-1)+1):
            w=1.0-abs(gi-xc)/step
            if w>0: rows.append(gi-1); cols.append(c); vals.append(w)
    return sp.csr_matrix((vals,(rows,cols)),shape=(npts,len(nodes)))
for (ax,az) in ((32,8),(16,8),(16,4),(8,4),(8,8),(4,4)):
    Px=hat_matrix(ni,ax); Pz=hat_matrix(nj,az)
    P=sp.kron(Pz,Px).tocsr()
    Ac=(P.T@L@P).toarray()
    Aci=np.linalg.inv(Ac)
    prec=lambda r: Mlu.solve(r)+P@(Aci@(P.T@r))
    lmin,lmax=extremes(prec)
    print("bilinear %dx%d (coarse %d): lambda in [%.3e, %.3f], kappa=%.0f, sqrt=%.1f"%(ax,az,P.shape[1],lmin,lmax,lmax/lmin,np.sqrt(lmax/lmin)))
print("--- Chebyshev sweeps to 1e-12 (rms residual / rms f), heating RHS")
from tests import map_oracle as MO
dr,dz=LR[1]/(nr-1),LZ[1]/(nz-1)
lat=W.heating_lattice(64,64,LR,LZ,2*dr,2*dz)
Q=MO.heat_field(lat[2080],g,np.float64); _,f=O.rhs_thermal(Q,d)
fv=-f[1:-1,1:-1].ravel()          # L here is -L_ref
def cheb_solve(prec,lmin,lmax,tol=1e-12,maxit=20000):
    x=np.zeros(n); r=fv-L@x; rms0=np.sqrt((fv**2).mean())
    th=(lmax+lmin)/2; de=(lmax-lmin)/2; sig=th/de
    rho_k=1/sig; z=prec(r); dvec=z/th
    for k in range(1,maxit+1):
        x=x+dvec; r=fv-L@x
        if np.sqrt((r**2).mean())<tol*rms0: return k
        z=prec(r); rho_n=1/(2*sig-rho_k)
        dvec=rho_n*rho_k*dvec+2*rho_n/de*z; rho_k=rho_n
    return maxit
lmin,lmax=extremes(lambda r: Mlu.solve(r))
print("block-line only:", cheb_solve(lambda r: Mlu.solve(r),lmin*0.95,lmax*1.01))
for (ax,az) in ((32,8),(8,8),(4,4)):
    Px=hat_matrix(ni,ax); Pz=hat_matrix(nj,az); P=sp.kron(Pz,Px).tocsr()
    Aci=np.linalg.inv((P.T@L@P).toarray())
    prec=lambda r: Mlu.solve(r)+P@(Aci@(P.T@r))
    lmin,lmax=extremes(prec)
    print("two-level bilinear %dx%d:"%(ax,az), cheb_solve(prec,lmin*0.95,lmax*1.01))
