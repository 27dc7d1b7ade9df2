import numpy as np, sys
sys.path.insert(0,'/root/repo')
from bmfr_b200 import synth
from oracle.oracle import Oracle
from tests import util
w,h,frames=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3])
pl,nl=synth.limits()
o=Oracle("port",w,h,position_limit_squared=pl,normal_limit_squared=nl,keep_tmp=1)
for fr in util.sequence(w,h,frames):
    o.frame(*fr)
tmp=o.buffer("tmp_data").astype(np.float32)   # [NB,13,32,32] as K1 left it
wts=o.buffer("weights")                        # [NB,10,3]
mm=o.buffer("mins_maxs")
noise=o.buffer("noise_tile")                   # [9,1024] float64
NB=tmp.shape[0]
A=tmp.reshape(NB,13,1024).transpose(0,2,1).copy()  # [NB,1024,13]
f32=np.float32
worst=dict(chol=0,chol_c=0,house=0)
errs=[]
def chol32(G):
    n=G.shape[0]; L=np.zeros_like(G)
    for j in range(n):
        s=G[j,j]-np.sum(L[j,:j]*L[j,:j],dtype=f32)
        L[j,j]=np.sqrt(max(s,f32(1e-30)))
        for i in range(j+1,n):
            L[i,j]=(G[i,j]-np.sum(L[i,:j]*L[j,:j],dtype=f32))/L[j,j]
    return L
def solve_gram(X,Y,shift):
    # X [1024,9] features (cols1..9), Y [1024,3]; fp32 Gram with column shifts, Cholesky, back-substitution, all fp32
    Xs=(X-shift[None,:]).astype(f32)
    M=np.concatenate([np.ones((1024,1),f32),Xs,Y.astype(f32)],axis=1)   # 13 columns
    # emulate per-thread 8-row partial sums then tree: use float32 matmul in chunks of 8 rows
    G=np.zeros((13,13),f32)
    parts=M.reshape(128,8,13)
    P=np.einsum('tri,trj->tij',parts,parts).astype(f32)     # per-thread partials (fp32)
    while P.shape[0]>1:
        P=(P[0::2]+P[1::2]).astype(f32)
    G=P[0]
    L=chol32(G[:10,:10].astype(f32))
    # solve L L^T w = G[:10,10:13]
    B=G[:10,10:13].astype(f32)
    z=np.zeros((10,3),f32)
    for i in range(10):
        z[i]=(B[i]-(L[i,:i,None]*z[:i]).sum(0,dtype=f32))/L[i,i]
    wv=np.zeros((10,3),f32)
    for i in range(9,-1,-1):
        wv[i]=(z[i]-(L[i+1:,i,None]*wv[i+1:]).sum(0,dtype=f32))/L[i,i]
    # undo shift: w0 -= sum w_j shift_j
    wv[0]=wv[0]-(wv[1:]*shift[:,None]).sum(0,dtype=f32)
    return wv
rng=np.random.default_rng(0)
sel=rng.choice(NB,size=min(NB,int(sys.argv[4])),replace=False)
for g in sel:
    a=A[g].astype(np.float64)
    mn=mm[g,:,0].astype(np.float64); mx=mm[g,:,1].astype(np.float64)
    clean=a.copy()
    for k in range(6):
        d=mx[k]-mn[k]
        clean[:,4+k]=(clean[:,4+k]-mn[k])/(d if abs(d)>1 else 1.0)
    noisy=clean.copy()
    noisy[:,1:10]+=noise.T
    X64=noisy[:,:10]; Y64=noisy[:,10:13]
    w64=np.linalg.lstsq(X64,Y64,rcond=None)[0]
    # shifts: scaled features 0.5; normals: the value at pixel 0
    shift=np.concatenate([noisy[0,1:4],np.full(6,0.5)]).astype(f32)
    wc=solve_gram(noisy[:,1:10].astype(f32),Y64,shift)
    w0=solve_gram(noisy[:,1:10].astype(f32),Y64,np.zeros(9,f32))
    fit64=clean[:,:10]@w64
    fitc=clean[:,:10].astype(f32)@wc; fit0=clean[:,:10].astype(f32)@w0; fith=clean[:,:10].astype(f32)@wts[g]
    def rel(f): return float(np.max(np.abs(f-fit64)/np.maximum(np.abs(fit64),1e-2)))
    errs.append((rel(fith),rel(fitc),rel(fit0),np.linalg.cond(X64)))
e=np.array(errs)
print("blocks",len(e),"cond median/max",np.median(e[:,3]),e[:,3].max())
for i,n in enumerate(("householder fp32 (oracle)","gram+chol shifted","gram+chol unshifted")):
    print(f"{n:28s} max rel err vs fp64 lstsq fit: median {np.median(e[:,i]):.2e}  p99 {np.quantile(e[:,i],0.99):.2e}  max {e[:,i].max():.2e}")

# which part dominates: forming G in fp32, or the fp32 Cholesky / solves?
def solve_from_G(G,shift,dt):
    G=G.astype(dt)
    L=np.linalg.cholesky(G[:10,:10].astype(np.float64)).astype(dt) if dt==np.float64 else chol32(G[:10,:10])
    B=G[:10,10:13]
    z=np.zeros((10,3),dt)
    for i in range(10):
        z[i]=(B[i]-(L[i,:i,None]*z[:i]).sum(0,dtype=dt))/L[i,i]
    wv=np.zeros((10,3),dt)
    for i in range(9,-1,-1):
        wv[i]=(z[i]-(L[i+1:,i,None]*wv[i+1:]).sum(0,dtype=dt))/L[i,i]
    wv[0]=wv[0]-(wv[1:]*shift[:,None].astype(dt)).sum(0,dtype=dt)
    return wv
errs2=[]
for g in sel[:150]:
    a=A[g].astype(np.float64)
    mn=mm[g,:,0].astype(np.float64); mx=mm[g,:,1].astype(np.float64)
    clean=a.copy()
    for k in range(6):
        d=mx[k]-mn[k]
        clean[:,4+k]=(clean[:,4+k]-mn[k])/(d if abs(d)>1 else 1.0)
    noisy=clean.copy(); noisy[:,1:10]+=noise.T
    noisy32=noisy.astype(f32)
    w64=np.linalg.lstsq(noisy32[:,:10].astype(np.float64),noisy32[:,10:13].astype(np.float64),rcond=None)[0]
    fit64=clean[:,:10]@w64
    shift=np.concatenate([noisy32[0,1:4],np.full(6,0.5,f32)]).astype(f32)
    M=np.concatenate([np.ones((1024,1),f32),(noisy32[:,1:10]-shift[None,:]).astype(f32),noisy32[:,10:13]],axis=1)
    Gex=(M.astype(np.float64).T@M.astype(np.float64))
    parts=M.reshape(128,8,13); P=np.einsum('tri,trj->tij',parts,parts).astype(f32)
    G32tree64=P.astype(np.float64).sum(0)
    while P.shape[0]>1: P=(P[0::2]+P[1::2]).astype(f32)
    G32=P[0]
    def rel(wv): 
        f=clean[:,:10]@wv.astype(np.float64); return float(np.max(np.abs(f-fit64)/np.maximum(np.abs(fit64),1e-2)))
    errs2.append((rel(solve_from_G(Gex,shift,f32)),rel(solve_from_G(G32,shift,np.float64)),rel(solve_from_G(G32tree64,shift,np.float64)),rel(solve_from_G(G32,shift,f32))))
e2=np.array(errs2)
for i,n in enumerate(("exact G + fp32 chol","fp32 G + fp64 chol","fp32 partials/fp64 tree + fp64 chol","fp32 G + fp32 chol")):
    print(f"{n:38s} median {np.median(e2[:,i]):.2e} p99 {np.quantile(e2[:,i],0.99):.2e} max {e2[:,i].max():.2e}")

errs3=[]
for g in sel[:150]:
    a=A[g].astype(np.float64)
    mn=mm[g,:,0].astype(np.float64); mx=mm[g,:,1].astype(np.float64)
    clean=a.copy()
    for k in range(6):
        d=mx[k]-mn[k]
        clean[:,4+k]=(clean[:,4+k]-mn[k])/(d if abs(d)>1 else 1.0)
    noisy=clean.copy(); noisy[:,1:10]+=noise.T
    noisy32=noisy.astype(f32)
    w64=np.linalg.lstsq(noisy32[:,:10].astype(np.float64),noisy32[:,10:13].astype(np.float64),rcond=None)[0]
    fit64=clean[:,:10]@w64
    def rel(wv):
        f=clean[:,:10]@wv.astype(np.float64); return float(np.max(np.abs(f-fit64)/np.maximum(np.abs(fit64),1e-2)))
    out=[]
    for mode in ("mean_all","mean_feat"):
        mean=noisy32[:,1:13].mean(0,dtype=f32)   # fp32 means of the 12 non-constant columns
        if mode=="mean_feat": mean[9:]=0
        Mc=(noisy32[:,1:13]-mean[None,:]).astype(f32)
        M=np.concatenate([np.ones((1024,1),f32),Mc],axis=1)
        parts=M.reshape(128,8,13); P=np.einsum('tri,trj->tij',parts,parts).astype(f32)
        while P.shape[0]>1: P=(P[0::2]+P[1::2]).astype(f32)
        G=P[0]
        L=chol32(G[:10,:10]); B=G[:10,10:13]
        z=np.zeros((10,3),f32)
        for i in range(10): z[i]=(B[i]-(L[i,:i,None]*z[:i]).sum(0,dtype=f32))/L[i,i]
        wv=np.zeros((10,3),f32)
        for i in range(9,-1,-1): wv[i]=(z[i]-(L[i+1:,i,None]*wv[i+1:]).sum(0,dtype=f32))/L[i,i]
        # un-centre: y - my = w0 + sum w_j (x_j - m_j)  ->  intercept
        wv[0]=wv[0]+mean[9:]-(wv[1:]*mean[:9,None]).sum(0,dtype=f32)
        out.append(rel(wv))
    errs3.append(out)
e3=np.array(errs3)
for i,n in enumerate(("true-mean centred (all 12), fp32 everything","true-mean centred (features only)")):
    print(f"{n:48s} median {np.median(e3[:,i]):.2e} p99 {np.quantile(e3[:,i],0.99):.2e} max {e3[:,i].max():.2e}")

errs4=[]
for g in sel[:150]:
    a=A[g].astype(np.float64)
    mn=mm[g,:,0].astype(np.float64); mx=mm[g,:,1].astype(np.float64)
    clean=a.copy()
    for k in range(6):
        d=mx[k]-mn[k]
        clean[:,4+k]=(clean[:,4+k]-mn[k])/(d if abs(d)>1 else 1.0)
    noisy=clean.copy(); noisy[:,1:10]+=noise.T
    noisy32=noisy.astype(f32)
    w64=np.linalg.lstsq(noisy32[:,:10].astype(np.float64),noisy32[:,10:13].astype(np.float64),rcond=None)[0]
    fit64=clean[:,:10]@w64
    def rel(wv):
        f=clean[:,:10]@wv.astype(np.float64); return float(np.max(np.abs(f-fit64)/np.maximum(np.abs(fit64),1e-2)))
    mean=noisy32[:,1:13].mean(0,dtype=f32)
    Mc=(noisy32[:,1:13]-mean[None,:]).astype(f32)
    M=np.concatenate([np.ones((1024,1),f32),Mc],axis=1)
    parts=M.reshape(128,8,13); P=np.einsum('tri,trj->tij',parts,parts).astype(f32)
    # warp-level fp32 tree (32 lanes), then 4 warps combined in fp64
    Pw=P.reshape(4,32,13,13)
    while Pw.shape[1]>1: Pw=(Pw[:,0::2]+Pw[:,1::2]).astype(f32)
    Gd=Pw[:,0].astype(np.float64).sum(0)
    G32=Pw[:,0].sum(0,dtype=f32)
    out=[]
    for G,dt in ((G32,np.float64),(Gd,np.float64)):
        G=G.astype(dt)
        L=np.linalg.cholesky(G[:10,:10]); B=G[:10,10:13]
        z=np.linalg.solve(L,B); wv=np.linalg.solve(L.T,z)
        wv[0]=wv[0]+mean[9:]-(wv[1:]*mean[:9,None]).sum(0)
        out.append(rel(wv))
    errs4.append(out)
e4=np.array(errs4)
for i,n in enumerate(("centred, fp32 G, fp64 chol","centred, fp32 warp trees + fp64 across warps + fp64 chol")):
    print(f"{n:58s} median {np.median(e4[:,i]):.2e} p99 {np.quantile(e4[:,i],0.99):.2e} max {e4[:,i].max():.2e}")
