"""CPU simulation of the multi-GPU item-replica exchange with the reference loop (oracle) as the epoch kernel: `world` user shards
train one BPR epoch each from the same item factors, then the replicas are combined with rule "sum" (V_start + sum of the changes) or
"mean" (V_start + the changes averaged over the shards that changed the row; parallel.py's default).  Prints pairwise accuracy and
the largest |V|, |U|, |B| per epoch: the sum diverges from 4 shards on, the mean tracks the single-process run.
    python tools/sim_localsgd.py"""
import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests')); from conftest import synth_csr
def run(world, mode, epochs=11, k=16, lr=0.05, reg=0.01, n_users=8000, n_items=1000, nnz=400000):
    indptr, indices = synth_csr(n_users, n_items, nnz, seed=1, zipf=1.0)
    rng = np.random.RandomState(0)
    U = ((rng.rand(n_users,k)-0.5)/k).astype(np.float32); V=((rng.rand(n_items,k)-0.5)/k).astype(np.float32); B=np.zeros(n_items,np.float32)
    # shards: contiguous users with equal nnz
    nnz=len(indices); cuts=[np.searchsorted(indptr, (r*nnz)//world) for r in range(world)]+[n_users]
    for e in range(epochs):
        Vs, Bs = V.copy(), B.copy()
        dV=np.zeros_like(V); dB=np.zeros_like(B); cV=np.zeros(n_items); 
        for r in range(world):
            lo,hi=cuts[r],cuts[r+1]
            ip=(indptr[lo:hi+1]-indptr[lo]).astype(np.int32); ix=indices[indptr[lo]:indptr[hi]]
            n=len(ix)
            g=np.random.RandomState(100*e+r)
            ii=g.randint(n,size=n).astype(np.int64); jj=g.randint(n_items,size=n).astype(np.int32)
            Vr,Br=Vs.copy(),Bs.copy(); Ur=U[lo:hi].copy()
            O.bpr_replay(ii,jj,ip,ix,Ur,Vr,Br,lr,reg,True)
            U[lo:hi]=Ur
            d=Vr-Vs; dV+=d; dB+=Br-Bs; cV+=(np.abs(d).sum(1)>0)
        if mode=="sum": V=Vs+dV; B=Bs+dB
        else:
            c=np.maximum(cV,1)[:,None].astype(np.float32); V=(Vs+dV/c).astype(np.float32); B=(Bs+dB/np.maximum(cV,1).astype(np.float32)).astype(np.float32)
        rows=np.repeat(np.arange(n_users),np.diff(indptr)); g2=np.random.RandomState(5); pk=g2.randint(len(indices),size=20000); uu,pi,pj=rows[pk],indices[pk],g2.randint(n_items,size=20000); acc=np.mean(np.einsum("nk,nk->n",U[uu],V[pi]-V[pj])+B[pi]-B[pj]>0)
        print(world,mode,e,"acc %.3f"%acc,"max|V| %.3g max|U| %.3g max|B| %.3g nan=%s"%(np.abs(V).max(),np.abs(U).max(),np.abs(B).max(),np.isnan(V).any()))
        if not np.isfinite(V).all(): break
for w,m in ((1,"sum"),(8,"mean"),(4,"mean"),(2,"mean"),(4,"sum")):
    run(w,m)
