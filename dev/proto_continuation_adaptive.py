"""Development prototype: continuation with a per-mode stopping test |d mu| <= 1e-11 |mu| (at most 8\niterations), anchors every 45th frequency of the cfg1 sweep, each continued to its 8 neighbours.\nPrints the mean / max iteration count per mode and the parity count.  Not imported by the product."""
import sys, warnings, time
warnings.filterwarnings('ignore')
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'axisym-safe-python_b200'))
import numpy as np, scipy.linalg as la
from oracle import safe_oracle as so
from axisafe_b200 import cases
import multiprocessing as mp
nf=2000
def adaptive(S,k0,X0,tol=1e-11,maxit=8):
    n=S.shape[0]; k=k0.copy(); its=np.zeros(n,int)
    for j in range(n):
        mu=k[j]; x=X0[:,j]/np.linalg.norm(X0[:,j])
        for it in range(maxit):
            y=la.solve(S-mu*np.eye(n),x)
            d=np.vdot(x,y); dm=1.0/d if d!=0 else 0
            mu=mu+dm; x=y/np.linalg.norm(y); its[j]=it+1
            if abs(dm)<=tol*abs(mu): break
        k[j]=mu
    return k,its
def work(a):
    c=cases.pipe_soil_pml(nf=nf)
    G=so.assemble(so.build_mesh(c['sets'],c['fmax'],c['PML'],c['PML_props']),c['n']); op=so.operators(G,c['n'])
    w=2*np.pi*c['f']
    k,X=la.eig(so.S_matrix(op,w[a])); out=[]
    for j in (-4,-3,-2,-1,1,2,3,4):
        i=a+j
        if i<0 or i>=nf: continue
        S=so.S_matrix(op,w[i]); ref=la.eigvals(S)
        k2,its=adaptive(S,k,X)
        perm,err=so.match_eigenvalues(ref,k2)
        ok=int((abs(ref-k2[perm])<=so.eig_tolerance(S,ref)).sum())
        out.append((i,abs(j),its.mean(),its.max(),ok))
    return out
if __name__=='__main__':
    anchors=list(range(4,nf,45))
    with mp.Pool(8) as p: res=p.map(work,anchors)
    flat=[r for rr in res for r in rr]
    for jj in (1,2,3,4):
        sel=[r for r in flat if r[1]==jj]
        print('|j|=%d: %d frequencies, mean iterations per mode %.2f, max %d, all 126 within the bar: %d of %d'%(jj,len(sel),np.mean([r[2] for r in sel]),max(r[3] for r in sel),sum(r[4]==126 for r in sel),len(sel)))
    print('overall mean iterations %.2f'%np.mean([r[2] for r in flat]))
