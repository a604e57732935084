"""CPU model (fp64) of a LOCAL depth bound for the march's early cut-off: suffix / prefix maxima of the per-column (per-row) depth
along the axis the light is closer to, against the one-bound-per-face cut-off that is built.  `python tools/sim_march_horizon.py
[synthetic|bench]` (bench = /tmp/bench_depth.npy, the CNN's depth for the bench's noise images): kept in-mask samples 0.75 -> 0.55 on
synthetic faces, 0.39 -> 0.35 on the bench's maps - not built (5 more instructions per sample would eat the gain on the bench)."""
import numpy as np, sys, os
HERE=os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(HERE,'sim_march_cutoff.py')).read().split("for use_mask_range in")[0])
which=sys.argv[1] if len(sys.argv)>1 else 'synthetic'
if which=='bench':
    D=np.load('/tmp/bench_depth.npy')[0,0].astype(np.float64)
Zmax=max(D[md].max(),0.0); Zmin=min(D[md].min(),0.0)
Dm=np.where(md,D,-np.inf)
colmax=Dm.max(0); rowmax=Dm.max(1)
def dil(a):
    b=a.copy(); b[1:]=np.maximum(b[1:],a[:-1]); b[:-1]=np.maximum(b[:-1],a[1:]); return b
cold=dil(colmax); rowd=dil(rowmax)
tot_in=tot_g=tot_l=0
for li in range(18):
    L=np.array(LIGHTS_18[li]); L=L/np.linalg.norm(L)*4013.0
    ex,ey=ray_end(x,y,L[0],L[1])
    dx=(ex-x).astype(np.float64); dy=(ey-y).astype(np.float64)
    px=x[...,None]+t*dx[...,None]; py=y[...,None]+t*dy[...,None]
    ci=np.rint(px).astype(int)+W//2; ri=H//2-np.rint(py).astype(int)
    inside=m[ri.clip(0,H-1),ci.clip(0,W-1)]
    u=px+128-1e-4; v=128-py-1e-4
    uf=np.floor(u).astype(int); vf=np.floor(v).astype(int); uc=np.ceil(u).astype(int); vc=np.ceil(v).astype(int)
    zi=(D[vf,uf]*(uc-u)+D[vf,uc]*(u-uf))*(vc-v)+(D[vc,uf]*(uc-u)+D[vc,uc]*(u-uf))*(v-vf)
    z=D[...,None]
    bax=u-128-x[...,None]; bay=128-v-y[...,None]; baz=zi-z
    bcx=(L[0]-x)[...,None]; bcy=(L[1]-y)[...,None]; bcz=(L[2]-D)[...,None]
    c0=bay*bcz-baz*bcy; c1=baz*bcx-bax*bcz; c2=bax*bcy-bay*bcx
    q=np.where(inside,c0**2+c1**2+c2**2,np.inf)
    run=np.minimum.accumulate(q,axis=-1)
    A=(bcx**2+bcy**2)[...,0]; S1=dx*bcx[...,0]+dy*bcy[...,0]; bz=bcz[...,0]
    nxy=np.abs(bcx[...,0])+np.abs(bcy[...,0])
    E=2.0**-21*((np.abs(bz)+nxy)*512+2*(Zmax-Zmin+max(abs(Zmax),abs(Zmin)))*nxy)
    sig=2e-4*nxy
    thr=(np.sqrt(run)+4*E[...,None])/np.sqrt(A)[...,None]     # needed height gap, known after sample k
    thr_prev=np.concatenate([np.full(thr.shape[:-1]+(1,),np.inf),thr[...,:-1]],-1)
    hk=(bz/A)[...,None]*(t*S1[...,None]-sig[...,None])          # line height above the pixel (lower bound), ascending case
    ok=((bz>0)&(S1>0))[...,None]
    # global bound
    gdone=ok&(hk-(Zmax-z)>=thr_prev)
    # local 1-D table: major axis by light, suffix/prefix by sign
    if abs(L[0])>=abs(L[1]):
        T=np.maximum.accumulate(cold[::-1])[::-1] if L[0]>0 else np.maximum.accumulate(cold)
        idx=ci.clip(0,W-1); match=(np.sign(dx)==np.sign(L[0]))[...,None]
    else:
        # light above (+y) -> rays move to smaller rows
        T=np.maximum.accumulate(rowd) if L[1]>0 else np.maximum.accumulate(rowd[::-1])[::-1]
        idx=ri.clip(0,H-1); match=(np.sign(dy)==np.sign(L[1]))[...,None]
    ldone=ok&match&(hk-(T[idx]-z)>=thr_prev)
    gd=np.maximum.accumulate(gdone,axis=-1); ld=np.maximum.accumulate(ldone|gdone,axis=-1)
    keptg=inside&~gd; keptl=inside&~ld
    # exactness
    prevrun=np.concatenate([np.full(q.shape[:-1]+(1,),np.inf),run[...,:-1]],-1)
    assert not ((inside&ld)&(q<prevrun)).any()
    tot_in+=inside.sum(); tot_g+=keptg.sum(); tot_l+=keptl.sum()
    print(li,"in %.1f global %.1f local %.1f"%(inside.sum()/65536,keptg.sum()/65536,keptl.sum()/65536))
print("total kept: global %.3f local %.3f"%(tot_g/tot_in,tot_l/tot_in))
