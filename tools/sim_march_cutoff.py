"""CPU model of the march's exact early cut-off on the synthetic bench faces (fp64): fraction of in-mask samples / warp-iterations
that survive, and a check that no skipped sample ever beats the running minimum."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)) if '__file__' in dir() else HERE, 'sim_march_warp_shape.py')).read().split("def run(shape):")[0])
from geomconsistentfr_b200.synthetic import synthetic_face
depth,_=synthetic_face(seed=0,noise=2.0); D=depth.numpy().astype(np.float64)
# dilated-mask depth range
from scipy.ndimage import binary_dilation
md=binary_dilation(m,structure=np.ones((3,3),bool)); md[:,-1]=True; md[-1,:]=True
for use_mask_range in (True,False):
    Zmax=max(D[md].max() if use_mask_range else D.max(),0.0); Zmin=min(D[md].min() if use_mask_range else D.min(),0.0)
    tot_in=tot_kept=0; wi_old=wi_new=0; walk_old=walk_new=0
    for li in range(8):
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
        q=c0**2+c1**2+c2**2
        q=np.where(inside,q,np.inf)
        run=np.minimum.accumulate(q,axis=-1)          # qmin after sample k
        A=(bcx**2+bcy**2)[...,0]; S1=dx*bcx[...,0]+dy*bcy[...,0]
        bz=bcz[...,0]
        E=2.0**-21*(np.abs(bz)*512+2*(Zmax-Zmin+max(abs(Zmax),abs(Zmin)))*(np.abs(bcx[...,0])+np.abs(bcy[...,0])))
        sig=2e-4*(np.abs(bcx[...,0])+np.abs(bcy[...,0]))
        # cutoff index after each k: kc(k)
        h=(np.sqrt(run)+4*E[...,None])/np.sqrt(A)[...,None]
        bazmax=(Zmax-D)[...,None]*(1+1e-6)+1e-3
        with np.errstate(divide='ignore',invalid='ignore'):
            tcut=((bazmax+h)*A[...,None]/bz[...,None]+sig[...,None])/S1[...,None]
        ok=(bz>0)[...,None]&(S1>0)[...,None]&np.isfinite(tcut)
        kc=np.where(ok,np.ceil((tcut*(1+1e-5)-0.025)/0.005)+1,10**6)
        kcrun=np.minimum.accumulate(kc,axis=-1)       # lane cutoff known after processing sample k
        kk=np.arange(len(t))
        # a sample k is skipped for the lane if k >= cutoff known from samples < k
        prev=np.concatenate([np.full(kc.shape[:-1]+(1,),10**6),kcrun[...,:-1]],-1)
        kept=inside&(kk<prev)
        # exactness check in sim: skipped samples never beat the running min
        viol=(inside&~kept)&(q<np.concatenate([np.full(q.shape[:-1]+(1,),np.inf),run[...,:-1]],-1))
        assert not viol.any(), viol.sum()
        tot_in+=inside.sum(); tot_kept+=kept.sum()
        # warp level (4x8): iterations with any lane kept vs any lane inside ; walked range: until all lanes past cutoff
        def warp(a): return a.reshape(H//4,4,W//8,8,len(t)).transpose(0,2,1,3,4).reshape(-1,32,len(t))
        wi_old+=warp(inside).any(1).sum(); 
        lane_alive=warp(kk<prev)  # lane still wants samples at k
        alive=lane_alive.any(1)
        wi_new+=(warp(inside).any(1)&alive).sum()
    print("mask-range" if use_mask_range else "global-range","lane samples kept %.1f%%"%(100*tot_kept/tot_in),"warp in-mask iterations %d -> %d (%.1f%%)"%(wi_old,wi_new,100*wi_new/wi_old))
