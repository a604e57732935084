"""CPU estimate of the ray march instruction volume per warp for different warp shapes on the bench masks: warp-iterations walked
(sample-range union over the 32 rays) and warp-iterations with at least one in-mask lane, x 17 / 67 instructions (DESIGN K1)."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomconsistentfr_b200.synthetic import synthetic_face, LIGHTS_18
H=W=256
t=np.arange(0.025,0.825,0.005)
_,m=synthetic_face(seed=0); m=m.numpy().astype(bool)
rows,cols=np.nonzero(m); c_lo,c_hi,r_lo,r_hi=cols.min(),cols.max(),rows.min(),rows.max()
col=np.arange(W)[None,:].repeat(H,0); row=np.arange(H)[:,None].repeat(W,1)
x=(col-W/2).astype(np.float32); y=(H/2-row).astype(np.float32)
xmin,xmax,ymin,ymax=-W/2,W/2-1,1-H/2,H/2
def ray_end(x,y,Lx,Ly):
    mm=(Ly-y)/((Lx-x)+1e-4); b=Ly-mm*Lx
    sx=-1 if Lx<xmin else (0 if Lx<=xmax else 1); sy=-1 if Ly<ymin else (0 if Ly<=ymax else 1)
    xe=xmin if sx<0 else xmax; ye=ymin if sy<0 else ymax
    exy=mm*xe+b; eyx=(ye-b)/(mm+1e-4)
    if sx!=0 and sy!=0:
        hit=(eyx>=xmin)&(eyx<=xmax); ex=np.where(hit,eyx,xe); ey=np.where(hit,ye,exy)
    elif sx!=0: ex=np.full_like(x,xe); ey=exy
    elif sy!=0: ex=eyx; ey=np.full_like(x,ye)
    else: ex=np.full_like(x,Lx); ey=np.full_like(x,Ly)
    return np.clip(ex,xmin,xmax),np.clip(ey,ymin,ymax)
def run(shape):
    wh,ww=shape
    tot_out=tot_in=0; n_iter_lane_in=0; n_pair=0
    for li in range(8):
        L=np.array(LIGHTS_18[li]); L=L/np.linalg.norm(L)*4013.0
        ex,ey=ray_end(x,y,L[0],L[1])
        dx=(ex-x).astype(np.float64); dy=(ey-y).astype(np.float64)
        px=x[...,None]+t*dx[...,None]; py=y[...,None]+t*dy[...,None]
        ci=np.rint(px).astype(int)+W//2; ri=H//2-np.rint(py).astype(int)
        inside=m[ri.clip(0,H-1),ci.clip(0,W-1)]
        # bbox interval per ray: sample k could be in bbox
        inb=(ci>=c_lo)&(ci<=c_hi)&(ri>=r_lo)&(ri<=r_hi)
        # per ray interval [first,last] in bbox
        kk=np.arange(len(t))
        first=np.where(inb.any(-1), inb.argmax(-1), len(t)); last=np.where(inb.any(-1), len(t)-1-inb[...,::-1].argmax(-1), -1)
        # warps
        f=first.reshape(H//wh,wh,W//ww,ww).transpose(0,2,1,3).reshape(-1,wh*ww)
        l=last.reshape(H//wh,wh,W//ww,ww).transpose(0,2,1,3).reshape(-1,wh*ww)
        ins=inside.reshape(H//wh,wh,W//ww,ww,len(t)).transpose(0,2,1,3,4).reshape(-1,wh*ww,len(t))
        kb=f.min(1); ke=l.max(1)
        walked=(kk[None,:]>=kb[:,None])&(kk[None,:]<=ke[:,None])
        anyin=ins.any(1)&walked
        tot_out+=walked.sum(); tot_in+=anyin.sum(); n_iter_lane_in+=ins.sum()
    nw=8*H*W//32
    print(shape,"walked/warp %.1f  anyin/warp %.1f  lane-in avg/ray %.1f  cost/warp %.0f"%(tot_out/nw,tot_in/nw,n_iter_lane_in/(8*H*W),(tot_out*17+tot_in*67)/nw))
for s in [(1,32),(2,16),(4,8),(8,4)]: run(s)
