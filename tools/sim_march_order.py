"""CPU model of the ray march's CTA load balance: per-CTA cost (slowest warp: 22 instructions per walked + 91 per in-mask
warp-iteration) list-scheduled in launch order onto 148 SMs x 6 resident CTAs, makespan / ideal for several CTA orders."""
import numpy as np, sys, heapq, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sim_march_warp_shape.py')).read().split("def run(shape):")[0])
def percost(shape):
    wh,ww=shape
    out=[]
    for li in range(8):
        L=np.array(LIGHTS_18[li]); L=L/np.linalg.norm(L)*4013.0
        ex,ey=ray_end(x,y,L[0],L[1])
        dx=(ex-x).astype(np.float64); dy=(ey-y).astype(np.float64)
        px=x[...,None]+t*dx[...,None]; py=y[...,None]+t*dy[...,None]
        ci=np.rint(px).astype(int)+W//2; ri=H//2-np.rint(py).astype(int)
        inside=m[ri.clip(0,H-1),ci.clip(0,W-1)]
        inb=(ci>=c_lo)&(ci<=c_hi)&(ri>=r_lo)&(ri<=r_hi)
        kk=np.arange(len(t))
        first=np.where(inb.any(-1), inb.argmax(-1), len(t)); last=np.where(inb.any(-1), len(t)-1-inb[...,::-1].argmax(-1), -1)
        f=first.reshape(H//wh,wh,W//ww,ww).transpose(0,2,1,3).reshape(H//wh,W//ww,wh*ww)
        l=last.reshape(H//wh,wh,W//ww,ww).transpose(0,2,1,3).reshape(H//wh,W//ww,wh*ww)
        ins=inside.reshape(H//wh,wh,W//ww,ww,len(t)).transpose(0,2,1,3,4).reshape(H//wh,W//ww,wh*ww,len(t))
        kb=f.min(-1); ke=l.max(-1)
        walked=(kk>=kb[...,None])&(kk<=ke[...,None])
        anyin=ins.any(2)&walked
        out.append(walked.sum(-1)*22+anyin.sum(-1)*91+400)   # [H/wh, W/ww] cost per warp
    return np.stack(out)
c=percost((4,8))       # [8, 64, 32] warps; CTA = 4 adjacent warps along x -> [8,64,8]
cta=c.reshape(8,64,8,4)
cta_time=cta.max(-1)   # latency model: CTA done when slowest warp done (warps run concurrently)
cta_work=cta.sum(-1)
print("warp cost mean %.0f max %d; cta max/mean %.2f"%(c.mean(),c.max(),cta_time.max()/cta_time.mean()))
# schedule: order z (b) slowest, then y, then x fastest; 148 SMs x 6 slots; each slot executes CTA with duration ~ cta_time (assume throughput share equal)
order=cta_time.reshape(-1)  # b, y, x order -> blockIdx.x fastest: matches (b,y,x) flatten
slots=[0.0]*(148*6); heapq.heapify(slots)
for d in order:
    s=heapq.heappop(slots); heapq.heappush(slots,s+d)
mk=max(slots); ideal=order.sum()/(148*6)
print("makespan %.0f ideal %.0f ratio %.3f"%(mk,ideal,mk/ideal))
def makespan(order):
    slots=[0.0]*(148*6); heapq.heapify(slots)
    for d in order:
        s=heapq.heappop(slots); heapq.heappush(slots,s+d)
    return max(slots)
T=cta_time  # [b,y,x]
print("LPT", makespan(sorted(T.reshape(-1),reverse=True))/ideal)
# center-out rows, batch fastest
ys=sorted(range(64), key=lambda yy: abs(yy-31.5))
o=[T[b,yy,xx] for yy in ys for xx in range(8) for b in range(8)]
print("rows center-out, x, batch fastest", makespan(o)/ideal)
xs=sorted(range(8), key=lambda xx: abs(xx-3.5))
o=[T[b,yy,xx] for yy in ys for xx in xs for b in range(8)]
print("rows+cols center-out", makespan(o)/ideal)
# by distance of tile center from image center
tiles=sorted([(yy,xx) for yy in range(64) for xx in range(8)], key=lambda p: ((p[0]*4+2-128)/100.0)**2+((p[1]*32+16-128)/80.0)**2)
o=[T[b,yy,xx] for (yy,xx) in tiles for b in range(8)]
print("elliptic distance", makespan(o)/ideal)
# 16-thread... smaller CTA: 2 warps (64 thr)
def light_order(major_by_light=True):
    o=[]
    per_b=[]
    for b in range(8):
        L=np.array(LIGHTS_18[b])
        tl=[(yy,xx) for yy in range(64) for xx in range(8)]
        # tile center coords: x = xx*32+16-128, y = 128-(yy*4+2)
        key=lambda p: ((p[1]*32+16-128)*L[0]+(128-(p[0]*4+2))*L[1])   # projection along light dir: small = far from light
        per_b.append(sorted(tl,key=key))
    for i in range(512):
        for b in range(8):
            yy,xx=per_b[b][i]; o.append(T[b,yy,xx])
    return o
print("sorted by projection on light dir (far side first), batch fastest", makespan(light_order())/ideal)
def simple_flip():
    o=[]
    for i in range(512):
        for b in range(8):
            L=np.array(LIGHTS_18[b])
            if abs(L[0])>=abs(L[1]):   # x major: iterate x slowest
                xx=i//64; yy=i%64
                if L[0]>0: pass
                else: xx=7-xx
                # light from +x: far side = small x first
            else:
                yy=i//8; xx=i%8
                if L[1]>0: yy=63-yy   # light from top (+y = small row): far side = bottom rows first
            o.append(T[b,yy,xx])
    return o
print("axis flip", makespan(simple_flip())/ideal)
for b in range(8):
    print(b, LIGHTS_18[b], (T[b].sum(0)/1000).astype(int))
def flip_face_slowest():
    o=[]
    for b in range(8):
        L=np.array(LIGHTS_18[b])
        for i in range(512):
            if abs(L[0])>=abs(L[1]):
                xx=i//64; yy=i%64
                if L[0]<=0: xx=7-xx
            else:
                yy=i//8; xx=i%8
                if L[1]>0: yy=63-yy
            o.append(T[b,yy,xx])
    return o
print("axis flip, batch slowest", makespan(flip_face_slowest())/ideal)
