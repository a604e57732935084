"""Dev tool: CUDA-event timing of the full eval forward at B (not the bench)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomconsistentfr_b200 import RelightNet, intrinsic_matrix
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
PREC = int(sys.argv[2]) if len(sys.argv) > 2 else 2
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
f = np.load(os.path.join(G, "ffhq.npz"))
net = RelightNet(); net.load_state_dict(torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu")); net = net.cuda().eval(); net.tc_precision = PREC
idx = [i % 10 for i in range(B)]
x = torch.from_numpy(f["q"][idx] / 1020.0).float().cuda()
m = torch.from_numpy(f["masks"][0].astype(np.float32).reshape(256, 256, 1)).cuda()
tl = torch.from_numpy(f["lights"][idx]).view(B, 3, 1, 1).cuda()
K = intrinsic_matrix().cuda(); amb = torch.full((B, 1, 1), 0.5).cuda()
def step(): return net(x, 200, K, m, tl, amb, None)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("eager forward B=%d: %.3f ms -> %.1f faces/s" % (B, ms, B / ms * 1e3))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step(); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        out = step()
torch.cuda.synchronize()
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("graph forward B=%d: %.3f ms -> %.1f faces/s" % (B, ms, B / ms * 1e3))
