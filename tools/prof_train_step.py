"""A few eager training iterations (TRAIN:617-656, B = 16) for an ncu launch list: 2 warm-up iterations, then j = 0 (discriminator
update) and j = 1.  Prints the launch counts so the list can be cut per iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geomconsistentfr_b200 import PatchGAN, RelightNet, intrinsic_matrix, ops
from geomconsistentfr_b200.synthetic import LIGHTS_18, synthetic_face
from geomconsistentfr_b200.trainer import TrainStep
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
B = 16
net = RelightNet(batch_size=B)
net.load_state_dict(torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu"), strict=True)
net = net.float().cuda().train()
if len(sys.argv) > 1:
    net.train_precision = int(sys.argv[1])
torch.manual_seed(0)
step = TrainStep(net, PatchGAN().cuda(), intrinsic_matrix().cuda())
g = torch.Generator().manual_seed(0)
img = torch.rand(B, 256, 256, 3, generator=g).cuda()
faces = [synthetic_face(seed=i) for i in range(B)]
mf = torch.stack([f[1] for f in faces]).float().cuda()
depth_gt = (torch.stack([f[0] for f in faces]) * 0.5).cuda()
albedo_gt = torch.rand(B, 256, 256, generator=g).cuda()
light_gt = torch.tensor([[0.5, *LIGHTS_18[i % 18]] for i in range(B)], dtype=torch.float32).cuda()
batch = (mf, mf, depth_gt, albedo_gt, light_gt)
marks = []
for j in (0, 1, 0, 1):
    n0 = ops.launch_count()
    step.step(img, 200, *batch, j=j)
    torch.cuda.synchronize()
    marks.append(ops.launch_count() - n0)
print("library launches per iteration (j = 0, 1, 0, 1):", marks)
