"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck): smoke() (march + one eval forward through the P16
tcgen05 convs) and one full training iteration (TRAIN:617-656, B = 2) in 3xTF32 and in bf16 (tensor-core wgrad)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("smoke", "all"):
    import __graft_entry__ as g
    g.smoke()
if what in ("train", "all"):
    from geomconsistentfr_b200 import PatchGAN, RelightNet, intrinsic_matrix
    from geomconsistentfr_b200.synthetic import LIGHTS_18, synthetic_face
    from geomconsistentfr_b200.trainer import TrainStep
    G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    B = 2
    for prec in (3, 4):
        net = RelightNet(batch_size=B)
        net.load_state_dict(torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu"), strict=True)
        net = net.float().cuda().train()
        net.train_precision = prec
        torch.manual_seed(0)
        step = TrainStep(net, PatchGAN().cuda(), intrinsic_matrix().cuda())
        g_ = torch.Generator().manual_seed(0)
        img = torch.rand(B, 256, 256, 3, generator=g_).cuda()
        faces = [synthetic_face(seed=i) for i in range(B)]
        mf = torch.stack([f[1] for f in faces]).float().cuda()
        depth_gt = (torch.stack([f[0] for f in faces]) * 0.5).cuda()
        albedo_gt = torch.rand(B, 256, 256, generator=g_).cuda()
        light_gt = torch.tensor([[0.5, *LIGHTS_18[i % 18]] for i in range(B)], dtype=torch.float32).cuda()
        total, _ = step.step(img, 200, mf, mf, depth_gt, albedo_gt, light_gt, j=0)
        torch.cuda.synchronize()
        print("train step precision", prec, "loss", float(total))
