"""TEST INFRASTRUCTURE — generates tests/golden/lpips.npz by running the reference's VENDORED, unmodified lpips module
(/root/reference/PerceptualSimilarity/lpips, `lpips.LPIPS(net='alex', spatial=True, pnet_rand=True)`: lpips.py:21-144) in the
authoring container.  The AlexNet trunk's ImageNet weights are not available offline, so the trunk is filled by `trunk_init`
(numpy RandomState: the test rebuilds the same weights); the linear heads are the SHIPPED lpips/weights/v0.1/alex.pth, stored in
the fixture (1 152 floats).  Stored: the spatial map for two seeded image pairs, the masked metric of test_network.py:41-45, and
the gradient of that metric w.r.t. the predicted image (every 2nd pixel).

    python -m oracle.make_golden_lpips
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def trunk_init(module, seed=3):
    """He-style deterministic fill of every conv of the AlexNet trunk (numpy RandomState, independent of torch's RNG)."""
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if ".slice" not in "." + name:
                continue
            if p.dim() == 4:
                fan_in = p.shape[1] * p.shape[2] * p.shape[3]
                v = rs.standard_normal(p.shape) * np.sqrt(2.0 / fan_in)
            else:
                v = 0.05 * rs.standard_normal(p.shape)
            p.copy_(torch.from_numpy(v.astype(np.float32)))
    return module


def case():
    rs = np.random.RandomState(8)
    ref = rs.uniform(-1, 1, (2, 3, 256, 256)).astype(np.float32)
    pred = np.clip(ref + 0.3 * rs.standard_normal(ref.shape), -1, 1).astype(np.float32)
    yy, xx = np.mgrid[0:256, 0:256]
    mask = ((((xx - 128) / 80.0) ** 2 + ((yy - 128) / 100.0) ** 2) < 1.0).astype(np.float32)
    return torch.from_numpy(ref), torch.from_numpy(pred), torch.from_numpy(mask)


def main():
    sys.path.insert(0, "/root/reference/PerceptualSimilarity")
    import lpips
    torch.set_num_threads(os.cpu_count())
    m = lpips.LPIPS(net="alex", spatial=True, pnet_rand=True, verbose=False)          # loads the shipped linear heads (lpips.py:109)
    trunk_init(m.net)
    ref, pred, mask = case()
    pred = pred.clone().requires_grad_()
    ex = m.forward(ref, pred)                                                           # test_network.py:39
    metric = torch.stack([torch.sum(mask * ex[i, 0]) / torch.sum(mask * ex[i, 0] > 0) for i in range(2)])      # test_network.py:45
    metric.sum().backward()
    heads = {"lin%d" % k: m.lins[k].model[1].weight.detach().numpy().reshape(-1) for k in range(5)}
    np.savez_compressed(os.path.join(OUT, "lpips.npz"), map=ex.detach().numpy(), metric=metric.detach().numpy(),
                        grad_pred_s2=pred.grad.numpy()[:, :, ::2, ::2].copy(), **heads)
    print("lpips.npz", os.path.getsize(os.path.join(OUT, "lpips.npz")), metric.detach().numpy())


if __name__ == "__main__":
    main()
