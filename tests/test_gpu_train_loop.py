"""GPU: the training driver (train_loop.Trainer, TRAIN:592-685) on a tiny synthetic data set with the reference's batch
size 3: reference-named checkpoint files that the reference's own loaders accept, graph-replayed steps equal to eager
steps, and save -> resume -> continue equal to an uninterrupted run."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _data(n=12):
    from geomconsistentfr_b200 import train_loop as TL
    from geomconsistentfr_b200.synthetic import LIGHTS_18, synthetic_face
    f = np.load(os.path.join(G, "ffhq.npz"))
    g = np.random.default_rng(0)
    images = np.stack([f["q"][i % 10] / 1020.0 for i in range(n)])
    faces = [synthetic_face(seed=i, noise=2.0) for i in range(n)]
    depths = np.stack([d.numpy() for d, _ in faces])[..., None].astype(np.float64) * 0.5
    masks = np.stack([f["masks"][i % 10] for i in range(n)])[..., None].astype(np.float64)
    fill = TL.binarise_fill_mask(masks, np.stack([m.numpy() * 255 for _, m in faces])[..., None].astype(np.float64))
    light = np.array([[0.5, *LIGHTS_18[(3 * i + 4) % 18]] for i in range(n)])
    albedo = g.integers(60, 200, (n, 256, 256)).astype(np.float64)
    return TL.TrainingArrays(images, light, depths, masks, albedo, fill)


def _models(seed=0):
    from geomconsistentfr_b200 import PatchGAN, RelightNet
    net = RelightNet(batch_size=3)
    net.load_state_dict(torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu"), strict=True)
    torch.manual_seed(seed)
    return net.float().cuda().train(), PatchGAN().cuda()


def _flat(tr):
    return torch.cat([tr.step.opt.flat, tr.step.opt_d.flat]).clone()


def test_trainer_writes_reference_files_and_resumes(tmp_path):
    import scipy.io
    from geomconsistentfr_b200 import train_loop as TL
    data = _data()
    ITERS = 6                                      # j = 0 and j = 5 update the discriminator (GD_ratio 5, TRAIN:624)

    # uninterrupted: epochs 8 and 9 (the s1 gate opens at epoch 9 -> a re-capture in the middle of the run)
    net, D = _models()
    a = TL.Trainer(net, D, data, str(tmp_path / "a"), iters_per_epoch=ITERS, max_epoch=10, seed=3, log=None)
    a.start_epoch = 8
    p0 = _flat(a)
    hist = a.train()
    assert len(hist) == 2 and set(hist[0]) == set(TL.LOSS_NAMES)
    assert all(np.isfinite(list(h.values())).all() for h in hist)
    assert float((_flat(a) - p0).abs().max()) > 1e-5                                   # it trained
    paths = TL.checkpoint_paths(str(tmp_path / "a"), 9)
    mat = scipy.io.loadmat(paths["losses"])
    assert all(abs(float(mat[k]) - hist[1][k]) < 1e-12 for k in TL.LOSS_NAMES)          # TRAIN:671-683
    # the reference's loaders accept the files (TEST1:509 load_state_dict; torch.optim.Adam for the added state)
    from oracle import relight_oracle as O
    o = O.RelightNetOracle()
    o.load_state_dict(torch.load(paths["model"], map_location="cpu"), strict=True)
    st = torch.load(paths["trainer"], map_location="cpu")
    torch.optim.Adam(o.parameters(), lr=1.0).load_state_dict(st["optimizer"])
    assert float(st["optimizer"]["state"][0]["step"]) == 2 * ITERS
    assert float(st["optimizer_patchgan"]["state"][0]["step"]) == 2 * 2                # D stepped at j = 0, 5 of each epoch

    # interrupted after epoch 8, resumed in a fresh process state, continued through epoch 9
    net, D = _models()
    b = TL.Trainer(net, D, data, str(tmp_path / "b"), iters_per_epoch=ITERS, max_epoch=9, seed=3, log=None)
    b.start_epoch = 8
    b.train()
    net, D = _models(seed=123)                                                          # different init: must be overwritten
    c = TL.Trainer(net, D, data, str(tmp_path / "b"), iters_per_epoch=ITERS, max_epoch=10, seed=99, log=None)
    assert c.resume() == 8 and c.start_epoch == 9 and c.seed == 3
    h2 = c.train()
    # same trajectory up to the run-to-run noise of the atomics in the weight-gradient kernels
    # (Adam turns the ~1e-8 rounding noise of analytically-zero gradients - conv biases in front of batch-stat BN - into
    # +-lr updates, so a few parameters may differ by up to 2 * lr * steps; everything else agrees to 2e-5)
    d = (_flat(c) - _flat(a)).abs()
    assert float(d.max()) <= 2 * 1e-4 * 2 * ITERS + 1e-6 and float((d > 2e-5).float().mean()) < 0.02, (float(d.max()), float((d > 2e-5).float().mean()))
    for k in TL.LOSS_NAMES:
        assert abs(h2[0][k] - hist[1][k]) <= 2e-3 * max(1.0, abs(hist[1][k])), k
    for (k, v), (_, w) in zip(sorted(a.net.state_dict().items()), sorted(c.net.state_dict().items())):
        assert torch.allclose(v.float(), w.float(), atol=3e-3, rtol=1e-3), k             # BN running statistics too


def test_graph_replay_equals_eager_steps(tmp_path):
    from geomconsistentfr_b200 import train_loop as TL
    data = _data(6)
    res = {}
    for graph in (True, False):
        net, D = _models()
        t = TL.Trainer(net, D, data, str(tmp_path / ("g%d" % graph)), iters_per_epoch=2, max_epoch=1, seed=1, use_graph=graph, log=None)
        res[graph] = (t.train()[0], _flat(t))
    for k in TL.LOSS_NAMES:
        assert abs(res[True][0][k] - res[False][0][k]) <= 1e-3 * max(1.0, abs(res[False][0][k])), k
    d = (res[True][1] - res[False][1]).abs()
    assert float(d.max()) <= 2 * 1e-4 * 2 + 1e-6 and float((d > 2e-5).float().mean()) < 0.02
