"""CPU: pins the TRAINING side of the oracle against what the reference's own code produced (tests/golden/patchgan.npz and
train_iter.npz, made by oracle/make_golden_train_iter.py from the UNMODIFIED TRAIN script):
  * `PatchGANOracle` == `TRAIN.PatchGAN` (TRAIN:15-35): logits, input gradient, parameter gradients, BN buffers;
  * one iteration (j = 0, epoch 15: every gate open, discriminator updated) of the oracle training loop == the first
    iteration of the reference's `main()`: all 11 printed loss terms."""
import os

import numpy as np
import torch

from oracle import relight_oracle as O
from oracle.make_golden_train_iter import inputs, patchgan_case, patchgan_init

G = os.path.join(os.path.dirname(__file__), "golden")


def test_patchgan_oracle_equals_reference_patchgan():
    f = np.load(os.path.join(G, "patchgan.npz"))
    D = patchgan_init(O.PatchGANOracle()).train()
    img, gl = patchgan_case()
    x = img.clone().requires_grad_()
    logits = D(x)
    (logits * gl).sum().backward()
    assert np.abs(logits.detach().numpy() - f["logits"]).max() <= 1e-6
    gi = f["grad_input_s2"]
    assert np.abs(x.grad.numpy()[:, :, ::2, ::2] - gi).max() <= 1e-6 * np.abs(gi).max() + 1e-9
    for n, p in D.named_parameters():
        g = p.grad.reshape(-1)
        ref = f["gsample_" + n]
        got = g[:: max(1, g.numel() // 4096)].numpy()
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-7, n
        assert abs(float(g.double().abs().sum()) - float(f["gabs_" + n])) <= 1e-5 * float(f["gabs_" + n]) + 1e-7, n
    for n, b in D.named_buffers():
        if "num_batches" not in n:
            assert np.abs(b.numpy() - f["buf_" + n]).max() <= 1e-6, n


def test_oracle_training_iteration_equals_reference_main():
    from _train_iter_driver import NAMES, oracle_iteration
    from geomconsistentfr_b200.train_loop import TrainingArrays
    f = np.load(os.path.join(G, "train_iter.npz"))
    assert tuple(f["loss_names"]) == NAMES
    torch.set_num_threads(os.cpu_count() or 1)
    data = TrainingArrays(*inputs())
    net = O.RelightNetOracle()
    net.load_state_dict(torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu"))
    net.train()
    D = patchgan_init(O.PatchGANOracle()).train()
    opt, opt_d = torch.optim.Adam(net.parameters(), lr=1e-4), torch.optim.Adam(D.parameters(), lr=1e-4)
    got = oracle_iteration(net, D, opt, opt_d, data.batch(0, 3), 15, 0, O.intrinsic_matrix())
    want = f["terms_e15"][0]
    for k, a, b in zip(NAMES, got, want):
        assert abs(a - b) <= 2e-5 * max(abs(b), 1e-3), (k, a, b)
