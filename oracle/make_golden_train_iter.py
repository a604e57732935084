"""TEST INFRASTRUCTURE — generates tests/golden/train_iter.npz and tests/golden/patchgan.npz by running the UNMODIFIED
reference training script (/root/reference/train_raytracing_relighting_CelebAHQ_DSSIM_8x.py = TRAIN, imported through
oracle/ref_shims.py) in the authoring container.

    python -m oracle.make_golden_train_iter          # from the repo root, ~5 min on 8 cores

train_iter.npz — the reference's own `main()` (TRAIN:560-685) executed for THREE iterations (j = 0, 1, 2; j = 0 updates
  the discriminator, TRAIN:624) of one epoch, twice: epoch 0 (every encoder-skip gate closed) and epoch 15 (every gate
  open, TRAIN:245,258,271,283).  Nothing of the loop body is restated: `main` runs as shipped, with only its
  environment intercepted —
      load_data()            -> 9 deterministic samples (`inputs()` below; the test rebuilds the same arrays)
      model.cuda()           -> (TRAIN:563) loads the shipped epoch-99 weights into the reference's RelightNet
      patchgan.cuda()        -> (TRAIN:567) sets the reference PatchGAN's parameters by `patchgan_init` (numpy RandomState,
                                so the test can rebuild the same start without an 11 MB fixture)
      np.random.shuffle      -> identity (batches 0, 1, 2 in order)
      range(1000), range(700)-> one epoch (0 or 15), three iterations
      print / scipy.io.savemat / torch.save -> captured: the 11 printed loss terms of every iteration, the epoch's
                                `losses` dict, the final state_dicts of both networks
  Stored: per-iteration loss terms [2 epochs, 3 iterations, 11], the final generator parameters as deltas from epoch 99
  (fp16 of delta * 1e4: |delta| <= 3 lr), its BatchNorm buffers, every 4th element of the final discriminator deltas.

patchgan.npz — `TRAIN.PatchGAN` (TRAIN:15-35) alone: logits, input gradient (every 2nd pixel) and parameter-gradient
  samples for a seeded input, B = 2, train mode.
"""
import builtins
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

LOSS_PRINT_ORDER = ("total", "recon", "depth", "ambient", "lighting", "albedo", "generator", "discriminator",
                    "discriminator_real", "discriminator_fake", "DSSIM")             # the print order of TRAIN:657-667
EPOCHS = (0, 15)
N_ITERS = 3
D_STRIDE = 4


def inputs():
    """`load_data()`'s six arrays (TRAIN:527-558 layouts and value ranges) for 9 samples, built from committed fixtures and
    closed-form generators only, so tests rebuild exactly what the reference trained on."""
    from oracle import relight_oracle as O
    f = np.load(os.path.join(OUT, "ffhq.npz"))
    n = 9
    idx = [(3 * i + 1) % 10 for i in range(n)]
    images = f["q"][idx].astype(np.float64) / 1020.0
    lightings = np.zeros((n, 4))
    lightings[:, 0] = 0.5                                                            # TRAIN:541
    lightings[:, 1:4] = np.array([O.LIGHTS_18[(5 * i + 3) % 18] for i in range(n)])
    faces = [O.synthetic_face(seed=20 + i, noise=2.0) for i in range(n)]
    depths = np.stack([d.numpy().astype(np.float64) for d, _ in faces])[..., None] * 0.5
    skin = f["masks"][idx].astype(np.float64)[..., None]                             # values {0, 64, 128, 255}
    masks = np.where(skin > 0, 255.0, 0.0)
    masks[:, 100:140, 90:170] = 0.0                                                  # depth masks exclude nose / mouth
    tmp = np.maximum(np.stack([m.numpy().astype(np.float64) * 255.0 for _, m in faces])[..., None], masks)
    fill = np.where(tmp > 128, 255.0, 0.0)                                           # TRAIN:552-556
    albedo = np.random.RandomState(11).randint(60, 200, (n, 256, 256)).astype(np.float64)
    return images, lightings, depths, masks, albedo, fill


def patchgan_init(module, seed=5):
    """Deterministic PatchGAN start, independent of torch's RNG stream: N(0, 0.02) weights, small biases, BN gamma ~ 1."""
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.startswith("bn") and name.endswith("weight"):
                v = 1.0 + 0.05 * rs.standard_normal(p.shape)
            elif name.endswith("bias"):
                v = 0.01 * rs.standard_normal(p.shape)
            else:
                v = 0.02 * rs.standard_normal(p.shape)
            p.copy_(torch.from_numpy(v.astype(np.float32)))
    return module


def _run_reference_main(epoch):
    """-> (per-iteration loss terms [N_ITERS, 11], the epoch `losses` dict, generator state_dict, patchgan state_dict)."""
    import scipy.io
    from oracle import ref_shims
    TRAIN = ref_shims.load_reference("TRAIN")
    printed, saved, mats = [], {}, {}
    RefNet, RefD = TRAIN.RelightNet, TRAIN.PatchGAN

    def net_cuda(self):                       # `model = model.cuda()` (TRAIN:563): where the start weights are put in
        self.load_state_dict(torch.load(os.path.join(OUT, "model_epoch99.pth"), map_location="cpu"))
        return self

    def fake_range(n):
        if n == 1000:
            return iter([epoch])
        if n == 700:
            return builtins.range(N_ITERS)
        return builtins.range(n)

    def fake_print(*a, **k):
        if len(a) == 1 and isinstance(a[0], str) and " loss: " in a[0]:
            printed.append(float(a[0].split(": ")[1]))

    def fake_save(obj, path, *a, **k):
        saved["patchgan" if "patchgan" in os.path.basename(path) else "model"] = {n: v.detach().clone() for n, v in obj.items()}

    RefNet.cuda, RefD.cuda = net_cuda, (lambda self: patchgan_init(self))            # TRAIN:563 / 567
    patched = dict(load_data=inputs, range=fake_range, print=fake_print)
    old_shuffle, old_savemat, old_tsave, old_load = np.random.shuffle, scipy.io.savemat, torch.save, TRAIN.load_data
    np.random.shuffle = lambda a: None
    scipy.io.savemat = lambda path, d, *a, **k: mats.update(d)
    torch.save = fake_save
    for k, v in patched.items():
        setattr(TRAIN, k, v)
    try:
        TRAIN.main()                                              # the reference's own training loop, as shipped
    finally:
        np.random.shuffle, scipy.io.savemat, torch.save = old_shuffle, old_savemat, old_tsave
        delattr(TRAIN, "range")
        delattr(TRAIN, "print")
        del RefNet.cuda, RefD.cuda
        TRAIN.load_data = old_load
    terms = np.array(printed, np.float64).reshape(N_ITERS, len(LOSS_PRINT_ORDER))
    return terms, mats, saved["model"], saved["patchgan"]


def make_train_iter():
    sd0 = torch.load(os.path.join(OUT, "model_epoch99.pth"), map_location="cpu")
    from oracle import relight_oracle as O
    d0 = patchgan_init(O.PatchGANOracle()).state_dict()
    out = {"epochs": np.array(EPOCHS), "loss_names": np.array(LOSS_PRINT_ORDER)}
    for e in EPOCHS:
        terms, mats, g_sd, d_sd = _run_reference_main(e)
        print("epoch", e, "terms\n", terms)
        out["terms_e%d" % e] = terms
        out["epoch_mean_e%d" % e] = np.array([float(mats[k]) for k in LOSS_PRINT_ORDER])       # TRAIN:671-682 (sum / 700)
        g_delta, bufs, d_delta = [], {}, []
        for k, v in g_sd.items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                bufs["g_e%d_%s" % (e, k)] = v.numpy().astype(np.float32)
            elif not k.endswith("num_batches_tracked"):
                g_delta.append(((v - sd0[k]).double() * 1e4).reshape(-1))
        out["g_delta_e%d" % e] = torch.cat(g_delta).numpy().astype(np.float16)
        out.update(bufs)
        for k, v in d_sd.items():
            if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")):
                d_delta.append(((v - d0[k]).double() * 1e4).reshape(-1)[::D_STRIDE])
        out["d_delta_e%d" % e] = torch.cat(d_delta).numpy().astype(np.float16)
    np.savez_compressed(os.path.join(OUT, "train_iter.npz"), **out)
    print("train_iter.npz", os.path.getsize(os.path.join(OUT, "train_iter.npz")))


def patchgan_case():
    """Seeded input / output-gradient of the PatchGAN fixture (also rebuilt by the tests)."""
    rs = np.random.RandomState(21)
    img = torch.from_numpy(rs.uniform(0, 1, (2, 3, 256, 256)).astype(np.float32))
    gl = torch.from_numpy(rs.standard_normal((2, 1, 15, 15)).astype(np.float32))
    return img, gl


def make_patchgan():
    from oracle import ref_shims
    TRAIN = ref_shims.load_reference("TRAIN")
    D = patchgan_init(TRAIN.PatchGAN()).float()                  # train mode, like TRAIN:564-568
    img, gl = patchgan_case()
    x = img.clone().requires_grad_()
    logits = D(x)
    (logits * gl).sum().backward()
    out = {"logits": logits.detach().numpy(), "grad_input_s2": x.grad.numpy()[:, :, ::2, ::2].copy()}
    for n, p in D.named_parameters():
        g = p.grad.reshape(-1)
        out["gsum_" + n] = np.float64(g.double().sum())
        out["gabs_" + n] = np.float64(g.double().abs().sum())
        out["gsample_" + n] = g[:: max(1, g.numel() // 4096)].numpy().copy()
    for n, b in D.named_buffers():
        if "num_batches" not in n:
            out["buf_" + n] = b.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "patchgan.npz"), **out)
    print("patchgan.npz", os.path.getsize(os.path.join(OUT, "patchgan.npz")))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    make_patchgan()
    make_train_iter()
