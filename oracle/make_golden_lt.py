"""TEST INFRASTRUCTURE — generates tests/golden/lt.npz and copies model_epoch106.pth by running the UNMODIFIED
lighting-transfer reference (test_relight_single_image_lighting_transfer.py = TEST_LT, via oracle/ref_shims.py) in the
authoring container, exactly as its main() does (TEST_LT:516-556): pass 1 on the REFERENCE image with a zero target
light yields the estimated light / ambient; pass 2 relights the INPUT image with them.

    python -m oracle.make_golden_lt          # from the repo root, ~20 s

Inputs come from tests/golden/ffhq.npz (input 00295, reference image 00110 — lit from the right —, the input's skin
mask), resized to 256x256 as TEST1:515 does; TEST_LT itself reads 256x256 files named on its command line.
A third forward with a hand-picked nearby light (|L_xy| < 4*256 after the 4013 scale: the `+5` branch of TEST_LT:503)
and one with a far light (no bonus) cover both sides of the lighting-transfer bonus rule."""
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    torch.set_num_threads(os.cpu_count())
    shutil.copyfile(os.path.join(ref_shims.REFERENCE_ROOT, "model_lighting_transfer", "model_epoch106.pth"),
                    os.path.join(OUT, "model_epoch106.pth"))
    f = np.load(os.path.join(OUT, "ffhq.npz"))
    names = list(f["names"])
    i_in, i_ref = names.index("00295"), names.index("00110")
    net = ref_shims.reference_model("TEST_LT").eval()
    K = np.zeros((1, 3, 3))                                                # TEST_LT:528-535
    K[:, 0, 0] = K[:, 1, 1] = 700.0
    K[:, 2, 2] = 1.0
    K[:, 0, 2] = K[:, 1, 2] = 128.0
    K = torch.from_numpy(K)
    x_in = torch.from_numpy((f["q"][i_in] / 1020.0)[None]).float()
    x_ref = torch.from_numpy((f["q"][i_ref] / 1020.0)[None]).float()
    m = torch.from_numpy(f["masks"][i_in].astype(np.float64).reshape(256, 256, 1)) / 255.0      # TEST_LT:542
    zero = torch.zeros(1, 4)
    out = {"input": "00295", "reference": "00110"}
    with torch.no_grad():
        r1 = net(x_ref, 200, K, m, zero[:, 1:4].float().view(1, 3, 1, 1), zero[:, 0].float().view(1, 1, 1))   # TEST_LT:543
        est_l, est_a = r1[10].clone(), r1[11].clone()
        r2 = net(x_in, 200, K, m, est_l.float().view(1, 3, 1, 1), est_a.float().view(1, 1, 1))              # TEST_LT:545
        out.update(est_light=est_l.numpy().reshape(3).astype(np.float32), est_ambient=est_a.numpy().reshape(1).astype(np.float32))
        for key, idx in (("depth", 1), ("shadow", 2), ("rendered", 5), ("unit_light", 6),
                         ("ambient", 7), ("final", 8), ("normals", 9), ("est_light_in", 10), ("est_ambient_in", 11)):
            out["transfer_" + key] = r2[idx].numpy().astype(np.float32)
        for tag, L, amb in (("near", (0.12, -0.2, 0.9724), 0.45), ("far", (-0.7574, 0.0, 0.6529), 0.3)):
            tl = torch.tensor(L, dtype=torch.float32).view(1, 3, 1, 1)
            r = net(x_in, 200, K, m, tl, torch.full((1, 1, 1), amb))
            out[tag + "_light"] = np.array(L, np.float32)
            out[tag + "_ambient"] = np.float32(amb)
            out[tag + "_shadow"] = r[2].numpy().astype(np.float32)
            print(tag, "light point", (4013 * torch.nn.functional.normalize(tl, dim=1)).view(3).tolist())
    np.savez_compressed(os.path.join(OUT, "lt.npz"), **out)

    # ---- lt_train.npz: the UNMODIFIED train_lighting_transfer.py (TRAIN_LT) model, train() mode, B = 3 (hard-coded at
    # TRAIN_LT:359), forward + autograd gradients of the recon + DSSIM terms (TRAIN_LT:640,650), epoch-106 weights
    from oracle import relight_oracle as O
    tnet = ref_shims.reference_model("TRAIN_LT")                       # stays in train() mode like TRAIN_LT:568-570
    sel = [names.index(n) for n in ("00104", "00322", "00508")]
    x = torch.from_numpy(f["q"][sel] / 1020.0).float()
    mk = (f["masks"][sel] > 128).astype(np.float64).reshape(3, 256, 256, 1)
    mt = torch.from_numpy(mk)
    r = tnet(x, 200, O.intrinsic_matrix(), mt)
    albedo, depth_t, shadow, _, _, rendered, unit_l, amb_v = r
    depth_t.retain_grad()
    m3 = mt.permute(0, 3, 1, 2).repeat(1, 3, 1, 1).float()
    target = x.permute(0, 3, 1, 2)
    comp = rendered * m3 + (1.0 - m3) * target
    loss = 20.0 * torch.sum((rendered * m3 - target * m3) ** 2) / torch.sum(m3) \
        + 8.0 * (1 - O.ssim(comp, target, data_range=1.0, size_average=True, nonnegative_ssim=True)) / 2.0
    loss.backward()
    np.savez_compressed(
        os.path.join(OUT, "lt_train.npz"), sel=np.array(sel), masks01=mk.astype(np.uint8)[..., 0],
        depth=depth_t.detach().numpy(), shadow=shadow.detach().numpy().astype(np.float16),
        rendered=rendered.detach().numpy().astype(np.float16), unit_light=unit_l.detach().numpy().reshape(3, 3),
        ambient=amb_v.detach().numpy().reshape(3), loss=np.float64(loss.item()),
        albedo_mean=albedo.detach().numpy().mean(axis=(2, 3)),
        grad_depth=depth_t.grad.numpy().astype(np.float16),
        grad_shortcut_h1=tnet.conv_shortcut_h1_out.weight.grad.numpy(),
        grad_shortcut_h3=tnet.conv_shortcut_h3_out.weight.grad.numpy(),
        grad_deconv_depth_shortcut_h6=tnet.deconv_depth_shortcut_h6_out.weight.grad.numpy(),
        grad_sl2_w=tnet.linear_SL2.weight.grad.numpy(), grad_depth_head_w=tnet.conv_depth_c2_o.weight.grad.numpy())
    print("train_lt done, loss", loss.item())
    print("est light", out["est_light"], "est ambient", out["est_ambient"])
    for fn in ("lt.npz", "lt_train.npz", "model_epoch106.pth"):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
