"""CPU: host logic of the training driver (geomconsistentfr_b200/train_loop.py, TRAIN:527-558 and 592-685): the batch
schedule, epoch gates, checkpoint naming / discovery, array conversion, and optimiser-state interchange with
torch.optim.Adam (the reference's optimiser, TRAIN:589-590)."""
import os

import numpy as np
import pytest
import torch

from geomconsistentfr_b200 import train_loop as TL
from geomconsistentfr_b200.autograd import FlatAdam


def test_epoch_schedule_world1_is_a_prefix_of_one_shuffle():
    s = TL.epoch_schedule(n_samples=29890, batch_size=3, iters=700, world=1, seed=7, epoch=4)
    assert s.shape == (700, 1)
    order = np.arange(29890 // 3)
    np.random.RandomState(11).shuffle(order)                 # TRAIN:593 with a reproducible per-epoch seed
    assert np.array_equal(s[:, 0], order[:700])
    assert len(set(s[:, 0].tolist())) == 700                 # no batch twice within an epoch
    assert not np.array_equal(s, TL.epoch_schedule(29890, 3, 700, 1, 7, 5))


def test_epoch_schedule_ranks_are_disjoint_and_cover_the_same_shuffle():
    w = 4
    s = TL.epoch_schedule(n_samples=1600, batch_size=16, iters=25, world=w, seed=0, epoch=0)
    assert s.shape == (25, w)
    assert sorted(s.reshape(-1).tolist()) == list(range(100))        # 25 x 4 = every batch exactly once
    one = TL.epoch_schedule(1600, 16, 100, 1, 0, 0)[:, 0]
    assert np.array_equal(s.reshape(-1), one)                        # rank r takes entry j*world + r
    with pytest.raises(ValueError):
        TL.epoch_schedule(10, 16, 5, 1, 0, 0)


def test_gate_signature_flips_where_the_reference_gates_do():
    flips = [e for e in range(1, 30) if TL.gate_signature(e) != TL.gate_signature(e - 1)]
    assert flips == [9, 11, 13, 15]                                  # `epoch > 8/10/12/14`, TRAIN:245,258,271,283
    assert TL.gate_signature(0) == (False,) * 4 and TL.gate_signature(200) == (True,) * 4


def test_checkpoint_names_follow_the_reference(tmp_path):
    p = TL.checkpoint_paths(str(tmp_path), 12)
    assert os.path.basename(p["model"]) == "model_epoch12.pth"       # TRAIN:684
    assert os.path.basename(p["patchgan"]) == "patchgan_epoch12.pth"  # TRAIN:685
    assert os.path.basename(p["losses"]) == "losses_epoch12.mat"     # TRAIN:683
    assert TL.latest_epoch(str(tmp_path)) is None
    os.makedirs(os.path.dirname(p["model"]))
    for e, complete in ((3, True), (5, True), (7, False)):
        q = TL.checkpoint_paths(str(tmp_path), e)
        open(q["trainer"], "w").close()
        if complete:
            open(q["model"], "w").close(); open(q["patchgan"], "w").close()
    assert TL.latest_epoch(str(tmp_path)) == 5                       # epoch 7 was cut off mid-save
    assert len(TL.LOSS_NAMES) == 11 and TL.LOSS_NAMES[0] == "total" and TL.LOSS_NAMES[-1] == "DSSIM"


def test_training_arrays_convert_like_the_reference():
    g = np.random.default_rng(0)
    N, H, W = 7, 8, 8
    face = g.integers(0, 256, (N, H, W, 1)).astype(np.float64)
    dmask = g.integers(0, 256, (N, H, W, 1)).astype(np.float64)
    fill = TL.binarise_fill_mask(face, dmask)
    assert set(np.unique(fill)) <= {0.0, 255.0}
    assert np.array_equal(fill == 255.0, np.maximum(face, dmask) > 128)            # TRAIN:553-555 (128 itself -> 0)
    d = TL.TrainingArrays(images=g.random((N, H, W, 3)), lightings=g.random((N, 4)), depths=g.random((N, H, W, 1)) * 100,
                          masks=dmask, albedo=g.integers(0, 256, (N, H, W)).astype(np.float64), masks_fill=fill)
    img, mf, m, dep, alb, lit = d.batch(1, 3)
    assert img.shape == (3, H, W, 3) and img.dtype == torch.float32
    assert torch.equal(img, torch.from_numpy(d.images[3:6]).float())               # TRAIN:618 `.float()`
    assert torch.equal(mf, torch.from_numpy(fill[3:6, :, :, 0] / 255.0).float()) and set(mf.unique().tolist()) <= {0.0, 1.0}
    assert torch.equal(m, torch.from_numpy(dmask[3:6, :, :, 0] / 255.0).float())   # TRAIN:610
    assert torch.equal(alb, torch.from_numpy(d.albedo[3:6] / 255.0).float())       # TRAIN:615
    assert dep.shape == (3, H, W) and lit.shape == (3, 4) and len(d) == N


def test_flat_adam_state_dict_interchanges_with_torch_adam():
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ref = torch.optim.Adam(ps, lr=1e-4)
    for _ in range(3):
        for p in ps:
            p.grad = torch.randn_like(p)
        ref.step()
    mine = FlatAdam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=3e-4)
    mine.load_state_dict(ref.state_dict())                            # the reference's optimiser state loads here
    assert mine.lr == 1e-4 and mine.seg_state[:, 0].tolist() == [3.0, 3.0]
    assert abs(float(mine.seg_state[0, 1]) - (1 - 0.9 ** 3)) < 1e-6 and abs(float(mine.seg_state[1, 2]) - (1 - 0.999 ** 3) ** 0.5) < 1e-6
    assert torch.equal(mine.exp_avg[:12].view(4, 3), ref.state[ps[0]]["exp_avg"])
    assert torch.equal(mine.exp_avg_sq[12:], ref.state[ps[1]]["exp_avg_sq"])
    back = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1.0)
    back.load_state_dict(mine.state_dict())                           # ... and this one's loads into torch.optim.Adam
    st = back.state_dict()["state"]
    assert float(st[0]["step"]) == 3.0 and torch.equal(st[1]["exp_avg"], ref.state[ps[1]]["exp_avg"])
    assert back.param_groups[0]["lr"] == 1e-4
    fresh = FlatAdam([torch.nn.Parameter(torch.zeros(2))])
    assert fresh.state_dict()["state"] == {}                          # like an unstepped torch.optim.Adam
    fresh.load_state_dict(fresh.state_dict())
    assert float(fresh.seg_state[0, 0]) == 0.0
    # per-parameter steps: a torch.optim.Adam in which one parameter started late (an epoch-gated skip block,
    # TRAIN:245) loads, keeps its differing step counts, and goes back out unchanged
    late = torch.optim.Adam(ps, lr=1e-4)
    ps[1].grad = None
    ps[0].grad = torch.randn_like(ps[0]); late.step(); late.step()
    ps[1].grad = torch.randn_like(ps[1]); late.step()
    mine2 = FlatAdam([torch.nn.Parameter(p.detach().clone()) for p in ps])
    mine2.load_state_dict(late.state_dict())
    assert mine2.seg_state[:, 0].tolist() == [3.0, 1.0]
    out = mine2.state_dict()["state"]
    assert float(out[0]["step"]) == 3.0 and float(out[1]["step"]) == 1.0
    never = torch.optim.Adam(ps, lr=1e-4)
    ps[1].grad = None
    never.step()
    mine3 = FlatAdam([torch.nn.Parameter(p.detach().clone()) for p in ps])
    mine3.load_state_dict(never.state_dict())                          # parameter 1 has no state at all
    assert mine3.seg_state[:, 0].tolist() == [1.0, 0.0] and set(mine3.state_dict()["state"]) == {0}
