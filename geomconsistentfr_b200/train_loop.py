"""The reference's training driver (TRAIN:560-685) around `trainer.TrainStep`: epochs x iterations over in-memory
arrays shaped like `load_data()`'s (TRAIN:527-558), epoch-gated skip connections, G/D alternation, per-epoch mean
losses -> `losses_epoch<i>.mat`, `model_epoch<i>.pth`, `patchgan_epoch<i>.pth` (the reference's file names and
formats, TRAIN:671-685), plus what the reference lacks: optimiser/RNG state (`trainer_epoch<i>.pth`) and `resume()`.

What differs from the reference loop, none of it in the numbers:
  * a step is a CUDA-graph replay (re-captured when an epoch gate of TRAIN:245,258,271,283 flips: epochs 9, 11, 13, 15);
  * the reference's 11 `.item()` device syncs per iteration (TRAIN:627-668) become one running sum on the device, read
    back once per epoch (or every `log_every` iterations);
  * data parallel: rank r of `world` takes batch `batch_list[j * world + r]` of the shared shuffle; gradients are
    averaged by the single flat all-reduce inside the step.  world = 1 is the reference's schedule exactly.
TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py."""
import os
from dataclasses import dataclass

import numpy as np
import torch

from .relightnet import _EPOCH_GATES, intrinsic_matrix

# the reference's loss names, in the order of TRAIN:672-682
LOSS_NAMES = ("total", "recon", "depth", "ambient", "lighting", "albedo", "generator", "discriminator",
              "discriminator_real", "discriminator_fake", "DSSIM")


@dataclass
class TrainingArrays:
    """`load_data()`'s six arrays (TRAIN:527-558), any float/uint8 dtype, values as the reference stores them:
    images [N,H,W,3] in [0,1]; lightings [N,4] = (0.5, lx, ly, lz); depths [N,H,W,1]; masks [N,H,W,1] in 0..255;
    albedo [N,H,W] in 0..255; masks_fill [N,H,W,1] in {0,255} (binarised at 128, TRAIN:552-556)."""
    images: np.ndarray
    lightings: np.ndarray
    depths: np.ndarray
    masks: np.ndarray
    albedo: np.ndarray
    masks_fill: np.ndarray

    def __len__(self):
        return self.images.shape[0]

    def batch(self, k, B):
        """Batch k = rows [k*B, (k+1)*B) converted as TRAIN:607-615 does (masks and albedo / 255), as host tensors in
        the layout TrainStep takes."""
        s = slice(k * B, (k + 1) * B)
        H, W = self.images.shape[1:3]
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return (t(self.images[s]), t(self.masks_fill[s].reshape(-1, H, W) / 255.0), t(self.masks[s].reshape(-1, H, W) / 255.0),
                t(self.depths[s].reshape(-1, H, W)), t(self.albedo[s].reshape(-1, H, W) / 255.0), t(self.lightings[s]))


def binarise_fill_mask(face_mask, depth_mask):
    """TRAIN:552-556: max of the face mask and the depth mask, > 128 -> 255 else 0."""
    tmp = np.maximum(face_mask, depth_mask).astype(np.float64)
    return np.where(tmp > 128, 255.0, 0.0)


def gate_signature(epoch):
    """Which encoder-skip blocks are active at this epoch (TRAIN:245,258,271,283): the captured graphs depend on it."""
    return tuple(epoch > g for _, g in sorted(_EPOCH_GATES.items()))


def epoch_schedule(n_samples, batch_size, iters, world, seed, epoch):
    """The batch index every (iteration j, rank r) trains on in `epoch`: row [j, r] of the result.
    The reference shuffles `np.arange(N // B)` in place at the start of every epoch with the global numpy RNG
    (TRAIN:593) and walks its first 700 entries; here the permutation of epoch e is `RandomState(seed + e)` so that a
    resumed run reproduces it, and with world > 1 rank r takes entry j*world + r (wrapping if the epoch is longer than
    the data)."""
    n_batches = n_samples // batch_size
    if n_batches < 1:
        raise ValueError("dataset smaller than one batch")
    order = np.arange(n_batches)
    np.random.RandomState((seed + epoch) % (2 ** 32)).shuffle(order)
    idx = (np.arange(iters)[:, None] * world + np.arange(world)[None, :]) % n_batches
    return order[idx]


def checkpoint_paths(out_dir, epoch):
    """The reference's per-epoch files (TRAIN:683-685) + the trainer state it does not save."""
    return {"losses": os.path.join(out_dir, "losses", "losses_epoch%d.mat" % epoch),
            "model": os.path.join(out_dir, "saved_epochs", "model_epoch%d.pth" % epoch),
            "patchgan": os.path.join(out_dir, "saved_epochs", "patchgan_epoch%d.pth" % epoch),
            "trainer": os.path.join(out_dir, "saved_epochs", "trainer_epoch%d.pth" % epoch)}


def latest_epoch(out_dir):
    d = os.path.join(out_dir, "saved_epochs")
    if not os.path.isdir(d):
        return None
    done = [int(f[len("trainer_epoch"):-4]) for f in os.listdir(d) if f.startswith("trainer_epoch") and f.endswith(".pth")]
    done = [e for e in done if all(os.path.isfile(p) for k, p in checkpoint_paths(out_dir, e).items() if k != "losses")]
    return max(done) if done else None


class Trainer:
    def __init__(self, net, patchgan, data, out_dir, iters_per_epoch=700, max_epoch=1000, seed=0, use_graph=True,
                 group=None, log_every=0, log=print):
        from .trainer import TrainStep
        import torch.distributed as dist
        self.net, self.D, self.data, self.out_dir = net, patchgan, data, out_dir
        self.B = net.batch_size                                       # TRAIN:41: the model owns the batch size
        self.iters, self.max_epoch, self.seed = iters_per_epoch, max_epoch, seed
        self.use_graph, self.log_every, self.log = use_graph, log_every, log
        self.rank, self.world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)
        dev = net.device
        if dev.type != "cuda":
            raise RuntimeError("Trainer needs the model on a CUDA device (there is no CPU path)")
        H, W = data.images.shape[1:3]
        self.step = TrainStep(net, patchgan, intrinsic_matrix(H, W).to(dev), group=group)
        self.start_epoch = 0
        self._captured_for = None
        self._pinned = None

    # ---- checkpoints: the reference's files + resumable trainer state
    def save(self, epoch, losses):
        import scipy.io
        p = checkpoint_paths(self.out_dir, epoch)
        if self.rank != 0:
            return p
        for f in p.values():
            os.makedirs(os.path.dirname(f), exist_ok=True)
        scipy.io.savemat(p["losses"], {k: float(v) for k, v in losses.items()})                       # TRAIN:683
        torch.save({k: v.detach().cpu() for k, v in self.net.state_dict().items()}, p["model"])        # TRAIN:684
        torch.save({k: v.detach().cpu() for k, v in self.D.state_dict().items()}, p["patchgan"])       # TRAIN:685
        torch.save({"epoch": epoch, "seed": self.seed, "optimizer": self.step.opt.state_dict(),
                    "optimizer_patchgan": self.step.opt_d.state_dict(), "losses": dict(losses)}, p["trainer"])
        return p

    def resume(self, epoch=None):
        """Load the checkpoint of `epoch` (default: the latest complete one in out_dir); training continues at epoch+1.
        Returns the epoch loaded, or None if there is nothing to resume from."""
        epoch = latest_epoch(self.out_dir) if epoch is None else epoch
        if epoch is None:
            return None
        p = checkpoint_paths(self.out_dir, epoch)
        self.net.load_state_dict(torch.load(p["model"], map_location="cpu"), strict=True)
        self.D.load_state_dict(torch.load(p["patchgan"], map_location="cpu"), strict=True)
        st = torch.load(p["trainer"], map_location="cpu")
        self.step.opt.load_state_dict(st["optimizer"])
        self.step.opt_d.load_state_dict(st["optimizer_patchgan"])
        self.seed = st["seed"]
        self.start_epoch = epoch + 1
        self._captured_for = None                                     # parameters were rewritten in place; graphs stay valid,
        return epoch                                                  # but re-capture anyway if the gates moved

    # ---- one epoch (TRAIN:592-682)
    def _stage(self, k, slot, stream):
        """Host batch k -> pinned staging slot -> device (async on `stream`); two slots alternate, guarded by events."""
        if self._pinned is None:
            self._pinned = [[torch.empty_like(t).pin_memory() for t in self.data.batch(0, self.B)] for _ in range(2)]
            self._pinned_free = [None, None]
        if self._pinned_free[slot] is not None:
            self._pinned_free[slot].synchronize()                       # the H2D that last used this slot has finished
        host = self.data.batch(k, self.B)
        out = []
        for dst, src in zip(self._pinned[slot], host):
            dst.copy_(src)
            out.append(dst.to(self.net.device, non_blocking=True))
        ev = torch.cuda.Event()
        ev.record(stream)
        self._pinned_free[slot] = ev
        return out

    def run_epoch(self, epoch):
        sched = epoch_schedule(len(self.data), self.B, self.iters, self.world, self.seed, epoch)[:, self.rank]
        dev = self.net.device
        graphed = self.use_graph
        if graphed and self._captured_for != gate_signature(epoch):
            self._snapshot_and_capture([t.to(dev) for t in self.data.batch(int(sched[0]), self.B)], epoch)
        stream = self.step._stream if graphed else torch.cuda.current_stream()
        with torch.cuda.stream(stream):
            acc = torch.zeros(len(LOSS_NAMES), dtype=torch.float64, device=dev)
            for j in range(self.iters):
                cur = self._stage(int(sched[j]), j & 1, stream)         # host work of step j+1 overlaps the replay of step j
                if graphed:
                    total, terms = self.step.step_graphed(*cur, j=j)
                else:
                    total, terms = self.step.step(cur[0], epoch, *cur[1:], j=j)
                acc += torch.stack([total.double()] + [terms[k].double() for k in LOSS_NAMES[1:]])
                if self.log_every and (j + 1) % self.log_every == 0:
                    v = (acc / (j + 1)).tolist()                          # one device sync per log line
                    self.log("Epoch: %d, Batch: %d  " % (epoch, j) + "  ".join("%s %.5f" % kv for kv in zip(LOSS_NAMES, v)))
            mean = (acc / self.iters).tolist()
        return dict(zip(LOSS_NAMES, mean))                              # TRAIN:671-682

    def _snapshot_and_capture(self, batch, epoch):
        """Graph capture runs warm-up steps that would move the parameters, BN buffers and Adam state: snapshot and
        restore them around it so that capturing is invisible to the training trajectory."""
        opt, opt_d = self.step.opt, self.step.opt_d
        saved = [t.clone() for t in (opt.flat, opt.exp_avg, opt.exp_avg_sq, opt.seg_state, opt_d.flat, opt_d.exp_avg, opt_d.exp_avg_sq, opt_d.seg_state)]
        bufs = [b for m in (self.net, self.D) for b in m.buffers()]
        saved_bufs = [b.clone() for b in bufs]
        self.step.capture(batch[0], epoch, *batch[1:])
        torch.cuda.synchronize()
        with torch.no_grad():
            for dst, src in zip((opt.flat, opt.exp_avg, opt.exp_avg_sq, opt.seg_state, opt_d.flat, opt_d.exp_avg, opt_d.exp_avg_sq, opt_d.seg_state), saved):
                dst.copy_(src)
            for dst, src in zip(bufs, saved_bufs):
                dst.copy_(src)
        self._captured_for = gate_signature(epoch)

    def train(self):
        """TRAIN:592-685 from `start_epoch` (0, or the epoch after the one `resume()` loaded)."""
        history = []
        for epoch in range(self.start_epoch, self.max_epoch):
            losses = self.run_epoch(epoch)
            self.save(epoch, losses)
            history.append(losses)
            if self.rank == 0 and self.log:
                self.log("epoch %d: " % epoch + "  ".join("%s %.5f" % kv for kv in losses.items()))
        return history
