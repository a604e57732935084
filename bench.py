#!/usr/bin/env python
"""bench.py — relit faces/s at 256x256 (ray-march shadow + shading + RelightNet CNN), BASELINE.json configs[1]:
batch 8 per GPU, full relight forward, fp32, eval mode, epoch-99 weights, synthetic inputs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A "step" is one forward over one batch of 8 faces per GPU.
  value      faces/s, whole job, inputs resident in HBM: the K steps rotate over the runner lanes and a 151 MB pool of
             distinct device batches (> L2), one CUDA-event pair around all K steps, max over ranks
  latency    the same forward on one lane, CUDA events per step, L2 flushed between steps
  e2e        faces/s through RelightRunner.relight_host: pinned host image/mask/light -> H2D -> forward -> composite ->
             D2H of the 8-bit BGR images the reference's driver stores (TEST1:590-620), all inside the timed region
  roofline   the kernel of the step that does the ray march (one fused launch: march + normals + Lambert + render):
             algorithmic bytes / CUDA-event duration, L2 flushed between launches; the stand-alone march beside it
  train      (sub-block of the same line) configs[2]/[3]: the reference's full training iteration TRAIN:617-656, B = 16
             per GPU, ONE NCCL all-reduce of the flat generator gradient per step (+ the discriminator's every 5th)
  sweep      (sub-block) configs[4]: 18 light directions x 8 faces per GPU per step, one CNN pass per face
  cpu_baseline / --impl reference: the UNMODIFIED reference forward (TEST1.RelightNet.forward imported from
             oracle/_ref through oracle/ref_shims.py, torch CPU, all host threads) on a bounded sample; the oracle port
             when oracle/_ref is absent
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

B_PER_GPU = 8
B_TRAIN = 16
H = W = 256
MARCH_BYTES_PER_FACE = 786432          # depth f32 + mask f32 + d_min f32
# the fused march+shade launch: depth + mask in (d_min stays on the SM), + albedo in, + shadow/full/final + rendered + normals out
FUSED_BYTES_PER_FACE = 2 * 262144 + 786432 + 3 * 262144 + 2 * 786432
MARCH_SAMPLES_PER_FACE = 160 * H * W
CNN_FLOP_PER_FACE = 4.54e9
METRIC = "relit faces/sec @256x256 (shadow+CNN)"
WORKLOAD = ("configs[1]: batch 8 per GPU, full relight forward 256x256 fp32 "
            "(RelightNet CNN + normals + 160-sample ray-march + Lambert render), eval, epoch-99 weights")


def make_config(world):
    """The `config` object of BOTH arms (ours and --impl reference): the workload, not how an arm executes it."""
    return {"workload": WORKLOAD, "global_batch": world * B_PER_GPU,
            "parallelism": "dp%d (faces sharded, no collective)" % world,
            "l2": "every step reads a different batch of a 151 MB pool (> 126 MB L2); latency + roofline legs write a "
                  "256 MiB flush between timed launches (untimed)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def synthetic_batch(B, seed):
    """Synthetic 256x256 faces: image U(0,1) (seeded), elliptical face mask, one of the 18 light directions."""
    from geomconsistentfr_b200.synthetic import synthetic_batch as sb
    return sb(B, seed, H, W)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self._stop_evt.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j] == "Active" for r in self.rows)]
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm (CPU) and the reference on one GPU — the only places that execute anything under oracle/
# ----------------------------------------------------------------------------------------------------------------
def _reference_forward_fn(device, ref_batch):
    """-> (fn(img, mask01, light) running ONE forward of `ref_batch` faces, kind).  kind "reference": the unmodified
    TEST1.RelightNet.forward from oracle/_ref (or /root/reference in the authoring container) through oracle/ref_shims;
    "port": oracle/relight_oracle.py when no copy of the reference is present."""
    from oracle import ref_shims as R
    from oracle import relight_oracle as O
    K = O.intrinsic_matrix()
    if R.reference_available():
        net = R.reference_model("TEST1", batch_size=ref_batch, cuda_identity=(device == "cpu")).eval()
        if device != "cpu":
            net = net.cuda()                                  # TEST1:511
            K = K.cuda()

        def fn(img, m01, light):                              # the call at TEST1:588
            B = img.shape[0]
            if device != "cpu":
                img, m01, light = img.cuda(), m01.cuda(), light.cuda()
            return net(img, 200, K, m01, light.view(B, 3, 1, 1), torch.full((B, 1, 1), 0.5, device=img.device),
                       m01.repeat(B, 1, 1, 1))
        return fn, "reference"
    net = O.RelightNetOracle()
    net.load_state_dict(torch.load(os.path.join(GOLDEN, "model_epoch99.pth"), map_location="cpu"))
    net = net.eval()
    if device != "cpu":
        net, K = net.cuda(), K.cuda()

    def fn(img, m01, light):
        if device != "cpu":
            img, m01, light = img.cuda(), m01.cuda(), light.cuda()
        return net.forward_test(img, 200, K, m01, light)
    return fn, "port"


def reference_times(n_steps, faces_per_step, ref_batch, device="cpu"):
    """Wall-clock seconds of each of `n_steps` steps; a step relights `faces_per_step` faces in forwards of `ref_batch`
    (the reference bakes batch_size = 1 into its pixel grids, TEST1:15,25-26, so its B = 8 is 8 forwards)."""
    fn, kind = _reference_forward_fn(device, ref_batch)
    times = []
    with torch.no_grad():
        for i in range(n_steps):
            img, mask, light = synthetic_batch(faces_per_step, 100 + i)
            m01 = mask.view(H, W, 1).double() / 255.0
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            for b in range(0, faces_per_step, ref_batch):
                out = fn(img[b:b + ref_batch], m01, light[b:b + ref_batch])
                float(out[5].sum())                           # the reference's `.cpu().numpy()` read-back (TEST1:591)
            if device != "cpu":
                torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    return times, kind


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host threads, rank 0 only.  Each step is
    a bounded sample of the workload (`--ref-faces` faces, default 1, relit one per forward like TEST1:15)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    gpu = args.impl == "reference-gpu"
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    total = args.warmup + args.steps
    times, kind = reference_times(total, args.ref_faces, args.ref_batch, "cuda" if gpu else "cpu")
    times = times[args.warmup:]
    ms = 1e3 * sum(times) / len(times)
    v = args.ref_faces * 1e3 / ms
    what = ("the UNMODIFIED reference TEST1.RelightNet.forward (oracle/_ref via oracle/ref_shims.py)" if kind == "reference"
            else "torch oracle port of TEST1:169-505 (oracle/_ref absent)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "faces/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": make_config(max(1, args.gpus)),
        "cpu_baseline": {"value": v, "unit": "faces/s", "cores": 0 if gpu else threads, "kind": kind,
                         "device": "cuda:0 (torch eager, the reference's own .cuda() calls)" if gpu else "cpu",
                         "sample": "%d steps x %d face(s), forwards of B=%d (the reference bakes batch_size = 1, TEST1:15): %s"
                                   % (len(times), args.ref_faces, args.ref_batch, what)},
        "e2e": {"value": v, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def reference_subprocess(impl, steps, warmup, extra=()):
    """Runs `bench.py --impl reference[-gpu]` in a child process (the reference needs torch's `.cuda()` patched to the
    identity on the CPU — that must not happen inside this process) and returns its parsed JSON line."""
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    if impl == "reference":
        env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", impl, "--steps", str(steps), "--warmup", str(warmup)] + list(extra)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("reference arm failed: %s" % (r.stderr.strip()[-300:] or r.stdout.strip()[-300:]))


# ----------------------------------------------------------------------------------------------------------------
def _max_over_ranks(ms, dist):
    if dist is None:
        return ms
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_train(args, rank, world, dist):
    """configs[2]/[3]: the reference's training iteration (TRAIN:617-656), B = 16 per GPU, synthetic batch resident on the
    device: train-mode RelightNet forward, PatchGAN x3, all seven loss terms, backward, ONE NCCL all-reduce of the flat
    generator gradient buffer per step (+ the discriminator's on every GD_ratio-th step, TRAIN:624), fused Adam; the whole
    iteration replayed from two CUDA graphs (D-update and no-D-update iterations)."""
    from geomconsistentfr_b200 import RelightNet, intrinsic_matrix, ops
    from geomconsistentfr_b200 import PatchGAN
    from geomconsistentfr_b200.trainer import GeneratorStep, TrainStep
    from geomconsistentfr_b200.synthetic import LIGHTS_18, synthetic_face
    B = B_TRAIN
    net = RelightNet(batch_size=B)
    net.load_state_dict(torch.load(os.path.join(GOLDEN, "model_epoch99.pth"), map_location="cpu"), strict=True)
    net = net.float().cuda().train()
    if args.train_precision:
        net.train_precision = args.train_precision
    torch.manual_seed(0)
    full = not args.no_gan
    step = TrainStep(net, PatchGAN().cuda(), intrinsic_matrix().cuda()) if full else GeneratorStep(net, intrinsic_matrix().cuda())
    g = torch.Generator().manual_seed(rank)
    img = torch.rand(B, H, W, 3, generator=g).cuda()
    faces = [synthetic_face(seed=rank * B + i) for i in range(B)]
    mf = torch.stack([f[1] for f in faces]).float().cuda()
    depth_gt = (torch.stack([f[0] for f in faces]) * 0.5).cuda()
    albedo_gt = torch.rand(B, H, W, generator=g).cuda()
    light_gt = torch.tensor([[0.5, *LIGHTS_18[(rank + i) % 18]] for i in range(B)], dtype=torch.float32).cuda()
    batch = (mf, mf, depth_gt, albedo_gt, light_gt)

    opts = [step.opt] + ([step.opt_d] if full else [])
    if world > 1:
        for o in opts:                                       # CUDA events around the (captured) NCCL all-reduce node
            o.ar_events = (torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))
    n_pre = ops.launch_count()
    step.step(img, 200, *batch)
    launches_d_step = ops.launch_count() - n_pre
    n_pre = ops.launch_count()
    if full:
        step.step(img, 200, *batch, j=1)
    launches_g_step = ops.launch_count() - n_pre if full else launches_d_step
    it = [0]                                                 # iteration counter: the discriminator updates every GD_ratio-th (TRAIN:624)

    def kw():
        it[0] += 1
        return dict(j=it[0] - 1) if full else {}

    if args.no_graph:
        stream = torch.cuda.current_stream()
        one = lambda: step.step(img, 200, *batch, **kw())
    else:
        step.capture(img, 200, *batch)
        stream = step._stream
        one = lambda: step.step_graphed(img, *batch, **kw())
    warm = max(args.warmup, 3)
    for _ in range(warm):
        one()
    it[0] = 0
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        total, _ = one()
    e1.record(stream)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = _max_over_ranks(e0.elapsed_time(e1), dist)
    n_d = len([j for j in range(args.steps) if j % 5 == 0]) if full else 0
    launches = n_d * launches_d_step + (args.steps - n_d) * launches_g_step

    ar = {"bytes_per_step_generator": step.opt.grad.numel() * 4,
          "bytes_every_5th_step_discriminator": step.opt_d.grad.numel() * 4 if full else 0}
    if world > 1:
        def ar_us(o):
            """CUDA-event time of the all-reduce inside the last replay; a stand-alone all-reduce of the same buffer when
            event-record nodes of a graph cannot be read back."""
            try:
                return 1e3 * o.ar_events[0].elapsed_time(o.ar_events[1]), "events around the NCCL node inside the last replayed step graph"
            except Exception:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    dist.all_reduce(o.grad)
                    a.record(stream)
                    for _ in range(10):
                        dist.all_reduce(o.grad)
                    b.record(stream)
                stream.synchronize()
                return 1e2 * a.elapsed_time(b), "10 back-to-back all-reduces of the same buffer (stand-alone)"
        gus, how = ar_us(step.opt)
        ar.update(generator_us=_max_over_ranks(gus, dist), how=how, backend="nccl", world=world)
        if full:
            ar["discriminator_us"] = _max_over_ranks(ar_us(step.opt_d)[0], dist)
    prec = getattr(net, "train_precision", 3)
    return {"metric": "training faces/sec @256x256 (%s)" % ("full iteration TRAIN:617-656: generator + PatchGAN" if full else "generator step"),
            "value": world * B * args.steps * 1e3 / ms, "unit": "faces/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "dtype": {3: "f32 (3xTF32 convs)", 1: "tf32 convs, fp32 march", 4: "bf16 CNN / fp32 ray-march"}.get(prec, str(prec)),
            "data": "synthetic",
            "config": {"workload": "configs[2]/[3]: training iteration, batch 16 per GPU, 256x256: train-mode RelightNet fwd + masked "
                                   "recon/depth/albedo + ambient + light + DSSIM losses + backward + flat-gradient all-reduce + fused Adam"
                                   + ("; PatchGAN x3 passes, discriminator Adam step every 5th iteration (TRAIN:617-656)" if full
                                      else "; PatchGAN terms skipped (--no-gan)"),
                       "global_batch": world * B, "parallelism": "dp%d, one all_reduce of %.1f MB per step" % (world, step.opt.grad.numel() * 4 / 1e6),
                       "working_set": "activations of one step (> L2) are rewritten every step", "cuda_graph": not args.no_graph},
            "allreduce": ar, "gpu_launches": int(launches), "final_loss": float(total)}


def bench_sweep(args, rank, world, dist):
    """configs[4]: the 18-light Multi-PIE sweep (TESTB:565-583) - every face relit under 18 light directions with ONE CNN
    pass per face (`RelightNet.relight_sweep`, lights_per_face = 18 in the march/shade launch); faces are sharded over the
    ranks with no collective.  One step = 8 faces x 18 lights = 144 relit images per GPU, captured in a CUDA graph."""
    from geomconsistentfr_b200 import RelightNet, intrinsic_matrix, ops
    from geomconsistentfr_b200.synthetic import LIGHTS_18
    net = RelightNet()
    net.load_state_dict(torch.load(os.path.join(GOLDEN, "model_epoch99.pth"), map_location="cpu"), strict=True)
    net = net.float().cuda().eval()
    B, L = B_PER_GPU, 18
    K = intrinsic_matrix()
    lights = torch.tensor(LIGHTS_18, dtype=torch.float32).cuda()
    pool = [tuple(t.cuda() for t in synthetic_batch(B, 2000 * rank + i)) for i in range(8)]
    img, mask = pool[0][0].clone(), pool[0][1].clone()
    stream = torch.cuda.Stream()
    warm = max(args.warmup, 3)
    with torch.cuda.stream(stream):
        for _ in range(2):
            n0 = ops.launch_count()
            out = net.relight_sweep(img, 200, K, mask.view(H, W, 1), lights)
            per_step = ops.launch_count() - n0
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            out = net.relight_sweep(img, 200, K, mask.view(H, W, 1), lights)

        def one(i):
            img.copy_(pool[i % len(pool)][0], non_blocking=True)
            mask.copy_(pool[i % len(pool)][1], non_blocking=True)
            graph.replay()

        for i in range(warm):
            one(i)
        stream.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            one(i)
        e1.record(stream)
        stream.synchronize()
    if dist is not None:
        dist.barrier()
    ms = _max_over_ranks(e0.elapsed_time(e1), dist)
    return {"metric": "relit images/sec @256x256, 18-light sweep (one CNN pass per face)", "value": world * B * L * args.steps * 1e3 / ms,
            "unit": "images/s", "faces_per_s": world * B * args.steps * 1e3 / ms, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "configs[4]: 18 Multi-PIE light directions x 8 faces per GPU per step (512 faces = 64 steps of one GPU "
                                   "or 8 steps of eight), CNN once per face, march + shade for 144 (face, light) pairs in one launch",
                       "global_batch": world * B, "lights": L, "parallelism": "dp%d (faces sharded, no collective)" % world,
                       "working_set": "144 relit images = 170 MB of outputs per step (> L2)", "cuda_graph": True},
            "gpu_launches": int(per_step * args.steps), "rendered_shape": list(out["rendered"].shape)}


def bench_forward(args, rank, world, local, dist):
    from geomconsistentfr_b200 import RelightNet, RelightRunner, ops
    net = RelightNet()
    net.load_state_dict(torch.load(os.path.join(GOLDEN, "model_epoch99.pth"), map_location="cpu"), strict=True)
    net = net.float().cuda().eval()
    B = B_PER_GPU
    runner = RelightRunner(net, B, use_graph=not args.no_graph, lanes=args.lanes)
    stream = runner.stream

    # distinct synthetic batches, host-pinned (for e2e) and device-resident (for value)
    n_pool = 24                                                          # 24 x 6.3 MB of images = 151 MB > L2 (126 MB)
    batches = [synthetic_batch(B, 1000 * rank + i) for i in range(n_pool)]
    host = [tuple(t.pin_memory() for t in b) for b in batches[:4]]
    dev = [tuple(t.cuda() for t in b) for b in batches]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """CUDA events around each step on the launching stream; L2 flushed (untimed) between steps."""
        for i in range(warmup):
            fn(i)
        barrier()
        pairs = []
        with torch.cuda.stream(stream):
            for i in range(steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn(i)
                e1.record(stream)
                pairs.append((e0, e1))
        barrier()
        return _max_over_ranks(sum(a.elapsed_time(b) for a, b in pairs), dist)

    # ---- value: whole-job throughput, inputs resident in HBM.  The K steps rotate over the runner lanes (the latency-
    # bound low-resolution layers of one forward overlap the machine-filling layers of another) and over a pool of
    # distinct device-resident batches that is larger than L2 (24 x 6.3 MB = 151 MB > 126 MB), so no step finds its
    # inputs cached; one event pair brackets all K steps (start on lane 0, every lane waits for it; end = last lane).
    # `latency` is the same forward on ONE lane, one event pair per step, L2 flushed (untimed) between steps.
    def step_device(i):
        runner.set_inputs(*dev[i % n_pool])
        runner.run()

    def timed_lanes(submit, steps, warmup):
        for i in range(warmup):
            submit(i)
        runner.synchronize()
        barrier()
        start = torch.cuda.Event(enable_timing=True)
        start.record(runner.lanes[0].stream)
        for lane in runner.lanes[1:]:
            lane.stream.wait_event(start)
        for i in range(steps):
            submit(warmup + i)
        ends = []
        for lane in runner.lanes:
            e = torch.cuda.Event(enable_timing=True)
            e.record(lane.stream)
            ends.append(e)
        runner.synchronize()
        barrier()
        return _max_over_ranks(max(start.elapsed_time(e) for e in ends), dist)

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    n0 = ops.launch_count()
    latency_ms = timed(step_device, args.steps, args.warmup) / args.steps
    ms_per_step = timed_lanes(lambda i: runner.relight_resident(*dev[i % n_pool]), args.steps, args.warmup) / args.steps
    launches = runner.launches_per_run * args.steps if runner.graph is not None else \
        (ops.launch_count() - n0) * args.steps // (2 * (args.steps + args.warmup))
    value = world * B * 1e3 / ms_per_step

    # ---- e2e: host buffers through the public runner API (RelightRunner.relight_host), copies inside the timed
    # region.  The runner rotates over its lanes, so the H2D of step i+1 / D2H of step i-1 overlap the kernels of step
    # i; the K steps are bracketed by one event pair (start on lane 0, every lane waits for it; end = the last lane to
    # finish).  Every step streams a different host batch; the per-step working set (~1.1 GB of activations) exceeds L2.
    # The result read back is what the reference's driver keeps of a forward: the 8-bit BGR composite (TEST1:590-620).
    out_kind = "rendered_f32" if args.e2e_f32 else "bgr_u8"
    e2e_ms = timed_lanes(lambda i: runner.relight_host(*host[i % len(host)], output=out_kind), args.steps, args.warmup) / args.steps
    clocks = sampler.stop()
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = B * 3 * H * W * (4 if args.e2e_f32 else 1)

    # ---- roofline: the kernel of the step that does the ray march.  In the forward it is ONE fused launch
    # (march_shade_fwd_kernel: every thread marches its ray, then shades its pixel; d_min never leaves the SM), timed
    # here on the depth / albedo maps of the last forward with CUDA events on the launching stream, L2 flushed between
    # launches.  The host is allowed to run ahead of the device (a device-side sleep is queued first), so the event
    # pairs bracket device time only.  The stand-alone march kernel (the operator ShadowMarch uses) is timed beside it.
    albedo, depth = runner.out[0], runner.out[1]
    bits = ops.mask_pack(dev[0][1].view(1, H, W))
    light_pt = (net.light_distance * torch.nn.functional.normalize(dev[0][2].view(B, 3), dim=1)).contiguous()
    amb = torch.full((B,), 0.4, device="cuda")

    def step_fused(i):
        ops.march_shade_fwd(albedo, depth, bits, light_pt, amb, inside_bonus=5.0)

    def step_march(i):
        ops.shadow_march_fwd(depth, bits, light_pt, inside_bonus=5.0, variant=0)

    def timed_kernel(fn, steps):
        for i in range(3):
            fn(i)
        barrier()
        pairs = []
        with torch.cuda.stream(stream):
            torch.cuda._sleep(40_000_000)                 # ~20 ms: the whole loop below is enqueued before it ends
            for i in range(steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn(i)
                e1.record(stream)
                pairs.append((e0, e1))
        barrier()
        return statistics.median(a.elapsed_time(b) for a, b in pairs)

    n_k = max(args.steps, 20)
    fused_ms = timed_kernel(step_fused, n_k)
    march_ms = timed_kernel(step_march, n_k)

    # ---- the CNN's representative tensor-core layer, timed the same way: 16 -> 16 channels at 256^2 (three of them per
    # decoder: 46 % of the CNN's MACs sit at this resolution) on P16 activations; algorithmic bytes = the
    # activation read once + written once (4 B per element each), weights negligible
    g = torch.Generator(device="cuda").manual_seed(7)
    xa = ops.nchw_to_p16(torch.randn(B, 16, H, W, device="cuda", generator=g))
    wl = torch.randn(16, 16, 3, 3, device="cuda", generator=g) / 12.0
    bl = torch.zeros(16, device="cuda")
    wpk, wsc = ops.conv_p16_pack_weights(wl, 16, 2)
    conv_ms = timed_kernel(lambda i: ops.conv3x3_p16_fwd(xa, wpk, bl, 16, (16, 2, 2), wsc), n_k)
    ops.conv_p16_config(2)           # the same launch with two CTAs per SM: fastest alone, slower in the overlapped step
    conv_ms_2 = timed_kernel(lambda i: ops.conv3x3_p16_fwd(xa, wpk, bl, 16, (16, 2, 2), wsc), n_k)
    ops.conv_p16_config(0)
    conv_bytes = 2 * B * 16 * H * W * 4
    conv_flop = 2.0 * B * H * W * 16 * 16 * 9
    bf16_peak = None
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(mp):
        bf16_peak = json.load(open(mp)).get("bf16_tflops")
    del xa
    hbm_peak, peak_src = peaks()
    achieved = FUSED_BYTES_PER_FACE * B / (fused_ms * 1e-3) / 1e9
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "march_traffic.json")
    if os.path.isfile(tp):
        traffic = json.load(open(tp))

    # the bound that actually binds: warp-instruction issue (ncu instruction count of the same launch / live duration
    # against 148 SMs x 4 schedulers x SM clock)
    issue = None
    if traffic.get("fused_inst_executed") and B == 8:
        peak_ginst = 148 * 4 * (clocks.get("sm_mhz") or 1965.0) * 1e-3
        ach = traffic["fused_inst_executed"] / (fused_ms * 1e-3) / 1e9
        issue = {"bound": "warp-instruction issue", "warp_inst_per_launch": traffic["fused_inst_executed"],
                 "achieved": ach, "peak": peak_ginst, "unit": "G warp-inst/s", "frac": ach / peak_ginst,
                 "source": traffic.get("source", "smsp__inst_executed.sum, profiles/march_traffic.json")}

    cfg = make_config(world)
    line = {
        "metric": METRIC, "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "execution": {"lanes": len(runner.lanes), "cnn": net.cnn_impl, "tc_precision": net.tc_precision,
                      "cuda_graph": runner.graph is not None},
        "e2e": {"value": world * B * 1e3 / e2e_ms, "unit": "faces/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "lanes": len(runner.lanes),
                "note": "RelightRunner.relight_host: pinned host image/mask/light -> H2D -> forward -> "
                        + ("D2H rendered (fp32)" if args.e2e_f32 else "device composite -> D2H of the 8-bit BGR images the reference's "
                           "driver stores (TEST1:590-620)") + ", steps pipelined over the runner lanes, one event pair around all K steps"},
        "latency": {"ms_per_step": latency_ms, "faces_per_s": world * B * 1e3 / latency_ms,
                    "note": "one forward of 8 faces on ONE lane, CUDA events per step, L2 flushed between steps"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "march_shade_fwd_kernel (ray march + normals + Lambert + render, one launch)", "bound": "hbm",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic.get("fused_dram_bytes_per_launch"), "peak_source": peak_src,
                     "ms_per_launch": fused_ms, "share_of_step": fused_ms / latency_ms,
                     "algorithmic_bytes_per_face": FUSED_BYTES_PER_FACE,
                     "note": "the kernel is instruction-issue bound by construction (160 samples x ~90 instructions per "
                             "in-mask pixel against 56 algorithmic bytes per pixel), not HBM bound - DESIGN.md 3/K1; "
                             "gsamples_per_s counts the reference's 160 samples for every pixel",
                     "gsamples_per_s": MARCH_SAMPLES_PER_FACE * B / (fused_ms * 1e-3) / 1e9,
                     "issue": issue,
                     "march_only": {"kernel": "shadow_march_fwd_fast", "ms_per_launch": march_ms,
                                    "algorithmic_bytes_per_face": MARCH_BYTES_PER_FACE,
                                    "achieved": MARCH_BYTES_PER_FACE * B / (march_ms * 1e-3) / 1e9,
                                    "frac": MARCH_BYTES_PER_FACE * B / (march_ms * 1e-3) / 1e9 / hbm_peak,
                                    "traffic": traffic.get("dram_bytes_per_launch"),
                                    "gsamples_per_s": MARCH_SAMPLES_PER_FACE * B / (march_ms * 1e-3) / 1e9}},
        "cnn": {"flop_per_face": CNN_FLOP_PER_FACE,
                "note": "tensor-pipe utilisation of the conv kernel: profiles/ (ncu sm__inst_executed_pipe_tensor / pipe_tensor cycles active)"},
        "roofline_cnn": {"kernel": "conv3x3_p16_kernel<16,2,2> (tcgen05, 16 -> 16 channels @256^2, B = %d, L2 flushed between launches)" % B,
                         "bound": "hbm", "achieved": conv_bytes / (conv_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": conv_bytes / (conv_ms * 1e-3) / 1e9 / hbm_peak, "traffic": 34_100_000, "peak_source": peak_src,
                         "ms_per_launch": conv_ms, "algorithmic_bytes_per_launch": conv_bytes,
                         "grid": "one persistent CTA per SM (the step's default: the free slot overlaps another lane's kernel)",
                         "two_ctas_per_sm": {"ms_per_launch": conv_ms_2, "achieved": conv_bytes / (conv_ms_2 * 1e-3) / 1e9,
                                             "frac": conv_bytes / (conv_ms_2 * 1e-3) / 1e9 / hbm_peak},
                         "tensor": {"achieved": conv_flop / (conv_ms * 1e-3) / 1e12, "peak": bf16_peak, "unit": "TFLOP/s",
                                    "frac": (conv_flop / (conv_ms * 1e-3) / 1e12 / bf16_peak) if bf16_peak else None,
                                    "note": "useful MACs only; the fp16-pair split issues 3 products per MAC (hi*W1, hi*W2, lo*W1), "
                                            "so the tensor pipe does 3x this"},
                         "note": "traffic: ncu dram__bytes_read + write of the same launch, profiles/r02_ncu_conv_p16_full.csv (33.6 MB read, "
                                 "0.5 MB written when the kernel ends: the output is still dirty in L2); the layer is bound by the MMAs' "
                                 "shared-memory operand reads (sm__pipe_tc_cycles_active 69.5 % of the cycles an SM is active), DESIGN.md 3/K3"},
    }
    del runner, dev, flush
    torch.cuda.empty_cache()
    return line


def pin_rank_cores(local, world):
    """N > 1: every rank keeps to its own slice of the host cores (the box's GPUs all report the same CPU affinity, so 8
    ranks x (launch thread + copy threads + NCCL proxy) otherwise migrate over the same cores)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if per >= 2:
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
            return per
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--cpu-faces", type=int, default=5, help="faces in the bounded CPU-baseline sample")
    ap.add_argument("--ref-faces", type=int, default=1, help="--impl reference: faces relit per step (bounded sample of the 8-face batch)")
    ap.add_argument("--ref-batch", type=int, default=1, help="--impl reference: faces per forward (the reference bakes 1; 8 patches its grids)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-gpu-ref", action="store_true", help="skip the informational run of the reference on cuda:0 "
                                                              "(~50k tiny launches; always skip it under ncu)")
    ap.add_argument("--lanes", type=int, default=3, help="runner lanes the e2e (host-buffer) path rotates over")
    ap.add_argument("--e2e-f32", action="store_true", help="e2e reads back fp32 rendered_images (6.3 MB) instead of the 8-bit BGR composite")
    ap.add_argument("--no-gan", action="store_true", help="train workload without the PatchGAN terms")
    ap.add_argument("--train-precision", type=int, default=4, help="train-mode conv operand precision: 4 = bf16 operands / fp32 accumulation "
                                                                   "(BASELINE configs[2]: 'bf16 CNN / fp32 ray-march', the default), 3 = 3xTF32 "
                                                                   "(fp32-grade: what the parity tests run), 1 = TF32")
    ap.add_argument("--no-pin", action="store_true", help="do not pin each rank to its own host cores")
    ap.add_argument("--workload", default="all", choices=["all", "forward", "train", "sweep"],
                    help="all (default) = the configs[1] forward line with `train` (configs[2]/[3]) and `sweep` (configs[4]) "
                         "sub-blocks; forward / train / sweep = that leg alone")
    args = ap.parse_args()
    if args.impl != "ours":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dist = None
    pinned = None
    if world > 1:
        if not args.no_pin:
            pinned = pin_rank_cores(local, world)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.workload == "train":
        line = bench_train(args, rank, world, dist)
    elif args.workload == "sweep":
        line = bench_sweep(args, rank, world, dist)
    else:
        line = bench_forward(args, rank, world, local, dist)
        if pinned:
            line["execution"]["host_cores_per_rank"] = pinned
        if args.workload == "all":
            line["sweep"] = bench_sweep(args, rank, world, dist)
            line["train"] = bench_train(args, rank, world, dist)
        if rank == 0 and world == 1:
            try:
                r = reference_subprocess("reference", max(args.cpu_faces, 2), 1)
                line["cpu_baseline"] = r["cpu_baseline"]
            except Exception as e:
                line["cpu_baseline"] = {"unavailable": str(e)[:300]}
            if not args.no_gpu_ref:
                try:
                    r = reference_subprocess("reference-gpu", 6, 2)
                    line["reference_gpu"] = dict(r["cpu_baseline"], note="the reference's forward on cuda:0 (torch eager, cuDNN convs + the "
                                                 "per-sample tensor program of TEST1:351-498 with its per-image host syncs): the "
                                                 "north star's 'reference 1-GPU PyTorch throughput'")
                except Exception as e:                      # informational only
                    line["reference_gpu"] = {"unavailable": str(e)[:300]}
    if dist is not None:
        # the step graphs hold captured NCCL kernels: tearing the communicator down under them was observed to hang at
        # interpreter exit on 2 GPUs, so synchronise, agree that everybody is done, print, and leave without finalisers
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps(line), flush=True)          # the last thing on stdout (NCCL_DEBUG=INFO also writes there)
    if dist is not None:
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
