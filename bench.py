#!/usr/bin/env python
"""bench.py — relit faces/s at 256x256 (ray-march shadow + shading + RelightNet CNN), BASELINE.json configs[1]:
batch 8 per GPU, full relight forward, fp32, eval mode, epoch-99 weights, synthetic inputs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A "step" is one forward over one batch of 8 faces per GPU.
  value      faces/s, whole job, inputs resident in HBM: the K steps rotate over the runner lanes and a 151 MB pool of
             distinct device batches (> L2), one CUDA-event pair around all K steps, max over ranks
  latency    the same forward on one lane, CUDA events per step, L2 flushed between steps
  e2e        faces/s through RelightRunner.relight_host: pinned host image/mask/light -> H2D -> forward -> D2H of
             rendered_images, all inside the timed region
  roofline   the kernel of the step that does the ray march (one fused launch: march + normals + Lambert + render):
             algorithmic bytes / CUDA-event duration, L2 flushed between launches; the stand-alone march beside it
  cpu_baseline / --impl reference: the CPU oracle port of the reference forward (oracle/relight_oracle.py, torch
             CPU, all host threads) on a bounded sample
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
H = W = 256
MARCH_BYTES_PER_FACE = 786432          # depth f32 + mask f32 + d_min f32
# the fused march+shade launch: depth + mask in (d_min stays on the SM), + albedo in, + shadow/full/final + rendered + normals out
FUSED_BYTES_PER_FACE = 2 * 262144 + 786432 + 3 * 262144 + 2 * 786432
MARCH_SAMPLES_PER_FACE = 160 * H * W
CNN_FLOP_PER_FACE = 4.54e9
METRIC = "relit faces/sec @256x256 (shadow+CNN)"
WORKLOAD = ("configs[1]: batch 8 per GPU, full relight forward 256x256 fp32 "
            "(RelightNet CNN + normals + 160-sample ray-march + Lambert render), eval, epoch-99 weights")


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


def cpu_forward_faces_per_s(n_faces, threads):
    """The CPU oracle port of the reference forward (TEST1:169-505), B=1 per call like the reference."""
    from oracle import relight_oracle as O
    torch.set_num_threads(threads)
    net = O.RelightNetOracle()
    net.load_state_dict(torch.load(os.path.join(GOLDEN, "model_epoch99.pth"), map_location="cpu"))
    net.eval()
    K = O.intrinsic_matrix()
    times = []
    with torch.no_grad():
        for i in range(n_faces):
            img, mask, light = synthetic_batch(1, 100 + i)
            m = mask.view(H, W, 1).double() / 255.0
            t0 = time.perf_counter()
            net.forward_test(img, 200, K, m, light[:1])
            times.append(time.perf_counter() - t0)
    return times


def gpu_port_faces_per_s(n_faces):
    """The same oracle port on cuda:0 (torch eager: cuDNN convs + the per-sample tensor program of TEST1:351-498, with its
    per-image host syncs), B = 1 per call like the reference: the closest stand-in for "the reference's 1-GPU PyTorch
    throughput" that can run on this box (the reference itself cannot travel here)."""
    from oracle import relight_oracle as O
    net = O.RelightNetOracle()
    net.load_state_dict(torch.load(os.path.join(GOLDEN, "model_epoch99.pth"), map_location="cpu"))
    net = net.cuda().eval()
    K = O.intrinsic_matrix().cuda()
    times = []
    with torch.no_grad():
        for i in range(n_faces + 2):
            img, mask, light = synthetic_batch(1, 100 + i)
            img, light = img.cuda(), light.cuda()
            m = (mask.view(H, W, 1).double() / 255.0).cuda()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            net.forward_test(img, 200, K, m, light[:1])
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    return times[2:]


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port, kind 'port'), rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    total = args.warmup + args.steps
    times = cpu_forward_faces_per_s(total, threads)[args.warmup:]
    ms = 1e3 * sum(times) / len(times)
    v = 1e3 / ms
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "faces/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "step": "bounded sample of that workload: 1 face per step (the reference's own batch size, TEST1:15), "
                           "CPU oracle port on all host threads"},
        "cpu_baseline": {"value": v, "unit": "faces/s", "cores": threads, "kind": "port",
                         "sample": "%d faces, B=1 each, torch CPU oracle port of TEST1:169-505" % len(times)},
        "e2e": {"value": v, "unit": "faces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_train(args, rank, world, local, dist):
    """configs[2]/[3]: generator training step (TRAIN:618, 633-645 minus the PatchGAN terms, 655-656), B = 16 per GPU,
    synthetic batch resident on the device, one NCCL all-reduce of the flat gradient buffer per step."""
    from geomconsistentfr_b200 import RelightNet, intrinsic_matrix, ops
    from geomconsistentfr_b200 import PatchGAN
    from geomconsistentfr_b200.trainer import GeneratorStep, TrainStep
    from geomconsistentfr_b200.synthetic import LIGHTS_18, synthetic_face
    B = 16
    net = RelightNet(batch_size=B)
    net.load_state_dict(torch.load(os.path.join(GOLDEN, "model_epoch99.pth"), map_location="cpu"), strict=True)
    net = net.float().cuda().train()
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
    n_pre = ops.launch_count()
    step.step(img, 200, *batch)
    launches_per_step = ops.launch_count() - n_pre
    it = [0]                                             # iteration counter: the discriminator updates every GD_ratio-th (TRAIN:624)

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
    for _ in range(max(args.warmup, 3)):
        one()
    it[0] = 0
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    n0 = ops.launch_count() - launches_per_step * args.steps if not args.no_graph else ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        total, _ = one()
    e1.record(stream)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "metric": "training faces/sec @256x256 (%s)" % ("full iteration: generator + PatchGAN" if full else "generator step"), "value": world * B * args.steps * 1e3 / ms, "unit": "faces/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 convs)", "data": "synthetic",
            "config": {"workload": "configs[2]/[3]: generator training step, batch 16 per GPU, 256x256: train-mode RelightNet fwd + "
                                   "masked recon/depth/albedo + ambient + light + DSSIM losses + backward + flat-gradient "
                                   "all-reduce + fused Adam" + ("; PatchGAN x3 passes, discriminator Adam step every 5th iteration (TRAIN:617-656)" if full else
                                                               "; PatchGAN terms skipped (--no-gan)"), "global_batch": world * B,
                       "parallelism": "dp%d, one all_reduce of %.1f MB per step" % (world, step.opt.grad.numel() * 4 / 1e6),
                       "working_set": "activations of one step (> L2) are rewritten every step", "cuda_graph": not args.no_graph},
            "gpu_launches": int(ops.launch_count() - n0), "final_loss": float(total)}), flush=True)
    if dist is not None:
        # the step graphs hold captured NCCL kernels: tearing the communicator down under them was observed to hang at
        # interpreter exit on 2 GPUs, so synchronise, agree that everybody is done, and leave without finalisers
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def run_sweep(args, rank, world, local, dist):
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

        for i in range(max(args.warmup, 3)):
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
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "metric": "relit images/sec @256x256, 18-light sweep (one CNN pass per face)", "value": world * B * L * args.steps * 1e3 / ms,
            "unit": "images/s", "faces_per_s": world * B * args.steps * 1e3 / ms, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[4]: 18 Multi-PIE light directions x 8 faces per GPU per step (512 faces = 64 steps of one GPU "
                                   "or 8 steps of eight), CNN once per face, march + shade for 144 (face, light) pairs in one launch",
                       "global_batch": world * B, "lights": L, "parallelism": "dp%d (faces sharded, no collective)" % world,
                       "working_set": "144 relit images = 170 MB of outputs per step (> L2)", "cuda_graph": True},
            "gpu_launches": int(per_step * args.steps), "rendered_shape": list(out["rendered"].shape)}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-faces", type=int, default=6, help="faces in the bounded CPU-baseline sample")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-gpu-port", action="store_true", help="skip the informational torch-eager GPU run of the oracle port "
                                                               "(~50k tiny launches; always skip it under ncu)")
    ap.add_argument("--lanes", type=int, default=3, help="runner lanes the e2e (host-buffer) path rotates over")
    ap.add_argument("--no-gan", action="store_true", help="train workload without the PatchGAN terms")
    ap.add_argument("--workload", default="forward", choices=["forward", "train", "sweep"],
                    help="forward = configs[1] (the default bench line); train = configs[2]/[3]: generator training step, "
                         "B=16 per GPU, one flat-gradient all-reduce per step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.workload == "train":
        return run_train(args, rank, world, local, dist)
    if args.workload == "sweep":
        return run_sweep(args, rank, world, local, dist)

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
        total_ms = sum(a.elapsed_time(b) for a, b in pairs)
        if dist is not None:
            t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms

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
        total_ms = max(start.elapsed_time(e) for e in ends)
        if dist is not None:
            t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms

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
    e2e_ms = timed_lanes(lambda i: runner.relight_host(*host[i % len(host)]), args.steps, args.warmup) / args.steps
    clocks = sampler.stop()
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = B * 3 * H * W * 4

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
                 "source": "smsp__inst_executed.sum from profiles/r01_ncu_march_shade_full.csv"}

    line = {
        "metric": METRIC, "value": value, "unit": "faces/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch": world * B, "parallelism": "dp%d (faces sharded, no collective)" % world,
                   "l2": "value/e2e: every step reads a different batch of a 151 MB device pool (> 126 MB L2) / streams it from "
                         "the host, and rewrites ~1.1 GB of activations; latency + roofline: 256 MiB flush written between "
                         "timed launches (untimed)", "lanes": len(runner.lanes), "cnn": net.cnn_impl,
                   "cuda_graph": runner.graph is not None},
        "e2e": {"value": world * B * 1e3 / e2e_ms, "unit": "faces/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "lanes": len(runner.lanes),
                "note": "RelightRunner.relight_host: pinned host image/mask/light -> H2D -> forward -> D2H rendered, "
                        "steps pipelined over the runner lanes, one event pair around all K steps"},
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
    }
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        t = cpu_forward_faces_per_s(max(args.cpu_faces, 2), threads)[1:]
        line["cpu_baseline"] = {"value": len(t) / sum(t), "unit": "faces/s", "cores": threads, "kind": "port",
                                "sample": "%d faces (1 warm-up dropped), B=1 each, torch CPU oracle port of TEST1:169-505"
                                          % len(t)}
        try:
            if args.no_gpu_port:
                raise RuntimeError("skipped (--no-gpu-port)")
            tg = gpu_port_faces_per_s(8)
            line["reference_gpu_port"] = {"value": len(tg) / sum(tg), "unit": "faces/s", "kind": "port",
                                          "sample": "%d faces, B=1 each, the torch oracle port (TEST1:169-505) run eagerly on cuda:0" % len(tg)}
        except Exception as e:                      # informational only
            line["reference_gpu_port"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
