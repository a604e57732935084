"""CPU, world_size 2, gloo: the host logic of the data-parallel path — every rank gets a disjoint shard of
the faces (no data-path collective in the forward), and training gradients live in ONE flat buffer that is summed with
a single all_reduce per optimiser step and scaled by 1/world inside the Adam kernel."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geomconsistentfr_b200 import FlatAdam
    from geomconsistentfr_b200.runner import shard_faces
    torch.manual_seed(0)                                  # identical parameters on every replica
    params = [torch.randn(4, 3).requires_grad_(), torch.randn(5).requires_grad_()]
    opt = FlatAdam(params, lr=1e-4)
    assert params[0].data_ptr() == opt.flat.data_ptr() and params[1].grad.data_ptr() == opt.grad[12:].data_ptr()
    opt.zero_grad()
    params[0].grad += float(rank + 1)                     # per-replica gradients
    params[1].grad += 10.0 * float(rank + 1)
    scale = opt.all_reduce_grads()
    assert abs(scale - 1.0 / world) < 1e-12
    assert torch.allclose(opt.grad[:12], torch.full((12,), 3.0)) and torch.allclose(opt.grad[12:], torch.full((5,), 30.0))
    lo, hi = shard_faces(10, rank, world)
    got = [None] * world
    dist.all_gather_object(got, (lo, hi))
    if rank == 0:
        out.put(got)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_and_face_sharding():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    shards = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert shards == [(0, 5), (5, 10)]
