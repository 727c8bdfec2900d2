"""CPU, world_size 2 over gloo: the data-parallel reduction logic of nrw/train.py - every .grad is a view of ONE flat
buffer, a single all-reduce averages all of them, per-rank loss normalisers stay per-rank; TrainSystem.reduce_grads itself on the same
buffers; disjoint ray-cache shards per rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nrw import _lib

    table, total = _lib.param_table(100, 48)
    flat = torch.zeros(total)
    flat_grad = torch.zeros(total)
    params = []
    for name, shape, off, numel in table:
        p = torch.nn.Parameter(flat[off:off + numel].view(shape))
        p.grad = flat_grad[off:off + numel].view(shape)          # what TrainSystem.training_step installs
        params.append((name, p, off, numel))
    # rank-dependent "gradients"
    for i, (name, p, off, numel) in enumerate(params):
        p.grad.fill_(float(rank + 1) * (i + 1))
    dist.all_reduce(flat_grad)                                      # ONE collective for all parameters
    flat_grad.div_(world)
    ok = True
    for i, (name, p, off, numel) in enumerate(params):
        want = (i + 1) * sum(r + 1 for r in range(world)) / world
        ok &= bool(torch.all(p.grad == want)) and p.grad.data_ptr() == flat_grad.data_ptr() + off * 4
    # padding between tensors is reduced too and must stay zero (identical layouts on every rank)
    used = torch.zeros(total, dtype=torch.bool)
    for _, _, off, numel in table:
        used[off:off + numel] = True
    ok &= bool(torch.all(flat_grad[~used] == 0))
    # the product's own reduction (nrw/train.py::TrainSystem.reduce_grads, the world_size > 1 branch of training_step) on the
    # same buffers: DDP mean of the flat gradient and of the dense embedding gradient, in place
    import types

    from nrw.raycache import local_split
    from nrw.train import TrainSystem

    for i, (name, p, off, numel) in enumerate(params):
        p.grad.fill_(float(rank + 1) * (i + 1))
    emb_grad = torch.full((100, 48), float(10 * (rank + 1)))
    ptr = flat_grad.data_ptr()
    TrainSystem.reduce_grads(types.SimpleNamespace(world_size=world), flat_grad, emb_grad)
    ok &= flat_grad.data_ptr() == ptr and bool(torch.all(emb_grad == 10.0 * sum(r + 1 for r in range(world)) / world))
    for i, (name, p, off, numel) in enumerate(params):
        ok &= bool(torch.all(p.grad == (i + 1) * sum(r + 1 for r in range(world)) / world))
    # ray-cache shards: every rank derives the same permutation and takes a disjoint slice (datasets/data.py:83-100)
    mine = local_split([f"split_{i}" for i in range(64)], world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    flat_names = [n for part in gathered for n in part]
    ok &= len(flat_names) == 64 and len(set(flat_names)) == 64 and all(len(part) == 64 // world for part in gathered)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
