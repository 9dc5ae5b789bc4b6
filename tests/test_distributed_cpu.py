"""CPU, gloo, world_size 2: the data-parallel plumbing (flat gradient bucket, one all-reduce, identical clip on every
rank) and rank-sharding of conformations / ensemble members.  No kernels involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from coarsegrainingvae_b200.train import FlatGrads, used_parameters


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # identical replicas
    net = nn.ModuleDict({"used": nn.Linear(6, 4), "also": nn.Linear(4, 1), "never": nn.Linear(3, 3)})
    data = torch.arange(24, dtype=torch.float32).reshape(4, 6) / 10.0     # global batch of 4 "conformations"
    shard = data[rank::world]                              # rank r takes conformations r::world

    def run():
        net["also"](torch.tanh(net["used"](shard))).sum().backward()

    used = used_parameters(net, run)
    assert [k for k, _ in used] == ["used.weight", "used.bias", "also.weight", "also.bias"]
    flat = FlatGrads([p for _, p in used])
    assert net["never"].weight.grad is None
    flat.zero_()
    run()
    assert net["used"].weight.grad.data_ptr() == flat.flat.data_ptr()     # grads live inside the bucket
    flat.allreduce_mean_()
    norm = flat.clip_(0.01)
    out.put((rank, flat.flat.clone(), float(norm)))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([out.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process result on the same global batch: sum over the batch -> mean over ranks == (1/world) * total
    torch.manual_seed(0)
    net = nn.ModuleDict({"used": nn.Linear(6, 4), "also": nn.Linear(4, 1), "never": nn.Linear(3, 3)})
    data = torch.arange(24, dtype=torch.float32).reshape(4, 6) / 10.0
    net["also"](torch.tanh(net["used"](data))).sum().backward()
    want = torch.cat([p.grad.reshape(-1) for k, p in net.named_parameters() if p.grad is not None]) / world
    norm = torch.linalg.vector_norm(want)
    want = want * torch.clamp(0.01 / (norm + 1e-6), max=1.0)
    for rank, flat, n in results:
        assert torch.allclose(flat, want, rtol=1e-6, atol=1e-9), rank
        assert abs(n - float(norm)) < 1e-6 * max(1.0, float(norm))
    assert torch.equal(results[0][1], results[1][1])       # every rank holds identical clipped gradients
