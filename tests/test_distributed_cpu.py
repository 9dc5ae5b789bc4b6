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


def _train_worker(rank, world, port, out):
    """data-parallel TrainStep on the torch-optimiser path with the TEST-ONLY kernel emulator: ragged shards, global
    denominators (dp_norms), the loss slot riding in the all-reduced buffer, identical skip decision on every rank."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    import coarsegrainingvae_b200 as cg
    from coarsegrainingvae_b200 import ops, synthetic
    from coarsegrainingvae_b200.factory import build_cgvae
    from coarsegrainingvae_b200.train import TrainStep, training_loss
    from oracle import graph_oracle as gorc
    from tests import emulator
    for name in emulator._NAMES:
        setattr(ops, name, getattr(emulator, name))
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(n_basis=16, enc_nconv=1, dec_nconv=1, atom_cutoff=4.0)          # short cutoff: ragged edge / bond counts
    rad = lambda xyz, c: gorc.radius_graph(np.asarray(xyz, dtype=np.float32), c)
    samples = [synthetic.cgvae_sample(cfg, 77 + k, rad) for k in range(3)]     # 3 conformations over 2 ranks: 2 + 1
    mine = samples[rank::world]
    local = cg.CG_collate(mine)
    glob = cg.CG_collate(samples)
    F, ncg = cfg["n_basis"], cfg["n_cgs"]
    eps_all = torch.randn(3, ncg, F, generator=torch.Generator().manual_seed(3))
    eps_local, eps_glob = eps_all[rank::world].reshape(-1, F), eps_all.reshape(-1, F)

    def make():
        torch.manual_seed(11)
        return build_cgvae(F, cfg["n_rbf"], 1, 1, cfg["atom_cutoff"], cfg["cg_cutoff"], ncg)

    ref = make()
    o = ref(glob, eps=eps_glob)
    gl = training_loss(o, o[4], glob["bond_edge_list"], cfg["beta"], cfg["gamma"])[0]
    gl.backward()
    want = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}

    model = make()
    tr = TrainStep(model, cfg["beta"], cfg["gamma"], lr=1e-3, optimizer="torch")
    lb = tr.update_dp_norms(dict(local))
    counts = lb["dp_norms"] * world
    assert int(counts[0]) == glob["nxyz"].shape[0] and int(counts[2]) == glob["bond_edge_list"].shape[0]
    tr.prepare(lb, eps_local)
    tr.forward_backward(lb, eps_local)
    tr.exchange_gradients()                                                 # torch path: mean over ranks applied here
    worst = 0.0
    for k, p in model.named_parameters():
        if k in want:
            worst = max(worst, float((p.grad - want[k]).abs().max() / want[k].abs().max().clamp_min(1e-30)))
    loss_mean = float(tr.flat.loss_slot)                                    # all-reduced: mean over ranks of the local losses
    # with global denominators the mean over ranks of the local losses IS the single-process loss of the global batch
    out.put((rank, worst, loss_mean, float(gl)))
    tr.loss_limit = loss_mean * 0.5                                         # every rank must skip (same all-reduced value)
    before = [p.detach().clone() for p in tr.flat.params]
    tr.apply_gradients()
    assert all(torch.equal(a, p) for a, p in zip(before, tr.flat.params)) and tr.skipped_steps() == 1
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_train_step_matches_single_process_on_ragged_shards():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([out.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for rank, worst, loss_mean, loss_global in results:
        assert worst < 2e-5, (rank, worst)                                  # fp32 emulator, different summation order
        assert abs(loss_mean - loss_global) <= 1e-5 * abs(loss_global), (rank, loss_mean, loss_global)
    assert results[0][2] == results[1][2]
