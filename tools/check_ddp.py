"""Data-parallel check, run under torchrun on >= 2 GPUs:
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp.py
(1) the gradient exchange of train.TrainStep with factor all-gather (all-reduce of the head of the flat buffer +
    all-gather of the deferred-gradient factors + grouped launch over the rows of all ranks) equals the all-reduced
    plain-autograd gradients; (2) after graphed steps every rank holds bit-identical parameters; (3) timing of both
    exchange modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import coarsegrainingvae_b200 as cg
from coarsegrainingvae_b200 import ops, synthetic
from coarsegrainingvae_b200.factory import build_cgvae
from coarsegrainingvae_b200.train import GraphedTrainStep, TrainStep, to_static_batch, training_loss

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
small = os.environ.get("CHECK_SMALL", "0") == "1"
cfg = dict(synthetic.CONFIGS["c2_chignolin"])
if small:
    cfg.update(n_basis=64, dec_nconv=3)
rad = lambda xyz, c: ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=dev), c).cpu().numpy()
raw = [synthetic.cgvae_batch(cfg, rank * 1000 + i, rad, cg.CG_collate) for i in range(2)]
B, n, ncg = cfg["batch"], cfg["n_atoms"], cfg["n_cgs"]
caps = {"nbr_list": B * n * (n - 1) // 2, "CG_nbr_list": B * ncg * (ncg - 1) // 2,
        "bond_edge_list": max(b["bond_edge_list"].shape[0] for b in raw) + 64}
cap_t = torch.tensor([caps["bond_edge_list"]], device=dev)
dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
caps["bond_edge_list"] = int(cap_t)
batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in to_static_batch(b, caps).items()} for b in raw]
eps = torch.randn(B * ncg, cfg["n_basis"], generator=torch.Generator().manual_seed(7)).to(dev)


def make():
    torch.manual_seed(123)
    return build_cgvae(cfg["n_basis"], cfg["n_rbf"], cfg["enc_nconv"], cfg["dec_nconv"], cfg["atom_cutoff"], cfg["cg_cutoff"],
                       cfg["n_cgs"]).to(dev)


# ---- (1) gradients
model = make()
out = model(batches[0], eps=eps)
training_loss(out, out[4], batches[0]["bond_edge_list"], cfg["beta"], cfg["gamma"], batches[0].get("bond_count"))[0].backward()
ref = {}
for k, p in model.named_parameters():
    if p.grad is not None:
        g = p.grad.detach().clone()
        dist.all_reduce(g)
        ref[k] = g
model2 = make()
tr = TrainStep(model2, cfg["beta"], cfg["gamma"], capturable=True)
tr.prepare(batches[0], eps)
assert tr.gather_factors, "factor exchange not active"
tr.forward_backward(batches[0], eps)
tr.exchange_gradients()
ops.wgrad_grouped_table(tr._factor_table, tr._factor_out_floats)
torch.cuda.synchronize()
worst = 0.0
for k, p in model2.named_parameters():
    if k in ref:
        err = float((p.grad - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30))
        worst = max(worst, err)
        assert err < 5e-6, (k, err)
head_mb = tr.n_reduce * 4 / 1e6
print("rank %d: exchanged gradients == all-reduced plain gradients, worst rel err %.2e; all-reduce %.1f MB + all-gather %.1f MB "
      "per rank instead of %.1f MB" % (rank, worst, head_mb, tr._arena.numel() * 4 / 1e6, tr.flat.flat.numel() * 4 / 1e6), flush=True)

# ---- (2) + (3) graphed steps in both modes
for mode in ("1", "0"):
    os.environ["CGVAE_GATHER_FACTORS"] = mode          # "1" forces the factor exchange at any world size
    m = make()
    t = TrainStep(m, cfg["beta"], cfg["gamma"], capturable=True)
    t.prepare(batches[0], eps)
    g = GraphedTrainStep(t, batches[0], eps)
    for i in range(5):
        g.step(batches[i % 2])
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        loss = g.step(batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    flat = t.flat_p
    chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(torch.equal(allc[0], c) for c in allc)
    if rank == 0:
        print("gather_factors=%s: %.3f ms per step (max over ranks), loss %.5f, parameters identical on all ranks: %s" %
              (mode, float(ms), float(loss), same), flush=True)
    assert same and torch.isfinite(loss)
    named = {k: p.detach().clone() for k, p in m.named_parameters()}      # by NAME: the two modes order the flat buffer differently
    if mode == "1":
        p_gather = named
    else:
        rel = max(float((named[k] - p_gather[k]).abs().max() / named[k].abs().max().clamp_min(1e-30)) for k in named)
        if rank == 0:
            print("parameters after 25 steps, factor exchange vs full all-reduce: max rel diff %.2e" % rel, flush=True)
        assert rel < 1e-3
    t.flat.release()
dist.barrier()
dist.destroy_process_group()
