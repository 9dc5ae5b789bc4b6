"""Worker of tests/test_gpu_multi.py, launched under torchrun on N GPUs.

SURVEY.md section 8(e) gate: the data-parallel gradients on N GPUs (rank r holds conformations r::N of the global batch,
ONE all-reduce of the flat gradient buffer, 1/N inside the fused optimiser) equal the single-GPU gradients of the same
GLOBAL batch within 1e-5 -- also for the factor-exchange mode (2 ranks) -- and after optimiser steps every rank holds
bit-identical parameters.  Prints one line `DIST_GRAD_PARITY_OK world=N ...` on rank 0.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import coarsegrainingvae_b200 as cg
from coarsegrainingvae_b200 import ops, synthetic
from coarsegrainingvae_b200.factory import build_cgvae
from coarsegrainingvae_b200.train import TrainStep, sample_ensemble_sharded, sample_single, training_loss

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
per_rank = 2
cfg.update(batch=per_rank * world, n_basis=64, enc_nconv=2, dec_nconv=2, atom_cutoff=4.5)   # short cutoff: ragged edge / bond counts
rad = lambda xyz, c: ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=dev), c).cpu().numpy()
samples = [synthetic.cgvae_sample(cfg, 4321 + k, rad) for k in range(cfg["batch"])]
to_dev = lambda b: {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
global_batch = to_dev(cg.CG_collate(samples))
local_batch = to_dev(cg.CG_collate(samples[rank::world]))
F, ncg = cfg["n_basis"], cfg["n_cgs"]
eps_all = torch.randn(cfg["batch"], ncg, F, generator=torch.Generator().manual_seed(7)).to(dev)
eps_global = eps_all.reshape(-1, F)
eps_local = eps_all[rank::world].reshape(-1, F)


def make():
    torch.manual_seed(123)
    return build_cgvae(F, cfg["n_rbf"], 2, 2, cfg["atom_cutoff"], cfg["cg_cutoff"], ncg).to(dev)


# single-GPU gradients of the GLOBAL batch (every rank computes them itself)
ref_model = make()
out = ref_model(global_batch, eps=eps_global)
training_loss(out, out[4], global_batch["bond_edge_list"], cfg["beta"], cfg["gamma"])[0].backward()
ref = {k: p.grad.detach().clone() for k, p in ref_model.named_parameters() if p.grad is not None}

worst_all = 0.0
modes = ["0", "1"] if world == 2 else ["0"]
os.environ["CGVAE_SHARD_OPT"] = "0"          # this section reads the fully reduced gradient buffer on every rank
for mode in modes:
    os.environ["CGVAE_GATHER_FACTORS"] = mode
    model = make()
    tr = TrainStep(model, cfg["beta"], cfg["gamma"], lr=1e-3)
    lb = tr.update_dp_norms(dict(local_batch))
    tr.prepare(lb, eps_local)
    tr.forward_backward(lb, eps_local)
    tr.exchange_gradients()
    if tr.gather_factors:
        ops.wgrad_grouped_table(tr._factor_table, tr._factor_out_floats)
    torch.cuda.synchronize()
    worst = 0.0
    for k, p in model.named_parameters():
        if k in ref:
            got = p.grad / world                                  # the 1/world factor lives in the fused optimiser
            err = float((got - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30))
            worst = max(worst, err)
            assert err < 1e-5, (mode, k, err)
    worst_all = max(worst_all, worst)
    # optimiser steps: identical parameters on every rank
    for _ in range(3):
        tr.step(lb, eps_local)
    chk = torch.stack([tr.flat_p.double().sum(), tr.flat_p.double().abs().sum()])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    assert all(torch.equal(allc[0], c) for c in allc), "parameters diverged across ranks"
    # ... and equal to single-GPU steps on the global batch
    if mode == "0":
        solo = TrainStep(ref_model, cfg["beta"], cfg["gamma"], lr=1e-3, group=None)
        solo._world = lambda: 1
        solo.prepare(global_batch, eps_global)
        for _ in range(3):
            solo.forward_backward(global_batch, eps_global)
            solo.apply_gradients()
        named_solo = dict(ref_model.named_parameters())
        # Adam divides by sqrt(v): an entry whose gradient sits at rounding-noise level moves by up to lr per step in EITHER
        # run, so the gate is the relative L2 distance of the whole parameter vector plus a bound on the largest single
        # deviation as a fraction of the largest possible movement (3 steps x lr); the 1e-5 gradient gate above is the
        # parity statement proper
        num = sum(float((p.detach() - named_solo[k].detach()).double().pow(2).sum()) for k, p in model.named_parameters() if k in ref)
        den = sum(float(named_solo[k].detach().double().pow(2).sum()) for k in ref)
        dmax = max(float((p.detach() - named_solo[k].detach()).abs().max()) for k, p in model.named_parameters() if k in ref)
        assert (num / den) ** 0.5 < 1e-4, (num / den) ** 0.5
        assert dmax < 0.1 * 3 * 1e-3, dmax
        solo.flat.release()
    tr.flat.release()

# the whole step as CUDA graph replays: overlapped exchange (decoder block all-reduced on the communication stream while the
# encoder's backward runs, NCCL captured inside the graph) == the two-graph form with one eager all-reduce in between
from coarsegrainingvae_b200.train import GraphedTrainStep, to_static_batch
cpu_local = cg.CG_collate(samples[rank::world])
caps = {"nbr_list": cpu_local["nbr_list"].shape[0] + 64, "CG_nbr_list": cpu_local["CG_nbr_list"].shape[0] + 8,
        "bond_edge_list": cpu_local["bond_edge_list"].shape[0] + 16}
cap_t = torch.tensor([caps["nbr_list"], caps["CG_nbr_list"], caps["bond_edge_list"]], device=dev)
dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
caps = dict(zip(("nbr_list", "CG_nbr_list", "bond_edge_list"), [int(x) for x in cap_t.tolist()]))
static_local = to_dev(to_static_batch(cpu_local, caps))
finals = {}
for overlap in ("0", "1", "shard"):
    # "0": one all-reduce between two graphs; "1": overlapped decoder block, NCCL captured; "shard": reduce-scatter ->
    # Adam on 1/world of the buffers -> all-gather of the parameters (the default)
    os.environ["CGVAE_DP_OVERLAP"] = "1" if overlap == "1" else "0"
    os.environ["CGVAE_SHARD_OPT"] = "1" if overlap == "shard" else "0"
    os.environ["CGVAE_GATHER_FACTORS"] = "0"
    model = make()
    init = {k: p.detach().clone() for k, p in model.named_parameters()}
    tr = TrainStep(model, cfg["beta"], cfg["gamma"], lr=1e-3, capturable=True)
    sb = tr.update_dp_norms(dict(static_local))
    tr.prepare(sb, eps_local)
    assert (tr.n_overlap > 0) == (overlap == "1") and tr.shard_opt == (overlap == "shard")
    graphed = GraphedTrainStep(tr, sb, eps_local)
    with torch.no_grad():                                   # undo the warm-up steps of the capture
        for k, p in model.named_parameters():
            p.copy_(init[k])
        tr.exp_avg.zero_(); tr.exp_avg_sq.zero_(); tr.step_count.zero_()
        if tr.flat_p.numel() > tr.flat.flat.numel():
            tr.flat_p[tr.flat.flat.numel():].zero_()
    losses = [float(graphed.step(static_local)) for _ in range(3)]
    torch.cuda.synchronize()
    finals[overlap] = ({k: p.detach().clone() for k, p in model.named_parameters()}, losses)
    graphed = None
    tr.flat.release()
pa, la = finals["0"]
for other in ("1", "shard"):
    pb, lb_ = finals[other]
    num = sum(float((pa[k] - pb[k]).double().pow(2).sum()) for k in pa)
    den = sum(float(pa[k].double().pow(2).sum()) for k in pa)
    assert (num / den) ** 0.5 < 1e-5, (other, "vs plain exchange", (num / den) ** 0.5)
    assert max(abs(x - y) for x, y in zip(la, lb_)) <= 1e-5 * max(abs(x) for x in la), (other, la, lb_)
    # every rank holds the same parameters
    chk = torch.stack([sum(p.double().sum() for p in pb.values()), sum(p.double().abs().sum() for p in pb.values())])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    assert all(torch.equal(allc[0], c) for c in allc), (other, "parameters diverged across ranks")
os.environ.pop("CGVAE_DP_OVERLAP"); os.environ.pop("CGVAE_SHARD_OPT")

# ensemble members sharded over the ranks + all_gather == the single-device loop, member by member
model = make()
one = to_dev(cg.CG_collate(samples[:1]))
n_ens = 8
eps_m = torch.randn(n_ens, ncg, F, generator=torch.Generator().manual_seed(11)).to(dev)
want = sample_single(model, one, n_ens, eps_m, reconstruct=False)[0]
got = sample_ensemble_sharded(model, one, n_ens, eps_m)
assert got.shape == want.shape and torch.equal(got, want), float((got - want).abs().max())
# ... and as one CUDA graph per rank (ShardedGraphedSampler) over a static-capacity batch
from coarsegrainingvae_b200.train import ShardedGraphedSampler
cpu_one = cg.CG_collate(samples[:1])
caps1 = {"nbr_list": 22 * 21 // 2, "CG_nbr_list": ncg * (ncg - 1) // 2 + 1, "bond_edge_list": cpu_one["bond_edge_list"].shape[0] + 4}
static_one = to_dev(to_static_batch(cpu_one, caps1))
ss = ShardedGraphedSampler(model, static_one, n_ens)
got_g = ss.sample(static_one, eps_m)
assert got_g.shape == want.shape and float((got_g - want).abs().max() / want.abs().max()) < 1e-6
ss = None
dist.barrier()
if rank == 0:
    print("DIST_GRAD_PARITY_OK world=%d worst_rel_err=%.2e modes=%s" % (world, worst_all, ",".join(modes)), flush=True)
dist.destroy_process_group()
