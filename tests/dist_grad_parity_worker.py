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
        rel = max(float((p - named_solo[k]).abs().max() / named_solo[k].abs().max().clamp_min(1e-30))
                  for k, p in model.named_parameters() if k in ref)
        assert rel < 1e-4, rel
        solo.flat.release()
    tr.flat.release()

# ensemble members sharded over the ranks + all_gather == the single-device loop, member by member
model = make()
one = to_dev(cg.CG_collate(samples[:1]))
n_ens = 8
eps_m = torch.randn(n_ens, ncg, F, generator=torch.Generator().manual_seed(11)).to(dev)
want = sample_single(model, one, n_ens, eps_m, reconstruct=False)[0]
got = sample_ensemble_sharded(model, one, n_ens, eps_m)
assert got.shape == want.shape and torch.equal(got, want), float((got - want).abs().max())
dist.barrier()
if rank == 0:
    print("DIST_GRAD_PARITY_OK world=%d worst_rel_err=%.2e modes=%s" % (world, worst_all, ",".join(modes)), flush=True)
dist.destroy_process_group()
