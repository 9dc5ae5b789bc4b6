"""GPU (B200), round 2: the rows the round-1 review found unpinned -- ensemble sampling against the REAL reference's
loop, ``CGDataset.generate_neighbor_list``, the skip guard / validation branch of the training loop, index validation
(the reference's IndexError at the lifting gather), bidirectional edge lists in static (CUDA-graph) mode -- and the
tensor-core message kernels."""
import copy

import numpy as np
import pytest
import torch

import coarsegrainingvae_b200 as cg
from coarsegrainingvae_b200 import ops, synthetic
from oracle import cgvae_oracle as orc
from oracle import graph_oracle as gorc
from tests import parity_cases as pc
from tests.golden_util import load, rel_err, section

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _to(batch, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}


def _gpu_radius(xyz, cutoff):
    return ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=DEV), cutoff).cpu().numpy()


# ------------------------------------------------------------------------------------------ sampling (a25)

@pytest.mark.parametrize("tag", ["sym", "nosym"])
def test_sampling_matches_reference_loop(tag):
    """train.sample_single and train.GraphedSampler reproduce, member by member and for the same noise, the geometries
    the REAL reference produces through the loop of scripts/sampling.py:265-293 (tests/golden/sampling_small.npz)."""
    pc.sampling_case(DEV, tag, torch.float32, TOL, gold_tol=5e-5, graphed=True)


# ------------------------------------------------------------------------------------------ dataset graph builder (a2)

@pytest.mark.parametrize("tag,cg_cut,und", [("radius", 4.0, True), ("radius_dir", 4.0, False), ("bond", None, True)])
def test_generate_neighbor_list_matches_reference(tag, cg_cut, und):
    """CGDataset.generate_neighbor_list (one batched kernel call per graph kind, split back with searchsorted) == the real
    reference's per-frame python loop (data.py:207-252) on ragged frames, bit for bit; ``cg_cutoff=None``: bond graph."""
    z = load("dataset_lists.npz")
    sizes = [int(x) for x in z["meta/sizes"]]
    keys = ("nxyz", "CG_nxyz", "CG_mapping", "num_atoms", "num_CGs", "bond_edge_list")
    props = {k: [torch.from_numpy(np.asarray(z["props/%d/%s" % (i, k)])) for i in range(len(sizes))] for k in keys}
    ds = cg.CGDataset(props)
    ds.generate_neighbor_list(atom_cutoff=float(z["meta/atom_cutoff"]), cg_cutoff=cg_cut, device=DEV, undirected=und)
    for i in range(len(sizes)):
        for k in ("nbr_list", "CG_nbr_list"):
            got = ds.props[k][i]
            want = torch.from_numpy(np.asarray(z["%s/%d/%s" % (tag, i, k)]))
            assert got.dtype == torch.int64 and got.device.type == "cpu"
            assert got.shape == want.shape and torch.equal(got, want), (tag, i, k)


def test_generate_neighbor_list_with_empty_frames():
    g = torch.Generator().manual_seed(2)
    frames = [torch.cat([torch.ones(n, 1), torch.randn(n, 3, generator=g) * s], 1) for n, s in ((6, 1.0), (1, 1.0), (9, 40.0), (4, 0.5))]
    props = {"nxyz": frames, "CG_nxyz": [f[: max(1, f.shape[0] // 2)] for f in frames]}
    ds = cg.CGDataset(props)
    ds.generate_neighbor_list(atom_cutoff=2.0, cg_cutoff=2.0, device=DEV)
    for k, src in (("nbr_list", props["nxyz"]), ("CG_nbr_list", props["CG_nxyz"])):
        for f, got in zip(src, ds.props[k]):
            want = gorc.radius_graph(f[:, 1:4].numpy(), 2.0, True)
            assert np.array_equal(got.numpy(), want)


# ------------------------------------------------------------------------------------------ loop control flow (a24)

def _small_trainer(lr=1e-3, **kw):
    from coarsegrainingvae_b200.factory import build_cgvae
    from coarsegrainingvae_b200.train import TrainStep
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(batch=3, n_basis=64, enc_nconv=2, dec_nconv=2)
    raw = synthetic.cgvae_batch(cfg, 0, _gpu_radius, cg.CG_collate)
    torch.manual_seed(5)
    model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], 2, 2, cfg["atom_cutoff"], cfg["cg_cutoff"], 3).to(DEV)
    eps = torch.randn(9, 64, device=DEV)
    tr = TrainStep(model, cfg["beta"], cfg["gamma"], lr=lr, **kw)
    return cfg, raw, model, eps, tr


def test_skip_guard_and_validation_mode_match_reference_loop():
    """scripts/utils.py:145-160 on the device: a loss >= 200*gamma or NaN leaves parameters, moments and the step
    counter untouched (and is counted); the validation branch runs backward without an optimiser step."""
    cfg, raw, model, eps, tr = _small_trainer()
    batch = _to(raw, DEV)
    tr.prepare(batch, eps)
    snap = lambda: (tr.flat_p.clone(), tr.exp_avg.clone(), tr.exp_avg_sq.clone(), float(tr.step_count))
    p0 = snap()
    loss = tr.step(batch, eps, train=False)                     # validation: loss.backward() only
    assert orc.train_loop_step(loss, cfg["gamma"], train=False) == (True, False)
    p1 = snap()
    assert all(torch.equal(a, b) for a, b in zip(p0[:3], p1[:3])) and p1[3] == p0[3]
    assert float(tr.flat.flat.abs().max()) > 0                  # gradients were produced
    tr.loss_limit = float(loss) * 0.5                           # past the guard: skipped on the device
    tr.step(batch, eps)
    p2 = snap()
    assert all(torch.equal(a, b) for a, b in zip(p0[:3], p2[:3])) and p2[3] == p0[3] and tr.skipped_steps() == 1
    tr.loss_limit = cfg["gamma"] * 200.0                        # the reference's limit: the step is taken
    loss = tr.step(batch, eps)
    assert orc.train_loop_step(loss, cfg["gamma"]) == (True, True)
    p3 = snap()
    assert not torch.equal(p3[0], p0[0]) and p3[3] == p0[3] + 1 and tr.skipped_steps() == 1
    # NaN loss (poisoned input coordinates): no update, parameters stay finite
    bad = dict(batch)
    bad["nxyz"] = batch["nxyz"].clone()
    bad["nxyz"][0, 1] = float("nan")
    lnan = tr.step(bad, eps)
    assert orc.train_loop_step(lnan, cfg["gamma"]) == (False, False)
    p4 = snap()
    assert all(torch.equal(a, b) for a, b in zip(p3[:3], p4[:3])) and p4[3] == p3[3] and tr.skipped_steps() == 2
    assert bool(torch.isfinite(tr.flat_p).all())
    tr.flat.release()


def test_skip_guard_inside_cuda_graph_replay():
    """the guard is data-dependent but takes no host read: the SAME captured graph skips a poisoned batch and steps on a
    good one."""
    from coarsegrainingvae_b200.train import GraphedTrainStep, to_static_batch
    cfg, raw, model, eps, tr = _small_trainer(capturable=True)
    caps = {"nbr_list": 3 * 22 * 21 // 2, "CG_nbr_list": 9, "bond_edge_list": raw["bond_edge_list"].shape[0] + 8}
    good = _to(to_static_batch(raw, caps), DEV)
    tr.prepare(good, eps)
    g = GraphedTrainStep(tr, good, eps)
    bad = dict(good)
    bad["nxyz"] = good["nxyz"].clone()
    bad["nxyz"][3, 2] = float("nan")
    before, steps = tr.flat_p.clone(), float(tr.step_count)
    g.step(bad)
    assert torch.equal(tr.flat_p, before) and float(tr.step_count) == steps and tr.skipped_steps() == 1
    g.step(good)
    assert not torch.equal(tr.flat_p, before) and float(tr.step_count) == steps + 1 and tr.skipped_steps() == 1
    tr.flat.release()


def test_fused_adam_nonfinite_norm_is_a_noop():
    n = 4096 + 3
    g = torch.Generator().manual_seed(0)
    p, grad = torch.randn(n, generator=g).to(DEV), torch.randn(n, generator=g).to(DEV)
    m, v, step, skipped = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    grad[17] = float("inf")
    p0 = p.clone()
    ops.adam_clip_step(p, grad, m, v, step, 0.01, 1e-3, skipped=skipped)
    assert torch.equal(p, p0) and float(m.abs().max()) == 0 and float(v.abs().max()) == 0
    assert float(step) == 0 and float(skipped) == 1
    grad[17] = 0.5
    ops.adam_clip_step(p, grad, m, v, step, 0.01, 1e-3, skipped=skipped)
    assert not torch.equal(p, p0) and float(step) == 1 and float(skipped) == 1


# ------------------------------------------------------------------------------------------ index validation

def test_bead_with_more_atoms_than_channels_raises_like_the_reference():
    """cg_v[mapping, CG2atomChannel] raises IndexError in the reference when a bead holds more atoms than channels
    (cgvae.py:473); here the lifting kernels never touch memory out of range and the forward raises the same error."""
    from coarsegrainingvae_b200.factory import build_cgvae
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(batch=2, n_basis=8, n_rbf=4, enc_nconv=1, dec_nconv=1)       # 8 atoms in bead 0, F = 8: just fits
    torch.manual_seed(1)
    model = build_cgvae(8, 4, 1, 1, cfg["atom_cutoff"], cfg["cg_cutoff"], 3).to(DEV)
    batch = _to(synthetic.cgvae_batch(cfg, 0, _gpu_radius, cg.CG_collate), DEV)
    out = model(batch)
    assert bool(torch.isfinite(out[5]).all())
    bad = dict(batch)
    bad["CG_mapping"] = batch["CG_mapping"].clone()
    bad["CG_mapping"][8] = 0                                                # ninth atom of bead 0
    with pytest.raises(IndexError):
        model(bad)
    bad = dict(batch)
    bad["nxyz"] = batch["nxyz"].clone()
    bad["nxyz"][0, 0] = 250.0                                               # atomic number outside nn.Embedding(100, F)
    with pytest.raises(IndexError):
        model(bad)
    out2 = model(batch)                                                     # the flag is cleared: good batches still run
    assert torch.equal(out2[5], out[5])


# ------------------------------------------------------------------------------------------ static mode, both-direction lists

def test_graphed_step_with_bond_derived_cg_graph_matches_eager():
    """a CG list that already holds both directions (cg_cutoff=None: data.bond_cg_graph) must not be doubled by the
    static-capacity CSR builder: graph replay == eager on the same batches."""
    from coarsegrainingvae_b200.data import bond_cg_graph
    from coarsegrainingvae_b200.factory import build_cgvae
    from coarsegrainingvae_b200.train import GraphedTrainStep, TrainStep, to_static_batch
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(batch=3, n_basis=64, enc_nconv=2, dec_nconv=2)
    raws = []
    for i in range(3):
        samples = [synthetic.cgvae_sample(cfg, 900 + 10 * i + k, _gpu_radius) for k in range(cfg["batch"])]
        for s_ in samples:
            s_["CG_nbr_list"] = bond_cg_graph(s_["bond_edge_list"], s_["CG_mapping"], 3)
            assert s_["CG_nbr_list"].shape[0] > 0
        raws.append(cg.CG_collate(samples))
    caps = {"nbr_list": 3 * 22 * 21 // 2, "CG_nbr_list": 3 * 6, "bond_edge_list": max(b["bond_edge_list"].shape[0] for b in raws) + 8}
    static = [to_static_batch(b, caps) for b in raws]
    assert static[0]["CG_nbr_symmetrize"] is False and static[0]["nbr_symmetrize"] is True
    static = [_to(b, DEV) for b in static]
    torch.manual_seed(5)
    model_a = build_cgvae(64, cfg["n_rbf"], 2, 2, cfg["atom_cutoff"], cfg["cg_cutoff"], 3).to(DEV)
    model_b = copy.deepcopy(model_a)
    eps = torch.randn(9, 64, device=DEV)
    eager = TrainStep(model_a, cfg["beta"], cfg["gamma"], lr=1e-3)
    eager.prepare(_to(raws[0], DEV), eps)
    gtr = TrainStep(model_b, cfg["beta"], cfg["gamma"], lr=1e-3, capturable=True)
    gtr.prepare(static[0], eps)
    graphed = GraphedTrainStep(gtr, static[0], eps)
    for _ in range(3):
        eager.step(_to(raws[0], DEV), eps)
    for i in (1, 2, 0):
        la = eager.step(_to(raws[i], DEV), eps)
        lb = graphed.step(static[i])
        assert rel_err(lb, la) < 1e-5, (i, float(la), float(lb))
    gtr.flat.release()
    eager.flat.release()
