"""CPU: the real Python modules / autograd nodes of coarsegrainingvae_b200, run on top of the TEST-ONLY
kernel emulator (tests/emulator.py), against the oracle's autograd (cases in tests/parity_cases.py).

float64 runs pin the hand-derived backward formulas the CUDA kernels implement (tolerance 1e-10);
float32 runs against the frozen reference outputs (tests/golden) pin layouts, residual fusion,
state_dict compatibility.  The kernels themselves are checked on the GPU (test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

import coarsegrainingvae_b200 as cg
from tests import emulator, parity_cases as pc
from tests.golden_util import load, section

DTYPES = [(torch.float64, 1e-10), (torch.float32, 2e-5)]


@pytest.fixture(autouse=True)
def _emulated(monkeypatch):
    emulator.install(monkeypatch)


@pytest.mark.parametrize("tag,cls", [("k3", "EquiMessageBlock"), ("k4", "EquiMessageCross")])
@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_message_block_api(tag, cls, dtype, tol):
    pc.message_block_api("cpu", tag, cls, dtype, tol)


def test_message_block_edge_weight():
    pc.message_block_edge_weight("cpu")


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_pseudo_block_api(dtype, tol):
    pc.pseudo_block_api("cpu", dtype, tol)


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_update_and_contraction_api(dtype, tol):
    pc.update_block_api("cpu", dtype, tol)
    pc.contraction_block_api("cpu", dtype, tol)


@pytest.mark.parametrize("act", pc.ACTIVATIONS)
def test_blocks_with_every_registry_activation(act):
    """activation plumbing of the autograd nodes (forward code + derivative code) in float64 against the oracle"""
    pc.blocks_with_activation("cpu", act, torch.float64, 1e-10)


@pytest.mark.parametrize("tag", ["vae_sym", "vae_nosym", "vae_noneq"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 5e-5)])
def test_cgvae_model(tag, dtype, tol):
    pc.cgvae_model("cpu", tag, dtype, tol)


@pytest.mark.parametrize("tag", ["pcn_cross", "pcn_plain"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 5e-5)])
def test_pcn_model(tag, dtype, tol):
    pc.pcn_model("cpu", tag, dtype, tol)


@pytest.mark.parametrize("tag", ["sym", "nosym"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 5e-5)])
def test_sampling_loop(tag, dtype, tol):
    """train.sample_single on the emulator vs the real reference's sampling loop (frozen) and the oracle"""
    pc.sampling_case("cpu", tag, dtype, tol, gold_tol=5e-5 if dtype == torch.float32 else 5e-6)


def test_static_batch_keeps_bidirectional_lists_single():
    """to_static_batch takes make_directed's two host reads (conv.py:12-13): a list that already holds both directions (the
    bond-derived CG graph of cg_cutoff=None, or dir_mp=True) must not get the flipped half appended in static mode."""
    from coarsegrainingvae_b200.train import to_static_batch, validate_batch
    from coarsegrainingvae_b200.cgvae import BatchGraphs
    one_way = torch.tensor([[0, 1], [0, 2], [1, 2]])
    both = torch.cat([one_way, one_way.flip(1)])
    base = {"nbr_list": one_way, "CG_nbr_list": both, "bond_edge_list": one_way}
    sb = to_static_batch(base, {"nbr_list": 8, "CG_nbr_list": 8, "bond_edge_list": 8})
    assert sb["nbr_symmetrize"] is True and sb["CG_nbr_symmetrize"] is False
    g = BatchGraphs.for_batch(sb)
    atom = g.directed_graph("atom", sb["nbr_list"], 3)
    cgg = g.directed_graph("cg", sb["CG_nbr_list"], 3)
    assert int(atom.rowptr[-1]) == 6 and int(cgg.rowptr[-1]) == 6          # 3 undirected edges -> 6 directed, both ways
    assert torch.equal(cgg.rowptr, atom.rowptr)
    # dir_mp=True: already_directed wins even for a one-directional list
    assert int(g.directed_graph("atom", sb["nbr_list"], 3, already_directed=True).rowptr[-1]) == 3
    full = {"nxyz": torch.tensor([[1., 0, 0, 0], [6., 1, 0, 0], [8., 0, 1, 0]]), "CG_nxyz": torch.zeros(3, 4),
            "CG_mapping": torch.tensor([0, 0, 1]), **sb}
    validate_batch(full, n_basis=2)
    with pytest.raises(IndexError):
        validate_batch(full, n_basis=1)                 # bead 0 holds 2 atoms > 1 channel (cgvae.py:473)
    bad = dict(full, CG_mapping=torch.tensor([0, 0, 3]))
    with pytest.raises(IndexError):
        validate_batch(bad, n_basis=8)
    bad = dict(full, nxyz=torch.tensor([[1., 0, 0, 0], [6., 1, 0, 0], [108., 0, 1, 0]]))
    with pytest.raises(IndexError):
        validate_batch(bad, n_basis=8)


def test_train_step_skip_guard_and_validation_mode_cpu():
    """TrainStep on the torch-optimiser path (CPU): the reference loop's control flow (scripts/utils.py:145-160) -- no
    optimiser step for loss >= 200*gamma or NaN, validation mode runs backward without a step."""
    from coarsegrainingvae_b200.train import TrainStep
    from oracle import cgvae_oracle as orc
    z = load("cgvae_small.npz")
    sec = section(z, "vae_sym")
    F, R, enc, dec, acut, ccut, breaksym, beta, gamma = [float(x) for x in sec["meta"]]
    model = pc.build_vae(int(F), int(R), int(enc), int(dec), acut, ccut, bool(breaksym))
    pc._load_params(model, sec, torch.float32, "cpu")
    batch = {k[len("batch/"):]: v for k, v in sec.items() if k.startswith("batch/")}
    eps = sec["eps"]
    tr = TrainStep(model, beta, gamma, lr=1e-3, optimizer="torch")
    tr.prepare(batch, eps)
    before = [p.detach().clone() for p in tr.flat.params]
    loss = tr.step(batch, eps, train=False)                      # validation: backward, no step
    assert orc.train_loop_step(loss, gamma, train=False) == (True, False)
    assert all(torch.equal(a, p) for a, p in zip(before, tr.flat.params)) and float(tr.flat.flat.abs().max()) > 0
    tr.loss_limit = float(loss) * 0.5                            # pretend the loss is past the guard
    tr.step(batch, eps)
    assert all(torch.equal(a, p) for a, p in zip(before, tr.flat.params)) and tr.skipped_steps() == 1
    tr.loss_limit = gamma * 200.0
    loss = tr.step(batch, eps)
    assert orc.train_loop_step(loss, gamma) == (True, True)
    assert any(not torch.equal(a, p) for a, p in zip(before, tr.flat.params)) and tr.skipped_steps() == 1


def test_product_refuses_cpu_tensors(monkeypatch):
    monkeypatch.undo()                                 # drop the emulator: this is the shipped behaviour
    blk = cg.UpdateBlock(feat_dim=8, activation="swish", dropout=0.0)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        blk(torch.zeros(2, 8), torch.zeros(2, 8, 3))
    with pytest.raises(RuntimeError, match="GPU only"):
        cg.get_neighbor_list(np.zeros((3, 3)), "cpu", 1.0)


def test_collate_matches_reference_layout():
    z = load("cgvae_small.npz")
    sec = section(z, "vae_sym")
    n_at, n_cg = 9, 3
    samples = []
    for k in range(3):
        a = slice(k * n_at, (k + 1) * n_at)
        nb = sec["batch/nbr_list"]
        nb = nb[(nb[:, 0] >= k * n_at) & (nb[:, 0] < (k + 1) * n_at)] - k * n_at
        cnb = sec["batch/CG_nbr_list"]
        cnb = cnb[(cnb[:, 0] >= k * n_cg) & (cnb[:, 0] < (k + 1) * n_cg)] - k * n_cg
        be = sec["batch/bond_edge_list"]
        be = be[(be[:, 0] >= k * n_at) & (be[:, 0] < (k + 1) * n_at)] - k * n_at
        samples.append({"nxyz": sec["batch/nxyz"][a], "CG_nxyz": sec["batch/CG_nxyz"][k * n_cg:(k + 1) * n_cg],
                        "CG_mapping": sec["batch/CG_mapping"][a] - k * n_cg, "nbr_list": nb, "CG_nbr_list": cnb,
                        "bond_edge_list": be, "num_atoms": torch.tensor(n_at), "num_CGs": torch.tensor(n_cg)})
    batch = cg.CG_collate(samples)
    for key in ("nxyz", "CG_nxyz", "CG_mapping", "nbr_list", "CG_nbr_list", "bond_edge_list", "num_atoms", "num_CGs"):
        assert torch.equal(batch[key], sec["batch/" + key]), key


def test_deferred_gradients_merge_bias_into_weight_problem(monkeypatch):
    """ops.flush_deferred: the bias problem of a Dense layer joins the weight problem on the same gy; a second weight
    problem on the same gy (filter weight + filter bias of the 9-split block) stays separate."""
    import torch
    from coarsegrainingvae_b200 import ops
    seen = []
    monkeypatch.setattr(ops, "wgrad_grouped", lambda problems: seen.append(list(problems)))
    gy, gy2, x = torch.randn(12, 8), torch.randn(12, 8), torch.randn(12, 5)
    dW, db, dW2, db2, dW3 = torch.empty(8, 5), torch.empty(8), torch.empty(8, 5), torch.empty(8), torch.empty(8, 1)
    ops.flush_deferred([(gy, x, dW, None), (gy, None, None, db), (gy2, None, None, db2), (gy2, x, dW2, None),
                        (gy, x[:, :1], dW3, None)])
    (merged,) = seen
    assert len(merged) == 3
    assert merged[0][0] is gy and merged[0][2] is dW and merged[0][3] is db
    assert merged[1][0] is gy2 and merged[1][1] is x and merged[1][2] is dW2 and merged[1][3] is db2
    assert merged[2][2] is dW3 and merged[2][3] is None
    assert not ops.deferring(12)
    with ops.DeferredGrads():
        assert ops.deferring(12) and not ops.deferring(ops.DEFER_MAX_ROWS + 1)
    assert len(seen) == 2 and seen[1] == []


def test_torch_library_ops_are_registered_and_cuda_only():
    """torch.ops.cgvae_b200.*: schemas exist, fake kernels give the right shapes without a GPU, CPU tensors are refused
    (there is no CPU kernel behind the ops)."""
    import pytest
    import torch
    from coarsegrainingvae_b200 import torch_ops
    for name in torch_ops.OPS:
        assert hasattr(torch.ops.cgvae_b200, name), name
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x, W, b = torch.empty(12, 600, device="cuda"), torch.empty(1800, 600, device="cuda"), torch.empty(1800, device="cuda")
        assert tuple(torch.ops.cgvae_b200.dense(x, W, b).shape) == (12, 1800)
        y, a1, z1 = torch.ops.cgvae_b200.mlp2(x, torch.empty(600, 600, device="cuda"), torch.empty(600, device="cuda"), W, b, 1)
        assert tuple(y.shape) == (12, 1800) and tuple(a1.shape) == (12, 600) == tuple(z1.shape)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.cgvae_b200.dense(torch.zeros(2, 4), torch.zeros(3, 4), None)


def test_pcn_train_step_and_static_batch_cpu():
    """PCNTrainStep on the torch-optimiser path (emulated kernels): its loss is the oracle's restatement of the PCN loop
    (scripts/pcn_utils.py:160-183: MSE + gamma * bond term + kappa * dihedral term), the guard sits at gamma * 300, and
    to_static_pcn_batch pads every index list with its live count (the padded batch gives the same loss)."""
    from coarsegrainingvae_b200 import synthetic, train
    from coarsegrainingvae_b200.factory import build_pcn
    from oracle import cgvae_oracle as orc
    from oracle import graph_oracle as gorc
    from tests.golden_util import rel_err
    cfg = dict(synthetic.CONFIGS["c4_protein"])
    cfg.update(n_res=12, n_basis=32)
    rad = lambda xyz, c: gorc.radius_graph(np.asarray(xyz, dtype=np.float32), c)
    raw = synthetic.pcn_batch(cfg, 0, rad, n_proteins=2)
    torch.manual_seed(0)
    model = build_pcn(cfg["n_basis"], cfg["n_rbf"], cfg["cg_cutoff"], 2)
    gamma, kappa = 2.0, 0.5
    tr = train.PCNTrainStep(model, gamma, kappa, optimizer="torch")
    assert tr.loss_limit == gamma * 300.0 and tr.beta == 0.0
    out = model(raw)
    want = orc.pcn_loss(out[5].detach().double(), raw["xyz"].double(), raw["bond_edge_list"], raw["dihe_idxs"], gamma, kappa)[0]
    got = tr._loss(raw, None)
    assert rel_err(got, want) < 1e-5
    caps = {"CG_nbr_list": raw["CG_nbr_list"].shape[0] + 7, "bond_edge_list": raw["bond_edge_list"].shape[0] + 5,
            "dihe_idxs": raw["dihe_idxs"].shape[0] + 3, "ca_idx": raw["ca_idx"].shape[0] + 2}
    sb = train.to_static_pcn_batch(raw, caps)
    for key, cnt in (("CG_nbr_list", "CG_nbr_count"), ("bond_edge_list", "bond_count"), ("dihe_idxs", "dihe_count"), ("ca_idx", "ca_count")):
        n = int(sb[cnt])
        assert sb[key].shape[0] == caps[key] and n == raw[key].shape[0]
        assert torch.equal(sb[key][:n], raw[key]) and int(sb[key][n:].abs().sum()) == 0
    assert "seq" not in sb and sb["CG_nbr_symmetrize"] is True
    assert rel_err(tr._loss(sb, None), want) < 1e-5
    # one optimiser step moves the parameters; a loss past the guard does not
    tr.prepare(sb, None)
    before = [p.detach().clone() for p in tr.flat.params]
    tr.step(sb, None)
    assert any(not torch.equal(a, p) for a, p in zip(before, tr.flat.params))
    after = [p.detach().clone() for p in tr.flat.params]
    tr.loss_limit = 0.0
    tr.step(sb, None)
    assert all(torch.equal(a, p) for a, p in zip(after, tr.flat.params)) and tr.skipped_steps() == 1


def test_device_dataset_and_metrics_refuse_cpu():
    """the round-2 device-side helpers have no CPU path either."""
    from coarsegrainingvae_b200 import metrics
    from coarsegrainingvae_b200.data import DeviceCGDataset
    xyz = np.zeros((2, 4, 3), dtype=np.float32)
    with pytest.raises(RuntimeError):
        DeviceCGDataset(np.ones(4), xyz, np.array([0, 0, 1, 1]), np.array([[0, 1]]), 3.0, 5.0, device="cpu")
    with pytest.raises(RuntimeError):
        metrics.get_bond_graphs(torch.zeros(4, 3), torch.ones(4, dtype=torch.int64))
