"""CPU: the real Python modules / autograd nodes of coarsegrainingvae_b200, run on top of the TEST-ONLY
kernel emulator (tests/emulator.py), against the oracle's autograd.

float64 runs pin the hand-derived backward formulas the CUDA kernels implement (tolerance 1e-10);
float32 runs against the frozen reference outputs (tests/golden) pin layouts, residual fusion,
state_dict compatibility.  The kernels themselves are checked on the GPU (test_gpu_*.py)."""
import numpy as np
import pytest
import torch
from torch import nn

import coarsegrainingvae_b200 as cg
from oracle import cgvae_oracle as orc
from tests import emulator
from tests.golden_util import load, section, rel_err


@pytest.fixture(autouse=True)
def _emulated(monkeypatch):
    emulator.install(monkeypatch)


def _load_params(module, sec, dtype):
    sd = {k[2:]: v.to(dtype) if v.dtype.is_floating_point else v for k, v in sec.items() if k.startswith("P/")}
    module.to(dtype)
    module.load_state_dict(sd, strict=True)           # key set and shapes == the reference's state_dict
    return {k: v for k, v in sec.items() if k.startswith("G/")}


def _oracle_params(module):
    return {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in module.state_dict().items()}


def _check_grads(module, P, tol):
    n = 0
    for k, p in module.named_parameters():
        og = P[k].grad
        if og is None or float(og.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        else:
            assert p.grad is not None, k
            assert rel_err(p.grad, og) < tol, (k, rel_err(p.grad, og))
            n += 1
    assert n > 0


@pytest.mark.parametrize("tag,cls", [("k3", "EquiMessageBlock"), ("k4", "EquiMessageCross")])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_message_block_api(tag, cls, dtype, tol):
    z = load("blocks_small.npz")
    sec = section(z, tag)
    F, R, cutoff = int(z["meta/F"]), int(z["meta/R"]), float(z["meta/cutoff"])
    blk = getattr(cg, cls)(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0)
    G = _load_params(blk, sec, dtype)
    nbrs = torch.from_numpy(z["in/nbrs"])
    r = torch.from_numpy(z["in/r"]).to(dtype)
    s = sec["s"].to(dtype).requires_grad_()
    v = sec["v"].to(dtype).requires_grad_()
    ds, dv = blk(s, v, r, nbrs)
    assert rel_err(ds, sec["ds"]) < 2e-5 and rel_err(dv, sec["dv"]) < 2e-5
    ((ds * sec["gs"].to(dtype)).sum() + (dv * sec["gv"].to(dtype)).sum()).backward()
    # oracle autograd in the same dtype
    P = {"blk." + k: p for k, p in _oracle_params(blk).items()}
    so = sec["s"].to(dtype).requires_grad_()
    vo = sec["v"].to(dtype).requires_grad_()
    fn = orc.equi_message if tag == "k3" else orc.equi_message_cross
    ods, odv = fn(P, "blk", so, vo, r, nbrs, R, cutoff)
    ((ods * sec["gs"].to(dtype)).sum() + (odv * sec["gv"].to(dtype)).sum()).backward()
    assert rel_err(ds, ods) < tol and rel_err(dv, odv) < tol
    assert rel_err(s.grad, so.grad) < tol and rel_err(v.grad, vo.grad) < tol
    _check_grads(blk, {k[4:]: p for k, p in P.items()}, tol)
    if dtype == torch.float32:
        for k, g in G.items():
            assert rel_err(dict(blk.named_parameters())[k[2:]].grad, g) < 2e-5, k


def test_message_block_edge_weight():
    z = load("blocks_small.npz")
    sec = section(z, "k3w")
    blk = cg.EquiMessageBlock(feat_dim=int(z["meta/F"]), activation="swish", n_rbf=int(z["meta/R"]),
                              cutoff=float(z["meta/cutoff"]), dropout=0.0)
    _load_params(blk, sec, torch.float32)
    ds, dv = blk(sec["s"], sec["v"], torch.from_numpy(z["in/r"]), torch.from_numpy(z["in/nbrs"]), edge_wgt=sec["w"])
    assert rel_err(ds, sec["ds"]) < 2e-5 and rel_err(dv, sec["dv"]) < 2e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_pseudo_block_api(dtype, tol):
    z = load("blocks_small.npz")
    sec = section(z, "k9")
    F, R, cutoff = int(z["meta/F"]), int(z["meta/R"]), float(z["meta/cutoff"])
    blk = cg.EquiMessagePsuedo(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0)
    _load_params(blk, sec, dtype)
    nbrs = torch.from_numpy(z["in/nbrs"])
    r = torch.from_numpy(z["in/r"]).to(dtype)
    names = ("s", "sbar", "v", "vbar")
    onames = ("ds", "dsbar", "dv", "dvbar")
    ins = [sec[k].to(dtype).requires_grad_() for k in names]
    outs = blk(*ins, r, nbrs)
    sum((o * sec["g_" + k].to(dtype)).sum() for o, k in zip(outs, onames)).backward()
    P = {"blk." + k: p for k, p in _oracle_params(blk).items()}
    oins = [sec[k].to(dtype).requires_grad_() for k in names]
    oouts = orc.equi_message_pseudo(P, "blk", *oins, r, nbrs, R, cutoff)
    sum((o * sec["g_" + k].to(dtype)).sum() for o, k in zip(oouts, onames)).backward()
    for a, b, k in zip(outs, oouts, onames):
        assert rel_err(a, b) < tol, k
        assert rel_err(a, sec[k]) < 2e-5, k
    for a, b, k in zip(ins, oins, names):
        assert rel_err(a.grad, b.grad) < tol, k
    _check_grads(blk, {k[4:]: p for k, p in P.items()}, tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_update_and_contraction_api(dtype, tol):
    z = load("blocks_small.npz")
    F, R = int(z["meta/F"]), int(z["meta/R"])
    sec = section(z, "upd")
    blk = cg.UpdateBlock(feat_dim=F, activation="swish", dropout=0.0)
    _load_params(blk, sec, dtype)
    s = sec["s"].to(dtype).requires_grad_()
    v = sec["v"].to(dtype).requires_grad_()
    ds, dv = blk(s, v)
    ((ds * sec["gs"].to(dtype)).sum() + (dv * sec["gv"].to(dtype)).sum()).backward()
    P = {"blk." + k: p for k, p in _oracle_params(blk).items()}
    so, vo = sec["s"].to(dtype).requires_grad_(), sec["v"].to(dtype).requires_grad_()
    ods, odv = orc.update_block(P, "blk", so, vo)
    ((ods * sec["gs"].to(dtype)).sum() + (odv * sec["gv"].to(dtype)).sum()).backward()
    assert rel_err(ds, ods) < tol and rel_err(dv, odv) < tol and rel_err(ds, sec["ds"]) < 2e-5
    assert rel_err(s.grad, so.grad) < tol and rel_err(v.grad, vo.grad) < tol
    _check_grads(blk, {k[4:]: p for k, p in P.items()}, tol)

    sec = section(z, "con")
    blk = cg.ContractiveMessageBlock(feat_dim=F, activation="swish", n_rbf=R, cutoff=20.0, dropout=0.0)
    _load_params(blk, sec, dtype)
    s = sec["s"].to(dtype).requires_grad_()
    v = sec["v"].to(dtype).requires_grad_()
    dS, dV = blk(s, v, sec["r_iI"].to(dtype), sec["mapping"])
    ((dS * sec["gS"].to(dtype)).sum() + (dV * sec["gV"].to(dtype)).sum()).backward()
    assert rel_err(dS, sec["dS"]) < 2e-5 and rel_err(dV, sec["dV"]) < 2e-5
    assert rel_err(s.grad, sec["grad_s"]) < 2e-5 and rel_err(v.grad, sec["grad_v"]) < 2e-5
    P = {"blk." + k: p for k, p in _oracle_params(blk).items()}
    so, vo = sec["s"].to(dtype).requires_grad_(), sec["v"].to(dtype).requires_grad_()
    odS, odV = orc.contractive_message(P, "blk", so, vo, sec["r_iI"].to(dtype), sec["mapping"], R)
    ((odS * sec["gS"].to(dtype)).sum() + (odV * sec["gV"].to(dtype)).sum()).backward()
    assert rel_err(s.grad, so.grad) < tol and rel_err(v.grad, vo.grad) < tol
    _check_grads(blk, {k[4:]: p for k, p in P.items()}, tol)


def _build_vae(F, R, enc, dec, acut, ccut, breaksym):
    dec_net = cg.EquivariantPsuedoDecoder(n_atom_basis=F, n_rbf=R, cutoff=acut, num_conv=dec, activation="swish",
                                          breaksym=breaksym)
    enc_net = cg.EquiEncoder(n_conv=enc, n_atom_basis=F, n_rbf=R, cutoff=ccut, activation="swish", cg_mp=False, dir_mp=False)
    prior = cg.CGprior(n_conv=enc, n_atom_basis=F, n_rbf=R, cutoff=ccut, activation="swish", dir_mp=False)
    mu = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
    sg = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
    return cg.CGequiVAE(enc_net, dec_net, mu, sg, 3, feature_dim=F, prior_net=prior, det=False, equivariant=True)


@pytest.mark.parametrize("tag", ["vae_sym", "vae_nosym"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 5e-5)])
def test_cgvae_model(tag, dtype, tol):
    z = load("cgvae_small.npz")
    sec = section(z, tag)
    F, R, enc, dec, acut, ccut, breaksym, beta, gamma = [float(x) for x in sec["meta"]]
    model = _build_vae(int(F), int(R), int(enc), int(dec), acut, ccut, bool(breaksym))
    G = _load_params(model, sec, dtype)
    batch = {k[len("batch/"):]: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in sec.items()
             if k.startswith("batch/")}
    out = model(batch, eps=sec["eps"].to(dtype))
    loss = orc.training_loss(out, batch, beta, gamma)[0]
    loss.backward()
    spec = dict(n_basis=int(F), n_rbf=int(R), enc_nconv=int(enc), dec_nconv=int(dec), atom_cutoff=acut, cg_cutoff=ccut,
                decoder="pseudo", breaksym=bool(breaksym), activation="swish")
    P = _oracle_params(model)
    oout = orc.cgvae_forward(P, spec, batch, eps=sec["eps"].to(dtype))
    oloss = orc.training_loss(oout, batch, beta, gamma)[0]
    oloss.backward()
    for a, b, k in zip(out, oout, ("mu", "sigma", "pmu", "pstd", "xyz", "xyz_recon")):
        assert rel_err(a, b) < tol, k
    assert rel_err(loss, oloss) < tol
    _check_grads(model, P, tol)
    assert rel_err(out[5], sec["xyz_recon"]) < 5e-5 and rel_err(loss, sec["loss"]) < 5e-5
    assert torch.equal(model.CG2ChannelIdx(batch["CG_mapping"]), sec["chan"])
    if dtype == torch.float32:
        params = dict(model.named_parameters())
        for k, g in G.items():
            assert rel_err(params[k[2:]].grad, g) < 5e-5, k


@pytest.mark.parametrize("tag", ["pcn_cross", "pcn_plain"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 5e-5)])
def test_pcn_model(tag, dtype, tol):
    z = load("pcn_small.npz")
    sec = section(z, tag)
    F, R, dec, cutoff, cross_flag = [float(x) for x in sec["meta"]]
    net = cg.EquivariantDecoder(n_atom_basis=int(F), n_rbf=int(R), cutoff=cutoff, num_conv=int(dec), activation="swish",
                                cross_flag=bool(cross_flag))
    model = cg.PCN(net, feature_dim=int(F), offset=False)
    G = _load_params(model, sec, dtype)
    batch = {k[len("batch/"):]: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in sec.items()
             if k.startswith("batch/")}
    batch["seq"] = ["A" * batch["ca_xyz"].shape[0]]
    out = model(batch)
    loss = (out[5] - out[4]).pow(2).mean()
    loss.backward()
    spec = dict(n_basis=int(F), n_rbf=int(R), dec_nconv=int(dec), atom_cutoff=cutoff,
                decoder="cross" if cross_flag else "plain", activation="swish")
    P = _oracle_params(model)
    oout = orc.pcn_forward(P, spec, batch)
    oloss = (oout[5] - oout[4]).pow(2).mean()
    oloss.backward()
    assert rel_err(out[5], oout[5]) < tol and rel_err(loss, oloss) < tol
    assert rel_err(out[5], sec["xyz_recon"]) < 5e-5
    _check_grads(model, P, tol)
    if dtype == torch.float32:
        params = dict(model.named_parameters())
        for k, g in G.items():
            assert rel_err(params[k[2:]].grad, g) < 5e-5, k


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
