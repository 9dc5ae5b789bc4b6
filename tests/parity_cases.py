"""Parity cases shared by the CPU host-logic tests (kernel emulator, float64 + float32) and the GPU
tests (real sm_100a kernels through the C ABI, float32).  The product side runs on ``dev``; the oracle
always runs on CPU with the same weights (state_dict copy), inputs and eps.

Metric: tests.golden_util.rel_err = max|a-b| / max|b| (scale-relative).  Tolerances are passed in by
the caller: 1e-10 for float64-on-emulator (pins the backward derivations), 1e-5 for fp32 kernels vs
the fp32 oracle (the north-star gate), 2e-5..5e-5 vs the frozen reference outputs where two fp32
roundings stack up.
"""
import torch
from torch import nn

import coarsegrainingvae_b200 as cg
from oracle import cgvae_oracle as orc
from tests.golden_util import load, section, rel_err


def _load_params(module, sec, dtype, dev):
    sd = {k[2:]: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in sec.items() if k.startswith("P/")}
    module.to(dtype)
    module.load_state_dict(sd, strict=True)           # key set and shapes == the reference's state_dict
    module.to(dev)
    return {k[2:]: v for k, v in sec.items() if k.startswith("G/")}


def _oracle_params(module, prefix=""):
    return {prefix + k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point)
            for k, v in module.state_dict().items()}


def _check_grads(module, P, tol, prefix="", P64=None):
    """every parameter gradient vs the oracle's autograd.  With ``P64`` (the same oracle evaluated in float64 = ground
    truth) the criterion becomes conditioning-aware: per tensor, the kernel's error against the truth must be below
    ``tol`` OR below ten times the error the fp32 oracle itself makes on that tensor (summation-order luck on an
    ill-conditioned quantity); and the WHOLE gradient vector (what clip_grad_norm_ / Adam consume) must be within ``tol``
    of the truth in relative L2 norm.  The raw fp32-vs-fp32 difference is returned for reporting.
    (Needed for the UpdateBlock norm path: d sqrt(sum(Vv^2+1e-10)) / dVv divides by ~1e-5 when v is still ~0, which
    amplifies fp32 rounding in *both* implementations -- conv.py:600.)"""
    n = 0
    worst = 0.0
    num = den = 0.0
    for k, p in module.named_parameters():
        og = P[prefix + k].grad
        if og is None or float(og.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        raw = rel_err(p.grad, og)
        if P64 is None:
            assert raw < tol, (k, raw)
        else:
            truth = P64[prefix + k].grad
            e_kernel, e_ref = rel_err(p.grad, truth), rel_err(og, truth)
            assert e_kernel < max(tol, 10.0 * e_ref), (k, e_kernel, e_ref)
            num += float((p.grad.detach().cpu().double() - truth).pow(2).sum())
            den += float(truth.pow(2).sum())
        worst = max(worst, raw)
        n += 1
    assert n > 0
    if P64 is not None:
        assert (num / den) ** 0.5 < tol, ("global gradient L2 error", (num / den) ** 0.5)
    return worst


def _check_golden_grads(module, G, tol):
    params = dict(module.named_parameters())
    for k, g in G.items():
        assert rel_err(params[k].grad, g) < tol, (k, rel_err(params[k].grad, g))


def _leaf(t, dtype, dev):
    return t.to(dtype).to(dev).requires_grad_()


def message_block_api(dev, tag, cls, dtype, tol, gold_tol=2e-5):
    z = load("blocks_small.npz")
    sec = section(z, tag)
    F, R, cutoff = int(z["meta/F"]), int(z["meta/R"]), float(z["meta/cutoff"])
    blk = getattr(cg, cls)(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0)
    G = _load_params(blk, sec, dtype, dev)
    nbrs = torch.from_numpy(z["in/nbrs"])
    r = torch.from_numpy(z["in/r"]).to(dtype)
    gs, gv = sec["gs"].to(dtype), sec["gv"].to(dtype)
    s, v = _leaf(sec["s"], dtype, dev), _leaf(sec["v"], dtype, dev)
    ds, dv = blk(s, v, r.to(dev), nbrs.to(dev))
    ((ds * gs.to(dev)).sum() + (dv * gv.to(dev)).sum()).backward()
    P = _oracle_params(blk, "blk.")
    so, vo = _leaf(sec["s"], dtype, "cpu"), _leaf(sec["v"], dtype, "cpu")
    fn = orc.equi_message if tag == "k3" else orc.equi_message_cross
    ods, odv = fn(P, "blk", so, vo, r, nbrs, R, cutoff)
    ((ods * gs).sum() + (odv * gv).sum()).backward()
    assert rel_err(ds, ods) < tol and rel_err(dv, odv) < tol
    assert rel_err(s.grad, so.grad) < tol and rel_err(v.grad, vo.grad) < tol
    _check_grads(blk, P, tol, "blk.")
    assert rel_err(ds, sec["ds"]) < gold_tol and rel_err(dv, sec["dv"]) < gold_tol
    if dtype == torch.float32:
        _check_golden_grads(blk, G, gold_tol)


def message_block_edge_weight(dev, gold_tol=2e-5):
    z = load("blocks_small.npz")
    sec = section(z, "k3w")
    blk = cg.EquiMessageBlock(feat_dim=int(z["meta/F"]), activation="swish", n_rbf=int(z["meta/R"]),
                              cutoff=float(z["meta/cutoff"]), dropout=0.0)
    _load_params(blk, sec, torch.float32, dev)
    ds, dv = blk(sec["s"].to(dev), sec["v"].to(dev), torch.from_numpy(z["in/r"]).to(dev),
                 torch.from_numpy(z["in/nbrs"]).to(dev), edge_wgt=sec["w"].to(dev))
    assert rel_err(ds, sec["ds"]) < gold_tol and rel_err(dv, sec["dv"]) < gold_tol


def pseudo_block_api(dev, dtype, tol, gold_tol=2e-5):
    z = load("blocks_small.npz")
    sec = section(z, "k9")
    F, R, cutoff = int(z["meta/F"]), int(z["meta/R"]), float(z["meta/cutoff"])
    blk = cg.EquiMessagePsuedo(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0)
    G = _load_params(blk, sec, dtype, dev)
    nbrs = torch.from_numpy(z["in/nbrs"])
    r = torch.from_numpy(z["in/r"]).to(dtype)
    names, onames = ("s", "sbar", "v", "vbar"), ("ds", "dsbar", "dv", "dvbar")
    ins = [_leaf(sec[k], dtype, dev) for k in names]
    outs = blk(*ins, r.to(dev), nbrs.to(dev))
    sum((o * sec["g_" + k].to(dtype).to(dev)).sum() for o, k in zip(outs, onames)).backward()
    P = _oracle_params(blk, "blk.")
    oins = [_leaf(sec[k], dtype, "cpu") for k in names]
    oouts = orc.equi_message_pseudo(P, "blk", *oins, r, nbrs, R, cutoff)
    sum((o * sec["g_" + k].to(dtype)).sum() for o, k in zip(oouts, onames)).backward()
    for a, b, k in zip(outs, oouts, onames):
        assert rel_err(a, b) < tol, k
        assert rel_err(a, sec[k]) < gold_tol, k
    for a, b, k in zip(ins, oins, names):
        assert rel_err(a.grad, b.grad) < tol, k
    _check_grads(blk, P, tol, "blk.")
    if dtype == torch.float32:
        _check_golden_grads(blk, G, gold_tol)


def update_block_api(dev, dtype, tol, gold_tol=2e-5):
    z = load("blocks_small.npz")
    F = int(z["meta/F"])
    sec = section(z, "upd")
    blk = cg.UpdateBlock(feat_dim=F, activation="swish", dropout=0.0)
    G = _load_params(blk, sec, dtype, dev)
    gs, gv = sec["gs"].to(dtype), sec["gv"].to(dtype)
    s, v = _leaf(sec["s"], dtype, dev), _leaf(sec["v"], dtype, dev)
    ds, dv = blk(s, v)
    ((ds * gs.to(dev)).sum() + (dv * gv.to(dev)).sum()).backward()
    P = _oracle_params(blk, "blk.")
    so, vo = _leaf(sec["s"], dtype, "cpu"), _leaf(sec["v"], dtype, "cpu")
    ods, odv = orc.update_block(P, "blk", so, vo)
    ((ods * gs).sum() + (odv * gv).sum()).backward()
    assert rel_err(ds, ods) < tol and rel_err(dv, odv) < tol and rel_err(ds, sec["ds"]) < gold_tol
    assert rel_err(s.grad, so.grad) < tol and rel_err(v.grad, vo.grad) < tol
    _check_grads(blk, P, tol, "blk.")
    if dtype == torch.float32:
        _check_golden_grads(blk, G, gold_tol)


def contraction_block_api(dev, dtype, tol, gold_tol=2e-5):
    z = load("blocks_small.npz")
    F, R = int(z["meta/F"]), int(z["meta/R"])
    sec = section(z, "con")
    blk = cg.ContractiveMessageBlock(feat_dim=F, activation="swish", n_rbf=R, cutoff=20.0, dropout=0.0)
    G = _load_params(blk, sec, dtype, dev)
    gS, gV = sec["gS"].to(dtype), sec["gV"].to(dtype)
    r_iI = sec["r_iI"].to(dtype)
    s, v = _leaf(sec["s"], dtype, dev), _leaf(sec["v"], dtype, dev)
    dS, dV = blk(s, v, r_iI.to(dev), sec["mapping"].to(dev))
    ((dS * gS.to(dev)).sum() + (dV * gV.to(dev)).sum()).backward()
    P = _oracle_params(blk, "blk.")
    so, vo = _leaf(sec["s"], dtype, "cpu"), _leaf(sec["v"], dtype, "cpu")
    odS, odV = orc.contractive_message(P, "blk", so, vo, r_iI, sec["mapping"], R)
    ((odS * gS).sum() + (odV * gV).sum()).backward()
    assert rel_err(dS, odS) < tol and rel_err(dV, odV) < tol
    assert rel_err(dS, sec["dS"]) < gold_tol and rel_err(dV, sec["dV"]) < gold_tol
    assert rel_err(s.grad, so.grad) < tol and rel_err(v.grad, vo.grad) < tol
    _check_grads(blk, P, tol, "blk.")
    if dtype == torch.float32:
        _check_golden_grads(blk, G, gold_tol)


def build_vae(F, R, enc, dec, acut, ccut, breaksym, n_cgs=3, equivariant=True):
    """the constructor calls of scripts/run_ala.py:184-209"""
    dec_net = cg.EquivariantPsuedoDecoder(n_atom_basis=F, n_rbf=R, cutoff=acut, num_conv=dec, activation="swish",
                                          breaksym=breaksym)
    enc_net = cg.EquiEncoder(n_conv=enc, n_atom_basis=F, n_rbf=R, cutoff=ccut, activation="swish", cg_mp=False, dir_mp=False)
    prior = cg.CGprior(n_conv=enc, n_atom_basis=F, n_rbf=R, cutoff=ccut, activation="swish", dir_mp=False)
    mu = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
    sg = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
    return cg.CGequiVAE(enc_net, dec_net, mu, sg, n_cgs, feature_dim=F, prior_net=prior, det=False, equivariant=equivariant)


def _batch_to(batch, dtype, dev):
    return {k: ((v.to(dtype) if v.dtype.is_floating_point else v).to(dev) if torch.is_tensor(v) else v)
            for k, v in batch.items()}


def cgvae_model(dev, tag, dtype, tol, gold_tol=5e-5):
    equivariant = tag != "vae_noneq"                      # cgvae_noneq.npz: the Linear(F, 3F) decoder head
    z = load("cgvae_small.npz" if equivariant else "cgvae_noneq.npz")
    sec = section(z, tag)
    F, R, enc, dec, acut, ccut, breaksym, beta, gamma = [float(x) for x in sec["meta"]]
    model = build_vae(int(F), int(R), int(enc), int(dec), acut, ccut, bool(breaksym), equivariant=equivariant)
    G = _load_params(model, sec, dtype, dev)
    batch = {k[len("batch/"):]: v for k, v in sec.items() if k.startswith("batch/")}
    cpu_batch = _batch_to(batch, dtype, "cpu")
    dev_batch = _batch_to(batch, dtype, dev)
    eps = sec["eps"].to(dtype)
    out = model(dev_batch, eps=eps.to(dev))
    loss = orc.training_loss(out, dev_batch, beta, gamma)[0]
    loss.backward()
    spec = dict(n_basis=int(F), n_rbf=int(R), enc_nconv=int(enc), dec_nconv=int(dec), atom_cutoff=acut, cg_cutoff=ccut,
                decoder="pseudo", breaksym=bool(breaksym), activation="swish", equivariant=equivariant)
    P = _oracle_params(model)
    oout = orc.cgvae_forward(P, spec, cpu_batch, eps=eps)
    oloss = orc.training_loss(oout, cpu_batch, beta, gamma)[0]
    oloss.backward()
    for a, b, k in zip(out, oout, ("mu", "sigma", "pmu", "pstd", "xyz", "xyz_recon")):
        assert rel_err(a, b) < tol, (k, rel_err(a, b))
    assert rel_err(loss, oloss) < tol
    _check_grads(model, P, tol)
    assert rel_err(out[5], sec["xyz_recon"]) < gold_tol and rel_err(loss, sec["loss"]) < gold_tol
    assert torch.equal(model.CG2ChannelIdx(dev_batch["CG_mapping"]).cpu(), sec["chan"])
    if dtype == torch.float32:
        _check_golden_grads(model, G, gold_tol)


def pcn_model(dev, tag, dtype, tol, gold_tol=5e-5):
    z = load("pcn_small.npz")
    sec = section(z, tag)
    F, R, dec, cutoff, cross_flag = [float(x) for x in sec["meta"]]
    net = cg.EquivariantDecoder(n_atom_basis=int(F), n_rbf=int(R), cutoff=cutoff, num_conv=int(dec), activation="swish",
                                cross_flag=bool(cross_flag))
    model = cg.PCN(net, feature_dim=int(F), offset=False)
    G = _load_params(model, sec, dtype, dev)
    batch = {k[len("batch/"):]: v for k, v in sec.items() if k.startswith("batch/")}
    batch["seq"] = ["A" * batch["ca_xyz"].shape[0]]
    cpu_batch, dev_batch = _batch_to(batch, dtype, "cpu"), _batch_to(batch, dtype, dev)
    out = model(dev_batch)
    loss = (out[5] - out[4]).pow(2).mean()
    loss.backward()
    spec = dict(n_basis=int(F), n_rbf=int(R), dec_nconv=int(dec), atom_cutoff=cutoff,
                decoder="cross" if cross_flag else "plain", activation="swish")
    P = _oracle_params(model)
    oout = orc.pcn_forward(P, spec, cpu_batch)
    oloss = (oout[5] - oout[4]).pow(2).mean()
    oloss.backward()
    assert rel_err(out[5], oout[5]) < tol and rel_err(loss, oloss) < tol
    assert rel_err(out[5], sec["xyz_recon"]) < gold_tol
    _check_grads(model, P, tol)
    if dtype == torch.float32:
        _check_golden_grads(model, G, gold_tol)


ACTIVATIONS = ("ReLU", "Tanh", "sigmoid", "shifted_softplus", "LeakyReLU", "ELU")


def blocks_with_activation(dev, act, dtype, tol):
    """every block type built with a non-default entry of the reference's layer_types registry (modules.py:32-42): outputs,
    input gradients and parameter gradients vs the oracle run with the same activation (golden inputs and weights)."""
    z = load("blocks_small.npz")
    F, R, cutoff = int(z["meta/F"]), int(z["meta/R"]), float(z["meta/cutoff"])
    nbrs = torch.from_numpy(z["in/nbrs"])
    r = torch.from_numpy(z["in/r"]).to(dtype)
    # 3-split message block
    sec = section(z, "k3")
    blk = cg.EquiMessageBlock(feat_dim=F, activation=act, n_rbf=R, cutoff=cutoff, dropout=0.0)
    _load_params(blk, sec, dtype, dev)
    gs, gv = sec["gs"].to(dtype), sec["gv"].to(dtype)
    s, v = _leaf(sec["s"], dtype, dev), _leaf(sec["v"], dtype, dev)
    ds, dv = blk(s, v, r.to(dev), nbrs.to(dev))
    ((ds * gs.to(dev)).sum() + (dv * gv.to(dev)).sum()).backward()
    P = _oracle_params(blk, "blk.")
    so, vo = _leaf(sec["s"], dtype, "cpu"), _leaf(sec["v"], dtype, "cpu")
    ods, odv = orc.equi_message(P, "blk", so, vo, r, nbrs, R, cutoff, act=act)
    ((ods * gs).sum() + (odv * gv).sum()).backward()
    assert rel_err(ds, ods) < tol and rel_err(dv, odv) < tol
    assert rel_err(s.grad, so.grad) < tol and rel_err(v.grad, vo.grad) < tol
    _check_grads(blk, P, tol, "blk.")
    # 9-split block
    sec = section(z, "k9")
    blk = cg.EquiMessagePsuedo(feat_dim=F, activation=act, n_rbf=R, cutoff=cutoff, dropout=0.0)
    _load_params(blk, sec, dtype, dev)
    names, onames = ("s", "sbar", "v", "vbar"), ("ds", "dsbar", "dv", "dvbar")
    ins = [_leaf(sec[k], dtype, dev) for k in names]
    outs = blk(*ins, r.to(dev), nbrs.to(dev))
    sum((o * sec["g_" + k].to(dtype).to(dev)).sum() for o, k in zip(outs, onames)).backward()
    P = _oracle_params(blk, "blk.")
    oins = [_leaf(sec[k], dtype, "cpu") for k in names]
    oouts = orc.equi_message_pseudo(P, "blk", *oins, r, nbrs, R, cutoff, act=act)
    sum((o * sec["g_" + k].to(dtype)).sum() for o, k in zip(oouts, onames)).backward()
    for a, b, k in zip(outs, oouts, onames):
        assert rel_err(a, b) < tol, k
    for a, b, k in zip(ins, oins, names):
        assert rel_err(a.grad, b.grad) < tol, k
    _check_grads(blk, P, tol, "blk.")
    # update block
    sec = section(z, "upd")
    blk = cg.UpdateBlock(feat_dim=F, activation=act, dropout=0.0)
    _load_params(blk, sec, dtype, dev)
    gs, gv = sec["gs"].to(dtype), sec["gv"].to(dtype)
    s, v = _leaf(sec["s"], dtype, dev), _leaf(sec["v"], dtype, dev)
    ds, dv = blk(s, v)
    ((ds * gs.to(dev)).sum() + (dv * gv.to(dev)).sum()).backward()
    P = _oracle_params(blk, "blk.")
    so, vo = _leaf(sec["s"], dtype, "cpu"), _leaf(sec["v"], dtype, "cpu")
    ods, odv = orc.update_block(P, "blk", so, vo, act=act)
    ((ods * gs).sum() + (odv * gv).sum()).backward()
    assert rel_err(ds, ods) < tol and rel_err(dv, odv) < tol
    assert rel_err(s.grad, so.grad) < tol and rel_err(v.grad, vo.grad) < tol
    _check_grads(blk, P, tol, "blk.")
    # atoms -> beads contraction
    sec = section(z, "con")
    blk = cg.ContractiveMessageBlock(feat_dim=F, activation=act, n_rbf=R, cutoff=20.0, dropout=0.0)
    _load_params(blk, sec, dtype, dev)
    gS, gV = sec["gS"].to(dtype), sec["gV"].to(dtype)
    r_iI = sec["r_iI"].to(dtype)
    s, v = _leaf(sec["s"], dtype, dev), _leaf(sec["v"], dtype, dev)
    dS, dV = blk(s, v, r_iI.to(dev), sec["mapping"].to(dev))
    ((dS * gS.to(dev)).sum() + (dV * gV.to(dev)).sum()).backward()
    P = _oracle_params(blk, "blk.")
    so, vo = _leaf(sec["s"], dtype, "cpu"), _leaf(sec["v"], dtype, "cpu")
    odS, odV = orc.contractive_message(P, "blk", so, vo, r_iI, sec["mapping"], R, act=act)
    ((odS * gS).sum() + (odV * gV).sum()).backward()
    assert rel_err(dS, odS) < tol and rel_err(dV, odV) < tol
    assert rel_err(s.grad, so.grad) < tol and rel_err(v.grad, vo.grad) < tol
    _check_grads(blk, P, tol, "blk.")


def sampling_case(dev, tag, dtype, tol, gold_tol=5e-5, graphed=False):
    """train.sample_single (and train.GraphedSampler when ``graphed``) against the loop of scripts/sampling.py:265-293
    run on the REAL reference (tests/golden/sampling_small.npz, fixed noise) and against oracle.sample_ensemble."""
    from coarsegrainingvae_b200 import train
    sec = section(load("sampling_small.npz"), tag)
    F, R, enc, dec, acut, ccut, breaksym, n_ens = [float(x) for x in sec["meta"]]
    model = build_vae(int(F), int(R), int(enc), int(dec), acut, ccut, bool(breaksym), n_cgs=4)
    _load_params(model, sec, dtype, dev)
    batch = {k[len("batch/"):]: v for k, v in sec.items() if k.startswith("batch/")}
    cpu_batch, dev_batch = _batch_to(batch, dtype, "cpu"), _batch_to(batch, dtype, dev)
    eps_m, eps_r = sec["eps_members"].to(dtype), sec["eps_recon"].to(dtype)
    members, xyz_recon, mu, sigma = train.sample_single(model, dev_batch, int(n_ens), eps_m.to(dev), eps_r.to(dev))
    spec = dict(n_basis=int(F), n_rbf=int(R), enc_nconv=int(enc), dec_nconv=int(dec), atom_cutoff=acut, cg_cutoff=ccut,
                decoder="pseudo", breaksym=bool(breaksym), activation="swish")
    with torch.no_grad():
        P = _oracle_params(model)
        omu, osig, oxyz = orc.sample_ensemble(P, spec, cpu_batch, eps_m)
        orecon = orc.cgvae_forward(P, spec, cpu_batch, eps=eps_r)[5]
    assert rel_err(mu, omu) < tol and rel_err(sigma, osig) < tol
    assert rel_err(members, oxyz) < tol, rel_err(members, oxyz)
    assert rel_err(xyz_recon, orecon) < tol
    # the frozen geometries of the real reference, member by member
    assert rel_err(members, sec["members"]) < gold_tol and rel_err(xyz_recon, sec["xyz_recon"]) < gold_tol
    assert rel_err(mu, sec["prior_mu"]) < gold_tol and rel_err(sigma, sec["prior_sigma"]) < gold_tol
    if graphed:
        caps = {"nbr_list": batch["nbr_list"].shape[0] + 5, "CG_nbr_list": batch["CG_nbr_list"].shape[0] + 3,
                "bond_edge_list": batch["bond_edge_list"].shape[0] + 2}
        static = _batch_to(train.to_static_batch(batch, caps), dtype, dev)
        sampler = train.GraphedSampler(model, static, int(n_ens))
        got = sampler.sample(static, eps_m.to(dev))
        assert rel_err(got, oxyz) < tol and rel_err(got, sec["members"]) < gold_tol
