"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference read-only through oracle/ref_shim.py (torch_scatter / ase stubs),
runs the reference modules on small seeded inputs and freezes inputs, weights
(state_dict), outputs and autograd gradients.  The reference has no golden vectors of
its own; these fixtures are what pins the oracle -- and through it
the CUDA path -- to the reference's behaviour.  The fixtures travel to the GPU box; the
reference does not.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_shim  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _sym_edges(n, p, gen):
    """random symmetric directed edge list in make_directed() layout: [i<j half ; flipped half]."""
    iu = torch.triu_indices(n, n, 1)
    keep = torch.rand(iu.shape[1], generator=gen) < p
    half = iu[:, keep].t().contiguous()
    return half  # undirected (j > i), sorted row-major like get_neighbor_list


def _save(name, store):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **store)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def _state(prefix, module, store):
    for k, v in module.state_dict().items():
        store["%s/P/%s" % (prefix, k)] = _np(v)


def _grads(prefix, module, store):
    for k, p in module.named_parameters():
        if p.grad is not None:
            store["%s/G/%s" % (prefix, k)] = _np(p.grad)


def blocks(ref_conv):
    gen = torch.Generator().manual_seed(7)
    F, R, N, cutoff = 16, 4, 11, 4.0
    store = {"meta/F": F, "meta/R": R, "meta/N": N, "meta/cutoff": cutoff}
    xyz = torch.randn(N, 3, generator=gen) * 1.5
    half = _sym_edges(N, 0.6, gen)
    nbrs, _ = ref_conv.make_directed(half)
    r = xyz[nbrs[:, 1]] - xyz[nbrs[:, 0]]
    store["in/xyz"], store["in/half"], store["in/nbrs"], store["in/r"] = map(_np, (xyz, half, nbrs, r))

    def rnd(*shape):
        return torch.randn(*shape, generator=gen)

    torch.manual_seed(11)
    # --- EquiMessageBlock (K=3) and EquiMessageCross (K=4) ---------------------------------
    for tag, cls in (("k3", ref_conv.EquiMessageBlock), ("k4", ref_conv.EquiMessageCross)):
        blk = cls(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0)
        for p in blk.parameters():          # biases are zero-initialised: make them matter
            if p.dim() == 1:
                p.data.normal_(0, 0.3, generator=gen)
        s = rnd(N, F).requires_grad_()
        v = rnd(N, F, 3).requires_grad_()
        ds, dv = blk(s, v, r, nbrs)
        gs, gv = rnd(N, F), rnd(N, F, 3)
        ((ds * gs).sum() + (dv * gv).sum()).backward()
        _state(tag, blk, store)
        _grads(tag, blk, store)
        for k, t in (("s", s), ("v", v), ("ds", ds), ("dv", dv), ("gs", gs), ("gv", gv),
                     ("grad_s", s.grad), ("grad_v", v.grad)):
            store["%s/%s" % (tag, k)] = _np(t)
    # edge-weighted variant of K=3 (diffpool caller, conv.py:528-533)
    blk = ref_conv.EquiMessageBlock(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0)
    s, v, w = rnd(N, F), rnd(N, F, 3), torch.rand(nbrs.shape[0], generator=gen)
    ds, dv = blk(s, v, r, nbrs, edge_wgt=w)
    _state("k3w", blk, store)
    for k, t in (("s", s), ("v", v), ("w", w), ("ds", ds), ("dv", dv)):
        store["k3w/%s" % k] = _np(t)
    # --- EquiMessagePsuedo (K=9) --------------------------------------------------------------
    blk = ref_conv.EquiMessagePsuedo(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0)
    for p in blk.parameters():
        if p.dim() == 1:
            p.data.normal_(0, 0.3, generator=gen)
    ins = [rnd(N, F).requires_grad_(), rnd(N, F).requires_grad_(),
           rnd(N, F, 3).requires_grad_(), rnd(N, F, 3).requires_grad_()]
    outs = blk(ins[0], ins[1], ins[2], ins[3], r, nbrs)
    gouts = [rnd(*o.shape) for o in outs]
    sum((o * g).sum() for o, g in zip(outs, gouts)).backward()
    _state("k9", blk, store)
    _grads("k9", blk, store)
    for k, t in zip(("s", "sbar", "v", "vbar"), ins):
        store["k9/%s" % k] = _np(t)
        store["k9/grad_%s" % k] = _np(t.grad)
    for k, t, g in zip(("ds", "dsbar", "dv", "dvbar"), outs, gouts):
        store["k9/%s" % k] = _np(t)
        store["k9/g_%s" % k] = _np(g)
    # --- UpdateBlock ---------------------------------------------------------------------------
    blk = ref_conv.UpdateBlock(feat_dim=F, activation="swish", dropout=0.0)
    for p in blk.parameters():
        if p.dim() == 1:
            p.data.normal_(0, 0.3, generator=gen)
    s = rnd(N, F).requires_grad_()
    v = rnd(N, F, 3).requires_grad_()
    ds, dv = blk(s, v)
    gs, gv = rnd(N, F), rnd(N, F, 3)
    ((ds * gs).sum() + (dv * gv).sum()).backward()
    _state("upd", blk, store)
    _grads("upd", blk, store)
    for k, t in (("s", s), ("v", v), ("ds", ds), ("dv", dv), ("gs", gs), ("gv", gv),
                 ("grad_s", s.grad), ("grad_v", v.grad)):
        store["upd/%s" % k] = _np(t)
    # --- ContractiveMessageBlock ---------------------------------------------------------------
    blk = ref_conv.ContractiveMessageBlock(feat_dim=F, activation="swish", n_rbf=R, cutoff=20.0, dropout=0.0)
    for p in blk.parameters():
        if p.dim() == 1:
            p.data.normal_(0, 0.3, generator=gen)
    mapping = torch.tensor([0, 0, 1, 0, 2, 2, 1, 1, 2, 0, 1])
    cg_xyz = torch.stack([xyz[mapping == b].mean(0) for b in range(3)])
    r_iI = xyz - cg_xyz[mapping]
    s = rnd(N, F).requires_grad_()
    v = rnd(N, F, 3).requires_grad_()
    dS, dV = blk(s, v, r_iI, mapping)
    gS, gV = rnd(*dS.shape), rnd(*dV.shape)
    ((dS * gS).sum() + (dV * gV).sum()).backward()
    _state("con", blk, store)
    _grads("con", blk, store)
    for k, t in (("s", s), ("v", v), ("r_iI", r_iI), ("mapping", mapping), ("dS", dS), ("dV", dV),
                 ("gS", gS), ("gV", gV), ("grad_s", s.grad), ("grad_v", v.grad)):
        store["con/%s" % k] = _np(t)
    _save("blocks_small.npz", store)


def blocks_activations(ref_conv):
    """blocks_act.npz: the K=3 message block and the UpdateBlock built by the REAL reference with every other entry of its
    layer_types registry (modules.py:32-42) -- pins the oracle's activation table, which the CUDA epilogue codes 2..7 are
    then tested against."""
    gen = torch.Generator().manual_seed(77)
    F, R, N, cutoff = 16, 4, 11, 4.0
    store = {"meta/F": F, "meta/R": R, "meta/N": N, "meta/cutoff": cutoff}
    xyz = torch.randn(N, 3, generator=gen) * 1.5
    half = _sym_edges(N, 0.6, gen)
    nbrs, _ = ref_conv.make_directed(half)
    r = xyz[nbrs[:, 1]] - xyz[nbrs[:, 0]]
    store["in/nbrs"], store["in/r"] = _np(nbrs), _np(r)

    def rnd(*shape):
        return torch.randn(*shape, generator=gen)

    torch.manual_seed(12)
    for act in ("ReLU", "Tanh", "sigmoid", "shifted_softplus", "LeakyReLU", "ELU"):
        blk = ref_conv.EquiMessageBlock(feat_dim=F, activation=act, n_rbf=R, cutoff=cutoff, dropout=0.0)
        for p in blk.parameters():
            if p.dim() == 1:
                p.data.normal_(0, 0.3, generator=gen)
        s, v = rnd(N, F).requires_grad_(), rnd(N, F, 3).requires_grad_()
        ds, dv = blk(s, v, r, nbrs)
        gs, gv = rnd(N, F), rnd(N, F, 3)
        ((ds * gs).sum() + (dv * gv).sum()).backward()
        tag = "msg_" + act
        _state(tag, blk, store)
        _grads(tag, blk, store)
        for k, t in (("s", s), ("v", v), ("ds", ds), ("dv", dv), ("gs", gs), ("gv", gv), ("grad_s", s.grad), ("grad_v", v.grad)):
            store["%s/%s" % (tag, k)] = _np(t)
        blk = ref_conv.UpdateBlock(feat_dim=F, activation=act, dropout=0.0)
        for p in blk.parameters():
            if p.dim() == 1:
                p.data.normal_(0, 0.3, generator=gen)
        s, v = rnd(N, F).requires_grad_(), rnd(N, F, 3).requires_grad_()
        ds, dv = blk(s, v)
        gs, gv = rnd(N, F), rnd(N, F, 3)
        ((ds * gs).sum() + (dv * gv).sum()).backward()
        tag = "upd_" + act
        _state(tag, blk, store)
        _grads(tag, blk, store)
        for k, t in (("s", s), ("v", v), ("ds", ds), ("dv", dv), ("gs", gs), ("gv", gv), ("grad_s", s.grad), ("grad_v", v.grad)):
            store["%s/%s" % (tag, k)] = _np(t)
    _save("blocks_act.npz", store)


def _molecule_batch(ref_data, gen, n_conf, n_atoms, mapping, atom_cutoff, cg_cutoff, spread):
    samples = []
    n_cg = int(mapping.max()) + 1
    for _ in range(n_conf):
        xyz = torch.randn(n_atoms, 3, generator=gen) * spread
        z = torch.randint(1, 9, (n_atoms,), generator=gen).float()
        cg_xyz = torch.stack([xyz[mapping == b].mean(0) for b in range(n_cg)])
        nbr = ref_data.get_neighbor_list(xyz, "cpu", atom_cutoff, True)
        cg_nbr = ref_data.get_neighbor_list(cg_xyz, "cpu", cg_cutoff, True)
        bonds = torch.stack([torch.arange(n_atoms - 1), torch.arange(1, n_atoms)], 1)
        samples.append({
            "nxyz": torch.cat([z[:, None], xyz], 1),
            "CG_nxyz": torch.cat([torch.arange(n_cg).float()[:, None], cg_xyz], 1),
            "CG_mapping": mapping.clone(), "nbr_list": nbr, "CG_nbr_list": cg_nbr,
            "bond_edge_list": bonds, "num_atoms": torch.tensor(n_atoms), "num_CGs": torch.tensor(n_cg),
        })
    return ref_data.CG_collate(samples)


def cgvae_small(ref_cgvae, ref_data, fname="cgvae_small.npz", cases=(("vae_sym", True, True), ("vae_nosym", False, True)),
                seed=21):
    """cases: (tag, breaksym, equivariant).  The default call writes cgvae_small.npz exactly as before; the second call in
    main() adds cgvae_noneq.npz for the non-equivariant decoder head (cgvae.py:469-471)."""
    from torch import nn
    gen = torch.Generator().manual_seed(seed)
    F, R, enc, dec = 24, 5, 2, 2
    atom_cutoff, cg_cutoff = 3.5, 6.0
    mapping = torch.tensor([0, 0, 0, 1, 1, 2, 2, 2, 2])
    store = {}
    for tag, breaksym, equivariant in cases:
        batch = _molecule_batch(ref_data, gen, 3, 9, mapping, atom_cutoff, cg_cutoff, 1.4)
        torch.manual_seed(123)
        dec_net = ref_cgvae.EquivariantPsuedoDecoder(n_atom_basis=F, n_rbf=R, cutoff=atom_cutoff,
                                                     num_conv=dec, activation="swish", breaksym=breaksym)
        enc_net = ref_cgvae.EquiEncoder(n_conv=enc, n_atom_basis=F, n_rbf=R, cutoff=cg_cutoff,
                                        activation="swish", cg_mp=False, dir_mp=False)
        prior = ref_cgvae.CGprior(n_conv=enc, n_atom_basis=F, n_rbf=R, cutoff=cg_cutoff,
                                  activation="swish", dir_mp=False)
        mu_net = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
        sg_net = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
        model = ref_cgvae.CGequiVAE(enc_net, dec_net, mu_net, sg_net, 3, feature_dim=F, prior_net=prior,
                                    det=False, equivariant=equivariant)
        # reparametrize draws torch.randn_like(sigma): reseed right before forward and
        # regenerate the same eps afterwards (same generator state, same shape).
        torch.manual_seed(999)
        mu, sigma, pmu, pstd, xyz, xyz_recon = model(batch)
        torch.manual_seed(999)
        eps = torch.randn_like(sigma)
        recon = (xyz_recon - xyz).pow(2).mean()
        kl = 0.5 * ((sigma.pow(2) / pstd.pow(2)).sum(-1) + ((mu - pmu).pow(2) / pstd).sum(-1)
                    + torch.log(pstd.pow(2)).sum(-1) - torch.log(sigma.pow(2)).sum(-1) - sigma.shape[-1]).mean()
        e = batch["bond_edge_list"]
        gd = ((xyz_recon[e[:, 0]] - xyz_recon[e[:, 1]]).pow(2).sum(-1) + 1e-6).sqrt()
        dd = ((xyz[e[:, 0]] - xyz[e[:, 1]]).pow(2).sum(-1) + 1e-6).sqrt()
        gl = (gd - dd).pow(2).mean()
        beta, gamma = 0.05, 25.0
        loss = recon + kl * beta + gl * gamma
        loss.backward()
        chan = model.CG2ChannelIdx(batch["CG_mapping"])
        _state(tag, model, store)
        _grads(tag, model, store)
        for k, t in batch.items():
            store["%s/batch/%s" % (tag, k)] = _np(t)
        for k, t in (("eps", eps), ("mu", mu), ("sigma", sigma), ("pmu", pmu), ("pstd", pstd),
                     ("xyz_recon", xyz_recon), ("loss", loss), ("recon", recon), ("kl", kl), ("graph", gl),
                     ("chan", chan)):
            store["%s/%s" % (tag, k)] = _np(t)
        store["%s/meta" % tag] = np.array([F, R, enc, dec, atom_cutoff, cg_cutoff, int(breaksym), beta, gamma])
    _save(fname, store)


def pcn_small(ref_cgvae, ref_data):
    gen = torch.Generator().manual_seed(33)
    F, R, dec, cutoff = 16, 4, 2, 9.0
    n_res, per = 7, 4
    store = {}
    for tag, cross_flag in (("pcn_cross", True), ("pcn_plain", False)):
        ca = torch.cumsum(torch.randn(n_res, 3, generator=gen) * 2.2, 0)
        mapping = torch.arange(n_res).repeat_interleave(per)
        xyz = ca[mapping] + torch.randn(n_res * per, 3, generator=gen)
        ca_idx = torch.arange(n_res) * per + 1
        xyz[ca_idx] = ca
        batch = {
            "xyz": xyz, "ca_xyz": ca, "res": torch.randint(0, 20, (n_res,), generator=gen),
            "cg_map": mapping, "bond_edge_list": torch.stack([torch.arange(n_res * per - 1),
                                                               torch.arange(1, n_res * per)], 1),
            "CG_nbr_list": ref_data.get_neighbor_list(ca, "cpu", cutoff, True),
            "seq": ["A" * n_res], "ca_idx": ca_idx,
        }
        torch.manual_seed(5)
        net = ref_cgvae.EquivariantDecoder(n_atom_basis=F, n_rbf=R, cutoff=cutoff, num_conv=dec,
                                           activation="swish", cross_flag=cross_flag)
        model = ref_cgvae.PCN(net, feature_dim=F, offset=False)
        out = model(batch)
        xyz_recon = out[5]
        loss = (xyz_recon - xyz).pow(2).mean()
        loss.backward()
        _state(tag, model, store)
        _grads(tag, model, store)
        for k, t in batch.items():
            if k != "seq":
                store["%s/batch/%s" % (tag, k)] = _np(t)
        store["%s/xyz_recon" % tag] = _np(xyz_recon)
        store["%s/loss" % tag] = _np(loss)
        store["%s/meta" % tag] = np.array([F, R, dec, cutoff, int(cross_flag)])
    _save("pcn_small.npz", store)


def graphs(ref_data, ref_cgvae):
    gen = torch.Generator().manual_seed(5)
    store = {}
    # jittered lattice (distances cluster near the cutoff -> stresses the <= comparison) + gaussian cloud
    g = torch.arange(7).float()
    lat = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) * 1.5
    lat = lat + (torch.rand(lat.shape, generator=gen) - 0.5) * 0.2
    cloud = torch.randn(257, 3, generator=gen) * 3.0
    exact = torch.stack(torch.meshgrid(g[:5], g[:5], g[:5], indexing="ij"), -1).reshape(-1, 3)  # ties at d == cutoff
    for name, pts, cuts in (("lattice", lat, (1.5, 2.6, 4.5)), ("cloud", cloud, (0.7, 2.0, 5.0)),
                            ("exact", exact, (1.0, 2.0, 3.0))):
        store["%s/xyz" % name] = _np(pts)
        for c in cuts:
            store["%s/und_%g" % (name, c)] = _np(ref_data.get_neighbor_list(pts, "cpu", c, True))
            store["%s/dir_%g" % (name, c)] = _np(ref_data.get_neighbor_list(pts, "cpu", c, False))
    model = ref_cgvae.CGequiVAE.__new__(ref_cgvae.CGequiVAE)
    for i, m in enumerate((torch.tensor([0, 0, 1, 0, 2, 2, 1, 1, 2, 0, 1]),
                           torch.randint(0, 13, (200,), generator=gen),
                           torch.arange(40).repeat_interleave(5),
                           torch.tensor([3, 3, 3, 7, 7, 0]))):
        store["chan/%d/mapping" % i] = _np(m)
        store["chan/%d/index" % i] = _np(ref_cgvae.CGequiVAE.CG2ChannelIdx(model, m))
    _save("graphs.npz", store)


def sampling(ref_cgvae, ref_data):
    """sampling_small.npz: the ensemble loop of scripts/sampling.py:252-293 (``sample_single``) driven on the REAL reference
    model.  scripts/sampling.py itself does not import here (ase / mdshare / pyemma are absent), so its three model calls
    are issued verbatim -- ``model.prior_net`` once (:268), then per member ``sample_normal`` (:247-250: ``eps =
    torch.randn_like(sigma); z = eps.mul(sigma).add_(mu)``) and ``model.decoder(cg_xyz, CG_nbr_list, H, H, mapping,
    num_CGs)`` (:277-279), then the reconstruction forward ``model(batch)`` (:293) -- under one fixed torch seed; the noise
    of every draw is recovered by replaying the generator in the same call order."""
    from torch import nn
    gen = torch.Generator().manual_seed(41)
    F, R, enc, dec = 24, 5, 2, 3
    atom_cutoff, cg_cutoff = 3.5, 6.0
    n_ens = 5
    mapping = torch.tensor([0, 0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3])
    store = {}
    for tag, breaksym in (("sym", True), ("nosym", False)):
        batch = _molecule_batch(ref_data, gen, 1, 12, mapping, atom_cutoff, cg_cutoff, 1.6)   # batch_size 1 (sampling.py:335)
        torch.manual_seed(321)
        dec_net = ref_cgvae.EquivariantPsuedoDecoder(n_atom_basis=F, n_rbf=R, cutoff=atom_cutoff,
                                                     num_conv=dec, activation="swish", breaksym=breaksym)
        enc_net = ref_cgvae.EquiEncoder(n_conv=enc, n_atom_basis=F, n_rbf=R, cutoff=cg_cutoff,
                                        activation="swish", cg_mp=False, dir_mp=False)
        prior = ref_cgvae.CGprior(n_conv=enc, n_atom_basis=F, n_rbf=R, cutoff=cg_cutoff,
                                  activation="swish", dir_mp=False)
        mu_net = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
        sg_net = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
        model = ref_cgvae.CGequiVAE(enc_net, dec_net, mu_net, sg_net, 4, feature_dim=F, prior_net=prior,
                                    det=False, equivariant=True)
        for p in model.parameters():         # default-initialised biases are zero: make them matter
            if p.dim() == 1:
                p.data.normal_(0, 0.2, generator=gen)
        torch.manual_seed(2024)
        z, cg_z, xyz, cg_xyz, nbr_list, CG_nbr_list, mp, num_CGs = model.get_inputs(batch)
        H_prior_mu, H_prior_sigma = model.prior_net(cg_z, cg_xyz, CG_nbr_list)
        members = []
        for _ in range(n_ens):
            eps = torch.randn_like(H_prior_sigma)                       # sample_normal, sampling.py:247-250
            H = eps.mul(H_prior_sigma).add_(H_prior_mu)
            members.append(model.decoder(cg_xyz, CG_nbr_list, H, H, mp, num_CGs).detach())
        S_mu, S_sigma, pm, ps, xyz_out, xyz_recon = model(batch)        # sampling.py:293
        # replay the generator: same seed, same shapes, same order
        torch.manual_seed(2024)
        eps_members = torch.stack([torch.randn_like(H_prior_sigma) for _ in range(n_ens)])
        eps_recon = torch.randn_like(S_sigma)
        _state(tag, model, store)
        for k, t in batch.items():
            store["%s/batch/%s" % (tag, k)] = _np(t)
        for k, t in (("prior_mu", H_prior_mu), ("prior_sigma", H_prior_sigma), ("eps_members", eps_members),
                     ("members", torch.stack(members)), ("eps_recon", eps_recon), ("xyz_recon", xyz_recon),
                     ("mu", S_mu), ("sigma", S_sigma)):
            store["%s/%s" % (tag, k)] = _np(t)
        store["%s/meta" % tag] = np.array([F, R, enc, dec, atom_cutoff, cg_cutoff, int(breaksym), n_ens])
    _save("sampling_small.npz", store)


def dataset_lists(ref_data):
    """dataset_lists.npz: CGDataset.generate_neighbor_list (data.py:207-252) of the REAL reference on ragged frames: the
    radius branch (atom and CG graphs) and the bond-derived CG graph of ``cg_cutoff=None`` (:227-248)."""
    gen = torch.Generator().manual_seed(63)
    sizes = [9, 14, 5, 11]
    props = {"nxyz": [], "CG_nxyz": [], "CG_mapping": [], "num_atoms": [], "num_CGs": [], "bond_edge_list": []}
    for n in sizes:
        xyz = torch.randn(n, 3, generator=gen) * 1.7
        n_cg = max(2, n // 4)
        mapping = (torch.arange(n) * n_cg) // n
        cg_xyz = torch.stack([xyz[mapping == b].mean(0) for b in range(n_cg)])
        props["nxyz"].append(torch.cat([torch.ones(n, 1), xyz], 1))
        props["CG_nxyz"].append(torch.cat([torch.arange(n_cg).float()[:, None], cg_xyz], 1))
        props["CG_mapping"].append(mapping)
        props["num_atoms"].append(torch.tensor(n))
        props["num_CGs"].append(torch.tensor(n_cg))
        chain = torch.stack([torch.arange(n - 1), torch.arange(1, n)], 1)
        extra = torch.stack([torch.arange(n - 3), torch.arange(3, n)], 1)[::2]
        props["bond_edge_list"].append(torch.cat([chain, extra], 0))
    store = {"meta/sizes": np.array(sizes), "meta/atom_cutoff": 2.5, "meta/cg_cutoff": 4.0}
    for i in range(len(sizes)):
        for k in ("nxyz", "CG_nxyz", "CG_mapping", "num_atoms", "num_CGs", "bond_edge_list"):
            store["props/%d/%s" % (i, k)] = _np(props[k][i])
    for tag, cg_cut, und in (("radius", 4.0, True), ("radius_dir", 4.0, False), ("bond", None, True)):
        ds = ref_data.CGDataset({k: list(v) for k, v in props.items()})
        ds.generate_neighbor_list(atom_cutoff=2.5, cg_cutoff=cg_cut, device="cpu", undirected=und)
        for i in range(len(sizes)):
            store["%s/%d/nbr_list" % (tag, i)] = _np(ds.props["nbr_list"][i])
            store["%s/%d/CG_nbr_list" % (tag, i)] = _np(ds.props["CG_nbr_list"][i])
    _save("dataset_lists.npz", store)


def pcn_dihedral():
    """compute_dihe of the PCN loop (scripts/pcn_utils.py:114-132) and its dihedral loss term (:178-180), value and gradient.
    pcn_utils.py imports ase / networkx / sklearn at module level, so the function is lifted out of the UNMODIFIED source
    with ast and executed as it stands (its only globals: torch, EPS)."""
    import ast
    src_path = os.path.join(ref_shim.REFERENCE_ROOT, "scripts", "pcn_utils.py")
    tree = ast.parse(open(src_path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "compute_dihe"][0]
    eps = [n for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == "EPS"][0]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[eps, fn], type_ignores=[]), src_path, "exec"), ns)
    compute_dihe = ns["compute_dihe"]
    gen = torch.Generator().manual_seed(31)
    n_atoms, n_dihe = 64, 90
    xyz = torch.randn(n_atoms, 3, generator=gen) * 2.0
    rec = (xyz + 0.4 * torch.randn(n_atoms, 3, generator=gen)).requires_grad_(True)
    idx = torch.stack([torch.randperm(n_atoms, generator=gen)[:4] for _ in range(n_dihe)])
    gen_d, dat_d = compute_dihe(rec, idx), compute_dihe(xyz, idx)
    loss = (gen_d - dat_d).pow(2).mean()
    loss.backward()
    _save("pcn_dihedral.npz", {"xyz": _np(xyz), "xyz_rec": _np(rec), "idx": _np(idx), "gen_dihe": _np(gen_d), "data_dihe": _np(dat_d),
                               "loss": _np(loss), "g_xyz_rec": _np(rec.grad)})


def sample_quality():
    """eval_sample_qualities and its helpers (scripts/sampling.py:120-194,220-239,324-333) lifted out of the UNMODIFIED source
    with ast (the module imports mdshare / pyemma at the top) and run on a minimal ``Atoms`` stand-in (positions + numbers):
    bond adjacency of the reference conformation, per-sample difference counts / ratios and RMSDs."""
    import ast
    src_path = os.path.join(ref_shim.REFERENCE_ROOT, "scripts", "sampling.py")
    tree = ast.parse(open(src_path).read())
    want = {"compute_bond_cutoff", "compute_distance_mat", "dropH", "compare_graph", "get_bond_graphs", "count_valid_graphs",
            "compute_rmsd", "eval_sample_qualities"}
    body = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name in want)
            or (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == "COVCUTOFFTABLE")]

    class Atoms(object):
        def __init__(self, numbers, positions):
            self.numbers, self.positions = np.asarray(numbers), np.asarray(positions, dtype=np.float64)

        def get_positions(self):
            return self.positions

        def get_atomic_numbers(self):
            return self.numbers

        def __len__(self):
            return len(self.numbers)

    ns = {"torch": torch, "np": np, "Atoms": Atoms}
    exec(compile(ast.Module(body=body, type_ignores=[]), src_path, "exec"), ns)
    rng = np.random.default_rng(77)
    # a chain molecule with hydrogens: heavy atoms 1.5 A apart, one or two hydrogens at ~1.0 A
    n_heavy = 24
    heavy = np.cumsum(rng.normal(size=(n_heavy, 3)) * 0.2 + np.array([1.45, 0.3, 0.1]), 0)
    z, pos = [], []
    for i in range(n_heavy):
        z.append(int(rng.choice([6, 7, 8, 16])))
        pos.append(heavy[i])
        for _ in range(int(rng.integers(0, 3))):
            d = rng.normal(size=3)
            z.append(1)
            pos.append(heavy[i] + d / np.linalg.norm(d) * 1.0)
    z, pos = np.asarray(z), np.asarray(pos, dtype=np.float32).astype(np.float64)
    ref = Atoms(z, pos)
    samples = []
    for s, noise in enumerate([0.0, 0.01, 0.03, 0.08, 0.15, 0.3, 0.02, 0.0]):
        p = pos + rng.normal(size=pos.shape) * noise
        if s == 7:
            p = pos.copy()
            p[z == 1] += rng.normal(size=(int((z == 1).sum()), 3)) * 0.6      # hydrogens scrambled, heavy graph intact
        samples.append(np.asarray(p, dtype=np.float32).astype(np.float64))
    atoms_list = [Atoms(z, p) for p in samples]
    all_rmsds, heavy_rmsds, valid_ratio, valid_allatom_ratio, graph_val_ratio, graph_allatom_val_ratio = ns["eval_sample_qualities"](ref, atoms_list)
    store = {"z": z, "ref_xyz": pos.astype(np.float32), "samples": np.stack(samples).astype(np.float32),
             "ref_bonds": ns["get_bond_graphs"](ref).numpy(), "sample3_bonds": ns["get_bond_graphs"](atoms_list[3]).numpy(),
             "diff_counts": np.asarray([ns["compare_graph"](ref, a, 1.3) for a in atoms_list]),
             "all_rmsds": np.zeros((0, 2)) if all_rmsds is None else all_rmsds,
             "heavy_rmsds": np.zeros((0, 2)) if heavy_rmsds is None else heavy_rmsds,
             "valid_ratio": np.asarray(valid_ratio), "valid_allatom_ratio": np.asarray(valid_allatom_ratio),
             "graph_val_ratio": np.asarray(graph_val_ratio), "graph_allatom_val_ratio": np.asarray(graph_allatom_val_ratio)}
    print("valid heavy %.3f all-atom %.3f" % (valid_ratio, valid_allatom_ratio), store["diff_counts"])
    _save("sample_quality.npz", store)


def main():
    torch.set_num_threads(1)
    ref_modules, ref_conv, ref_cgvae, ref_data = ref_shim.import_reference()
    blocks(ref_conv)
    blocks_activations(ref_conv)
    cgvae_small(ref_cgvae, ref_data)
    cgvae_small(ref_cgvae, ref_data, fname="cgvae_noneq.npz", cases=(("vae_noneq", True, False),), seed=22)
    pcn_small(ref_cgvae, ref_data)
    graphs(ref_data, ref_cgvae)
    sampling(ref_cgvae, ref_data)
    dataset_lists(ref_data)
    pcn_dihedral()
    sample_quality()


if __name__ == "__main__":
    main()
