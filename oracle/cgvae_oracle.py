"""Functional CPU restatement of the CGVAE hot path (TEST INFRASTRUCTURE, see package docstring).

Everything here is a pure function of ``(P, inputs)`` where ``P`` is a flat
``{state_dict key: tensor}`` mapping laid out exactly like the reference's
``state_dict``, so weights can be exchanged with the product
modules and with the reference itself.  Tensors keep the reference layout:
scalars ``[N, F]``, vectors ``[N, F, 3]`` (xyz innermost), edge lists int64
``[E, 2]`` with column 0 = receiver i, column 1 = sender j.

Works in fp32 (parity / CPU baseline) and fp64 (finite-difference checks).
Citations are into /root/reference/CoarseGrainingVAE unless a path is given.
"""
import math

import numpy as np
import torch

# ----------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------

_ACT = {
    # modules.py:32-42 (layer_types) -- only element-wise activations are listed
    "swish": lambda x: x * torch.sigmoid(x),                 # modules.py:16-21
    "ReLU": torch.relu,
    "Tanh": torch.tanh,
    "sigmoid": torch.sigmoid,
    "shifted_softplus": lambda x: torch.nn.functional.softplus(x) - math.log(2.0),  # modules.py:9-14
    "LeakyReLU": torch.nn.functional.leaky_relu,
    "ELU": torch.nn.functional.elu,
}


def activation(name):
    return _ACT[name]


def affine(P, key, x, act=None):
    """``Dense`` = Linear (+dropout p=0) (+activation): modules.py:75-114."""
    y = x @ P[key + ".weight"].t()
    b = P.get(key + ".bias")
    if b is not None:
        y = y + b
    return act(y) if act is not None else y


def make_directed(nbr_list):
    """conv.py:10-20: append the flipped list unless both orientations occur."""
    up = bool((nbr_list[:, 0] > nbr_list[:, 1]).any())
    down = bool((nbr_list[:, 1] > nbr_list[:, 0]).any())
    if up and down:
        return nbr_list, True
    return torch.cat([nbr_list, nbr_list.flip(1)], dim=0), False


def rbf_coefficients(n_rbf, cutoff, dtype=torch.float32):
    """modules.py:145,156: ``n * pi / cutoff`` evaluated in the tensor dtype."""
    n = torch.arange(1, n_rbf + 1).to(dtype)
    return n * np.pi / cutoff


def edge_geometry(r, n_rbf, cutoff):
    """conv.py:25-29 (eps 1e-8 per component), modules.py:148-172, modules.py:52-58.

    returns d [E], unit [E,3], rbf [E,R], env [E]
    """
    d = ((r ** 2 + 1e-8).sum(-1)) ** 0.5
    unit = r / d.reshape(-1, 1)
    dd = d.unsqueeze(-1)
    coef = rbf_coefficients(n_rbf, cutoff, r.dtype)
    one = torch.ones((), dtype=r.dtype)
    zero = torch.zeros((), dtype=r.dtype)
    denom = torch.where(dd == 0, one, dd)
    num = torch.where(dd == 0, coef, torch.sin(coef * dd))
    rbf = torch.where(dd >= cutoff, zero, num / denom)
    env = 0.5 * (torch.cos(np.pi * d / cutoff) + 1)
    env = torch.where(d >= cutoff, zero, env)
    return d, unit, rbf, env


def distance_embed(P, key, d_rbf, env):
    """modules.py:192-197: (Dense(rbf)) * envelope, bias inside the envelope."""
    return affine(P, key + ".block.1", d_rbf) * env.reshape(-1, 1)


def segment_sum(src, index, size):
    """torch_scatter 2.0.9 ``scatter_add(dim=0)`` (see oracle/ref_shim.py)."""
    out = torch.zeros((size,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, index, src)


def segment_mean(src, index, size):
    """torch_scatter 2.0.9 ``scatter_mean(dim=0)``: sum / count.clamp(min=1)."""
    total = segment_sum(src, index, size)
    count = torch.zeros(size, dtype=src.dtype).index_add_(
        0, index, torch.ones(index.shape[0], dtype=src.dtype)).clamp_(min=1)
    return total / count.view((-1,) + (1,) * (src.dim() - 1))


def cross(a, b):
    """conv.py:213,218-219,379 ``torch.cross`` on [E,F,3] == cross over the xyz axis
    (unless E == 3 or F == 3, a reference quirk that is not reproduced)."""
    return torch.linalg.cross(a, b, dim=-1)


# ----------------------------------------------------------------------------
# blocks (conv.py)
# ----------------------------------------------------------------------------

def invariant_message(P, key, s, r, nbrs, n_rbf, cutoff, act, n_split):
    """conv.py:63-75: phi(s)[sender] * DistanceEmbed(dist) -> [E, n_split, F] + unit."""
    _, unit, rbf, env = edge_geometry(r, n_rbf, cutoff)
    phi = affine(P, key + ".inv_dense.1",
                 affine(P, key + ".inv_dense.0", s, act))
    w = distance_embed(P, key + ".dist_embed", rbf, env)
    m = phi[nbrs[:, 1]] * w
    return m.reshape(m.shape[0], n_split, s.shape[-1]), unit


def equi_message(P, key, s, v, r, nbrs, n_rbf, cutoff, act="swish", edge_wgt=None):
    """``EquiMessageBlock.forward`` conv.py:505-563 (3 filter splits) -> (ds, dv)."""
    m, unit = invariant_message(P, key + ".inv_message", s, r, nbrs, n_rbf, cutoff,
                                activation(act), 3)
    i, j = nbrs[:, 0], nbrs[:, 1]
    dv_e = m[:, 2, :, None] * unit[:, None, :] + m[:, 0, :, None] * v[j]
    ds_e = m[:, 1, :]
    if edge_wgt is not None:                                  # conv.py:528-533
        dv_e = dv_e * edge_wgt[:, None, None]
        ds_e = ds_e * edge_wgt[:, None]
    n = s.shape[0]
    return segment_sum(ds_e, i, n), segment_sum(dv_e, i, n)


def equi_message_cross(P, key, s, v, r, nbrs, n_rbf, cutoff, act="swish", edge_wgt=None):
    """``EquiMessageCross.forward`` conv.py:358-402 (4 splits) -> (ds, dv)."""
    m, unit = invariant_message(P, key + ".inv_message", s, r, nbrs, n_rbf, cutoff,
                                activation(act), 4)
    i, j = nbrs[:, 0], nbrs[:, 1]
    dv_e = (m[:, 2, :, None] * unit[:, None, :] + m[:, 0, :, None] * v[j]
            + m[:, 3, :, None] * cross(v[i], v[j]))
    ds_e = m[:, 1, :]
    if edge_wgt is not None:
        dv_e = dv_e * edge_wgt[:, None, None]
        ds_e = ds_e * edge_wgt[:, None]
    n = s.shape[0]
    return segment_sum(ds_e, i, n), segment_sum(dv_e, i, n)


def equi_message_pseudo(P, key, s, sbar, v, vbar, r, nbrs, n_rbf, cutoff, act="swish"):
    """``EquiMessagePsuedo.forward`` conv.py:180-242 (9 splits) -> (ds, dsbar, dv, dvbar)."""
    m, unit = invariant_message(P, key + ".inv_message", s, r, nbrs, n_rbf, cutoff,
                                activation(act), 9)
    i, j = nbrs[:, 0], nbrs[:, 1]
    sp = [m[:, k, :, None] for k in range(9)]
    ds_e = m[:, 0, :] * s[i]
    dsbar_e = (v[i] * vbar[j]).sum(-1)
    sb_i = sbar[i][:, :, None]
    dv_e = (sp[1] * unit[:, None, :] + sp[2] * v[j] + sp[3] * cross(v[i], vbar[j])
            + sp[4] * sb_i * vbar[j])
    dvbar_e = (sp[5] * vbar[j] + sp[6] * sb_i * v[j] + sp[7] * cross(v[i], v[j])
               + sp[8] * cross(vbar[i], vbar[j]))
    n = s.shape[0]
    return (segment_sum(ds_e, i, n), segment_sum(dsbar_e, i, n),
            segment_sum(dv_e, i, n), segment_sum(dvbar_e, i, n))


def update_block(P, key, s, v, act="swish"):
    """``UpdateBlock.forward`` conv.py:588-616 -> (ds, dv)."""
    n, f = s.shape
    vt = v.transpose(1, 2).reshape(-1, f)                      # [3N, F], rows (n, xyz)
    uv = (vt @ P[key + ".u_mat.weight"].t()).reshape(n, 3, f).transpose(1, 2)
    vv = (vt @ P[key + ".v_mat.weight"].t()).reshape(n, 3, f).transpose(1, 2)
    nrm = ((vv ** 2 + 1e-10).sum(-1)) ** 0.5
    x = torch.cat([s, nrm], dim=-1)
    q = affine(P, key + ".s_dense.1", affine(P, key + ".s_dense.0", x, activation(act)))
    q = q.reshape(n, 3, f)
    dv = uv * q[:, 0, :, None]
    ds = (uv * vv).sum(-1) * q[:, 1, :] + q[:, 2, :]
    return ds, dv


def contractive_message(P, key, s, v, r_iI, mapping, n_rbf, cutoff=20.0, act="swish",
                        n_beads=None):
    """``ContractiveMessageBlock.forward`` conv.py:703-733 -> (dS [Nc,F], dV [Nc,F,3])."""
    _, unit, rbf, env = edge_geometry(r_iI, n_rbf, cutoff)
    a = activation(act)
    phi = affine(P, key + ".inv_dense.1", affine(P, key + ".inv_dense.0", s, a))
    m = (phi * distance_embed(P, key + ".dist_embed", rbf, env)).reshape(s.shape[0], 3, -1)
    dv_n = m[:, 2, :, None] * unit[:, None, :] + m[:, 0, :, None] * v
    ds_n = m[:, 1, :]
    if n_beads is None:
        n_beads = int(mapping.max()) + 1                        # scatter_add without dim_size
    return segment_sum(ds_n, mapping, n_beads), segment_sum(dv_n, mapping, n_beads)


# ----------------------------------------------------------------------------
# stacks and models (cgvae.py)
# ----------------------------------------------------------------------------

def encoder_forward(P, key, spec, z, xyz, cg_xyz, mapping, nbr_list, cg_nbr_list):
    """``EquiEncoder.forward`` cgvae.py:266-331 -> (H [Nc,F], h [N,F])."""
    if not spec.get("dir_mp", False):
        nbr_list, _ = make_directed(nbr_list)
    h = torch.nn.functional.embedding(z.long(), P[key + ".atom_embed.weight"], padding_idx=0)
    v = torch.zeros(h.shape[0], h.shape[1], 3, dtype=h.dtype)
    r_ij = xyz[nbr_list[:, 1]] - xyz[nbr_list[:, 0]]
    r_iI = xyz - cg_xyz[mapping]
    n_beads = int(mapping.max()) + 1
    H = V = None
    for l in range(spec["enc_nconv"]):
        ds, dv = equi_message(P, "%s.message_blocks.%d" % (key, l), h, v, r_ij, nbr_list,
                              spec["n_rbf"], spec["cg_cutoff"], spec["activation"])
        h = h + ds
        v = v + dv
        if l == 0:                                             # cgvae.py:296-298
            H = segment_mean(h, mapping, n_beads)
            V = segment_mean(v, mapping, n_beads)
        dH, dV = contractive_message(P, "%s.cgmessage_layers.%d" % (key, l), h, v, r_iI,
                                     mapping, spec["n_rbf"], 20.0, spec["activation"], n_beads)
        H = H + dH
        V = V + dV
    return H, h


def prior_forward(P, key, spec, cg_z, cg_xyz, cg_nbr_list):
    """``CGprior.forward`` cgvae.py:374-403 -> (mu, std)."""
    cg_nbr_list, _ = make_directed(cg_nbr_list)
    h = torch.nn.functional.embedding(cg_z.long(), P[key + ".atom_embed.weight"], padding_idx=0)
    v = torch.zeros(h.shape[0], h.shape[1], 3, dtype=h.dtype)
    r = cg_xyz[cg_nbr_list[:, 1]] - cg_xyz[cg_nbr_list[:, 0]]
    for l in range(spec["enc_nconv"]):
        ds, dv = equi_message(P, "%s.message_blocks.%d" % (key, l), h, v, r, cg_nbr_list,
                              spec["n_rbf"], spec["cg_cutoff"], spec["activation"])
        h = h + ds
        v = v + dv
    mu = affine(P, key + ".mu.2", affine(P, key + ".mu.0", h, torch.tanh))
    logvar = affine(P, key + ".sigma.2", affine(P, key + ".sigma.0", h, torch.tanh))
    return mu, 1e-9 + torch.exp(logvar / 2)


def decoder_stack_forward(P, key, spec, cg_xyz, cg_nbr_list, S):
    """``EquivariantPsuedoDecoder.forward`` cgvae.py:85-125 (decoder == 'pseudo') or
    ``EquivariantDecoder.forward`` cgvae.py:163-191 ('cross' / 'plain') -> (S, V)."""
    cg_nbr_list, _ = make_directed(cg_nbr_list)
    r = cg_xyz[cg_nbr_list[:, 1]] - cg_xyz[cg_nbr_list[:, 0]]
    n, f = S.shape
    V = torch.zeros(n, f, 3, dtype=S.dtype)
    kind = spec["decoder"]
    R, cut, act = spec["n_rbf"], spec["atom_cutoff"], spec["activation"]
    if kind == "pseudo":
        Sbar = (torch.ones if spec.get("breaksym", False) else torch.zeros)(n, f, dtype=S.dtype)
        Vbar = torch.zeros(n, f, 3, dtype=S.dtype)
    for l in range(spec["dec_nconv"]):
        mkey = "%s.message_blocks.%d" % (key, l)
        if kind == "pseudo":
            dS, dSbar, dV, dVbar = equi_message_pseudo(P, mkey, S, Sbar, V, Vbar, r,
                                                       cg_nbr_list, R, cut, act)
            S, Sbar, V, Vbar = S + dS, Sbar + dSbar, V + dV, Vbar + dVbar
        elif kind == "cross":
            dS, dV = equi_message_cross(P, mkey, S, V, r, cg_nbr_list, R, cut, act)
            S, V = S + dS, V + dV
        else:
            dS, dV = equi_message(P, mkey, S, V, r, cg_nbr_list, R, cut, act)
            S, V = S + dS, V + dV
        dS, dV = update_block(P, "%s.update_blocks.%d" % (key, l), S, V, act)
        S, V = S + dS, V + dV
    return S, V


def channel_index(mapping):
    """``CG2ChannelIdx`` cgvae.py:451-460: rank of each atom among the atoms of its bead."""
    m = np.asarray(mapping, dtype=np.int64)
    out = np.zeros_like(m)
    seen = {}
    for n, b in enumerate(m.tolist()):
        out[n] = seen.get(b, 0)
        seen[b] = out[n] + 1
    return torch.from_numpy(out)


def lift(V, cg_xyz, mapping, offset=True, ca_idx=None):
    """cgvae.py:466-482 (CGequiVAE.decoder tail) / cgvae.py:556-576 (PCN.decoder tail)."""
    chan = channel_index(mapping)
    rel = V[mapping, chan, :]
    if ca_idx is not None:                                     # PCN: pin C-alpha to its bead
        # cgvae.py:569-571: rel[ca] -= clone(rel[ca])  ==  value 0 and gradient 0 there
        if int(ca_idx[-1]) < rel.shape[0]:
            rel = rel.index_fill(0, ca_idx, 0.0)
    elif offset:
        n_beads = int(mapping.max()) + 1
        rel = rel - segment_mean(rel, mapping, n_beads)[mapping]
    return rel + cg_xyz[mapping]


def cgvae_forward(P, spec, batch, eps=None):
    """``CGequiVAE.forward`` cgvae.py:486-513.  ``eps`` replaces ``torch.randn_like``
    (cgvae.py:446) so both sides of a parity test see the same noise."""
    nxyz, cg_nxyz = batch["nxyz"], batch["CG_nxyz"]
    z, xyz = nxyz[:, 0], nxyz[:, 1:]
    cg_z, cg_xyz = cg_nxyz[:, 0], cg_nxyz[:, 1:]
    mapping = batch["CG_mapping"]
    H, h = encoder_forward(P, "encoder", spec, z, xyz, cg_xyz, mapping,
                           batch["nbr_list"], batch["CG_nbr_list"])
    pmu, pstd = prior_forward(P, "prior_net", spec, cg_z, cg_xyz, batch["CG_nbr_list"])
    mu = affine(P, "atom_munet.2", affine(P, "atom_munet.0", H, torch.relu))
    logvar = affine(P, "atom_sigmanet.2", affine(P, "atom_sigmanet.0", H, torch.relu))
    sigma = 1e-12 + torch.exp(logvar / 2)
    if spec.get("det", False):
        zs = H
    else:
        if eps is None:
            eps = torch.randn_like(sigma)
        zs = eps * sigma + mu
    S, V = decoder_stack_forward(P, "equivaraintconv", spec, cg_xyz, batch["CG_nbr_list"], zs)
    if spec.get("equivariant", True) is False:
        # cgvae.py:469-471: the vectors come from a plain Linear(F, 3F) on the scalars instead of the equivariant channel
        V = affine(P, "euclidean", S).reshape(S.shape[0], S.shape[1], 3)
    xyz_recon = lift(V, cg_xyz, mapping, offset=spec.get("offset", True))
    return mu, sigma, pmu, pstd, xyz, xyz_recon


def cgvae_decode(P, spec, cg_xyz, cg_nbr_list, S, mapping):
    """``CGequiVAE.decoder`` cgvae.py:462-484: decoder stack + lifting."""
    S2, V = decoder_stack_forward(P, "equivaraintconv", spec, cg_xyz, cg_nbr_list, S)
    if spec.get("equivariant", True) is False:
        V = affine(P, "euclidean", S2).reshape(S2.shape[0], S2.shape[1], 3)
    return lift(V, cg_xyz, mapping, offset=spec.get("offset", True))


def sample_ensemble(P, spec, batch, eps_members):
    """The ensemble loop of ``sample_single`` scripts/sampling.py:265-284: the prior once (:268), then per member
    ``H = eps*sigma + mu`` (``sample_normal`` :247-250) and ``model.decoder(cg_xyz, CG_nbr_list, H, H, mapping, num_CGs)``
    (:277-279).  ``eps_members [n_ensemble, Nc, F]`` replaces the ``torch.randn_like`` draws.
    Returns (prior_mu, prior_sigma, xyz [n_ensemble, N, 3])."""
    cg_nxyz = batch["CG_nxyz"]
    cg_z, cg_xyz = cg_nxyz[:, 0], cg_nxyz[:, 1:]
    mu, sigma = prior_forward(P, "prior_net", spec, cg_z, cg_xyz, batch["CG_nbr_list"])
    out = []
    for eps in eps_members:
        H = eps * sigma + mu
        out.append(cgvae_decode(P, spec, cg_xyz, batch["CG_nbr_list"], H, batch["CG_mapping"]))
    return mu, sigma, torch.stack(out)


def train_loop_step(loss, gamma, train=True):
    """The control flow of one iteration of ``loop`` scripts/utils.py:145-160 as data: returns (backward, optimiser_step).
    ``loss >= gamma*200`` or NaN -> ``continue`` (no backward, no step); validation still runs ``loss.backward()``."""
    l = float(loss)
    if l >= gamma * 200.0 or l != l:
        return False, False
    return True, bool(train)


def pcn_forward(P, spec, batch):
    """``PCN.forward`` cgvae.py:588-594."""
    S = torch.nn.functional.embedding(batch["res"].long(), P["embedding.weight"], padding_idx=0)
    S2, V = decoder_stack_forward(P, "equivaraintconv", spec, batch["ca_xyz"],
                                  batch["CG_nbr_list"], S)
    xyz_recon = lift(V, batch["ca_xyz"], batch["cg_map"], offset=False, ca_idx=batch["ca_idx"])
    return None, None, None, None, batch["xyz"], xyz_recon


# ----------------------------------------------------------------------------
# losses (/root/reference/scripts/utils.py)
# ----------------------------------------------------------------------------

def kl_divergence(mu1, std1, mu2, std2):
    """scripts/utils.py:81-86 (the ``/ std2`` -- not ``std2**2`` -- quirk is kept); ``mu2 is None``: KL against the
    standard normal (:82-83)."""
    if mu2 is None:
        return -0.5 * torch.sum(1 + torch.log(std1.pow(2)) - mu1.pow(2) - std1.pow(2), dim=-1).mean()
    return 0.5 * ((std1.pow(2) / std2.pow(2)).sum(-1) + ((mu1 - mu2).pow(2) / std2).sum(-1)
                  + torch.log(std2.pow(2)).sum(-1) - torch.log(std1.pow(2)).sum(-1)
                  - std1.shape[-1]).mean()


def graph_loss(xyz_recon, xyz, edge_list, eps=1e-6):
    """scripts/utils.py:130-133 (EPS = 1e-6, scripts/utils.py:27)."""
    a, b = edge_list[:, 0], edge_list[:, 1]
    gen = ((xyz_recon[a] - xyz_recon[b]).pow(2).sum(-1) + eps).sqrt()
    dat = ((xyz[a] - xyz[b]).pow(2).sum(-1) + eps).sqrt()
    return (gen - dat).pow(2).mean()


def training_loss(outputs, batch, beta, gamma):
    """scripts/utils.py:117-141: recon + beta*KL + gamma*graph."""
    mu, sigma, pmu, pstd, xyz, xyz_recon = outputs
    recon = (xyz_recon - xyz).pow(2).mean()
    kl = kl_divergence(mu, sigma, pmu, pstd) if mu is not None else torch.zeros(())
    g = graph_loss(xyz_recon, xyz, batch["bond_edge_list"]) if gamma != 0.0 else torch.zeros(())
    return recon + kl * beta + g * gamma, recon, kl, g


EPS = 1e-6      # scripts/pcn_utils.py:21, scripts/utils.py:27


def compute_dihe(xyz, indices):
    """scripts/pcn_utils.py:114-132: theta = arctan(p1 / (p2 + EPS)) with b1..b3 the three bond vectors of the quadruple,
    c1 = b2 x b3, c2 = b1 x b2, p1 = (b1 . c1) * sqrt(b2 . b2 + EPS), p2 = c1 . c2 (EPS = 1e-6, pcn_utils.py:21)."""
    b1 = xyz[indices[:, 1]] - xyz[indices[:, 0]]
    b2 = xyz[indices[:, 2]] - xyz[indices[:, 1]]
    b3 = xyz[indices[:, 3]] - xyz[indices[:, 2]]
    c1 = torch.linalg.cross(b2, b3)
    c2 = torch.linalg.cross(b1, b2)
    p1 = (b1 * c1).sum(-1) * ((b2 * b2).sum(-1) + EPS) ** 0.5
    p2 = (c1 * c2).sum(-1)
    return torch.arctan(p1 / (p2 + EPS))


def pcn_loss(xyz_recon, xyz, bond_edge_list, dihe_idxs, gamma, kappa):
    """the loss of the PCN loop, scripts/pcn_utils.py:160-183 (no KL: PCN.forward returns S_mu = None):
    mean((rec - xyz)^2) + gamma * mean over bonds (|d_rec|_e - |d_xyz|_e)^2 + kappa * mean over dihedrals (theta_rec - theta_xyz)^2."""
    recon = (xyz_recon - xyz).pow(2).mean()
    a, b = bond_edge_list[:, 0], bond_edge_list[:, 1]
    gen = ((xyz_recon[a] - xyz_recon[b]).pow(2).sum(-1) + EPS).sqrt()
    dat = ((xyz[a] - xyz[b]).pow(2).sum(-1) + EPS).sqrt()
    graph = (gen - dat).pow(2).mean() if gamma != 0.0 else torch.zeros((), dtype=xyz.dtype)
    dihe = (compute_dihe(xyz_recon, dihe_idxs) - compute_dihe(xyz, dihe_idxs)).pow(2).mean()
    return recon + graph * gamma + dihe * kappa, recon, graph, dihe


def pcn_loop_step(loss, gamma, train=True):
    """scripts/pcn_utils.py:191-206: ``loss >= gamma * 300`` or NaN -> ``continue``; validation still runs ``loss.backward()``."""
    l = float(loss)
    if l >= gamma * 300.0 or l != l:
        return False, False
    return True, bool(train)
