"""TEST-ONLY stand-in for the C-ABI kernels, used to exercise the host-side logic on CPU.

The product has no CPU path.  This module lives under tests/ and is installed by monkeypatching
``coarsegrainingvae_b200.ops`` inside a pytest fixture; it restates, in plain torch, what each
KERNEL computes -- including the hand-derived backward formulas in the same
algebraic form the CUDA code uses (B-accumulator form of the filter gradient, q-form of the cross
term, receiver- and sender-side terms of the 9-split block).  Running the real Python modules /
autograd nodes on top of it in float64 and comparing with the oracle's autograd validates

  * the backward derivations the kernels implement,
  * the autograd wiring, layouts, residual fusion and state_dict handling of the host code.

The CUDA kernels themselves are validated separately on the GPU (tests marked ``gpu``).
"""
import numpy as np
import torch

from coarsegrainingvae_b200 import ops


def _act(code, z):
    import math
    F = torch.nn.functional
    return {0: lambda t: t, 1: lambda t: t * torch.sigmoid(t), 2: torch.relu, 3: torch.tanh, 4: torch.sigmoid,
            5: lambda t: F.softplus(t) - math.log(2.0), 6: F.leaky_relu, 7: F.elu}[code](z)


def _dact(code, z):
    if code == 1:
        sig = torch.sigmoid(z)
        return sig * (1 + z * (1 - sig))
    if code == 2:
        return (z > 0).to(z.dtype)
    if code == 3:
        return 1 - torch.tanh(z) ** 2
    if code == 4:
        sig = torch.sigmoid(z)
        return sig * (1 - sig)
    if code == 5:
        return torch.sigmoid(z)
    if code == 6:
        return torch.where(z > 0, torch.ones_like(z), torch.full_like(z, 0.01))
    if code == 7:
        return torch.where(z > 0, torch.ones_like(z), torch.exp(z))
    return torch.ones_like(z)


def radius_graph(xyz, cutoff, undirected=True, frame_ptr=None, use_cells=None):
    from oracle import graph_oracle
    x = xyz.detach().cpu().numpy()
    if frame_ptr is None:
        return torch.from_numpy(graph_oracle.radius_graph(x, cutoff, undirected))
    fp = frame_ptr.cpu().numpy()
    parts = [graph_oracle.radius_graph(x[fp[k]:fp[k + 1]], cutoff, undirected) + fp[k] for k in range(len(fp) - 1)]
    return torch.from_numpy(np.concatenate(parts, 0)) if parts else torch.zeros((0, 2), dtype=torch.int64)


def edge_orientation(pairs):
    return bool((pairs[:, 0] > pairs[:, 1]).any()), bool((pairs[:, 1] > pairs[:, 0]).any())


def build_graph(pairs, n_recv, n_send=None, symmetrize=False, n_edges_dev=None):
    if n_send is None:
        n_send = n_recv
    pairs = pairs.to(torch.int64)
    if n_edges_dev is not None:
        pairs = pairs[:int(n_edges_dev)]
    if symmetrize:
        pairs = torch.cat([pairs, pairs.flip(1)], 0)
    E = pairs.shape[0]
    eid = torch.argsort(pairs[:, 0], stable=True)
    rowptr = torch.zeros(n_recv + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(pairs[:, 0], minlength=n_recv), 0)
    col = pairs[eid, 1]
    slot_of_edge = torch.empty(E, dtype=torch.int64)
    slot_of_edge[eid] = torch.arange(E)
    eid_s = torch.argsort(pairs[:, 1], stable=True)
    rowptr_t = torch.zeros(n_send + 1, dtype=torch.int64)
    rowptr_t[1:] = torch.cumsum(torch.bincount(pairs[:, 1], minlength=n_send), 0)
    return ops.Graph(n_recv, n_send, E, rowptr, col, eid, rowptr_t, pairs[eid_s, 0], slot_of_edge[eid_s])


def build_segments(mapping, n_beads):
    mapping = mapping.to(torch.int64)
    n = mapping.shape[0]
    atoms = torch.argsort(mapping, stable=True)
    rowptr = torch.zeros(n_beads + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(mapping, minlength=n_beads), 0)
    slot = torch.empty(n, dtype=torch.int64)
    slot[atoms] = torch.arange(n)
    rank = slot - rowptr[mapping]
    return ops.Segments(n_beads, mapping, rowptr, atoms, slot, rank)


def contraction_graph(seg):
    n = seg.n
    return ops.Graph(seg.n_beads, n, n, seg.rowptr, seg.atoms, seg.atoms, torch.arange(n + 1), seg.mapping, seg.slot)


def _recv_of_slot(graph):
    counts = graph.rowptr[1:] - graph.rowptr[:-1]
    return torch.repeat_interleave(torch.arange(graph.n_recv), counts)


def edge_geometry(graph, xyz_send, xyz_recv, n_rbf, cutoff, edge_wgt=None, r_edge=None):
    if r_edge is not None:
        r = r_edge[graph.eid]
    else:
        r = xyz_send[graph.col] - xyz_recv[_recv_of_slot(graph)]
    dt = r.dtype
    d = ((r ** 2 + 1e-8).sum(-1)) ** 0.5
    unit = r / d[:, None]
    coef = (torch.arange(1, n_rbf + 1).to(dt) * np.pi / cutoff)
    inside = (d < cutoff).to(dt)
    env = 0.5 * (torch.cos(np.pi * d / cutoff) + 1) * inside
    rbf = torch.sin(coef[None, :] * d[:, None]) / d[:, None] * inside[:, None]
    rb = ops.rb_for(n_rbf)
    basis = torch.zeros(r.shape[0], rb, dtype=dt)
    basis[:, :n_rbf] = rbf * env[:, None]
    basis[:, n_rbf] = env
    if edge_wgt is not None:
        basis = basis * edge_wgt[graph.eid][:, None]
    return ops.Geometry(graph, basis, torch.cat([unit, d[:, None]], 1), n_rbf, rb, cutoff)


def gemm(form, A, B, M, N, K, bias=None, act=0, z_out=False, z_in=None, dact=0, add=None, out=None):
    if form == ops.GEMM_NT:
        C = A[:M, :K] @ B[:N, :K].t()
    elif form == ops.GEMM_NN:
        C = A[:M, :K] @ B[:K, :N]
    else:
        C = A[:K, :M].t() @ B[:K, :N]
    if bias is not None:
        C = C + bias
    Z = C.clone() if z_out else None
    C = _act(act, C)
    if z_in is not None:
        C = C * _dact(dact, z_in)
    if add is not None:
        C = C + add
    if out is not None:
        out.copy_(C)
        C = out
    return (C, Z) if z_out else C


def colsum(X, param=None):
    return X.sum(0)


def _filter_matrix(Wf, bf, K, F, rb):
    """[K, RB, F]: rows r < R from Wf[(kF+f), r], row R = bias, rest zero -- filter_entry() of message.cu."""
    R = Wf.shape[1]
    W = torch.zeros(K, rb, F, dtype=Wf.dtype)
    W[:, :R, :] = Wf.view(K, F, R).permute(0, 2, 1)
    W[:, R, :] = bf.view(K, F)
    return W


def _cross(a, b):
    return torch.linalg.cross(a, b, dim=1)          # planar [*, 3, F]


def message_fwd(n_split, phi, v_send, v_recv, geom, Wf, bf, res_s, res_v, want_q=False):
    g = geom.graph
    F = phi.shape[-1]
    W = _filter_matrix(Wf, bf, n_split, F, geom.rb)
    w = torch.einsum("er,krf->ekf", geom.basis, W)
    i, j = _recv_of_slot(g), g.col
    m = phi[j] * w
    u = geom.unit[:, :3]
    out_s = torch.zeros(g.n_recv, F, dtype=phi.dtype).index_add_(0, i, m[:, 1])
    dv = m[:, 2, None, :] * u[:, :, None]
    q = None
    if v_send is not None:
        dv = dv + m[:, 0, None, :] * v_send[j]
    out_v = torch.zeros(g.n_recv, 3, F, dtype=phi.dtype).index_add_(0, i, dv)
    if n_split == 4 and v_send is not None:
        q = torch.zeros(g.n_recv, 3, F, dtype=phi.dtype).index_add_(0, i, m[:, 3, None, :] * v_send[j])
        out_v = out_v + _cross(v_recv, q)
    if res_s is not None:
        out_s = out_s + res_s
    if res_v is not None:
        out_v = out_v + res_v
    return out_s, out_v, (q if want_q else None)


def message_bwd(n_split, phi, v_send, v_recv, q, geom, Wf, bf, g_out_s, g_out_v, residual, sink=True):
    g = geom.graph
    F = phi.shape[-1]
    K = n_split
    W = _filter_matrix(Wf, bf, K, F, geom.rb)
    # iterate the SENDER csr like the kernel: slot t -> receiver col_t[t], edge data at perm_t[t]
    counts = g.rowptr_t[1:] - g.rowptr_t[:-1]
    j = torch.repeat_interleave(torch.arange(g.n_send), counts)
    i, e = g.col_t, g.perm_t
    b, u = geom.basis[e], geom.unit[e, :3]
    gs, gv = g_out_s[i], g_out_v[i]
    gm = torch.zeros(e.shape[0], K, F, dtype=phi.dtype)
    gm[:, 1] = gs
    gm[:, 2] = (gv * u[:, :, None]).sum(1)
    vj = v_send[j] if v_send is not None else torch.zeros_like(gv)
    gm[:, 0] = (gv * vj).sum(1)
    w0 = b @ W[0]
    m0 = phi[j, 0] * w0
    contrib = m0[:, None, :] * gv
    if K == 4:
        vi = v_recv[i] if v_send is not None else torch.zeros_like(gv)
        gq = _cross(gv, vi)
        gm[:, 3] = (gq * vj).sum(1)
        m3 = phi[j, 3] * (b @ W[3])
        contrib = contrib + m3[:, None, :] * gq
    g_v_send = torch.zeros(g.n_send, 3, F, dtype=phi.dtype).index_add_(0, j, contrib)
    B = torch.zeros(g.n_send, K, geom.rb, F, dtype=phi.dtype).index_add_(0, j, gm[:, :, None, :] * b[:, None, :, None])
    g_phi = (W[None] * B).sum(2)
    dW = (phi[:, :, None, :] * B).sum(0)                      # [K, RB, F]
    R = Wf.shape[1]
    dWf = dW[:, :R, :].permute(0, 2, 1).reshape(K * F, R)
    dbf = dW[:, R, :].reshape(K * F)
    if residual:
        g_v_send = g_v_send + g_out_v
    if K == 4 and q is not None and v_send is not None:
        g_v_send = g_v_send + _cross(q, g_out_v)
    return g_phi, g_v_send, dWf, dbf


def message9_fwd(phi, s, sbar, v, vbar, geom, Wf, bf, residual):
    g = geom.graph
    n, F = s.shape
    W = _filter_matrix(Wf, bf, 9, F, geom.rb)
    w = torch.einsum("er,krf->ekf", geom.basis, W)
    i, j = _recv_of_slot(g), g.col
    m = phi[j] * w
    u = geom.unit[:, :3]
    mk = lambda k: m[:, k, None, :]
    sb_i = sbar[i][:, None, :]
    ds = m[:, 0] * s[i]
    dsb = (v[i] * vbar[j]).sum(1)
    dv = mk(1) * u[:, :, None] + mk(2) * v[j] + mk(3) * _cross(v[i], vbar[j]) + mk(4) * sb_i * vbar[j]
    dvb = mk(5) * vbar[j] + mk(6) * sb_i * v[j] + mk(7) * _cross(v[i], v[j]) + mk(8) * _cross(vbar[i], vbar[j])
    z2 = lambda: torch.zeros(n, F, dtype=s.dtype)
    z3 = lambda: torch.zeros(n, 3, F, dtype=s.dtype)
    outs = [z2().index_add_(0, i, ds), z2().index_add_(0, i, dsb), z3().index_add_(0, i, dv), z3().index_add_(0, i, dvb)]
    if residual:
        outs = [outs[0] + s, outs[1] + sbar, outs[2] + v, outs[3] + vbar]
    return outs


def message9_bwd(phi, s, sbar, v, vbar, geom, Wf, bf, residual, g_s, g_sbar, g_v, g_vbar):
    g = geom.graph
    n, F = s.shape
    W = _filter_matrix(Wf, bf, 9, F, geom.rb)
    w_all = torch.einsum("er,krf->ekf", geom.basis, W)      # per receiver-csr slot
    dot = lambda a, b: (a * b).sum(1)
    # ---- sender role
    counts = g.rowptr_t[1:] - g.rowptr_t[:-1]
    j = torch.repeat_interleave(torch.arange(n), counts)
    i, e = g.col_t, g.perm_t
    w, u = w_all[e], geom.unit[e, :3]
    ph = phi[j]
    gm = torch.stack([
        g_s[i] * s[i],
        dot(g_v[i], u[:, :, None].expand(-1, -1, F)),
        dot(g_v[i], v[j]),
        dot(g_v[i], _cross(v[i], vbar[j])),
        sbar[i] * dot(g_v[i], vbar[j]),
        dot(g_vbar[i], vbar[j]),
        sbar[i] * dot(g_vbar[i], v[j]),
        dot(g_vbar[i], _cross(v[i], v[j])),
        dot(g_vbar[i], _cross(vbar[i], vbar[j])),
    ], 1)
    g_phi = torch.zeros(n, 9, F, dtype=s.dtype).index_add_(0, j, gm * w)
    gw = torch.zeros(g.n_edges, 9, F, dtype=s.dtype)
    gw[e] = gm * ph
    m = ph * w
    mk = lambda k: m[:, k, None, :]
    sb_i = sbar[i][:, None, :]
    d_v = mk(2) * g_v[i] + mk(6) * sb_i * g_vbar[i] + mk(7) * _cross(g_vbar[i], v[i])
    d_vb = (g_sbar[i][:, None, :] * v[i] + mk(3) * _cross(g_v[i], v[i]) + mk(4) * sb_i * g_v[i] + mk(5) * g_vbar[i]
            + mk(8) * _cross(g_vbar[i], vbar[i]))
    gi_v = torch.zeros(n, 3, F, dtype=s.dtype).index_add_(0, j, d_v)
    gi_vb = torch.zeros(n, 3, F, dtype=s.dtype).index_add_(0, j, d_vb)
    # ---- receiver role
    i2, j2 = _recv_of_slot(g), g.col
    m2 = phi[j2] * w_all
    mk2 = lambda k: m2[:, k, None, :]
    gi_s = torch.zeros(n, F, dtype=s.dtype).index_add_(0, i2, g_s[i2] * m2[:, 0])
    gi_sb = torch.zeros(n, F, dtype=s.dtype).index_add_(
        0, i2, m2[:, 4] * dot(g_v[i2], vbar[j2]) + m2[:, 6] * dot(g_vbar[i2], v[j2]))
    gi_v.index_add_(0, i2, g_sbar[i2][:, None, :] * vbar[j2] + mk2(3) * _cross(vbar[j2], g_v[i2])
                    + mk2(7) * _cross(v[j2], g_vbar[i2]))
    gi_vb.index_add_(0, i2, mk2(8) * _cross(vbar[j2], g_vbar[i2]))
    if residual:
        gi_s, gi_sb, gi_v, gi_vb = gi_s + g_s, gi_sb + g_sbar, gi_v + g_v, gi_vb + g_vbar
    R = geom.n_rbf
    gw2 = gw.reshape(g.n_edges, 9 * F)
    dWf = gw2.t() @ geom.basis[:, :R]
    dbf = gw2.t() @ geom.basis[:, R]
    return gi_s, gi_sb, gi_v, gi_vb, g_phi, dWf, dbf


def update_norm_fwd(s, Vv):
    return torch.cat([s, ((Vv ** 2 + 1e-10).sum(1)) ** 0.5], 1)


def update_combine_fwd(s, v, Uv, Vv, q, residual):
    ds = (Uv * Vv).sum(1) * q[:, 1] + q[:, 2]
    dv = Uv * q[:, 0, None, :]
    return (s + ds, v + dv) if residual else (ds, dv)


def update_combine_bwd(Uv, Vv, q, g_s, g_v, cat=False):
    inner = (Uv * Vv).sum(1)
    gq = torch.stack([(g_v * Uv).sum(1), g_s * inner, g_s], 1)
    t = (g_s * q[:, 1])[:, None, :]
    return gq, g_v * q[:, 0, None, :] + t * Vv, t * Uv


def update_norm_bwd(x, Vv, gx, g_s, gVv, residual):
    F = Vv.shape[-1]
    gVv += (gx[:, F:] / x[:, F:])[:, None, :] * Vv
    return gx[:, :F] + g_s if residual else gx[:, :F].clone()


def segment_reduce_fwd(X, seg, mean, param=None):
    out = torch.zeros((seg.n_beads,) + tuple(X.shape[1:]), dtype=X.dtype).index_add_(0, seg.mapping, X)
    if mean:
        cnt = (seg.rowptr[1:] - seg.rowptr[:-1]).clamp(min=1).to(X.dtype)
        out = out / cnt.view((-1,) + (1,) * (X.dim() - 1))
    return out


def segment_reduce_bwd(g_out, seg, mean):
    g = g_out[seg.mapping]
    if mean:
        cnt = (seg.rowptr[1:] - seg.rowptr[:-1]).clamp(min=1).to(g_out.dtype)
        g = g / cnt[seg.mapping].view((-1,) + (1,) * (g_out.dim() - 1))
    return g


def gather_rows(table, idx):
    return table[idx.to(torch.int64)]


def lift_fwd(V, cg_xyz, seg, mode, pin):
    rel = V[seg.mapping, :, seg.rank]
    if mode == 1:
        rel = rel - segment_reduce_fwd(rel, seg, True)[seg.mapping]
    elif mode == 2 and pin is not None:
        rel = rel * (pin == 0).to(rel.dtype)[:, None]
    return rel + cg_xyz[seg.mapping]


def lift_bwd(g_xyz, seg, F, mode, pin):
    g = g_xyz
    if mode == 1:
        g = g - segment_reduce_fwd(g, seg, True)[seg.mapping]
    elif mode == 2 and pin is not None:
        g = g * (pin == 0).to(g.dtype)[:, None]
    g_V = torch.zeros(seg.n_beads, 3, F, dtype=g_xyz.dtype)
    g_V[seg.mapping, :, seg.rank] = g
    return g_V


def vec_to_planar(v):
    return v.permute(0, 2, 1).contiguous()


def vec_from_planar(v):
    return v.permute(0, 2, 1).contiguous()


def fill(shape, value, like):
    return torch.full(shape, float(value), dtype=like.dtype, device=like.device)


def vae_latent_fwd(mu, logvar, eps):
    sigma = 1e-12 + torch.exp(logvar / 2)
    return (eps * sigma + mu if eps is not None else None), sigma


def vae_latent_bwd(g_z, g_sigma, eps, sigma):
    g_s = torch.zeros_like(sigma) if g_sigma is None else g_sigma.clone()
    if g_z is not None:
        g_s = g_s + g_z * eps
    return g_z, g_s * 0.5 * (sigma - 1e-12)


def pin_mask(idx, n_atoms, count=None):
    pin = torch.zeros(n_atoms, dtype=torch.uint8, device=idx.device)
    n = int(count) if count is not None else idx.numel()
    if n and int(idx[n - 1]) < n_atoms:
        pin[idx[:n]] = 1
    return pin


def std_logvar_fwd(x, c):
    return c + torch.exp(x / 2)


def std_logvar_bwd(gy, y, c):
    return gy * 0.5 * (y - c)


_NAMES = ["fill", "pin_mask", "vae_latent_fwd", "vae_latent_bwd", "std_logvar_fwd", "std_logvar_bwd", "radius_graph", "edge_orientation", "build_graph", "build_segments", "contraction_graph", "edge_geometry",
          "gemm", "colsum", "message_fwd", "message_bwd", "message9_fwd", "message9_bwd", "update_norm_fwd",
          "update_combine_fwd", "update_combine_bwd", "update_norm_bwd", "segment_reduce_fwd", "segment_reduce_bwd",
          "gather_rows", "lift_fwd", "lift_bwd", "vec_to_planar", "vec_from_planar"]


def install(monkeypatch):
    """patch coarsegrainingvae_b200.ops for the duration of one test."""
    for name in _NAMES:
        monkeypatch.setattr(ops, name, globals()[name])
