"""`torch.library` registration of the hot-path entry points: ``torch.ops.cgvae_b200.*``.

The C ABI (include/cgvae_b200.h) is the drop-in boundary; this module exposes its core kernels to the PyTorch
dispatcher as custom ops with fake (meta) kernels and registered autograd formulas, so that they compose with the
rest of a PyTorch program (autograd, `torch.library.opcheck`, tracing front ends) like any ATen op:

    radius_graph(xyz, cutoff, undirected)                      data.py:65-82                    (no gradient)
    dense(x, W, b) / mlp2(x, W1, b1, W2, b2, act)              modules.py:99-112 (Dense), conv.py:41-49 (phi-MLP)
    message_layer(n_split, phi, v, <CSR + geometry>, Wf, bf, res_s, res_v) -> (s, v, q)
                                                               conv.py:505-563 / 358-402 / 703-733
    segment_reduce(X, rowptr, atoms, mapping, mean)            scatter_mean / scatter_add cgvae.py:297-298

The model classes (modules / conv / cgvae) call the same kernels through `functions.py` autograd nodes, which add
what a dispatcher op cannot express cheaply -- gradients written into a caller-owned flat buffer, deferred grouped
weight gradients, forked graph branches -- and skip the per-call dispatcher cost; results are identical
(tests/test_gpu_parity.py::test_torch_library_ops_match_autograd_nodes).  CUDA tensors only: there is no CPU kernel.
"""
from typing import Optional, Tuple

import torch
from torch.library import custom_op, register_autograd

from . import ops

NS = "cgvae_b200"


def _graph(rowptr, col, rowptr_t, col_t, perm_t, n_send):
    return ops.Graph(rowptr.shape[0] - 1, n_send, col.shape[0], rowptr, col, None, rowptr_t, col_t, perm_t)


def _geom(graph, basis, unit, n_rbf):
    return ops.Geometry(graph, basis, unit, n_rbf, basis.shape[1], 0.0)


# ------------------------------------------------------------------------------------------------ radius graph

@custom_op(NS + "::radius_graph", mutates_args=(), device_types="cuda")
def radius_graph(xyz: torch.Tensor, cutoff: float, undirected: bool) -> torch.Tensor:
    return ops.radius_graph(xyz, cutoff, undirected=undirected)


@radius_graph.register_fake
def _(xyz, cutoff, undirected):
    n_edges = torch.library.get_ctx().new_dynamic_size()
    return xyz.new_empty((n_edges, 2), dtype=torch.int64)


# ------------------------------------------------------------------------------------------------ dense / phi-MLP

@custom_op(NS + "::dense", mutates_args=(), device_types="cuda")
def dense(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """x W^T + b   (nn.Linear / Dense without activation)"""
    return ops.linear_fwd(x.contiguous(), W, b, 0)


@dense.register_fake
def _(x, W, b):
    return x.new_empty((x.shape[0], W.shape[0]))


@custom_op(NS + "::dense_backward", mutates_args=(), device_types="cuda")
def dense_backward(gy: torch.Tensor, x: torch.Tensor, W: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    gy = gy.contiguous()
    return ops.linear_bwd_input(gy, W), ops.linear_bwd_weight(gy, x.contiguous()), ops.colsum(gy)


@dense_backward.register_fake
def _(gy, x, W):
    return torch.empty_like(x), torch.empty_like(W), W.new_empty((W.shape[0],))


def _dense_setup(ctx, inputs, output):
    x, W, b = inputs
    ctx.save_for_backward(x, W)
    ctx.has_bias = b is not None


def _dense_bwd(ctx, gy):
    x, W = ctx.saved_tensors
    gx, gW, gb = torch.ops.cgvae_b200.dense_backward(gy, x, W)
    return gx, gW, (gb if ctx.has_bias else None)


register_autograd(NS + "::dense", _dense_bwd, setup_context=_dense_setup)


@custom_op(NS + "::mlp2", mutates_args=(), device_types="cuda")
def mlp2(x: torch.Tensor, W1: torch.Tensor, b1: torch.Tensor, W2: torch.Tensor, b2: torch.Tensor, act: int
         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Dense2(act(Dense1(x))): the phi-MLP of InvariantMessage (conv.py:41-49); act: 1 swish, 2 ReLU, 3 tanh.
    Returns (y, a1, z1): output, hidden activation and its pre-activation (saved for the backward)."""
    x = x.contiguous()
    a1, z1 = ops.linear_fwd(x, W1, b1, act, save_pre=True)
    return ops.linear_fwd(a1, W2, b2, 0), a1, z1


@mlp2.register_fake
def _(x, W1, b1, W2, b2, act):
    h = x.new_empty((x.shape[0], W1.shape[0]))
    return x.new_empty((x.shape[0], W2.shape[0])), h, torch.empty_like(h)


@custom_op(NS + "::mlp2_backward", mutates_args=(), device_types="cuda")
def mlp2_backward(gy: torch.Tensor, x: torch.Tensor, a1: torch.Tensor, z1: torch.Tensor, W1: torch.Tensor, W2: torch.Tensor,
                  act: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    gy = gy.contiguous()
    gz1 = ops.linear_bwd_input(gy, W2, z_in=z1, dact=act)          # (gy W2) * act'(z1): fused GEMM epilogue
    gW2, gb2 = ops.linear_bwd_weight(gy, a1), ops.colsum(gy)
    gx = ops.linear_bwd_input(gz1, W1)
    gW1, gb1 = ops.linear_bwd_weight(gz1, x.contiguous()), ops.colsum(gz1)
    return gx, gW1, gb1, gW2, gb2


@mlp2_backward.register_fake
def _(gy, x, a1, z1, W1, W2, act):
    return (torch.empty_like(x), torch.empty_like(W1), W1.new_empty((W1.shape[0],)), torch.empty_like(W2),
            W2.new_empty((W2.shape[0],)))


def _mlp2_setup(ctx, inputs, output):
    x, W1, b1, W2, b2, act = inputs
    ctx.save_for_backward(x, output[1], output[2], W1, W2)
    ctx.act = act


def _mlp2_bwd(ctx, gy, ga1_unused, gz1_unused):
    x, a1, z1, W1, W2 = ctx.saved_tensors
    gx, gW1, gb1, gW2, gb2 = torch.ops.cgvae_b200.mlp2_backward(gy, x, a1, z1, W1, W2, ctx.act)
    return gx, gW1, gb1, gW2, gb2, None


register_autograd(NS + "::mlp2", _mlp2_bwd, setup_context=_mlp2_setup)


# ------------------------------------------------------------------------------------------------ message layer

@custom_op(NS + "::message_layer", mutates_args=(), device_types="cuda")
def message_layer(n_split: int, phi: torch.Tensor, v: Optional[torch.Tensor], rowptr: torch.Tensor, col: torch.Tensor,
                  rowptr_t: torch.Tensor, col_t: torch.Tensor, perm_t: torch.Tensor, basis: torch.Tensor, unit: torch.Tensor,
                  Wf: torch.Tensor, bf: torch.Tensor, res_s: Optional[torch.Tensor], res_v: Optional[torch.Tensor], n_rbf: int
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """one fused message layer: phi [n_send, n_split, F] (the phi-MLP output), v [n_send, 3, F] planar vectors or None
    (all zero), receiver / sender CSR and per-edge geometry as built by ops.build_graph / ops.edge_geometry,
    Wf [n_split*F, n_rbf], bf [n_split*F], optional residual state on the receiver set.
    Returns (s [n_recv, F], v [n_recv, 3, F], q) -- q is the saved cross-term sum of the 4-split block (empty otherwise)."""
    g = _graph(rowptr, col, rowptr_t, col_t, perm_t, phi.shape[0])
    geom = _geom(g, basis, unit, n_rbf)
    vs = v.contiguous() if v is not None else None
    out_s, out_v, q = ops.message_fwd(n_split, phi.contiguous(), vs, vs if n_split == 4 else None, geom, Wf, bf, res_s, res_v,
                                      want_q=(n_split == 4 and v is not None))
    return out_s, out_v, (q if q is not None else phi.new_empty((0,)))


@message_layer.register_fake
def _(n_split, phi, v, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf, res_s, res_v, n_rbf):
    n_recv, F = rowptr.shape[0] - 1, phi.shape[-1]
    q = phi.new_empty((n_recv, 3, F)) if (n_split == 4 and v is not None) else phi.new_empty((0,))
    return phi.new_empty((n_recv, F)), phi.new_empty((n_recv, 3, F)), q


@custom_op(NS + "::message_layer_backward", mutates_args=(), device_types="cuda")
def message_layer_backward(n_split: int, phi: torch.Tensor, v: Optional[torch.Tensor], q: torch.Tensor, rowptr: torch.Tensor,
                           col: torch.Tensor, rowptr_t: torch.Tensor, col_t: torch.Tensor, perm_t: torch.Tensor,
                           basis: torch.Tensor, unit: torch.Tensor, Wf: torch.Tensor, bf: torch.Tensor, g_s: torch.Tensor,
                           g_v: torch.Tensor, n_rbf: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    g = _graph(rowptr, col, rowptr_t, col_t, perm_t, phi.shape[0])
    geom = _geom(g, basis, unit, n_rbf)
    vs = v.contiguous() if v is not None else None
    qq = q if q.numel() else None
    g_phi, g_vs, dWf, dbf = ops.message_bwd(n_split, phi.contiguous(), vs, vs if n_split == 4 else None, qq, geom, Wf, bf,
                                            g_s.contiguous(), g_v.contiguous(), False, sink=False)
    return g_phi, g_vs, dWf, dbf


@message_layer_backward.register_fake
def _(n_split, phi, v, q, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf, g_s, g_v, n_rbf):
    return torch.empty_like(phi), phi.new_empty((phi.shape[0], 3, phi.shape[-1])), torch.empty_like(Wf), torch.empty_like(bf)


def _message_setup(ctx, inputs, output):
    n_split, phi, v, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf, res_s, res_v, n_rbf = inputs
    ctx.save_for_backward(phi, v, output[2], rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf)
    ctx.n_split, ctx.n_rbf = n_split, n_rbf
    ctx.has_v, ctx.has_res_s, ctx.has_res_v = v is not None, res_s is not None, res_v is not None


def _message_bwd(ctx, g_s, g_v, g_q_unused):
    phi, v, q, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf = ctx.saved_tensors
    g_phi, g_vs, dWf, dbf = torch.ops.cgvae_b200.message_layer_backward(ctx.n_split, phi, v, q, rowptr, col, rowptr_t, col_t,
                                                                         perm_t, basis, unit, Wf, bf, g_s, g_v, ctx.n_rbf)
    return (None, g_phi, g_vs if ctx.has_v else None, None, None, None, None, None, None, None, dWf, dbf,
            g_s if ctx.has_res_s else None, g_v if ctx.has_res_v else None, None)


register_autograd(NS + "::message_layer", _message_bwd, setup_context=_message_setup)


# ------------------------------------------------------------------------------------------------ bead pooling

@custom_op(NS + "::segment_reduce", mutates_args=(), device_types="cuda")
def segment_reduce(X: torch.Tensor, rowptr: torch.Tensor, atoms: torch.Tensor, mapping: torch.Tensor, mean: bool) -> torch.Tensor:
    """scatter_mean / scatter_add over beads: X [N, ...] -> [n_beads, ...] with the bead CSR of ops.build_segments."""
    seg = ops.Segments(rowptr.shape[0] - 1, mapping, rowptr, atoms, None, None)
    return ops.segment_reduce_fwd(X.contiguous(), seg, mean)


@segment_reduce.register_fake
def _(X, rowptr, atoms, mapping, mean):
    return X.new_empty((rowptr.shape[0] - 1,) + tuple(X.shape[1:]))


@custom_op(NS + "::segment_reduce_backward", mutates_args=(), device_types="cuda")
def segment_reduce_backward(g: torch.Tensor, rowptr: torch.Tensor, atoms: torch.Tensor, mapping: torch.Tensor, mean: bool
                            ) -> torch.Tensor:
    seg = ops.Segments(rowptr.shape[0] - 1, mapping, rowptr, atoms, None, None)
    return ops.segment_reduce_bwd(g.contiguous(), seg, mean)


@segment_reduce_backward.register_fake
def _(g, rowptr, atoms, mapping, mean):
    return g.new_empty((mapping.shape[0],) + tuple(g.shape[1:]))


def _segment_setup(ctx, inputs, output):
    X, rowptr, atoms, mapping, mean = inputs
    ctx.save_for_backward(rowptr, atoms, mapping)
    ctx.mean = mean


def _segment_bwd(ctx, g):
    rowptr, atoms, mapping = ctx.saved_tensors
    return torch.ops.cgvae_b200.segment_reduce_backward(g, rowptr, atoms, mapping, ctx.mean), None, None, None, None


register_autograd(NS + "::segment_reduce", _segment_bwd, setup_context=_segment_setup)


# ------------------------------------------------------------------------------------------------ 9-split (pseudo-scalar) layer

@custom_op(NS + "::message9_layer", mutates_args=(), device_types="cuda")
def message9_layer(phi: torch.Tensor, s: torch.Tensor, sbar: torch.Tensor, v: torch.Tensor, vbar: torch.Tensor, rowptr: torch.Tensor,
                   col: torch.Tensor, rowptr_t: torch.Tensor, col_t: torch.Tensor, perm_t: torch.Tensor, basis: torch.Tensor,
                   unit: torch.Tensor, Wf: torch.Tensor, bf: torch.Tensor, residual: bool, n_rbf: int
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """EquiMessagePsuedo (conv.py:180-242) on one homogeneous graph: phi [n, 9, F] (phi-MLP output), scalar / pseudo-scalar
    states s, sbar [n, F], planar vector / pseudo-vector states v, vbar [n, 3, F]; residual: the adds of cgvae.py:108-111."""
    g = _graph(rowptr, col, rowptr_t, col_t, perm_t, phi.shape[0])
    outs = ops.message9_fwd(phi.contiguous(), s.contiguous(), sbar.contiguous(), v.contiguous(), vbar.contiguous(),
                            _geom(g, basis, unit, n_rbf), Wf, bf, residual)
    return outs[0], outs[1], outs[2], outs[3]


@message9_layer.register_fake
def _(phi, s, sbar, v, vbar, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf, residual, n_rbf):
    return torch.empty_like(s), torch.empty_like(sbar), torch.empty_like(v), torch.empty_like(vbar)


@custom_op(NS + "::message9_layer_backward", mutates_args=(), device_types="cuda")
def message9_layer_backward(phi: torch.Tensor, s: torch.Tensor, sbar: torch.Tensor, v: torch.Tensor, vbar: torch.Tensor,
                            rowptr: torch.Tensor, col: torch.Tensor, rowptr_t: torch.Tensor, col_t: torch.Tensor, perm_t: torch.Tensor,
                            basis: torch.Tensor, unit: torch.Tensor, Wf: torch.Tensor, bf: torch.Tensor, residual: bool, n_rbf: int,
                            g_s: torch.Tensor, g_sbar: torch.Tensor, g_v: torch.Tensor, g_vbar: torch.Tensor
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    g = _graph(rowptr, col, rowptr_t, col_t, perm_t, phi.shape[0])
    geom = _geom(g, basis, unit, n_rbf)
    fork, ops.FORK_ENABLED = ops.FORK_ENABLED, False          # a dispatcher op returns finished tensors: no forked branch
    try:
        outs = ops.message9_bwd(phi.contiguous(), s.contiguous(), sbar.contiguous(), v.contiguous(), vbar.contiguous(), geom, Wf, bf,
                                residual, g_s.contiguous(), g_sbar.contiguous(), g_v.contiguous(), g_vbar.contiguous())
    finally:
        ops.FORK_ENABLED = fork
    return tuple(outs)


@message9_layer_backward.register_fake
def _(phi, s, sbar, v, vbar, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf, residual, n_rbf, g_s, g_sbar, g_v, g_vbar):
    return (torch.empty_like(s), torch.empty_like(sbar), torch.empty_like(v), torch.empty_like(vbar), torch.empty_like(phi),
            torch.empty_like(Wf), torch.empty_like(bf))


def _message9_setup(ctx, inputs, output):
    phi, s, sbar, v, vbar, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf, residual, n_rbf = inputs
    ctx.save_for_backward(phi, s, sbar, v, vbar, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf)
    ctx.residual, ctx.n_rbf = residual, n_rbf


def _message9_bwd(ctx, g_s, g_sbar, g_v, g_vbar):
    phi, s, sbar, v, vbar, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf = ctx.saved_tensors
    gi_s, gi_sbar, gi_v, gi_vbar, g_phi, dWf, dbf = torch.ops.cgvae_b200.message9_layer_backward(
        phi, s, sbar, v, vbar, rowptr, col, rowptr_t, col_t, perm_t, basis, unit, Wf, bf, ctx.residual, ctx.n_rbf, g_s, g_sbar, g_v,
        g_vbar)
    return g_phi, gi_s, gi_sbar, gi_v, gi_vbar, None, None, None, None, None, None, None, dWf, dbf, None, None


register_autograd(NS + "::message9_layer", _message9_bwd, setup_context=_message9_setup)


# ------------------------------------------------------------------------------------------------ update block

@custom_op(NS + "::update_block", mutates_args=(), device_types="cuda")
def update_block(s: torch.Tensor, v: torch.Tensor, U: torch.Tensor, V: torch.Tensor, A0: torch.Tensor, c0: torch.Tensor,
                 A1: torch.Tensor, c1: torch.Tensor, act: int, residual: bool
                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """UpdateBlock.forward (conv.py:588-616) on planar vectors: s [N, F], v [N, 3, F]; U, V the bias-free u_mat / v_mat
    weights, (A0, c0), (A1, c1) the two Dense layers of s_dense; residual: s + ds, v + dv (cgvae.py:122-123).
    Returns (s_out, v_out, Uv, Vv, x, h, z, q) -- the last six are the tensors the backward needs."""
    s, v = s.contiguous(), v.contiguous()
    N, F = s.shape
    Uv, Vv = ops.linear_pair_fwd(v.view(3 * N, F), U, V)
    Uv, Vv = Uv.view(N, 3, F), Vv.view(N, 3, F)
    x = ops.update_norm_fwd(s, Vv)
    h, z = ops.linear_fwd(x, A0, c0, act, save_pre=True)
    q = ops.linear_fwd(h, A1, c1, 0).view(N, 3, F)
    s_out, v_out = ops.update_combine_fwd(s, v, Uv, Vv, q, residual)
    return s_out, v_out, Uv, Vv, x, h, z, q


@update_block.register_fake
def _(s, v, U, V, A0, c0, A1, c1, act, residual):
    N, F = s.shape
    h = s.new_empty((N, A0.shape[0]))
    return (torch.empty_like(s), torch.empty_like(v), torch.empty_like(v), torch.empty_like(v), s.new_empty((N, A0.shape[1])), h,
            torch.empty_like(h), torch.empty_like(v))


@custom_op(NS + "::update_block_backward", mutates_args=(), device_types="cuda")
def update_block_backward(g_s: torch.Tensor, g_v: torch.Tensor, v: torch.Tensor, Uv: torch.Tensor, Vv: torch.Tensor, x: torch.Tensor,
                          h: torch.Tensor, z: torch.Tensor, q: torch.Tensor, U: torch.Tensor, V: torch.Tensor, A0: torch.Tensor,
                          A1: torch.Tensor, act: int, residual: bool
                          ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    g_s, g_v = g_s.contiguous(), g_v.contiguous()
    N, _, F = v.shape
    gq, gUv, gVv = ops.update_combine_bwd(Uv, Vv, q, g_s, g_v)
    gq2, v2 = gq.view(N, 3 * F), v.view(3 * N, F)
    gUv2, gVv2 = gUv.view(3 * N, F), gVv.view(3 * N, F)
    gz = ops.linear_bwd_input(gq2, A1, z_in=z, dact=act)
    gA1, gc1 = ops.linear_bwd_weight(gq2, h), ops.colsum(gq2)
    gU = ops.linear_bwd_weight(gUv2, v2)
    gx = ops.linear_bwd_input(gz, A0)
    gA0, gc0 = ops.linear_bwd_weight(gz, x), ops.colsum(gz)
    gs_in = ops.update_norm_bwd(x, Vv, gx, g_s, gVv, residual)               # adds the norm path into gVv in place
    gv_in = ops.linear_bwd_input(gUv2, U, add=g_v.view(3 * N, F) if residual else None)
    gv_in = ops.linear_bwd_input(gVv2, V, add=gv_in).view(N, 3, F)
    gV = ops.linear_bwd_weight(gVv2, v2)
    return gs_in, gv_in, gU, gV, gA0, gc0, gA1, gc1


@update_block_backward.register_fake
def _(g_s, g_v, v, Uv, Vv, x, h, z, q, U, V, A0, A1, act, residual):
    return (torch.empty_like(g_s), torch.empty_like(v), torch.empty_like(U), torch.empty_like(V), torch.empty_like(A0),
            A0.new_empty((A0.shape[0],)), torch.empty_like(A1), A1.new_empty((A1.shape[0],)))


def _update_setup(ctx, inputs, output):
    s, v, U, V, A0, c0, A1, c1, act, residual = inputs
    ctx.save_for_backward(v, output[2], output[3], output[4], output[5], output[6], output[7], U, V, A0, A1)
    ctx.act, ctx.residual = act, residual


def _update_bwd(ctx, g_s, g_v, *unused):
    v, Uv, Vv, x, h, z, q, U, V, A0, A1 = ctx.saved_tensors
    gs_in, gv_in, gU, gV, gA0, gc0, gA1, gc1 = torch.ops.cgvae_b200.update_block_backward(g_s, g_v, v, Uv, Vv, x, h, z, q, U, V, A0, A1,
                                                                                          ctx.act, ctx.residual)
    return gs_in, gv_in, gU, gV, gA0, gc0, gA1, gc1, None, None


register_autograd(NS + "::update_block", _update_bwd, setup_context=_update_setup)


# ------------------------------------------------------------------------------------------------ bead -> atom lifting

@custom_op(NS + "::lift", mutates_args=(), device_types="cuda")
def lift(V: torch.Tensor, cg_xyz: torch.Tensor, mapping: torch.Tensor, rank: torch.Tensor, rowptr: torch.Tensor, atoms: torch.Tensor,
         pin: Optional[torch.Tensor], mode: int) -> torch.Tensor:
    """xyz [N, 3] = cg_xyz[mapping] + V[mapping, :, rank] (cgvae.py:466-482); mode 1 subtracts the bead mean (offset=True), mode 2
    zeroes the displacement of pinned atoms (PCN C-alpha re-anchoring, cgvae.py:569-571); bead CSR / ranks from build_segments."""
    seg = ops.Segments(rowptr.shape[0] - 1, mapping, rowptr, atoms, None, rank)
    return ops.lift_fwd(V.contiguous(), cg_xyz.contiguous(), seg, mode, pin)


@lift.register_fake
def _(V, cg_xyz, mapping, rank, rowptr, atoms, pin, mode):
    return cg_xyz.new_empty((mapping.shape[0], 3))


@custom_op(NS + "::lift_backward", mutates_args=(), device_types="cuda")
def lift_backward(g: torch.Tensor, mapping: torch.Tensor, rank: torch.Tensor, rowptr: torch.Tensor, atoms: torch.Tensor,
                  pin: Optional[torch.Tensor], mode: int, F: int) -> torch.Tensor:
    seg = ops.Segments(rowptr.shape[0] - 1, mapping, rowptr, atoms, None, rank)
    return ops.lift_bwd(g.contiguous(), seg, F, mode, pin)


@lift_backward.register_fake
def _(g, mapping, rank, rowptr, atoms, pin, mode, F):
    return g.new_empty((rowptr.shape[0] - 1, 3, F))


def _lift_setup(ctx, inputs, output):
    V, cg_xyz, mapping, rank, rowptr, atoms, pin, mode = inputs
    ctx.save_for_backward(mapping, rank, rowptr, atoms, pin)
    ctx.mode, ctx.F = mode, V.shape[-1]


def _lift_bwd(ctx, g):
    mapping, rank, rowptr, atoms, pin = ctx.saved_tensors
    return torch.ops.cgvae_b200.lift_backward(g, mapping, rank, rowptr, atoms, pin, ctx.mode, ctx.F), None, None, None, None, None, None, None


register_autograd(NS + "::lift", _lift_bwd, setup_context=_lift_setup)


# ------------------------------------------------------------------------------------------------ clip + Adam

@custom_op(NS + "::adam_clip_step", mutates_args=("p", "m", "v", "step"), device_types="cuda")
def adam_clip_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: torch.Tensor, max_norm: float, lr: float,
                   beta1: float, beta2: float, eps: float, grad_scale: float) -> None:
    """clip_grad_norm_(max_norm) + Adam.step() (scripts/utils.py:151-157) on flat fp32 buffers, in place; step: float [1]."""
    ops.adam_clip_step(p, g, m, v, step, max_norm, lr, betas=(beta1, beta2), eps=eps, grad_scale=grad_scale)


OPS = ("radius_graph", "dense", "dense_backward", "mlp2", "mlp2_backward", "message_layer", "message_layer_backward", "segment_reduce",
       "segment_reduce_backward", "message9_layer", "message9_layer_backward", "update_block", "update_block_backward", "lift",
       "lift_backward", "adam_clip_step")
