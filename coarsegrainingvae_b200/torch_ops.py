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

OPS = ("radius_graph", "dense", "dense_backward", "mlp2", "mlp2_backward", "message_layer", "message_layer_backward", "segment_reduce",
       "segment_reduce_backward")
