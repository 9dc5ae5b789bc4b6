"""Parameter holders with the reference's names and state_dict layout (modules.py of the reference).

These classes own nn.Parameters so that ``state_dict()`` keys / shapes / init match
``CoarseGrainingVAE.modules`` (Dense: xavier-uniform weight, zero bias -- modules.py:75-98); the
arithmetic itself runs in the sm_100a kernels via ``functions.py``.
"""
import torch
from torch import nn
from torch.nn.init import xavier_uniform_, zeros_

from . import ops

# activation names of the reference's ``layer_types`` registry (modules.py:32-42) -> kernel codes.
# The hot path ships swish / ReLU / Tanh / linear; the remaining registry entries are listed in
# DESIGN.md as "next".
ACTIVATION_CODES = {"swish": 1, "ReLU": 2, "Tanh": 3, "linear": 0}


def activation_code(name):
    if name not in ACTIVATION_CODES:
        raise NotImplementedError("activation %r has no sm_100a epilogue yet (have: %s)" % (name, sorted(ACTIVATION_CODES)))
    return ACTIVATION_CODES[name]


class Swish(nn.Module):
    """x * sigmoid(x) (modules.py:16-21); parameter-free marker, fused into GEMM epilogues."""
    code = 1


class Dense(nn.Linear):
    """Linear with xavier-uniform weight and zero bias (modules.py:75-98)."""

    def __init__(self, in_features, out_features, bias=True, activation=None, dropout_rate=0.0):
        if dropout_rate:
            raise NotImplementedError("dropout > 0 is never used by the reference drivers (dropout=0.0 everywhere)")
        super().__init__(in_features, out_features, bias)
        self.activation = activation

    def reset_parameters(self):
        xavier_uniform_(self.weight)
        if self.bias is not None:
            zeros_(self.bias)


class PainnRadialBasis(nn.Module):
    """sin(n pi d / c) / d, n = 1..n_rbf (modules.py:139-172).  ``n`` and ``cutoff`` are plain attributes, not
    buffers, exactly like the reference; evaluated inside cgvae_edge_geometry."""

    def __init__(self, n_rbf, cutoff):
        super().__init__()
        self.n = torch.arange(1, n_rbf + 1).float()
        self.cutoff = cutoff


class CosineEnvelope(nn.Module):
    """0.5 (cos(pi d / c) + 1), zero beyond the cutoff (modules.py:45-58)."""

    def __init__(self, cutoff):
        super().__init__()
        self.cutoff = cutoff


class DistanceEmbed(nn.Module):
    """(Dense(rbf(d))) * envelope(d) (modules.py:175-197); keys ``block.1.{weight,bias}``."""

    def __init__(self, n_rbf, cutoff, feat_dim, dropout):
        super().__init__()
        self.block = nn.Sequential(PainnRadialBasis(n_rbf=n_rbf, cutoff=cutoff),
                                   Dense(in_features=n_rbf, out_features=feat_dim, bias=True, dropout_rate=dropout))
        self.f_cut = CosineEnvelope(cutoff=cutoff)
        self.n_rbf, self.cutoff = n_rbf, cutoff

    @property
    def filter_weight(self):
        return self.block[1].weight

    @property
    def filter_bias(self):
        return self.block[1].bias


def make_directed(nbr_list):
    """conv.py:10-20 on the device: append the flipped list unless both orientations already occur."""
    up, down = ops.edge_orientation(nbr_list)
    if up and down:
        return nbr_list, True
    return torch.cat([nbr_list, nbr_list.flip(1)], dim=0), False
