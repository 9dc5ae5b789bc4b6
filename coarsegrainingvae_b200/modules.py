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
# "Dropout" is not an activation (the drivers run dropout = 0.0 everywhere).
ACTIVATION_CODES = {"linear": 0, "swish": 1, "ReLU": 2, "Tanh": 3, "sigmoid": 4, "shifted_softplus": 5, "LeakyReLU": 6,
                    "ELU": 7}


def activation_code(name):
    if name not in ACTIVATION_CODES:
        raise NotImplementedError("activation %r has no sm_100a epilogue yet (have: %s)" % (name, sorted(ACTIVATION_CODES)))
    return ACTIVATION_CODES[name]


class Swish(nn.Module):
    """x * sigmoid(x) (modules.py:16-21); parameter-free marker, fused into GEMM epilogues."""
    code = 1


class shifted_softplus(nn.Module):
    """softplus(x) - ln 2 (modules.py:8-14); parameter-free marker, fused into GEMM epilogues."""
    code = 5


# module class of an activation marker -> kernel code (the heads atom_munet / atom_sigmanet / prior are nn.Sequential
# (Linear, activation module, Linear) built by the drivers: scripts/run_ala.py:184-185)
MODULE_CODES = {"Swish": 1, "ReLU": 2, "Tanh": 3, "Sigmoid": 4, "shifted_softplus": 5, "LeakyReLU": 6, "ELU": 7}


def module_activation_code(module):
    name = module.__class__.__name__
    if name not in MODULE_CODES:
        raise NotImplementedError("activation module %s has no sm_100a epilogue" % name)
    if name == "LeakyReLU" and abs(module.negative_slope - 0.01) > 0 or name == "ELU" and module.alpha != 1.0:
        raise NotImplementedError("%s with non-default parameters has no sm_100a epilogue" % name)
    return MODULE_CODES[name]


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
