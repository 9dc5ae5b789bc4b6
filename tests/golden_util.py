"""Helpers to read tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


def section(z, prefix):
    """{subkey: tensor} for all keys 'prefix/subkey'."""
    pre = prefix + "/"
    return {k[len(pre):]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith(pre)}


def rel_err(a, b):
    """scale-relative error: max|a-b| / max(|b|, tiny). The parity metric used everywhere:
    the north-star tolerance '1e-5 relative' is read as relative to the tensor's scale
    (element-wise relative error is meaningless at the zeros of a signed quantity)."""
    a = torch.as_tensor(a).detach().to(torch.float64).cpu()
    b = torch.as_tensor(b).detach().to(torch.float64).cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
