"""Import the *real* reference read-only (build container only).

Used by ``tests/golden/make_golden.py`` to generate fixtures and by the optional
"live reference" checks in the CPU test-suite.  ``/root/reference`` does not exist
on the GPU box, so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py``
calls this.

Two third-party modules the reference imports are absent from this image and are
stubbed:

* ``torch_scatter`` (pinned ``torch_scatter==2.0.9``, reference ``requirements.txt:18``).
  In 2.0.9 ``scatter_sum/scatter_add`` is pure Python on ATen: broadcast the
  index to ``src``, allocate ``zeros`` with ``size[dim] = dim_size or index.max()+1``
  and call ``Tensor.scatter_add_``; ``scatter_mean`` divides by the per-segment count
  clamped to >= 1.  The stub restates exactly that published algorithm.
* ``ase`` -- only imported, never used, by ``CoarseGrainingVAE/data.py:8-9``.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("CGVAE_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "CoarseGrainingVAE"))


def _scatter_add(src, index, dim=-1, out=None, dim_size=None):
    if dim < 0:
        dim += src.dim()
    idx = index
    if idx.dim() == 1 and src.dim() > 1:
        view = [1] * src.dim()
        view[dim] = -1
        idx = idx.view(view)
    idx = idx.expand_as(src)
    size = list(src.shape)
    if dim_size is not None:
        size[dim] = dim_size
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    res = torch.zeros(size, dtype=src.dtype, device=src.device)
    return res.scatter_add_(dim, idx, src)


def _scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    total = _scatter_add(src, index, dim, None, dim_size)
    if dim < 0:
        dim += src.dim()
    ones = torch.ones(index.shape, dtype=src.dtype, device=src.device)
    count = _scatter_add(ones, index, 0, None, total.shape[dim]).clamp_(min=1)
    view = [1] * total.dim()
    view[dim] = -1
    return total / count.view(view)


def install_stubs():
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")
        ts.scatter_add = _scatter_add
        ts.scatter_sum = _scatter_add
        ts.scatter_mean = _scatter_mean
        sys.modules["torch_scatter"] = ts
    if "ase" not in sys.modules:
        ase = types.ModuleType("ase")
        ase.Atoms = object
        nl = types.ModuleType("ase.neighborlist")
        nl.neighbor_list = None
        ase.neighborlist = nl
        sys.modules["ase"] = ase
        sys.modules["ase.neighborlist"] = nl


def import_reference():
    """Returns the reference's (modules, conv, cgvae, data) python modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import CoarseGrainingVAE.modules as ref_modules
    import CoarseGrainingVAE.conv as ref_conv
    import CoarseGrainingVAE.cgvae as ref_cgvae
    import CoarseGrainingVAE.data as ref_data
    return ref_modules, ref_conv, ref_cgvae, ref_data
