"""Graph builders and batch assembly with the reference's names (data.py of the reference):
``get_neighbor_list``, ``CGDataset.generate_neighbor_list``, ``CG_collate``, ``batch_to``.
The radius search runs on the GPU (cell lists for large frames) and returns the reference's edge set
bit for bit, in the reference's order.
"""
import numpy as np
import torch
from torch.utils.data import Dataset as TorchDataset

from . import ops


def batch_to(batch, device):
    """data.py:16-20."""
    return {key: (val.to(device) if hasattr(val, 'to') else val) for key, val in batch.items()}


def get_neighbor_list(xyz, device='cuda', cutoff=5, undirected=True):
    """data.py:65-82.  ``device`` must be a CUDA device: the search is an sm_100a kernel.  Returns an int64
    [E,2] tensor on that device, sorted like ``torch.nonzero`` (i ascending, then j ascending)."""
    dev = torch.device(device)
    if dev.type != 'cuda':
        raise RuntimeError("get_neighbor_list runs on the GPU only (device=%r); no CPU fallback" % (device,))
    xyz = torch.as_tensor(np.asarray(xyz) if not torch.is_tensor(xyz) else xyz, dtype=torch.float32).to(dev)
    return ops.radius_graph(xyz, cutoff, undirected=undirected)


def get_neighbor_list_batch(xyz, num_atoms, cutoff, undirected=True):
    """All frames of a batch in one launch: equals the per-frame lists of data.py:207-225 after the
    CG_collate offsets (data.py:255-270).  xyz [sum(num_atoms),3] CUDA, num_atoms [B]."""
    counts = torch.as_tensor(num_atoms, dtype=torch.int64)
    frame_ptr = torch.zeros(counts.numel() + 1, dtype=torch.int64)
    frame_ptr[1:] = torch.cumsum(counts.cpu(), 0)
    use_cells = bool(counts.max().item() > 1024) if counts.numel() else False
    return ops.radius_graph(xyz, cutoff, undirected=undirected, frame_ptr=frame_ptr.to(xyz.device), use_cells=use_cells)


def bond_cg_graph(bond_edge_list, mapping, n_cgs):
    """CG graph derived from the bond graph (data.py:227-248): beads I != J are neighbours iff some bond joins an atom of
    I to an atom of J.  The reference forms assign^T . adj . assign as dense float matrices and takes ``nonzero()``; this
    is the same set in the same (row-major, both directions) order from integer keys -- dataset-preparation time, on the
    device the lists live on (the reference keeps them on the CPU)."""
    bond = torch.as_tensor(bond_edge_list, dtype=torch.int64)
    mapping = torch.as_tensor(mapping, dtype=torch.int64)
    a, b = mapping[bond[:, 0]], mapping[bond[:, 1]]
    keys = torch.cat([a * n_cgs + b, b * n_cgs + a])
    keys = torch.unique(keys[torch.cat([a != b, a != b])])          # sorted: row-major order of nonzero()
    return torch.stack([keys // n_cgs, keys % n_cgs], 1)


class CGDataset(TorchDataset):
    """data.py:186-252."""

    def __init__(self, props, check_props=True):
        self.props = props

    def __len__(self):
        return len(self.props['nxyz'])

    def __getitem__(self, idx):
        return {key: val[idx] for key, val in self.props.items()}

    def generate_neighbor_list(self, atom_cutoff, cg_cutoff, device='cuda', undirected=True, use_bond=False):
        """Builds props['nbr_list'] / props['CG_nbr_list'] for every frame.  All frames go through ONE batched
        kernel call per graph kind instead of the reference's python loop (data.py:215-225); results are split
        back into per-frame CPU tensors with frame-local indices, as the reference stores them."""
        if use_bond:
            nbr_list = self.props['bond_edge_list']
        else:
            nbr_list = self._batched(self.props['nxyz'], atom_cutoff, device, undirected)
        if cg_cutoff is None:
            cg_nbr_list = [bond_cg_graph(bond, self.props['CG_mapping'][i], int(self.props['num_CGs'][i]))
                           for i, bond in enumerate(self.props['bond_edge_list'])]
        else:
            cg_nbr_list = self._batched(self.props['CG_nxyz'], cg_cutoff, device, undirected)
        self.props['nbr_list'] = nbr_list
        self.props['CG_nbr_list'] = cg_nbr_list

    @staticmethod
    def _batched(frames, cutoff, device, undirected):
        counts = [int(f.shape[0]) for f in frames]
        xyz = torch.cat([torch.as_tensor(f)[:, 1:4].float() for f in frames], 0).to(device)
        pairs = get_neighbor_list_batch(xyz, counts, cutoff, undirected).cpu()
        offsets = np.cumsum([0] + counts)
        first = torch.searchsorted(pairs[:, 0].contiguous(), torch.as_tensor(offsets, dtype=torch.int64))
        return [pairs[first[k]:first[k + 1]] - int(offsets[k]) for k in range(len(counts))]


def CG_collate(dicts):
    """data.py:255-289: offset the index tensors of every sample by the running atom / bead counts, then
    concatenate (tensors with a shape) or stack (0-d tensors); strings are gathered in lists."""
    atom_off = np.cumsum([0] + [int(d['num_atoms']) for d in dicts])[:-1]
    bead_off = np.cumsum([0] + [int(d['num_CGs']) for d in dicts])[:-1]
    shifted = []
    for d, a, b in zip(dicts, atom_off, bead_off):
        e = dict(d)
        e['nbr_list'] = d['nbr_list'] + int(a)
        e['bond_edge_list'] = d['bond_edge_list'] + int(a)
        e['CG_mapping'] = d['CG_mapping'] + int(b)
        e['CG_nbr_list'] = d['CG_nbr_list'] + int(b)
        shifted.append(e)
    batch = {}
    for key, val in shifted[0].items():
        if hasattr(val, 'shape') and len(val.shape) > 0:
            batch[key] = torch.cat([s[key] for s in shifted], dim=0)
        elif isinstance(val, str):
            batch[key] = [s[key] for s in shifted]
        else:
            batch[key] = torch.stack([torch.as_tensor(s[key]) for s in shifted], dim=0)
    return batch
