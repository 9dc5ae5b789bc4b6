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


class DeviceCGDataset(object):
    """On-GPU batch assembly for the trajectory datasets of the reference: every frame is the
    same molecule (datasets.py:434-500: one atomic-number vector, one CG mapping, one bond list; only the coordinates
    differ), so the whole trajectory stays resident in HBM as ``xyz [T, n, 3]`` and a batch is assembled ON THE DEVICE from
    a vector of frame indices:

    * ``nxyz`` / ``CG_nxyz`` rows gathered, ``CG_mapping`` / ``bond_edge_list`` offset by the running atom / bead counts
      (``CG_collate``, data.py:255-289);
    * ``CG_xyz = scatter_mean(xyz, mapping)`` (datasets.py:487) by the bead-pooling kernel when no CG trajectory is given;
    * ``nbr_list`` / ``CG_nbr_list`` built on the fly by ONE batched radius-graph launch each (``get_neighbor_list`` per
      frame in ``generate_neighbor_list``, data.py:207-225) -- no stored edge lists, no per-frame CPU pre-pass, no H2D of
      index tensors; ``cg_cutoff=None`` takes the bond-derived CG graph (data.py:227-248), identical in every frame.

    ``collate(idx)`` returns the dict ``CG_collate([dataset[i] for i in idx])`` would (same keys, dtypes, order: bit-exact
    index tensors); ``collate(idx, static=True)`` returns the static-capacity form of ``train.to_static_batch`` (padded lists
    + device-side live counts) with NO host synchronisation, ready for ``GraphedTrainStep.step``."""

    def __init__(self, atomic_nums, xyz, mapping, bond_edge_list, atom_cutoff, cg_cutoff, device='cuda', undirected=True,
                 cg_xyz=None):
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise RuntimeError("DeviceCGDataset lives on a CUDA device (device=%r); no CPU fallback" % (device,))
        self.device = dev
        self.xyz = torch.as_tensor(np.asarray(xyz) if not torch.is_tensor(xyz) else xyz, dtype=torch.float32).to(dev).contiguous()
        self.n_frames, self.n_atoms = int(self.xyz.shape[0]), int(self.xyz.shape[1])
        self.z = torch.as_tensor(atomic_nums, dtype=torch.float32).to(dev)
        self.mapping = torch.as_tensor(mapping, dtype=torch.int64).to(dev)
        self.n_cgs = int(self.mapping.max().item()) + 1
        self.bonds = torch.as_tensor(bond_edge_list, dtype=torch.int64).to(dev)
        self.atom_cutoff, self.cg_cutoff, self.undirected = atom_cutoff, cg_cutoff, undirected
        self.seg = ops.build_segments(self.mapping, self.n_cgs)
        self.cg_xyz = None if cg_xyz is None else torch.as_tensor(cg_xyz, dtype=torch.float32).to(dev).contiguous()
        self.cg_bond_graph = bond_cg_graph(self.bonds, self.mapping, self.n_cgs) if cg_cutoff is None else None

    def __len__(self):
        return self.n_frames

    def capacities(self, batch_size):
        """worst-case list capacities of a static batch of ``batch_size`` frames."""
        n, nb, div = self.n_atoms, self.n_cgs, (2 if self.undirected else 1)
        cg = batch_size * nb * (nb - 1) // div if self.cg_cutoff is not None else batch_size * int(self.cg_bond_graph.shape[0])
        return {"nbr_list": batch_size * n * (n - 1) // div, "CG_nbr_list": cg, "bond_edge_list": batch_size * int(self.bonds.shape[0])}

    def collate(self, idx, static=False):
        dev, n, nb = self.device, self.n_atoms, self.n_cgs
        idx = torch.as_tensor(idx, dtype=torch.int64, device=dev)
        B = int(idx.shape[0])
        xyz = self.xyz.index_select(0, idx)                                        # [B, n, 3]
        flat = xyz.reshape(B * n, 3)
        nxyz = torch.cat([self.z.repeat(B)[:, None], flat], 1)
        if self.cg_xyz is not None:
            cg_flat = self.cg_xyz.index_select(0, idx).reshape(B * nb, 3)
        else:
            # scatter_mean per frame == bead pooling over the block-diagonal batch mapping
            cg_flat = ops.segment_reduce_fwd(flat.contiguous(), self._batch_segments(B), True)
        frame = torch.arange(B, dtype=torch.int64, device=dev)
        CG_nxyz = torch.cat([torch.arange(nb, dtype=torch.float32, device=dev).repeat(B)[:, None], cg_flat], 1)
        CG_mapping = (self.mapping[None, :] + frame[:, None] * nb).reshape(-1)
        bond = (self.bonds[None, :, :] + (frame * n)[:, None, None]).reshape(-1, 2)
        atom_ptr = torch.arange(B + 1, dtype=torch.int64, device=dev) * n
        bead_ptr = torch.arange(B + 1, dtype=torch.int64, device=dev) * nb
        batch = {'nxyz': nxyz, 'CG_nxyz': CG_nxyz, 'num_atoms': torch.full((B,), n, dtype=torch.int64, device=dev),
                 'num_CGs': torch.full((B,), nb, dtype=torch.int64, device=dev), 'CG_mapping': CG_mapping,
                 'bond_edge_list': bond}
        cap = self.capacities(B) if static else None
        kw = lambda key, mf: dict(capacity=cap[key], max_frame=mf) if static else dict(max_frame=mf)
        nbr = ops.radius_graph(flat, self.atom_cutoff, self.undirected, frame_ptr=atom_ptr, use_cells=n > 1024, **kw("nbr_list", n))
        if self.cg_cutoff is not None:
            cg_nbr = ops.radius_graph(cg_flat.contiguous(), self.cg_cutoff, self.undirected, frame_ptr=bead_ptr, use_cells=nb > 1024,
                                      **kw("CG_nbr_list", nb))
        else:
            lst = (self.cg_bond_graph[None, :, :] + (frame * nb)[:, None, None]).reshape(-1, 2)
            cg_nbr = (lst, torch.full((1,), int(lst.shape[0]), dtype=torch.int64, device=dev)) if static else lst
        if static:
            batch['nbr_list'], batch['nbr_count'] = nbr
            batch['CG_nbr_list'], batch['CG_nbr_count'] = cg_nbr
            batch['bond_count'] = torch.full((1,), int(bond.shape[0]), dtype=torch.int64, device=dev)
            # radius graphs are one-directional (j > i) when undirected: the flipped half is generated by the CSR kernels;
            # the bond-derived CG graph already holds both directions (nonzero() of a symmetric matrix)
            batch['nbr_symmetrize'] = bool(self.undirected)
            batch['CG_nbr_symmetrize'] = bool(self.undirected) if self.cg_cutoff is not None else False
        else:
            batch['nbr_list'], batch['CG_nbr_list'] = nbr, cg_nbr
        return batch

    def _batch_segments(self, B):
        cache = getattr(self, "_seg_cache", None)
        if cache is None or cache[0] != B:
            frame = torch.arange(B, dtype=torch.int64, device=self.device)
            mapping = (self.mapping[None, :] + frame[:, None] * self.n_cgs).reshape(-1)
            self._seg_cache = (B, ops.build_segments(mapping, B * self.n_cgs))
        return self._seg_cache[1]


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
