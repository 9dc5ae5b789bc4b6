"""Integer / index part of the oracle (TEST INFRASTRUCTURE, see package docstring).

numpy restatement of the reference graph builders and batch assembly.  All of it is
bit-exact work: edge sets, offsets and ranks must be reproduced exactly.
Citations are into /root/reference/CoarseGrainingVAE.
"""
import numpy as np


def radius_graph(xyz, cutoff, undirected=True, row_block=2048):
    """``get_neighbor_list`` data.py:65-82.

    The reference builds the dense fp32 matrix ``(x_j - x_i).pow(2).sum(2).sqrt()`` and
    keeps ``dist <= cutoff`` (cutoff rounded to fp32 by the tensor-scalar comparison),
    removes the diagonal and, when ``undirected``, keeps ``j > i``; ``torch.nonzero``
    returns the pairs row-major, i.e. sorted by (i asc, j asc).

    fp32 recipe that reproduces torch bit-for-bit (hard part 3):
    ``d2 = (dx*dx + dy*dy) + dz*dz`` with every product and sum rounded to fp32 (no FMA),
    then an IEEE sqrt.  Row-blocked so n = 20 000 needs MBs, not the reference's 4.8 GB.
    """
    x = np.ascontiguousarray(np.asarray(xyz, dtype=np.float32).reshape(-1, 3))
    n = x.shape[0]
    cut = np.float32(cutoff)
    out = []
    for lo in range(0, n, row_block):
        hi = min(n, lo + row_block)
        d = x[None, :, :] - x[lo:hi, None, :]                   # [rows, n, 3] = x_j - x_i
        sq = d * d
        d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
        mask = np.sqrt(d2) <= cut
        rows = np.arange(lo, hi)
        mask[rows - lo, rows] = False
        if undirected:
            mask &= np.arange(n)[None, :] > rows[:, None]
        i, j = np.nonzero(mask)
        out.append(np.stack([i + lo, j], axis=1))
    if not out:
        return np.zeros((0, 2), dtype=np.int64)
    return np.concatenate(out, axis=0).astype(np.int64)


def make_directed(nbr_list):
    """conv.py:10-20."""
    nbr_list = np.asarray(nbr_list, dtype=np.int64)
    if (nbr_list[:, 0] > nbr_list[:, 1]).any() and (nbr_list[:, 1] > nbr_list[:, 0]).any():
        return nbr_list
    return np.concatenate([nbr_list, nbr_list[:, ::-1]], axis=0)


def channel_index(mapping):
    """``CG2ChannelIdx`` cgvae.py:451-460: rank of atom n among atoms m < n of the same bead."""
    mapping = np.asarray(mapping, dtype=np.int64)
    order = np.argsort(mapping, kind="stable")
    sorted_map = mapping[order]
    start = np.r_[0, np.flatnonzero(np.diff(sorted_map)) + 1]
    seg_id = np.cumsum(np.r_[0, np.diff(sorted_map) != 0])
    rank_sorted = np.arange(mapping.size) - start[seg_id] if mapping.size else np.zeros(0, np.int64)
    out = np.empty_like(mapping)
    out[order] = rank_sorted
    return out


def receiver_csr(nbrs, n_nodes):
    """Receiver-major CSR of a directed edge list, edges of one receiver kept in list
    order (the order ``scatter_add_`` visits them on CPU).  returns rowptr, col, eid."""
    nbrs = np.asarray(nbrs, dtype=np.int64)
    order = np.argsort(nbrs[:, 0], kind="stable")
    counts = np.bincount(nbrs[:, 0], minlength=n_nodes)
    rowptr = np.r_[0, np.cumsum(counts)].astype(np.int64)
    return rowptr, nbrs[order, 1], order


def collate(samples):
    """``CG_collate`` data.py:255-289: offset index tensors by the cumulative atom / bead
    counts and concatenate.  ``samples``: list of dicts of numpy arrays with keys
    nxyz, CG_nxyz, CG_mapping, nbr_list, CG_nbr_list, bond_edge_list, num_atoms, num_CGs."""
    atom_off = np.cumsum([0] + [int(s["num_atoms"]) for s in samples])[:-1]
    bead_off = np.cumsum([0] + [int(s["num_CGs"]) for s in samples])[:-1]
    batch = {}
    for key in samples[0]:
        parts = []
        for s, a, b in zip(samples, atom_off, bead_off):
            val = np.asarray(s[key])
            if key in ("nbr_list", "bond_edge_list"):
                val = val + a
            elif key in ("CG_mapping", "CG_nbr_list"):
                val = val + b
            parts.append(val)
        if parts[0].ndim > 0:
            batch[key] = np.concatenate(parts, axis=0)
        else:
            batch[key] = np.stack(parts, axis=0)
    return batch
