"""Synthetic inputs for the BASELINE.json configurations: jittered-lattice
conformations, CG mappings, batch dicts with the reference's keys.  Host-side data preparation only
(numpy / torch CPU): no network, no MD trajectories.  The neighbour search is injected by the caller
(``radius_fn(xyz_numpy, cutoff) -> int64 [E,2] undirected``) so the product arm can use the GPU builder
and the CPU reference arm its own.
"""
import math

import numpy as np
import torch

# name -> model / data hyper-parameters (README.md:49,55 of the reference; scripts/run_pdb.py:330-333,470-475)
CONFIGS = {
    # alanine dipeptide, 22 atoms, n_cgs 3, batch 32 (CPU-runnable reference case)
    "c1_dipeptide": dict(kind="cgvae", n_atoms=22, n_cgs=3, batch=32, spacing=1.6, n_basis=600, n_rbf=8, enc_nconv=4,
                         dec_nconv=5, atom_cutoff=8.5, cg_cutoff=9.5, beta=0.05, gamma=25.0,
                         mapping=[0] * 8 + [1] * 7 + [2] * 7, z=[6] * 6 + [1] * 12 + [7] * 2 + [8] * 2),
    # chignolin all-atom, 175 atoms, n_cgs 6, batch 2: the configuration the headline metric is quoted on
    "c2_chignolin": dict(kind="cgvae", n_atoms=175, n_cgs=6, batch=2, spacing=2.2, n_basis=600, n_rbf=10, enc_nconv=2,
                         dec_nconv=9, atom_cutoff=12.0, cg_cutoff=25.0, beta=0.05, gamma=50.0),
    # ~2000-atom protein, alpha-carbon CG (PCN / run_pdb path)
    "c4_protein": dict(kind="pcn", n_res=250, atoms_per_res=8, batch=64, spacing=5.2, n_basis=512, n_rbf=8, dec_nconv=9,
                       cg_cutoff=15.5),
    # one large graph, single message layer + radius-graph sweep
    "c5_large": dict(kind="layer", n_atoms=20000, spacing=2.2, n_basis=600, n_rbf=8, cutoffs=(4.5, 6.0, 8.5, 12.0)),
}


def _rng(seed):
    return np.random.default_rng(seed)


def random_rotation(rng):
    """proper rotation from a random unit quaternion (mirrors the per-frame rotation of datasets.py:65-71,475)."""
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]], dtype=np.float64)


def lattice_points(n, spacing, rng, rotate=True):
    """n sites of a jittered cubic lattice (jitter +-0.3 spacing) nearest the origin, ordered lexicographically so that
    contiguous index ranges are spatially compact, then rotated."""
    side = int(math.ceil(n ** (1.0 / 3.0))) + 2
    g = np.arange(side) - (side - 1) / 2.0
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    keep = np.argsort((pts ** 2).sum(1), kind="stable")[:n]
    pts = pts[np.sort(keep)]
    pts = (pts + rng.uniform(-0.3, 0.3, size=pts.shape)) * spacing
    if rotate:
        pts = pts @ random_rotation(rng).T
    return pts.astype(np.float32)


def bead_means(xyz, mapping, n_beads):
    """CG coordinates = per-bead mean of the atom coordinates (datasets.py:487)."""
    out = np.zeros((n_beads, 3), dtype=np.float64)
    np.add.at(out, mapping, xyz.astype(np.float64))
    cnt = np.maximum(np.bincount(mapping, minlength=n_beads), 1)
    return (out / cnt[:, None]).astype(np.float32)


def cgvae_sample(cfg, seed, radius_fn):
    rng = _rng(seed)
    n, n_cgs = cfg["n_atoms"], cfg["n_cgs"]
    xyz = lattice_points(n, cfg["spacing"], rng)
    mapping = np.asarray(cfg["mapping"], dtype=np.int64) if "mapping" in cfg else (np.arange(n) * n_cgs // n).astype(np.int64)
    z = np.asarray(cfg["z"], dtype=np.float32) if "z" in cfg else rng.choice([1, 6, 7, 8], size=n, p=[0.5, 0.3, 0.1, 0.1]).astype(np.float32)
    cg_xyz = bead_means(xyz, mapping, n_cgs)
    bonds = radius_fn(xyz, 1.15 * cfg["spacing"])
    if bonds.shape[0] == 0:
        bonds = np.stack([np.arange(n - 1), np.arange(1, n)], 1)
    return {
        "nxyz": torch.from_numpy(np.concatenate([z[:, None], xyz], 1)),
        "CG_nxyz": torch.from_numpy(np.concatenate([np.arange(n_cgs, dtype=np.float32)[:, None], cg_xyz], 1)),
        "CG_mapping": torch.from_numpy(mapping),
        "nbr_list": torch.from_numpy(np.asarray(radius_fn(xyz, cfg["atom_cutoff"]), dtype=np.int64)),
        "CG_nbr_list": torch.from_numpy(np.asarray(radius_fn(cg_xyz, cfg["cg_cutoff"]), dtype=np.int64)),
        "bond_edge_list": torch.from_numpy(np.asarray(bonds, dtype=np.int64)),
        "num_atoms": torch.tensor(n), "num_CGs": torch.tensor(n_cgs),
    }


def cgvae_batch(cfg, step, radius_fn, collate_fn, config_id=1):
    """one collated batch of cfg['batch'] conformations; seeds 1234 + 100*config + conformation."""
    base = 1234 + 100 * config_id + step * cfg["batch"]
    return collate_fn([cgvae_sample(cfg, base + k, radius_fn) for k in range(cfg["batch"])])


def pcn_batch(cfg, step, radius_fn, n_proteins=None, config_id=4):
    """SCNCG_collate-shaped batch (data.py:368-398): keys xyz, ca_xyz, res, cg_map, bond_edge_list, CG_nbr_list, seq, ca_idx,
    dihe_idxs."""
    n_prot = cfg["batch"] if n_proteins is None else n_proteins
    n_res, per = cfg["n_res"], cfg["atoms_per_res"]
    xyz_l, ca_l, res_l, map_l, bond_l, nbr_l, ca_idx_l, seqs, dihe_l = [], [], [], [], [], [], [], [], []
    a_off = r_off = 0
    for p in range(n_prot):
        rng = _rng(1234 + 100 * config_id + step * n_prot + p)
        ca = lattice_points(n_res, cfg["spacing"], rng)
        mapping = np.repeat(np.arange(n_res), per)
        xyz = ca[mapping] + rng.normal(scale=1.5, size=(n_res * per, 3)).astype(np.float32)
        ca_idx = np.arange(n_res) * per + 1
        xyz[ca_idx] = ca
        chain = np.arange(n_res * per - 2)
        bonds = np.concatenate([np.stack([chain, chain + 1], 1), np.stack([chain, chain + 2], 1)], 0)
        xyz_l.append(xyz); ca_l.append(ca); res_l.append(rng.integers(0, 20, size=n_res))
        map_l.append(mapping + r_off); bond_l.append(bonds + a_off)
        nbr_l.append(np.asarray(radius_fn(ca, cfg["cg_cutoff"]), dtype=np.int64) + r_off)
        ca_idx_l.append(ca_idx + a_off); seqs.append("A" * n_res)
        # backbone-like dihedrals: (N, CA, C) of a residue and the N of the next one, and the shifted quadruple
        r = np.arange(n_res - 1) * per
        dihe_l.append(np.concatenate([np.stack([r, r + 1, r + 2, r + per], 1), np.stack([r + 1, r + 2, r + per, r + per + 1], 1)], 0) + a_off)
        a_off += n_res * per
        r_off += n_res
    cat = lambda xs, dt: torch.from_numpy(np.concatenate(xs, 0).astype(dt))
    return {"xyz": cat(xyz_l, np.float32), "ca_xyz": cat(ca_l, np.float32), "res": cat(res_l, np.int64),
            "cg_map": cat(map_l, np.int64), "bond_edge_list": cat(bond_l, np.int64), "CG_nbr_list": cat(nbr_l, np.int64),
            "seq": seqs, "ca_idx": cat(ca_idx_l, np.int64), "dihe_idxs": cat(dihe_l, np.int64)}
