"""TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu arm) -- never imported by the product.

CPU restatement of the sample-quality metrics of the reference's sampling loop (scripts/sampling.py), on plain arrays
instead of ASE ``Atoms``.  Pinned by tests/golden/sample_quality.npz, which tests/golden/make_golden.py produced by executing the
reference's own functions (lifted from the unmodified source) on the same inputs."""
import numpy as np
import torch


def bond_cutoff(radii, scale=1.3):
    """compute_bond_cutoff, sampling.py:120-126: (r[None, :] + r[:, None]) * scale in float32."""
    r = torch.as_tensor(radii, dtype=torch.float32)
    return (r[None, :] + r[:, None]) * scale


def distance_mat(xyz):
    """compute_distance_mat, sampling.py:128-133: float32 positions, sqrt of the summed squared differences."""
    x = torch.as_tensor(np.asarray(xyz), dtype=torch.float32)
    return (x[:, None, :] - x[None, :, :]).pow(2).sum(-1).sqrt()


def get_bond_graphs(xyz, radii, scale=1.3):
    """get_bond_graphs, sampling.py:158-167: dist < cutoff with the diagonal cleared, int64."""
    bond = distance_mat(xyz) < bond_cutoff(radii, scale)
    bond.fill_diagonal_(False)
    return bond.to(torch.long)


def count_valid_graphs(ref_xyz, samples, z, radii, heavy_only=True, scale=1.3):
    """count_valid_graphs, sampling.py:171-196 (dropH, sampling.py:135-146, = the z != 1 sub-block)."""
    z = np.asarray(z)
    keep = (z != 1) if heavy_only else np.ones(len(z), dtype=bool)
    r = np.asarray(radii)[keep]
    ref_graph = get_bond_graphs(np.asarray(ref_xyz)[keep], r, scale)
    valid, ratios = [], []
    for i, x in enumerate(samples):
        gen = get_bond_graphs(np.asarray(x)[keep], r, scale)
        if int((gen != ref_graph).sum()) == 0:                       # compare_graph, sampling.py:148-156
            valid.append(i)
        ratios.append(float((ref_graph - gen).sum().abs() / ref_graph.sum()))
    return valid, len(valid) / len(samples), ratios


def compute_rmsd(ref_xyz, samples, z, valid_ids):
    """compute_rmsd, sampling.py:220-239: [all-atom, heavy-atom] RMSD (no alignment) of the valid samples, float64."""
    z = np.asarray(z)
    heavy = z != 1
    rows = []
    for i, x in enumerate(samples):
        d = np.asarray(x, dtype=np.float64) - np.asarray(ref_xyz, dtype=np.float64)
        aa = np.sqrt(np.power(d, 2).sum(-1).mean())
        hv = np.sqrt(np.power(d[heavy], 2).sum(-1).mean())
        if i in valid_ids:
            rows.append([aa, hv])
    return np.array(rows) if len(valid_ids) else None


def eval_sample_qualities(ref_xyz, samples, z, radii, scale=1.3):
    """eval_sample_qualities, sampling.py:324-333."""
    valid_ids, valid_ratio, graph_val_ratio = count_valid_graphs(ref_xyz, samples, z, radii, True, scale)
    valid_all_ids, valid_allatom_ratio, graph_allatom_val_ratio = count_valid_graphs(ref_xyz, samples, z, radii, False, scale)
    heavy_rmsds = compute_rmsd(ref_xyz, samples, z, valid_ids)
    all_rmsds = compute_rmsd(ref_xyz, samples, z, valid_all_ids)
    return all_rmsds, heavy_rmsds, valid_ratio, valid_allatom_ratio, graph_val_ratio, graph_allatom_val_ratio
