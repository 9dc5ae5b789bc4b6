"""Sample-quality metrics of the ensemble-sampling loop on the GPU (scripts/sampling.py:120-194 ``get_bond_graphs`` /
``count_valid_graphs``, :220-239 ``compute_rmsd``, :324-333 ``eval_sample_qualities``): the reference builds 4 dense [N, N]
distance / cutoff matrices per sample on the host through ASE objects; here all samples of an ensemble are scored by two
launches (csrc/metrics.cu) with a bit-exact bond adjacency.  Inputs are tensors: ``z`` atomic numbers [N], positions [N, 3]."""
import torch

from . import _lib, ops

# covalent cutoff radii by atomic number (index 0 unused): the data of COVCUTOFFTABLE, scripts/sampling.py:12-118
COV_CUTOFF = (
    0, 0.23, 0.93, 0.68, 0.35, 0.83, 0.68, 0.68, 0.68, 0.64, 1.12, 0.97, 1.1, 1.35, 1.2, 0.75, 1.02, 0.99, 1.57,
    1.33, 0.99, 1.44, 1.47, 1.33, 1.35, 1.35, 1.34, 1.33, 1.5, 1.52, 1.45, 1.22, 1.17, 1.21, 1.22, 1.21, 1.91, 1.47,
    1.12, 1.78, 1.56, 1.48, 1.47, 1.35, 1.4, 1.45, 1.5, 1.59, 1.69, 1.63, 1.46, 1.46, 1.47, 1.4, 1.98, 1.67, 1.34,
    1.87, 1.83, 1.82, 1.81, 1.8, 1.8, 1.99, 1.79, 1.76, 1.75, 1.74, 1.73, 1.72, 1.94, 1.72, 1.57, 1.43, 1.37, 1.35,
    1.37, 1.32, 1.5, 1.5, 1.7, 1.55, 1.54, 1.54, 1.68, 1.7, 2.4, 2, 1.9, 1.88, 1.79, 1.61, 1.58, 1.55, 1.53, 1.51,
    1.5, 1.5, 1.5, 1.5, 1.5, 1.5, 1.5, 1.5, 1.57, 1.49, 1.43, 1.41,
)


def cutoff_radii(z):
    """float32 [N] radii for atomic numbers z (``compute_bond_cutoff`` builds (r_i + r_j) * scale from them, sampling.py:120-126)."""
    table = torch.tensor(COV_CUTOFF, dtype=torch.float32, device=z.device)
    return table[z.to(torch.int64)]


def get_bond_graphs(xyz, z, scale=1.3):
    """``get_bond_graphs`` (sampling.py:158-167): int64 [N, N] adjacency of one conformation (on the device)."""
    ops._need_cuda(xyz)
    lib = _lib.load()
    xyz = xyz.to(torch.float32).contiguous()
    n = xyz.shape[0]
    bond = torch.empty((n, n), dtype=torch.uint8, device=xyz.device)
    _lib.check(lib.cgvae_bond_graph(ops._p(xyz), ops._p(cutoff_radii(z)), n, float(scale), ops._p(bond), ops._stream()), "bond_graph")
    return bond.to(torch.long)


def sample_quality_counts(ref_xyz, z, samples, scale=1.3):
    """(counts int32 [S, 6], rmsd float32 [S, 2]) -- see cgvae_sample_quality in the header."""
    ops._need_cuda(ref_xyz, samples)
    lib = _lib.load()
    ref_xyz = ref_xyz.to(torch.float32).contiguous()
    samples = samples.to(torch.float32).contiguous()
    S, n = samples.shape[0], samples.shape[1]
    heavy = (z != 1).to(torch.uint8).contiguous()
    counts = torch.empty((S, 6), dtype=torch.int32, device=samples.device)
    rmsd = torch.empty((S, 2), dtype=torch.float32, device=samples.device)
    _lib.check(lib.cgvae_sample_quality(ops._p(ref_xyz), ops._p(samples), ops._p(cutoff_radii(z)), ops._p(heavy), n, S, float(scale),
                                        ops._p(counts), ops._p(rmsd), ops._stream()), "sample_quality")
    return counts, rmsd


def eval_sample_qualities(ref_xyz, z, samples, scale=1.3):
    """``eval_sample_qualities`` (sampling.py:324-333) for samples [S, N, 3] of one conformation.  Returns, like the reference,
    (all_rmsds, heavy_rmsds, valid_ratio, valid_allatom_ratio, graph_val_ratio, graph_allatom_val_ratio): the rmsd arrays hold
    [all-atom, heavy-atom] rows for the samples whose all-atom / heavy-atom bond graph equals the reference's (None when
    there is none), the ratios are fractions of valid samples and the per-sample |sum(ref - gen)| / sum(ref) lists."""
    counts, rmsd = sample_quality_counts(ref_xyz, z, samples, scale)
    counts, rmsd = counts.cpu(), rmsd.cpu().numpy()              # ONE device -> host read per ensemble
    S = counts.shape[0]
    valid_all = [i for i in range(S) if int(counts[i, 0]) == 0]
    valid_heavy = [i for i in range(S) if int(counts[i, 3]) == 0]
    # the reference divides two int64 tensors: torch's true division in float32, then .item() (sampling.py:189-190)
    ratio = lambda d, r: [float(x) for x in (d.to(torch.int64).abs() / r.to(torch.int64))]
    graph_allatom_val_ratio = ratio(counts[:, 1], counts[:, 2])
    graph_val_ratio = ratio(counts[:, 4], counts[:, 5])
    all_rmsds = rmsd[valid_all] if valid_all else None
    heavy_rmsds = rmsd[valid_heavy] if valid_heavy else None
    return all_rmsds, heavy_rmsds, len(valid_heavy) / S, len(valid_all) / S, graph_val_ratio, graph_allatom_val_ratio
