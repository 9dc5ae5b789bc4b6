"""GPU (B200), round 2: the rows the round-1 review found unpinned -- ensemble sampling against the REAL reference's
loop, ``CGDataset.generate_neighbor_list``, the skip guard / validation branch of the training loop, index validation
(the reference's IndexError at the lifting gather), bidirectional edge lists in static (CUDA-graph) mode -- and the
tensor-core message kernels."""
import copy

import numpy as np
import pytest
import torch

import coarsegrainingvae_b200 as cg
import coarsegrainingvae_b200 as cg_mod
from coarsegrainingvae_b200 import ops, synthetic
from oracle import cgvae_oracle as orc
from oracle import graph_oracle as gorc
from tests import parity_cases as pc
from tests.golden_util import load, rel_err, section

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _to(batch, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}


def _gpu_radius(xyz, cutoff):
    return ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=DEV), cutoff).cpu().numpy()


# ------------------------------------------------------------------------------------------ sampling (a25)

@pytest.mark.parametrize("tag", ["sym", "nosym"])
def test_sampling_matches_reference_loop(tag):
    """train.sample_single and train.GraphedSampler reproduce, member by member and for the same noise, the geometries
    the REAL reference produces through the loop of scripts/sampling.py:265-293 (tests/golden/sampling_small.npz)."""
    pc.sampling_case(DEV, tag, torch.float32, TOL, gold_tol=5e-5, graphed=True)


# ------------------------------------------------------------------------------------------ dataset graph builder (a2)

@pytest.mark.parametrize("tag,cg_cut,und", [("radius", 4.0, True), ("radius_dir", 4.0, False), ("bond", None, True)])
def test_generate_neighbor_list_matches_reference(tag, cg_cut, und):
    """CGDataset.generate_neighbor_list (one batched kernel call per graph kind, split back with searchsorted) == the real
    reference's per-frame python loop (data.py:207-252) on ragged frames, bit for bit; ``cg_cutoff=None``: bond graph."""
    z = load("dataset_lists.npz")
    sizes = [int(x) for x in z["meta/sizes"]]
    keys = ("nxyz", "CG_nxyz", "CG_mapping", "num_atoms", "num_CGs", "bond_edge_list")
    props = {k: [torch.from_numpy(np.asarray(z["props/%d/%s" % (i, k)])) for i in range(len(sizes))] for k in keys}
    ds = cg.CGDataset(props)
    ds.generate_neighbor_list(atom_cutoff=float(z["meta/atom_cutoff"]), cg_cutoff=cg_cut, device=DEV, undirected=und)
    for i in range(len(sizes)):
        for k in ("nbr_list", "CG_nbr_list"):
            got = ds.props[k][i]
            want = torch.from_numpy(np.asarray(z["%s/%d/%s" % (tag, i, k)]))
            assert got.dtype == torch.int64 and got.device.type == "cpu"
            assert got.shape == want.shape and torch.equal(got, want), (tag, i, k)


def test_generate_neighbor_list_with_empty_frames():
    g = torch.Generator().manual_seed(2)
    frames = [torch.cat([torch.ones(n, 1), torch.randn(n, 3, generator=g) * s], 1) for n, s in ((6, 1.0), (1, 1.0), (9, 40.0), (4, 0.5))]
    props = {"nxyz": frames, "CG_nxyz": [f[: max(1, f.shape[0] // 2)] for f in frames]}
    ds = cg.CGDataset(props)
    ds.generate_neighbor_list(atom_cutoff=2.0, cg_cutoff=2.0, device=DEV)
    for k, src in (("nbr_list", props["nxyz"]), ("CG_nbr_list", props["CG_nxyz"])):
        for f, got in zip(src, ds.props[k]):
            want = gorc.radius_graph(f[:, 1:4].numpy(), 2.0, True)
            assert np.array_equal(got.numpy(), want)


# ------------------------------------------------------------------------------------------ loop control flow (a24)

def _small_trainer(lr=1e-3, **kw):
    from coarsegrainingvae_b200.factory import build_cgvae
    from coarsegrainingvae_b200.train import TrainStep
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(batch=3, n_basis=64, enc_nconv=2, dec_nconv=2)
    raw = synthetic.cgvae_batch(cfg, 0, _gpu_radius, cg.CG_collate)
    torch.manual_seed(5)
    model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], 2, 2, cfg["atom_cutoff"], cfg["cg_cutoff"], 3).to(DEV)
    eps = torch.randn(9, 64, device=DEV)
    tr = TrainStep(model, cfg["beta"], cfg["gamma"], lr=lr, **kw)
    return cfg, raw, model, eps, tr


def test_skip_guard_and_validation_mode_match_reference_loop():
    """scripts/utils.py:145-160 on the device: a loss >= 200*gamma or NaN leaves parameters, moments and the step
    counter untouched (and is counted); the validation branch runs backward without an optimiser step."""
    cfg, raw, model, eps, tr = _small_trainer()
    batch = _to(raw, DEV)
    tr.prepare(batch, eps)
    snap = lambda: (tr.flat_p.clone(), tr.exp_avg.clone(), tr.exp_avg_sq.clone(), float(tr.step_count))
    p0 = snap()
    loss = tr.step(batch, eps, train=False)                     # validation: loss.backward() only
    assert orc.train_loop_step(loss, cfg["gamma"], train=False) == (True, False)
    p1 = snap()
    assert all(torch.equal(a, b) for a, b in zip(p0[:3], p1[:3])) and p1[3] == p0[3]
    assert float(tr.flat.flat.abs().max()) > 0                  # gradients were produced
    tr.loss_limit = float(loss) * 0.5                           # past the guard: skipped on the device
    tr.step(batch, eps)
    p2 = snap()
    assert all(torch.equal(a, b) for a, b in zip(p0[:3], p2[:3])) and p2[3] == p0[3] and tr.skipped_steps() == 1
    tr.loss_limit = cfg["gamma"] * 200.0                        # the reference's limit: the step is taken
    loss = tr.step(batch, eps)
    assert orc.train_loop_step(loss, cfg["gamma"]) == (True, True)
    p3 = snap()
    assert not torch.equal(p3[0], p0[0]) and p3[3] == p0[3] + 1 and tr.skipped_steps() == 1
    # NaN loss (poisoned input coordinates): no update, parameters stay finite
    bad = dict(batch)
    bad["nxyz"] = batch["nxyz"].clone()
    bad["nxyz"][0, 1] = float("nan")
    lnan = tr.step(bad, eps)
    assert orc.train_loop_step(lnan, cfg["gamma"]) == (False, False)
    p4 = snap()
    assert all(torch.equal(a, b) for a, b in zip(p3[:3], p4[:3])) and p4[3] == p3[3] and tr.skipped_steps() == 2
    assert bool(torch.isfinite(tr.flat_p).all())
    tr.flat.release()


def test_skip_guard_inside_cuda_graph_replay():
    """the guard is data-dependent but takes no host read: the SAME captured graph skips a poisoned batch and steps on a
    good one."""
    from coarsegrainingvae_b200.train import GraphedTrainStep, to_static_batch
    cfg, raw, model, eps, tr = _small_trainer(capturable=True)
    caps = {"nbr_list": 3 * 22 * 21 // 2, "CG_nbr_list": 9, "bond_edge_list": raw["bond_edge_list"].shape[0] + 8}
    good = _to(to_static_batch(raw, caps), DEV)
    tr.prepare(good, eps)
    g = GraphedTrainStep(tr, good, eps)
    bad = dict(good)
    bad["nxyz"] = good["nxyz"].clone()
    bad["nxyz"][3, 2] = float("nan")
    before, steps = tr.flat_p.clone(), float(tr.step_count)
    g.step(bad)
    assert torch.equal(tr.flat_p, before) and float(tr.step_count) == steps and tr.skipped_steps() == 1
    g.step(good)
    assert not torch.equal(tr.flat_p, before) and float(tr.step_count) == steps + 1 and tr.skipped_steps() == 1
    tr.flat.release()


def test_fused_adam_nonfinite_norm_is_a_noop():
    n = 4096 + 3
    g = torch.Generator().manual_seed(0)
    p, grad = torch.randn(n, generator=g).to(DEV), torch.randn(n, generator=g).to(DEV)
    m, v, step, skipped = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    grad[17] = float("inf")
    p0 = p.clone()
    ops.adam_clip_step(p, grad, m, v, step, 0.01, 1e-3, skipped=skipped)
    assert torch.equal(p, p0) and float(m.abs().max()) == 0 and float(v.abs().max()) == 0
    assert float(step) == 0 and float(skipped) == 1
    grad[17] = 0.5
    ops.adam_clip_step(p, grad, m, v, step, 0.01, 1e-3, skipped=skipped)
    assert not torch.equal(p, p0) and float(step) == 1 and float(skipped) == 1


# ------------------------------------------------------------------------------------------ index validation

def test_bead_with_more_atoms_than_channels_raises_like_the_reference():
    """cg_v[mapping, CG2atomChannel] raises IndexError in the reference when a bead holds more atoms than channels
    (cgvae.py:473); here the lifting kernels never touch memory out of range and the forward raises the same error."""
    from coarsegrainingvae_b200.factory import build_cgvae
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(batch=2, n_basis=8, n_rbf=4, enc_nconv=1, dec_nconv=1)       # 8 atoms in bead 0, F = 8: just fits
    torch.manual_seed(1)
    model = build_cgvae(8, 4, 1, 1, cfg["atom_cutoff"], cfg["cg_cutoff"], 3).to(DEV)
    batch = _to(synthetic.cgvae_batch(cfg, 0, _gpu_radius, cg.CG_collate), DEV)
    eps = torch.randn(6, 8, device=DEV)
    out = model(batch, eps=eps)
    assert bool(torch.isfinite(out[5]).all())
    bad = dict(batch)
    bad["CG_mapping"] = batch["CG_mapping"].clone()
    bad["CG_mapping"][8] = 0                                                # ninth atom of bead 0
    with pytest.raises(IndexError):
        model(bad)
    bad = dict(batch)
    bad["nxyz"] = batch["nxyz"].clone()
    bad["nxyz"][0, 0] = 250.0                                               # atomic number outside nn.Embedding(100, F)
    with pytest.raises(IndexError):
        model(bad)
    out2 = model(batch, eps=eps)                                            # the flag is cleared: good batches still run
    assert torch.equal(out2[5], out[5])


# ------------------------------------------------------------------------------------------ static mode, both-direction lists

def test_graphed_step_with_bond_derived_cg_graph_matches_eager():
    """a CG list that already holds both directions (cg_cutoff=None: data.bond_cg_graph) must not be doubled by the
    static-capacity CSR builder: graph replay == eager on the same batches."""
    from coarsegrainingvae_b200.data import bond_cg_graph
    from coarsegrainingvae_b200.factory import build_cgvae
    from coarsegrainingvae_b200.train import GraphedTrainStep, TrainStep, to_static_batch
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(batch=3, n_basis=64, enc_nconv=2, dec_nconv=2)
    raws = []
    for i in range(3):
        samples = [synthetic.cgvae_sample(cfg, 900 + 10 * i + k, _gpu_radius) for k in range(cfg["batch"])]
        for s_ in samples:
            s_["CG_nbr_list"] = bond_cg_graph(s_["bond_edge_list"], s_["CG_mapping"], 3)
            assert s_["CG_nbr_list"].shape[0] > 0
        raws.append(cg.CG_collate(samples))
    caps = {"nbr_list": 3 * 22 * 21 // 2, "CG_nbr_list": 3 * 6, "bond_edge_list": max(b["bond_edge_list"].shape[0] for b in raws) + 8}
    static = [to_static_batch(b, caps) for b in raws]
    assert static[0]["CG_nbr_symmetrize"] is False and static[0]["nbr_symmetrize"] is True
    static = [_to(b, DEV) for b in static]
    torch.manual_seed(5)
    model_a = build_cgvae(64, cfg["n_rbf"], 2, 2, cfg["atom_cutoff"], cfg["cg_cutoff"], 3).to(DEV)
    model_b = copy.deepcopy(model_a)
    eps = torch.randn(9, 64, device=DEV)
    eager = TrainStep(model_a, cfg["beta"], cfg["gamma"], lr=1e-3)
    eager.prepare(_to(raws[0], DEV), eps)
    gtr = TrainStep(model_b, cfg["beta"], cfg["gamma"], lr=1e-3, capturable=True)
    gtr.prepare(static[0], eps)
    graphed = GraphedTrainStep(gtr, static[0], eps)
    for _ in range(3):
        eager.step(_to(raws[0], DEV), eps)
    for i in (1, 2, 0):
        la = eager.step(_to(raws[i], DEV), eps)
        lb = graphed.step(static[i])
        assert rel_err(lb, la) < 1e-5, (i, float(la), float(lb))
    gtr.flat.release()
    eager.flat.release()


# ------------------------------------------------------------------------------------------ tensor-core message kernels

NB = 64        # columns per batch record (csrc/message_tc.cu)


def _tile_reference(rowptr, col, basis, unit, rc):
    """numpy restatement of the column-tile layout (cgvae_msg_tiles_build): per chunk of rc rows, the sorted union of the
    partner nodes; column (g, rr) = edge (row chunk*rc+rr <- partner union[g]); NB columns per batch."""
    n_rows = len(rowptr) - 1
    gb = NB // rc
    batches = []          # per batch: (gcol list, dense [32, RB] basis, dense [32, 3] unit)
    bptr = [0]
    ngroups = []
    for c0 in range(0, n_rows, rc):
        rows = range(c0, min(n_rows, c0 + rc))
        union = sorted({int(col[e]) for i in rows for e in range(rowptr[i], rowptr[i + 1])})
        rank = {j: g for g, j in enumerate(union)}
        nb = (len(union) + gb - 1) // gb
        local = [([0] * gb, np.zeros((NB, basis.shape[1]), np.float32), np.zeros((NB, 3), np.float32)) for _ in range(nb)]
        for g, j in enumerate(union):
            local[g // gb][0][g % gb] = j
        for i in rows:
            for e in range(rowptr[i], rowptr[i + 1]):
                g = rank[int(col[e])]
                c = (g % gb) * rc + (i - c0)
                local[g // gb][1][c] = basis[e]
                local[g // gb][2][c] = unit[e, :3]
        batches += local
        bptr.append(bptr[-1] + nb)
        ngroups.append(len(union))
    return bptr, ngroups, batches


@pytest.mark.parametrize("rc", [4, 8, 16])
def test_message_tiles_layout_exact(rc):
    """the column tiles the tensor-core kernels consume: partner ids, batch offsets and group counts bit-exact, tf32 hi + lo
    == the fp32 basis to 2^-21, zero rows for non-edges, unit vectors exact -- forward (receiver chunks) and backward
    (sender chunks through perm_t) tiles of a ragged, non-symmetric graph."""
    n, R, cutoff = 70, 10, 6.0
    g = torch.Generator().manual_seed(5 + rc)
    xyz = torch.rand(n, 3, generator=g) * 6.0
    pairs = torch.randint(0, n, (900, 2), generator=g)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    pairs = torch.unique(pairs, dim=0)
    pairs = pairs[torch.randperm(pairs.shape[0], generator=g)]             # arbitrary edge-list order
    pairs = pairs[(pairs[:, 0] < 50) | (pairs[:, 0] > 58)]                   # a run of receivers without edges
    graph = ops.build_graph(pairs.to(DEV), n)
    geom = ops.edge_geometry(graph, xyz.to(DEV), xyz.to(DEV), R, cutoff)
    basis, unit = geom.basis.cpu().numpy(), geom.unit.cpu().numpy()
    for transposed in (False, True):
        tiles = ops.message_tiles(geom, transposed, rc)
        if transposed:
            rowptr, col = graph.rowptr_t.cpu().numpy(), graph.col_t.cpu().numpy()
            perm = graph.perm_t.cpu().numpy()
            b_src, u_src = basis[perm], unit[perm]
        else:
            rowptr, col, b_src, u_src = graph.rowptr.cpu().numpy(), graph.col.cpu().numpy(), basis, unit
        bptr, ngroups, batches = _tile_reference(rowptr, col, b_src, u_src, rc)
        assert tiles.bptr.cpu().tolist() == bptr and tiles.ngroups.cpu().tolist() == ngroups
        tile = NB * 16 * 4
        rec_bytes = (2 * tile + 3 * NB * 4 + (NB // 4) * 4 + 127) // 128 * 128
        rec = tiles.rec.cpu().numpy().reshape(-1, rec_bytes)
        gb = NB // rc
        for b, (gcol, bb, uu) in enumerate(batches):
            r = rec[b]
            hi = r[0:tile].view(np.float32).reshape(NB // 8, 4, 8, 4)      # [row group][k chunk][row in group][k in chunk]
            lo = r[tile:2 * tile].view(np.float32).reshape(NB // 8, 4, 8, 4)
            hi = hi.transpose(0, 2, 1, 3).reshape(NB, 16)
            lo = lo.transpose(0, 2, 1, 3).reshape(NB, 16)
            want = np.zeros((NB, 16), np.float32)
            want[:, :bb.shape[1]] = bb
            assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)   # tf32 values
            assert np.abs((hi.astype(np.float64) + lo) - want).max() <= 2.0 ** -21 * max(1e-30, np.abs(want).max())
            assert np.all(hi[np.all(want == 0, axis=1)] == 0)
            un = r[2 * tile:2 * tile + 12 * NB].view(np.float32).reshape(3, NB).T
            assert np.array_equal(un, uu)
            got_gcol = r[2 * tile + 12 * NB:2 * tile + 12 * NB + 4 * gb].view(np.int32).tolist()
            assert got_gcol == gcol, (b, got_gcol, gcol)


def _tc_vs_simt_case(n_split, rc, n, F, R, cutoff, seed, v_zero=False, residual=True, sym=True):
    g = torch.Generator().manual_seed(seed)
    xyz = synthetic.lattice_points(n, 2.2, np.random.default_rng(seed), rotate=False)
    half = ops.radius_graph(torch.as_tensor(xyz, device=DEV), cutoff, True)
    pairs = torch.cat([half, half.flip(1)], 0) if sym else half
    graph = ops.build_graph(pairs, n)
    x = torch.as_tensor(xyz, device=DEV)
    geom = ops.edge_geometry(graph, x, x, R, cutoff)
    phi = torch.randn(n, n_split, F, generator=g).to(DEV)
    v = None if v_zero else torch.randn(n, 3, F, generator=g).to(DEV)
    Wf = (torch.randn(n_split * F, R, generator=g) * 0.5).to(DEV)
    bf = (torch.randn(n_split * F, generator=g) * 0.5).to(DEV)
    res_s = torch.randn(n, F, generator=g).to(DEV) if residual else None
    res_v = torch.randn(n, 3, F, generator=g).to(DEV) if residual else None
    old_tc, old_rc = ops.MSG_TC, ops.MSG_TC_RC
    try:
        ops.MSG_TC = "0"
        want = ops.message_fwd(n_split, phi, v, v if n_split == 4 else None, geom, Wf, bf, res_s, res_v, want_q=n_split == 4)
        ops.MSG_TC, ops.MSG_TC_RC = "1", rc
        got = ops.message_fwd(n_split, phi, v, v if n_split == 4 else None, geom, Wf, bf, res_s, res_v, want_q=n_split == 4)
        again = ops.message_fwd(n_split, phi, v, v if n_split == 4 else None, geom, Wf, bf, res_s, res_v, want_q=n_split == 4)
    finally:
        ops.MSG_TC, ops.MSG_TC_RC = old_tc, old_rc
    # float64 ground truth of the layer from the SAME per-edge records (basis, unit): conv.py:505-563 / 358-402
    col = graph.col.long()
    recv = torch.repeat_interleave(torch.arange(n, device=DEV), (graph.rowptr[1:] - graph.rowptr[:-1]).long())
    E = int(graph.rowptr[-1])
    col, recv = col[:E], recv[:E]
    B = geom.basis[:E].double()
    Wfull = torch.zeros(n_split * F, geom.rb, dtype=torch.float64, device=DEV)
    Wfull[:, :R] = Wf.double()
    Wfull[:, R] = bf.double()
    w = (B @ Wfull.t()).view(E, n_split, F)
    m = phi.double()[col] * w
    u = geom.unit[:E, :3].double()
    ds = torch.zeros(n, F, dtype=torch.float64, device=DEV).index_add_(0, recv, m[:, 1])
    dv_e = m[:, 2, None, :] * u[:, :, None]
    if v is not None:
        dv_e = dv_e + m[:, 0, None, :] * v.double()[col]
        if n_split == 4:
            dv_e = dv_e + m[:, 3, None, :] * torch.linalg.cross(v.double()[recv], v.double()[col], dim=1)
    dv = torch.zeros(n, 3, F, dtype=torch.float64, device=DEV).index_add_(0, recv, dv_e)
    if residual:
        ds, dv = ds + res_s.double(), dv + res_v.double()
    for a, b, name in ((got[0], ds, "s"), (got[1], dv, "v")):
        assert rel_err(a, b) < TOL, (name, rel_err(a, b))
    assert rel_err(got[0], want[0]) < TOL and rel_err(got[1], want[1]) < TOL
    if n_split == 4 and v is not None:
        assert rel_err(got[2], want[2]) < TOL
    assert all(torch.equal(a, b) for a, b in zip(got[:2], again[:2]))       # deterministic
    return E


@pytest.mark.parametrize("n_split,rc,n,F,R,cutoff", [
    (3, 8, 350, 600, 10, 12.0),      # chignolin-like density, F not a multiple of 128
    (3, 4, 300, 128, 8, 5.0),
    (3, 16, 200, 64, 8, 7.0),        # F < 128: partial channel slice
    (4, 8, 250, 512, 8, 7.0),        # cross block at the PCN width
    (4, 4, 97, 96, 4, 4.0),          # node count not a multiple of the chunk
])
def test_message_tc_forward_matches_simt_and_float64(n_split, rc, n, F, R, cutoff):
    """cgvae_message_tc_fwd (filter on tcgen05, column tiles) == cgvae_message_fwd (fp32 SIMT) == float64 ground truth within
    1e-5 scale-relative, bitwise repeatable."""
    E = _tc_vs_simt_case(n_split, rc, n, F, R, cutoff, seed=3 + n_split + rc)
    assert E > 0


def test_message_tc_forward_first_layer_and_delta_modes():
    _tc_vs_simt_case(3, 8, 180, 192, 10, 9.0, seed=11, v_zero=True, residual=True)
    _tc_vs_simt_case(4, 8, 180, 192, 8, 9.0, seed=12, v_zero=True, residual=False)
    _tc_vs_simt_case(3, 8, 180, 192, 8, 9.0, seed=13, v_zero=False, residual=False, sym=False)   # one-directional list


@pytest.mark.parametrize("tag,cls", [("k3", "EquiMessageBlock"), ("k4", "EquiMessageCross")])
def test_message_block_api_on_tensor_cores(tag, cls):
    """the frozen reference block outputs / gradients (tests/golden/blocks_small.npz) with the tensor-core path forced"""
    old = ops.MSG_TC
    try:
        ops.MSG_TC = "1"
        pc.message_block_api(DEV, tag, cls, torch.float32, TOL)
    finally:
        ops.MSG_TC = old


# ------------------------------------------------------------------------------------------ fused loss / latent kernels

def _loss_inputs(seed, n_atoms, n_beads, F, n_bonds, dtype, dev):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.randn(n_atoms, 3, generator=g) * 2
    rec = xyz + 0.3 * torch.randn(n_atoms, 3, generator=g)
    a = torch.randint(0, n_atoms, (n_bonds,), generator=g)
    b = (a + 1 + torch.randint(0, n_atoms - 1, (n_bonds,), generator=g)) % n_atoms
    bonds = torch.stack([a, b], 1)
    mu, pmu = torch.randn(n_beads, F, generator=g), torch.randn(n_beads, F, generator=g)
    sigma = 0.2 + torch.rand(n_beads, F, generator=g)
    pstd = 0.2 + torch.rand(n_beads, F, generator=g)
    cast = lambda t: t.to(dtype).to(dev)
    return cast(xyz), cast(rec), bonds.to(dev), cast(mu), cast(sigma), cast(pmu), cast(pstd)


@pytest.mark.parametrize("mode", ["full", "std_normal_prior", "no_kl", "no_bonds", "dp_norms", "static_bonds", "tiny"])
def test_fused_training_loss_matches_reference_expression(mode):
    """csrc/loss.cu (two launches each way) == the loop's loss (scripts/utils.py:81-141) evaluated in float64 by the
    oracle's expression: value, the three components and the gradients of every input, <= 1e-5."""
    from coarsegrainingvae_b200 import train
    n_atoms, n_beads, F, n_bonds = (5, 2, 3, 4) if mode == "tiny" else (3001, 411, 96, 2950)
    beta, gamma = 0.05, (0.0 if mode == "no_bonds" else 25.0)
    xyz, rec, bonds, mu, sigma, pmu, pstd = _loss_inputs(3, n_atoms, n_beads, F, n_bonds, torch.float32, DEV)
    if mode == "std_normal_prior":
        pmu = pstd = None
    if mode == "no_kl":
        mu = sigma = pmu = pstd = None
    norms = torch.tensor([n_atoms * 1.25, n_beads * 0.75, n_bonds * 1.5], device=DEV) if mode == "dp_norms" else None
    bond_count = None
    bonds_in = bonds
    if mode == "static_bonds":
        bonds_in = torch.cat([bonds, torch.zeros(200, 2, dtype=torch.int64, device=DEV)])
        bond_count = torch.tensor(n_bonds, dtype=torch.int64, device=DEV)

    def run(fused, dtype):
        train.FUSED_LOSS = fused
        try:
            leaves = [t.detach().to(dtype).requires_grad_(True) if t is not None else None for t in (rec, mu, sigma, pmu, pstd)]
            out = (leaves[1], leaves[2], leaves[3], leaves[4], None, leaves[0])
            loss, r, k, gph = train.training_loss(out, xyz.to(dtype), bonds_in, beta, gamma, bond_count=bond_count,
                                                  norms=norms.to(dtype) if norms is not None else None)
            (loss * 1.7).backward()
            return [loss, r, k, gph], [t.grad if t is not None else None for t in leaves]
        finally:
            train.FUSED_LOSS = True

    got_v, got_g = run(True, torch.float32)
    want_v, want_g = run(False, torch.float64)
    for name, a, b in zip(("loss", "recon", "kl", "graph"), got_v, want_v):
        assert (a is None) == (b is None), name
        if a is not None:
            assert rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy()) <= TOL, name
    for name, a, b in zip(("g_rec", "g_mu", "g_sigma", "g_pmu", "g_pstd"), got_g, want_g):
        assert (a is None) == (b is None), name
        if a is not None:
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= TOL, name
    # deterministic: fixed-order reductions, per-atom gather backward
    again_v, again_g = run(True, torch.float32)
    assert torch.equal(again_v[0], got_v[0]) and torch.equal(again_g[0], got_g[0])


def test_vae_latent_and_prior_std_kernels():
    """(z, sigma) of cgvae.py:445-449,500-507 and the prior std of cgvae.py:401: values and gradients vs float64 torch."""
    from coarsegrainingvae_b200 import functions as fn
    g = torch.Generator().manual_seed(0)
    mu, lv, eps = (torch.randn(777, 36, generator=g).to(DEV) for _ in range(3))
    a, b = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    z, sigma = fn.VAELatent.apply(a, b, eps)
    w1, w2 = torch.randn_like(z), torch.randn_like(z)
    ((z * w1).sum() + (sigma * w2).sum()).backward()
    a64, b64 = mu.double().requires_grad_(True), lv.double().requires_grad_(True)
    s64 = 1e-12 + torch.exp(b64 / 2)
    z64 = eps.double() * s64 + a64
    ((z64 * w1.double()).sum() + (s64 * w2.double()).sum()).backward()
    for got, want in ((z, z64), (sigma, s64), (a.grad, a64.grad), (b.grad, b64.grad)):
        assert rel_err(got.detach().cpu().numpy(), want.detach().cpu().numpy()) <= 1e-6
    x = lv.clone().requires_grad_(True)
    y = fn.StdLogvar.apply(x, 1e-9)
    (y * w1).sum().backward()
    x64 = lv.double().requires_grad_(True)
    y64 = 1e-9 + torch.exp(x64 / 2)
    (y64 * w1.double()).sum().backward()
    assert rel_err(y.detach().cpu().numpy(), y64.detach().cpu().numpy()) <= 1e-6
    assert rel_err(x.grad.cpu().numpy(), x64.grad.cpu().numpy()) <= 1e-6
    f = ops.fill((1000, 7), 1.0, mu)
    assert f.shape == (1000, 7) and bool((f == 1.0).all())


# ------------------------------------------------------------------------------------------ PCN: dihedral loss, device predicate, graph

def test_dihedral_loss_kernel_matches_reference_golden():
    """csrc/loss.cu dihedral kernels (analytic backward, per-atom CSR gather) == the reference's compute_dihe + autograd
    (tests/golden/pcn_dihedral.npz, generated by executing scripts/pcn_utils.py:114-132 from the reference source)."""
    from coarsegrainingvae_b200 import train
    z = load("pcn_dihedral.npz")
    xyz = torch.from_numpy(np.asarray(z["xyz"])).to(DEV)
    idx = torch.from_numpy(np.asarray(z["idx"])).to(DEV)
    rec = torch.from_numpy(np.asarray(z["xyz_rec"])).to(DEV).requires_grad_(True)
    loss = train.dihedral_loss(rec, xyz, idx)
    (loss * 1.0).backward()
    assert rel_err(loss, z["loss"]) <= TOL and rel_err(rec.grad, z["g_xyz_rec"]) <= TOL
    # float64 torch expression on a larger random case + static-capacity (padded) list + a data-parallel denominator
    g = torch.Generator().manual_seed(5)
    n, D = 4000, 7001
    x = (torch.randn(n, 3, generator=g) * 3).to(DEV)
    r0 = (x.cpu() + 0.3 * torch.randn(n, 3, generator=g)).to(DEV)
    q = torch.stack([torch.randperm(n, generator=g)[:4] for _ in range(D)]).to(DEV)
    r = r0.clone().requires_grad_(True)
    pad = torch.cat([q, torch.zeros(99, 4, dtype=torch.int64, device=DEV)])
    cnt = torch.tensor([D], dtype=torch.int64, device=DEV)
    norm = torch.tensor([D * 1.5], device=DEV)
    l1 = train.dihedral_loss(r, x, pad, count=cnt, norm=norm)
    l1.backward()
    r64 = r0.double().requires_grad_(True)
    l64 = (train.compute_dihe(r64, q) - train.compute_dihe(x.double(), q)).pow(2).sum() / (D * 1.5)
    l64.backward()
    assert rel_err(l1, l64) <= TOL and rel_err(r.grad, r64.grad) <= TOL
    r2 = r0.clone().requires_grad_(True)
    train.dihedral_loss(r2, x, pad, count=cnt, norm=norm).backward()
    assert torch.equal(r2.grad, r.grad)                      # deterministic


def test_pcn_pin_mask_device_predicate():
    """cgvae.py:569-571 without the host read: mask set iff ca_idx[-1] < n_atoms; live count honoured; bad index flagged."""
    idx = torch.tensor([1, 9, 17, 25], device=DEV)
    m = ops.pin_mask(idx, 30).cpu().numpy()
    assert m.sum() == 4 and m[[1, 9, 17, 25]].all()
    assert int(ops.pin_mask(idx, 25).sum()) == 0               # last index overruns the atoms: the reference skips the re-anchoring
    padded = torch.tensor([1, 9, 17, 25, 0, 0], device=DEV)
    m = ops.pin_mask(padded, 30, torch.tensor([4], device=DEV)).cpu().numpy()
    assert m.sum() == 4 and m[0] == 0


def test_pcn_step_graphed_matches_eager_and_reference_loss():
    """PCNTrainStep (gamma * bond loss + kappa * dihedral loss, skip guard at gamma * 300) on a reduced c4 batch: the loss equals
    the oracle's restatement of scripts/pcn_utils.py:160-183, and the CUDA-graph replay over static-capacity buffers
    reproduces the eager steps bit for bit (losses and parameters after 3 steps)."""
    from coarsegrainingvae_b200 import train
    from coarsegrainingvae_b200.factory import build_pcn
    cfg = dict(synthetic.CONFIGS["c4_protein"])
    cfg["n_res"], cfg["n_basis"] = 40, 128
    raws = [synthetic.pcn_batch(cfg, i, _gpu_radius, n_proteins=2) for i in range(3)]
    caps = {"CG_nbr_list": max(b["CG_nbr_list"].shape[0] for b in raws) + 16, "bond_edge_list": raws[0]["bond_edge_list"].shape[0] + 8,
            "dihe_idxs": raws[0]["dihe_idxs"].shape[0] + 5, "ca_idx": raws[0]["ca_idx"].shape[0] + 3}
    statics = [_to(train.to_static_pcn_batch(b, caps), DEV) for b in raws]
    gamma, kappa = 2.0, 0.5

    def make():
        torch.manual_seed(3)
        return build_pcn(cfg["n_basis"], cfg["n_rbf"], cfg["cg_cutoff"], 3).to(DEV)

    # loss of the step == oracle (float64) on the unpadded batch
    m0 = make()
    tr0 = train.PCNTrainStep(m0, gamma, kappa)
    out = m0(_to(raws[0], DEV))
    want = orc.pcn_loss(out[5].detach().double().cpu(), raws[0]["xyz"].double(), raws[0]["bond_edge_list"], raws[0]["dihe_idxs"],
                        gamma, kappa)[0]
    got = tr0._loss(statics[0], None)
    assert rel_err(got, want) <= TOL
    assert tr0.loss_limit == gamma * 300.0
    # eager vs graph
    ma, mb = make(), make()
    ta = train.PCNTrainStep(ma, gamma, kappa, capturable=True)
    ta.prepare(statics[0], None)
    tb = train.PCNTrainStep(mb, gamma, kappa, capturable=True)
    tb.prepare(statics[0], None)
    graphed = train.GraphedTrainStep(tb, statics[0], None)
    mb.load_state_dict(ma.state_dict())
    for name in ("exp_avg", "exp_avg_sq", "step_count"):
        getattr(tb, name).zero_()
    # the warm-up steps of the capture advanced tb; reset it to ta's state (parameters are views of the flat buffers)
    tb.flat_p.copy_(ta.flat_p)
    la, lb = [], []
    for b in statics:
        la.append(float(ta.step(b, None)))
        lb.append(float(graphed.step(b)))
    assert la == lb, (la, lb)
    assert torch.equal(ta.flat_p, tb.flat_p)


# ------------------------------------------------------------------------------------------ sample-quality metrics (f3)

def test_sample_quality_metrics_match_reference():
    """csrc/metrics.cu == the reference's eval_sample_qualities (tests/golden/sample_quality.npz): bond adjacency bit-exact,
    difference counts and valid sets exact, ratios exact, RMSDs to 1e-6; plus a 3000-atom case against the oracle."""
    from coarsegrainingvae_b200 import metrics
    from oracle import metrics_oracle as morc
    g = load("sample_quality.npz")
    z = torch.from_numpy(np.asarray(g["z"])).to(DEV)
    ref = torch.from_numpy(np.asarray(g["ref_xyz"])).to(DEV)
    samples = torch.from_numpy(np.asarray(g["samples"])).to(DEV)
    assert np.array_equal(metrics.get_bond_graphs(ref, z).cpu().numpy(), g["ref_bonds"])
    assert np.array_equal(metrics.get_bond_graphs(samples[3], z).cpu().numpy(), g["sample3_bonds"])
    counts, _ = metrics.sample_quality_counts(ref, z, samples)
    assert np.array_equal(counts[:, 0].cpu().numpy(), g["diff_counts"])
    all_r, heavy_r, vr, var, gvr, gavr = metrics.eval_sample_qualities(ref, z, samples)
    assert vr == float(g["valid_ratio"]) and var == float(g["valid_allatom_ratio"])
    assert np.allclose(gvr, g["graph_val_ratio"], rtol=0, atol=1e-12) and np.allclose(gavr, g["graph_allatom_val_ratio"], rtol=0, atol=1e-12)
    assert np.allclose(all_r, g["all_rmsds"], rtol=1e-6) and np.allclose(heavy_r, g["heavy_rmsds"], rtol=1e-6)
    # larger: 3000 atoms on a jittered lattice (many pairs at the cutoff boundary), 5 samples, vs the CPU oracle
    rng = np.random.default_rng(3)
    n = 3000
    xyz = synthetic.lattice_points(n, 1.45, rng).astype(np.float32)
    zz = rng.choice([1, 6, 7, 8], size=n, p=[0.4, 0.4, 0.1, 0.1])
    radii = np.asarray(metrics.COV_CUTOFF, dtype=np.float32)[zz]
    smp = np.stack([xyz + rng.normal(size=xyz.shape).astype(np.float32) * s for s in (0.0, 0.005, 0.02, 0.05, 0.2)]).astype(np.float32)
    want = morc.eval_sample_qualities(xyz, list(smp), zz, radii)
    got = metrics.eval_sample_qualities(torch.from_numpy(xyz).to(DEV), torch.from_numpy(zz).to(DEV), torch.from_numpy(smp).to(DEV))
    assert np.array_equal(metrics.get_bond_graphs(torch.from_numpy(smp[2]).to(DEV), torch.from_numpy(zz).to(DEV)).cpu().numpy(),
                          morc.get_bond_graphs(smp[2], radii).numpy())
    assert got[2] == want[2] and got[3] == want[3]
    assert np.allclose(got[4], want[4], rtol=0, atol=1e-12) and np.allclose(got[5], want[5], rtol=0, atol=1e-12)
    for a, b in ((got[0], want[0]), (got[1], want[1])):
        assert (a is None) == (b is None) and (a is None or np.allclose(a, b, rtol=1e-6))


# ------------------------------------------------------------------------------------------ on-GPU batch assembly (f2)

def _trajectory(T, n, n_cgs, seed):
    rng = np.random.default_rng(seed)
    base = synthetic.lattice_points(n, 1.6, rng).astype(np.float32)
    xyz = np.stack([base + rng.normal(size=base.shape).astype(np.float32) * 0.25 for _ in range(T)]).astype(np.float32)
    z = rng.choice([1, 6, 7, 8], size=n).astype(np.float32)
    mapping = (np.arange(n) * n_cgs // n).astype(np.int64)
    bonds = np.asarray(gorc.radius_graph(base, 1.9), dtype=np.int64)
    return xyz, z, mapping, bonds


@pytest.mark.parametrize("cg_cutoff", [6.5, None])
def test_device_collate_matches_host_pipeline(cg_cutoff):
    """DeviceCGDataset.collate (trajectory resident in HBM, graphs built on the fly, bead means by the pooling kernel) ==
    the reference pipeline: per-frame props (datasets.py:470-500) -> generate_neighbor_list (data.py:207-252) -> CG_collate
    (data.py:255-289); index tensors bit for bit, also in the static-capacity form (== train.to_static_batch) with no host read."""
    from coarsegrainingvae_b200 import train
    T, n, n_cgs, atom_cutoff = 7, 60, 5, 4.0
    xyz, z, mapping, bonds = _trajectory(T, n, n_cgs, 11)
    props = {k: [] for k in ("nxyz", "CG_nxyz", "num_atoms", "num_CGs", "CG_mapping", "bond_edge_list")}
    for t in range(T):
        cg = np.stack([xyz[t][mapping == b].astype(np.float32).mean(0) for b in range(n_cgs)]).astype(np.float32)
        props["nxyz"].append(torch.from_numpy(np.concatenate([z[:, None], xyz[t]], 1)))
        props["CG_nxyz"].append(torch.from_numpy(np.concatenate([np.arange(n_cgs, dtype=np.float32)[:, None], cg], 1)))
        props["num_atoms"].append(torch.LongTensor([n]))
        props["num_CGs"].append(torch.LongTensor([n_cgs]))
        props["CG_mapping"].append(torch.from_numpy(mapping))
        props["bond_edge_list"].append(torch.from_numpy(bonds))
    ds = cg_mod.CGDataset(props)
    ds.generate_neighbor_list(atom_cutoff=atom_cutoff, cg_cutoff=cg_cutoff, device=DEV, undirected=True)
    dds = cg_mod.DeviceCGDataset(z, xyz, mapping, bonds, atom_cutoff, cg_cutoff, device=DEV)
    for idx in ([3, 0, 6], [5], [1, 1, 2, 4]):
        want = cg_mod.CG_collate([ds[i] for i in idx])
        got = dds.collate(idx)
        assert set(got.keys()) == set(want.keys())
        for k, w in want.items():
            g = got[k].cpu()
            assert g.shape == w.shape and g.dtype == w.dtype, (k, g.shape, w.shape, g.dtype, w.dtype)
            if w.dtype.is_floating_point:
                assert rel_err(g, w) <= 1e-6, k                       # CG_nxyz: bead means (summation order)
            else:
                assert torch.equal(g, w), k
        caps = dds.capacities(len(idx))
        swant = train.to_static_batch(want, caps)
        sgot = dds.collate(idx, static=True)
        for k, w in swant.items():
            if torch.is_tensor(w):
                g = sgot[k].cpu()
                assert g.shape == w.shape, (k, g.shape, w.shape)
                assert (rel_err(g, w) <= 1e-6) if w.dtype.is_floating_point else torch.equal(g, w), k
            else:
                assert sgot[k] == w, k


def test_train_step_on_device_collated_batches():
    """a graphed training step fed by DeviceCGDataset.collate(static=True) (no host data path at all) gives the losses of the
    same step fed by host-collated static batches."""
    from coarsegrainingvae_b200 import train
    from coarsegrainingvae_b200.factory import build_cgvae
    T, n, n_cgs, atom_cutoff, cg_cutoff = 6, 44, 4, 4.0, 7.0
    xyz, z, mapping, bonds = _trajectory(T, n, n_cgs, 5)
    dds = cg_mod.DeviceCGDataset(z, xyz, mapping, bonds, atom_cutoff, cg_cutoff, device=DEV)
    B = 2
    torch.manual_seed(2)
    model = build_cgvae(64, 6, 2, 2, atom_cutoff, cg_cutoff, n_cgs).to(DEV)
    eps = torch.randn(B * n_cgs, 64, generator=torch.Generator().manual_seed(3)).to(DEV)
    first = dds.collate([0, 1], static=True)
    tr = train.TrainStep(model, 0.05, 10.0, capturable=True)
    tr.prepare(first, eps)
    graphed = train.GraphedTrainStep(tr, first, eps)
    ref_model = build_cgvae(64, 6, 2, 2, atom_cutoff, cg_cutoff, n_cgs).to(DEV)
    for idx in ([2, 3], [5, 4], [0, 5]):
        sb = dds.collate(idx, static=True)
        ref_model.load_state_dict(model.state_dict())
        want = train.TrainStep(ref_model, 0.05, 10.0)._loss(dds.collate(idx), eps)
        got = graphed.step(sb)
        assert rel_err(got, want) <= TOL
