"""GPU (B200): parity of the sm_100a kernels, called through the C ABI, against the CPU oracle and the
frozen reference outputs.  Integer work (edge sets, CSR, ranks) is bit-exact; fp32 work is within
1e-5 scale-relative of the fp32 oracle (north-star gate), and within 2e-5..5e-5 of the frozen
reference outputs (two independent fp32 roundings)."""
import numpy as np
import pytest
import torch

import coarsegrainingvae_b200 as cg
from coarsegrainingvae_b200 import ops, synthetic
from oracle import cgvae_oracle as orc
from oracle import graph_oracle as gorc
from tests import parity_cases as pc
from tests.golden_util import load, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5
GEMM_TOL = 1e-5   # fp32 FMA accumulation over K up to 20000 vs a float64 reference


def _dev(x, dtype=None):
    t = torch.as_tensor(x)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


# ------------------------------------------------------------------------------------------ integer kernels

def test_native_library_is_loaded():
    from coarsegrainingvae_b200 import _lib
    assert _lib.load().cgvae_abi_version() == 1
    before = ops.launch_count()
    ops.radius_graph(torch.rand(64, 3, device=DEV), 0.5)
    assert ops.launch_count() > before


def test_radius_graph_golden_bit_exact():
    z = load("graphs.npz")
    n_checked = 0
    for key in z.files:
        parts = key.split("/")
        if len(parts) == 2 and parts[1][:4] in ("und_", "dir_"):
            cutoff = float(parts[1][4:])
            xyz = _dev(z[parts[0] + "/xyz"])
            for use_cells in (False, True):
                got = ops.radius_graph(xyz, cutoff, undirected=parts[1].startswith("und"), use_cells=use_cells)
                assert got.dtype == torch.int64
                assert np.array_equal(got.cpu().numpy(), z[key]), (key, use_cells)
            n_checked += 1
    assert n_checked == 18


@pytest.mark.parametrize("n,cutoff", [(3000, 4.5), (3000, 8.5), (1500, 12.0)])
def test_radius_graph_vs_oracle_large(n, cutoff):
    xyz = synthetic.lattice_points(n, 2.2, np.random.default_rng(n))
    want_u = gorc.radius_graph(xyz, cutoff, True)
    want_d = gorc.radius_graph(xyz, cutoff, False)
    for use_cells in (False, True):
        assert np.array_equal(ops.radius_graph(_dev(xyz), cutoff, True, use_cells=use_cells).cpu().numpy(), want_u)
        assert np.array_equal(ops.radius_graph(_dev(xyz), cutoff, False, use_cells=use_cells).cpu().numpy(), want_d)


def test_radius_graph_batched_ragged_and_empty():
    rng = np.random.default_rng(3)
    sizes = [22, 1, 0, 175, 2, 1300]
    frames = [rng.normal(size=(s, 3)).astype(np.float32) * 4 for s in sizes]
    xyz = np.concatenate(frames, 0)
    ptr = np.cumsum([0] + sizes)
    want = np.concatenate([gorc.radius_graph(f, 3.0, True) + o for f, o in zip(frames, ptr[:-1])], 0)
    for use_cells in (False, True):
        got = ops.radius_graph(_dev(xyz), 3.0, True, frame_ptr=torch.as_tensor(ptr), use_cells=use_cells)
        assert np.array_equal(got.cpu().numpy(), want), use_cells
    got = cg.get_neighbor_list_batch(_dev(xyz), sizes, 3.0)
    assert np.array_equal(got.cpu().numpy(), want)
    # no atoms / no pairs within the cutoff
    assert ops.radius_graph(torch.zeros((0, 3), device=DEV), 1.0).shape == (0, 2)
    far = _dev(np.arange(30, dtype=np.float32).reshape(10, 3) * 100)
    assert ops.radius_graph(far, 1.0).shape == (0, 2)
    # a row longer than the shared-memory sort capacity (2048) exercises the ordered fallback
    blob = _dev(rng.normal(size=(2600, 3)).astype(np.float32) * 0.1)
    got = ops.radius_graph(blob, 5.0, False, use_cells=True).cpu().numpy()
    assert got.shape[0] == 2600 * 2599 and np.array_equal(got, gorc.radius_graph(blob.cpu().numpy(), 5.0, False))


def test_csr_and_segments_exact():
    rng = np.random.default_rng(11)
    n = 257
    half = gorc.radius_graph(rng.normal(size=(n, 3)).astype(np.float32) * 3, 2.5, True)
    pairs = gorc.make_directed(half)
    perm = rng.permutation(pairs.shape[0])
    pairs = pairs[perm]                                   # arbitrary edge order
    g = ops.build_graph(_dev(pairs), n)
    rowptr, col, eid = gorc.receiver_csr(pairs, n)
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(g.col.cpu().numpy(), col) and np.array_equal(g.eid.cpu().numpy(), eid)
    rowptr_t, col_t, eid_t = gorc.receiver_csr(pairs[:, ::-1], n)
    slot_of_edge = np.empty(pairs.shape[0], dtype=np.int64)
    slot_of_edge[eid] = np.arange(pairs.shape[0])
    assert np.array_equal(g.rowptr_t.cpu().numpy(), rowptr_t)
    assert np.array_equal(g.col_t.cpu().numpy(), col_t)
    assert np.array_equal(g.perm_t.cpu().numpy(), slot_of_edge[eid_t])
    assert ops.edge_orientation(_dev(half)) == (False, True)
    assert ops.edge_orientation(_dev(pairs)) == (True, True)
    nd, flag = cg.make_directed(_dev(half))
    assert not flag and np.array_equal(nd.cpu().numpy(), gorc.make_directed(half))
    z = load("graphs.npz")
    i = 0
    while "chan/%d/mapping" % i in z.files:
        m = z["chan/%d/mapping" % i]
        seg = ops.build_segments(_dev(m), int(m.max()) + 1)
        assert np.array_equal(seg.rank.cpu().numpy(), z["chan/%d/index" % i])
        order = np.argsort(m, kind="stable")
        assert np.array_equal(seg.atoms.cpu().numpy(), order)
        i += 1
    big = rng.integers(0, 50, size=20000)
    seg = ops.build_segments(_dev(big), 50)
    assert np.array_equal(seg.rank.cpu().numpy(), gorc.channel_index(big))


# ------------------------------------------------------------------------------------------ fp32 kernels

def test_edge_geometry_vs_oracle():
    rng = np.random.default_rng(5)
    xyz = synthetic.lattice_points(400, 2.2, rng)
    pairs = gorc.make_directed(gorc.radius_graph(xyz, 9.0, True))
    g = ops.build_graph(_dev(pairs), 400)
    for R, cutoff in ((8, 8.5), (10, 12.0), (4, 4.0)):    # cutoff < graph radius: exercises the d >= cutoff branch
        geom = ops.edge_geometry(g, _dev(xyz), _dev(xyz), R, cutoff)
        i, j = pairs[g.eid.cpu().numpy(), 0], pairs[g.eid.cpu().numpy(), 1]
        r = torch.from_numpy(xyz[j] - xyz[i])                 # fp32 difference, as the kernel forms it
        d, unit, rbf, env = orc.edge_geometry(r.double(), R, cutoff)   # float64 evaluation of the reference formulas
        assert float((geom.unit[:, :3] ** 2).sum(1).max()) <= 1.0 + 1e-6
        assert rel_err(geom.unit[:, :3], unit) < 1e-6 and rel_err(geom.unit[:, 3], d) < 1e-6  # geometry tolerances
        assert rel_err(geom.basis[:, :R], rbf * env[:, None]) < GEMM_TOL
        assert rel_err(geom.basis[:, R], env) < 1e-6
        assert float(geom.basis[:, R + 1:].abs().max()) == 0.0 if geom.rb > R + 1 else True


@pytest.mark.parametrize("M,N,K", [(12, 600, 600), (12, 5400, 600), (36, 600, 600), (350, 1800, 600), (97, 130, 77),
                                    (12, 600, 5400), (1000, 64, 20000), (5, 7, 3),
                                    (4000, 2048, 512), (2500, 600, 1200),    # large: tcgen05 3xTF32 path (gemm_tc.cu)
                                    # weight-streaming kernels (gemm_stream.cu): every row-count variant, ragged tails,
                                    # several K blocks, column counts that are not a multiple of the vector width
                                    (16, 1800, 1200), (13, 1030, 260), (40, 1030, 644), (48, 200, 1028), (1, 4096, 128),
                                    (33, 77, 64), (12, 1200, 600), (36, 1800, 600)])
def test_gemm_forms_and_epilogues(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    add = torch.randn(M, N, generator=g)
    ref = A.double() @ W.double().t() + b.double()
    y, zpre = ops.gemm(ops.GEMM_NT, _dev(A), _dev(W), M, N, K, bias=_dev(b), act=1, z_out=True)
    assert rel_err(zpre, ref) < GEMM_TOL
    assert rel_err(y, ref * torch.sigmoid(ref)) < GEMM_TOL
    # activation epilogues: the error budget is set by the pre-activation scale (|z| up to ~5 here), while tanh / relu
    # outputs are O(1): allow GEMM_TOL * max|z| / max|act(z)|
    for act, f in ((2, torch.relu), (3, torch.tanh)):
        tol = GEMM_TOL * max(1.0, float(ref.abs().max() / f(ref).abs().max()))
        assert rel_err(ops.gemm(ops.GEMM_NT, _dev(A), _dev(W), M, N, K, bias=_dev(b), act=act), f(ref)) < tol
    gy = torch.randn(M, N, generator=g)
    zin = torch.randn(M, K, generator=g)
    sig = torch.sigmoid(zin.double())
    want = (gy.double() @ W.double()) * (sig * (1 + zin.double() * (1 - sig)))
    got = ops.gemm(ops.GEMM_NN, _dev(gy), _dev(W), M, K, N, z_in=_dev(zin), dact=1)
    assert rel_err(got, want) < GEMM_TOL
    got = ops.gemm(ops.GEMM_NN, _dev(gy), _dev(W), M, K, N, add=_dev(zin))
    assert rel_err(got, gy.double() @ W.double() + zin.double()) < GEMM_TOL
    got = ops.gemm(ops.GEMM_TN, _dev(gy), _dev(A), N, K, M)
    assert rel_err(got, gy.double().t() @ A.double()) < GEMM_TOL
    assert rel_err(ops.colsum(_dev(gy)), gy.double().sum(0)) < GEMM_TOL
    del add


@pytest.mark.parametrize("tag,cls", [("k3", "EquiMessageBlock"), ("k4", "EquiMessageCross")])
def test_message_block_api(tag, cls):
    pc.message_block_api(DEV, tag, cls, torch.float32, TOL)


def test_message_block_edge_weight():
    pc.message_block_edge_weight(DEV)


def test_pseudo_block_api():
    pc.pseudo_block_api(DEV, torch.float32, TOL)


def test_update_and_contraction_api():
    pc.update_block_api(DEV, torch.float32, TOL)
    pc.contraction_block_api(DEV, torch.float32, TOL)


@pytest.mark.parametrize("tag", ["vae_sym", "vae_nosym", "vae_noneq"])
def test_cgvae_model_golden(tag):
    pc.cgvae_model(DEV, tag, torch.float32, TOL)


@pytest.mark.parametrize("tag", ["pcn_cross", "pcn_plain"])
def test_pcn_model_golden(tag):
    pc.pcn_model(DEV, tag, torch.float32, TOL)


# ------------------------------------------------------------------------------------------ config-sized cases

def _gpu_radius(xyz, cutoff):
    return ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=DEV), cutoff).cpu().numpy()


def _to(batch, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}


@pytest.mark.parametrize("name,n_conf", [("c1_dipeptide", 4), ("c2_chignolin", 2)])
def test_cgvae_step_at_config_width(name, n_conf):
    """full train step (fwd + loss + bwd) at the real feature width (F=600) vs the oracle; c2 at its full size."""
    cfg = dict(synthetic.CONFIGS[name])
    cfg["batch"] = n_conf
    batch = synthetic.cgvae_batch(cfg, 0, _gpu_radius, cg.CG_collate)
    for key, cut in (("nbr_list", cfg["atom_cutoff"]), ("CG_nbr_list", cfg["cg_cutoff"])):   # edge sets bit-exact
        col = "nxyz" if key == "nbr_list" else "CG_nxyz"
        per = cfg["n_atoms"] if key == "nbr_list" else cfg["n_cgs"]
        want = np.concatenate([gorc.radius_graph(batch[col][k * per:(k + 1) * per, 1:].numpy(), cut) + k * per
                               for k in range(n_conf)], 0)
        assert np.array_equal(batch[key].numpy(), want), key
    torch.manual_seed(123)
    model = pc.build_vae(cfg["n_basis"], cfg["n_rbf"], cfg["enc_nconv"], cfg["dec_nconv"], cfg["atom_cutoff"],
                         cfg["cg_cutoff"], cfg["n_cgs"] == 3, cfg["n_cgs"]).to(DEV)
    eps = torch.randn(n_conf * cfg["n_cgs"], cfg["n_basis"], generator=torch.Generator().manual_seed(7))
    out = model(_to(batch, DEV), eps=eps.to(DEV))
    loss = orc.training_loss(out, _to(batch, DEV), cfg["beta"], cfg["gamma"])[0]
    loss.backward()
    spec = dict(n_basis=cfg["n_basis"], n_rbf=cfg["n_rbf"], enc_nconv=cfg["enc_nconv"], dec_nconv=cfg["dec_nconv"],
                atom_cutoff=cfg["atom_cutoff"], cg_cutoff=cfg["cg_cutoff"], decoder="pseudo", breaksym=cfg["n_cgs"] == 3,
                activation="swish")
    P = pc._oracle_params(model)
    oout = orc.cgvae_forward(P, spec, batch, eps=eps)
    oloss = orc.training_loss(oout, batch, cfg["beta"], cfg["gamma"])[0]
    oloss.backward()
    for a, b, k in zip(out, oout, ("mu", "sigma", "pmu", "pstd", "xyz", "xyz_recon")):
        assert rel_err(a, b) < TOL, (k, rel_err(a, b))
    assert rel_err(loss, oloss) < TOL
    # gradients: the same oracle in float64 is the ground truth (see parity_cases._check_grads)
    P64 = {k: v.detach().double().requires_grad_(v.dtype.is_floating_point) for k, v in P.items()}
    b64 = {k: (v.double() if torch.is_tensor(v) and v.dtype.is_floating_point else v) for k, v in batch.items()}
    o64 = orc.cgvae_forward(P64, spec, b64, eps=eps.double())
    orc.training_loss(o64, b64, cfg["beta"], cfg["gamma"])[0].backward()
    worst = pc._check_grads(model, P, TOL, P64=P64)
    print("%s: worst fp32-vs-fp32 gradient difference %.2e" % (name, worst))


def test_pcn_step_reduced_protein_batch():
    """c4 model (PCN, cross decoder, F=512, 9 layers) on 2 proteins x 60 residues vs the oracle."""
    cfg = dict(synthetic.CONFIGS["c4_protein"])
    cfg["n_res"] = 60
    batch = synthetic.pcn_batch(cfg, 0, _gpu_radius, n_proteins=2)
    torch.manual_seed(123)
    net = cg.EquivariantDecoder(n_atom_basis=cfg["n_basis"], n_rbf=cfg["n_rbf"], cutoff=cfg["cg_cutoff"],
                                num_conv=cfg["dec_nconv"], activation="swish", cross_flag=True)
    model = cg.PCN(net, feature_dim=cfg["n_basis"], offset=False).to(DEV)
    out = model(_to(batch, DEV))
    loss = (out[5] - out[4]).pow(2).mean()
    loss.backward()
    spec = dict(n_basis=cfg["n_basis"], n_rbf=cfg["n_rbf"], dec_nconv=cfg["dec_nconv"], atom_cutoff=cfg["cg_cutoff"],
                decoder="cross", activation="swish")
    P = pc._oracle_params(model)
    oout = orc.pcn_forward(P, spec, batch)
    oloss = (oout[5] - oout[4]).pow(2).mean()
    oloss.backward()
    assert rel_err(out[5], oout[5]) < TOL and rel_err(loss, oloss) < TOL
    # gradients against the float64 evaluation of the same oracle (ground truth): nine layers deep, the v_mat gradients go
    # through d sqrt(sum(Vv^2) + 1e-10) / dVv and two fp32 implementations differ from EACH OTHER by up to 1.4e-5 there
    # while each is within 1e-5 (or 10x the fp32 oracle's own error) of the truth -- see parity_cases._check_grads
    P64 = {k: v.detach().double().requires_grad_(v.dtype.is_floating_point) for k, v in P.items()}
    b64 = {k: (v.double() if torch.is_tensor(v) and v.dtype.is_floating_point else v) for k, v in batch.items()}
    o64 = orc.pcn_forward(P64, spec, b64)
    (o64[5] - o64[4]).pow(2).mean().backward()
    pc._check_grads(model, P, TOL, P64=P64)


def _layer_inputs(n, cutoff, F, R, seed):
    xyz = synthetic.lattice_points(n, 2.2, np.random.default_rng(seed), rotate=False)
    half = ops.radius_graph(_dev(xyz), cutoff, True)
    nbrs = torch.cat([half, half.flip(1)], 0)
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(n, F, generator=g)
    v = torch.randn(n, F, 3, generator=g)
    return xyz, nbrs, s, v


def test_message_layer_medium_graph_vs_oracle_and_deterministic():
    """K=3 layer, N=3000, ~0.2 M edges, F=96: forward + all gradients vs the oracle; two runs are bit-identical."""
    n, F, R, cutoff = 3000, 96, 8, 6.0
    xyz, nbrs, s0, v0 = _layer_inputs(n, cutoff, F, R, 17)
    torch.manual_seed(1)
    blk = cg.EquiMessageBlock(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0).to(DEV)
    r = _dev(xyz)[nbrs[:, 1]] - _dev(xyz)[nbrs[:, 0]]
    runs = []
    for _ in range(2):
        blk.zero_grad()
        s, v = s0.to(DEV).requires_grad_(), v0.to(DEV).requires_grad_()
        ds, dv = blk(s, v, r, nbrs)
        (ds.sum() + (dv * dv).sum()).backward()
        runs.append([ds.detach().clone(), dv.detach().clone(), s.grad.clone(), v.grad.clone()] +
                    [p.grad.clone() for p in blk.parameters() if p.grad is not None])
    for a, b in zip(*runs):
        assert torch.equal(a, b)                           # deterministic: no atomics on the float path
    P = pc._oracle_params(blk, "blk.")
    so, vo = s0.clone().requires_grad_(), v0.clone().requires_grad_()
    ods, odv = orc.equi_message(P, "blk", so, vo, r.cpu(), nbrs.cpu(), R, cutoff)
    (ods.sum() + (odv * odv).sum()).backward()
    assert rel_err(runs[0][0], ods) < TOL and rel_err(runs[0][1], odv) < TOL
    assert rel_err(runs[0][2], so.grad) < TOL and rel_err(runs[0][3], vo.grad) < TOL
    pc._check_grads(blk, P, TOL, "blk.")


def test_message_layer_full_size_equivariance():
    """c5 scale (N=20000, F=600, cutoff 4.5 A => ~0.6 M directed edges): too big for the oracle in seconds, so check
    the size-independent property instead -- rotating the input geometry and vectors rotates dv and leaves ds unchanged."""
    n, F, R, cutoff = 20000, 600, 8, 4.5
    xyz, nbrs, s0, v0 = _layer_inputs(n, cutoff, F, R, 23)
    assert nbrs.shape[0] > 500000
    torch.manual_seed(2)
    blk = cg.EquiMessageBlock(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0).to(DEV)
    Q = torch.from_numpy(synthetic.random_rotation(np.random.default_rng(9))).float()
    x = _dev(xyz)
    with torch.no_grad():
        r = x[nbrs[:, 1]] - x[nbrs[:, 0]]
        ds, dv = blk(s0.to(DEV), v0.to(DEV), r, nbrs)
        ds_r, dv_r = blk(s0.to(DEV), (v0 @ Q.t()).to(DEV), r @ Q.t().to(DEV), nbrs)
    assert rel_err(ds_r, ds) < 1e-5
    assert rel_err(dv_r, dv @ Q.t().to(DEV)) < 1e-5


def test_train_step_gradient_sink_matches_plain_autograd():
    """train.TrainStep makes the weight-gradient kernels write straight into one flat buffer (ops.GRAD_SINK); the
    gradients must alias the buffer and equal the plain autograd path.  (Equal to rounding, not bitwise: with the parameters
    laid out in one buffer u_mat and v_mat are adjacent and the update-block backward contracts [gUv | gVv] with [U; V] in ONE
    launch over 2F instead of two launches over F -- a different summation order for everything upstream.)"""
    from coarsegrainingvae_b200.factory import build_cgvae
    from coarsegrainingvae_b200.train import TrainStep, training_loss
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(batch=3, n_basis=64, enc_nconv=2, dec_nconv=2)
    batch = _to(synthetic.cgvae_batch(cfg, 0, _gpu_radius, cg.CG_collate), DEV)
    torch.manual_seed(5)
    model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], cfg["enc_nconv"], cfg["dec_nconv"], cfg["atom_cutoff"],
                        cfg["cg_cutoff"], cfg["n_cgs"]).to(DEV)
    eps = torch.randn(9, 64, device=DEV)
    out = model(batch, eps=eps)
    training_loss(out, out[4], batch["bond_edge_list"], cfg["beta"], cfg["gamma"])[0].backward()
    plain = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    step = TrainStep(model, cfg["beta"], cfg["gamma"])
    step.defer_grads = False
    used = step.prepare(batch, eps)
    assert sorted(used) == sorted(plain)
    step.forward_backward(batch, eps)
    assert step.flat.check_adopted()
    for k, p in model.named_parameters():
        if k in plain:
            assert rel_err(p.grad, plain[k]) < 2e-6, (k, rel_err(p.grad, plain[k]))
    # deferred mode: the small-graph weight / bias gradients come from ONE grouped launch after the backward pass
    # (different summation order than the per-layer contraction: compare to rounding)
    step.defer_grads = True
    before = ops.launch_count()
    step.forward_backward(batch, eps)
    deferred_launches = ops.launch_count() - before
    assert step.flat.check_adopted()
    for k, p in model.named_parameters():
        if k in plain:
            assert rel_err(p.grad, plain[k]) < 2e-6, k
    step.defer_grads = False
    before = ops.launch_count()
    step.forward_backward(batch, eps)
    assert deferred_launches < ops.launch_count() - before
    step.defer_grads = True
    loss = step.step(batch, eps)
    assert torch.isfinite(loss)
    step.flat.release()
    assert not ops.GRAD_SINK


def test_wgrad_grouped_vs_float64():
    """cgvae_wgrad_grouped: every dW = gy^T x and db = colsum(gy) of a table of ragged problems (odd sizes, strided
    operands, bias-only and weight-only entries, more than one launch) against float64."""
    g = torch.Generator().manual_seed(11)
    shapes = [(12, 600, 600), (12, 5400, 600), (36, 600, 600), (1, 7, 5), (60, 5400, 10), (97, 130, 77), (33, 64, 64),
              (128, 200, 36), (12, 1800, 1200)] + [(5 + i % 9, 17 + 3 * i, 9 + 5 * i) for i in range(70)]
    problems, want = [], []
    for idx, (rows, n_out, n_in) in enumerate(shapes):
        gy_full = torch.randn(rows, n_out + 3, generator=g).to(DEV)
        x_full = torch.randn(rows, n_in + 2, generator=g).to(DEV)
        gy = gy_full[:, :n_out] if idx % 3 == 0 else gy_full[:, :n_out].contiguous()
        x = x_full[:, 1:n_in + 1] if idx % 4 == 0 else x_full[:, :n_in].contiguous()
        dW = torch.full((n_out, n_in), float("nan"), device=DEV) if idx % 5 != 1 else None
        db = torch.full((n_out,), float("nan"), device=DEV) if idx % 5 != 2 else None
        problems.append((gy, x if dW is not None else None, dW, db))
        want.append((gy.double().t() @ x.double(), gy.double().sum(0)))
    before = ops.launch_count()
    ops.wgrad_grouped(problems)
    assert ops.launch_count() - before == (len(problems) + ops.WGRAD_PROBLEMS_PER_LAUNCH - 1) // ops.WGRAD_PROBLEMS_PER_LAUNCH
    for (gy, x, dW, db), (w_ref, b_ref) in zip(problems, want):
        if dW is not None:
            assert rel_err(dW, w_ref) < GEMM_TOL, (tuple(gy.shape), tuple(dW.shape))
        if db is not None:
            assert rel_err(db, b_ref) < GEMM_TOL, tuple(gy.shape)
    first = [t.clone() for pr in problems for t in pr[2:] if t is not None]
    ops.wgrad_grouped(problems)
    again = [t for pr in problems for t in pr[2:] if t is not None]
    assert all(torch.equal(a, b) for a, b in zip(first, again))


def test_fused_clip_adam_matches_torch():
    """cgvae_adam_clip_step == torch.nn.utils.clip_grad_norm_ + torch.optim.Adam over several steps (same gradients)."""
    g = torch.Generator().manual_seed(3)
    shapes = [(600, 600), (600,), (1800, 10), (5, 7), (3,)]
    params_a = [torch.randn(*s, generator=g).to(DEV).requires_grad_() for s in shapes]
    n = sum(p.numel() for p in params_a)
    flat_p = torch.cat([p.detach().reshape(-1) for p in params_a]).clone()
    m, v = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
    step = torch.zeros(1, device=DEV)
    norm_out = torch.zeros(1, device=DEV)
    opt = torch.optim.Adam(params_a, lr=1e-3)
    for it in range(4):
        grads = [torch.randn(*s, generator=g).to(DEV) * (10.0 if it % 2 else 1e-3) for s in shapes]   # clipped / not clipped
        for p, gr in zip(params_a, grads):
            p.grad = gr.clone()
        flat_g = torch.cat([gr.reshape(-1) for gr in grads])
        want_norm = torch.nn.utils.clip_grad_norm_(params_a, 0.01 if it % 2 else 1e3)
        opt.step()
        ops.adam_clip_step(flat_p, flat_g, m, v, step, 0.01 if it % 2 else 1e3, 1e-3, norm_out=norm_out)
        assert rel_err(norm_out, want_norm.reshape(1)) < 1e-6
        want = torch.cat([p.detach().reshape(-1) for p in params_a])
        assert rel_err(flat_p, want) < 1e-6, it
    assert float(step) == 4.0 and n == flat_p.numel()
    # grad_scale: the data-parallel 1/world factor applied inside the kernel == pre-scaled gradients
    pa, pb = flat_p.clone(), flat_p.clone()
    ma, va, mb, vb = m.clone(), v.clone(), m.clone(), v.clone()
    sa, sb = step.clone(), step.clone()
    gsum = torch.randn(n, generator=g).to(DEV)
    ops.adam_clip_step(pa, gsum * 0.25, ma, va, sa, 0.01, 1e-3)
    ops.adam_clip_step(pb, gsum, mb, vb, sb, 0.01, 1e-3, grad_scale=0.25)
    assert rel_err(pb, pa) < 1e-6 and rel_err(mb, ma) < 1e-6 and rel_err(vb, va) < 1e-6


def test_graphed_train_step_matches_eager():
    """one CUDA graph per step over static-capacity buffers (device-side edge counts, on-the-fly symmetrisation) gives
    the same losses and the same parameters as eager launches on the plain batches, for batches with different edge
    counts replayed through the SAME graph."""
    import copy
    from coarsegrainingvae_b200.factory import build_cgvae
    from coarsegrainingvae_b200.train import GraphedTrainStep, TrainStep, to_static_batch
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(batch=3, n_basis=64, enc_nconv=2, dec_nconv=2, atom_cutoff=4.0)   # short cutoff: edge count varies
    raw = [synthetic.cgvae_batch(cfg, i, _gpu_radius, cg.CG_collate) for i in range(3)]
    assert len({b["nbr_list"].shape[0] for b in raw}) > 1
    caps = {"nbr_list": 3 * 22 * 21 // 2, "CG_nbr_list": 9, "bond_edge_list": max(b["bond_edge_list"].shape[0] for b in raw) + 8}
    static = [_to(to_static_batch(b, caps), DEV) for b in raw]
    torch.manual_seed(5)
    model_a = build_cgvae(cfg["n_basis"], cfg["n_rbf"], 2, 2, cfg["atom_cutoff"], cfg["cg_cutoff"], 3).to(DEV)
    model_b = copy.deepcopy(model_a)
    eps = torch.randn(9, 64, device=DEV)
    eager = TrainStep(model_a, cfg["beta"], cfg["gamma"], lr=1e-3)
    eager.prepare(_to(raw[0], DEV), eps)
    graph_tr = TrainStep(model_b, cfg["beta"], cfg["gamma"], lr=1e-3, capturable=True)
    graph_tr.prepare(static[0], eps)
    # the three warm-up steps inside GraphedTrainStep update model_b (the capture itself executes nothing):
    # take the same three steps on model_a
    graphed = GraphedTrainStep(graph_tr, static[0], eps)
    for _ in range(3):
        eager.step(_to(raw[0], DEV), eps)
    for n, i in enumerate((1, 2, 0, 1)):
        la = eager.step(_to(raw[i], DEV), eps)
        # dict of tensors (one copy per key) and the packed single-buffer form (one copy) must load the same batch
        lb = graphed.step(static[i] if n % 2 == 0 else graphed.pack({k: (v.cpu() if torch.is_tensor(v) else v)
                                                                     for k, v in static[i].items()}, pin=True))
        assert rel_err(lb, la) < 1e-5, (i, float(la), float(lb))
    for (k, pa), (_, pb) in zip(model_a.named_parameters(), model_b.named_parameters()):
        assert rel_err(pb, pa) < 1e-4, k
    graph_tr.flat.release()
    eager.flat.release()


def test_torch_library_ops_match_autograd_nodes():
    """torch.ops.cgvae_b200.* (torch.library custom ops with fake kernels and registered autograd) give the results of
    the autograd nodes / raw kernel calls the models use, and pass opcheck's schema / fake-tensor / autograd checks."""
    from coarsegrainingvae_b200 import functions, torch_ops  # noqa: F401  (registers the ops)
    T = torch.ops.cgvae_b200
    g = torch.Generator().manual_seed(21)
    rn = lambda *sh: torch.randn(*sh, generator=g).to(DEV)
    # dense vs float64 torch
    x, W, b = rn(12, 600).requires_grad_(), (rn(1800, 600) * 0.05).requires_grad_(), rn(1800).requires_grad_()
    gy = rn(12, 1800)
    y = T.dense(x, W, b)
    y.backward(gy)
    xd, Wd, bd = (t.detach().double().requires_grad_() for t in (x, W, b))
    yd = torch.nn.functional.linear(xd, Wd, bd)
    yd.backward(gy.double())
    for got, want in ((y, yd), (x.grad, xd.grad), (W.grad, Wd.grad), (b.grad, bd.grad)):
        assert rel_err(got, want) < GEMM_TOL
    # mlp2 == functions.MLP2 (same kernels, same order: bitwise)
    ps = [rn(36, 600), rn(600, 600) * 0.05, rn(600), rn(1800, 600) * 0.05, rn(1800)]
    a = [t.clone().requires_grad_() for t in ps]
    c = [t.clone().requires_grad_() for t in ps]
    gy = rn(36, 1800)
    T.mlp2(a[0], a[1], a[2], a[3], a[4], 1)[0].backward(gy)
    functions.MLP2.apply(1, *c).backward(gy)
    for u, w in zip(a, c):
        assert torch.equal(u.grad, w.grad)
    # message layer (3 and 4 splits) == raw kernel calls
    n, F, R = 300, 64, 8
    xyz = rn(n, 3) * 4.0
    pairs = ops.radius_graph(xyz, 6.0)
    gr = ops.build_graph(pairs, n, symmetrize=True)
    geom = ops.edge_geometry(gr, xyz, xyz, R, 6.0)
    for K in (3, 4):
        phi, v = rn(n, K, F).requires_grad_(), rn(n, 3, F).requires_grad_()
        Wf, bf = (rn(K * F, R) * 0.1).requires_grad_(), (rn(K * F) * 0.1).requires_grad_()
        rs, rv = rn(n, F).requires_grad_(), rn(n, 3, F).requires_grad_()
        gs, gv = rn(n, F), rn(n, 3, F)
        o_s, o_v, q = T.message_layer(K, phi, v, gr.rowptr, gr.col, gr.rowptr_t, gr.col_t, gr.perm_t, geom.basis, geom.unit,
                                      Wf, bf, rs, rv, R)
        torch.autograd.backward([o_s, o_v], [gs, gv])
        w_s, w_v, w_q = ops.message_fwd(K, phi.detach(), v.detach(), v.detach() if K == 4 else None, geom, Wf.detach(),
                                        bf.detach(), rs.detach(), rv.detach(), want_q=(K == 4))
        assert torch.equal(o_s, w_s) and torch.equal(o_v, w_v)
        g_phi, g_vs, dWf, dbf = ops.message_bwd(K, phi.detach(), v.detach(), v.detach() if K == 4 else None, w_q, geom,
                                                Wf.detach(), bf.detach(), gs, gv, False, sink=False)
        assert torch.equal(phi.grad, g_phi) and torch.equal(v.grad, g_vs)
        assert torch.equal(Wf.grad, dWf) and torch.equal(bf.grad, dbf)
        assert torch.equal(rs.grad, gs) and torch.equal(rv.grad, gv)
    # bead pooling vs index_add
    mapping = torch.sort(torch.randint(0, 40, (n,), generator=g))[0].to(DEV)
    seg = ops.build_segments(mapping, 40)
    X = rn(n, 3, F).requires_grad_()
    out = T.segment_reduce(X, seg.rowptr, seg.atoms, seg.mapping, True)
    gout = rn(40, 3, F)
    out.backward(gout)
    Xd = X.detach().double().requires_grad_()
    cnt = torch.bincount(mapping, minlength=40).clamp(min=1).double().view(40, 1, 1)
    ref = torch.zeros(40, 3, F, dtype=torch.float64, device=DEV).index_add_(0, mapping, Xd) / cnt
    ref.backward(gout.double())
    assert rel_err(out, ref) < TOL and rel_err(X.grad, Xd.grad) < TOL
    # radius graph op == ops.radius_graph
    assert torch.equal(T.radius_graph(xyz, 6.0, True), pairs)
    # dispatcher-level checks
    checks = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(T.dense.default, (x.detach().requires_grad_(), W.detach().requires_grad_(), b.detach().requires_grad_()),
                          test_utils=checks)
    torch.library.opcheck(T.segment_reduce.default, (X.detach().requires_grad_(), seg.rowptr, seg.atoms, seg.mapping, True),
                          test_utils=checks)
    torch.library.opcheck(T.mlp2.default, tuple(t.detach().requires_grad_() for t in ps) + (1,), test_utils=checks)


def test_torch_library_ops_round2():
    """the remaining entry points behind the dispatcher: contraction (message_layer on the atoms -> beads graph), 9-split
    layer, update block, lifting, clip + Adam -- each equal to the autograd node / raw kernels the models use (bitwise:
    same kernels in the same order) and passing opcheck."""
    from coarsegrainingvae_b200 import functions, torch_ops  # noqa: F401
    T = torch.ops.cgvae_b200
    g = torch.Generator().manual_seed(33)
    rn = lambda *sh: torch.randn(*sh, generator=g).to(DEV)
    checks = ("test_schema", "test_faketensor", "test_autograd_registration")
    n, nb, F, R = 120, 12, 64, 8
    xyz, cg_xyz = rn(n, 3) * 3.0, rn(nb, 3) * 3.0
    mapping = torch.sort(torch.randint(0, nb, (n,), generator=g))[0].to(DEV)
    seg = ops.build_segments(mapping, nb)
    # contraction: receivers = beads, senders = atoms, residual state on the receiver set (conv.py:703-733)
    cgraph = ops.contraction_graph(seg)
    cgeom = ops.edge_geometry(cgraph, xyz, cg_xyz, R, 6.0)
    phi, v = rn(n, 3, F).requires_grad_(), rn(n, 3, F).requires_grad_()
    Wf, bf = (rn(3 * F, R) * 0.1).requires_grad_(), (rn(3 * F) * 0.1).requires_grad_()
    H, V = rn(nb, F).requires_grad_(), rn(nb, 3, F).requires_grad_()
    o_s, o_v, _ = T.message_layer(3, phi, v, cgraph.rowptr, cgraph.col, cgraph.rowptr_t, cgraph.col_t, cgraph.perm_t, cgeom.basis,
                                  cgeom.unit, Wf, bf, H, V, R)
    w_s, w_v, _ = ops.message_fwd(3, phi.detach(), v.detach(), None, cgeom, Wf.detach(), bf.detach(), H.detach(), V.detach())
    assert o_s.shape == (nb, F) and torch.equal(o_s, w_s) and torch.equal(o_v, w_v)
    gs, gv = rn(nb, F), rn(nb, 3, F)
    torch.autograd.backward([o_s, o_v], [gs, gv])
    g_phi, g_vs, dWf, dbf = ops.message_bwd(3, phi.detach(), v.detach(), None, None, cgeom, Wf.detach(), bf.detach(), gs, gv, False,
                                            sink=False)
    assert torch.equal(phi.grad, g_phi) and torch.equal(v.grad, g_vs) and torch.equal(Wf.grad, dWf) and torch.equal(H.grad, gs)
    # 9-split layer on the bead graph == functions.Message9Block minus its phi-MLP
    pairs = ops.radius_graph(cg_xyz, 50.0)
    gr = ops.build_graph(pairs, nb, symmetrize=True)
    geom = ops.edge_geometry(gr, cg_xyz, cg_xyz, R, 50.0)
    ins = [rn(nb, 9, F), rn(nb, F), rn(nb, F), rn(nb, 3, F), rn(nb, 3, F), rn(9 * F, R) * 0.1, rn(9 * F) * 0.1]
    a = [t.clone().requires_grad_() for t in ins]
    outs = T.message9_layer(a[0], a[1], a[2], a[3], a[4], gr.rowptr, gr.col, gr.rowptr_t, gr.col_t, gr.perm_t, geom.basis, geom.unit,
                            a[5], a[6], True, R)
    want = ops.message9_fwd(*ins[:5], geom, ins[5], ins[6], True)
    gouts = [rn(nb, F), rn(nb, F), rn(nb, 3, F), rn(nb, 3, F)]
    torch.autograd.backward(list(outs), gouts)
    wb = ops.message9_bwd(*ins[:5], geom, ins[5], ins[6], True, *gouts)
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(outs, want))
    for got, ref in zip((a[1].grad, a[2].grad, a[3].grad, a[4].grad, a[0].grad, a[5].grad, a[6].grad), wb):
        assert rel_err(got, ref) <= 1e-6
    # update block == functions.UpdateBlockFn
    ps = [rn(n, F), rn(n, 3, F), rn(F, F) * 0.1, rn(F, F) * 0.1, rn(F, 2 * F) * 0.1, rn(F), rn(3 * F, F) * 0.1, rn(3 * F)]
    a = [t.clone().requires_grad_() for t in ps]
    c = [t.clone().requires_grad_() for t in ps]
    gs, gv = rn(n, F), rn(n, 3, F)
    o = T.update_block(*a, 1, True)
    torch.autograd.backward([o[0], o[1]], [gs, gv])
    w = functions.UpdateBlockFn.apply(True, 1, *c)
    torch.autograd.backward(list(w), [gs, gv])
    torch.cuda.synchronize()
    assert torch.equal(o[0], w[0]) and torch.equal(o[1], w[1])
    for u, x in zip(a, c):
        assert rel_err(u.grad, x.grad) <= 1e-6
    # lifting == functions.Lift (all three modes)
    pin = torch.zeros(n, dtype=torch.uint8, device=DEV)
    pin[::7] = 1
    for mode in (0, 1, 2):
        Vb = rn(nb, 3, F).requires_grad_()
        Vc = Vb.detach().clone().requires_grad_()
        got = T.lift(Vb, cg_xyz, seg.mapping, seg.rank, seg.rowptr, seg.atoms, pin if mode == 2 else None, mode)
        ref = functions.Lift.apply(seg, mode, pin if mode == 2 else None, Vc, cg_xyz)
        gx = rn(n, 3)
        got.backward(gx)
        ref.backward(gx)
        assert torch.equal(got, ref) and torch.equal(Vb.grad, Vc.grad)
    # clip + Adam (mutating op) == ops.adam_clip_step
    m = 10007
    p0, g0 = rn(m), rn(m) * 1e-2
    bufs_a = [p0.clone(), torch.zeros(m, device=DEV), torch.zeros(m, device=DEV), torch.zeros(1, device=DEV)]
    bufs_b = [t.clone() for t in bufs_a]
    for _ in range(3):
        T.adam_clip_step(bufs_a[0], g0, bufs_a[1], bufs_a[2], bufs_a[3], 0.01, 1e-3, 0.9, 0.999, 1e-8, 1.0)
        ops.adam_clip_step(bufs_b[0], g0, bufs_b[1], bufs_b[2], bufs_b[3], 0.01, 1e-3)
    assert all(torch.equal(x, y) for x, y in zip(bufs_a, bufs_b)) and float(bufs_a[3]) == 3.0
    # dispatcher-level checks
    torch.library.opcheck(T.update_block.default, tuple(t.detach().requires_grad_() for t in ps) + (1, True), test_utils=checks)
    torch.library.opcheck(T.lift.default, (rn(nb, 3, F).requires_grad_(), cg_xyz, seg.mapping, seg.rank, seg.rowptr, seg.atoms, None, 1),
                          test_utils=checks)
    torch.library.opcheck(T.message9_layer.default,
                          tuple(t.detach().requires_grad_() for t in ins[:5]) + (gr.rowptr, gr.col, gr.rowptr_t, gr.col_t, gr.perm_t,
                                                                                 geom.basis, geom.unit, ins[5].requires_grad_(),
                                                                                 ins[6].requires_grad_(), True, R), test_utils=checks)
    torch.library.opcheck(T.adam_clip_step.default, (p0.clone(), g0, torch.zeros(m, device=DEV), torch.zeros(m, device=DEV),
                                                     torch.zeros(1, device=DEV), 0.01, 1e-3, 0.9, 0.999, 1e-8, 1.0),
                          test_utils=("test_schema", "test_faketensor"))


def test_graphed_sampler_matches_eager_members():
    """train.GraphedSampler (prior + n_ensemble decoder passes of scripts/sampling.py:265-284 as one CUDA graph over
    static-capacity inputs) reproduces the eager member-by-member geometries for the same noise, for conformations
    replayed through the SAME graph."""
    from coarsegrainingvae_b200.factory import build_cgvae
    from coarsegrainingvae_b200.train import GraphedSampler, sample_ensemble_member, to_static_batch
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    cfg.update(batch=2, n_basis=64, enc_nconv=2, dec_nconv=3)
    raw = [synthetic.cgvae_batch(cfg, i, _gpu_radius, cg.CG_collate) for i in range(4)]
    caps = {"nbr_list": 2 * 22 * 21 // 2, "CG_nbr_list": 6, "bond_edge_list": max(b["bond_edge_list"].shape[0] for b in raw) + 8}
    static = [_to(to_static_batch(b, caps), DEV) for b in raw]
    torch.manual_seed(9)
    model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], 2, 3, cfg["atom_cutoff"], cfg["cg_cutoff"], 3).to(DEV)
    n_ens = 3
    sampler = GraphedSampler(model, static[0], n_ens)
    for i in (1, 2, 3, 0):
        eps = torch.randn(n_ens, 6, 64, generator=torch.Generator().manual_seed(i)).to(DEV)
        got = sampler.sample(static[i], eps).clone()
        b = _to(raw[i], DEV)
        with torch.no_grad():
            z, cg_z, xyz, cg_xyz, nbr, cg_nbr, mapping, num = model.get_inputs(b)
            graphs = cg.BatchGraphs()
            mu, sig = model.prior_net(cg_z, cg_xyz.contiguous(), cg_nbr, graphs=graphs)
            want = torch.stack([sample_ensemble_member(model, cg_xyz.contiguous(), cg_nbr, mapping, num, mu, sig, eps[m], graphs=graphs)
                                for m in range(n_ens)])
        assert got.shape == want.shape and rel_err(got, want) < 1e-6, i


@pytest.mark.parametrize("cls,K", [("EquiMessageBlock", 3), ("EquiMessageCross", 4)])
def test_message_layer_ragged_degrees_vs_oracle(cls, K):
    """Hand-built directed graph with isolated nodes, in- and out-degrees of exactly 1, 31, 32, 33, 64, 65 (the edge
    metadata is staged 32 edges at a time per warp) and one hub of degree 150: outputs and all gradients vs the oracle."""
    n, F, R, cutoff = 200, 64, 8, 9.0
    g = torch.Generator().manual_seed(41 + K)
    xyz = torch.rand(n, 3, generator=g) * 5.0
    degs = {3: 1, 10: 31, 11: 32, 12: 33, 40: 64, 41: 65, 77: 150}            # receiver -> in-degree; the rest: 0
    recv, send = [], []
    for i, d in degs.items():
        others = [j for j in range(n) if j != i]
        perm = torch.randperm(len(others), generator=g)[:d].tolist()
        for q in sorted(perm):
            recv.append(i); send.append(others[q])
    nbrs = torch.tensor([recv, send], dtype=torch.int64).t().contiguous()      # directed as given (not symmetrised)
    nbrs = torch.cat([nbrs, nbrs.flip(1)], 0)                                  # + reversed edges: ragged out-degrees too
    s0, v0 = torch.randn(n, F, generator=g), torch.randn(n, F, 3, generator=g)
    torch.manual_seed(3)
    blk = getattr(cg, cls)(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0).to(DEV)
    r = (xyz[nbrs[:, 1]] - xyz[nbrs[:, 0]]).to(DEV)
    s, v = s0.to(DEV).requires_grad_(), v0.to(DEV).requires_grad_()
    ds, dv = blk(s, v, r, nbrs.to(DEV))
    (ds.sum() + (dv * dv).sum()).backward()
    P = pc._oracle_params(blk, "blk.")
    so, vo = s0.clone().requires_grad_(), v0.clone().requires_grad_()
    fn_o = orc.equi_message if K == 3 else orc.equi_message_cross
    ods, odv = fn_o(P, "blk", so, vo, r.cpu(), nbrs, R, cutoff)
    (ods.sum() + (odv * odv).sum()).backward()
    assert rel_err(ds, ods) < TOL and rel_err(dv, odv) < TOL
    assert rel_err(s.grad, so.grad) < TOL and rel_err(v.grad, vo.grad) < TOL
    isolated = [i for i in range(n) if i not in degs and i not in set(send)]
    assert isolated and float(ds.detach()[isolated].abs().max()) == 0.0 and float(dv.detach()[isolated].abs().max()) == 0.0
    pc._check_grads(blk, P, TOL, "blk.")


@pytest.mark.parametrize("M,N1,N2,K", [(36, 600, 600, 600), (12, 600, 1800, 600), (40, 130, 70, 260), (200, 64, 64, 128)])
def test_dense_pair_forward(M, N1, N2, K):
    """cgvae_dense_pair_fwd: two bias-free Dense layers on one input, weights adjacent in memory, one launch (streaming
    kernel up to 48 rows, two tiled launches beyond); non-adjacent weights take the two-launch route in ops."""
    g = torch.Generator().manual_seed(M + N1)
    x = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N1 + N2, K, generator=g) / K ** 0.5).to(DEV)
    W1, W2 = W[:N1], W[N1:]
    before = ops.launch_count()
    y1, y2 = ops.linear_pair_fwd(x, W1, W2)
    launches = ops.launch_count() - before
    assert launches == (1 if M <= 48 else 2)
    assert rel_err(y1, x.double() @ W1.double().t()) < GEMM_TOL and rel_err(y2, x.double() @ W2.double().t()) < GEMM_TOL
    z1, z2 = ops.linear_pair_fwd(x, W1.clone(), W2.clone())          # not adjacent: same numbers from two launches
    assert rel_err(z1, y1) < 1e-6 and rel_err(z2, y2) < 1e-6


@pytest.mark.parametrize("act", pc.ACTIVATIONS)
def test_blocks_with_every_registry_activation(act):
    """every block type with the non-default activations of the reference's layer_types registry (epilogue codes 2..7 of the
    kernels) vs the oracle: outputs, input and parameter gradients"""
    pc.blocks_with_activation(DEV, act, torch.float32, TOL)
