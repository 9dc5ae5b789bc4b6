"""Secondary measurements for the other BASELINE.json configurations (c1, c3, c4, c5).  One JSON line per row.
    python tools/bench_configs.py [--only c5,c4,...]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import coarsegrainingvae_b200 as cg
from coarsegrainingvae_b200 import ops, synthetic
from coarsegrainingvae_b200.factory import build_cgvae, build_pcn
from coarsegrainingvae_b200.train import TrainStep, GraphedTrainStep, to_static_batch, training_loss

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="c1,c3,c4,c5")
ap.add_argument("--c4-batch", type=int, default=64)
args = ap.parse_args()
dev = torch.device("cuda", 0)
PEAK_TF32 = 0.5 * json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 795.0
rad = lambda xyz, c: ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=dev), c).cpu().numpy()


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(**kw):
    print(json.dumps(kw), flush=True)


only = set(args.only.split(","))

if "c1" in only:
    cfg = dict(synthetic.CONFIGS["c1_dipeptide"])
    raw = [synthetic.cgvae_batch(cfg, i, rad, cg.CG_collate) for i in range(2)]
    B, n, ncg = cfg["batch"], cfg["n_atoms"], cfg["n_cgs"]
    caps = {"nbr_list": B * n * (n - 1) // 2, "CG_nbr_list": B * ncg * (ncg - 1) // 2,
            "bond_edge_list": max(b["bond_edge_list"].shape[0] for b in raw) + 64}
    batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in to_static_batch(b, caps).items()} for b in raw]
    torch.manual_seed(123)
    model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], cfg["enc_nconv"], cfg["dec_nconv"], cfg["atom_cutoff"], cfg["cg_cutoff"], ncg).to(dev)
    eps = torch.randn(B * ncg, cfg["n_basis"], device=dev)
    tr = TrainStep(model, cfg["beta"], cfg["gamma"], capturable=True)
    tr.prepare(batches[0], eps)
    g = GraphedTrainStep(tr, batches[0], eps)
    it = [0]
    def step():
        g.step(batches[it[0] % 2]); it[0] += 1
    ms = timeit(step, 5, 30)
    emit(config="c1_dipeptide", what="train step (CUDA graph), batch 32 x 22 atoms, F=600, enc 4 / dec 5", ms_per_step=ms,
         conformations_per_s=B / ms * 1e3)
    tr.flat.release()
    del model, tr, g

if "c3" in only:
    # chignolin sampling: prior once per conformation, then n_ensemble decoder passes (scripts/sampling.py:268-279)
    cfg = dict(synthetic.CONFIGS["c2_chignolin"]); cfg["batch"] = 1
    torch.manual_seed(123)
    model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], cfg["enc_nconv"], cfg["dec_nconv"], cfg["atom_cutoff"], cfg["cg_cutoff"], cfg["n_cgs"]).to(dev)
    confs = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic.cgvae_batch(cfg, i, rad, cg.CG_collate).items()} for i in range(4)]
    n_ens = 8
    def sample_conf(b):
        with torch.no_grad():
            z, cg_z, xyz, cg_xyz, nbr, cg_nbr, mapping, num = model.get_inputs(b)
            graphs = cg.BatchGraphs()
            mu, sig = model.prior_net(cg_z, cg_xyz.contiguous(), cg_nbr, graphs=graphs)
            outs = []
            for m in range(n_ens):
                H = mu + torch.randn_like(sig) * sig
                outs.append(model.decoder(cg_xyz.contiguous(), cg_nbr, H, H, mapping, num, graphs=graphs))
            return torch.stack(outs)
    it = [0]
    def step():
        sample_conf(confs[it[0] % 4]); it[0] += 1
    ms = timeit(step, 3, 20)
    emit(config="c3_sampling", what="1 prior + 8 decoder passes per conformation (eager launches), chignolin model", ms_per_conformation=ms,
         decoder_passes_per_s=n_ens / ms * 1e3, conformation_members_per_s=n_ens / ms * 1e3)
    # the same as one CUDA graph per conformation (train.GraphedSampler), noise passed in
    from coarsegrainingvae_b200.train import GraphedSampler
    raw = [synthetic.cgvae_batch(cfg, i, rad, cg.CG_collate) for i in range(4)]
    n, ncg = cfg["n_atoms"], cfg["n_cgs"]
    caps = {"nbr_list": n * (n - 1) // 2, "CG_nbr_list": ncg * (ncg - 1) // 2,
            "bond_edge_list": max(b["bond_edge_list"].shape[0] for b in raw) + 64}
    sconfs = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in to_static_batch(b, caps).items()} for b in raw]
    sampler = GraphedSampler(model, sconfs[0], n_ens)
    eps = torch.randn(n_ens, ncg, cfg["n_basis"], device=dev)
    # same noise, same conformation: graph replay == eager members
    with torch.no_grad():
        z, cg_z, xyz, cg_xyz, nbr, cg_nbr, mapping, num = model.get_inputs(confs[1])
        graphs = cg.BatchGraphs()
        mu, sig = model.prior_net(cg_z, cg_xyz.contiguous(), cg_nbr, graphs=graphs)
        want = torch.stack([model.decoder(cg_xyz.contiguous(), cg_nbr, mu + eps[m] * sig, mu + eps[m] * sig, mapping, num, graphs=graphs)
                            for m in range(n_ens)])
    got = sampler.sample(sconfs[1], eps).clone()
    err = float((got - want).abs().max() / want.abs().max())
    def gstep():
        sampler.sample(sconfs[it[0] % 4], eps); it[0] += 1
    ms = timeit(gstep, 3, 30)
    emit(config="c3_sampling_graph", what="1 prior + 8 decoder passes per conformation as ONE CUDA graph replay (train.GraphedSampler)",
         ms_per_conformation=ms, decoder_passes_per_s=n_ens / ms * 1e3, max_rel_err_vs_eager=err)
    del model

if "c5" in only:
    cfg = synthetic.CONFIGS["c5_large"]
    F, R = cfg["n_basis"], cfg["n_rbf"]
    xyz_np = synthetic.lattice_points(cfg["n_atoms"], cfg["spacing"], np.random.default_rng(55), rotate=False)
    xyz = torch.as_tensor(xyz_np, device=dev)
    g0 = torch.Generator().manual_seed(5)
    s0 = torch.randn(cfg["n_atoms"], F, generator=g0).to(dev)
    v0 = torch.randn(cfg["n_atoms"], 3, F, generator=g0).to(dev)
    for n_sub in (2000, 5000, 10000, 20000):
        for cutoff in cfg["cutoffs"]:
            sub = xyz[:n_sub].contiguous()
            ms = timeit(lambda: ops.radius_graph(sub, cutoff, True), 2, 5)
            E = int(ops.radius_graph(sub, cutoff, True).shape[0])
            emit(config="c5_radius_graph", n_atoms=n_sub, cutoff=cutoff, undirected_edges=E, ms=ms, note="cell lists incl. count+scan+fill+sort and the host read of the edge count")
    for cutoff in cfg["cutoffs"]:
        half = ops.radius_graph(xyz, cutoff, True)
        graph = ops.build_graph(half, cfg["n_atoms"], symmetrize=True)
        geom = ops.edge_geometry(graph, xyz, xyz, R, cutoff)
        E = graph.n_edges
        torch.manual_seed(1)
        blk = cg.EquiMessageBlock(feat_dim=F, activation="swish", n_rbf=R, cutoff=cutoff, dropout=0.0).to(dev)
        W1, b1, W2, b2, Wf, bf = blk.inv_message.params()
        phi = torch.randn(cfg["n_atoms"], 3, F, device=dev)
        ms_f = timeit(lambda: ops.message_fwd(3, phi, v0, None, geom, Wf, bf, s0, v0), 2, 5)
        gs, gv = torch.randn_like(s0), torch.randn_like(v0)
        ms_b = timeit(lambda: ops.message_bwd(3, phi, v0, None, None, geom, Wf, bf, gs, gv, True, sink=False), 2, 5)
        s_in = s0.clone().requires_grad_(); v_in = v0.clone().requires_grad_()
        def layer():
            blk.zero_grad(set_to_none=True)
            o_s, o_v = blk.fused(s_in, v_in, geom)
            (o_s.sum() + o_v.sum()).backward()
        ms_l = timeit(layer, 2, 5)
        flops = 2.0 * (R + 1) * 3 * F * E
        emit(config="c5_message_layer", cutoff=cutoff, directed_edges=E, fused_fwd_kernel_ms=ms_f, fused_bwd_kernel_ms=ms_b,
             fwd_edges_per_s=E / ms_f * 1e3, bwd_edges_per_s=E / ms_b * 1e3, layer_fwd_bwd_ms=ms_l, layer_edges_per_s=E / ms_l * 1e3,
             fwd_filter_tflops=flops / ms_f / 1e9, fwd_frac_of_tf32_peak=flops / ms_f / 1e9 / PEAK_TF32,
             bwd_filter_tflops=2 * flops / ms_b / 1e9, bwd_frac_of_tf32_peak=2 * flops / ms_b / 1e9 / PEAK_TF32,
             gathered_bytes_fwd=E * 4.0 * 6 * F, fwd_gather_TBps=E * 4.0 * 6 * F / ms_f / 1e9,
             note="layer_fwd_bwd includes the phi-MLP GEMMs (N=20000 rows, SIMT fp32) and their gradients")
        del blk, graph, geom

if "c4" in only:
    cfg = dict(synthetic.CONFIGS["c4_protein"])
    nb = args.c4_batch
    batch = synthetic.pcn_batch(cfg, 0, rad, n_proteins=nb)
    dbatch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    torch.manual_seed(123)
    model = build_pcn(cfg["n_basis"], cfg["n_rbf"], cfg["cg_cutoff"], cfg["dec_nconv"]).to(dev)
    def run():
        out = model(dbatch)
        return (out[5] - out[4]).pow(2).mean()
    used = None
    from coarsegrainingvae_b200.train import FlatGrads, used_parameters
    up = used_parameters(model, lambda: run().backward())
    flat = FlatGrads([p for _, p in up])
    opt = torch.optim.Adam(flat.params, lr=1e-4, fused=True)
    def step():
        flat.zero_()
        run().backward()
        flat.clip_(0.01)
        opt.step()
    ms = timeit(step, 2, 5)
    emit(config="c4_protein", what="PCN train step (eager): %d proteins x 250 residues x 8 atoms, cross decoder 9 layers, F=512" % nb,
         atoms=int(batch["xyz"].shape[0]), beads=int(batch["ca_xyz"].shape[0]), directed_cg_edges=2 * int(batch["CG_nbr_list"].shape[0]),
         ms_per_step=ms, conformations_per_s=nb / ms * 1e3, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
