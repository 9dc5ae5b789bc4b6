#!/usr/bin/env python
"""bench.py -- headline benchmark of the CGVAE equivariant message-passing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--workload c2_chignolin|c1_dipeptide]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): chignolin all-atom CGVAE,
175 atoms, n_cgs 6, batch 2, enc_nconv 2 / dec_nconv 9, n_basis 600, n_rbf 10, cutoffs 12 / 25 A, synthetic
jittered-lattice conformations, random-init weights.  One "step" = one optimisation step of the reference's
training loop (scripts/utils.py:112-157): make_directed + CSR build + edge geometry + forward + loss + backward
(+ gradient all-reduce over NCCL when N > 1) + clip_grad_norm_(0.01) + Adam.

Prints ONE JSON line (rank 0).  `value` = conformations/s with the batch resident in HBM; `e2e` = the same through
the public API from pinned HOST buffers with the H2D copy and the D2H loss read inside the timed region;
`roofline` = the dominant kernel (fused message layer on the atom graph) timed live with CUDA events;
`cpu_baseline` = the CPU oracle port of the reference on this box's host cores, bounded sample.

--impl reference: times the reference algorithm's CPU implementation (the oracle port: the reference is pure
Python/PyTorch, so there is no oracle/_ref binary) on the host cores for the same metric / config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) at the chignolin shapes, from the committed
# ncu --set full capture profiles/r1b_ncu_full_hot_kernels_raw.csv (tools/profile_kernels.py)
NCU_DRAM_BYTES = {"message_fwd": 12.3e6, "message_bwd": 12.5e6}
NCU_SOURCE = "profiles/r1b_ncu_full_hot_kernels_raw.csv (tools/profile_kernels.py: same shapes, same build)"

METRIC = "conformations/s (fwd+bwd train step)"
UNIT = "conformations/s"


def _args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2_chignolin", choices=["c2_chignolin", "c1_dipeptide"])
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic batches cycled through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    return ap.parse_args()


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def _spec(cfg):
    return dict(n_basis=cfg["n_basis"], n_rbf=cfg["n_rbf"], enc_nconv=cfg["enc_nconv"], dec_nconv=cfg["dec_nconv"],
                atom_cutoff=cfg["atom_cutoff"], cg_cutoff=cfg["cg_cutoff"], decoder="pseudo", breaksym=cfg["n_cgs"] == 3,
                activation="swish")


def _config_json(name, cfg, n_gpus, extra=None):
    out = {"workload": "%s: CGVAE train step, %d atoms x batch %d per GPU, n_cgs %d, enc_nconv %d / dec_nconv %d, n_basis %d, "
                       "n_rbf %d, cutoffs %g/%g A" % (name, cfg["n_atoms"], cfg["batch"], cfg["n_cgs"], cfg["enc_nconv"],
                                                      cfg["dec_nconv"], cfg["n_basis"], cfg["n_rbf"], cfg["atom_cutoff"],
                                                      cfg["cg_cutoff"]),
           "batch_per_gpu": cfg["batch"], "global_batch": cfg["batch"] * n_gpus,
           "parallelism": "dp%d (conformations sharded by rank, one NCCL all-reduce of the used-gradient bucket)" % n_gpus,
           "step": "make_directed + CSR + geometry + fwd + loss + bwd + clip_grad_norm_ + Adam",
           "l2": "no explicit flush: parameters + gradients + Adam state streamed every step are >1 GB, far above the 126 MB L2"}
    if extra:
        out.update(extra)
    return out


class ClockSampler(object):
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU reference arm

def _oracle_step_fn(cfg, pool_size, threads):
    """the reference algorithm on CPU (oracle port): returns (step(i) -> loss, description)."""
    import torch
    from coarsegrainingvae_b200 import synthetic
    from oracle import cgvae_oracle as orc
    from oracle import graph_oracle as gorc
    from coarsegrainingvae_b200.factory import build_cgvae

    torch.set_num_threads(threads)
    torch.manual_seed(123)
    model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], cfg["enc_nconv"], cfg["dec_nconv"], cfg["atom_cutoff"], cfg["cg_cutoff"],
                        cfg["n_cgs"])                                      # parameter container only (CPU tensors)
    P = dict(model.named_parameters())
    spec = _spec(cfg)
    batches = [gorc.collate([{k: v.numpy() for k, v in synthetic.cgvae_sample(cfg, 1234 + 100 + i * cfg["batch"] + k,
                                                                               gorc.radius_graph).items()}
                             for k in range(cfg["batch"])]) for i in range(pool_size)]
    batches = [{k: torch.from_numpy(v) for k, v in b.items()} for b in batches]
    gen = torch.Generator().manual_seed(7)
    eps = torch.randn(cfg["batch"] * cfg["n_cgs"], cfg["n_basis"], generator=gen)
    params = None
    state = {}

    def step(i):
        b = batches[i % pool_size]
        for p in P.values():
            p.grad = None
        out = orc.cgvae_forward(P, spec, b, eps=eps)
        loss = orc.training_loss(out, b, cfg["beta"], cfg["gamma"])[0]
        loss.backward()
        if "opt" not in state:
            used = [p for p in P.values() if p.grad is not None]
            state["used"] = used
            state["opt"] = torch.optim.Adam(used, lr=1e-4)
        torch.nn.utils.clip_grad_norm_(state["used"], 0.01)
        state["opt"].step()
        return float(loss)

    del params
    return step


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _reference_step_fn(cfg, pool_size, threads):
    """The UNMODIFIED reference (pip-installed into the git-ignored baseline/_ref by tools/install_reference.sh; it travels
    to the GPU box with the snapshot) driven through its own public API on the host CPU: model construction as in
    scripts/run_ala.py:184-209, batches from the reference's get_neighbor_list (data.py:65-82) + CG_collate (data.py:255-289),
    and one iteration of the training loop of scripts/utils.py:112-157 per step (scripts/ itself needs ase / mdtraj, which
    are absent, so the dozen lines of that loop are restated here around the reference's model call).  Returns None when
    baseline/_ref is missing (then the oracle port is timed and reported as kind "port")."""
    if not os.path.isdir(os.path.join(REF_DIR, "CoarseGrainingVAE")):
        return None
    os.environ["CGVAE_REFERENCE_ROOT"] = REF_DIR
    import torch
    from torch import nn
    from coarsegrainingvae_b200 import synthetic
    from oracle import ref_shim
    ref_shim.REFERENCE_ROOT = REF_DIR
    try:
        _, _, ref_cgvae, ref_data = ref_shim.import_reference()
    except Exception:
        return None
    torch.set_num_threads(threads)
    torch.manual_seed(123)
    F, R, act = cfg["n_basis"], cfg["n_rbf"], "swish"
    decoder = ref_cgvae.EquivariantPsuedoDecoder(n_atom_basis=F, n_rbf=R, cutoff=cfg["atom_cutoff"], num_conv=cfg["dec_nconv"],
                                                 activation=act, breaksym=cfg["n_cgs"] == 3)
    encoder = ref_cgvae.EquiEncoder(n_conv=cfg["enc_nconv"], n_atom_basis=F, n_rbf=R, cutoff=cfg["cg_cutoff"], activation=act,
                                    cg_mp=False, dir_mp=False)
    prior = ref_cgvae.CGprior(n_conv=cfg["enc_nconv"], n_atom_basis=F, n_rbf=R, cutoff=cfg["cg_cutoff"], activation=act, dir_mp=False)
    atom_mu = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
    atom_sigma = nn.Sequential(nn.Linear(F, F), nn.ReLU(), nn.Linear(F, F))
    model = ref_cgvae.CGequiVAE(encoder, decoder, atom_mu, atom_sigma, cfg["n_cgs"], feature_dim=F, prior_net=prior, det=False,
                                equivariant=True)
    model.train()
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-4)            # scripts/run_ala.py:211

    def ref_radius(xyz, cutoff):
        return ref_data.get_neighbor_list(torch.as_tensor(xyz), "cpu", cutoff, True).numpy()

    batches = [synthetic.cgvae_batch(cfg, i, ref_radius, ref_data.CG_collate) for i in range(pool_size)]
    beta, gamma, EPS = cfg["beta"], cfg["gamma"], 1e-6

    def step(i):
        batch = batches[i % pool_size]
        S_mu, S_sigma, H_prior_mu, H_prior_sigma, xyz, xyz_recon = model(batch)                     # utils.py:114
        loss_kl = 0.5 * ((S_sigma.pow(2) / H_prior_sigma.pow(2)).sum(-1) + ((S_mu - H_prior_mu).pow(2) / H_prior_sigma).sum(-1)
                         + torch.log(H_prior_sigma.pow(2)).sum(-1) - torch.log(S_sigma.pow(2)).sum(-1) - S_sigma.shape[-1]).mean()
        loss_recon = (xyz_recon - xyz).pow(2).mean()                                                # utils.py:124
        edge_list = batch["bond_edge_list"]
        gen = ((xyz_recon[edge_list[:, 0]] - xyz_recon[edge_list[:, 1]]).pow(2).sum(-1) + EPS).sqrt()
        dat = ((xyz[edge_list[:, 0]] - xyz[edge_list[:, 1]]).pow(2).sum(-1) + EPS).sqrt()
        loss = loss_recon + loss_kl * beta + (gen - dat).pow(2).mean() * gamma                       # utils.py:141
        if loss.item() >= gamma * 200.0 or torch.isnan(loss):                                       # utils.py:145-148
            return float(loss)
        optimizer.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.01)
        optimizer.step()
        return float(loss)

    return step


def _cpu_step_fn(cfg, pool_size, threads):
    """(step, kind): the unmodified reference when baseline/_ref is present, else the oracle port."""
    step = _reference_step_fn(cfg, pool_size, threads)
    if step is not None:
        return step, "reference"
    return _oracle_step_fn(cfg, pool_size, threads), "port"


def run_reference(args, cfg):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step, kind = _cpu_step_fn(cfg, args.pool, threads)
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    ms = 1e3 * dt / max(args.steps, 1)
    value = cfg["batch"] / (ms / 1e3)
    sample = "%d full steps of the %s batch (%d conformations each) after %d warm-up" % (args.steps, args.workload,
                                                                                          cfg["batch"], args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": _config_json(args.workload, cfg, args.gpus),
            "reference_impl": ("unmodified wwang2/CoarseGrainingVAE from baseline/_ref on the host CPU cores, torch %s"
                               if kind == "reference" else "oracle port of the reference (baseline/_ref missing), torch %s")
                              % torch.__version__,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- CUDA arm

def run_cuda(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist
    import coarsegrainingvae_b200 as cg
    from coarsegrainingvae_b200 import ops, synthetic
    from coarsegrainingvae_b200.factory import build_cgvae
    from coarsegrainingvae_b200.train import GraphedTrainStep, TrainStep, to_static_batch

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl cuda needs a CUDA device: there is no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    def gpu_radius(xyz, cutoff):
        return ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=dev), cutoff).cpu().numpy()

    # synthetic data: `pool` distinct batches per rank (different conformations on every rank: weak scaling)
    raw_batches = [synthetic.cgvae_batch(cfg, rank * 1000 + i, gpu_radius, cg.CG_collate) for i in range(args.pool)]
    edges = int(np.mean([2 * b["nbr_list"].shape[0] for b in raw_batches]))
    use_graph = not args.no_graph
    if use_graph:
        # static capacities: dense upper bounds for the radius graphs, pool maximum (+ slack) for the bond list
        B, n, ncg = cfg["batch"], cfg["n_atoms"], cfg["n_cgs"]
        caps = {"nbr_list": B * n * (n - 1) // 2, "CG_nbr_list": max(B * ncg * (ncg - 1) // 2, 1),
                "bond_edge_list": max(b["bond_edge_list"].shape[0] for b in raw_batches) + 64}
        raw_batches = [to_static_batch(b, caps) for b in raw_batches]
    host_batches = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b.items()} for b in raw_batches]
    dev_batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()} for b in host_batches]
    h2d_bytes = int(np.mean([sum(v.numel() * v.element_size() for v in b.values() if torch.is_tensor(v)) for b in host_batches]))

    torch.manual_seed(123)                                   # identical replicas on every rank
    model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], cfg["enc_nconv"], cfg["dec_nconv"], cfg["atom_cutoff"], cfg["cg_cutoff"],
                        cfg["n_cgs"]).to(dev)
    eps = torch.randn(cfg["batch"] * cfg["n_cgs"], cfg["n_basis"], generator=torch.Generator().manual_seed(7)).to(dev)
    trainer = TrainStep(model, cfg["beta"], cfg["gamma"], lr=1e-4, max_norm=0.01, capturable=use_graph)
    used = trainer.prepare(dev_batches[0], eps)
    n_used = int(trainer.flat.flat.numel())
    if use_graph:
        graphed = GraphedTrainStep(trainer, dev_batches[0], eps)
        # batches packed once (dataset-preparation time) into the byte layout of the graph's static input buffer:
        # loading a batch is ONE copy -- D2D for `value`, H2D from pinned host memory for `e2e`
        packed_dev = [graphed.pack(b, device=dev) for b in dev_batches]
        packed_host = [graphed.pack(b, pin=True) for b in host_batches]
        h2d_bytes = int(packed_host[0].numel())
        run_step = lambda i: graphed.step(packed_dev[i])
    else:
        run_step = lambda i: trainer.step(dev_batches[i], eps)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: batch resident in HBM
    for i in range(max(args.warmup, 3)):
        run_step(i % args.pool)
    sync_all()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        run_step(i % args.pool)
    e1.record()
    sync_all()
    launches = (graphed.launches_per_step * args.steps) if use_graph else (ops.launch_count() - launches0)
    clock_info = clocks.stop()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms = float(ms_total) / max(args.steps, 1)
    value = cfg["batch"] * n_gpus / (ms / 1e3)

    # ---- e2e: pinned host batch -> H2D -> step -> D2H loss, every step
    def e2e_step(i):
        hb = host_batches[i % args.pool]
        if use_graph:
            loss = graphed.step(packed_host[i % args.pool])  # pinned host -> static device buffer (one async H2D), then replay
        else:
            db = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in hb.items()}
            loss = trainer.step(db, eps)
        return loss.item()                                   # the reference reads loss.item() every step (utils.py:145)

    for i in range(3):
        e2e_step(i)
    sync_all()
    e0.record()
    last = 0.0
    for i in range(args.steps):
        last = e2e_step(i)
    e1.record()
    sync_all()
    ms_e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e) / max(args.steps, 1)
    e2e_value = cfg["batch"] * n_gpus / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel (fused message layer on the ATOM graph): CUDA events around the individual
    # launches.  Kernel boundaries are not host-visible inside a graph replay, so these launches are timed in eager
    # steps of the same workload, same process, right after the timed region.
    ops.TIMER = ops.KernelTimer(["message_fwd", "message_bwd", "adam_clip"])
    n_prof = min(args.steps, 10)
    for i in range(n_prof):
        trainer.step(dev_batches[i % args.pool], eps)
    sync_all()
    timer, ops.TIMER = ops.TIMER, None
    peaks = _peaks()
    tf32_peak = 0.5 * peaks["bf16_sustained"]
    roof, other = None, {}
    summ = timer.summary()
    R, F = cfg["n_rbf"], cfg["n_basis"]
    for name, factor in (("message_fwd", 1.0), ("message_bwd", 2.0)):
        recs = [(t, m) for t, m in summ.get(name, []) if m["E"] >= edges // 2 and m["n_recv"] == m["n_send"]]
        if not recs:
            continue
        t_ms = float(np.mean([t for t, _ in recs]))
        E = float(edges)           # live directed edges (the static-capacity graph holds a few padded slots on top)
        flops = factor * 2.0 * (R + 1) * 3 * F * E                   # filter contraction incl. the bias column
        share = (sum(t for t, _ in recs) / n_prof) / ms
        entry = {"kernel": name + "_kernel<3,%d> (atom graph)" % (ops.rb_for(R) // 4), "bound": "tensor",
                 "achieved": flops / (t_ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                 "frac": flops / (t_ms * 1e-3) / 1e12 / tf32_peak, "traffic": NCU_DRAM_BYTES.get(name),
                 "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this launch in " + NCU_SOURCE,
                 "fp32_simt": {"achieved_tflops": 2.0 * (46.0 if factor == 1.0 else 61.0) * E * F / (t_ms * 1e-3) / 1e12,
                               "peak_tflops": 2.0 * 148 * 128 * 1.965e-3,
                               "note": "all fp32 FMAs of the kernel (filter + channel mixing, 46 / 61 per edge and channel "
                                       "forward / backward) against the SIMT FMA peak: the bound this implementation runs on"},
                 "peak_source": "derived: 0.5 x bf16_tflops_sustained of %s MEASURED_PEAKS (TF32 is not measured there)" % peaks["source"],
                 "avg_launch_us": 1e3 * t_ms, "launches_per_step": len(recs) / n_prof, "edges_per_launch": E,
                 "edges_per_s": E / (t_ms * 1e-3), "share_of_step": share,
                 "note": "fp32 SIMT implementation this round: the tensor pipe is idle, frac is measured against the "
                         "tensor roofline the north star names"}
        other[name] = entry
    if other:
        roof = max(other.values(), key=lambda e: e["share_of_step"])
    # HBM-bound streaming kernels of the step: algorithmic bytes / measured time against the measured copy bandwidth
    hbm_peak = peaks["hbm_gbs"]

    def hbm_entry(kernel, recs, bytes_of, note):
        if not recs:
            return None
        tot_ms = sum(t for t, _ in recs)
        tot_b = float(sum(bytes_of(m) for _, m in recs))
        return {"kernel": kernel, "bound": "hbm", "achieved": tot_b / (tot_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": tot_b / (tot_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "launches_per_step": len(recs) / n_prof,
                "avg_launch_us": 1e3 * tot_ms / len(recs), "algorithmic_bytes_per_step": tot_b / n_prof,
                "share_of_step": (tot_ms / n_prof) / ms, "peak_source": "hbm_gbs of %s MEASURED_PEAKS" % peaks["source"], "note": note}

    # Kernels of a few microseconds cannot be timed launch by launch in eager mode (the host is slower than the GPU, the
    # event pair would measure the wait for the next launch): the EXACT launches of one step -- same operands, same
    # weights, 270 MB of them so nothing stays in L2 between replays -- are re-issued back to back inside one CUDA graph
    # and the graph is timed with CUDA events.
    ops.TIMER = ops.KernelTimer(["gemm", "wgrad_grouped"], keep_operands=True)
    trainer.step(dev_batches[0], eps)
    sync_all()
    keep, ops.TIMER = ops.TIMER, None
    ksumm = keep.summary()

    def graph_time_ms(fn, reps=5):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            fn()
        g_.replay()
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(reps):
            g_.replay()
        b_.record()
        torch.cuda.synchronize()
        return a_.elapsed_time(b_) / reps

    calls = [m for _, m in ksumm.get("gemm", []) if m["M"] <= 16 and m["form"] != ops.GEMM_TN and m["N"] * m["K"] >= 4096]
    tables = [m["table"] for _, m in ksumm.get("wgrad_grouped", []) if m.get("table")]

    def replay_gemms():
        for m in calls:
            ops.gemm(m["form"], m["A"], m["B"], m["M"], m["N"], m["K"], bias=m["bias"], act=m["act"], z_in=m["z_in"],
                     dact=m["dact"], add=m["add"])

    def replay_wgrad():
        for tb in tables:
            ops.wgrad_grouped(tb)

    iso = []
    if calls:
        iso.append(("gemm_stream", "gemm_nt_stream / gemm_nn_stream (12-bead Dense layers, TMA weight streaming)",
                    graph_time_ms(replay_gemms), len(calls), float(sum(4.0 * m["N"] * m["K"] for m in calls)),
                    "bytes = the weight matrix once per launch; the %d launches of one step replayed back to back in a CUDA graph" % len(calls)))
    if tables:
        n_launch = sum((len(tb) + ops.WGRAD_PROBLEMS_PER_LAUNCH - 1) // ops.WGRAD_PROBLEMS_PER_LAUNCH for tb in tables)
        iso.append(("wgrad_grouped", "wgrad_grouped_kernel (all small-graph weight / bias gradients of the step)",
                    graph_time_ms(replay_wgrad), n_launch,
                    float(sum(4.0 * ((p[2].numel() if p[2] is not None else 0) + (p[3].numel() if p[3] is not None else 0))
                              for tb in tables for p in tb)),
                    "bytes = gradients written once; the launches of one step replayed in a CUDA graph"))
    for key, kernel, t_ms, n_l, tot_b, note in iso:
        other[key] = {"kernel": kernel, "bound": "hbm", "achieved": tot_b / (t_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": tot_b / (t_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "launches_per_step": n_l,
                      "avg_launch_us": 1e3 * t_ms / n_l, "algorithmic_bytes_per_step": tot_b, "share_of_step": t_ms / ms,
                      "peak_source": "hbm_gbs of %s MEASURED_PEAKS" % peaks["source"], "note": note}
    del keep, ksumm, calls, tables
    for key, ent in (
            ("adam_clip", hbm_entry("sumsq_partial + adam_clip_kernel (clip_grad_norm_ + Adam on flat buffers)",
                                    summ.get("adam_clip", []), lambda m: 32.0 * m["n"],
                                    "bytes = 4 (norm pass) + 28 (p, g, m, v read; p, m, v written) per parameter")),):
        if ent is not None:
            other[key] = ent

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": _config_json(args.workload, cfg, n_gpus),
            "workload_details": {
                "directed_atom_edges_per_batch": edges, "used_gradient_floats": n_used, "used_parameters": len(used),
                "launch_mode": ("one CUDA graph per step over static-capacity input buffers (edge counts read on the device)"
                                if use_graph else "eager launches")},
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "last_loss": last},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_kernels": other,
            "message_pass_edges_per_s": (other.get("message_fwd") or {}).get("edges_per_s")}

    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step, kind = _cpu_step_fn(cfg, 2, threads)
        step(0)
        t0 = time.perf_counter()
        n = 0
        while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 8):
            step(n + 1)
            n += 1
        dt = (time.perf_counter() - t0) / n
        line["cpu_baseline"] = {"value": cfg["batch"] / dt, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": "%d full train steps of the same %s batch on the host CPU (%s, torch %s), after 1 "
                                          "warm-up" % (n, args.workload, "the unmodified reference from baseline/_ref"
                                                       if kind == "reference" else "oracle port of the reference", torch.__version__)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = _args()
    from coarsegrainingvae_b200 import synthetic
    cfg = dict(synthetic.CONFIGS[args.workload])
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_cuda(args, cfg)


if __name__ == "__main__":
    main()
