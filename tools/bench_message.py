"""Per-launch timing of the fused message layer: SIMT kernel (message.cu) vs tensor-core kernel (message_tc.cu) for the
chunk sizes RC, at the chignolin atom-graph shape (c2) and the large-graph sweep (c5).  One JSON line per row.
    python tools/bench_message.py [--only c2,c5] [--bwd]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from coarsegrainingvae_b200 import ops, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="c2,c5")
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--rcs", default="4,8,16")
args = ap.parse_args()
dev = torch.device("cuda", 0)
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
PEAK_TF32 = 0.5 * PEAKS.get("bf16_tflops", 1590.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > L2: written between timed launches


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3      # median, min in microseconds


def case(tag, xyz_np, cutoff, F, R, n_split, frame_ptr=None):
    x = torch.as_tensor(xyz_np, device=dev)
    half = ops.radius_graph(x, cutoff, True, frame_ptr=frame_ptr)
    pairs = torch.cat([half, half.flip(1)], 0)
    n = x.shape[0]
    graph = ops.build_graph(pairs, n)
    geom = ops.edge_geometry(graph, x, x, R, cutoff)
    E = graph.n_edges
    g = torch.Generator().manual_seed(1)
    phi = torch.randn(n, n_split, F, generator=g).to(dev)
    v = torch.randn(n, 3, F, generator=g).to(dev)
    s = torch.randn(n, F, generator=g).to(dev)
    Wf = (torch.randn(n_split * F, R, generator=g) * 0.3).to(dev)
    bf = (torch.randn(n_split * F, generator=g) * 0.3).to(dev)
    gs, gv = torch.randn(n, F, generator=g).to(dev), torch.randn(n, 3, F, generator=g).to(dev)
    flops = 2.0 * (R + 1) * n_split * F * E
    rows = []
    for mode in ["simt"] + ["tc%s" % r for r in args.rcs.split(",")]:
        if mode == "simt":
            ops.MSG_TC = "0"
        else:
            ops.MSG_TC, ops.MSG_TC_RC = "1", int(mode[2:])
            if n_split == 4 and ops.MSG_TC_RC == 16:
                continue
        row = dict(case=tag, mode=mode, n=n, E=E, F=F, R=R, n_split=n_split)
        if mode != "simt":
            t = ops.message_tiles(geom, False)
            nb = int(t.bptr[-1])
            row.update(fill=E / (64.0 * nb), build_us=timeit(lambda: (geom.__dict__.pop("_tiles", None), ops.message_tiles(geom, False)), 5)[0])
        fwd = lambda: ops.message_fwd(n_split, phi, v, v if n_split == 4 else None, geom, Wf, bf, s, v, want_q=n_split == 4)
        med, mn = timeit(fwd)
        row.update(fwd_us=med, fwd_min_us=mn, fwd_Medges_s=E / med, fwd_tflops=flops / med * 1e-6, fwd_frac=flops / med * 1e-6 / PEAK_TF32)
        if args.bwd:
            q = fwd()[2]
            bwd = lambda: ops.message_bwd(n_split, phi, v, v if n_split == 4 else None, q, geom, Wf, bf, gs, gv, True, sink=False)
            med, mn = timeit(bwd)
            row.update(bwd_us=med, bwd_min_us=mn, bwd_Medges_s=E / med, bwd_tflops=2 * flops / med * 1e-6,
                       bwd_frac=2 * flops / med * 1e-6 / PEAK_TF32)
        print(json.dumps(row), flush=True)
        rows.append(row)
    return rows


only = set(args.only.split(","))
if "c2" in only:
    cfg = synthetic.CONFIGS["c2_chignolin"]
    pts = [synthetic.lattice_points(cfg["n_atoms"], cfg["spacing"], np.random.default_rng(1434 + k)) for k in range(cfg["batch"])]
    fp = torch.tensor([0, 175, 350], dtype=torch.int64, device=dev)
    case("c2_atom_graph", np.concatenate(pts, 0), cfg["atom_cutoff"], cfg["n_basis"], cfg["n_rbf"], 3, frame_ptr=fp)
if "c1" in only:
    cfg = synthetic.CONFIGS["c1_dipeptide"]
    pts = [synthetic.lattice_points(cfg["n_atoms"], cfg["spacing"], np.random.default_rng(1334 + k)) for k in range(cfg["batch"])]
    fp = torch.arange(0, 22 * 33, 22, dtype=torch.int64, device=dev)
    case("c1_atom_graph", np.concatenate(pts, 0), cfg["atom_cutoff"], cfg["n_basis"], cfg["n_rbf"], 3, frame_ptr=fp)
if "c4" in only:
    cfg = synthetic.CONFIGS["c4_protein"]
    pts = [synthetic.lattice_points(cfg["n_res"], cfg["spacing"], np.random.default_rng(1634 + k)) for k in range(cfg["batch"])]
    fp = torch.arange(0, 250 * 65, 250, dtype=torch.int64, device=dev)
    case("c4_ca_graph", np.concatenate(pts, 0), cfg["cg_cutoff"], cfg["n_basis"], cfg["n_rbf"], 4, frame_ptr=fp)
if "c5" in only:
    cfg = synthetic.CONFIGS["c5_large"]
    xyz = synthetic.lattice_points(cfg["n_atoms"], cfg["spacing"], np.random.default_rng(55), rotate=False)
    for cutoff in cfg["cutoffs"]:
        case("c5_%g" % cutoff, xyz, cutoff, cfg["n_basis"], cfg["n_rbf"], 3)
