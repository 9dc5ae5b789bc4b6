"""Kernel timeline of the graphed chignolin train step via torch.profiler (CUPTI): where does the step time go?
    python tools/trace_step.py [--steps 6] [--no-graph]
Writes gpurun_out/trace_summary.txt"""
import argparse, collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import coarsegrainingvae_b200 as cg
from coarsegrainingvae_b200 import ops, synthetic
from coarsegrainingvae_b200.factory import build_cgvae
from coarsegrainingvae_b200.train import GraphedTrainStep, TrainStep, to_static_batch

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--out", default="gpurun_out/trace_summary.txt")
args = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = dict(synthetic.CONFIGS["c2_chignolin"])
rad = lambda xyz, c: ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=dev), c).cpu().numpy()
raw = [synthetic.cgvae_batch(cfg, i, rad, cg.CG_collate) for i in range(2)]
B, n, ncg = cfg["batch"], cfg["n_atoms"], cfg["n_cgs"]
caps = {"nbr_list": B * n * (n - 1) // 2, "CG_nbr_list": B * ncg * (ncg - 1) // 2, "bond_edge_list": max(b["bond_edge_list"].shape[0] for b in raw) + 64}
batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in to_static_batch(b, caps).items()} for b in raw]
torch.manual_seed(123)
model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], cfg["enc_nconv"], cfg["dec_nconv"], cfg["atom_cutoff"], cfg["cg_cutoff"], cfg["n_cgs"]).to(dev)
eps = torch.randn(B * ncg, cfg["n_basis"], device=dev)
tr = TrainStep(model, cfg["beta"], cfg["gamma"], capturable=not args.no_graph)
tr.prepare(batches[0], eps)
if args.no_graph:
    run = lambda b: tr.step(b, eps)
else:
    g = GraphedTrainStep(tr, batches[0], eps)
    run = lambda b: g.step(b)
for i in range(5):
    run(batches[i % 2])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(args.steps):
        run(batches[i % 2])
        torch.cuda.synchronize()
path = "/tmp/trace.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
kern = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
kern.sort(key=lambda e: e["ts"])
# split into steps by large idle gaps (synchronize between steps)
# equal-sized chunks: every replay launches the same activities
per = len(kern) // args.steps
steps = [kern[i * per:(i + 1) * per] for i in range(args.steps)]
lines = []
P = lines.append
P("# torch.profiler (CUPTI) timeline of the c2_chignolin train step, %s, %d steps traced" % ("eager" if args.no_graph else "CUDA graph replay", len(steps)))
for si, s in enumerate(steps[-3:]):
    t0, t1 = s[0]["ts"], max(e["ts"] + e["dur"] for e in s)
    span = t1 - t0
    tot = sum(e["dur"] for e in s)
    # union busy
    busy, end = 0.0, t0
    conc = 0.0
    for e in s:
        a, b = e["ts"], e["ts"] + e["dur"]
        if b > end:
            busy += b - max(a, end); end = b
    P("step %d: %d activities, span %.1f us, sum of durations %.1f us, GPU busy (union) %.1f us, idle inside span %.1f us" % (si, len(s), span, tot, busy, span - busy))
s = steps[-1]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in s:
    nm = e["name"].split("(")[0].replace("void ", "")[:80]
    agg[nm][0] += 1; agg[nm][1] += e["dur"]
tot = sum(v[1] for v in agg.values())
P("last step by kernel:")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    P("  %-82s n=%4d tot=%8.1f us avg=%7.2f us %5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
# streams
by_stream = collections.defaultdict(float)
for e in s:
    by_stream[e.get("args", {}).get("stream", "?")] += e["dur"]
P("sum of durations per stream: %s" % dict(by_stream))
# gaps on the union timeline
gaps, where = [], []
end = s[0]["ts"] + s[0]["dur"]
prev = s[0]
for e in s[1:]:
    if e["ts"] > end:
        gaps.append(e["ts"] - end)
        where.append((e["ts"] - end, prev["name"].split("(")[0].replace("void ", "")[:60], e["name"].split("(")[0].replace("void ", "")[:60]))
    if e["ts"] + e["dur"] >= end:
        prev = e
    end = max(end, e["ts"] + e["dur"])
gaps = np.array(gaps) if gaps else np.zeros(1)
P("largest idle gaps (us, after kernel -> before kernel):")
for g_, a_, b_ in sorted(where, key=lambda w: -w[0])[:24]:
    P("  %6.2f  %-60s -> %s" % (g_, a_, b_))
P("idle gaps in last step: n=%d total=%.1f us median=%.2f us p90=%.2f us max=%.1f us" % (len(gaps), gaps.sum(), np.median(gaps), np.percentile(gaps, 90), gaps.max()))
os.makedirs(os.path.dirname(args.out), exist_ok=True)
open(args.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
