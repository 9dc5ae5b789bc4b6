"""Kernel timeline (torch.profiler / CUPTI) of the eager c4 PCN training step: per-kernel totals of one step.
    python tools/trace_c4.py [--proteins 64] [--out gpurun_out/trace_c4.txt]"""
import argparse, collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coarsegrainingvae_b200 import ops, synthetic
from coarsegrainingvae_b200.factory import build_pcn
from coarsegrainingvae_b200.train import TrainStep, training_loss
ap = argparse.ArgumentParser()
ap.add_argument("--proteins", type=int, default=64)
ap.add_argument("--out", default="gpurun_out/trace_c4.txt")
args = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = dict(synthetic.CONFIGS["c4_protein"])
rad = lambda xyz, c: ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=dev), c).cpu().numpy()
raw = synthetic.pcn_batch(cfg, 0, rad, n_proteins=args.proteins)
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in raw.items()}
torch.manual_seed(123)
model = build_pcn(cfg["n_basis"], cfg["n_rbf"], cfg["cg_cutoff"], cfg["dec_nconv"]).to(dev)
class Step(TrainStep):
    def _loss(self, b, eps):
        out = self.model(b)
        return training_loss(out, out[4], b["bond_edge_list"], 0.0, self.gamma, None, b.get("dp_norms"))[0]
tr = Step(model, 0.0, 1.0, lr=1e-4, loss_limit=None)
tr.prepare(batch, None)
for _ in range(2): tr.step(batch, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): tr.step(batch, None)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(batch, None); torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace_c4.json")
ev = [e for e in json.load(open("/tmp/trace_c4.json"))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    nm = e["name"].split("(")[0].replace("void ", "")[:90]
    agg[nm][0] += 1; agg[nm][1] += e["dur"]
tot = sum(v[1] for v in agg.values())
lines = ["# c4 PCN step, %d proteins x %d beads, eager: %.2f ms/step (CUDA events, 3 steps); sum of kernel durations %.2f ms" % (args.proteins, cfg["n_res"], ms, tot / 1e3)]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    lines.append("  %-92s n=%4d tot=%9.1f us avg=%8.2f us %5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
open(args.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
