"""GEMM calls of one eager chignolin training step: (form, M, N, K) -> count, mean CUDA-event time per call.
    python tools/gemm_shapes.py [--workload c2_chignolin]"""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coarsegrainingvae_b200 as cg
from coarsegrainingvae_b200 import ops, synthetic
from coarsegrainingvae_b200.factory import build_cgvae
from coarsegrainingvae_b200.train import TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2_chignolin")
args = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = dict(synthetic.CONFIGS[args.workload])
rad = lambda xyz, c: ops.radius_graph(torch.as_tensor(xyz, dtype=torch.float32, device=dev), c).cpu().numpy()
torch.manual_seed(1)
if cfg["kind"] == "pcn":
    from coarsegrainingvae_b200.factory import build_pcn
    from coarsegrainingvae_b200.train import training_loss
    b = synthetic.pcn_batch(cfg, 0, rad, n_proteins=cfg["batch"])
    model = build_pcn(cfg["n_basis"], cfg["n_rbf"], cfg["cg_cutoff"], cfg["dec_nconv"]).to(dev)

    class _Step(TrainStep):
        def _loss(self, bb, eps):
            out = self.model(bb)
            return training_loss(out, out[4], bb["bond_edge_list"], 0.0, self.gamma, None, bb.get("dp_norms"))[0]
    tr = _Step(model, 0.0, 1.0, lr=1e-4, loss_limit=None)
else:
    b = synthetic.cgvae_batch(cfg, 0, rad, cg.CG_collate)
    model = build_cgvae(cfg["n_basis"], cfg["n_rbf"], cfg["enc_nconv"], cfg["dec_nconv"], cfg["atom_cutoff"], cfg["cg_cutoff"], cfg["n_cgs"]).to(dev)
    tr = TrainStep(model, cfg["beta"], cfg["gamma"])
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
tr.prepare(b, None)
for _ in range(3):
    tr.step(b, None)
ops.FORK_ENABLED = False
ops.TIMER = ops.KernelTimer(["gemm"])
for _ in range(5):
    tr.step(b, None)
torch.cuda.synchronize()
agg = collections.defaultdict(list)
for ms, m in ops.TIMER.summary()["gemm"]:
    agg[(m["form"], m["M"], m["N"], m["K"], m.get("lda", 0), m.get("ldb", 0))].append(ms * 1e3)
names = {0: "NT", 1: "NN", 2: "TN"}
tot = 0.0
print("%-3s %7s %7s %7s %6s %6s %6s %9s %9s %8s" % ("op", "M", "N", "K", "lda", "ldb", "calls", "us/call", "us/step", "TFLOP/s"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    n = len(v) / 5
    us = sum(v) / len(v)
    tot += us * n
    print("%-3s %7d %7d %7d %6d %6d %6.0f %9.2f %9.1f %8.2f" % (names[k[0]], k[1], k[2], k[3], k[4], k[5], n, us, us * n, 2.0 * k[1] * k[2] * k[3] / us / 1e6))
print("total us/step (event-timed, includes launch gaps):", round(tot, 1))
