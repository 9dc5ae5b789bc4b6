"""Isolated timing of the 350-row (atom-level) contractions of the chignolin step, all three operand forms, inside a CUDA
graph.  CGVAE_MMA_GEMM=0 / CGVAE_TCGEN05=0 select the SIMT tiles for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coarsegrainingvae_b200 import ops
dev = "cuda"


def bench(fn_list, iters=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fn_list:
            f()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fn_list:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * len(fn_list)) * 1e3


print("env mma=%s tcgen05=%s" % (os.environ.get("CGVAE_MMA_GEMM", "1"), os.environ.get("CGVAE_TCGEN05", "1")))
for (M, N, K) in [(350, 1800, 600), (350, 600, 600), (2000, 1536, 512)]:
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev)
    gy = torch.randn(M, N, device=dev)
    nt = bench([lambda: ops.gemm(ops.GEMM_NT, A, W, M, N, K) for _ in range(8)])
    nn = bench([lambda: ops.gemm(ops.GEMM_NN, gy, W, M, K, N) for _ in range(8)])
    tn = bench([lambda: ops.gemm(ops.GEMM_TN, gy, A, N, K, M) for _ in range(8)])
    fl = 2.0 * M * N * K / 1e6
    print("M=%d N=%d K=%d  NT %.1f us (%.1f TF)  NN %.1f us (%.1f TF)  TN %.1f us (%.1f TF)" %
          (M, N, K, nt, fl / nt, nn, fl / nn, tn, fl / tn), flush=True)
