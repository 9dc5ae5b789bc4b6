"""tcgen05 3xTF32 GEMM vs float64, all operand forms; SIMT path timed beside it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coarsegrainingvae_b200 import ops
dev = "cuda"
def rel(a, b):
    return float((a.double().cpu() - b).abs().max() / b.abs().max())
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("tcgen05 enabled:", os.environ.get("CGVAE_TCGEN05", "1"))
for (M, N, K) in [(1024, 1024, 64), (1024, 1024, 512), (2000, 600, 600), (700, 1800, 600), (4000, 2048, 512), (16000, 2048, 512), (1000, 130, 77)]:
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    gy = torch.randn(M, N, generator=g)
    Ad, Wd, bd, gyd = A.to(dev), W.to(dev), b.to(dev), gy.to(dev)
    ref = A.double() @ W.double().t() + b.double()
    y, z = ops.gemm(ops.GEMM_NT, Ad, Wd, M, N, K, bias=bd, act=1, z_out=True)
    torch.cuda.synchronize()
    e_nt = rel(z, ref)
    got = ops.gemm(ops.GEMM_NN, gyd, Wd, M, K, N)
    e_nn = rel(got, gy.double() @ W.double())
    got = ops.gemm(ops.GEMM_TN, gyd, Ad, N, K, M)
    e_tn = rel(got, gy.double().t() @ A.double())
    ms_nt = t(lambda: ops.gemm(ops.GEMM_NT, Ad, Wd, M, N, K, bias=bd, act=1))
    ms_nn = t(lambda: ops.gemm(ops.GEMM_NN, gyd, Wd, M, K, N))
    ms_tn = t(lambda: ops.gemm(ops.GEMM_TN, gyd, Ad, N, K, M))
    fl = 2.0 * M * N * K / 1e9
    print("M%6d N%5d K%5d  err NT %.2e NN %.2e TN %.2e | NT %.3f ms %6.1f TF  NN %.3f ms %6.1f TF  TN %.3f ms %6.1f TF" %
          (M, N, K, e_nt, e_nn, e_tn, ms_nt, fl / ms_nt, ms_nn, fl / ms_nn, ms_tn, fl / ms_tn), flush=True)
