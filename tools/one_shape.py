"""time one GEMM shape / form / epilogue inside a CUDA graph.  python tools/one_shape.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coarsegrainingvae_b200 import ops
dev = "cuda"
def graph_time(fn, reps=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3 / reps * 1e3
for (M, N, K) in [(48000, 512, 512), (16000, 512, 512), (16000, 1024, 512), (16000, 512, 2048), (16000, 512, 1536)]:
    G = torch.randn(M, K, device=dev); W = torch.randn(K, N, device=dev); add = torch.randn(M, N, device=dev); z = torch.randn(M, N, device=dev)
    out = torch.empty(M, N, device=dev)
    fl = 2.0 * M * N * K / 1e6
    t0 = graph_time(lambda: ops.gemm(ops.GEMM_NN, G, W, M, N, K, out=out))
    t1 = graph_time(lambda: ops.gemm(ops.GEMM_NN, G, W, M, N, K, add=add, out=out))
    t2 = graph_time(lambda: ops.gemm(ops.GEMM_NN, G, W, M, N, K, z_in=z, dact=1, out=out))
    print("NN M%6d N%5d K%5d plain %7.1f us %6.1f TF | +add %7.1f us | +dswish(z_in) %7.1f us" % (M, N, K, t0, fl / t0, t1, t2), flush=True)
