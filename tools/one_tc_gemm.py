"""One tcgen05 GEMM per operand form at the protein-config shape, for `ncu --set full` captures.
    ncu --set full -k regex:gemm_tc --launch-skip 6 -c 6 python tools/one_tc_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coarsegrainingvae_b200 import ops
M, N, K = 16000, 2048, 512
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
G = torch.randn(M, N, device="cuda")
for _ in range(3):
    y = ops.gemm(ops.GEMM_NT, A, W, M, N, K, bias=b, act=1)
    gx = ops.gemm(ops.GEMM_NN, G, W, M, K, N)
torch.cuda.synchronize()
