import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coarsegrainingvae_b200 import ops
M, N, K = 16000, 2048, 512
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
for _ in range(3):
    y = ops.gemm(ops.GEMM_NT, A, W, M, N, K, bias=b, act=1)
torch.cuda.synchronize()
