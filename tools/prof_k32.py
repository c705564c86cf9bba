import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops
dev = "cuda"
M, N, K = 8200, 4096, 32
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev)
wt = ops.Weight(w, torch.zeros(N, device=dev), 1)
out = torch.empty(M, N, device=dev)
for _ in range(3):
    ops.gemm(x, wt, out=out, precision=1, a_rounded=True)
torch.cuda.synchronize()
