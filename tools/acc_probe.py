"""Accuracy of the GEMM precision modes vs an fp64 reference, as a function of K (run on the GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
for K in (256, 1024, 4096, 16384):
    M, N = 512, 512
    g = torch.Generator().manual_seed(K)
    x = torch.randn(M, K, generator=g).to(dev); w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    ref = (x.double() @ w.double().t())
    sc = ref.abs().max()
    e = {}
    e["cublas_fp32"] = float(((x @ w.t()).double() - ref).abs().max() / sc)
    e["simt_fp32"] = float((ops.gemm_simt(x, w).double() - ref).abs().max() / sc)
    for prec in (1, 3):
        y = ops.gemm(x, ops.Weight(w, None, prec), precision=prec)
        d = (y.double() - ref)
        e[f"tc_prec{prec}"] = float(d.abs().max() / sc)
        e[f"tc_prec{prec}_meanbias"] = float((d * ref.sign()).mean() / ref.abs().mean())
    print(K, {k: f"{v:.2e}" for k, v in e.items()}, flush=True)
