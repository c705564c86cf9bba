import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops
dev = "cuda"
lib = ops._lib.load()
for (M, N, K) in [(8200, 4096, 32), (8200, 4096, 1024), (2050, 3072, 1024)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev)
    wt = ops.Weight(w, torch.zeros(N, device=dev), 1)
    out = torch.empty(M, N, device=dev)
    dbg = torch.zeros(16, 8, dtype=torch.int64, device=dev)
    ops.gemm(x, wt, out=out, precision=1, a_rounded=True)
    lib.siu3r_gemm_debug_set(dbg.data_ptr())
    ops.gemm(x, wt, out=out, precision=1, a_rounded=True)
    torch.cuda.synchronize()
    lib.siu3r_gemm_debug_set(None)
    d = dbg.cpu()
    print(f"M={M} N={N} K={K}: stamps relative to CTA start (cycles): [setup_done, first_tma_issued, first_full, mma_all_issued, epi_start, epi_end, end]")
    for i in range(4):
        r = d[i]
        print("  cta", i, [int(r[j] - r[0]) for j in range(1, 8)])
