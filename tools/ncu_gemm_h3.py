"""A few launches of one h3 kernel shape for `ncu --set full -k regex:gemm_h3 -s 3 -c 1` (warm-up launches first).
usage: python tools/ncu_gemm_h3.py gemm M N K | conv H W Cin Cout k | flash B H N"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops
H3 = ops.PREC_H3
kind = sys.argv[1]
a = [int(v) for v in sys.argv[2:]]
if kind == "gemm":
    M, N, K = a
    x = ops.split(torch.randn(M, K, device="cuda"))
    wt = ops.Weight(torch.randn(N, K, device="cuda") / K ** 0.5, torch.randn(N, device="cuda"), H3)
    o = torch.empty(M, N, device="cuda")
    fn = lambda: ops.gemm(x, wt, out=o, precision=H3)
elif kind == "fc1":      # GELU + plane-pair output (the encoder's fc1 epilogue)
    M, N, K = a
    x = ops.split(torch.randn(M, K, device="cuda"))
    wt = ops.Weight(torch.randn(N, K, device="cuda") / K ** 0.5, torch.randn(N, device="cuda"), H3)
    o = ops.Split.empty(M, N, device="cuda")
    fn = lambda: ops.gemm(x, wt, out=o, precision=H3, act=1)
elif kind == "conv":
    h, w, cin, cout, k = a
    x = ops.split(torch.randn(1, h, w, cin, device="cuda"))
    wt = ops.Weight(torch.randn(cout, k * k * cin, device="cuda") / (k * k * cin) ** 0.5, torch.randn(cout, device="cuda"), H3)
    o = torch.empty(1, h, w, cout, device="cuda")
    fn = lambda: ops.conv2d(x, wt, k, k, pad=k // 2, out=o, precision=H3)
else:
    B, H, N = a
    C = H * 64
    qkv = torch.randn(B * N, 3 * C, device="cuda")
    qs = ops.split(qkv, unscaled=True)
    vt = ops.transpose_v_h3(qkv, 2 * C, N * 3 * C, 3 * C, B, N, H)
    o = ops.Split.empty(B * N, C, device="cuda")
    fn = lambda: ops.flash_attn_h3(qs, 0, qs, C, vt, 0, B, H, N, N, 0.125, out=o)
for _ in range(6):
    fn()
torch.cuda.synchronize()
print("done")
