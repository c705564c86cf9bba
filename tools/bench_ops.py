"""Kernel-level timing on the GPU box (CUDA events, L2 flushed between iterations).  Not the contract bench (bench.py)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


res = []
for prec in (1, 3):
    for (M, N, K) in [(2050, 3072, 1024), (2050, 1024, 1024), (2050, 4096, 1024), (2050, 1024, 4096), (1025, 2304, 768), (1025, 3072, 768),
                      (8200, 3072, 1024), (8200, 4096, 1024), (262144, 256, 160), (65536, 256, 256), (5376, 1024, 1024)]:
        x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
        wt = ops.Weight(w, torch.zeros(N, device=dev), prec)
        out = torch.empty(M, N, device=dev)
        if prec == 3:
            hi, lo = ops.split_tf32(x)
            lib = ops._lib.load()
            fn = lambda: lib.siu3r_gemm_tc(M, N, K, hi.data_ptr(), lo.data_ptr(), K, wt.w.data_ptr(), wt.w_lo.data_ptr(), K, out.data_ptr(), N,
                                           wt.bias.data_ptr(), None, 0, 0, 1.0, 3, ops._stream())
        else:
            fn = lambda: ops.gemm(x, wt, out=out, precision=1)
        ms = timeit(fn)
        torch.backends.cuda.matmul.allow_tf32 = True
        ms_ref = timeit(lambda: torch.nn.functional.linear(x, w, out=None))
        torch.backends.cuda.matmul.allow_tf32 = False
        res.append(dict(op="gemm", prec=prec, M=M, N=N, K=K, ms=ms, tflops=2 * M * N * K / ms / 1e9, cublas_tf32_ms=ms_ref,
                        cublas_tflops=2 * M * N * K / ms_ref / 1e9))
        print(res[-1], flush=True)
for prec in (1,):
    for (Nb, H, W, Cin, Cout) in [(1, 512, 512, 256, 256), (1, 256, 256, 256, 256), (1, 128, 128, 256, 256), (1, 512, 512, 128, 128), (1, 64, 64, 256, 256)]:
        x = torch.randn(Nb, H, W, Cin, device=dev); w = torch.randn(Cout, 9 * Cin, device=dev) / (9 * Cin) ** 0.5
        wt = ops.Weight(w, torch.zeros(Cout, device=dev), prec)
        out = torch.empty(Nb, H, W, Cout, device=dev)
        ms = timeit(lambda: ops.conv2d(x, wt, 3, 3, pad=1, out=out, precision=prec))
        fl = 2 * Nb * H * W * Cout * 9 * Cin
        res.append(dict(op="conv3x3", prec=prec, H=H, W=W, Cin=Cin, Cout=Cout, ms=ms, tflops=fl / ms / 1e9))
        print(res[-1], flush=True)
for prec in (1, 3):
    for (B, H, N) in [(2, 16, 1025), (1, 12, 1025), (8, 16, 1025)]:
        qkv = torch.randn(B, N, 3, H, 64, device=dev)
        out = torch.empty(B, N, H * 64, device=dev)
        bs, ts = N * 3 * H * 64, 3 * H * 64
        ms = timeit(lambda: ops.flash_attn_d64(qkv, 0, bs, ts, qkv, H * 64, bs, ts, qkv, 2 * H * 64, bs, ts, out, B, H, N, N, 0.125, prec))
        fl = 4 * B * H * N * N * 64
        res.append(dict(op="flash", prec=prec, B=B, H=H, N=N, ms=ms, tflops=fl / ms / 1e9))
        print(res[-1], flush=True)
for (B, H, N) in [(2, 16, 1025), (1, 12, 1025), (8, 16, 1025)]:
    C = H * 64
    qkv = torch.randn(B, N, 3, H, 64, device=dev)
    out = torch.empty(B, N, C, device=dev)
    ms = timeit(lambda: ops.flash_attn_tc(qkv, 0, N * 3 * C, 3 * C, 3 * C, qkv, C, N * 3 * C, 3 * C, 3 * C, qkv, 2 * C, N * 3 * C, 3 * C, out, B, H, N, N, 0.125))
    res.append(dict(op="flash_tc(+transpose)", B=B, H=H, N=N, ms=ms, tflops=4 * B * H * N * N * 64 / ms / 1e9))
    print(res[-1], flush=True)
# elementwise bandwidth sanity
x = torch.randn(64 * 1024 * 1024, device=dev)
ms = timeit(lambda: ops.eltwise(ops.ELT_RELU, x, out=x))
print(dict(op="relu_inplace", GBs=2 * x.numel() * 4 / ms / 1e6))
w_, b_ = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
xx = x.view(-1, 1024); yy = torch.empty_like(xx)
ms = timeit(lambda: ops.layernorm(xx, w_, b_, 1e-6, out=yy))
print(dict(op="layernorm1024", GBs=2 * x.numel() * 4 / ms / 1e6))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_ops.json", "w"), indent=1)
