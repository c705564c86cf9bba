"""Separate fixed per-launch cost from per-k-block cost of the tcgen05 GEMM (vary K; cold vs warm L2; back-to-back launches)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
lib = ops._lib.load()


def t_single(fn, cold, iters=15):
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def t_b2b(fn, n=20):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n


print("tc2 disabled:", os.environ.get("SIU3R_DISABLE_TC2"))
for (M, N) in [(2050, 1024), (2050, 3072), (8200, 4096)]:
    for K in (32, 256, 1024, 4096):
        x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
        wt = ops.Weight(w, torch.zeros(N, device=dev), 1)
        out = torch.empty(M, N, device=dev)
        fn = lambda: lib.siu3r_gemm_tc(M, N, K, x.data_ptr(), None, K, wt.w.data_ptr(), None, K, out.data_ptr(), N, wt.bias.data_ptr(), None, 0, 0, 1.0, 1,
                                       ops._stream())
        fn(); torch.cuda.synchronize()
        print(f"M={M} N={N} K={K:5d}: cold {t_single(fn, True):7.1f} us  warm {t_single(fn, False):7.1f} us  b2b {t_b2b(fn):7.1f} us", flush=True)
# empty-kernel baseline: an eltwise on 4 floats
z = torch.zeros(4, device=dev)
fn = lambda: ops.eltwise(ops.ELT_COPY, z, out=z)
print(f"tiny eltwise: cold {t_single(fn, True):.1f} warm {t_single(fn, False):.1f} b2b {t_b2b(fn):.1f} us")
