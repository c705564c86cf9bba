"""Tile / kernel sweep of the TF32 tensor-core GEMM on the shapes of the 512x512 forward: every kernel family forced in turn
(siu3r_gemm_force), checked against torch fp32 and timed cold-L2 (flush between iterations) and back-to-back."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops
dev = "cuda"
lib = ops._lib.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
NAMES = {0: "auto", 1: "tc3", 3: "tc2", 4: "tc1", 64: "tc3_tw64", 128: "tc3_tw128", 176: "tc3_tw176", 256: "tc3_tw256"}
SHAPES = [(2050, 3072, 1024), (2050, 1024, 1024), (2050, 4096, 1024), (2050, 1024, 4096), (1025, 2304, 768), (1025, 768, 768), (1025, 1536, 768),
          (1025, 3072, 768), (1025, 768, 3072), (10752, 1024, 1024), (10752, 256, 1024), (10752, 1024, 256), (2048, 1024, 1024), (10752, 768, 256), (10752, 1024, 1024),
          (32768, 100, 256), (8200, 3072, 1024), (8200, 1024, 4096)]


def timeit(fn, cold, iters=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def b2b(fn, n=20):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n


rows = []
for (M, N, K) in SHAPES:
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
    wt = ops.Weight(w, bias, 1)
    xr = ops.round_tf32(x)
    ref = (xr.double() @ wt.w.double().t() + bias.double()).float()
    out = torch.empty(M, N, device=dev)
    line = dict(M=M, N=N, K=K)
    for f in (4, 3, 1, 0):
        lib.siu3r_gemm_force(f)
        out.zero_()
        try:
            ops.gemm(xr, wt, out=out, precision=1, a_rounded=True)
            torch.cuda.synchronize()
            err = float((out - ref).abs().max())
            out.zero_()
            ops.gemm(xr, wt, out=out, precision=1, a_rounded=True, act=ops.ACT_GELU, residual=res)
            err2 = float((out - (torch.nn.functional.gelu(ref) + res)).abs().max())
            out.copy_(res)
            ops.gemm(xr, wt, out=out, precision=1, a_rounded=True, residual=out, round_out=True)     # in-place residual (+ split-K where chosen)
            err2 = max(err2, float((out - (ref + res)).abs().max()) - 2e-3 * float((ref + res).abs().max()) * 0.5)
            fn = lambda: ops.gemm(xr, wt, out=out, precision=1, a_rounded=True)
            tc, tw = timeit(fn, True), b2b(fn)
            line[NAMES[f]] = dict(us_cold=round(tc, 1), us_b2b=round(tw, 1), tflops_b2b=round(2 * M * N * K / tw / 1e6, 1), err=err, err_gelu_res=err2)
        except Exception as ex:
            line[NAMES[f]] = dict(error=repr(ex)[:200])
    lib.siu3r_gemm_force(0)
    # grouped launch (two problems, different operands) incl. in-place residual
    x2 = torch.randn(M, K, device=dev); w2 = torch.randn(N, K, device=dev) / K ** 0.5
    wt2 = ops.Weight(w2, bias.clone(), 1); x2r = ops.round_tf32(x2)
    ref2 = (x2r.double() @ wt2.w.double().t() + bias.double()).float()
    o1, o2 = res.clone(), res.clone()
    ops.gemm_group2([xr, x2r], [wt, wt2], outs=[o1, o2], residuals=[o1, o2], a_rounded=True)
    torch.cuda.synchronize()
    line["group2_err"] = max(float((o1 - (ref + res)).abs().max()), float((o2 - (ref2 + res)).abs().max()))
    fn = lambda: ops.gemm_group2([xr, x2r], [wt, wt2], outs=[o1, o2], a_rounded=True)
    line["group2_us_b2b"] = round(b2b(fn), 1)
    torch.backends.cuda.matmul.allow_tf32 = True
    fn = lambda: torch.nn.functional.linear(xr, wt.w, bias)
    line["cublas_tf32"] = dict(us_cold=round(timeit(fn, True), 1), us_b2b=round(b2b(fn), 1))
    torch.backends.cuda.matmul.allow_tf32 = False
    rows.append(line)
    print(json.dumps(line), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/gemm_sweep.json", "w"), indent=1)
