"""Micro-benchmarks of the h3 kernels on the model's own shapes, next to the TF32 kernels (CUDA events, L2-cold rotation of buffers).
python tools/h3_bench.py [gemm] [conv] [flash] [model] -> JSON lines on stdout + gpurun_out/h3_bench.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops  # noqa: E402

DEV = "cuda"
H3 = ops.PREC_H3


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3   # us


def bench_gemm(out):
    shapes = [(2050, 3072, 1024), (2050, 1024, 1024), (2050, 4096, 1024), (2050, 1024, 4096), (1025, 768, 3072), (1025, 3072, 768), (10752, 1024, 1024),
              (10752, 1024, 256), (10752, 256, 1024), (262144, 83, 256), (100, 2048, 256), (32768, 256, 256)]
    for (M, N, K) in shapes:
        x = torch.randn(M, K, device=DEV)
        w = torch.randn(N, K, device=DEV) / K ** 0.5
        b = torch.randn(N, device=DEV)
        wt3, wt1 = ops.Weight(w, b, H3), ops.Weight(w, b, 1)
        xs = ops.split(x)
        xr = ops.round_tf32(x)
        o = torch.empty(M, N, device=DEV)
        oh = ops.Split.empty(M, N, device=DEV)
        row = {"kind": "gemm", "M": M, "N": N, "K": K}
        row["tf32_us"] = timeit(lambda: ops.gemm(xr, wt1, out=o, precision=1, a_rounded=True))
        row["h3_us"] = timeit(lambda: ops.gemm(xs, wt3, out=o, precision=H3))
        row["h3_split_out_us"] = timeit(lambda: ops.gemm(xs, wt3, out=oh, precision=H3))
        lib = ops._lib.load()
        best = None
        for tw in (64, 96, 112, 128, 160, 192, 224, 256):
            lib.siu3r_gemm_h3_force(tw)
            t = timeit(lambda: ops.gemm(xs, wt3, out=o, precision=H3), iters=10, warm=2)
            row[f"h3_tw{tw}_us"] = t
            if best is None or t < best[1]:
                best = (tw, t)
        lib.siu3r_gemm_h3_force(0)
        row["h3_best_tw"] = best[0]
        fl = 2.0 * M * N * K
        row["tf32_tflops"] = fl / row["tf32_us"] / 1e6
        row["h3_tflops"] = fl / row["h3_us"] / 1e6
        torch.backends.cuda.matmul.allow_tf32 = True
        row["cublas_tf32_us"] = timeit(lambda: torch.addmm(b, x, w.t(), out=o))
        torch.backends.cuda.matmul.allow_tf32 = False
        print(json.dumps(row), flush=True)
        out.append(row)


def bench_conv(out):
    for (n, h, w_, cin, cout, k) in [(1, 512, 512, 256, 256, 3), (1, 256, 256, 256, 256, 3), (1, 128, 128, 256, 256, 3), (1, 64, 64, 256, 256, 3),
                                     (1, 32, 32, 256, 256, 3), (1, 16, 16, 256, 256, 3), (1, 512, 512, 128, 128, 3), (1, 256, 256, 256, 128, 3)]:
        x = torch.randn(n, h, w_, cin, device=DEV)
        w = torch.randn(cout, k * k * cin, device=DEV) / (k * k * cin) ** 0.5
        b = torch.randn(cout, device=DEV)
        wt3, wt1 = ops.Weight(w, b, H3), ops.Weight(w, b, 1)
        xs, xr = ops.split(x), ops.round_tf32(x)
        o = torch.empty(n, h, w_, cout, device=DEV)
        row = {"kind": "conv", "shape": [n, h, w_, cin, cout, k]}
        row["tf32_us"] = timeit(lambda: ops.conv2d(xr, wt1, k, k, pad=k // 2, out=o, precision=1, a_rounded=True))
        row["h3_us"] = timeit(lambda: ops.conv2d(xs, wt3, k, k, pad=k // 2, out=o, precision=H3))
        lib = ops._lib.load()
        for tw in (64, 128, 256):
            lib.siu3r_gemm_h3_force(tw)
            row[f"h3_tw{tw}_us"] = timeit(lambda: ops.conv2d(xs, wt3, k, k, pad=k // 2, out=o, precision=H3), iters=10, warm=2)
        lib.siu3r_gemm_h3_force(0)
        fl = 2.0 * n * h * w_ * cout * k * k * cin
        row["tf32_tflops"] = fl / row["tf32_us"] / 1e6
        row["h3_tflops"] = fl / row["h3_us"] / 1e6
        print(json.dumps(row), flush=True)
        out.append(row)


def bench_flash(out):
    for (B, H, N) in [(2, 16, 1025), (2, 12, 1025), (8, 16, 1025)]:
        C = H * 64
        qkv = torch.randn(B * N, 3 * C, device=DEV)
        qs = ops.split(qkv, unscaled=True)
        vt = ops.transpose_v_h3(qkv, 2 * C, N * 3 * C, 3 * C, B, N, H)
        o = ops.Split.empty(B * N, C, device=DEV)
        row = {"kind": "flash", "B": B, "H": H, "N": N}
        row["h3_us"] = timeit(lambda: ops.flash_attn_h3(qs, 0, qs, C, vt, 0, B, H, N, N, 0.125, out=o))
        qr = ops.round_tf32(qkv)
        of = torch.empty(B * N, C, device=DEV)
        row["tf32_us"] = timeit(lambda: ops.flash_attn_tc(qr, 0, N * 3 * C, 3 * C, 3 * C, qr, C, N * 3 * C, 3 * C, 3 * C, qr, 2 * C, N * 3 * C, 3 * C, of, B, H, N, N, 0.125))
        fl = 4.0 * B * H * N * N * 64
        row["h3_tflops"] = fl / row["h3_us"] / 1e6
        row["tf32_tflops"] = fl / row["tf32_us"] / 1e6
        print(json.dumps(row), flush=True)
        out.append(row)


def bench_model(out):
    from siu3r_b200 import synth
    from siu3r_b200.model import ModelCfg, SIU3RModel
    for prec in ("h3", "tf32"):
        model = SIU3RModel(ModelCfg(image_size=(512, 512)), precision=prec)
        model.load_state_dict(synth.make_state_dict())
        model.cuda()
        model.enable_cuda_graph()
        img, K = synth.pair_inputs(1, 2, 512)
        img, K = img.cuda(), K.cuda()
        for _ in range(3):
            model(img, K)
        torch.cuda.synchronize()
        t = timeit(lambda: model(img, K), iters=10, warm=1)
        row = {"kind": "model", "precision": prec, "ms_per_pair_serial": t / 1e3, "pairs_per_s": 1e6 / t}
        print(json.dumps(row), flush=True)
        out.append(row)
        del model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "conv", "flash", "model"]
    res = []
    for w in what:
        try:
            {"gemm": bench_gemm, "conv": bench_conv, "flash": bench_flash, "model": bench_model}[w](res)
        except Exception as ex:  # keep going: one failing section must not lose the others
            print(json.dumps({"kind": w, "error": repr(ex)}), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/h3_bench.json", "w") as f:
        json.dump(res, f, indent=1)
