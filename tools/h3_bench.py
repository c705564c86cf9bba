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
    """us per call, GPU time: the calls are captured into a CUDA graph and the replay is timed (the Python / ctypes / tensor-map encoding cost of a
    call, 20-60 us, would otherwise hide every kernel shorter than that)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (2 * iters) * 1e3   # us


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


def bench_epilogue(out):
    """epilogue variants of the encoder / decoder projections, back to back (GPU-bound timing), both tile orders"""
    lib = ops._lib.load()
    N_tok = 1025
    g = 32
    ys, xs_ = torch.meshgrid(torch.arange(g), torch.arange(g), indexing="ij")
    pos = torch.cat([torch.stack([ys.flatten(), xs_.flatten()], -1), torch.tensor([[g, 0]])], 0)[None].repeat(2, 1, 1).contiguous().to(DEV)
    tab = ops.rope2d_table(g + 1)
    for order in (0,):
        lib.siu3r_gemm_h3_order(order)
        for (M, N, K, what) in [(2050, 3072, 1024, "qkv"), (2050, 4096, 1024, "fc1"), (2050, 1024, 4096, "fc2"), (2050, 1024, 1024, "proj")]:
            x = ops.split(torch.randn(M, K, device=DEV))
            wt = ops.Weight(torch.randn(N, K, device=DEV) / K ** 0.5, torch.randn(N, device=DEV), H3)
            row = {"kind": "epilogue", "what": what, "order": order, "M": M, "N": N, "K": K}
            o = torch.empty(M, N, device=DEV)
            row["plain_us"] = timeit(lambda: ops.gemm(x, wt, out=o, precision=H3))
            stats = torch.zeros(M, 2, device=DEV, dtype=torch.int64)
            stats[:, 1] = K << 26          # mean 0, variance 1
            wln = wt.fold_ln(torch.ones(K, device=DEV), torch.zeros(K, device=DEV))
            if what == "qkv":
                C = 1024
                q = ops.Split.empty(M, 3 * C, device=DEV, unscaled=True)
                vth = ops.Split.empty(C, (M + 7) // 8 * 8, device=DEV, unscaled=True)
                row["rope_vt_split_us"] = timeit(lambda: ops.gemm(x, wt, out=q, precision=H3, rope=(pos.view(-1, 2), tab, 2 * C), vt=(vth, 2 * C, {}), unscaled=True))
                row["split_us"] = timeit(lambda: ops.gemm(x, wt, out=q, precision=H3, unscaled=True))
                row["rope_split_us"] = timeit(lambda: ops.gemm(x, wt, out=q, precision=H3, rope=(pos.view(-1, 2), tab, 2 * C), unscaled=True))
                row["vt_split_us"] = timeit(lambda: ops.gemm(x, wt, out=q, precision=H3, vt=(vth, 2 * C, {}), unscaled=True))
                row["ln_rope_vt_split_us"] = timeit(lambda: ops.gemm(x, wln, out=q, precision=H3, rope=(pos.view(-1, 2), tab, 2 * C), vt=(vth, 2 * C, {}),
                                                                     unscaled=True, ln_stats=stats))
                row["rope_f32_us"] = timeit(lambda: ops.gemm(x, wt, out=o, precision=H3, rope=(pos.view(-1, 2), tab, 2 * C)))
            elif what == "fc1":
                oh = ops.Split.empty(M, N, device=DEV)
                row["gelu_split_us"] = timeit(lambda: ops.gemm(x, wt, out=oh, precision=H3, act=1))
                row["ln_gelu_split_us"] = timeit(lambda: ops.gemm(x, wln, out=oh, precision=H3, act=1, ln_stats=stats))
                row["gelu_f32_us"] = timeit(lambda: ops.gemm(x, wt, out=o, precision=H3, act=1))
                row["relu_split_us"] = timeit(lambda: ops.gemm(x, wt, out=oh, precision=H3, act=2))
            else:
                r = torch.randn(M, N, device=DEV)
                row["residual_us"] = timeit(lambda: ops.gemm(x, wt, out=r, residual=r, precision=H3))
                rs_ = ops.Split.empty(M, N, device=DEV)
                row["residual_dual_us"] = timeit(lambda: ops.gemm(x, wt, out=rs_, out_f32=r, residual=r, precision=H3))
                row["residual_dual_stats_us"] = timeit(lambda: ops.gemm(x, wt, out=rs_, out_f32=r, residual=r, stats_out=stats, precision=H3))
            print(json.dumps(row), flush=True)
            out.append(row)
    lib.siu3r_gemm_h3_order(0)
    x = torch.randn(2050, 1024, device=DEV)
    w_, b_ = torch.randn(1024, device=DEV), torch.randn(1024, device=DEV)
    oh = ops.Split.empty(2050, 1024, device=DEV)
    out.append({"kind": "layernorm_h3", "us": timeit(lambda: ops.layernorm_h3([x], [(w_, b_)], 1e-6, outs=[oh]))})
    print(json.dumps(out[-1]), flush=True)


def bench_tiles(out):
    """Per-tile clock64 stamps of CTA pair 0 (siu3r_gemm_h3_debug_ts): where do the MMA warp and the epilogue warps wait?"""
    lib = ops._lib.load()
    M, N, K = 2050, 4096, 1024
    x = ops.split(torch.randn(M, K, device=DEV))
    wt = ops.Weight(torch.randn(N, K, device=DEV) / K ** 0.5, torch.randn(N, device=DEV), H3)
    o = torch.empty(M, N, device=DEV)
    oh = ops.Split.empty(M, N, device=DEV)
    ts = torch.zeros(64, dtype=torch.int64, device=DEV)
    for name, fn in (("plain", lambda: ops.gemm(x, wt, out=o, precision=H3)), ("gelu_split", lambda: ops.gemm(x, wt, out=oh, precision=H3, act=1)),
                     ("relu_split", lambda: ops.gemm(x, wt, out=oh, precision=H3, act=2)), ("gelu_f32", lambda: ops.gemm(x, wt, out=o, precision=H3, act=1))):
        for _ in range(3):
            fn()
        ts.zero_()
        lib.siu3r_gemm_h3_debug_ts(ts.data_ptr())
        fn()
        torch.cuda.synchronize()
        lib.siu3r_gemm_h3_debug_ts(None)
        t = ts.cpu().view(8, 8)
        t0 = int(t[0, 0])
        row = {"kind": "tiles", "what": name, "cols": "mma_wait_start, mma_start, mma_issued, epi_wait_start, epi_start, epi_end(w2), epi_end(w17)",
               "tiles": [[int(v) - t0 if int(v) else None for v in t[i, :7]] for i in range(8) if int(t[i, 0])]}
        print(json.dumps(row), flush=True)
        out.append(row)


def bench_twsweep(out):
    """Token tile width sweep WITH the model's epilogues (the cost model in pick_tw must see the exposed epilogue of single-buffered tiles)."""
    lib = ops._lib.load()
    g = 32
    ys, xs_ = torch.meshgrid(torch.arange(g), torch.arange(g), indexing="ij")
    pos = torch.cat([torch.stack([ys.flatten(), xs_.flatten()], -1), torch.tensor([[g, 0]])], 0)[None].repeat(2, 1, 1).contiguous().to(DEV)
    tab = ops.rope2d_table(g + 1)
    for (M, N, K, what) in [(2050, 3072, 1024, "qkv"), (2050, 4096, 1024, "fc1"), (2050, 1024, 4096, "fc2"), (2050, 1024, 1024, "proj"),
                            (2050, 2304, 768, "qkv"), (2050, 3072, 768, "fc1"), (2050, 768, 3072, "fc2"), (2050, 768, 768, "proj"), (2050, 1536, 768, "qkv_kv"),
                            (10752, 1024, 256, "fc1r"), (10752, 256, 1024, "fc2")]:
        x = ops.split(torch.randn(M, K, device=DEV))
        wt = ops.Weight(torch.randn(N, K, device=DEV) / K ** 0.5, torch.randn(N, device=DEV), H3)
        stats = torch.zeros(M, 2, device=DEV, dtype=torch.int64)
        stats[:, 1] = K << 26
        wln = wt.fold_ln(torch.ones(K, device=DEV), torch.zeros(K, device=DEV))
        if what.startswith("qkv"):
            C = N // 3 if what == "qkv" else N // 2
            rc = 2 * C if what == "qkv" else C
            q = ops.Split.empty(M, N, device=DEV, unscaled=True)
            vth = ops.Split.empty(C, (M + 7) // 8 * 8, device=DEV, unscaled=True)
            fn = lambda: ops.gemm(x, wln, out=q, precision=H3, rope=(pos.view(-1, 2)[:M], tab, rc), vt=(vth, rc, {}), unscaled=True, ln_stats=stats)
        elif what == "fc1":
            oh = ops.Split.empty(M, N, device=DEV)
            fn = lambda: ops.gemm(x, wln, out=oh, precision=H3, act=1, ln_stats=stats)
        elif what == "fc1r":
            oh = ops.Split.empty(M, N, device=DEV)
            fn = lambda: ops.gemm(x, wt, out=oh, precision=H3, act=2)
        else:
            r = torch.randn(M, N, device=DEV)
            rs_ = ops.Split.empty(M, N, device=DEV)
            fn = lambda: ops.gemm(x, wt, out=rs_, out_f32=r, residual=r, stats_out=stats, precision=H3)
        row = {"kind": "twsweep", "what": what, "M": M, "N": N, "K": K}
        row["auto_us"] = timeit(fn, iters=10, warm=2)
        for tw in (64, 96, 112, 128, 160, 192, 224, 256):
            lib.siu3r_gemm_h3_force(tw)
            row[f"tw{tw}"] = round(timeit(fn, iters=10, warm=2), 1)
        lib.siu3r_gemm_h3_force(0)
        print(json.dumps(row), flush=True)
        out.append(row)


def bench_limits(out):
    """Which side bounds the mainloop: the operand pipeline alone (no MMAs), the MMAs alone (no loads), both."""
    lib = ops._lib.load()
    for (M, N, K) in [(2050, 3072, 1024), (2050, 1024, 4096), (8192, 4096, 1024)]:
        x = ops.split(torch.randn(M, K, device=DEV))
        wt = ops.Weight(torch.randn(N, K, device=DEV) / K ** 0.5, torch.randn(N, device=DEV), H3)
        o = torch.empty(M, N, device=DEV)
        for tw in (64, 128, 192, 256):
            lib.siu3r_gemm_h3_force(tw)
            row = {"kind": "limits", "M": M, "N": N, "K": K, "tw": tw}
            for mode, name in ((0, "full_us"), (1, "tma_only_us"), (2, "mma_only_us")):
                lib.siu3r_gemm_h3_debug(mode)
                row[name] = timeit(lambda: ops.gemm(x, wt, out=o, precision=H3), iters=10, warm=2)
            lib.siu3r_gemm_h3_debug(0)
            print(json.dumps(row), flush=True)
            out.append(row)
    lib.siu3r_gemm_h3_force(0)


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


def bench_mhalf(out):
    """Cout <= 128 convs / linears: M = 128 MMAs (default) against the M = 256 kernel with half of its rows padding."""
    lib = ops._lib.load()
    for (n, h, w_, cin, cout, k) in [(2, 512, 512, 128, 128, 3), (2, 256, 256, 256, 128, 3), (2, 512, 512, 128, 83, 1), (2, 128, 128, 256, 128, 3)]:
        x = torch.randn(n, h, w_, cin, device=DEV)
        w = torch.randn(cout, k * k * cin, device=DEV) / (k * k * cin) ** 0.5
        wt3 = ops.Weight(w, torch.randn(cout, device=DEV), H3)
        xs = ops.split(x)
        o = torch.empty(n, h, w_, cout, device=DEV)
        row = {"kind": "mhalf", "shape": [n, h, w_, cin, cout, k]}
        for mode in (0, 1):
            lib.siu3r_gemm_h3_set_mhalf(mode)
            for tw in (0, 64, 128, 256):
                lib.siu3r_gemm_h3_force(tw)
                row[f"m{mode}_tw{tw}_us"] = timeit(lambda: ops.conv2d(xs, wt3, k, k, pad=k // 2, out=o, precision=H3, act=2), iters=10, warm=2)
        lib.siu3r_gemm_h3_force(0)
        lib.siu3r_gemm_h3_set_mhalf(1)
        row["tflops_m1"] = 2.0 * n * h * w_ * cout * k * k * cin / row["m1_tw0_us"] / 1e6
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
        torch.cuda.synchronize()
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(10):
            model(img, K)
        e0.record()
        torch.cuda.synchronize()
        t = s0.elapsed_time(e0) / 10 * 1e3
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
            {"gemm": bench_gemm, "conv": bench_conv, "flash": bench_flash, "model": bench_model, "epilogue": bench_epilogue, "limits": bench_limits, "mhalf": bench_mhalf, "tiles": bench_tiles, "twsweep": bench_twsweep}[w](res)
        except Exception as ex:  # keep going: one failing section must not lose the others
            print(json.dumps({"kind": w, "error": repr(ex)}), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/h3_bench.json", "w") as f:
        json.dump(res, f, indent=1)
