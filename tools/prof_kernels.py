"""Small driver for `ncu --set full` captures: a few launches of the top kernels at bench shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops, synth
from siu3r_b200.renderer import camera_matrices
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
if which == "gemm":
    for (M, N, K) in [(2050, 3072, 1024), (2050, 4096, 1024), (1025, 2304, 768)]:
        x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
        wt = ops.Weight(w, torch.zeros(N, device=dev), 1)
        out = torch.empty(M, N, device=dev)
        for _ in range(3):
            ops.gemm(x, wt, out=out, precision=1, a_rounded=True)
elif which == "conv":
    x = torch.randn(1, 512, 512, 256, device=dev); w = torch.randn(256, 9 * 256, device=dev) / 48
    wt = ops.Weight(w, torch.zeros(256, device=dev), 1)
    out = torch.empty(1, 512, 512, 256, device=dev)
    for _ in range(3):
        ops.conv2d(x, wt, 3, 3, pad=1, out=out, precision=1, a_rounded=True)
elif which == "flash":
    B, H, N = 2, 16, 1025
    qkv = torch.randn(B, N, 3, H, 64, device=dev)
    out = torch.empty(B, N, H * 64, device=dev)
    bs, ts = N * 3 * H * 64, 3 * H * 64
    C_ = H * 64
    for _ in range(3):
        ops.flash_attn_tc(qkv, 0, bs, ts, 3 * C_, qkv, C_, bs, ts, 3 * C_, qkv, 2 * C_, bs, ts, out, B, H, N, N, 0.125)
elif which == "raster":
    G, H, W = 500000, 512, 512
    sc = synth.raster_scene(G, H, W, seed=0, pixel_aligned=True)
    view, full, campos, tx, ty = camera_matrices(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
    a = [sc[k].to(dev) for k in ("means", "covariances", "harmonics", "opacities")]
    cam = [view[0].to(dev), full[0].to(dev), campos[0].to(dev), torch.zeros(3, device=dev)]
    for _ in range(3):
        ops.raster_forward(a[0], a[1], a[2], a[3], cam[0], cam[1], cam[2], cam[3], float(tx[0]), float(ty[0]), H, W, 4, sh_layout=1)
torch.cuda.synchronize()
