"""GPU parity tests of every C-ABI kernel against a plain PyTorch fp32 reference of the same op (and, for the rasterizer /
RoPE, against the C oracle).  Run on the B200 box: `python -m pytest tests -m gpu`.

Tolerances: TF32 tensor-core mode = the reference's own GPU numerics (croco/croco.py:13): ~1e-3 relative to the output
scale; 3xTF32 mode and the fp32 SIMT kernels: ~1e-5.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _no_tf32_in_reference():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(2050, 3072, 1024), (300, 768, 768), (100, 256, 2048), (128, 64, 32), (1025, 96, 1024), (77, 83, 256),
                                   (4096, 4096, 1024), (513, 21, 256)])
@pytest.mark.parametrize("prec", [1, 3])
def test_gemm_tc_plain(M, N, K, prec):
    from siu3r_b200 import ops
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    wt = ops.Weight(w, b, prec)
    y = ops.gemm(x, wt, precision=prec)
    ref = F.linear(x, w, b)
    tol = 3e-3 if prec == 1 else 2e-5
    assert rel_err(y, ref) < tol, (rel_err(y, ref), M, N, K, prec)
    # agreement with our own fp32 SIMT kernel (independent code path)
    y2 = ops.gemm_simt(x, w, b)
    assert rel_err(y2, ref) < 2e-5


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("prec", [1, 3])
def test_gemm_tc_epilogue_strided(act, prec):
    from siu3r_b200 import ops
    M, N, K = 777, 320, 512
    xbig = rnd(M, K + 64, seed=4)
    x = xbig[:, :K]  # lda = K + 64
    w, b = rnd(N, K, seed=5, scale=K ** -0.5), rnd(N, seed=6)
    res_big = rnd(M, N + 32, seed=7)
    res = res_big[:, :N]
    out_big = torch.full((M, N + 16), 7.0, device=DEV)
    out = out_big[:, :N]
    wt = ops.Weight(w, b, prec)
    ops.gemm(x if prec == 1 else x, wt, out=out, act=act, residual=res, alpha=0.5, precision=prec)
    ref = 0.5 * F.linear(x, w) + b
    ref = F.gelu(ref) if act == 1 else (F.relu(ref) if act == 2 else ref)
    ref = ref + res
    tol = 3e-3 if prec == 1 else 3e-5
    assert rel_err(out, ref) < tol
    assert torch.all(out_big[:, N:] == 7.0)  # ldc padding untouched


def test_gemm_inplace_residual():
    from siu3r_b200 import ops
    M, N, K = 1025, 1024, 1024
    x, w, b = rnd(M, K, seed=8), rnd(N, K, seed=9, scale=K ** -0.5), rnd(N, seed=10)
    r = rnd(M, N, seed=11)
    ref = F.linear(x, w, b) + r
    wt = ops.Weight(w, b, 3)
    ops.gemm(x, wt, out=r, residual=r, precision=3)
    assert rel_err(r, ref) < 3e-5


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 256, 3), (1, 32, 64, 96, 256, 3), (1, 128, 128, 256, 256, 3), (1, 64, 64, 768, 256, 3),
                                   (2, 8, 16, 32, 83, 1), (1, 256, 256, 128, 128, 3), (1, 8, 48, 64, 256, 3), (3, 8, 16, 32, 512, 3)])
@pytest.mark.parametrize("prec", [1, 3])
def test_conv2d_tc(shape, prec):
    from siu3r_b200 import ops
    N, H, W, Cin, Cout, k = shape
    x = rnd(N, H, W, Cin, seed=12)
    w = rnd(Cout, Cin, k, k, seed=13, scale=(Cin * k * k) ** -0.5)
    b = rnd(Cout, seed=14)
    res = rnd(N, H, W, Cout, seed=15)
    wt = ops.Weight(w.permute(0, 2, 3, 1).reshape(Cout, -1), b, prec)
    y = ops.conv2d(x, wt, k, k, stride=1, pad=k // 2, act=ops.ACT_RELU, residual=res, precision=prec)
    ref = F.relu(F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=k // 2)).permute(0, 2, 3, 1) + res
    tol = 3e-3 if prec == 1 else 3e-5
    assert rel_err(y, ref) < tol, rel_err(y, ref)


@pytest.mark.parametrize("cfg", [(1, 64, 64, 4, 64, 3, 2, 1), (2, 32, 32, 4, 256, 7, 1, 3), (1, 32, 32, 768, 768, 3, 2, 1), (1, 4, 4, 96, 256, 3, 1, 1),
                                 (2, 64, 64, 64, 128, 3, 2, 1)])
def test_conv2d_im2col(cfg):
    from siu3r_b200 import ops
    N, H, W, Cin, Cout, k, s, p = cfg
    x = rnd(N, H, W, Cin, seed=16)
    w = rnd(Cout, Cin, k, k, seed=17, scale=(Cin * k * k) ** -0.5)
    b = rnd(Cout, seed=18)
    wt = ops.Weight(w.permute(0, 2, 3, 1).reshape(Cout, -1), b, 3)
    y = ops.conv2d(x, wt, k, k, stride=s, pad=p, precision=3)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, stride=s, padding=p).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 1e-4  # K up to 6912: error grows with the MMA chain length (tools/acc_probe.py)


@pytest.mark.parametrize("prec", [1, 3])
def test_conv2d_rowpacked_7x7_image(prec):
    """7x7 conv on a 3(+1)-channel image (dpt_gs_head input_merger): 1x7 row packing + 7x1 implicit GEMM, fused ReLU + residual."""
    from siu3r_b200 import ops
    N, H, W, Cin, Cout, k = 2, 40, 48, 4, 256, 7
    x = rnd(N, H, W, Cin, seed=61)
    x[..., 3] = 0
    w = rnd(Cout, Cin, k, k, seed=62, scale=(Cin * k * k) ** -0.5)
    b = rnd(Cout, seed=63)
    res = rnd(N, H, W, Cout, seed=64)
    wt = ops.Weight(w.permute(0, 2, 3, 1).reshape(Cout, -1), b, prec)
    y = ops.conv2d(x, wt, k, k, pad=3, act=ops.ACT_RELU, residual=res, precision=prec)
    ref = F.relu(F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=3)).permute(0, 2, 3, 1) + res
    assert rel_err(y, ref) < (3e-3 if prec == 1 else 3e-5)


def test_conv_transpose_as_gemm_pixel_shuffle():
    from siu3r_b200 import ops
    N, H, W, Cin, Cout, s = 2, 8, 8, 96, 96, 4
    x = rnd(N, H, W, Cin, seed=19)
    w = rnd(Cin, Cout, s, s, seed=20, scale=Cin ** -0.5)  # torch ConvTranspose2d layout
    b = rnd(Cout, seed=21)
    add = rnd(N, H * s, W * s, Cout, seed=22)
    w2 = w.permute(2, 3, 1, 0).reshape(s * s * Cout, Cin)  # [(dy,dx,co), ci]
    wt = ops.Weight(w2, b.repeat(s * s), 3)
    g = ops.gemm(x.view(-1, Cin), wt, precision=3)
    y = ops.pixel_shuffle(g, N, H, W, Cout, s, add=add)
    ref = F.conv_transpose2d(x.permute(0, 3, 1, 2), w, b, stride=s).permute(0, 2, 3, 1) + add
    assert rel_err(y, ref) < 3e-5


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C_", [1024, 768, 256])
def test_layernorm(C_):
    from siu3r_b200 import ops
    x = rnd(1031, C_, seed=23, scale=3.0) + 0.5
    w, b = rnd(C_, seed=24), rnd(C_, seed=25)
    add = rnd(1031, C_, seed=26)
    y = ops.layernorm(x, w, b, 1e-6, add=add)
    ref = F.layer_norm(x, (C_,), w, b, 1e-6) + add
    assert float((y - ref).abs().max()) < 2e-5


def test_rope2d_vs_oracle_and_torch(oracle_lib):
    from siu3r_b200 import ops
    from oracle import raster_oracle as RO
    B, N, H, D = 2, 1025, 16, 64
    qkv = rnd(B, N, 3, H, D, seed=27)
    ys, xs = torch.meshgrid(torch.arange(32), torch.arange(32), indexing="ij")
    pos = torch.stack([ys.flatten(), xs.flatten()], -1)
    pos = torch.cat([pos, torch.tensor([[32, 0]])], 0)[None].repeat(B, 1, 1).contiguous()
    q_ref = RO.rope2d(qkv[:, :, 0].contiguous().cpu().numpy(), pos.numpy())
    k_ref = RO.rope2d(qkv[:, :, 1].contiguous().cpu().numpy(), pos.numpy())
    v_before = qkv[:, :, 2].clone()
    posd = pos.to(DEV)
    ops.rope2d_(qkv, 0, posd, B, N, H, D, N * 3 * H * D, 3 * H * D, nparts=2, part_stride=H * D)
    assert np.abs(qkv[:, :, 0].cpu().numpy() - q_ref).max() < 3e-5
    assert np.abs(qkv[:, :, 1].cpu().numpy() - k_ref).max() < 3e-5
    assert torch.equal(qkv[:, :, 2], v_before)
    lib_err = __import__("siu3r_b200._lib", fromlist=["x"]).load().siu3r_rope2d(qkv.data_ptr(), posd.data_ptr(), 1, 1, 1, 6, 6, 6, 100.0, 1.0, 1, 0, 0, None)
    assert lib_err == -1  # D % 4 != 0 -> invalid argument (kernels.cu:94 contract)


@pytest.mark.parametrize("M,N,K,cols", [(2050, 3072, 1024, 2048), (1025, 768, 768, 768), (1025, 1536, 768, 768), (2050, 4096, 256, 4096)])
@pytest.mark.parametrize("prec", [1, 3])
def test_gemm_rope_epilogue_matches_gemm_then_rope2d(M, N, K, cols, prec):
    """siu3r_gemm_tc_rope == siu3r_gemm_tc followed by siu3r_rope2d on the first `cols` columns (all tile variants)."""
    from siu3r_b200 import ops
    x = rnd(M, K, seed=71)
    wt = ops.Weight(rnd(N, K, seed=72) / K ** 0.5, rnd(N, seed=73), prec)
    n_tok = M // 2 if M % 2 == 0 else M
    Bn = M // n_tok
    g = int((n_tok - 1) ** 0.5)
    ys, xs = torch.meshgrid(torch.arange(g), torch.arange(g), indexing="ij")
    pos = torch.cat([torch.stack([ys.flatten(), xs.flatten()], -1), torch.tensor([[g, 0]])], 0)[None].repeat(Bn, 1, 1).contiguous().to(DEV)
    assert pos.shape[1] == n_tok
    tab = ops.rope2d_table(g + 1)
    ref = ops.gemm(x, wt, precision=prec)
    ops.rope2d_(ref, 0, pos, Bn, n_tok, cols // 64, 64, n_tok * N, N)
    got = ops.gemm(x, wt, precision=prec, rope=(pos.view(-1, 2), tab, cols))
    assert float((got - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))
    assert torch.equal(got[:, cols:], ref[:, cols:])


def _attn_ref(q, k, v, scale, mask=None):
    s = torch.einsum("bhqd,bhkd->bhqk", q, k) * scale
    if mask is not None:
        s = s.masked_fill(mask, float("-inf"))
    return torch.einsum("bhqk,bhkd->bhqd", s.softmax(-1), v)


@pytest.mark.parametrize("Nq,Nk,H", [(1025, 1025, 16), (257, 257, 12), (100, 3075, 12), (17, 17, 16)])
@pytest.mark.parametrize("prec", [1, 3])
def test_flash_attn_d64(Nq, Nk, H, prec):
    from siu3r_b200 import ops
    B, D = 2, 64
    q, k, v = rnd(B, Nq, H, D, seed=28), rnd(B, Nk, H, D, seed=29), rnd(B, Nk, H, D, seed=30)
    out = torch.empty(B, Nq, H * D, device=DEV)
    ops.flash_attn_d64(q, 0, Nq * H * D, H * D, k, 0, Nk * H * D, H * D, v, 0, Nk * H * D, H * D, out, B, H, Nq, Nk, D ** -0.5, prec)
    ref = _attn_ref(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), D ** -0.5).permute(0, 2, 1, 3).reshape(B, Nq, H * D)
    tol = 3e-3 if prec == 1 else 3e-5
    assert rel_err(out, ref) < tol, rel_err(out, ref)


@pytest.mark.parametrize("Nq,Nk,H", [(1025, 1025, 16), (257, 257, 12), (100, 3075, 12), (17, 17, 16), (128, 256, 2)])
def test_flash_attn_tc(Nq, Nk, H):
    from siu3r_b200 import ops
    B, D = 2, 64
    q, k, v = rnd(B, Nq, H, D, seed=58), rnd(B, Nk, H, D, seed=59), rnd(B, Nk, H, D, seed=60)
    out = torch.empty(B, Nq, H * D, device=DEV)
    ops.flash_attn_tc(q, 0, Nq * H * D, H * D, H * D, k, 0, Nk * H * D, H * D, H * D, v, 0, Nk * H * D, H * D, out, B, H, Nq, Nk, D ** -0.5)
    ref = _attn_ref(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), D ** -0.5).permute(0, 2, 1, 3).reshape(B, Nq, H * D)
    assert rel_err(out, ref) < 5e-3, rel_err(out, ref)  # raw fp32 operands: the tensor core truncates them (the engine rounds q, k, v first)


def test_flash_attn_tc_inside_qkv_buffer():
    from siu3r_b200 import ops
    B, N, H, D = 2, 1025, 16, 64
    C = H * D
    qkv = rnd(B, N, 3, H, D, seed=61)
    out = torch.empty(B, N, C, device=DEV)
    ops.flash_attn_tc(qkv, 0, N * 3 * C, 3 * C, 3 * C, qkv, C, N * 3 * C, 3 * C, 3 * C, qkv, 2 * C, N * 3 * C, 3 * C, out, B, H, N, N, 0.125)
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    ref = _attn_ref(q, k, v, 0.125).permute(0, 2, 1, 3).reshape(B, N, C)
    assert rel_err(out, ref) < 5e-3


def test_flash_attn_inside_qkv_buffer():
    from siu3r_b200 import ops
    B, N, H, D = 2, 1025, 16, 64
    qkv = rnd(B, N, 3, H, D, seed=31)
    out = torch.empty(B, N, H * D, device=DEV)
    bs, ts = N * 3 * H * D, 3 * H * D
    ops.flash_attn_d64(qkv, 0, bs, ts, qkv, H * D, bs, ts, qkv, 2 * H * D, bs, ts, out, B, H, N, N, 0.125, 3)
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    ref = _attn_ref(q, k, v, 0.125).permute(0, 2, 1, 3).reshape(B, N, H * D)
    assert rel_err(out, ref) < 3e-5


@pytest.mark.parametrize("Nk", [512, 2048, 100, 8192, 333])
def test_attn_small_d32_masked(Nk):
    from siu3r_b200 import ops
    B, H, Nq, D = 2, 8, 100, 32
    q, k, v = rnd(B, Nq, H * D, seed=32), rnd(B, Nk, H * D, seed=33), rnd(B, Nk, H * D, seed=34)
    g = torch.Generator(device="cpu"); g.manual_seed(35)
    mask = (torch.rand(B, Nq, Nk, generator=g) < 0.7)
    mask[0, 3] = True   # fully masked rows -> attend everywhere
    mask[1, 99] = True
    out = torch.empty(B, Nq, H * D, device=DEV)
    ops.attn_small_d32(q, Nq * H * D, H * D, k, Nk * H * D, H * D, v, Nk * H * D, H * D, out, Nq * H * D, H * D, mask.to(torch.uint8).to(DEV), B, H,
                       Nq, Nk, D ** -0.5)
    m = mask.clone()
    m[m.all(-1)] = False
    qq, kk, vv = [t.view(B, -1, H, D).permute(0, 2, 1, 3) for t in (q, k, v)]
    ref = _attn_ref(qq, kk, vv, D ** -0.5, m[:, None].to(DEV)).permute(0, 2, 1, 3).reshape(B, Nq, H * D)
    assert rel_err(out, ref) < 2e-5
    out2 = torch.empty_like(out)
    ops.attn_small_d32(q, Nq * H * D, H * D, k, Nk * H * D, H * D, v, Nk * H * D, H * D, out2, Nq * H * D, H * D, None, B, H, Nq, Nk, D ** -0.5)
    ref2 = _attn_ref(qq, kk, vv, D ** -0.5).permute(0, 2, 1, 3).reshape(B, Nq, H * D)
    assert rel_err(out2, ref2) < 2e-5


def _msda_ref(value, shapes, loc, attw):
    """value [B,Lin,nH,hd], loc [B,Lq,nH,L,P,2] normalised, attw [B,Lq,nH,L,P] -> [B,Lq,nH*hd] (grid_sample formulation)."""
    B, _, nH, hd = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    vals = value.split([h * w for h, w in shapes], dim=1)
    grids = 2 * loc - 1
    outs = []
    for l, (h, w) in enumerate(shapes):
        vl = vals[l].flatten(2).transpose(1, 2).reshape(B * nH, hd, h, w)
        gl = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)
        outs.append(F.grid_sample(vl, gl, mode="bilinear", padding_mode="zeros", align_corners=False))
    aw = attw.transpose(1, 2).reshape(B * nH, 1, Lq, L * P)
    out = (torch.stack(outs, dim=-2).flatten(-2) * aw).sum(-1).view(B, nH * hd, Lq)
    return out.transpose(1, 2).contiguous()


@pytest.mark.parametrize("cfg", [(16, 64, [(32, 32)], 4, 5376), (8, 32, [(16, 16), (32, 32), (64, 64)], 4, 5376), (8, 32, [(2, 2), (4, 4), (8, 8)], 4, 84)])
def test_msdeform_attn(cfg):
    from siu3r_b200 import ops
    nH, hd, shapes, P, Lq = cfg
    B, L = 2, len(shapes)
    Lin = sum(h * w for h, w in shapes)
    value = rnd(B, Lin, nH * hd, seed=36)
    noff, nw = nH * L * P * 2, nH * L * P
    ow = torch.cat([rnd(B * Lq, noff, seed=37, scale=2.0), rnd(B * Lq, nw, seed=38)], 1).contiguous()
    g = torch.Generator(device="cpu"); g.manual_seed(39)
    ref_pts = torch.rand(Lq, 2, generator=g).to(DEV)
    out = torch.empty(B * Lq, nH * hd, device=DEV)
    ops.msdeform_attn(value.view(B * Lin, nH * hd), Lin, ow, ref_pts, shapes, P, B, Lq, nH, hd, out)
    offs = ow[:, :noff].view(B, Lq, nH, L, P, 2)
    attw = ow[:, noff:].view(B, Lq, nH, L * P).softmax(-1).view(B, Lq, nH, L, P)
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32, device=DEV)
    loc = ref_pts[None, :, None, None, None, :] + offs / norm[None, None, None, :, None, :]
    ref = _msda_ref(value.view(B, Lin, nH, hd), shapes, loc, attw)
    assert rel_err(out.view(B, Lq, -1), ref) < 2e-5


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [(16, 16, 32, 32, True), (32, 32, 64, 64, True), (32, 32, 128, 128, False), (32, 32, 16, 16, False),
                                 (128, 128, 256, 256, False), (64, 64, 128, 128, False), (16, 16, 2, 2, False)])
def test_resize_bilinear(cfg):
    from siu3r_b200 import ops
    H, W, OH, OW, align = cfg
    x = rnd(2, H, W, 64, seed=40)
    y = ops.resize_bilinear(x, OH, OW, align)
    ref = F.interpolate(x.permute(0, 3, 1, 2), size=(OH, OW), mode="bilinear", align_corners=align).permute(0, 2, 3, 1)
    assert float((y - ref).abs().max()) < 1e-5
    base = rnd(2, OH, OW, 64, seed=41)
    acc = base.clone()
    ops.resize_bilinear(x, OH, OW, align, out=acc, accumulate=True)
    assert float((acc - (ref + base)).abs().max()) < 1e-5


def test_eltwise_affine_misc():
    from siu3r_b200 import ops
    a, b = rnd(1000, 260, seed=42), rnd(1000, 260, seed=43)
    assert torch.allclose(ops.eltwise(ops.ELT_RELU, a), F.relu(a))
    assert torch.allclose(ops.eltwise(ops.ELT_ADD, a, b), a + b)
    assert torch.allclose(ops.eltwise(ops.ELT_ADD_RELU, a, b), F.relu(a + b))
    assert torch.allclose(ops.eltwise(ops.ELT_GELU, a), F.gelu(a), atol=1e-6)
    assert torch.allclose(ops.eltwise(ops.ELT_SIGMOID, a), torch.sigmoid(a), atol=1e-6)
    assert torch.allclose(ops.eltwise(ops.ELT_CLAMP01, a), a.clamp(0, 1))
    sc, sh = rnd(260, seed=44), rnd(260, seed=45)
    y = ops.rows_affine(a, scale=sc, shift=sh, add=b, relu=True)
    assert torch.allclose(y, F.relu(a * sc + sh + b), atol=1e-6)
    c = a.clone()
    ops.scale_(c, 10.0)
    assert torch.allclose(c, a * 10.0)
    hi, lo = ops.split_tf32(a)
    assert float((hi + lo - a).abs().max() / a.abs().max()) < 1e-6
    assert torch.all((hi.view(torch.int32) & 0x1FFF) == 0)


def test_layout_pool_dwconv_groupnorm():
    from siu3r_b200 import ops
    x = rnd(2, 3, 32, 48, seed=46)
    y = ops.nchw_to_nhwc(x, 4)
    assert torch.equal(y[..., :3], x.permute(0, 2, 3, 1)) and torch.all(y[..., 3] == 0)
    z = ops.nhwc_to_nchw(y, 3)
    assert torch.equal(z, x)
    f = rnd(2, 33, 47, 64, seed=47)
    mp = ops.maxpool3x3s2(f)
    ref = F.max_pool2d(f.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(mp, ref)
    # depthwise 3x3 on a token range with leading dim / batch stride, GELU fused
    B, Hh, Ww, Cc, ntok = 2, 16, 16, 256, 400
    tok = rnd(B, ntok, Cc, seed=48)
    w, bb = rnd(Cc, 1, 3, 3, seed=49), rnd(Cc, seed=50)
    out = torch.zeros(B, ntok, Cc, device=DEV)
    start = 100
    ops.dwconv3x3(tok.data_ptr() + 4 * start * Cc, Cc, ntok * Cc, B, Hh, Ww, Cc, w.view(Cc, 9).t().contiguous(), bb,
                  out.data_ptr() + 4 * start * Cc, Cc, ntok * Cc, True)
    xin = tok[:, start:start + Hh * Ww].transpose(1, 2).reshape(B, Cc, Hh, Ww)
    ref = F.gelu(F.conv2d(xin, w, bb, padding=1, groups=Cc)).flatten(2).transpose(1, 2)
    assert float((out[:, start:start + Hh * Ww] - ref).abs().max()) < 1e-5
    assert torch.all(out[:, :start] == 0)
    gx = rnd(2, 16 * 16, 256, seed=51, scale=2.0)
    gw, gb = rnd(256, seed=52), rnd(256, seed=53)
    gn = ops.groupnorm(gx, 32, gw, gb, 1e-5, True)
    ref = F.relu(F.group_norm(gx.transpose(1, 2).reshape(2, 256, 16, 16), 32, gw, gb, 1e-5)).flatten(2).transpose(1, 2)
    assert float((gn - ref).abs().max()) < 2e-5


def test_depth_exp_and_gaussian_adapter():
    from siu3r_b200 import ops
    n = 5000
    xyz = rnd(n, 4, seed=54)
    pts = ops.depth_exp(xyz, n, 4)
    v = xyz[:, :3]
    d = v.norm(dim=-1, keepdim=True)
    ref = v / d.clip(min=1e-8) * d.expm1()
    assert float((pts - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    G = 1000 + 37
    raw = rnd(G, 83, seed=55, scale=2.0)
    cov, harm, opac, scales, rots = ops.gaussian_adapter(raw)
    o, s, r, sh = raw.split((1, 3, 4, 75), dim=-1)
    assert torch.allclose(opac, o.sigmoid().squeeze(-1), atol=1e-6)
    s_ref = (0.001 * F.softplus(s)).clamp_max(0.3)
    assert torch.allclose(scales, s_ref, rtol=1e-5, atol=1e-9)
    assert torch.equal(rots, r)
    mask = torch.ones(25, device=DEV)
    for dgr in range(1, 5):
        mask[dgr * dgr:(dgr + 1) ** 2] = 0.1 * 0.25 ** dgr
    assert torch.allclose(harm, sh.reshape(G, 3, 25) * mask, rtol=1e-6, atol=0)
    rn = r / (r.norm(dim=-1, keepdim=True) + 1e-8)
    i, j, k, w = rn.unbind(-1)
    two_s = 2 / ((rn * rn).sum(-1) + 1e-8)
    R = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * w), two_s * (i * k + j * w), two_s * (i * j + k * w),
                     1 - two_s * (i * i + k * k), two_s * (j * k - i * w), two_s * (i * k - j * w), two_s * (j * k + i * w),
                     1 - two_s * (i * i + j * j)), -1).view(G, 3, 3)
    S = s_ref.diag_embed()
    cref = R @ S @ S.transpose(1, 2) @ R.transpose(1, 2)
    assert float((cov - cref).abs().max()) < 1e-5 * float(cref.abs().max())


def test_postprocess_kernels():
    from siu3r_b200 import ops
    N, h, w, Q = 2, 16, 16, 100
    logits = rnd(N, h, w, Q, seed=56, scale=3.0)
    idx = torch.tensor([3, 50, 7, 99], dtype=torch.int32, device=DEV)
    probs = ops.eltwise(ops.ELT_SIGMOID, ops.resize_bilinear(logits, 32, 32, False))
    sel = ops.resize_select(probs, idx, 64, 64)
    ref = F.interpolate(probs.permute(0, 3, 1, 2)[:, idx.long()], size=(64, 64), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert float((sel - ref).abs().max()) < 1e-5
    score = torch.tensor([0.9, 0.6, 0.8, 0.7], device=DEV)
    labels, area, orig = ops.argmax_area(sel, score, 0.5)
    wv = sel * score
    lref = wv.argmax(-1).flatten()
    assert torch.equal(labels.long(), lref)
    assert torch.equal(area.long(), torch.bincount(lref, minlength=4))
    assert torch.equal(orig.long(), (wv >= 0.5).flatten(0, 2).sum(0))
    seg_lut = torch.tensor([1, 0, 2, 2], dtype=torch.int32, device=DEV)
    sem_lut = torch.tensor([5, 0, 1, 1], dtype=torch.int32, device=DEV)
    seg, sem, inst = ops.label_lut(labels, seg_lut, sem_lut)
    assert torch.equal(seg, seg_lut[lref]) and torch.equal(sem, sem_lut[lref]) and torch.equal(inst, seg)
    keep = torch.tensor([0, 2], dtype=torch.int32, device=DEV)
    cls = rnd(2, 21, seed=57).softmax(-1)
    qc = ops.qc_logits(sel, keep, cls)
    ref = sel.flatten(0, 2)[:, keep.long(), None] * cls[None]
    assert torch.allclose(qc, ref, atol=1e-7)
    # attention mask
    T = 2
    m = ops.attn_mask_from_logits(logits, 1, T, h, w, Q, 8, 8)
    lg = logits.permute(0, 3, 1, 2)  # [T, Q, h, w]
    am = F.interpolate(lg, size=(8, 8), mode="bilinear", align_corners=False).sigmoid() < 0.5  # [T,Q,8,8]
    ref = am.permute(1, 0, 2, 3).reshape(1, Q, T * 64)
    mism = (m.bool() != ref).float().mean()
    assert float(mism) < 1e-4  # (threshold ties at fp rounding only)


# ---------------------------------------------------------------------------------------------------------------
def _raster_case(G, H, W, seed, pixel_aligned):
    from siu3r_b200 import synth
    from siu3r_b200.renderer import camera_matrices
    sc = synth.raster_scene(G, H, W, seed=seed, pixel_aligned=pixel_aligned)
    view, full, campos, tx, ty = camera_matrices(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
    return sc, view[0], full[0], campos[0], float(tx[0]), float(ty[0])


@pytest.mark.parametrize("G,H,W,pa", [(3000, 64, 64, False), (20000, 128, 128, False), (50000, 256, 256, True), (5000, 100, 180, False), (1, 32, 32, False)])
def test_raster_vs_oracle(G, H, W, pa, oracle_lib):
    from siu3r_b200 import ops
    from oracle import raster_oracle as RO
    sc, view, full, campos, tx, ty = _raster_case(G, H, W, 3, pa)
    row, col = torch.triu_indices(3, 3)
    cov6 = sc["covariances"][:, row, col].contiguous()
    shs = sc["harmonics"].permute(0, 2, 1).contiguous()
    ref = RO.rasterize(sc["means"].numpy(), cov6.numpy(), shs.numpy(), sc["opacities"].numpy(), view.numpy(), full.numpy(), campos.numpy(), tx, ty,
                       H, W, 4)
    bg = torch.zeros(3, device=DEV)
    # product layouts: full 3x3 covariances + [G,3,25] harmonics (no gather / transpose passes)
    res = ops.raster_forward(sc["means"].to(DEV), sc["covariances"].to(DEV), sc["harmonics"].to(DEV), sc["opacities"].to(DEV), view.to(DEV),
                             full.to(DEV), campos.to(DEV), bg, tx, ty, H, W, 4, sh_layout=1, debug=True)
    D = ref["num_rendered"]
    assert res["num_rendered"] == D
    assert np.array_equal(res["radii"].cpu().numpy(), ref["radii"])
    assert np.array_equal(res["tiles"].cpu().numpy().astype(np.uint32), ref["tiles"])
    assert np.array_equal(res["offsets"].cpu().numpy().astype(np.uint32), ref["offsets"])
    assert np.array_equal(res["keys"][:D].cpu().numpy().astype(np.uint64), ref["keys"])
    assert np.array_equal(res["values"][:D].cpu().numpy().astype(np.uint32), ref["values"])
    assert np.array_equal(res["ranges"].cpu().numpy().astype(np.uint32), ref["ranges"])
    assert np.abs(res["color"].cpu().numpy() - ref["color"]).max() < 1e-3
    assert np.abs(res["opacity"].cpu().numpy() - ref["opacity"]).max() < 1e-3
    assert np.abs(res["depth"].cpu().numpy() - ref["depth"]).max() < 1e-3 * max(1.0, float(ref["depth"].max()))
    nt = res["n_touched"].cpu().numpy()
    # n_touched depends on a T > 0.5 threshold evaluated with a different exp(): allow isolated off-by-ones
    assert np.abs(nt - ref["n_touched"]).max() <= 2 and (nt != ref["n_touched"]).mean() < 1e-3
    # the reference layouts (cov6 + [G,M,3] SH) must give the identical result
    res2 = ops.raster_forward(sc["means"].to(DEV), cov6.to(DEV), shs.to(DEV), sc["opacities"].to(DEV), view.to(DEV), full.to(DEV), campos.to(DEV), bg,
                              tx, ty, H, W, 4, sh_layout=0, count_touched=False)
    assert torch.equal(res2["color"], res["color"]) and torch.equal(res2["radii"], res["radii"])


def test_raster_properties_large():
    """BASELINE config-5 size: properties that do not need the (slow) oracle."""
    from siu3r_b200 import ops
    G, H, W = 500000, 512, 512
    sc, view, full, campos, tx, ty = _raster_case(G, H, W, 5, True)
    bg = torch.zeros(3, device=DEV)
    res = ops.raster_forward(sc["means"].to(DEV), sc["covariances"].to(DEV), sc["harmonics"].to(DEV), sc["opacities"].to(DEV), view.to(DEV),
                             full.to(DEV), campos.to(DEV), bg, tx, ty, H, W, 4, sh_layout=1, debug=True)
    D = res["num_rendered"]
    keys = res["keys"][:D]
    assert bool((keys[1:] >= keys[:-1]).all())                       # sortedness
    assert int(res["tiles"].long().sum()) == D                       # scan total
    assert int(res["offsets"][-1]) == D
    rg = res["ranges"].long()
    assert int((rg[:, 1] - rg[:, 0]).sum()) == D                     # ranges partition the list
    op = res["opacity"]
    assert float(op.min()) >= 0.0 and float(op.max()) <= 1.0
    assert torch.isfinite(res["color"]).all()
    # idempotence
    res2 = ops.raster_forward(sc["means"].to(DEV), sc["covariances"].to(DEV), sc["harmonics"].to(DEV), sc["opacities"].to(DEV), view.to(DEV),
                              full.to(DEV), campos.to(DEV), bg, tx, ty, H, W, 4, sh_layout=1)
    assert torch.equal(res2["color"], res["color"]) and torch.equal(res2["depth"], res["depth"])
    # exact sub-tile culling: bit-identical to the unculled per-pixel loop (colour, depth, opacity, n_touched)
    from siu3r_b200 import _lib
    lib = _lib.load()
    for pa in (True, False):
        sc, view, full, campos, tx, ty = _raster_case(G // 4 if not pa else G, H, W, 11, pa)
        sc["opacities"][::7] *= 1e-3        # some records below the 1/255 floor
        args = (sc["means"].to(DEV), sc["covariances"].to(DEV), sc["harmonics"].to(DEV), sc["opacities"].to(DEV), view.to(DEV), full.to(DEV),
                campos.to(DEV), bg, tx, ty, H, W, 4)
        a = ops.raster_forward(*args, sh_layout=1)
        lib.siu3r_raster_set_culling(0)
        try:
            b = ops.raster_forward(*args, sh_layout=1)
        finally:
            lib.siu3r_raster_set_culling(1)
        for kk in ("color", "depth", "opacity", "n_touched"):
            assert torch.equal(a[kk], b[kk]), (pa, kk)


# ---------------------------------------------------------------------------------------------------------------
# N-channel feature splatting (gsplat.rasterization semantics; oracle/gsplat_ref.py, parity unpinned vs gsplat itself)
def _feature_scene(G, H, W, C, seed):
    rng = np.random.default_rng(seed)
    z = rng.uniform(1.5, 30.0, G)
    means = np.stack([rng.uniform(-1, 1, G) * z * 0.5, rng.uniform(-1, 1, G) * z * 0.5, z], -1).astype("f4")
    A = (rng.standard_normal((G, 3, 3)) * (0.02 * z)[:, None, None]).astype("f4")
    cov = (A @ A.transpose(0, 2, 1) + 1e-6 * np.eye(3, dtype="f4")).astype("f4")
    op = rng.random(G).astype("f4")
    op[::11] = 0.001                                  # below the 1/255 floor
    means[::13, 2] = 0.5                              # in front of the near plane (1.0)
    feats = rng.standard_normal((G, C)).astype("f4")
    ang = 0.2
    V = np.eye(4, dtype="f4")
    V[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], dtype="f4")
    V[:3, 3] = [0.1, -0.05, 0.3]
    return means, cov, op, feats, V


@pytest.mark.parametrize("G,H,W,C", [(3000, 64, 64, 42), (8000, 96, 160, 21), (500, 48, 80, 70), (1, 32, 32, 5)])
def test_raster_features_vs_oracle(G, H, W, C):
    from siu3r_b200 import ops
    from oracle import gsplat_ref as GR
    means, cov, op, feats, V = _feature_scene(G, H, W, C, 5)
    fx, fy, cx, cy = 1.242 * W, 1.242 * H, 0.5 * W, 0.47 * H
    ref, ref_alpha, pr = GR.rasterize(means, cov, op, feats, V, fx, fy, cx, cy, W, H, 1.0, 1000.0)
    t = lambda a: torch.from_numpy(a).to(DEV)
    res = ops.raster_features_forward(t(means), t(cov), t(op), t(feats), t(V), (fx, fy, cx, cy), 1.0, 1000.0, H, W, want_radii=True)
    radii = res["radii"].cpu().numpy()
    # extents go through log(): CUDA logf and numpy log may differ in the last bit -> allow isolated off-by-one radii
    assert (np.abs(radii - pr["radii"]).max() <= 1) and (radii != pr["radii"]).mean() < 2e-3
    scale = max(1.0, float(np.abs(ref).max()))
    assert np.abs(res["features"].cpu().numpy() - ref).max() < 1e-3 * scale
    assert np.abs(res["alpha"].cpu().numpy() - ref_alpha).max() < 1e-3
    # cov6 layout gives the identical result; so does a second call (idempotence)
    row, col = np.triu_indices(3)
    res2 = ops.raster_features_forward(t(means), t(np.ascontiguousarray(cov[:, row, col])), t(op), t(feats), t(V), (fx, fy, cx, cy), 1.0, 1000.0, H, W)
    assert torch.equal(res2["features"], res["features"])
    # linearity in the features (the blend weights do not depend on them)
    res3 = ops.raster_features_forward(t(means), t(cov), t(op), t(2 * feats), t(V), (fx, fy, cx, cy), 1.0, 1000.0, H, W)
    assert float((res3["features"] - 2 * res["features"]).abs().max()) < 1e-4 * scale
    # the no-host-sync binned path renders the identical frame and reports the duplicate count on the device
    status = torch.zeros(4, device=DEV, dtype=torch.int32)
    alpha4 = torch.empty(H, W, device=DEV)
    res4 = ops.raster_features_forward_nosync(t(means), t(cov), t(op), t(feats), t(V), (fx, fy, cx, cy), 1.0, 1000.0, H, W, status, alpha=alpha4)
    st = status.cpu()
    assert int(st[2]) == 0 and int(st[0]) == res["num_rendered"]
    assert torch.equal(res4["features"], res["features"]) and torch.equal(alpha4, res["alpha"])
    if res["num_rendered"] > 8:   # capacity overflow: flagged, nothing rendered
        status.zero_()
        out5 = torch.full((H, W, C), 7.0, device=DEV)
        ops.raster_features_forward_nosync(t(means), t(cov), t(op), t(feats), t(V), (fx, fy, cx, cy), 1.0, 1000.0, H, W, status, dup_capacity=4, out=out5)
        st = status.cpu()
        assert int(st[2]) & 1 and int(st[0]) == res["num_rendered"] and torch.all(out5 == 7.0)


def test_splatting_render_qc_logits_surface():
    """SplattingCUDA.forward(render_qc_logits=True): list over the batch of [V, q, c, H, W] (gaussian_renderer.py:75-110), consistent with a
    direct call of the feature rasterizer, colour path unchanged."""
    from siu3r_b200 import ops
    from siu3r_b200.gaussians import Gaussians
    from siu3r_b200.renderer import SplattingCUDA
    means, cov, op, feats, V = _feature_scene(4000, 64, 64, 42, 9)
    t = lambda a: torch.from_numpy(a).to(DEV)
    g = Gaussians(means=t(means)[None] / 10, covariances=t(cov)[None] / 100, harmonics=torch.randn(1, 4000, 3, 25, device=DEV) * 0.1, opacities=t(op)[None])
    g.seg_query_class_logits = [t(feats).view(4000, 2, 21)]
    E = torch.eye(4)[None, None].repeat(1, 2, 1, 1)
    E[0, 1, 0, 3] = 0.02
    K = torch.tensor([[1.242, 0, 0.5], [0, 1.242, 0.5], [0, 0, 1.0]])[None, None].repeat(1, 2, 1, 1)
    out = SplattingCUDA()(g, E, K, (64, 64), render_color=True, render_qc_logits=True)
    qc = out["render_qc_logits"]
    assert len(qc) == 1 and tuple(qc[0].shape) == (2, 2, 21, 64, 64) and tuple(out["render_color"].shape) == (1, 2, 3, 64, 64)
    # the renderer rescaled the scene x10 in place (gaussian_renderer.py:43-46); view 0 has the identity pose
    direct = ops.raster_features_forward(g.means[0].contiguous(), g.covariances[0].contiguous(), g.opacities[0].contiguous(), t(feats),
                                         torch.eye(4, device=DEV), (1.242 * 64, 1.242 * 64, 32.0, 32.0), 1.0, 1000.0, 64, 64)
    assert torch.equal(qc[0][0], direct["features"].view(64, 64, 2, 21).permute(2, 3, 0, 1))
    assert torch.isfinite(qc[0]).all() and float(qc[0].abs().max()) > 0


@pytest.mark.parametrize("G,H,W,pa", [(3000, 64, 64, False), (50000, 256, 256, True), (500000, 512, 512, True), (5000, 100, 180, False),
                                      (200000, 1080, 1920, True), (1, 32, 32, False), (60000, 48, 48, False)])
def test_raster_binned_sort_equals_global_sort(G, H, W, pa):
    """The per-tile binned sort (tile counts -> scatter -> shared-memory bitonic sort of (depth bits, id)) must reproduce the sorted list of the
    reference-shaped pipeline (global stable radix sort of (tile << 32 | depth)): identical images, depth, opacity, radii and n_touched.
    The last case puts ~26 000 records per tile (above the binned path's tile limit) and exercises the fallback to the global sort."""
    from siu3r_b200 import _lib, ops
    lib = _lib.load()
    sc, view, full, campos, tx, ty = _raster_case(G, H, W, 21, pa)
    bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    args = (sc["means"].to(DEV), sc["covariances"].to(DEV), sc["harmonics"].to(DEV), sc["opacities"].to(DEV), view.to(DEV), full.to(DEV),
            campos.to(DEV), bg, tx, ty, H, W, 4)
    a = ops.raster_forward(*args, sh_layout=1)
    lib.siu3r_raster_set_binning(0)
    try:
        b = ops.raster_forward(*args, sh_layout=1)
    finally:
        lib.siu3r_raster_set_binning(1)
    assert a["num_rendered"] == b["num_rendered"]
    for kk in ("color", "depth", "opacity", "radii", "n_touched"):
        assert torch.equal(a[kk], b[kk]), kk


@pytest.mark.parametrize("H,W", [(64, 64), (96, 160), (512, 512)])
def test_conv_rows_up2x_fused_matches_unfused(H, W):
    """input_merger of the GS head fused (7x7 image conv + ReLU + bilinear x2 of the trunk output in one launch) against the unfused
    resize_bilinear + conv2d path and against torch (fp32, TF32-rounded operands)."""
    import torch.nn.functional as F
    from siu3r_b200 import ops
    torch.manual_seed(H)
    N, Cout = 1, 256
    x = torch.zeros(N, H, W, 4, device=DEV)
    x[..., :3] = torch.rand(N, H, W, 3, device=DEV)
    low = torch.randn(N, H // 2, W // 2, Cout, device=DEV)
    w = torch.randn(Cout, 7, 7, 4, device=DEV) / 12
    w[..., 3] = 0
    b = torch.randn(Cout, device=DEV) * 0.1
    wt = ops.Weight(w.reshape(Cout, -1), b, ops.PREC_TF32)
    fused = ops.conv_kxk_up2x(x, wt, 7, 7, low, act=ops.ACT_RELU, round_out=False)
    assert fused is not None
    up = ops.resize_bilinear(low, H, W, True)
    unfused = ops.conv2d(x, wt, 7, 7, pad=3, act=ops.ACT_RELU, residual=up)
    assert float((fused - unfused).abs().max()) < 2e-4 * max(1.0, float(unfused.abs().max()))
    xr, wr = ops.round_tf32(x.contiguous()), wt.w[:, :196].reshape(Cout, 7, 7, 4)
    ref = F.relu(F.conv2d(xr.permute(0, 3, 1, 2), wr.permute(0, 3, 1, 2), b, padding=3)) + F.interpolate(low.permute(0, 3, 1, 2), size=(H, W),
                                                                                                        mode="bilinear", align_corners=True)
    assert float((fused - ref.permute(0, 2, 3, 1)).abs().max()) < 2e-3 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("M0,M1,N,K,col0", [(1025, 1025, 2304, 768, 1536), (257, 257, 2304, 768, 1536), (1025, 1025, 1536, 768, 768), (2050, 0, 3072, 1024, 2048)])
def test_gemm_vt_emission(M0, M1, N, K, col0):
    """V columns of a qkv / k|v projection written as V^T by the persistent kernel's epilogue (single and grouped launch): exact against the
    row-major result, pad columns zero, C untouched for those columns, tile-edge fragments do not clobber their neighbours."""
    from siu3r_b200 import ops
    torch.manual_seed(M0 + N)
    Ms = [M0] + ([M1] if M1 else [])
    xs = [ops.round_tf32(torch.randn(m, K, device=DEV)) for m in Ms]
    ws = [ops.Weight(torch.randn(N, K, device=DEV) / K ** 0.5, torch.randn(N, device=DEV), 1) for _ in Ms]
    outs = [torch.zeros(m, N, device=DEV) for m in Ms]
    stride = (M0 + 3) // 4 * 4
    buf = torch.full((N - col0, len(Ms) * stride), float("nan"), device=DEV)
    st = {}
    if M1:
        ops.gemm_group2(xs, ws, outs=outs, a_rounded=True, round_out=True, vt=([buf[:, :stride], buf[:, stride:]], [stride, stride], col0, st))
    else:
        ops.gemm(xs[0], ws[0], out=outs[0], a_rounded=True, round_out=True, vt=(buf, col0, st))
    assert st["ok"]
    for g, m in enumerate(Ms):
        ref = (xs[g].double() @ ws[g].w.double().t() + ws[g].bias.double()).float()
        assert float((outs[g][:, :col0] - ref[:, :col0]).abs().max()) < 4e-3
        w = buf[:, g * stride: g * stride + m]
        assert float((w - ref[:, col0:].t()).abs().max()) < 4e-3
        assert float(buf[:, g * stride + m: (g + 1) * stride].abs().sum()) == 0.0
        assert float(outs[g][:, col0:].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K", [(100, 256, 256), (100, 2048, 256), (100, 256, 2048), (100, 21, 256), (7, 83, 100), (128, 512, 260)])
def test_gemm_skinny(M, N, K):
    from siu3r_b200 import _lib, ops
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    b = torch.randn(N, device=DEV)
    r = torch.randn(M, N, device=DEV)
    lib = _lib.load()
    for act, res in ((0, None), (1, r), (2, None), (2 | 4, r)):
        out = torch.empty(M, N, device=DEV)
        code = lib.siu3r_gemm_skinny(M, N, K, x.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, b.data_ptr(), None if res is None else res.data_ptr(),
                                     N, act, 1.0, ops._stream())
        assert code == 0
        ref = (x.double() @ w.double().t() + b.double()).float()
        if act & 3 == 1:
            ref = torch.nn.functional.gelu(ref)
        elif act & 3 == 2:
            ref = torch.relu(ref)
        if res is not None:
            ref = ref + res
        tol = 3e-3 if act & 4 else 2e-5
        assert float((out - ref).abs().max()) < tol * max(1.0, float(ref.abs().max())), (act, float((out - ref).abs().max()))


@pytest.mark.parametrize("G,H,W,pa", [(3000, 64, 64, False), (500000, 512, 512, True), (5000, 100, 180, False), (12000, 48, 48, False),
                                      (2000000, 512, 512, True), (1, 32, 32, False)])
def test_raster_nosync_equals_sync_and_global_sort(G, H, W, pa):
    """siu3r_raster_forward_nosync (no host round trip, register-resident tile sorts of all three size classes: (12000, 48, 48) puts ~5000 records
    in a tile, the 2 M case ~4300) gives bit-identical images to the synchronising entry point with the reference-shaped global radix sort."""
    from siu3r_b200 import _lib, ops
    lib = _lib.load()
    sc, view, full, campos, tx, ty = _raster_case(G, H, W, 23, pa)
    bg = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    args = (sc["means"].to(DEV), sc["covariances"].to(DEV), sc["harmonics"].to(DEV), sc["opacities"].to(DEV), view.to(DEV), full.to(DEV),
            campos.to(DEV), bg, tx, ty, H, W, 4)
    lib.siu3r_raster_set_binning(0)
    try:
        ref = ops.raster_forward(*args, sh_layout=1)
    finally:
        lib.siu3r_raster_set_binning(1)
    status = torch.zeros(4, device=DEV, dtype=torch.int32)
    got = ops.raster_forward_nosync(*args, sh_layout=1, status=status, count_touched=True, dup_capacity=max(1 << 16, int(ref["num_rendered"] * 1.1)))
    st = status.cpu().tolist()
    assert st[0] == ref["num_rendered"] and st[2] == 0, st
    for kk in ("color", "depth", "opacity", "radii", "n_touched"):
        assert torch.equal(got[kk], ref[kk]), kk
    # the round-1 shared-memory sort and the register-resident one order every tile identically
    a = ops.raster_forward(*args, sh_layout=1)
    lib.siu3r_raster_set_regsort(0)
    try:
        b = ops.raster_forward(*args, sh_layout=1)
    finally:
        lib.siu3r_raster_set_regsort(1)
    for kk in ("color", "depth", "n_touched"):
        assert torch.equal(a[kk], b[kk]) and torch.equal(a[kk], ref[kk]), kk


def test_raster_nosync_flags_overflow_and_render_cuda_recovers():
    """Too small a duplicate capacity, or a tile above 8192 records: status flag raised, nothing written out of bounds; render_cuda re-renders."""
    from siu3r_b200 import ops
    from siu3r_b200.renderer import render_cuda
    G, H, W = 60000, 48, 48          # ~26 000 records per tile
    sc, view, full, campos, tx, ty = _raster_case(G, H, W, 29, False)
    bg = torch.zeros(3, device=DEV)
    args = (sc["means"].to(DEV), sc["covariances"].to(DEV), sc["harmonics"].to(DEV), sc["opacities"].to(DEV), view.to(DEV), full.to(DEV),
            campos.to(DEV), bg, tx, ty, H, W, 4)
    ref = ops.raster_forward(*args, sh_layout=1, count_touched=False)
    status = torch.zeros(4, device=DEV, dtype=torch.int32)
    ops.raster_forward_nosync(*args, sh_layout=1, status=status, dup_capacity=1 << 16)
    st = status.cpu().tolist()
    assert st[0] == ref["num_rendered"] and (st[2] & 1) == (1 if ref["num_rendered"] > (1 << 16) else 0) and (st[2] & 2) == 2, st
    color, depth = render_cuda(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"], (H, W), torch.zeros(1, 3), sc["means"][None].to(DEV),
                               sc["covariances"][None].to(DEV), sc["harmonics"][None].to(DEV), sc["opacities"][None].to(DEV))
    assert torch.equal(color[0], ref["color"]) and torch.equal(depth[0], ref["depth"])
