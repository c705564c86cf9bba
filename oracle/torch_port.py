"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path (siu3r_b200/).

Plain-PyTorch (CPU, fp32, functional) restatement of SIU3RModel.forward for the two-view path, consuming the reference
state_dict unchanged.  It exists because the reference itself is Python under /root/reference and cannot travel to the
GPU box: this port does, and is
  * PINNED against golden vectors generated from the unmodified reference (tests/golden/model_S*.npz, made by
    oracle/make_golden.py) in tests/test_oracle_cpu.py, and
  * the `cpu_baseline` / `--impl reference` arm of bench.py (kind = "port").

Each function cites the reference code it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _ln(x, sd, p, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


# ---- RoPE2D, pure PyTorch form (croco/pos_embed.py:126-179) -------------------------------------------------------
def rope2d(tokens, positions, base=100.0):
    D = tokens.shape[-1] // 2
    inv_freq = 1.0 / (base ** (torch.arange(0, D, 2).float() / D))
    t = torch.arange(int(positions.max()) + 1, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    freqs = torch.cat((freqs, freqs), dim=-1)
    cos, sin = freqs.cos(), freqs.sin()

    def rot_half(x):
        x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
        return torch.cat((-x2, x1), dim=-1)

    def rope1d(tok, pos1d):
        c = F.embedding(pos1d, cos)[:, None]
        s = F.embedding(pos1d, sin)[:, None]
        return tok * c + rot_half(tok) * s

    y, x = tokens.chunk(2, dim=-1)
    return torch.cat((rope1d(y, positions[:, :, 0]), rope1d(x, positions[:, :, 1])), dim=-1)


# ---- CroCo blocks (croco/blocks.py:94-112,127-130,149-169,186-191) ---------------------------------------------------
def _attention(x, pos, sd, p, nh):
    B, N, C = x.shape
    qkv = _lin(x, sd, p + ".qkv").reshape(B, N, 3, nh, C // nh).transpose(1, 3)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    q, k = rope2d(q, pos), rope2d(k, pos)
    a = ((q @ k.transpose(-2, -1)) * (C // nh) ** -0.5).softmax(dim=-1)
    return _lin((a @ v).transpose(1, 2).reshape(B, N, C), sd, p + ".proj")


def _cross_attention(q_in, kv_in, qpos, kpos, sd, p, nh):
    B, Nq, C = q_in.shape
    q = _lin(q_in, sd, p + ".projq").reshape(B, Nq, nh, C // nh).permute(0, 2, 1, 3)
    k = _lin(kv_in, sd, p + ".projk").reshape(B, -1, nh, C // nh).permute(0, 2, 1, 3)
    v = _lin(kv_in, sd, p + ".projv").reshape(B, -1, nh, C // nh).permute(0, 2, 1, 3)
    q, k = rope2d(q, qpos), rope2d(k, kpos)
    a = ((q @ k.transpose(-2, -1)) * (C // nh) ** -0.5).softmax(dim=-1)
    return _lin((a @ v).transpose(1, 2).reshape(B, Nq, C), sd, p + ".proj")


def _mlp(x, sd, p):
    return _lin(F.gelu(_lin(x, sd, p + ".fc1")), sd, p + ".fc2")


def _enc_block(x, pos, sd, p):
    x = x + _attention(_ln(x, sd, p + ".norm1", 1e-6), pos, sd, p + ".attn", 16)
    return x + _mlp(_ln(x, sd, p + ".norm2", 1e-6), sd, p + ".mlp")


def _dec_block(x, y, xpos, ypos, sd, p):
    x = x + _attention(_ln(x, sd, p + ".norm1", 1e-6), xpos, sd, p + ".attn", 12)
    y_ = _ln(y, sd, p + ".norm_y", 1e-6)
    x = x + _cross_attention(_ln(x, sd, p + ".norm2", 1e-6), y_, xpos, ypos, sd, p + ".cross_attn", 12)
    return x + _mlp(_ln(x, sd, p + ".norm3", 1e-6), sd, p + ".mlp")


def backbone(sd, images, intrinsics):
    """AsymmetricCroCo.forward (backbone_croco.py:263-339) -> feat-lists with the intrinsics token still attached."""
    B, V, _, H, W = images.shape
    p = "backbone."
    img = torch.cat((images[:, 0], images[:, 1]), 0)
    x = F.conv2d(img, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=16)
    gh, gw = x.shape[2], x.shape[3]
    x = x.flatten(2).transpose(1, 2)
    emb = F.linear(intrinsics.flatten(2), sd[p + "intrinsic_encoder.weight"], sd[p + "intrinsic_encoder.bias"])  # [B,2,1024]
    x = torch.cat((x, torch.cat((emb[:, 0], emb[:, 1]), 0)[:, None]), dim=1)
    pos = torch.cartesian_prod(torch.arange(gh), torch.arange(gw))
    pos = torch.cat((pos, torch.tensor([[gh, 0]])), 0)[None].expand(2 * B, -1, -1)
    all_feat = []
    for i in range(24):
        x = _enc_block(x, pos, sd, p + f"enc_blocks.{i}")
        all_feat.append(x)
    feat = _ln(x, sd, p + "enc_norm", 1e-6)
    f1, f2 = feat[:B], feat[B:]
    pos1 = pos[:B]
    dec1, dec2 = [f1], [f2]
    f1, f2 = _lin(f1, sd, p + "decoder_embed"), _lin(f2, sd, p + "decoder_embed")
    for i in range(12):
        n1 = _dec_block(f1, f2, pos1, pos1, sd, p + f"dec_blocks.{i}")
        n2 = _dec_block(f2, f1, pos1, pos1, sd, p + f"dec_blocks2.{i}")
        f1, f2 = n1, n2
        dec1.append(f1)
        dec2.append(f2)
    dec1[-1] = _ln(dec1[-1], sd, p + "dec_norm", 1e-6)
    dec2[-1] = _ln(dec2[-1], sd, p + "dec_norm", 1e-6)
    strip = lambda t: t[:, :-1]
    return dict(all_feat=[strip(t) for t in all_feat], dec1=[strip(t) for t in dec1], dec2=[strip(t) for t in dec2], gh=gh, gw=gw)


def backbone_multi(sd, images, intrinsics):
    """AsymmetricCroCoMulti.forward (backbone_croco.py:537-590) + _decoder (:487-535): V >= 2 views.  View 0 runs dec_blocks,
    views 1..V-1 run dec_blocks2; the cross-attention memory of view i is the concatenation (in view order) of the other
    V-1 views' previous-layer tokens, intrinsics tokens included, with their own positions (generate_ctx_views :500-506)."""
    B, V, _, H, W = images.shape
    p = "backbone."
    img = images.flatten(0, 1)                                    # (b v) ordering, :552
    x = F.conv2d(img, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=16)
    gh, gw = x.shape[2], x.shape[3]
    x = x.flatten(2).transpose(1, 2)
    emb = F.linear(intrinsics.flatten(2), sd[p + "intrinsic_encoder.weight"], sd[p + "intrinsic_encoder.bias"])  # [B,V,1024]
    x = torch.cat((x, emb.flatten(0, 1)[:, None]), dim=1)
    pos = torch.cartesian_prod(torch.arange(gh), torch.arange(gw))
    pos = torch.cat((pos, torch.tensor([[gh, 0]])), 0)[None].expand(B * V, -1, -1)
    all_feat = []
    for i in range(24):
        x = _enc_block(x, pos, sd, p + f"enc_blocks.{i}")
        all_feat.append(x)
    N = x.shape[1]
    feat = _ln(x, sd, p + "enc_norm", 1e-6).view(B, V, N, -1)
    pose = pos.reshape(B, V, N, 2)

    def ctx(t):  # [B,V,L,C] -> [B,V,(V-1)L,C]: for view i, all views j != i in order
        return torch.stack([torch.cat([t[:, j] for j in range(V) if j != i], dim=1) for i in range(V)], dim=1)

    pos_ctx = ctx(pose)
    outs = [feat]
    f = _lin(feat, sd, p + "decoder_embed")
    for i in range(12):
        c = ctx(f)
        n1 = _dec_block(f[:, 0], c[:, 0], pose[:, 0], pos_ctx[:, 0], sd, p + f"dec_blocks.{i}")
        n2 = _dec_block(f[:, 1:].flatten(0, 1), c[:, 1:].flatten(0, 1), pose[:, 1:].flatten(0, 1), pos_ctx[:, 1:].flatten(0, 1), sd,
                        p + f"dec_blocks2.{i}")
        f = torch.cat((n1[:, None], n2.view(B, V - 1, N, -1)), dim=1)
        outs.append(f)
    outs[-1] = _ln(outs[-1], sd, p + "dec_norm", 1e-6)
    all_feat = [t.view(B, V, N, -1)[:, :, :-1] for t in all_feat]
    return dict(all_feat=all_feat, dec=[t[:, :, :-1] for t in outs], gh=gh, gw=gw)


# ---- DPT heads (heads/dpt_block.py, dpt_head.py:36-79, dpt_gs_head.py:121-171, postprocess.py:46-61) --------------------
def _conv(x, sd, p, **kw):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), **kw)


def _rcu(x, sd, p):
    out = _conv(F.relu(x), sd, p + ".conv1", padding=1)
    out = _conv(F.relu(out), sd, p + ".conv2", padding=1)
    return out + x


def _fusion(sd, p, x0, x1=None):
    out = x0
    if x1 is not None:
        out = out + _rcu(x1, sd, p + ".resConfUnit1")
    out = _rcu(out, sd, p + ".resConfUnit2")
    out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    return _conv(out, sd, p + ".out_conv")


def _dpt_trunk(sd, p, dec, gh, gw):
    hooks = [0, 6, 9, 12]
    layers = [dec[h].transpose(1, 2).reshape(dec[h].shape[0], -1, gh, gw) for h in hooks]
    a = p + "act_postprocess."
    l0 = F.conv_transpose2d(_conv(layers[0], sd, a + "0.0"), sd[a + "0.1.weight"], sd[a + "0.1.bias"], stride=4)
    l1 = F.conv_transpose2d(_conv(layers[1], sd, a + "1.0"), sd[a + "1.1.weight"], sd[a + "1.1.bias"], stride=2)
    l2 = _conv(layers[2], sd, a + "2.0")
    l3 = _conv(_conv(layers[3], sd, a + "3.0"), sd, a + "3.1", stride=2, padding=1)
    ls = [_conv(l, sd, p + f"scratch.layer_rn.{i}", padding=1) for i, l in enumerate((l0, l1, l2, l3))]
    p4 = _fusion(sd, p + "scratch.refinenet4", ls[3])
    p3 = _fusion(sd, p + "scratch.refinenet3", p4, ls[2])
    p2 = _fusion(sd, p + "scratch.refinenet2", p3, ls[1])
    return _fusion(sd, p + "scratch.refinenet1", p2, ls[0])


def center_head(sd, name, dec, gh, gw):
    p = name + ".dpt."
    x = _conv(_dpt_trunk(sd, p, dec, gh, gw), sd, p + "head.0", padding=1)
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = _conv(F.relu(_conv(x, sd, p + "head.2", padding=1)), sd, p + "head.4")
    xyz = x.permute(0, 2, 3, 1)
    d = xyz.norm(dim=-1, keepdim=True)
    return xyz / d.clip(min=1e-8) * d.expm1()  # [B,H,W,3]


def gs_head(sd, name, dec, img, gh, gw):
    p = name + ".dpt."
    p1 = F.interpolate(_dpt_trunk(sd, p, dec, gh, gw), scale_factor=2, mode="bilinear", align_corners=True)
    p1 = p1 + F.relu(_conv(img, sd, p + "input_merger.0", padding=3))
    out = _conv(F.relu(_conv(p1, sd, p + "head.0", padding=1)), sd, p + "head.4")
    return out  # [B,83,H,W]


def gaussian_adapter(means, raw):
    """UnifiedGaussianAdapter.forward (gaussian_adapter.py:81-110); raw [..., 83]."""
    o, s, r, sh = raw.split((1, 3, 4, 75), dim=-1)
    opac = o.sigmoid().squeeze(-1)
    scales = (0.001 * F.softplus(s)).clamp_max(0.3)
    rn = r / (r.norm(dim=-1, keepdim=True) + 1e-8)
    mask = torch.ones(25)
    for dg in range(1, 5):
        mask[dg * dg:(dg + 1) ** 2] = 0.1 * 0.25 ** dg
    harm = sh.reshape(*sh.shape[:-1], 3, 25) * mask
    i, j, k, w = rn.unbind(-1)
    two_s = 2 / ((rn * rn).sum(-1) + 1e-8)
    R = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * w), two_s * (i * k + j * w), two_s * (i * j + k * w),
                     1 - two_s * (i * i + k * k), two_s * (j * k - i * w), two_s * (i * k - j * w), two_s * (j * k + i * w),
                     1 - two_s * (i * i + j * j)), -1).reshape(*rn.shape[:-1], 3, 3)
    S = scales.diag_embed()
    cov = R @ S @ S.transpose(-1, -2) @ R.transpose(-1, -2)
    return dict(means=means, covariances=cov, harmonics=harm, opacities=opac, scales=scales, rotations=r)


# ---- deformable attention (vit_adapter/blocks.py:171-267) ---------------------------------------------------------------
def _msda(value, shapes, loc, attw):
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


def _ref_points(shapes):
    out = []
    for h, w in shapes:
        ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h), torch.linspace(0.5, w - 0.5, w), indexing="ij")
        out.append(torch.stack((rx.reshape(-1) / w, ry.reshape(-1) / h), -1))
    return torch.cat(out, 0)


def _deform_attn(sd, p, query, ref, feat, shapes, nH, P):
    B, Lq, C = query.shape
    L = len(shapes)
    value = _lin(feat, sd, p + ".value_proj").view(B, -1, nH, C // nH)
    offs = _lin(query, sd, p + ".sampling_offsets").view(B, Lq, nH, L, P, 2)
    aw = _lin(query, sd, p + ".attention_weights").view(B, Lq, nH, L * P).softmax(-1).view(B, Lq, nH, L, P)
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
    loc = ref[None, :, None, None, None, :] + offs / norm[None, None, None, :, None, :]
    return _lin(_msda(value, shapes, loc, aw), sd, p + ".output_proj")


# ---- ViT adapter (vit_adapter/vit_adapter.py:200-441) ---------------------------------------------------------------------
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def adapter(sd, img, all_feat, gh, gw):
    p = "adapter."
    x = img
    for i in (0, 3, 6):
        x = F.relu(_bn(F.conv2d(x, sd[p + f"spm.stem.{i}.weight"], None, stride=2 if i == 0 else 1, padding=1), sd, p + f"spm.stem.{i + 1}"))
    c1 = F.max_pool2d(x, 3, 2, 1)
    c2 = F.relu(_bn(F.conv2d(c1, sd[p + "spm.conv2.0.weight"], None, stride=2, padding=1), sd, p + "spm.conv2.1"))
    c3 = F.relu(_bn(F.conv2d(c2, sd[p + "spm.conv3.0.weight"], None, stride=2, padding=1), sd, p + "spm.conv3.1"))
    c4 = F.relu(_bn(F.conv2d(c3, sd[p + "spm.conv4.0.weight"], None, stride=2, padding=1), sd, p + "spm.conv4.1"))
    c1 = _conv(c1, sd, p + "spm.fc1")
    B, dim = c1.shape[:2]
    lvl = sd[p + "level_embed"]
    toks = [_conv(c, sd, p + f"spm.fc{i + 2}").view(B, dim, -1).transpose(1, 2) + lvl[i] for i, c in enumerate((c2, c3, c4))]
    n2, n3 = toks[0].shape[1], toks[1].shape[1]
    c = torch.cat(toks, dim=1)
    ref = _ref_points([(2 * gh, 2 * gw), (gh, gw), (gh // 2, gw // 2)])

    def extractor(q, feat, e):
        attn = _deform_attn(sd, e + ".attn", _ln(q, sd, e + ".query_norm", 1e-6), ref, _ln(feat, sd, e + ".feat_norm", 1e-6), [(gh, gw)], 16, 4)
        q = q + attn
        t = _lin(_ln(q, sd, e + ".ffn_norm", 1e-6), sd, e + ".ffn.fc1")
        n = t.shape[1] // 21
        Cc = t.shape[2]
        parts = []
        for (a0, a1, hh, ww) in ((0, 16 * n, 2 * gh, 2 * gw), (16 * n, 20 * n, gh, gw), (20 * n, t.shape[1], gh // 2, gw // 2)):
            m = t[:, a0:a1].transpose(1, 2).reshape(B, Cc, hh, ww)
            m = F.conv2d(m, sd[e + ".ffn.dwconv.dwconv.weight"], sd[e + ".ffn.dwconv.dwconv.bias"], padding=1, groups=Cc)
            parts.append(m.flatten(2).transpose(1, 2))
        return q + _lin(F.gelu(torch.cat(parts, 1)), sd, e + ".ffn.fc2")

    outs = []
    for i, idx in enumerate((5, 11, 17, 23)):
        xf = all_feat[idx]
        c = extractor(c, xf, p + f"interactions.{i}.extractor")
        if i == 3:
            for j in range(2):
                c = extractor(c, xf, p + f"interactions.{i}.extra_extractors.{j}")
        outs.append(xf.transpose(1, 2).reshape(B, dim, gh, gw))
    c2 = c[:, :n2].transpose(1, 2).reshape(B, dim, 2 * gh, 2 * gw)
    c3 = c[:, n2:n2 + n3].transpose(1, 2).reshape(B, dim, gh, gw)
    c4 = c[:, n2 + n3:].transpose(1, 2).reshape(B, dim, gh // 2, gw // 2)
    c1 = F.conv_transpose2d(c2, sd[p + "up.weight"], sd[p + "up.bias"], stride=2) + c1
    c1 = c1 + F.interpolate(outs[0], scale_factor=4, mode="bilinear", align_corners=False)
    c2 = c2 + F.interpolate(outs[1], scale_factor=2, mode="bilinear", align_corners=False)
    c3 = c3 + outs[2]
    c4 = c4 + F.interpolate(outs[3], scale_factor=0.5, mode="bilinear", align_corners=False)
    return [_bn(c, sd, p + f"norm{i + 1}") for i, c in enumerate((c1, c2, c3, c4))]


# ---- Mask2Former (mask2former/video_seg_decoder.py) ---------------------------------------------------------------------
def _sine2d(h, w, F_=128, temp=10000.0):
    y = torch.arange(1, h + 1, dtype=torch.float32)[:, None].expand(h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32)[None, :].expand(h, w)
    y = y / (y[-1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, -1:] + 1e-6) * (2 * math.pi)
    dim_t = temp ** (2 * torch.div(torch.arange(F_, dtype=torch.float32), 2, rounding_mode="floor") / F_)
    px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(h * w, 2 * F_)


def _sine3d(t, h, w, F_=128, temp=10000.0):
    z = torch.arange(1, t + 1, dtype=torch.float32)[:, None, None].expand(t, h, w)
    y = torch.arange(1, h + 1, dtype=torch.float32)[None, :, None].expand(t, h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32)[None, None, :].expand(t, h, w)
    y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    z = z / (z[-1:, :, :] + 1e-6) * (2 * math.pi)
    dim_t = temp ** (2 * torch.div(torch.arange(F_, dtype=torch.float32), 2, rounding_mode="floor") / F_)
    dim_tz = temp ** (2 * torch.div(torch.arange(2 * F_, dtype=torch.float32), 2, rounding_mode="floor") / (2 * F_))
    px, py, pz = x[..., None] / dim_t, y[..., None] / dim_t, z[..., None] / dim_tz
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    pz = torch.stack((pz[..., 0::2].sin(), pz[..., 1::2].cos()), dim=4).flatten(3)
    return (torch.cat((py, px), dim=3) + pz).reshape(t * h * w, 2 * F_)


def mask2former(sd, feats, B, T=2, Q=100):
    """feats: 4 maps [B*T, 1024, h, w] (strides 4..32, frame index b*T+t) -> class logits [B,Q,21], mask logits [B,Q,T,h4,w4]."""
    pd = "mask2former.model.pixel_decoder."
    BT = B * T
    embeds, shapes = [], []
    for i, f in enumerate((feats[3], feats[2], feats[1])):
        e = F.group_norm(_conv(f, sd, pd + f"input_projections.{i}.0"), 32, sd[pd + f"input_projections.{i}.1.weight"],
                         sd[pd + f"input_projections.{i}.1.bias"], 1e-5)
        shapes.append((e.shape[2], e.shape[3]))
        embeds.append(e.flatten(2).transpose(1, 2))
    x = torch.cat(embeds, 1)
    pos = torch.cat([_sine2d(h, w) + sd[pd + "level_embed"][i][None] for i, (h, w) in enumerate(shapes)], 0)[None]
    ref = _ref_points(shapes)
    for i in range(6):
        p = pd + f"encoder.layers.{i}"
        a = _deform_attn(sd, p + ".self_attn", x + pos, ref, x, shapes, 8, 4)
        x = _ln(x + a, sd, p + ".self_attn_layer_norm", 1e-5)
        x = _ln(x + _lin(F.relu(_lin(x, sd, p + ".fc1")), sd, p + ".fc2"), sd, p + ".final_layer_norm", 1e-5)
    starts = [0, shapes[0][0] * shapes[0][1], shapes[0][0] * shapes[0][1] + shapes[1][0] * shapes[1][1]]
    lvl_maps = []
    for i, (h, w) in enumerate(shapes):
        lvl_maps.append(x[:, starts[i]:starts[i] + h * w].transpose(1, 2).reshape(BT, 256, h, w))
    cur = F.group_norm(F.conv2d(feats[0], sd[pd + "adapter_1.0.weight"]), 32, sd[pd + "adapter_1.1.weight"], sd[pd + "adapter_1.1.bias"], 1e-5)
    out = cur + F.interpolate(lvl_maps[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
    out = F.relu(F.group_norm(F.conv2d(out, sd[pd + "layer_1.0.weight"], padding=1), 32, sd[pd + "layer_1.1.weight"], sd[pd + "layer_1.1.bias"], 1e-5))
    mask_feat = _conv(out, sd, pd + "mask_projection")  # [BT,256,h4,w4]
    h4, w4 = mask_feat.shape[-2:]
    mask_feat = mask_feat.view(B, T, 256, h4, w4)
    tm = "mask2former.model.transformer_module."
    src, srcpos = [], []
    for i, (h, w) in enumerate(shapes):
        s = lvl_maps[i].flatten(2) + sd[tm + "level_embed.weight"][i][None, :, None]   # [BT,256,hw]
        s = s.view(B, T, 256, h * w).permute(0, 1, 3, 2).reshape(B, T * h * w, 256)
        src.append(s)
        srcpos.append(s + _sine3d(T, h, w)[None])
    hidden = sd[tm + "queries_features.weight"][None].expand(B, -1, -1)
    qpos = sd[tm + "queries_embedder.weight"][None].expand(B, -1, -1)
    dl = tm + "decoder."

    def predict(hid, target):
        inter = _ln(hid, sd, dl + "layernorm", 1e-5)
        e = inter
        for j in range(3):
            e = _lin(e, sd, dl + f"mask_predictor.mask_embedder.{j}.0")
            if j < 2:
                e = F.relu(e)
        logits = torch.einsum("bqc,btchw->bqthw", e, mask_feat)
        am = None
        if target is not None:
            am = F.interpolate(logits.flatten(0, 1), size=target, mode="bilinear", align_corners=False).view(B, Q, T, *target)
            am = am.sigmoid().flatten(2) < 0.5  # [B,Q,T*h*w]
        return inter, logits, am

    def mha(q, k, v, mask):
        nh, E = 8, 256
        qh = q.view(B, -1, nh, E // nh).transpose(1, 2)
        kh = k.view(B, -1, nh, E // nh).transpose(1, 2)
        vh = v.view(B, -1, nh, E // nh).transpose(1, 2)
        s = (qh @ kh.transpose(-2, -1)) * (E // nh) ** -0.5
        if mask is not None:
            s = s.masked_fill(mask[:, None], float("-inf"))
        return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, -1, E)

    inter, logits, am = predict(hidden, shapes[0])
    for idx in range(9):
        p = dl + f"layers.{idx}"
        li = idx % 3
        am = am.clone()
        am[am.all(-1)] = False
        W_, b_ = sd[p + ".cross_attn.in_proj_weight"], sd[p + ".cross_attn.in_proj_bias"]
        q = F.linear(hidden + qpos, W_[:256], b_[:256])
        k = F.linear(srcpos[li], W_[256:512], b_[256:512])
        v = F.linear(src[li], W_[512:], b_[512:])
        hidden = _ln(hidden + _lin(mha(q, k, v, am), sd, p + ".cross_attn.out_proj"), sd, p + ".cross_attn_layer_norm", 1e-5)
        hq = hidden + qpos
        a = mha(_lin(hq, sd, p + ".self_attn.q_proj"), _lin(hq, sd, p + ".self_attn.k_proj"), _lin(hidden, sd, p + ".self_attn.v_proj"), None)
        hidden = _ln(hidden + _lin(a, sd, p + ".self_attn.out_proj"), sd, p + ".self_attn_layer_norm", 1e-5)
        hidden = _ln(hidden + _lin(F.relu(_lin(hidden, sd, p + ".fc1")), sd, p + ".fc2"), sd, p + ".final_layer_norm", 1e-5)
        inter, logits, am = predict(hidden, None if idx == 8 else shapes[(idx + 1) % 3])
    cls = _lin(inter, sd, "mask2former.class_predictor")
    return cls, logits, dict(mask_features=mask_feat, ms=[m.view(B, T, 256, *m.shape[-2:]) for m in lvl_maps])


# ---- panoptic post-process (image_processing_video_mask2former.py:1238-1481 + model.py:231-312) --------------------------
def post_process(cls_logits, mask_logits, H, W, threshold=0.5, fuse=(0, 1)):
    B, Q, T, h, w = mask_logits.shape
    num_labels = cls_logits.shape[-1] - 1
    ml = mask_logits.permute(0, 2, 1, 3, 4).reshape(B * T, Q, h, w)
    ml = F.interpolate(ml, size=(256, 256), mode="bilinear", align_corners=False).view(B, T, Q, 256, 256)
    mask_probs = ml.sigmoid()
    class_probs = cls_logits.softmax(-1)
    scores, labels = class_probs.max(-1)
    results = []
    for i in range(B):
        keep = labels[i].ne(num_labels) & (scores[i] > threshold)
        mp, sc, lb, cp = mask_probs[i][:, keep], scores[i][keep], labels[i][keep], class_probs[i][keep]
        if sc.shape[0] == 0:
            qc = torch.zeros(T, 1, num_labels + 1, H, W)
            qc[:, 0, -1] = 1
            results.append(dict(segmentation=torch.zeros(T, H, W) - 1, segments_info=[], query_class_logits=qc, query_scores=[0.0]))
            continue
        seg = torch.zeros(T, H, W, dtype=torch.int32)
        mp = F.interpolate(mp, size=(H, W), mode="bilinear", align_corners=False)
        weighted = mp * sc[None, :, None, None]
        mlab = weighted.argmax(1)
        segments, keepq, keeps, cur, stuff = [], [], [], 0, {}
        for k in range(lb.shape[0]):
            pc = lb[k].item()
            should_fuse = pc in fuse
            mk = mlab == k
            area, orig = mk.sum(), (weighted[:, k] >= 0.5).sum()
            exists = bool(area > 0 and orig > 0)
            if exists and not (area / orig).item() > 0.8:
                exists = False
            if exists:
                if pc in stuff:
                    fid = stuff[pc]
                else:
                    cur += 1
                    fid = cur
                sid = cur if not should_fuse else fid
                seg[mk] = sid
                s6 = round(sc[k].item(), 6)
                segments.append({"id": sid, "label_id": pc, "was_fused": should_fuse, "score": s6})
                keepq.append(k)
                keeps.append(s6)
                if should_fuse and pc not in stuff:
                    stuff[pc] = cur
        qc = (cp[None, :, :, None, None] * mp[:, :, None])[:, keepq]
        if qc.shape[1] <= 0:
            qc = torch.zeros(T, 1, num_labels + 1, h, w)
            qc[:, 0, -1] = 1
        results.append(dict(segmentation=seg, segments_info=segments, query_class_logits=qc, query_scores=keeps))
    return results


@torch.no_grad()
def forward(sd, images, intrinsics, lift=True, stages=None):
    """SIU3RModel.forward (model.py:314-389).  Returns a dict of outputs in the reference's layouts."""
    B, V, _, H, W = images.shape
    bb = backbone(sd, images, intrinsics)
    gh, gw = bb["gh"], bb["gw"]
    ms1 = adapter(sd, images[:, 0], [t[:B] for t in bb["all_feat"]], gh, gw)
    ms2 = adapter(sd, images[:, 1], [t[B:] for t in bb["all_feat"]], gh, gw)
    feats = [torch.stack([a, b], dim=1).flatten(0, 1) for a, b in zip(ms1, ms2)]
    p1 = center_head(sd, "downstream_head1", bb["dec1"], gh, gw)
    p2 = center_head(sd, "downstream_head2", bb["dec2"], gh, gw)
    r1 = gs_head(sd, "gaussian_param_head1", bb["dec1"], images[:, 0], gh, gw)
    r2 = gs_head(sd, "gaussian_param_head2", bb["dec2"], images[:, 1], gh, gw)
    pts = torch.stack((p1.reshape(B, -1, 3), p2.reshape(B, -1, 3)), dim=1)
    raw = torch.stack((r1.flatten(2).transpose(1, 2), r2.flatten(2).transpose(1, 2)), dim=1)
    g = gaussian_adapter(pts, raw)
    cls, masks, aux = mask2former(sd, feats, B)
    res = post_process(cls, masks, H, W)
    sem = torch.zeros(B, 2, H, W, dtype=torch.int32)
    inst = torch.zeros(B, 2, H, W, dtype=torch.int32)
    for b, r in enumerate(res):
        for s in r["segments_info"]:
            m = r["segmentation"] == s["id"]
            sem[b][m] = s["label_id"] + 1
            inst[b][m] = s["id"]
    out = {k: v.flatten(1, 2) for k, v in g.items()}
    out.update(class_queries_logits=cls, masks_queries_logits=masks, semantic_labels=sem.flatten(1), instance_labels=inst.flatten(1),
               seg_masks=[r["segmentation"] for r in res], seg_infos=[r["segments_info"] for r in res],
               query_scores=[r["query_scores"] for r in res],
               seg_query_class_logits=[r["query_class_logits"].permute(0, 3, 4, 1, 2).flatten(0, 2) for r in res])
    if stages is not None:
        stages.update(enc={i: bb["all_feat"][i] for i in (5, 11, 17, 23)}, dec1=bb["dec1"], dec2=bb["dec2"], adapter=[ms1, ms2], gs_raw=[r1, r2],
                      pts3d=[p1, p2], m2f=aux)
    return out


@torch.no_grad()
def forward_multi(sd, images, intrinsics, lift=True, stages=None):
    """SIU3RMultiViewModel.forward (model_multi.py:310-392): head1 / gaussian_param_head1 on view 0, head2 / gaussian_param_head2
    on every other view (:175-215); adapter per view (:337), Mask2Former over the V frames (:355-359)."""
    B, V, _, H, W = images.shape
    bb = backbone_multi(sd, images, intrinsics)
    gh, gw = bb["gh"], bb["gw"]
    ms = [adapter(sd, images[:, v], [t[:, v] for t in bb["all_feat"]], gh, gw) for v in range(V)]
    feats = [torch.stack([ms[v][l] for v in range(V)], dim=1).flatten(0, 1) for l in range(4)]
    pts, raws = [], []
    for v in range(V):
        hn = "1" if v == 0 else "2"
        dec = [t[:, v] for t in bb["dec"]]
        pts.append(center_head(sd, "downstream_head" + hn, dec, gh, gw).reshape(B, -1, 3))
        raws.append(gs_head(sd, "gaussian_param_head" + hn, dec, images[:, v], gh, gw).flatten(2).transpose(1, 2))
    g = gaussian_adapter(torch.stack(pts, dim=1), torch.stack(raws, dim=1))
    cls, masks, aux = mask2former(sd, feats, B, T=V)
    res = post_process(cls, masks, H, W)
    sem = torch.zeros(B, V, H, W, dtype=torch.int32)
    inst = torch.zeros(B, V, H, W, dtype=torch.int32)
    for b, r in enumerate(res):
        for s in r["segments_info"]:
            m = r["segmentation"] == s["id"]
            sem[b][m] = s["label_id"] + 1
            inst[b][m] = s["id"]
    out = {k: v.flatten(1, 2) for k, v in g.items()}
    out.update(class_queries_logits=cls, masks_queries_logits=masks, semantic_labels=sem.flatten(1), instance_labels=inst.flatten(1),
               seg_masks=[r["segmentation"] for r in res], seg_infos=[r["segments_info"] for r in res],
               query_scores=[r["query_scores"] for r in res],
               seg_query_class_logits=[r["query_class_logits"].permute(0, 3, 4, 1, 2).flatten(0, 2) for r in res])
    if stages is not None:
        stages.update(enc={i: bb["all_feat"][i] for i in (5, 11, 17, 23)}, dec=bb["dec"], adapter=ms, gs_raw=raws, pts3d=pts, m2f=aux)
    return out
