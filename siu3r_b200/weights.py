"""One-time repack of the reference state_dict (SURVEY.md Appendix C key set, unchanged) into kernel-friendly layouts.

Host-side only (runs once at load): conv weights [Cout,Cin,KH,KW] -> [Cout,(KH,KW,Cin)] K-major GEMM operands, eval-mode
BatchNorm folded into the preceding conv, ConvTranspose(k = s) -> [(dy,dx,Cout), Cin] GEMM operand with tiled bias,
q/k/v style projections that share an input concatenated into one operand, 3-channel image convs padded to 4 channels.
"""
from __future__ import annotations

import torch

from .ops import Weight


class Packer:
    def __init__(self, sd: dict, device, precision: int):
        # accept both raw SIU3RModel keys and Lightning-checkpoint keys ("model." prefix, pipeline.py:30)
        if not any(k.startswith("backbone.") for k in sd) and any(k.startswith("model.backbone.") for k in sd):
            sd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
        self.sd = sd
        self.dev = device
        self.prec = precision

    def t(self, key: str) -> torch.Tensor:
        return self.sd[key].detach().float()

    def has(self, key: str) -> bool:
        return key in self.sd

    def vec(self, key: str) -> torch.Tensor:
        return self.t(key).contiguous().to(self.dev)

    def W(self, w: torch.Tensor, b: torch.Tensor | None) -> Weight:
        return Weight(w.to(self.dev), None if b is None else b.to(self.dev), self.prec)

    def linear(self, prefix: str, bias: bool = True) -> Weight:
        return self.W(self.t(prefix + ".weight"), self.t(prefix + ".bias") if bias and self.has(prefix + ".bias") else None)

    def linear_cat(self, prefixes: list[str]) -> Weight:
        w = torch.cat([self.t(p + ".weight") for p in prefixes], 0)
        b = torch.cat([self.t(p + ".bias") for p in prefixes], 0)
        return self.W(w, b)

    def linear_rows(self, wkey: str, bkey: str, r0: int, r1: int) -> Weight:
        return self.W(self.t(wkey)[r0:r1], self.t(bkey)[r0:r1])

    def bn_scale_shift(self, prefix: str, eps: float = 1e-5):
        g, b = self.t(prefix + ".weight"), self.t(prefix + ".bias")
        m, v = self.t(prefix + ".running_mean"), self.t(prefix + ".running_var")
        scale = g / torch.sqrt(v + eps)
        return scale, b - m * scale

    def conv(self, prefix: str, bn: str | None = None, pad_cin_to: int | None = None, extra_bias: torch.Tensor | None = None) -> Weight:
        w = self.t(prefix + ".weight")  # [Cout, Cin, KH, KW]
        b = self.t(prefix + ".bias") if self.has(prefix + ".bias") else None
        if bn is not None:
            scale, shift = self.bn_scale_shift(bn)
            w = w * scale[:, None, None, None]
            b = shift if b is None else b * scale + shift
        if pad_cin_to is not None and w.shape[1] < pad_cin_to:
            wp = torch.zeros(w.shape[0], pad_cin_to, w.shape[2], w.shape[3])
            wp[:, : w.shape[1]] = w
            w = wp
        if extra_bias is not None:
            b = extra_bias if b is None else b + extra_bias
        w2 = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()
        return self.W(w2, b)

    def conv_transpose(self, prefix: str) -> tuple[Weight, int]:
        w = self.t(prefix + ".weight")  # [Cin, Cout, s, s]
        b = self.t(prefix + ".bias")
        s = w.shape[2]
        w2 = w.permute(2, 3, 1, 0).reshape(s * s * w.shape[1], w.shape[0]).contiguous()
        return self.W(w2, b.repeat(s * s)), s

    def dwconv(self, prefix: str):
        w = self.t(prefix + ".weight")  # [C,1,3,3]
        return w.view(w.shape[0], 9).t().contiguous().to(self.dev), self.vec(prefix + ".bias")


def sine_pos_2d(h: int, w: int, num_pos_feats: int = 128, temperature: float = 10000.0) -> torch.Tensor:
    """Normalised 2-D sine embedding of mask2former/video_seg_decoder.py:704-735 for an unpadded h x w map -> [h*w, 2F]."""
    scale, eps = 2 * torch.pi, 1e-6
    y = torch.arange(1, h + 1, dtype=torch.float32)[:, None].expand(h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32)[None, :].expand(h, w)
    y = y / (y[-1:, :] + eps) * scale
    x = x / (x[:, -1:] + eps) * scale
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)
    px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(h * w, 2 * num_pos_feats).contiguous()


def sine_pos_3d(t: int, h: int, w: int, num_pos_feats: int = 128, temperature: float = 10000.0) -> torch.Tensor:
    """3-D (frame, y, x) sine embedding of video_seg_decoder.py:628-679 -> [t*h*w, 2F] in (t, h, w) order."""
    scale, eps = 2 * torch.pi, 1e-6
    z = torch.arange(1, t + 1, dtype=torch.float32)[:, None, None].expand(t, h, w)
    y = torch.arange(1, h + 1, dtype=torch.float32)[None, :, None].expand(t, h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32)[None, None, :].expand(t, h, w)
    y = y / (y[:, -1:, :] + eps) * scale
    x = x / (x[:, :, -1:] + eps) * scale
    z = z / (z[-1:, :, :] + eps) * scale
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)
    iz = torch.arange(num_pos_feats * 2, dtype=torch.float32)
    dim_tz = temperature ** (2 * torch.div(iz, 2, rounding_mode="floor") / (num_pos_feats * 2))
    px, py, pz = x[..., None] / dim_t, y[..., None] / dim_t, z[..., None] / dim_tz
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    pz = torch.stack((pz[..., 0::2].sin(), pz[..., 1::2].cos()), dim=4).flatten(3)
    return (torch.cat((py, px), dim=3) + pz).reshape(t * h * w, 2 * num_pos_feats).contiguous()


def reference_points(shapes) -> torch.Tensor:
    """Normalised (x, y) centres of every cell of every level, concatenated (vit_adapter/blocks.py:10-24,
    video_seg_decoder.py:1848-1881 with valid_ratio = 1) -> [sum(h*w), 2]."""
    out = []
    for h, w in shapes:
        ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h, dtype=torch.float32), torch.linspace(0.5, w - 0.5, w, dtype=torch.float32),
                                indexing="ij")
        out.append(torch.stack((rx.reshape(-1) / w, ry.reshape(-1) / h), -1))
    return torch.cat(out, 0).contiguous()
