"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch; gloo in CPU tests).

The per-pair path shards with NO data-path collective: pairs are independent units (every norm layer is
per-sample in eval mode).  The single collective of the north star is the all-gather of the packed per-Gaussian render
record, issued only when Gaussians of several pairs must be merged into one scene for joint rasterisation.
There is no precedent in the reference (its inference is single-GPU: inference.py:116-124).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

RECORD_FLOATS = 3 + 9 + 75 + 1  # means, covariances (3x3), harmonics (3x25), opacity


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block partition: rank r owns items [r*ceil.., ...) (batch 32 on 8 GPUs -> pairs 4r .. 4r+3)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return range(lo, min(n_items, lo + per))


def pack_render_record(g) -> torch.Tensor:
    """Gaussians (batch b) -> [b, G, 88] fp32 render record (what the rasterizer consumes)."""
    b, G = g.means.shape[:2]
    return torch.cat([g.means.reshape(b, G, 3), g.covariances.reshape(b, G, 9), g.harmonics.reshape(b, G, 75), g.opacities.reshape(b, G, 1)], dim=-1).contiguous()


def unpack_render_record(rec: torch.Tensor):
    b, G = rec.shape[:2]
    return (rec[..., 0:3].contiguous(), rec[..., 3:12].reshape(b, G, 3, 3).contiguous(), rec[..., 12:87].reshape(b, G, 3, 25).contiguous(),
            rec[..., 87].contiguous())


def all_gather_gaussians(rec: torch.Tensor, group=None) -> torch.Tensor:
    """ONE all-gather of the packed records: [b_local, G, 88] per rank -> [world * b_local, G, 88] on every rank."""
    world = dist.get_world_size(group)
    out = torch.empty((world * rec.shape[0],) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    return out
