"""Synthetic inputs of the BASELINE.json workloads: seeded, host-generated, no reference dependency."""
from __future__ import annotations

import json
import os
import zlib

import torch

SHAPES_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "state_shapes.json")  # reference state_dict key set: name -> [shape, dtype]


def pair_inputs(batch: int, views: int, size, seed: int = 0):
    """images rand(B,V,3,H,W) in [0,1] (size = S for square frames or (H, W)); K = inference.py defaults (fx = fy = 318/256, cx = cy = 0.5) normalised."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    H, W = (size, size) if isinstance(size, int) else size
    img = torch.rand(batch, views, 3, H, W, generator=g)
    K = torch.tensor([[318 / 256, 0, 0.5], [0, 318 / 256, 0.5], [0, 0, 1.0]])
    return img, K[None, None].repeat(batch, views, 1, 1).contiguous()


def _quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    i, j, k, r = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def raster_scene(G: int, H: int, W: int, seed: int = 0, pixel_aligned: bool = False):
    """Config 5 of BASELINE.json: camera at the origin looking down +z, normalised focal 318/256, scene already in the
    renderer's x10 units (z ~ U[5, 80]).  Returns CPU tensors: means [G,3], covariances [G,3,3], harmonics [G,3,25],
    opacities [G], extrinsics [1,4,4] (c2w), intrinsics [1,3,3] (normalised), near, far."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7000 + seed)
    f = 318 / 256
    z = 5 + 75 * torch.rand(G, generator=g)
    u = torch.rand(G, generator=g) * 2 - 1  # NDC in [-1, 1]
    v = torch.rand(G, generator=g) * 2 - 1
    x = u * z * (0.5 / f)
    y = v * z * (0.5 / f)  # fy is normalised by H, so NDC y = v as well
    means = torch.stack((x, y, z), -1)
    if pixel_aligned:
        sigma = (z / (f * W))[:, None] * (0.5 + torch.rand(G, 3, generator=g))
    else:
        sigma = (0.01 * torch.nn.functional.softplus(torch.randn(G, 3, generator=g))).clamp_max(3.0)
    q = torch.randn(G, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    R = _quat_to_rot(q)
    cov = R @ torch.diag_embed(sigma * sigma) @ R.transpose(-1, -2)
    opac = torch.sigmoid(torch.randn(G, generator=g))
    mask = torch.ones(25)
    for d in range(1, 5):
        mask[d * d:(d + 1) * (d + 1)] = 0.1 * 0.25 ** d
    harm = torch.randn(G, 3, 25, generator=g) * mask
    E = torch.eye(4)[None].clone()
    K = torch.tensor([[f, 0, 0.5], [0, f, 0.5], [0, 0, 1.0]])[None]
    return dict(means=means.contiguous(), covariances=cov.contiguous(), harmonics=harm.contiguous(), opacities=opac.contiguous(),
                extrinsics=E, intrinsics=K, near=torch.tensor([1.0]), far=torch.tensor([1000.0]))


def load_state_shapes(path: str = SHAPES_JSON) -> dict:
    with open(path) as f:
        return json.load(f)


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def make_state_dict(shapes: dict | None = None, seed: int = 0, populated: bool = False) -> dict:
    """Seeded synthetic weights for the full reference key set (655.5 M params).

    No checkpoint is available offline (README.md:35,84 of the reference are downloads), so
    parity is checked with random weights.  The rules keep every path numerically alive
    (non-zero biases, spread deformable offsets, peaked class logits so that the panoptic
    post-process takes its populated branch) and keep activations O(1) so that the
    north-star tolerances are meaningful.

    populated=False (parity fixtures): every Mask2Former query ends up with nearly the same state (27 unit-gain post-norm sublayers wash the query
    identity out), the mask logits are far from their thresholds and the network is well conditioned end to end -- the segmentation logits can be
    held to the 1e-4 north-star tolerance -- but the panoptic post-process takes its empty branch at 256^2 / 512^2.
    populated=True (bench.py, tests/test_populated_gpu.py): unit-variance learned queries + damped decoder sublayers keep the 100 queries distinct and a
    class bias makes most of them void; at 512^2 19 queries pass the score test, 13 fail the area test and 6 survive, four of them fused into one
    "floor" segment (inference.py's workload).  The price is conditioning: the masked-attention decoder thresholds ~10^7 mask logits into boolean
    attention masks, some of them sit within 1e-4 of the threshold, and ANY two evaluations that differ in the last bits (our modes, the fp32 port
    on CPU vs GPU) flip a few mask bits and move the class logits by 1e-3..1e-2 -- so that preset is checked on everything but the decoder logits.
    """
    if shapes is None:
        shapes = load_state_shapes()
    sd = {}
    for key, (shape, dtype) in shapes.items():
        g = _gen(key, seed)
        if dtype == "torch.int64":
            sd[key] = torch.zeros(shape, dtype=torch.int64)
            continue
        if key.startswith("mask2former.criterion"):
            # training-only buffer (empty_weight); value irrelevant for the forward path
            sd[key] = torch.ones(shape, dtype=torch.float32)
            continue
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":  # LayerNorm / BatchNorm / GroupNorm scale
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:  # biases
            if "sampling_offsets" in key:
                t = 2.0 * torch.randn(shape, generator=g)
            else:
                t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            gain = 1.0
            if "sampling_offsets" in key:
                gain = 0.5
            if key.startswith("mask2former.class_predictor"):
                gain = 12.0  # peaked class distribution -> scores > 0.5 for some queries
            if ".dpt.head.4." in key:
                # centre heads: keep ||xyz|| (the argument of expm1) ~1 so |means| stays in a room-scale range (< ~10)
                gain = 0.12 if key.startswith("downstream") else 0.3
            if leaf in ("level_embed",) or "queries_" in key or "level_embed" in key:
                gain = 1.0
            # populated preset (see the docstring): distinct queries
            if populated and "queries_features" in key:
                gain = 16.0
            if populated and "transformer_module.decoder.layers" in key and (".out_proj." in key or ".fc2." in key):
                gain = 0.3
            t = gain * torch.randn(shape, generator=g) / (fan_in ** 0.5)
        if populated and key == "mask2former.class_predictor.bias":
            t[20] += 30.0   # most queries void ...
            t[0] += 8.0     # ... and the two stuff classes (wall, floor) favoured so that fused segments occur
            t[1] += 8.0
        sd[key] = t.to(torch.float32)
    return sd
