"""Splatting front-end with the reference's call surface.

  SplattingCUDA.forward  <-> /root/reference/src/models/gaussian_renderer.py:29-116
  render_cuda            <-> /root/reference/src/models/cuda_splatting.py:46-122
  get_projection_matrix  <-> cuda_splatting.py:16-43 ;  get_fov <-> src/utils/projection.py:247-261

Camera set-up is host-side float32 arithmetic on tiny [b,4,4] matrices (plumbing); all per-Gaussian / per-pixel work
runs in the sm_100a rasterizer (csrc/raster.cu) through the C ABI.  `render_qc_logits=True` splats the per-Gaussian query x class
logits with the N-channel kernel that replaces gsplat.rasterization (gaussian_renderer.py:92-106).  Differences from the reference, by design:
  * no per-camera copy of the Gaussians (the reference `repeat`s them per view, gaussian_renderer.py:57-60);
  * covariances are consumed as the full [g,3,3] tensor and harmonics as [g,3,d_sh] -- the upper-triangle gather
    (cuda_splatting.py:107,115) and the SH transpose (:65) happen inside the preprocess kernel's loads.
"""
from __future__ import annotations

from math import isqrt

import torch

from . import ops
from .gaussians import Gaussians


def get_fov(intrinsics: torch.Tensor) -> torch.Tensor:
    """Normalised intrinsics [b,3,3] -> (fov_x, fov_y) [b,2] (host float32)."""
    K = intrinsics.detach().float().cpu()
    Kinv = torch.linalg.inv(K)

    def ray(v):
        d = Kinv @ torch.tensor(v, dtype=torch.float32)
        return d / d.norm(dim=-1, keepdim=True)

    left, right = ray([0.0, 0.5, 1.0]), ray([1.0, 0.5, 1.0])
    top, bottom = ray([0.5, 0.0, 1.0]), ray([0.5, 1.0, 1.0])
    fov_x = (left * right).sum(-1).acos()
    fov_y = (top * bottom).sum(-1).acos()
    return torch.stack((fov_x, fov_y), dim=-1)


def get_projection_matrix(near: torch.Tensor, far: torch.Tensor, fov_x: torch.Tensor, fov_y: torch.Tensor) -> torch.Tensor:
    """Frustum -> (-1,1) x (-1,1) x (0,1), z not flipped to (-1,1) (cuda_splatting.py:16-43)."""
    tx, ty = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    top, right = ty * near, tx * near
    b = near.shape[0]
    P = torch.zeros(b, 4, 4, dtype=torch.float32)
    P[:, 0, 0] = 2 * near / (right + right)
    P[:, 1, 1] = 2 * near / (top + top)
    P[:, 3, 2] = 1
    P[:, 2, 2] = far / (far - near)
    P[:, 2, 3] = -(far * near) / (far - near)
    return P


def camera_matrices(extrinsics: torch.Tensor, intrinsics: torch.Tensor, near: torch.Tensor, far: torch.Tensor):
    """Host-side camera set-up of render_cuda (cuda_splatting.py:68-77): returns float32 CPU tensors
    view [b,4,4] (= inverse(c2w)^T), full [b,4,4] (= view @ proj^T), campos [b,3], tan_fov_x/y [b]."""
    E = extrinsics.detach().float().cpu()
    fov = get_fov(intrinsics)
    fov_x, fov_y = fov[:, 0], fov[:, 1]
    tan_x, tan_y = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    proj = get_projection_matrix(near.detach().float().cpu(), far.detach().float().cpu(), fov_x, fov_y).transpose(1, 2)
    view = torch.linalg.inv(E).transpose(1, 2).contiguous()
    full = (view @ proj).contiguous()
    return view, full, E[:, :3, 3].contiguous(), tan_x, tan_y


def render_cuda(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means, gaussian_covariances,
                gaussian_sh_coefficients, gaussian_opacities, use_sh: bool = True, return_aux: bool = False):
    """Same contract as the reference's render_cuda: returns (color [b,3,h,w], depth [b,h,w]).
    Gaussian tensors may have batch size b (one set per camera) or 1 (shared by all cameras)."""
    assert use_sh, "colors_precomp path is not on the SIU3R hot path"
    b = extrinsics.shape[0]
    h, w = image_shape
    n = gaussian_sh_coefficients.shape[-1]
    degree = isqrt(n) - 1
    dev = gaussian_means.device
    view, full, campos, tan_x, tan_y = camera_matrices(extrinsics, intrinsics, near, far)
    cam = torch.cat([view.reshape(b, 16), full.reshape(b, 16), campos, background_color.detach().float().cpu().reshape(b, 3)], dim=1).to(dev)
    color = torch.empty(b, 3, h, w, device=dev)
    depth = torch.empty(b, h, w, device=dev)
    args = lambda i: (gaussian_means[i if gaussian_means.shape[0] == b else 0].contiguous(), gaussian_covariances[i if gaussian_covariances.shape[0] == b else 0].contiguous(),
                      gaussian_sh_coefficients[i if gaussian_sh_coefficients.shape[0] == b else 0].contiguous(),
                      gaussian_opacities[i if gaussian_opacities.shape[0] == b else 0].contiguous(), cam[i, 0:16], cam[i, 16:32], cam[i, 32:35], cam[i, 35:38],
                      float(tan_x[i]), float(tan_y[i]), h, w, degree)
    if return_aux:   # reference 5-tuple incl. n_touched / duplicate counts: the synchronising entry point
        aux = []
        for i in range(b):
            res = ops.raster_forward(*args(i), sh_layout=1, count_touched=True)
            color[i], depth[i] = res["color"], res["depth"]
            aux.append(res)
        return color, depth, aux
    # All cameras are enqueued back to back on one workspace without a host round trip (the reference rasterizer synchronises once per camera to size
    # its buffers); the per-camera status words are read ONCE afterwards, and a camera that overflowed the duplicate capacity is re-rendered.
    # n_touched / radii / opacity are discarded like in the reference (cuda_splatting.py:109,122).
    status = torch.zeros(b, 4, device=dev, dtype=torch.int32)
    ws = None
    G = gaussian_means.shape[1]
    scratch = dict(opacity=torch.empty(h, w, device=dev), radii=torch.empty(G, device=dev, dtype=torch.int32), n_touched=None)
    for i in range(b):
        r = ops.raster_forward_nosync(*args(i), sh_layout=1, status=status[i], ws=ws, out=dict(color=color[i], depth=depth[i], **scratch))
        ws = r["ws"]
    st = status.cpu()
    for i in range(b):
        if int(st[i, 2]) != 0:
            res = ops.raster_forward(*args(i), sh_layout=1, count_touched=False, dup_capacity=int(int(st[i, 0]) * 1.05) + 1024)
            color[i], depth[i] = res["color"], res["depth"]
    return color, depth


class SplattingCUDA:
    """Drop-in for the reference's nn.Module of the same name (no parameters; `background_color` is a non-persistent buffer)."""

    def __init__(self) -> None:
        self.near = 0.1
        self.far = 100.0
        self.scale_factor = 1 / self.near
        self.background_color = torch.tensor([0.0, 0.0, 0.0], dtype=torch.float32)

    def eval(self):
        return self

    def cuda(self):
        return self

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def forward(self, gaussians: Gaussians, extrinsics, intrinsics, image_shape, render_color: bool = True, render_feature: bool = False,
                render_id: bool = False, render_qc_logits: bool = False, cam_rot_delta=None, cam_trans_delta=None):
        assert cam_rot_delta is None and cam_trans_delta is None, "pose deltas are a training-only (backward) feature"
        b, v = extrinsics.shape[:2]
        E = extrinsics.detach().float().cpu().clone()
        E[..., :3, 3] = E[..., :3, 3] * self.scale_factor
        # NOTE: like the reference (gaussian_renderer.py:45-46) this rescales the Gaussians IN PLACE
        ops.scale_(gaussians.covariances, self.scale_factor ** 2)
        ops.scale_(gaussians.means, self.scale_factor)
        near, far = 1.0, self.far * self.scale_factor
        color = depth = None
        if render_color:
            means = gaussians.means
            cols, deps = [], []
            for bi in range(b):
                c, d = render_cuda(E[bi], intrinsics[bi], torch.full((v,), near), torch.full((v,), far), image_shape,
                                   self.background_color[None].repeat(v, 1), means[bi:bi + 1], gaussians.covariances[bi:bi + 1],
                                   gaussians.harmonics[bi:bi + 1], gaussians.opacities[bi:bi + 1])
                cols.append(c)
                deps.append(d)
            color = torch.stack(cols)
            depth = torch.stack(deps)
            color = ops.eltwise(ops.ELT_CLAMP01, color.contiguous())
        qc = None
        if render_qc_logits:
            # gaussian_renderer.py:75-110: per sample, splat the [n, q*c] query-class logits into every target view (gsplat semantics)
            h, w = image_shape
            qc = []
            for bi in range(b):
                logits = gaussians.seg_query_class_logits[bi]            # [n, q, c]
                n, q, c = logits.shape
                feats = logits.reshape(n, q * c).contiguous()
                Ks = intrinsics[bi].detach().float().cpu()
                viewmats = torch.linalg.inv(E[bi]).contiguous().to(gaussians.means.device)
                # all V cameras are enqueued on one workspace without a host round trip (binned per-tile sort); the status words are read once and a
                # camera that overflowed is re-rendered through the sizing path
                mg, cg, og = gaussians.means[bi].contiguous(), gaussians.covariances[bi].contiguous(), gaussians.opacities[bi].contiguous()
                intrs = [(float(Ks[vi, 0, 0]) * w, float(Ks[vi, 1, 1]) * h, float(Ks[vi, 0, 2]) * w, float(Ks[vi, 1, 2]) * h) for vi in range(v)]
                stacked = torch.empty(v, h, w, q * c, device=mg.device)
                status = torch.zeros(v, 4, device=mg.device, dtype=torch.int32)
                ws = None
                for vi in range(v):
                    ws = ops.raster_features_forward_nosync(mg, cg, og, feats, viewmats[vi], intrs[vi], near, far, h, w, status[vi], ws=ws, out=stacked[vi])["ws"]
                st = status.cpu()
                for vi in range(v):
                    if int(st[vi, 2]) != 0:
                        stacked[vi] = ops.raster_features_forward(mg, cg, og, feats, viewmats[vi], intrs[vi], near, far, h, w, want_alpha=False,
                                                                  dup_capacity=int(int(st[vi, 0]) * 1.05) + 1024)["features"]
                qc.append(stacked.view(v, h, w, q, c).permute(0, 3, 4, 1, 2))     # "n h w (q c) -> n q c h w" as a view
        return {"render_color": color, "render_depth": depth, "render_qc_logits": qc}
