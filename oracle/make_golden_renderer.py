"""Generates tests/golden/renderer_frontend.npz by running the reference's OWN splatting front-end, unmodified:

  SplattingCUDA.forward   /root/reference/src/models/gaussian_renderer.py:29-116   (x10 / x100 in-place rescale, near / far, clamp, Ks scaling)
  render_cuda             /root/reference/src/models/cuda_splatting.py:46-122      (get_fov, projection / view / full matrices, SH transpose,
                                                                                   upper-triangle covariance gather)
  get_fov                 /root/reference/src/utils/projection.py:247-261

The two THIRD-PARTY rasterizers it imports (diff_gaussian_rasterization, gsplat: un-vendored CUDA packages, absent here) are replaced at the
import boundary by recorders that (a) store every argument the reference hands over -- that is the front-end's whole output -- and (b) answer
with the CPU oracles (oracle/raster_ref.c, oracle/gsplat_ref.py; both still "parity unpinned" against the packages themselves).  The reference
hard-codes device="cuda" for the near / far scalars (gaussian_renderer.py:52-53); torch.tensor is wrapped for the duration of the call so that
they land on the CPU.  Stored: the recorded per-camera arguments, the rescaled Gaussians, render_color / render_depth / render_qc_logits.

    python oracle/make_golden_renderer.py            (needs /root/reference; CPU only)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gsplat_ref, raster_oracle  # noqa: E402
from oracle.ref_model import _import_reference  # noqa: E402

G, H, W, V, QC = 3000, 64, 96, 3, (2, 21)
RECORD = {"dgr": [], "gsplat": []}


def scene():
    """Room-scale scene in the MODEL's units (the renderer rescales x10): what SIU3RModel.forward hands to the renderer."""
    from siu3r_b200 import synth
    sc = synth.raster_scene(G, H, W, seed=5)
    g = torch.Generator().manual_seed(99)
    means = sc["means"] / 10.0
    cov = sc["covariances"] / 100.0
    E = torch.eye(4)[None, None].repeat(1, V, 1, 1).clone()
    E[0, 1, :3, 3] = torch.tensor([0.05, -0.02, 0.1])
    a = 0.05
    E[0, 2, :3, :3] = torch.tensor([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], dtype=torch.float32)
    E[0, 2, :3, 3] = torch.tensor([-0.08, 0.03, -0.05])
    K = torch.tensor([[318 / 256, 0, 0.5], [0, 318 / 256 * W / H, 0.5], [0, 0, 1.0]])[None, None].repeat(1, V, 1, 1).clone()
    K[0, 1, 0, 2] = 0.47
    qc = torch.rand(G, *QC, generator=g)
    return means[None].contiguous(), cov[None].contiguous(), sc["harmonics"][None].contiguous(), sc["opacities"][None].contiguous(), E, K, qc


def install_stubs():
    dgr = types.ModuleType("diff_gaussian_rasterization")

    class GaussianRasterizationSettings:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class GaussianRasterizer:
        def __init__(self, raster_settings):
            self.s = raster_settings

        def __call__(self, means3D, means2D, shs, colors_precomp, opacities, cov3D_precomp, theta=None, rho=None):
            s = self.s
            rec = dict(viewmatrix=s.viewmatrix.numpy().copy(), projmatrix=s.projmatrix.numpy().copy(), projmatrix_raw=s.projmatrix_raw.numpy().copy(),
                       campos=s.campos.numpy().copy(), bg=s.bg.numpy().copy(), tanfovx=np.float32(s.tanfovx), tanfovy=np.float32(s.tanfovy),
                       sh_degree=np.int32(s.sh_degree), image_hw=np.array([s.image_height, s.image_width], np.int32),
                       means3D=means3D.numpy().copy(), shs=shs.numpy().copy(), opacities=opacities.numpy().copy(), cov3D=cov3D_precomp.numpy().copy())
            RECORD["dgr"].append(rec)
            r = raster_oracle.rasterize(rec["means3D"], rec["cov3D"], rec["shs"], rec["opacities"][:, 0], rec["viewmatrix"], rec["projmatrix"], rec["campos"],
                                        float(s.tanfovx), float(s.tanfovy), s.image_height, s.image_width, int(s.sh_degree), bg=rec["bg"], debug=False)
            t = torch.from_numpy
            return t(r["color"]), t(r["radii"]), t(r["depth"])[None], t(r["opacity"])[None], t(r["n_touched"])

    dgr.GaussianRasterizationSettings, dgr.GaussianRasterizer = GaussianRasterizationSettings, GaussianRasterizer
    sys.modules["diff_gaussian_rasterization"] = dgr

    gs = types.ModuleType("gsplat")

    def rasterization(means, quats, scales, covars, opacities, colors, viewmats, Ks, width, height, sh_degree, near_plane, far_plane):
        assert quats is None and scales is None and sh_degree is None
        RECORD["gsplat"].append(dict(viewmats=viewmats.numpy().copy(), Ks=Ks.numpy().copy(), near=np.float32(near_plane), far=np.float32(far_plane),
                                     wh=np.array([width, height], np.int32)))
        outs = []
        for i in range(viewmats.shape[0]):
            Ki = Ks[i].numpy()
            r = gsplat_ref.rasterize(means.numpy(), covars.numpy(), opacities.numpy(), colors.numpy(), viewmats[i].numpy(), float(Ki[0, 0]), float(Ki[1, 1]),
                                     float(Ki[0, 2]), float(Ki[1, 2]), width, height, float(near_plane), float(far_plane))
            outs.append(torch.from_numpy(np.asarray(r[0] if isinstance(r, tuple) else r["features"], dtype=np.float32)))
        return torch.stack(outs), None, None

    gs.rasterization = rasterization
    sys.modules["gsplat"] = gs


class _TorchOnCpu:
    """torch, except that torch.tensor(..., device="cuda") lands on the CPU (gaussian_renderer.py:52-53 hard-codes the device)."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def tensor(*a, **k):
        k.pop("device", None)
        return torch.tensor(*a, **k)


def main():
    _import_reference()
    install_stubs()
    import src.models.gaussian_renderer as GR
    from src.utils.gaussians_types import Gaussians
    from src.utils.projection import get_fov
    GR.torch = _TorchOnCpu()
    means, cov, harm, opac, E, K, qc = scene()
    g = Gaussians(means=means.clone(), covariances=cov.clone(), harmonics=harm.clone(), opacities=opac.clone(), scales=torch.zeros(1, G, 3),
                  rotations=torch.zeros(1, G, 4))
    g.seg_query_class_logits = [qc.clone()]
    with torch.no_grad():
        out = GR.SplattingCUDA()(g, E.clone(), K.clone(), (H, W), render_color=True, render_qc_logits=True)
    def samples(a, k=4096):
        a = np.ascontiguousarray(a).reshape(-1)
        i = np.arange(min(k, a.size), dtype=np.int64)
        return a[(i * 2654435761 + 12345) % a.size]

    # (the scene is regenerated by the tests from scene(): only what the reference produced is stored)
    qcl = out["render_qc_logits"][0].numpy()
    res = dict(E=E.numpy(), K=K.numpy(), out_means_samples=samples(g.means.numpy()), out_cov_samples=samples(g.covariances.numpy()),
               render_color=out["render_color"].numpy(), render_depth=out["render_depth"].numpy(), render_qc_logits_samples=samples(qcl, 16384),
               render_qc_logits_shape=np.array(qcl.shape, np.int32), render_qc_logits_sum=np.float64(qcl.astype(np.float64).sum()),
               fov=get_fov(K[0]).numpy())
    for i, rec in enumerate(RECORD["dgr"]):
        for k in ("viewmatrix", "projmatrix", "projmatrix_raw", "campos", "bg", "tanfovx", "tanfovy", "sh_degree", "image_hw"):
            res[f"cam{i}_{k}"] = rec[k]
        if i == 0:   # the Gaussian operands are the same for every camera of the sample
            for k in ("means3D", "shs", "opacities", "cov3D"):
                res[f"dgr_{k}_samples"], res[f"dgr_{k}_shape"] = samples(rec[k]), np.array(rec[k].shape, np.int32)
    gsr = RECORD["gsplat"][0]
    res.update(gs_viewmats=gsr["viewmats"], gs_Ks=gsr["Ks"], gs_near=gsr["near"], gs_far=gsr["far"], gs_wh=gsr["wh"])
    path = os.path.join(ROOT, "tests", "golden", "renderer_frontend.npz")
    np.savez_compressed(path, **res)
    print("cameras recorded:", len(RECORD["dgr"]), "color range", float(out["render_color"].min()), float(out["render_color"].max()),
          "depth max", float(out["render_depth"].max()))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
