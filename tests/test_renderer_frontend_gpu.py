"""R1 on the GPU: siu3r_b200.renderer.SplattingCUDA.forward (render_color + render_qc_logits) against tests/golden/renderer_frontend.npz = the
outputs of the reference's OWN SplattingCUDA.forward -> render_cuda front-end (gaussian_renderer.py:29-116, cuda_splatting.py:46-122) run
unmodified on the same scene, with the two third-party rasterizers replaced at the import boundary by the CPU oracles
(oracle/make_golden_renderer.py).  Covers the in-place x10 / x100 rescale, near / far, the per-view loop over rectangular 64 x 96 frames with
three different poses and an off-centre principal point, the [0, 1] clamp, and the n q c h w layout of the rendered logits."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _samples(a, k=4096):
    a = np.ascontiguousarray(a).reshape(-1)
    i = np.arange(min(k, a.size), dtype=np.int64)
    return a[(i * 2654435761 + 12345) % a.size]


def test_splatting_forward_matches_reference_frontend():
    from oracle import make_golden_renderer as MG
    from siu3r_b200.gaussians import Gaussians
    from siu3r_b200.renderer import SplattingCUDA
    z = np.load(os.path.join(GOLD, "renderer_frontend.npz"))
    means, cov, harm, opac, E, K, qc = MG.scene()
    g = Gaussians(means=means.cuda(), covariances=cov.cuda(), harmonics=harm.cuda(), opacities=opac.cuda(), scales=torch.zeros(1, MG.G, 3).cuda(),
                  rotations=torch.zeros(1, MG.G, 4).cuda())
    g.seg_query_class_logits = [qc.cuda()]
    out = SplattingCUDA()(g, E, K, (MG.H, MG.W), render_color=True, render_qc_logits=True)
    torch.cuda.synchronize()
    # in-place rescale of the caller's Gaussians, exactly like the reference (gaussian_renderer.py:45-46)
    assert np.array_equal(_samples(g.means.cpu().numpy()), z["out_means_samples"])
    assert np.array_equal(_samples(g.covariances.cpu().numpy()), z["out_cov_samples"])
    color, depth = out["render_color"].cpu().numpy(), out["render_depth"].cpu().numpy()
    assert color.shape == z["render_color"].shape and depth.shape == z["render_depth"].shape
    assert color.min() >= 0.0 and color.max() <= 1.0
    assert np.abs(color - z["render_color"]).max() < 1e-3                      # north star: rendered RGB within 1e-3 abs
    assert np.abs(depth - z["render_depth"]).max() < 1e-3 * max(1.0, float(z["render_depth"].max()))
    ql = out["render_qc_logits"][0]
    assert list(ql.shape) == list(z["render_qc_logits_shape"])
    assert np.abs(_samples(ql.contiguous().cpu().numpy(), 16384) - z["render_qc_logits_samples"]).max() < 1e-3
    assert abs(float(ql.double().sum()) - float(z["render_qc_logits_sum"])) < 1e-4 * abs(float(z["render_qc_logits_sum"]))
