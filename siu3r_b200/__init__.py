"""siu3r_b200 -- B200-native (sm_100a) engine for the SIU3R per-image-pair hot path.

Public surface (mirrors the reference):
  siu3r_b200.SIU3RModel        <-> /root/reference/src/models/model.py:31  (forward :314-389)
  siu3r_b200.SIU3RMultiViewModel <-> /root/reference/src/models/model_multi.py:28 (forward :310-392; V >= 2 context views)
  siu3r_b200.SplattingCUDA     <-> /root/reference/src/models/gaussian_renderer.py:15 (forward :29-116)
  siu3r_b200.render_cuda       <-> /root/reference/src/models/cuda_splatting.py:46-122
  siu3r_b200.Gaussians         <-> /root/reference/src/utils/gaussians_types.py:4-38
  siu3r_b200.labels_from_qc_logits <-> the 2-D label extraction of /root/reference/src/pipeline.py:132-193 (viewer.py:422-435: viewer_labels)
  siu3r_b200.curope.{rope_2d, cuRoPE2D} <-> /root/reference/src/models/croco/curope/{curope.cpp:49-65, curope2d.py:32-40}
  siu3r_b200.PairPipeline      <-> the upload / forward / detach_cpu_copy loop of /root/reference/inference.py:119-141, overlapped
"""
from .gaussians import Gaussians  # noqa: F401


def __getattr__(name):
    if name in ("SIU3RModel", "SIU3RMultiViewModel", "ModelCfg"):
        from . import model as _m
        return getattr(_m, name)
    if name in ("SplattingCUDA", "render_cuda"):
        from . import renderer as _r
        return getattr(_r, name)
    if name in ("labels_from_qc_logits", "viewer_labels"):
        from . import labels2d as _l
        return getattr(_l, name)
    if name == "PairPipeline":
        from . import serving as _s
        return _s.PairPipeline
    raise AttributeError(name)
