"""Generates tests/golden/postprocess_S{64,512}.npz by running the reference's OWN panoptic post-process on crafted logits.

Reference functions executed unmodified (imported from /root/reference; CPU fp32):
  VideoMask2FormerImageProcessor.post_process_panoptic_segmentation   src/models/mask2former/image_processing_video_mask2former.py:1238-1481
  SIU3RModel.post_process_gaussians                                   src/models/model.py:231-312 (called unbound with a stand-in `self`
                                                                      that carries .processor and .cfg.mask2former, nothing else is read)
Inputs come from oracle/postprocess_cases.py (deterministic, exactly representable).  Stored per case: segments_info, query scores, the full
segmentation map, the per-Gaussian semantic / instance labels, and shape + 4096 samples + sum of the lifted query-class logits.

    python oracle/make_golden_postprocess.py            (needs /root/reference; CPU only; ~1 min)
"""
from __future__ import annotations

import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import postprocess_cases as PC  # noqa: E402
from oracle.ref_model import _import_reference  # noqa: E402


def sample_idx(n: int, k: int = 4096) -> np.ndarray:
    i = np.arange(min(k, n), dtype=np.int64)
    return (i * 2654435761 + 12345) % n


def run_reference(cls, masks, S):
    _import_reference()
    from src.models.mask2former.image_processing_video_mask2former import VideoMask2FormerImageProcessor
    from src.models.model import SIU3RModel
    from src.utils.gaussians_types import Gaussians
    from src.utils.scannet_constant import STUFF_CLASSES
    fake = SimpleNamespace(processor=VideoMask2FormerImageProcessor(),
                           cfg=SimpleNamespace(mask2former=SimpleNamespace(seg_threshold=0.5, label_ids_to_fuse=STUFF_CLASSES)))
    B = 1
    G = S * S
    z = lambda *s: torch.zeros(B, 2, G, *s)
    g = Gaussians(means=z(3), covariances=z(3, 3), harmonics=z(3, 25), opacities=z(), scales=z(3), rotations=z(4))
    seg_out = SimpleNamespace(class_queries_logits=cls, masks_queries_logits=masks)
    with torch.no_grad():
        g, _, seg_masks, seg_infos, qscores = SIU3RModel.post_process_gaussians(fake, B, S, S, g, seg_out, enable_query_class_logit_lift=True)
    return g, seg_masks, seg_infos, qscores


def main():
    for S in (64, 512):
        out = {}
        meta = {}
        for name in PC.CASES:
            cls, masks = PC.make_case(name, S)
            g, seg_masks, seg_infos, qscores = run_reference(cls, masks, S)
            sm = seg_masks[0]
            qc = g.seg_query_class_logits[0]
            out[f"{name}__seg_mask"] = sm.numpy().astype(np.int16)
            out[f"{name}__sem"] = g.semantic_labels[0].numpy().astype(np.int8)
            out[f"{name}__inst"] = g.instance_labels[0].numpy().astype(np.int8)
            flat = qc.reshape(-1).numpy()
            out[f"{name}__qc_samples"] = flat[sample_idx(flat.size)]
            meta[name] = dict(seg_infos=seg_infos[0], query_scores=[float(s) for s in qscores[0]], qc_shape=list(qc.shape),
                              qc_sum=float(qc.double().sum()), seg_mask_dtype=str(sm.dtype).replace("torch.", ""),
                              n_segments=len(seg_infos[0]))
            print(S, name, "segments", [(s["id"], s["label_id"], s["was_fused"]) for s in seg_infos[0]], "qc", list(qc.shape))
        out["meta"] = np.array(json.dumps(meta))
        path = os.path.join(ROOT, "tests", "golden", f"postprocess_S{S}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
