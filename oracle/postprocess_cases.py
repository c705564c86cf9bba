"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path (siu3r_b200/).

Crafted Mask2Former logits that drive every data-dependent branch of the panoptic post-process
(/root/reference/src/models/mask2former/image_processing_video_mask2former.py:1238-1481 + src/models/model.py:231-312):
kept queries, stuff fusing of classes {0, 1}, area-ratio rejection, "kept but nothing survives" (:1468-1472) and "no mask found"
(:1351-1375).  The logits are built from integer random fields with exactly representable arithmetic only (no interpolation, no
transcendental), so the generator yields bit-identical tensors on every machine; oracle/make_golden_postprocess.py feeds them to the
reference's own functions and stores what comes back in tests/golden/postprocess_S*.npz.
"""
from __future__ import annotations

import torch

Q, T, NUM_LABELS = 100, 2, 20
CASES = ("fused", "area_reject", "no_survivor", "empty", "mixed")


def _field(gen, h, w, cell):
    """blocky integer field in {-10, -8, ..., -2} (background: every sigmoid below 0.12)"""
    gy, gx = h // cell, w // cell
    f = torch.randint(1, 6, (Q, T, gy, gx), generator=gen).float() * -2.0
    return f.repeat_interleave(cell, 2).repeat_interleave(cell, 3)


def _rect(gen, h, w, cell):
    gy, gx = h // cell, w // cell
    y0 = int(torch.randint(0, gy - 1, (1,), generator=gen))
    x0 = int(torch.randint(0, gx - 1, (1,), generator=gen))
    y1 = y0 + 1 + int(torch.randint(0, gy - y0 - 1, (1,), generator=gen))
    x1 = x0 + 1 + int(torch.randint(0, gx - x0 - 1, (1,), generator=gen))
    return y0 * cell, y1 * cell, x0 * cell, x1 * cell


def make_case(name: str, S: int):
    """-> class logits [1, Q, 21], mask logits [1, Q, T, S/4, S/4] (CPU fp32, deterministic)."""
    assert name in CASES and S % 32 == 0
    h = w = S // 4
    cell = max(h // 8, 1)
    gen = torch.Generator().manual_seed(1234 + CASES.index(name) * 17 + S)
    cls = torch.full((1, Q, NUM_LABELS + 1), -3.0)
    cls[:, :, NUM_LABELS] = 3.0                               # every query void unless set below
    masks = _field(gen, h, w, cell)[None].clone()

    def keep(q, c, level):                                    # query q predicts class c with logit `level` (probability > 0.5 from level >= 2)
        cls[0, q, NUM_LABELS] = -3.0
        cls[0, q, c] = level

    def paint(q, value=8.0, frames=(0, 1)):
        y0, y1, x0, x1 = _rect(gen, h, w, cell)
        for t in frames:
            masks[0, q, t, y0:y1, x0:x1] = value
        return y0, y1, x0, x1

    if name == "fused":
        for q, c, lv in ((3, 0, 4.0), (10, 0, 4.5), (20, 1, 5.0), (33, 7, 5.5), (47, 7, 6.0), (60, 12, 6.5)):
            keep(q, c, lv)
            paint(q)
    elif name == "area_reject":
        keep(1, 5, 7.0)
        y0, y1, x0, x1 = paint(1)
        keep(2, 6, 2.5)                                       # lower score, the same region plus one extra cell column: loses most of its area to query 1
        masks[0, 2] = -10.0                                   # (and never wins a background pixel: lowest logit, lowest score)
        masks[0, 2, :, y0:y1, x0:x1] = 8.0
        xe = min(x1 + cell, w)
        masks[0, 2, :, y0:y1, x1:xe] = 8.0
        keep(9, 0, 5.0)
        paint(9)
    elif name == "no_survivor":
        keep(5, 3, 6.0)                                       # confident class, but its mask never reaches 0.5
        keep(6, 1, 4.0)
        masks[0, 5] = -9.0
        masks[0, 6] = -9.0
    elif name == "empty":
        pass
    elif name == "mixed":
        classes = [0, 0, 1, 1, 2, 5, 5, 9, 13, 17, 19, 4]
        for i, c in enumerate(classes):
            q = 7 * i + 2
            keep(q, c, 2.0 + 0.5 * i)
            paint(q, 8.0 if i % 3 else 4.0, frames=(0, 1) if i % 4 else (0,))
    return cls, masks
