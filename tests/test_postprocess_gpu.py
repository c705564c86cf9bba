"""H13 on the GPU engine: SIU3RModel._post_process (host decisions on 100 scalars + ops_head.cu kernels) driven with the crafted logits of
oracle/postprocess_cases.py at 64^2 and at the BASELINE size 512^2, lift = True, against goldens produced by the reference's own
post_process_panoptic_segmentation + SIU3RModel.post_process_gaussians (tests/golden/postprocess_S*.npz, oracle/make_golden_postprocess.py;
reference: image_processing_video_mask2former.py:1238-1481, model.py:231-312).  Branches: stuff fusing, area-ratio rejection,
kept-but-nothing-survives (:1468-1472), no mask found (:1351-1375)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("S", [64, 512])
def test_engine_post_process_branches_match_reference(S):
    from oracle import postprocess_cases as PC
    from siu3r_b200.model import ModelCfg, SIU3RModel
    z = np.load(os.path.join(GOLD, f"postprocess_S{S}.npz"))
    meta = json.loads(str(z["meta"]))
    model = SIU3RModel(ModelCfg(image_size=(S, S)))
    model.dev = torch.device("cuda", torch.cuda.current_device())
    for name in PC.CASES:
        cls, masks = PC.make_case(name, S)
        B, Q, T, h, w = masks.shape
        ml = masks.permute(0, 2, 3, 4, 1).reshape(B * T, h, w, Q).contiguous().cuda()     # the engine's pixel-major mask logits
        seg_masks, seg_infos, qc_list, qscores, sem, inst = model._post_process(cls.cuda(), ml, B, S, S, True, T)
        torch.cuda.synchronize()
        m = meta[name]
        assert [(s["id"], s["label_id"], s["was_fused"]) for s in seg_infos[0]] == [(s["id"], s["label_id"], s["was_fused"]) for s in m["seg_infos"]], name
        assert np.allclose([s["score"] for s in seg_infos[0]], [s["score"] for s in m["seg_infos"]], atol=2e-6), name
        assert np.allclose(qscores[0], m["query_scores"], atol=2e-6), name
        sm = seg_masks[0].cpu().numpy()
        ref = z[f"{name}__seg_mask"]
        # the kernels evaluate the two bilinear resizes + sigmoid in fp32 like ATen but not in the same association order: a pixel whose two
        # best weighted probabilities tie to the last ulp may flip; everything else must be identical
        mism = float((sm.astype(np.int16) != ref).mean())
        assert mism < 1e-4, (name, mism)
        assert float((sem[0].cpu().numpy().astype(np.int8) != z[f"{name}__sem"]).mean()) < 1e-4, name
        assert float((inst[0].cpu().numpy().astype(np.int8) != z[f"{name}__inst"]).mean()) < 1e-4, name
        qc = qc_list[0]
        assert list(qc.shape) == m["qc_shape"], (name, qc.shape, m["qc_shape"])
        flat = qc.reshape(-1)
        i = torch.arange(min(4096, flat.numel()), dtype=torch.int64, device=flat.device)
        got = flat[(i * 2654435761 + 12345) % flat.numel()].cpu().numpy()
        assert np.abs(got - z[f"{name}__qc_samples"]).max() < 1e-5, name
        assert abs(float(qc.double().sum()) - m["qc_sum"]) < 1e-5 * max(1.0, abs(m["qc_sum"])), name
