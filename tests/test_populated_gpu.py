"""The bench preset (synth.make_state_dict(populated=True)): at 512^2 the panoptic post-process takes its populated branch inside the real forward --
19 queries pass the score test, 13 fail the area test, 6 survive, four of them fused into one "floor" segment -- with enable_query_class_logit_lift=True,
as the reference's inference.py runs it.  The engine (h3 mode) is compared, on full tensors, with the oracle port evaluated in fp32 on the GPU with the
same weights.  Everything up to the masked-attention decoder is held to the usual tolerances; the decoder logits are chaotic under this preset (a
mask logit within ~1e-4 of its threshold flips a boolean attention-mask bit: any two fp32 evaluations that differ in the last bits disagree by
1e-3 .. 1e-2 there, see synth.make_state_dict), so they are bounded loosely, and the discrete outputs are compared as such."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S", [256, 512])
def test_populated_preset_forward_with_lift(S):
    from oracle import torch_port as TP
    from siu3r_b200 import synth
    from siu3r_b200.model import ModelCfg, SIU3RModel
    sd = synth.make_state_dict(populated=True)
    img, K = synth.pair_inputs(1, 2, S)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.device("cuda"):
            ref = TP.forward({k: v.cuda() for k, v in sd.items()}, img.cuda(), K.cuda())
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    model = SIU3RModel(ModelCfg(image_size=(S, S)), precision="h3")
    model.load_state_dict(sd)
    model.cuda()
    g, seg_out, seg_masks, seg_infos, qscores = model(img.cuda(), K.cuda(), enable_query_class_logit_lift=True)
    torch.cuda.synchronize()
    for n in ("means", "covariances", "harmonics", "opacities", "scales", "rotations"):
        assert float((getattr(g, n) - ref[n]).abs().max()) < 1e-3, n
    ml, rl = seg_out.masks_queries_logits, ref["masks_queries_logits"]
    cl, rc = seg_out.class_queries_logits, ref["class_queries_logits"]
    e_m = float((ml - rl).abs().max() / rl.abs().max())
    e_c = float((cl - rc).abs().max() / rc.abs().max())
    print(f"\n[populated preset, S={S}] mask-logit rel err {e_m:.3e}, class-logit rel err {e_c:.3e}, segments {[(s['id'], s['label_id'], s['was_fused']) for s in seg_infos[0]]}")
    assert e_m < 5e-2 and e_c < 5e-2
    assert len(seg_infos[0]) >= (5 if S == 512 else 3) and any(s["was_fused"] for s in seg_infos[0])
    assert [(s["id"], s["label_id"], s["was_fused"]) for s in seg_infos[0]] == [(s["id"], s["label_id"], s["was_fused"]) for s in ref["seg_infos"][0]]
    assert float((g.semantic_labels.flatten() != ref["semantic_labels"].flatten()).float().mean()) < 2e-3
    assert float((g.instance_labels.flatten() != ref["instance_labels"].flatten()).float().mean()) < 2e-3
    qc, rq = g.seg_query_class_logits[0], ref["seg_query_class_logits"][0]
    assert qc.shape == rq.shape and float((qc - rq).abs().max()) < 5e-2
