"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path (siu3r_b200/).

Model-path oracle: the reference's own PyTorch modules, imported
unmodified from /root/reference and run on CPU in fp32.  /root/reference only exists in
the build container, so this module is used there to (a) validate our engine directly
and (b) generate the golden fixtures under tests/golden/ that travel to the GPU box
(oracle/make_golden.py).  The weight generator below is plain torch-CPU code with no
dependency on the reference and is what both sides use to obtain identical weights.

Reference entry points exercised (file:line relative to /root/reference):
  src/models/model.py:314-389            SIU3RModel.forward
  src/models/backbone_croco.py:263-339   AsymmetricCroCo.forward
  src/models/vit_adapter/vit_adapter.py:393-441
  src/models/mask2former/video_seg_decoder.py:2351-2477
  src/models/mask2former/image_processing_video_mask2former.py:1238-1481
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

REFERENCE_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def _import_reference():
    if not reference_available():
        raise RuntimeError("/root/reference is not present (GPU box?) -- use tests/golden fixtures")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    stubs = os.path.join(_HERE, "stubs")
    if stubs not in sys.path:
        sys.path.append(stubs)


from siu3r_b200.synth import SHAPES_JSON, load_state_shapes, make_state_dict  # noqa: E402,F401  (reference-free generator)


def synthetic_inputs(batch: int, views: int, size, seed: int = 0):
    """images = rand(B,V,3,H,W) seeded (size = S or (H, W)); K = inference.py defaults (318/256, .5)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    H, W = (size, size) if isinstance(size, int) else size
    img = torch.rand(batch, views, 3, H, W, generator=g)
    K = torch.tensor([[318 / 256, 0, 0.5], [0, 318 / 256, 0.5], [0, 0, 1.0]])
    K = K[None, None].repeat(batch, views, 1, 1).contiguous()
    return img, K


# --------------------------------------------------------------------------------------
# Reference construction + staged forward
# --------------------------------------------------------------------------------------
def build_reference(size, state_dict: dict | None = None, multiview: bool = False):
    _import_reference()
    from src.config import CrocoCfg, GaussianHeadCfg, Mask2formerCfg, ModelCfg
    from src.utils.scannet_constant import PANOPTIC_SEMANTIC2NAME, STUFF_CLASSES

    cfg = ModelCfg(
        croco=CrocoCfg(),
        gaussian_head=GaussianHeadCfg(),
        image_size=[size, size] if isinstance(size, int) else list(size),
        pretrained_weights_path=None,
        mask2former=Mask2formerCfg(id2label=PANOPTIC_SEMANTIC2NAME, label_ids_to_fuse=STUFF_CLASSES),
    )
    if multiview:
        from src.models.model_multi import SIU3RMultiViewModel as Model
    else:
        from src.models.model import SIU3RModel as Model
    torch.manual_seed(0)
    model = Model(cfg).eval()
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
        assert all("criterion" in k or k == "backbone.mask_token" for k in missing), missing   # mask_token: unused by forward
    return model


def dump_state_shapes(path: str = SHAPES_JSON):
    model = build_reference(64)
    sd = model.state_dict()
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump({k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}, f)
    return path


@torch.no_grad()
def run_reference_stages(model, img, K) -> dict:
    """Run SIU3RModel.forward and capture stage-boundary tensors with forward hooks."""
    st: dict = {}
    hooks = []

    def grab(name, idx=None):
        def fn(_m, _inp, out):
            o = out if idx is None else out[idx]
            st[name] = o.detach().clone() if torch.is_tensor(o) else o
        return fn

    bb = model.backbone
    hooks.append(bb.patch_embed.register_forward_hook(grab("patch_embed", 0)))
    for i in (0, 5, 11, 17, 23):
        hooks.append(bb.enc_blocks[i].register_forward_hook(grab(f"enc{i}")))
    hooks.append(bb.enc_norm.register_forward_hook(grab("enc_norm")))
    for i in (0, 5, 11):
        hooks.append(bb.dec_blocks[i].register_forward_hook(grab(f"dec1_{i}", 0)))
        hooks.append(bb.dec_blocks2[i].register_forward_hook(grab(f"dec2_{i}", 0)))
    n_adapter = [0]

    def adapter_hook(_m, _inp, out):
        v = n_adapter[0]
        for j, f in enumerate(out):
            st[f"adapter_v{v}_f{j + 1}"] = f.detach().clone()
        n_adapter[0] += 1

    hooks.append(model.adapter.register_forward_hook(adapter_hook))
    hooks.append(model.downstream_head1.register_forward_hook(lambda m, i, o: st.__setitem__("pts3d_1", o["pts3d"].detach().clone())))
    hooks.append(model.downstream_head2.register_forward_hook(lambda m, i, o: st.__setitem__("pts3d_2", o["pts3d"].detach().clone())))
    hooks.append(model.gaussian_param_head1.register_forward_hook(grab("gs_raw_1")))
    hooks.append(model.gaussian_param_head2.register_forward_hook(grab("gs_raw_2")))
    pd = model.mask2former.model.pixel_decoder

    def pd_hook(_m, _inp, out):
        st["m2f_mask_features"] = out.mask_features.detach().clone()
        for j, f in enumerate(out.multi_scale_features):
            st[f"m2f_ms{j}"] = f.detach().clone()

    hooks.append(pd.register_forward_hook(pd_hook))

    out = model(img, K, enable_query_class_logit_lift=True)
    for h in hooks:
        h.remove()
    g, seg_out, seg_masks, seg_infos, q_scores = out
    st["class_queries_logits"] = seg_out.class_queries_logits
    st["masks_queries_logits"] = seg_out.masks_queries_logits
    for name in ("means", "covariances", "harmonics", "opacities", "scales", "rotations",
                 "semantic_labels", "instance_labels"):
        st["g_" + name] = getattr(g, name)
    st["seg_masks"] = [m.clone() for m in seg_masks]
    st["seg_infos"] = seg_infos
    st["query_scores"] = q_scores
    st["seg_query_class_logits"] = g.seg_query_class_logits
    return st


@torch.no_grad()
def run_reference_stages_multi(model, img, K) -> dict:
    """Run SIU3RMultiViewModel.forward (model_multi.py:310-392) and capture stage-boundary tensors with forward hooks.
    Per-view modules are called once per view (head2 / gaussian_param_head2 V-1 times): their outputs are stored per call."""
    st: dict = {}
    hooks = []

    def grab(name, idx=None):
        def fn(_m, _inp, out):
            o = out if idx is None else out[idx]
            st[name] = o.detach().clone()
        return fn

    bb = model.backbone
    for i in (0, 5, 11, 17, 23):
        hooks.append(bb.enc_blocks[i].register_forward_hook(grab(f"enc{i}")))
    hooks.append(bb.enc_norm.register_forward_hook(grab("enc_norm")))
    for i in (0, 5, 11):
        hooks.append(bb.dec_blocks[i].register_forward_hook(grab(f"dec1_{i}", 0)))
        hooks.append(bb.dec_blocks2[i].register_forward_hook(grab(f"dec2_{i}", 0)))
    counters = {"adapter": 0, "pts": 0, "raw": 0}

    def adapter_hook(_m, _inp, out):
        v = counters["adapter"]
        for j, f in enumerate(out):
            st[f"adapter_v{v}_f{j + 1}"] = f.detach().clone()
        counters["adapter"] += 1

    def pts_hook(_m, _inp, out):
        st[f"pts3d_{counters['pts']}"] = out["pts3d"].detach().clone()
        counters["pts"] += 1

    def raw_hook(_m, _inp, out):
        st[f"gs_raw_{counters['raw']}"] = out.detach().clone()
        counters["raw"] += 1

    hooks.append(model.adapter.register_forward_hook(adapter_hook))
    hooks.append(model.downstream_head1.register_forward_hook(pts_hook))   # view 0 first, then views 1.. (model_multi.py:175-185)
    hooks.append(model.downstream_head2.register_forward_hook(pts_hook))
    hooks.append(model.gaussian_param_head1.register_forward_hook(raw_hook))
    hooks.append(model.gaussian_param_head2.register_forward_hook(raw_hook))
    pd = model.mask2former.model.pixel_decoder

    def pd_hook(_m, _inp, out):
        st["m2f_mask_features"] = out.mask_features.detach().clone()
        for j, f in enumerate(out.multi_scale_features):
            st[f"m2f_ms{j}"] = f.detach().clone()

    hooks.append(pd.register_forward_hook(pd_hook))
    out = model(img, K, enable_query_class_logit_lift=True)
    for h in hooks:
        h.remove()
    g, seg_out, seg_masks, seg_infos, q_scores = out
    st["class_queries_logits"] = seg_out.class_queries_logits
    st["masks_queries_logits"] = seg_out.masks_queries_logits
    for name in ("means", "covariances", "harmonics", "opacities", "scales", "rotations", "semantic_labels", "instance_labels"):
        st["g_" + name] = getattr(g, name)
    st["seg_masks"] = [m.clone() for m in seg_masks]
    st["seg_infos"] = seg_infos
    st["query_scores"] = q_scores
    st["seg_query_class_logits"] = g.seg_query_class_logits
    return st


def sample_indices(numel: int, n: int = 2048) -> torch.Tensor:
    """Fixed pseudo-random flat indices used by the golden fixtures (both sides)."""
    i = torch.arange(min(n, numel), dtype=torch.int64)
    return (i * 2654435761 + 12345) % numel


def summarize(t: torch.Tensor, n: int = 2048) -> dict:
    tf = t.detach().to(torch.float64).flatten()
    idx = sample_indices(tf.numel(), n)
    return {
        "shape": list(t.shape),
        "mean": float(tf.mean()),
        "absmean": float(tf.abs().mean()),
        "absmax": float(tf.abs().max()),
        "samples": t.detach().flatten()[idx].to(torch.float32 if t.is_floating_point() else torch.int64).numpy(),
    }
