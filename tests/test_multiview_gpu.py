"""BASELINE config 4 / SURVEY.md row M1: siu3r_b200.SIU3RMultiViewModel (V context views) against
  * golden fixtures generated from the UNMODIFIED reference SIU3RMultiViewModel (oracle/make_golden.py --views=4, CPU fp32), and
  * the oracle port (oracle/torch_port.py:forward_multi, itself pinned to those goldens in tests/test_oracle_cpu.py) on
    shapes the goldens do not cover (batch 2, V = 3: exercises the per-sample context assembly).

Tolerances as in tests/test_model_gpu.py: fp32x3 -> north-star (Gaussians 1e-3 abs, seg logits 1e-4 rel, labels exact);
tf32 -> the reference's own GPU numerics, looser bounds written below.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _samples(t: torch.Tensor, n=2048):
    t = t.contiguous().flatten()
    i = torch.arange(min(n, t.numel()), dtype=torch.int64, device=t.device)
    return t[(i * 2654435761 + 12345) % t.numel()].cpu().numpy()


def _build(S, precision):
    from siu3r_b200 import synth
    from siu3r_b200.model import ModelCfg, SIU3RMultiViewModel
    model = SIU3RMultiViewModel(ModelCfg(image_size=(S, S)), precision=precision)
    model.load_state_dict(synth.make_state_dict())
    model.cuda()
    return model


def _run_golden(S, V, precision, tol_stage, tol_gauss_abs, tol_logit_rel):
    from siu3r_b200 import synth
    path = os.path.join(GOLD, f"model_V{V}_S{S}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} missing")
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    model = _build(S, precision)
    model.capture = {}
    img, K = synth.pair_inputs(1, V, S)
    out = model(img.cuda(), K.cuda(), enable_query_class_logit_lift=True)
    torch.cuda.synchronize()
    cap = model.capture
    g, seg_out, seg_masks, seg_infos, qscores = out
    B, N = 1, (S // 16) ** 2 + 1
    d = {}
    for i in (5, 11, 17, 23):
        d[f"enc{i}"] = cap[f"enc{i}"].view(V * B, N, 1024)       # (b v) order == view-major for B = 1
    d["enc_norm"] = cap["enc_norm"].view(V * B, N, 1024)
    for i in (0, 5, 11):
        d[f"dec1_{i}"] = cap[f"dec1_{i}"].view(B, N, 768)
        d[f"dec2_{i}"] = cap[f"dec2_{i}"].view((V - 1) * B, N, 768)
    for v in range(V):
        for l in range(4):
            d[f"adapter_v{v}_f{l + 1}"] = cap["adapter_ms"][l][v::V].permute(0, 3, 1, 2)
        raw = cap["gs_raw"][0][0] if v == 0 else cap["gs_raw"][1][v - 1]
        d[f"gs_raw_{v}"] = raw.view(B, S, S, 83).permute(0, 3, 1, 2)
        d[f"pts3d_{v}"] = g.means.view(B, V, S, S, 3)[:, v]
    d["m2f_mask_features"] = cap["m2f_mask_features"].view(B, V, S // 4, S // 4, 256).permute(0, 1, 4, 2, 3)
    tok, off = cap["m2f_tokens"], 0
    for j, hw in enumerate((S // 32, S // 16, S // 8)):
        d[f"m2f_ms{j}"] = tok[:, off:off + hw * hw].reshape(B, V, hw, hw, 256).permute(0, 1, 4, 2, 3)
        off += hw * hw
    d["class_queries_logits"], d["masks_queries_logits"] = seg_out.class_queries_logits, seg_out.masks_queries_logits
    for name in ("means", "covariances", "harmonics", "opacities", "scales", "rotations"):
        d["g_" + name] = getattr(g, name)
    rep = []
    for name, t in d.items():
        assert list(t.shape) == meta[name]["shape"], (name, t.shape, meta[name]["shape"])
        e = float(np.abs(_samples(t) - z[name + "__samples"]).max())
        rep.append((name, e, e / max(meta[name]["absmax"], 1e-30)))
    msg = "\n".join(f"{n:28s} abs {e:.3e} rel {r:.3e}" for n, e, r in rep)
    print(f"\n[V={V} S={S} {precision}]\n{msg}")
    for n, e, r in rep:
        if n.startswith("g_") or n.startswith("pts3d"):
            assert e < tol_gauss_abs, (n, e, msg)
        elif n in ("class_queries_logits", "masks_queries_logits"):
            assert r < tol_logit_rel, (n, r, msg)
        else:
            assert r < tol_stage, (n, r, msg)
    return out, meta, z


@pytest.mark.parametrize("precision", ["h3", "fp32x3"])
@pytest.mark.parametrize("S", [64, 256, 512])
def test_multiview_fp32_grade_modes_meet_north_star(S, precision):
    out, meta, z = _run_golden(S, 4, precision, tol_stage=2e-4, tol_gauss_abs=1e-3, tol_logit_rel=1e-4)
    g, seg_out, seg_masks, seg_infos, qscores = out
    assert len(seg_infos[0]) == len(meta["seg_infos"][0])
    for a, b in zip(seg_infos[0], meta["seg_infos"][0]):
        assert (a["id"], a["label_id"], a["was_fused"]) == (b["id"], b["label_id"], b["was_fused"]) and abs(a["score"] - b["score"]) < 2e-4
    # label maps: equal up to pixels whose two best weighted mask probabilities tie to the last ulp (the reference's ATen resizes and our
    # kernels associate the bilinear sums differently): histograms within 1e-4 of the pixel count
    npix = g.semantic_labels.numel()
    sh = torch.bincount(g.semantic_labels.flatten().long(), minlength=22).tolist()
    ih = torch.bincount(g.instance_labels.flatten().long(), minlength=len(meta["inst_hist"])).tolist()
    assert len(sh) == len(meta["sem_hist"]) and sum(abs(a - b) for a, b in zip(sh, meta["sem_hist"])) <= 1e-4 * npix, (sh, meta["sem_hist"])
    assert len(ih) == len(meta["inst_hist"]) and sum(abs(a - b) for a, b in zip(ih, meta["inst_hist"])) <= 1e-4 * npix, (ih, meta["inst_hist"])
    sm = seg_masks[0]
    assert list(sm.shape) == meta["seg_mask0"]["shape"]
    assert float((_samples(sm).astype(np.int64) != z["seg_mask0__samples"].astype(np.int64)).mean()) < 2e-3
    qc = g.seg_query_class_logits[0]
    assert list(qc.shape) == meta["qc0"]["shape"] and np.abs(_samples(qc) - z["qc0__samples"]).max() < 1e-4


@pytest.mark.parametrize("S", [64, 256])
def test_multiview_tf32_reference_gpu_numerics(S):
    # TF32 mode = the reference's own GPU numerics: stage tensors ~1e-3 rel, Gaussians ~5e-3 abs (measured).  The query logits
    # pass through 9 masked-attention layers whose boolean masks are thresholds of the previous prediction: at 64x64 the key
    # maps are 2x2 / 4x4 / 8x8 pixels, a single flipped mask bit moves the logits by O(1) (measured 0.31 rel with 4 frames),
    # so at that size only the continuous stages are held to a tolerance; at 256x256 the logits are checked as well.
    _run_golden(S, 4, "tf32", tol_stage=1e-2, tol_gauss_abs=2e-2, tol_logit_rel=6e-2 if S >= 256 else 1.0)


def test_multiview_batch2_views3_matches_oracle_port():
    """B = 2, V = 3 at 64x64 (no golden): the context memories of every (sample, view) are assembled from the right rows."""
    from oracle import torch_port as TP
    from siu3r_b200 import synth
    S, B, V = 64, 2, 3
    img, K = synth.pair_inputs(B, V, S, seed=7)
    K = K.clone()
    K[1, :, 0, 0] *= 1.1      # different intrinsics tokens per sample
    K[:, 2, 1, 1] *= 0.9      # ... and per view
    ref = TP.forward_multi(synth.make_state_dict(), img, K)
    model = _build(S, "fp32x3")
    g, seg_out, seg_masks, seg_infos, _ = model(img.cuda(), K.cuda(), enable_query_class_logit_lift=True)
    for n in ("means", "covariances", "harmonics", "opacities", "scales", "rotations"):
        got, want = getattr(g, n).cpu(), ref[n]
        assert got.shape == want.shape, n
        assert float((got - want).abs().max()) < 1e-3, (n, float((got - want).abs().max()))
    ml, rl = seg_out.masks_queries_logits.cpu(), ref["masks_queries_logits"]
    assert ml.shape == rl.shape and float((ml - rl).abs().max()) < 1e-4 * float(rl.abs().max())
    cl, rc = seg_out.class_queries_logits.cpu(), ref["class_queries_logits"]
    assert float((cl - rc).abs().max()) < 1e-4 * float(rc.abs().max())
    # label maps: an argmax over score-weighted mask probabilities; pixels on a segment border can flip on 1e-6 differences
    for got, want in ((g.semantic_labels.cpu(), ref["semantic_labels"]), (g.instance_labels.cpu(), ref["instance_labels"])):
        assert got.shape == want.shape and float((got != want).float().mean()) < 2e-3
    assert [len(s) for s in seg_infos] == [len(s) for s in ref["seg_infos"]]


def test_multiview_v2_equals_pair_model_and_graph_replay():
    """With V = 2 the multi-view decoder degenerates to the pair decoder (same weights): both classes must agree; the CUDA-graph
    replay of the V = 4 forward reproduces the eager result exactly."""
    from siu3r_b200 import synth
    from siu3r_b200.model import ModelCfg, SIU3RModel
    S = 64
    img, K = synth.pair_inputs(1, 2, S, seed=3)
    pair = SIU3RModel(ModelCfg(image_size=(S, S)), precision="fp32x3")
    pair.load_state_dict(synth.make_state_dict())
    pair.cuda()
    g2 = pair(img.cuda(), K.cuda())[0]
    mv = _build(S, "fp32x3")
    gm = mv(img.cuda(), K.cuda())[0]
    for n in ("means", "covariances", "harmonics", "opacities"):
        assert float((getattr(g2, n) - getattr(gm, n)).abs().max()) < 1e-5, n
    img4, K4 = synth.pair_inputs(1, 4, S, seed=5)
    mt = _build(S, "tf32")
    eager = {n: getattr(mt(img4.cuda(), K4.cuda())[0], n).clone() for n in ("means", "harmonics", "opacities")}
    mt.enable_cuda_graph()
    for _ in range(2):
        gg = mt(img4.cuda(), K4.cuda())[0]
    torch.cuda.synchronize()
    for n, t in eager.items():
        assert float((getattr(gg, n) - t).abs().max()) < 1e-5, n
    with pytest.raises(AssertionError):
        pair(img4.cuda(), K4.cuda())   # the pair model refuses V != 2 (model.py:314-320 unpacks exactly two views)
