"""FULL-tensor parity of the engine on the GPU box, and the measured TF32 envelope of the reference's own GPU numerics.

The golden fixtures (tests/golden, generated from the unmodified reference on CPU) hold 2048 samples per tensor plus its mean / absmean;
a localized defect (tile edge, the last row of a 1025-token tile, an image border) could slip between samples.  Here the oracle port
(oracle/torch_port.py, pinned to those goldens in the CPU suite) is run ON THE GPU in plain fp32 (TF32 off: cuBLAS / cuDNN fp32 kernels,
test side only) and EVERY element of every stage tensor, Gaussian field and logit map is compared (max-abs over the whole tensor, i.e. over
every tile; the position of the worst element is printed); the fixtures' recorded means are asserted as well.

Envelope: the same port with torch.backends.*.allow_tf32 = True is what the reference runs on a GPU (croco/croco.py:13).  Its distance
from the fp32 evaluation is MEASURED here and the engine's "tf32" mode must stay within a small multiple of it, stage by stage.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_PORT = {}


def _port(S, tf32):
    """name -> tensor (reference layouts) of the oracle port evaluated on the GPU."""
    key = (S, tf32)
    if key in _PORT:
        return _PORT[key]
    from oracle import torch_port as TP
    from siu3r_b200 import synth
    sd = {k: v.cuda() for k, v in synth.make_state_dict().items()}
    img, K = synth.pair_inputs(1, 2, S)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        st = {}
        with torch.device("cuda"):
            out = TP.forward(sd, img.cuda(), K.cuda(), stages=st)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    d = {"g_" + n: out[n] for n in ("means", "covariances", "harmonics", "opacities", "scales", "rotations")}
    d["class_queries_logits"], d["masks_queries_logits"] = out["class_queries_logits"], out["masks_queries_logits"]
    for i in (5, 11, 17, 23):
        d[f"enc{i}"] = st["enc"][i]
    for i in (0, 5):
        d[f"dec1_{i}"], d[f"dec2_{i}"] = st["dec1"][i + 1], st["dec2"][i + 1]
    for v in range(2):
        for l in range(4):
            d[f"adapter_v{v}_f{l + 1}"] = st["adapter"][v][l]
        d[f"gs_raw_{v + 1}"], d[f"pts3d_{v + 1}"] = st["gs_raw"][v], st["pts3d"][v]
    d["m2f_mask_features"] = st["m2f"]["mask_features"]
    for j in range(3):
        d[f"m2f_ms{j}"] = st["m2f"]["ms"][j]
    d = {k: v.detach().float().cpu() for k, v in d.items()}
    d["_seg_infos"] = out["seg_infos"]
    d["_sem"] = out["semantic_labels"].cpu()
    d["_inst"] = out["instance_labels"].cpu()
    del sd
    torch.cuda.empty_cache()
    for k in [k for k in _PORT if k[0] != S]:
        del _PORT[k]
    _PORT[key] = d
    return d


def _engine(S, precision):
    import test_model_gpu as TM
    out, cap = TM._run(S, precision)
    d = TM._stage_tensors(out, cap, S)
    N = (S // 16) ** 2 + 1
    for k in list(d):
        if k.startswith("enc") and k != "enc_norm" or k.startswith("dec"):
            d[k] = d[k][:, : N - 1]     # the port strips the intrinsics token
    d.pop("enc_norm", None)
    for k in ("dec1_11", "dec2_11"):
        d.pop(k, None)
    return {k: v.detach().float().cpu() for k, v in d.items()}, out


def _compare(got: dict, ref: dict):
    rep = {}
    for name, r in ref.items():
        if name.startswith("_"):
            continue
        g = got[name]
        assert tuple(g.shape) == tuple(r.shape), (name, g.shape, r.shape)
        err = (g - r).abs()
        flat = int(err.argmax())
        rep[name] = (float(err.max()), float(err.max()) / max(float(r.abs().max()), 1e-30), flat, float(err.mean()))
    return rep


@pytest.mark.parametrize("S", [64, 256, 512])
def test_h3_full_tensors_vs_port_fp32_on_gpu(S):
    ref = _port(S, tf32=False)
    got, out = _engine(S, "h3")
    rep = _compare(got, ref)
    msg = "\n".join(f"{n:28s} abs {e:.3e} rel {r:.3e} (argmax {i}, mean abs err {m:.2e})" for n, (e, r, i, m) in rep.items())
    print(f"\n[full tensors, S={S}, h3 vs port-fp32-on-GPU]\n{msg}")
    # At 64^2 the masked-attention decoder is chaotic: its boolean attention masks are thresholded mask logits over 16 + 64 + 256 keys, and one flipped
    # bit moves the logits by O(1e-2).  The fp32 port ON THE GPU differs from the reference's CPU goldens by such a flip there (the engine agrees with
    # the CPU goldens: tests/test_model_gpu.py), so the logits are compared element by element only at 256^2 / 512^2, where no bit sits on the edge.
    chaotic = S == 64
    for n, (e, r, _, _) in rep.items():
        if n.startswith("g_") or n.startswith("pts3d"):
            assert e < 1e-3, (n, e, msg)
        elif n in ("class_queries_logits", "masks_queries_logits"):
            assert r < (2e-2 if chaotic else 1e-4), (n, r, msg)
        else:
            assert r < 2e-4, (n, r, msg)
    if chaotic:
        return
    # data-dependent branch: identical segments and label maps
    seg_infos = out[3]
    assert [(a["id"], a["label_id"], a["was_fused"]) for a in seg_infos[0]] == [(a["id"], a["label_id"], a["was_fused"]) for a in ref["_seg_infos"][0]]
    g = out[0]
    # label maps: identical up to pixels whose two best weighted mask probabilities tie to the last ulp (different association order of the two
    # bilinear resizes + sigmoid between ATen and our kernels)
    assert float((g.semantic_labels.cpu().flatten() != ref["_sem"].flatten()).float().mean()) < 1e-4
    assert float((g.instance_labels.cpu().flatten() != ref["_inst"].flatten()).float().mean()) < 1e-4
    # the golden fixtures' recorded first moments (from the unmodified reference on CPU)
    path = os.path.join(GOLD, f"model_S{S}.npz")
    if os.path.exists(path):
        meta = json.loads(str(np.load(path, allow_pickle=False)["meta"]))
        import test_model_gpu as TM
        full = TM._stage_tensors(*TM._run(S, "h3"), S)
        for name, t in full.items():
            if name not in meta or "mean" not in meta[name]:
                continue
            t = t.float()
            tol = 2e-4 * max(meta[name]["absmax"], 1e-30)
            assert abs(float(t.mean()) - meta[name]["mean"]) < tol, (name, float(t.mean()), meta[name]["mean"])
            assert abs(float(t.abs().mean()) - meta[name]["absmean"]) < tol, (name, float(t.abs().mean()), meta[name]["absmean"])


@pytest.mark.parametrize("S", [256, 512])
def test_tf32_mode_within_reference_gpu_envelope(S):
    """engine 'tf32' error vs fp32  <=  3 x (the port's own allow_tf32=True error vs fp32) + 1e-6 scale, per tensor (max-abs)."""
    ref = _port(S, tf32=False)
    env_t = _port(S, tf32=True)
    env = _compare({k: v for k, v in env_t.items() if not k.startswith("_")}, ref)
    got, _ = _engine(S, "tf32")
    rep = _compare(got, ref)
    lines = [f"{n:28s} engine-tf32 rel {rep[n][1]:.3e}   reference-tf32 envelope rel {env[n][1]:.3e}   ratio {rep[n][1] / max(env[n][1], 1e-30):.2f}" for n in rep]
    print(f"\n[TF32 envelope, S={S}]\n" + "\n".join(lines))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/tf32_envelope_S{S}.json", "w") as f:
        json.dump({n: {"engine_rel": rep[n][1], "engine_abs": rep[n][0], "reference_tf32_rel": env[n][1], "reference_tf32_abs": env[n][0]} for n in rep}, f, indent=1)
    for n in rep:
        assert rep[n][0] <= 3.0 * env[n][0] + 1e-6 * max(float(ref[n].abs().max()), 1.0), lines
