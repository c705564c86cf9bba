"""Input / output surface of the reference's inference script, kept call-compatible.

  preprocess_image    <-> /root/reference/inference.py:13-38   (PIL LANCZOS resize of the short side + centre crop, /255)
  default_intrinsics  <-> /root/reference/inference.py:107-115 (pixel intrinsics normalised by the crop size)
  export_ply          <-> /root/reference/src/utils/ply_export.py:30-97 (same signature, same vertex layout / header)
  load_checkpoint     <-> Pipeline.load_from_checkpoint(...).model weights (src/pipeline.py:30-39): "model." prefixed keys

preprocess_image is host-side image decoding exactly as in the reference (PIL); export_ply packs the vertex records on the GPU
(csrc/ops_head.cu: ply_pack_kernel) so that ONE device-to-host copy of the packed buffer is the body of the file -- the reference
instead downloads every attribute, concatenates them in float64 and converts G rows to Python tuples (ply_export.py:94-95).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import _lib, ops


def preprocess_image(image_path, size: int = 256) -> torch.Tensor:
    """-> [3, size, size] float32 in [0, 1].  The reference hard-codes size = 256."""
    from PIL import Image
    image = image_path if isinstance(image_path, Image.Image) else Image.open(image_path)
    image = image.convert("RGB")
    W, H = image.size
    if W < H:
        new_W, new_H = size, int(H * (size / W))
        image = image.resize((new_W, new_H), Image.Resampling.LANCZOS)
        top = (new_H - size) // 2
        image = image.crop((0, top, new_W, top + size))
    else:
        new_H, new_W = size, int(W * (size / H))
        image = image.resize((new_W, new_H), Image.Resampling.LANCZOS)
        left = (new_W - size) // 2
        image = image.crop((left, 0, left + size, new_H))
    arr = np.array(image).astype(np.float32)
    return torch.from_numpy(arr).permute(2, 0, 1) / 255.0


def default_intrinsics(fx: float = 318.0, fy: float = 318.0, cx: float = 128.0, cy: float = 128.0, size: float = 256.0, views: int = 2) -> torch.Tensor:
    """[1, views, 3, 3] normalised intrinsics."""
    K = torch.tensor([[[fx / size, 0, cx / size], [0, fy / size, cy / size], [0, 0, 1]]], dtype=torch.float32)
    return K.repeat(1, views, 1, 1)


def load_checkpoint(path) -> dict:
    """State dict of the model from a Lightning checkpoint (keys prefixed "model.") or a plain state_dict file."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = ckpt.get("state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    if any(k.startswith("model.") for k in sd):
        sd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    return sd


def ply_attributes(num_rest: int, labels: bool = True, qc_words: int = 0) -> list:
    """[(name, ply type)] in file order (ply_export.py:12-27,55-71)."""
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(num_rest)] + ["opacity"]
    names += [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    out = [(n, "float") for n in names]
    if labels:
        out += [("semantic_label", "int"), ("instance_label", "int")]
    out += [(f"seg_query_class_logits_{i}", "float") for i in range(qc_words)]
    return out


def ply_header(count: int, attrs: list) -> bytes:
    """Header as plyfile.PlyData([PlyElement.describe(elements, "vertex")]).write() emits it (binary little endian, no comments)."""
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {count}"] + [f"property {t} {n}" for n, t in attrs] + ["end_header"]
    return ("\n".join(lines) + "\n").encode("ascii")


def pack_ply_records(means, scales, rotations, harmonics, opacities, semantic_labels=None, instance_labels=None, seg_query_class_logits=None,
                     save_sh_dc_only: bool = True) -> torch.Tensor:
    """Device tensors -> packed vertex records [G, F] (int32 view of the little-endian words) on the same device."""
    dev = means.device
    assert dev.type == "cuda", "pack_ply_records runs on the GPU (upload first)"
    f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
    means, scales, rotations, harmonics, opacities = f32(means), f32(scales), f32(rotations), f32(harmonics), f32(opacities)
    G, d_sh = means.shape[0], harmonics.shape[-1]
    labels = semantic_labels is not None and instance_labels is not None
    sem = semantic_labels.detach().to(dev, torch.int32).contiguous() if labels else None
    inst = instance_labels.detach().to(dev, torch.int32).contiguous() if labels else None
    qc, qc_words = None, 0
    if seg_query_class_logits is not None:
        qc = f32(seg_query_class_logits).view(G, -1)
        qc_words = qc.shape[1]
    lib = _lib.load()
    F = lib.siu3r_ply_record_words(d_sh, 1 if save_sh_dc_only else 0, 1 if labels else 0, qc_words)
    out = torch.empty(G, F, device=dev, dtype=torch.int32)
    p = lambda t: None if t is None else t.data_ptr()
    _lib.check(lib.siu3r_ply_pack(p(means), p(scales), p(rotations), p(harmonics), p(opacities), p(sem), p(inst), p(qc), G, d_sh,
                                  1 if save_sh_dc_only else 0, qc_words, p(out), ops._stream()), "ply_pack")
    return out


def export_ply(means, scales, rotations, harmonics, opacities, semantic_labels, instance_labels, seg_query_class_logits, path,
               shift_and_scale: bool = False, save_sh_dc_only: bool = True):
    """Same contract as the reference's export_ply; tensors may live on the GPU (preferred: no per-attribute download) or the CPU."""
    if shift_and_scale:
        raise NotImplementedError("shift_and_scale=True is a viewer convenience of the reference (ply_export.py:43-50), not on the hot path")
    dev = means.device if means.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    up = lambda t: None if t is None else t.to(dev)
    rec = pack_ply_records(up(means), up(scales), up(rotations), up(harmonics), up(opacities), up(semantic_labels), up(instance_labels),
                           up(seg_query_class_logits), save_sh_dc_only)
    host = torch.empty(rec.shape, dtype=rec.dtype, pin_memory=True)
    host.copy_(rec, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    d_sh = harmonics.shape[-1]
    labels = semantic_labels is not None and instance_labels is not None
    qc_words = 0 if seg_query_class_logits is None else seg_query_class_logits.shape[1] * seg_query_class_logits.shape[2]
    attrs = ply_attributes(0 if save_sh_dc_only else 3 * (d_sh - 1), labels, qc_words)
    path = Path(path)
    path.parent.mkdir(exist_ok=True, parents=True)
    with open(path, "wb") as f:
        f.write(ply_header(means.shape[0], attrs))
        f.write(host.numpy().tobytes())
    return path
