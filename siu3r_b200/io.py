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

import functools
import math
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


class _Placeholder:
    """Stands in for a class of the reference's own packages that a checkpoint pickled by reference (the Lightning checkpoint stores
    hyper_parameters["cfg"] = RootCfg from src.config, with members from src.data.config etc.: pipeline.py:26,39): keeps whatever state it is given."""

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self._state = state


class _TolerantPickle:
    """pickle_module for torch.load: classes that cannot be imported (the reference tree is not on sys.path) become _Placeholder subclasses
    instead of failing the whole load -- only the tensors of "state_dict" are wanted."""
    import pickle as _pickle
    __name__ = "siu3r_b200_tolerant_pickle"
    load, loads, dump, dumps = _pickle.load, _pickle.loads, _pickle.dump, _pickle.dumps
    HIGHEST_PROTOCOL, DEFAULT_PROTOCOL = _pickle.HIGHEST_PROTOCOL, _pickle.DEFAULT_PROTOCOL
    PickleError, PicklingError, UnpicklingError, Pickler = _pickle.PickleError, _pickle.PicklingError, _pickle.UnpicklingError, _pickle.Pickler

    # Only what a tensor checkpoint needs is ever resolved to a real object; EVERY other global a pickle names -- the reference's config dataclasses,
    # Lightning callbacks, or something hostile -- becomes an inert _Placeholder subclass (never imported, never called with side effects).
    _SAFE_MODULES = ("torch._utils", "torch.storage", "torch._tensor", "torch.serialization", "collections", "numpy.core.multiarray",
                     "numpy._core.multiarray", "numpy", "_codecs")
    _SAFE_NAMES = {("torch", n) for n in ("FloatStorage", "HalfStorage", "BFloat16Storage", "DoubleStorage", "LongStorage", "IntStorage", "ShortStorage",
                                          "CharStorage", "ByteStorage", "BoolStorage", "Size", "device", "dtype", "Tensor", "float32", "float16", "bfloat16",
                                          "float64", "int64", "int32", "int16", "int8", "uint8", "bool")}
    _SAFE_NUMPY = ("dtype", "ndarray", "_reconstruct", "scalar", "float32", "float64", "int64", "int32", "bool_")

    class Unpickler(_pickle.Unpickler):
        def find_class(self, module, name):
            T = _TolerantPickle
            ok = (module, name) in T._SAFE_NAMES or (module in T._SAFE_MODULES and not name.startswith("__") and
                                                      (not module.startswith("numpy") or name in T._SAFE_NUMPY) and
                                                      (module != "_codecs" or name == "encode") and
                                                      (module != "collections" or name == "OrderedDict"))
            if ok:
                try:
                    return super().find_class(module, name)
                except (ImportError, AttributeError):
                    pass
            return type(name, (_Placeholder,), {"__module__": module})


def load_checkpoint(path) -> dict:
    """State dict of the model from a Lightning checkpoint (keys prefixed "model.") or a plain state_dict file.  The reference's checkpoint
    also pickles its config dataclasses (src.config.RootCfg, ...); they are not needed and are tolerated when that package is absent."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_TolerantPickle)
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


# ---- image ingest on the GPU ---------------------------------------------------------------------------
@functools.lru_cache(maxsize=256)
def lanczos_tables(in_size: int, out_size: int):
    """Window bounds [out, 2] (first tap, tap count), fixed-point coefficients [out, ksize] (int32, 22 fractional bits) and ksize of one
    LANCZOS pass, computed as Pillow does (libImaging/Resample.c: precompute_coeffs with support 3 + normalize_coeffs_8bpc; double
    precision, libm sin through math.sin).  in_size == out_size -> the identity pass (Pillow skips a pass that does not change the size)."""
    if in_size == out_size:
        bounds = np.stack([np.arange(out_size), np.ones(out_size, np.int64)], 1).astype(np.int32)
        return bounds, np.full((out_size, 1), 1 << 22, np.int32), 1
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale

    def sinc(x):
        if x == 0.0:
            return 1.0
        x = x * math.pi
        return math.sin(x) / x

    def lanczos(x):
        return sinc(x) * sinc(x / 3) if -3.0 <= x < 3.0 else 0.0

    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in enumerate(w):
            if ww != 0.0:
                v = v / ww
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def resize_plan(W: int, H: int, size: int = 256):
    """(new_W, new_H, crop_left, crop_top) of the reference recipe (inference.py:16-33), including its float rounding of the long side."""
    if W < H:
        new_W, new_H = size, int(H * (size / W))
        return new_W, new_H, 0, (new_H - size) // 2
    new_H, new_W = size, int(W * (size / H))
    return new_W, new_H, (new_W - size) // 2, 0


def ingest_plan(W: int, H: int, size: int = 256) -> dict:
    """Everything siu3r_resize_lanczos_u8 needs for a W x H frame, as host values: resized size, crop origin (may be negative: resize_plan),
    the two coefficient tables, and the range of source rows [row0, row0 + rows) that the cropped output rows read -- only those go through
    the horizontal pass (Pillow makes the same cut, ImagingResample's ybox_first / ybox_last)."""
    new_W, new_H, cx, cy = resize_plan(W, H, size)
    bx, kx, ksx = lanczos_tables(W, new_W)
    by, ky, ksy = lanczos_tables(H, new_H)
    y_first, y_last = max(cy, 0), min(cy + size, new_H) - 1               # output rows of the crop window that exist in the resized image
    row0 = int(by[y_first, 0])
    rows = int(by[y_last, 0] + by[y_last, 1]) - row0
    return dict(new_W=new_W, new_H=new_H, crop_x=cx, crop_y=cy, row0=row0, rows=rows, bounds_x=bx, kx=np.ascontiguousarray(kx), ksize_x=ksx,
                bounds_y=by, ky=np.ascontiguousarray(ky), ksize_y=ksy)


_DEVICE_PLANS = {}


def _device_plan(W, H, size, dev):
    key = (W, H, size, str(dev))
    if key not in _DEVICE_PLANS:
        plan = ingest_plan(W, H, size)
        for k in ("bounds_x", "kx", "bounds_y", "ky"):
            plan[k] = torch.from_numpy(plan[k]).to(dev)
        _DEVICE_PLANS[key] = plan
    return _DEVICE_PLANS[key]


def preprocess_image_cuda(image, size: int = 256, device=None, out: torch.Tensor = None) -> torch.Tensor:
    """GPU form of preprocess_image: the decoded 8-bit RGB frame is uploaded as it is ([H, W, 3] uint8: a path, a PIL image, a numpy array or
    a torch uint8 tensor, host or device) and resized / cropped / scaled on the device -> [3, size, size] float32, bit-identical to
    preprocess_image (csrc/resize.cu).  `out` (optional): a [3, size, size] float32 CUDA view to fill, e.g. a slot of a [B, V, 3, S, S] batch."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if isinstance(image, torch.Tensor):
        frame = image
    else:
        if not isinstance(image, np.ndarray):
            from PIL import Image
            pil = image if isinstance(image, Image.Image) else Image.open(image)
            image = np.array(pil.convert("RGB"))
        frame = torch.from_numpy(np.ascontiguousarray(image))
    assert frame.dtype == torch.uint8 and frame.dim() == 3 and frame.shape[2] == 3, "expected an [H, W, 3] uint8 RGB frame"
    frame = frame.to(dev, non_blocking=True).contiguous()
    H, W = int(frame.shape[0]), int(frame.shape[1])
    p = _device_plan(W, H, size, dev)
    tmp = torch.empty(p["rows"], size, 3, device=dev, dtype=torch.uint8)
    if out is None:
        out = torch.empty(3, size, size, device=dev, dtype=torch.float32)
    assert out.shape == (3, size, size) and out.dtype == torch.float32 and out.is_cuda and out.is_contiguous()
    _lib.check(lib.siu3r_resize_lanczos_u8(frame.data_ptr(), H, W, W * 3, p["bounds_x"].data_ptr(), p["kx"].data_ptr(), p["ksize_x"], p["new_W"],
                                           p["bounds_y"].data_ptr(), p["ky"].data_ptr(), p["ksize_y"], p["new_H"], p["crop_x"], p["crop_y"], size, size,
                                           p["row0"], p["rows"], tmp.data_ptr(), out.data_ptr(), ops._stream()), "resize_lanczos_u8")
    return out


def preprocess_views_cuda(images, size: int = 256, device=None) -> torch.Tensor:
    """Batched ingest: V frames -> [1, V, 3, size, size] (the `images` tensor of inference.py:101-105) without a host-side resize."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    batch = torch.empty(1, len(images), 3, size, size, device=dev, dtype=torch.float32)
    for v, im in enumerate(images):
        preprocess_image_cuda(im, size, dev, out=batch[0, v])
    return batch
