"""I/O surface of the reference's inference script: image pre-processing (inference.py:13-38), intrinsics (:107-115) and the PLY
vertex records of export_ply (src/utils/ply_export.py:30-97) packed on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_preprocess(image):
    """inference.py:13-38 restated literally (PIL calls and integer arithmetic as in the reference)."""
    from PIL import Image
    image = image.convert("RGB")
    W, H = image.size
    if W < H:
        new_W = 256
        new_H = int(H * (256 / W))
        image = image.resize((new_W, new_H), Image.Resampling.LANCZOS)
        left, top = 0, (new_H - 256) // 2
        image = image.crop((left, top, new_W, top + 256))
    else:
        new_H = 256
        new_W = int(W * (256 / H))
        image = image.resize((new_W, new_H), Image.Resampling.LANCZOS)
        left, top = (new_W - 256) // 2, 0
        image = image.crop((left, top, left + 256, new_H))
    image = np.array(image).astype(np.float32)
    return torch.from_numpy(image).permute(2, 0, 1) / 255.0


@pytest.mark.parametrize("wh", [(640, 480), (480, 640), (300, 300), (1296, 968), (257, 999)])
def test_preprocess_image_matches_reference_recipe(wh):
    from PIL import Image
    from siu3r_b200.io import default_intrinsics, preprocess_image
    rng = np.random.default_rng(wh[0])
    img = Image.fromarray(rng.integers(0, 256, size=(wh[1], wh[0], 3), dtype=np.uint8))
    got, want = preprocess_image(img), _reference_preprocess(img)
    assert got.shape == (3, 256, 256) and got.dtype == torch.float32
    assert torch.equal(got, want)
    assert float(got.min()) >= 0.0 and float(got.max()) <= 1.0
    K = default_intrinsics()
    assert K.shape == (1, 2, 3, 3) and torch.allclose(K[0, 0], torch.tensor([[318 / 256, 0, 0.5], [0, 318 / 256, 0.5], [0, 0, 1.0]]))


def test_ply_header_and_attribute_order_match_oracle():
    from oracle import ply_ref
    from siu3r_b200 import io
    rng = np.random.default_rng(0)
    G = 7
    b = ply_ref.export_ply_bytes(rng.standard_normal((G, 3)).astype("f4"), rng.random((G, 3)).astype("f4") + 0.1, rng.standard_normal((G, 4)).astype("f4"),
                                 rng.standard_normal((G, 3, 25)).astype("f4"), rng.random(G).astype("f4"), np.arange(G, dtype="i4"), np.arange(G, dtype="i4"),
                                 rng.random((G, 2, 21)).astype("f4"), save_sh_dc_only=False)
    attrs = io.ply_attributes(72, True, 42)
    hdr = io.ply_header(G, attrs)
    assert b.startswith(hdr)
    assert len(b) == len(hdr) + G * 4 * len(attrs)
    assert [n for n, _ in attrs[:9]] == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    assert [n for n, _ in attrs[81:91]] == ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3", "semantic_label", "instance_label"]
    assert io.ply_header(G, io.ply_attributes(0, True, 0)) == ply_ref.export_ply_bytes(
        rng.standard_normal((G, 3)).astype("f4"), rng.random((G, 3)).astype("f4") + 0.1, rng.standard_normal((G, 4)).astype("f4"),
        rng.standard_normal((G, 3, 25)).astype("f4"), rng.random(G).astype("f4"), np.arange(G, dtype="i4"), np.arange(G, dtype="i4"), None)[:-G * 4 * 19]


@pytest.mark.gpu
@pytest.mark.parametrize("dc_only,with_qc", [(False, True), (True, False), (True, True)])
def test_export_ply_gpu_pack_matches_oracle(tmp_path, dc_only, with_qc):
    from oracle import ply_ref
    from siu3r_b200 import io
    g = torch.Generator().manual_seed(3)
    G = 10007
    means = torch.randn(G, 3, generator=g)
    scales = torch.rand(G, 3, generator=g) * 0.3 + 1e-4
    rot = torch.randn(G, 4, generator=g)
    harm = torch.randn(G, 3, 25, generator=g)
    opac = torch.rand(G, generator=g)
    sem = torch.randint(0, 21, (G,), dtype=torch.int32, generator=g)
    inst = torch.randint(0, 9, (G,), dtype=torch.int32, generator=g)
    qc = torch.rand(G, 3, 21, generator=g) if with_qc else None
    path = io.export_ply(means.cuda(), scales.cuda(), rot.cuda(), harm.cuda(), opac.cuda(), sem.cuda(), inst.cuda(), None if qc is None else qc.cuda(),
                         tmp_path / "o.ply", save_sh_dc_only=dc_only)
    got = open(path, "rb").read()
    want = ply_ref.export_ply_bytes(means.numpy(), scales.numpy(), rot.numpy(), harm.numpy(), opac.numpy(), sem.numpy(), inst.numpy(),
                                    None if qc is None else qc.numpy(), save_sh_dc_only=dc_only)
    assert len(got) == len(want)
    hl = want.index(b"end_header\n") + len(b"end_header\n")
    assert got[:hl] == want[:hl]
    F = (len(want) - hl) // (4 * G)
    a = np.frombuffer(got[hl:], dtype="<u4").reshape(G, F)
    b = np.frombuffer(want[hl:], dtype="<u4").reshape(G, F)
    s0 = 9 + (0 if dc_only else 72) + 1          # scale_0..2: logf (CUDA, <= 1 ulp) vs np.log
    exact = np.ones(F, dtype=bool)
    exact[s0:s0 + 3] = False
    assert np.array_equal(a[:, exact], b[:, exact])
    sa, sb = a[:, s0:s0 + 3].copy().view("<f4"), b[:, s0:s0 + 3].copy().view("<f4")
    assert np.abs(sa - sb).max() <= 1e-6 * np.abs(sb).max()


@pytest.mark.gpu
def test_inference_cli_end_to_end(tmp_path):
    """inference.py flow on two synthetic JPEG-sized images with the seeded weights: file exists, header consistent with the outputs."""
    from PIL import Image
    from siu3r_b200 import inference
    rng = np.random.default_rng(5)
    p1, p2 = tmp_path / "a.png", tmp_path / "b.png"
    Image.fromarray(rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8)).save(p1)
    Image.fromarray(rng.integers(0, 256, size=(640, 480, 3), dtype=np.uint8)).save(p2)
    out = inference.main(["--image_path1", str(p1), "--image_path2", str(p2), "--output_path", str(tmp_path / "o"), "--synthetic_weights"])
    data = open(out, "rb").read()
    hl = data.index(b"end_header\n") + len(b"end_header\n")
    head = data[:hl].decode()
    assert "element vertex 131072" in head and "property int instance_label" in head
    nprops = head.count("property ")
    assert len(data) == hl + 131072 * 4 * nprops
    with pytest.raises(FileNotFoundError):
        inference.main(["--image_path1", str(tmp_path / "missing.jpg"), "--image_path2", str(p2), "--synthetic_weights"])


# ---- GPU ingest (csrc/resize.cu): tables + integer passes checked on the host against PIL -------------------------------------------------
INGEST_SIZES = [(1296, 968), (640, 480), (480, 640), (100, 80), (61, 97), (300, 300), (206, 206), (256, 256), (256, 300), (333, 256), (513, 777), (1920, 1080)]


def _test_frame(W, H, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    if seed % 2:
        img[::3] = 255                                   # hard edges: LANCZOS ringing runs into the 0 / 255 clipping
        img[1::3] = 0
    return img


@pytest.mark.parametrize("size", [256, 512])
def test_ingest_tables_and_integer_passes_match_pil_on_host(tmp_path, size):
    """siu3r_b200.io.lanczos_tables / resize_plan + the per-sample functions of csrc/resize_core.h (compiled for the host, driven with the
    kernels' index arithmetic) reproduce preprocess_image (PIL LANCZOS resize + crop + /255, inference.py:13-38) bit for bit -- including
    up-scaling, passes Pillow skips, and the square sizes whose resized side comes out one pixel short (black column from Image.crop)."""
    import ctypes
    import subprocess
    from PIL import Image
    from siu3r_b200 import io as sio
    so = str(tmp_path / "libresize_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-I", os.path.join(ROOT, "siu3r_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host_core", "resize_host.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.resize_host.argtypes = [P, I, I, ctypes.c_int64, P, P, I, I, P, P, I, I, I, I, I, I, I, I, P]
    short = False
    for n, (W, H) in enumerate(INGEST_SIZES):
        img = _test_frame(W, H, n)
        want = sio.preprocess_image(Image.fromarray(img), size).numpy()
        if size == 256:
            assert np.array_equal(want, _reference_preprocess(Image.fromarray(img)).numpy())
        p = sio.ingest_plan(W, H, size)                   # the same host plan preprocess_image_cuda hands to siu3r_resize_lanczos_u8
        short |= p["crop_x"] < 0
        assert 0 <= p["row0"] and p["row0"] + p["rows"] <= H
        got = np.empty((3, size, size), np.float32)
        rc = lib.resize_host(img.ctypes.data, H, W, W * 3, p["bounds_x"].ctypes.data, p["kx"].ctypes.data, p["ksize_x"], p["new_W"],
                             p["bounds_y"].ctypes.data, p["ky"].ctypes.data, p["ksize_y"], p["new_H"], p["crop_x"], p["crop_y"], size, size,
                             p["row0"], p["rows"], got.ctypes.data)
        assert rc == 0 and np.array_equal(got, want), (W, H, size, float(np.abs(got - want).max()))
    assert short or size != 256                          # 206 x 206 -> 255 wide at size 256: the out-of-image crop column was exercised


def test_load_checkpoint_tolerates_pickled_reference_config(tmp_path):
    """The reference's Lightning checkpoint pickles hyper_parameters["cfg"] = src.config.RootCfg (dataclasses from src.config / src.data.config,
    pipeline.py:26,39).  Without the reference tree on sys.path a plain torch.load raises ModuleNotFoundError;
    load_checkpoint must still return the "model."-stripped tensors (and drop non-model entries such as the lpips weights)."""
    import dataclasses
    import enum
    import pathlib
    import types
    from siu3r_b200.io import load_checkpoint
    mods = {n: types.ModuleType(n) for n in ("src", "src.config", "src.data", "src.data.config")}
    sys.modules.update(mods)
    try:
        @dataclasses.dataclass
        class DataCfg:
            root: pathlib.Path
            views: int = 2

        class Mode(enum.Enum):
            A = 1

        @dataclasses.dataclass
        class RootCfg:
            data: DataCfg
            mode: Mode
            name: str = "x"
        for cls, mod in ((DataCfg, "src.data.config"), (Mode, "src.config"), (RootCfg, "src.config")):
            cls.__module__, cls.__qualname__ = mod, cls.__name__
            setattr(mods[mod], cls.__name__, cls)
        ck = {"state_dict": {"model.backbone.w": torch.arange(6.0).view(2, 3), "lpips.net.x": torch.zeros(1)},
              "hyper_parameters": {"cfg": RootCfg(DataCfg(pathlib.Path("/d")), Mode.A)}, "epoch": 100}
        torch.save(ck, tmp_path / "fake.ckpt")
    finally:
        for n in mods:
            sys.modules.pop(n, None)
    with pytest.raises(ModuleNotFoundError):
        torch.load(tmp_path / "fake.ckpt", weights_only=False)
    sd = load_checkpoint(tmp_path / "fake.ckpt")
    assert list(sd) == ["backbone.w"] and torch.equal(sd["backbone.w"], torch.arange(6.0).view(2, 3))
    torch.save({"a": torch.ones(2)}, tmp_path / "plain.pt")
    assert torch.equal(load_checkpoint(tmp_path / "plain.pt")["a"], torch.ones(2))


def test_load_checkpoint_never_resolves_foreign_globals(tmp_path):
    """A checkpoint that names an arbitrary importable callable (os.system through __reduce__) must load as inert placeholders: nothing outside the
    tensor allow-list is imported or called."""
    import os
    import pickle

    from siu3r_b200.io import load_checkpoint
    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            return (os.system, (f"touch {marker}",))

    torch.save({"state_dict": {"model.backbone.x": torch.ones(3)}, "callbacks": Evil()}, tmp_path / "evil.ckpt", pickle_module=pickle)
    sd = load_checkpoint(tmp_path / "evil.ckpt")
    assert not marker.exists()
    assert torch.equal(sd["backbone.x"], torch.ones(3))
