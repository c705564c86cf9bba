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
