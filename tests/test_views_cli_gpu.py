"""siu3r_b200.inference_multiview (mirror of the reference's inference_multiview.py:40-153) end to end on three synthetic frames with the
seeded weights, host ingest and GPU ingest: preprocess_views_cuda is bit-identical to the PIL recipe (tests/test_resize_gpu.py), so both runs
see the same tensor and must write the same file up to the run-to-run noise of the few kernels that accumulate with atomics (GroupNorm
statistics in fp64): identical headers, < 0.1 % differing bytes.  Also the `--gpu_ingest` switch of the pair script."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(tmp_path, sizes):
    from PIL import Image
    rng = np.random.default_rng(9)
    d = tmp_path / "views"
    d.mkdir()
    for i, (w, h) in enumerate(sizes):
        ext = ("jpg", "png", "jpeg")[i % 3]
        Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(d / f"v{i}.{ext}")
    return d


def _same_file(da: bytes, db: bytes):
    hl = da.index(b"end_header\n") + len(b"end_header\n")
    assert da[:hl] == db[:hl] and len(da) == len(db)
    na, nb = np.frombuffer(da[hl:], np.uint8), np.frombuffer(db[hl:], np.uint8)
    assert float((na != nb).mean()) < 1e-3


def test_inference_multiview_cli_host_and_gpu_ingest_agree(tmp_path):
    from siu3r_b200 import inference_multiview
    d = _frames(tmp_path, [(640, 480), (480, 640), (300, 300)])
    a = inference_multiview.main(["--image_dir", str(d), "--output_path", str(tmp_path / "a"), "--synthetic_weights"])
    b = inference_multiview.main(["--image_dir", str(d), "--output_path", str(tmp_path / "b"), "--synthetic_weights", "--gpu_ingest"])
    da, db = open(a, "rb").read(), open(b, "rb").read()
    hl = da.index(b"end_header\n") + len(b"end_header\n")
    head = da[:hl].decode()
    assert f"element vertex {3 * 256 * 256}" in head and "property int instance_label" in head
    assert len(da) == hl + 3 * 256 * 256 * 4 * head.count("property ")
    _same_file(da, db)
    with pytest.raises(FileNotFoundError):
        inference_multiview.main(["--image_dir", str(tmp_path / "missing"), "--synthetic_weights"])
    (tmp_path / "one").mkdir()
    with pytest.raises(AssertionError):
        inference_multiview.main(["--image_dir", str(tmp_path / "one"), "--synthetic_weights"])


def test_inference_pair_cli_gpu_ingest_agrees(tmp_path):
    from siu3r_b200 import inference
    d = _frames(tmp_path, [(640, 480), (480, 640)])
    args = ["--image_path1", str(d / "v0.jpg"), "--image_path2", str(d / "v1.png"), "--synthetic_weights"]
    a = inference.main(args + ["--output_path", str(tmp_path / "a")])
    b = inference.main(args + ["--output_path", str(tmp_path / "b"), "--gpu_ingest"])
    _same_file(open(a, "rb").read(), open(b, "rb").read())
