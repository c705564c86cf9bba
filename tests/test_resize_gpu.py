"""image ingest on the GPU (siu3r_b200.io.preprocess_image_cuda -> csrc/resize.cu through the C-ABI) against the
reference recipe executed by PIL itself (siu3r_b200.io.preprocess_image = inference.py:13-38).  Byte / integer work: float32 outputs are u8 / 255
and must be IDENTICAL.  The same tables and per-sample functions are checked on the host in tests/test_io.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
SIZES = [(1296, 968), (640, 480), (480, 640), (100, 80), (61, 97), (300, 300), (206, 206), (256, 256), (256, 300), (333, 256), (513, 777), (1920, 1080)]


def _frame(W, H, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    if seed % 2:
        img[::3] = 255
        img[1::3] = 0
    return img


@pytest.mark.parametrize("size", [256, 512])
def test_ingest_matches_pil_bit_for_bit(size):
    from PIL import Image
    from siu3r_b200 import io as sio
    for n, (W, H) in enumerate(SIZES):
        img = _frame(W, H, n)
        want = sio.preprocess_image(Image.fromarray(img), size)
        got = sio.preprocess_image_cuda(img, size)
        assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == (3, size, size)
        assert torch.equal(got.cpu(), want), (W, H, size, float((got.cpu() - want).abs().max()))


def test_ingest_input_forms_batch_and_errors(tmp_path):
    from PIL import Image
    from siu3r_b200 import _lib
    from siu3r_b200 import io as sio
    a, b = _frame(640, 480, 1), _frame(480, 640, 2)
    Image.fromarray(a).save(tmp_path / "a.png")
    want_a, want_b = sio.preprocess_image(Image.fromarray(a)), sio.preprocess_image(Image.fromarray(b))
    for form in (a, Image.fromarray(a), str(tmp_path / "a.png"), torch.from_numpy(a), torch.from_numpy(a).pin_memory(), torch.from_numpy(a).to(DEV)):
        assert torch.equal(sio.preprocess_image_cuda(form).cpu(), want_a)
    batch = sio.preprocess_views_cuda([a, b])                    # the `images` tensor of inference.py:101-105
    assert tuple(batch.shape) == (1, 2, 3, 256, 256)
    assert torch.equal(batch.cpu(), torch.stack([want_a, want_b])[None])
    assert torch.equal(sio.preprocess_views_cuda([a, b]), batch)  # no state between calls
    with pytest.raises(AssertionError):
        sio.preprocess_image_cuda(torch.zeros(480, 640, 3))      # float frames are refused: the kernel resamples 8-bit samples like PIL
    with pytest.raises(AssertionError):
        sio.preprocess_image_cuda(np.zeros((480, 640), np.uint8))
    lib = _lib.load()
    d = torch.zeros(16, device=DEV, dtype=torch.int32)
    p = d.data_ptr()
    # crop window entirely outside the resized image / source rows outside the frame -> invalid argument, nothing is launched
    assert lib.siu3r_resize_lanczos_u8(p, 4, 4, 12, p, p, 1, 4, p, p, 1, 4, 9, 0, 2, 2, 0, 4, p, p, None) == -1
    assert lib.siu3r_resize_lanczos_u8(p, 4, 4, 12, p, p, 1, 4, p, p, 1, 4, 0, 0, 2, 2, 2, 4, p, p, None) == -1
