"""siu3r_b200.curope (drop-in for the reference's curope extension + cuRoPE2D wrapper, curope.cpp:49-65 / curope2d.py:32-40), used the way
croco/blocks.py:97-103 uses it: q and k are [B, H, N, D] views of the fused qkv projection, rotated in place.  Checked against the C oracle
(oracle/raster_ref.c, pinned to the reference's compiled curope.cpp and to its PyTorch fallback in tests/test_oracle_cpu.py); float
tolerance 3e-5 abs on N(0,1) tokens (sin / cos / pow of the device vs libm), v must stay untouched."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_curope_module_on_fused_qkv_views():
    from oracle import raster_oracle as RO
    from siu3r_b200.curope import cuRoPE2D, rope_2d
    B, N, H, D = 2, 1025, 16, 64
    g = torch.Generator().manual_seed(4)
    qkv = torch.randn(B, N, 3 * H * D, generator=g)
    ys, xs = torch.meshgrid(torch.arange(32), torch.arange(32), indexing="ij")
    pos = torch.cat([torch.stack([ys.flatten(), xs.flatten()], -1), torch.tensor([[32, 0]])], 0)[None].repeat(B, 1, 1)
    dq = qkv.to(DEV)
    q, k, v = dq.reshape(B, N, 3, H, D).transpose(1, 3).unbind(2)            # blocks.py:97: [B, H, N, D] views of the projection output
    v_before = v.clone()
    rope = cuRoPE2D(100.0)
    assert rope(q, pos.to(DEV)) is q
    rope(k, pos.to(DEV))
    torch.cuda.synchronize()
    host = qkv.view(B, N, 3, H, D)
    for i, t in ((0, q), (1, k)):
        want = RO.rope2d(host[:, :, i].contiguous().numpy(), pos.numpy())     # [B, N, H, D]
        assert np.abs(t.transpose(1, 2).cpu().numpy() - want).max() < 3e-5
    assert torch.equal(v, v_before)
    # inverse rotation (the fwd = -F0 call of the reference's backward) restores the tokens
    rope_2d(q.transpose(1, 2), pos.to(DEV), 100.0, -1.0)
    assert float((q.transpose(1, 2).cpu() - host[:, :, 0]).abs().max()) < 2e-5


def test_curope_argument_checks():
    from siu3r_b200.curope import rope_2d
    tok = torch.zeros(1, 4, 2, 8, device=DEV)
    pos = torch.zeros(1, 4, 2, dtype=torch.int64, device=DEV)
    for bad_tok, bad_pos, msg in [(tok[0], pos, "4 dimensions"), (tok, pos[0], "3 dimensions"), (tok, pos[:, :3], "seq_length"),
                                  (tok, pos.cpu(), "same device"), (tok.cpu(), pos.cpu(), "no CPU path"),
                                  (torch.zeros(1, 4, 8, 2, device=DEV).transpose(2, 3), pos, "not contiguous"),
                                  (torch.zeros(1, 4, 2, 6, device=DEV), pos, "multiple of 4"), (tok.double(), pos, "float32")]:
        with pytest.raises(RuntimeError, match=msg):
            rope_2d(bad_tok, bad_pos, 100.0, 1.0)
    rope_2d(tok, pos, 100.0, 1.0)                                             # position 0 -> identity
    assert float(tok.abs().max()) == 0.0
