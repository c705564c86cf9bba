"""Drop-in for the reference's `curope` extension and its wrapper module (native entry point 1):
  rope_2d(tokens, positions, base, fwd)   <-> curope.rope_2d   (croco/curope/curope.cpp:49-65 -> kernels.cu:84-108)
  cuRoPE2D(freq, F0).forward(tokens, pos) <-> croco/curope/curope2d.py:32-40 (called from croco/blocks.py:101-103,158-160)

The engine itself never calls these: inside SIU3RModel the rotation is fused into the epilogue of the qkv / q / k projections
(siu3r_gemm_tc_rope).  They exist so that reference code which still owns its attention block can switch its RoPE to this library, and so
that the entry point is tested through the interface the reference binds.  Same argument checks and messages as the reference's TORCH_CHECKs;
forward only (the autograd wrapper of curope2d.py:12-29 is training-side), float32 only, and CUDA only -- a host tensor is refused instead of
being rotated on the CPU.
"""
from __future__ import annotations

import torch

from . import _lib


def rope_2d(tokens: torch.Tensor, positions: torch.Tensor, base: float, fwd: float) -> None:
    """In place.  tokens [B, N, H, D] with stride(3) == 1 and stride(2) == D (any batch / token stride: a q or k slice of a fused qkv buffer
    qualifies); positions [B, N, 2] int64 contiguous (row, column)."""
    def check(cond, msg):
        if not cond:
            raise RuntimeError(msg)
    check(tokens.dim() == 4, "tokens must have 4 dimensions")                                   # curope.cpp:54-59
    check(positions.dim() == 3, "positions must have 3 dimensions")
    check(tokens.size(0) == positions.size(0), "batch size differs between tokens & positions")
    check(tokens.size(1) == positions.size(1), "seq_length differs between tokens & positions")
    check(positions.size(2) == 2, "positions.shape[2] must be equal to 2")
    check(tokens.is_cuda == positions.is_cuda, "tokens and positions are not on the same device")
    check(tokens.is_cuda, "siu3r_b200.curope.rope_2d has no CPU path: move tokens and positions to the GPU")
    B, N, H, D = tokens.shape
    check(tokens.stride(3) == 1 and tokens.stride(2) == D, "tokens are not contiguous")          # kernels.cu:91-94
    check(positions.is_contiguous(), "positions are not contiguous")
    check(D % 4 == 0, "token dim must be multiple of 4")
    check(tokens.dtype == torch.float32 and positions.dtype == torch.int64, "rope_2d: float32 tokens and int64 positions only")
    code = _lib.load().siu3r_rope2d(tokens.data_ptr(), positions.data_ptr(), B, N, H, D, tokens.stride(0), tokens.stride(1), float(base), float(fwd),
                                    1, 0, 0, torch.cuda.current_stream().cuda_stream)
    _lib.check(code, "rope2d")


class cuRoPE2D:
    """nn.Module-like (no parameters): forward(tokens [B, H, N, D], positions [B, N, 2]) rotates in place and returns tokens."""

    def __init__(self, freq: float = 100.0, F0: float = 1.0):
        self.base = freq
        self.F0 = F0

    def forward(self, tokens: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        rope_2d(tokens.transpose(1, 2), positions.contiguous(), self.base, self.F0)
        return tokens

    __call__ = forward
