"""Thin torch-tensor wrappers over the C ABI (include/siu3r_b200.h).

PyTorch is only the allocator / stream owner here: every function extracts raw device pointers and leading dimensions
and calls the hand-written sm_100a kernels in libsiu3r_b200.so on torch's current stream.  No torch compute op is used.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
ELT_RELU, ELT_ADD, ELT_ADD_RELU, ELT_GELU, ELT_COPY, ELT_SIGMOID, ELT_CLAMP01, ELT_RELU_RN, ELT_ROUND, ELT_ADD_RN = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
ACT_ROUND_TF32 = 4  # flag OR-ed into `act`: store RN_tf32(result)
PREC_TF32, PREC_FP32X3, PREC_H3 = 1, 3, 4   # 4 = fp32-grade on the fp16 tensor-core path (fp16 hi / lo plane pairs, csrc/h3.cuh)
SKINNY = False   # route M <= SMALL_M GEMMs (TF32 mode) to siu3r_gemm_skinny: measured no faster than the tensor-core kernel inside the
#                  captured graph (Mask2Former stage 3.85 vs 3.5 ms) -> off; the kernel stays available through the C ABI
SMALL_M, SMALL_NK = 128, 0   # GEMMs with at most SMALL_M rows and N*K <= SMALL_NK would run on the fp32 FFMA kernel; measured slower than
#                              the tensor-core kernel on the Mask2Former query GEMMs (few CTAs, serial K loop) -> disabled (0)


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Optional per-kernel-family timing (bench.py roofline pass): PROFILE = list -> (family, work, start_event, end_event)
PROFILE = None


class _Prof:
    __slots__ = ("fam", "work", "s", "tag")

    def __init__(self, fam, work, tag=None):
        self.fam, self.work, self.tag = fam, work, tag
        self.s = None

    def __enter__(self):
        if PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *a):
        if self.s is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            PROFILE.append((self.fam, self.work, self.s, e, self.tag))


def _p(t):
    return None if t is None else t.data_ptr()


def _chk_f32(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float32, (t.device, t.dtype)


class Split:
    """fp16 plane pair of the h3 mode (csrc/h3.cuh): t = [2, *shape] half tensor, t[0] = hi = fp16(x), t[1] = lo = fp16((x - hi) * 2^11)
    (or the unscaled fp16(x - hi) when `unscaled`: operands of the attention kernel).  Slicing / views apply to the logical dims."""

    __slots__ = ("t", "unscaled")

    def __init__(self, t: torch.Tensor, unscaled: bool = False):
        assert t.dtype == torch.float16 and t.shape[0] == 2 and t.is_cuda
        self.t, self.unscaled = t, unscaled

    @staticmethod
    def empty(*shape, device, unscaled: bool = False) -> "Split":
        return Split(torch.empty(2, *shape, device=device, dtype=torch.float16), unscaled)

    shape = property(lambda self: self.t.shape[1:])
    plane = property(lambda self: self.t.stride(0))
    device = property(lambda self: self.t.device)

    def dim(self):
        return self.t.dim() - 1

    def stride(self, i):
        return self.t.stride(i + 1 if i >= 0 else i)

    def data_ptr(self):
        return self.t.data_ptr()

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        return Split(self.t[(slice(None),) + idx], self.unscaled)

    def view(self, *shape):
        return Split(self.t.view(2, *shape), self.unscaled)

    def is_contiguous(self):
        return self.t[0].is_contiguous()

    def float(self) -> torch.Tensor:
        """hi + lo * 2^-11 as fp32 (tests / debugging)."""
        assert self.dim() == 2 and self.stride(1) == 1
        y = torch.empty(self.shape[0], self.shape[1], device=self.device, dtype=torch.float32)
        lib = _lib.load()
        _lib.check(lib.siu3r_merge_h3(self.data_ptr(), self.stride(0), self.plane, self.shape[0], self.shape[1], _p(y), y.stride(0), _stream()), "merge_h3")
        if self.unscaled:   # (the merge kernel applies the 2^-11 factor: undo it exactly on the lo plane instead of on the rounded sum)
            return self.t[0].float() + self.t[1].float()
        return y


def split(x, out: "Split | None" = None, unscaled: bool = False) -> "Split":
    """fp32 [rows, cols] (row-strided) or any contiguous fp32 tensor -> plane pair of the same logical shape."""
    if isinstance(x, Split):
        return x
    _chk_f32(x)
    if x.dim() != 2 or x.stride(1) != 1:
        assert x.is_contiguous()
        o = split(x.view(-1, x.shape[-1]), None if out is None else out.view(-1, x.shape[-1]), unscaled)
        return o.view(*x.shape) if out is None else out
    if out is None:   # row pitch padded to a multiple of 8 elements (16 bytes: TMA)
        out = Split.empty(x.shape[0], (x.shape[1] + 7) // 8 * 8, device=x.device, unscaled=unscaled)[:, : x.shape[1]]
    assert out.dim() == 2 and out.stride(1) == 1 and tuple(out.shape) == tuple(x.shape) and out.unscaled == unscaled
    _lib.check(_lib.load().siu3r_split_h3(_p(x), x.stride(0), x.shape[0], x.shape[1], out.data_ptr(), out.stride(0), out.plane, 1 if unscaled else 0,
                                          _stream()), "split_h3")
    return out


def launch_count() -> int:
    return int(_lib.load().siu3r_launch_count())


def reset_launch_count():
    _lib.load().siu3r_reset_launch_count()


def split_tf32(x: torch.Tensor):
    """x (contiguous, numel % 4 == 0) -> (hi, lo) tf32-exact planes with hi + lo ~= x to ~2^-21."""
    _chk_f32(x)
    assert x.is_contiguous() and x.numel() % 4 == 0
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    _lib.check(_lib.load().siu3r_split_tf32(_p(x), _p(hi), _p(lo), x.numel(), _stream()), "split_tf32")
    return hi, lo


class Weight:
    """A [N, K] K-major matrix (nn.Linear.weight layout) prepared for the tensor-core path."""

    __slots__ = ("w", "w_lo", "bias", "N", "K", "_rows", "h3", "ln_s")

    def __init__(self, w: torch.Tensor, bias: torch.Tensor | None, precision: int):
        w = w.contiguous().float()
        self.N, self.K = w.shape
        if self.K % 4 != 0:  # TMA needs 16-byte row pitch: zero-pad K
            kp = (self.K + 3) // 4 * 4
            wp = torch.zeros(self.N, kp, device=w.device, dtype=torch.float32)
            wp[:, : self.K] = w
            w = wp
        if precision == PREC_H3:
            self.w, self.w_lo = w, None   # exact fp32 copy for the SIMT fallback; the tensor cores read the fp16 plane pair below
        elif w.is_cuda and w.numel() % 4 == 0:
            hi, lo = split_tf32(w)
            self.w = hi  # round-to-nearest TF32 once at load (the tensor core would otherwise truncate)
            self.w_lo = lo if precision == PREC_FP32X3 else None
        else:
            self.w, self.w_lo = w, None
        self.bias = None if bias is None else bias.contiguous().float()
        self._rows = None
        self.h3 = None
        self.ln_s = None
        if precision == PREC_H3 and w.is_cuda:
            kp = (w.shape[1] + 7) // 8 * 8   # 16-byte row pitch in fp16
            self.h3 = Split(torch.zeros(2, self.N, kp, device=w.device, dtype=torch.float16))
            split(w, self.h3[:, : w.shape[1]])

    def fold_ln(self, gamma: torch.Tensor, beta: torch.Tensor) -> "Weight":
        """Linear(LayerNorm(x; gamma, beta)) with the LayerNorm folded in (h3 mode, fused-LayerNorm GEMM of siu3r_gemm_h3_ln):
        weights W * gamma, bias W beta + b, and ln_s[n] = sum_k (W gamma)[n, k] taken over the fp16 planes the tensor cores actually multiply."""
        assert self.h3 is not None
        w = self.w[:, : self.K].double()
        bias = w @ beta.double() + (self.bias.double() if self.bias is not None else 0.0)
        r = Weight((w * gamma.double()[None, :]).float(), bias.float(), PREC_H3)
        planes = r.h3.t[:, :, : self.K].double()
        r.ln_s = (planes[0] + planes[1] / 2048.0).sum(1).float().contiguous()
        return r

    def rowpacked(self, KH: int, KW: int, Cin: int) -> "Weight":
        """[Cout, KH*KW*Cin] conv weight re-laid as [Cout, KH*32]: the KW*Cin <= 32 taps of one filter row become one
        32-wide channel vector (zero padded), matching the activations packed by siu3r_im2col_nhwc(KH=1)."""
        if self._rows is None:
            assert KW * Cin <= 32 and self.K == KH * KW * Cin
            r = object.__new__(Weight)
            r.N, r.K, r.bias, r._rows, r.ln_s = self.N, KH * 32, self.bias, None, None

            def relay(w):
                if w is None:
                    return None
                o = torch.zeros(self.N, KH, 32, device=w.device, dtype=torch.float32)
                o[:, :, : KW * Cin] = w[:, : self.K].view(self.N, KH, KW * Cin)
                return o.view(self.N, KH * 32)
            r.w, r.w_lo = relay(self.w), relay(self.w_lo)
            r.h3 = None
            if self.h3 is not None:
                r.h3 = Split(torch.zeros(2, self.N, KH * 32, device=self.w.device, dtype=torch.float16))
                r.h3.t.view(2, self.N, KH, 32)[..., : KW * Cin] = self.h3.t[:, :, : self.K].reshape(2, self.N, KH, KW * Cin)
            self._rows = r
        return self._rows


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest TF32 copy of x (the A operand of a TF32 GEMM whose producer could not round it itself)."""
    xc = x if x.is_contiguous() else x.contiguous()
    if xc.numel() % 4 != 0 or xc.data_ptr() % 16 != 0:
        return xc
    return eltwise(ELT_ROUND, xc)


def gemm(x: torch.Tensor, wt: Weight, out: torch.Tensor | None = None, act: int = ACT_NONE, residual: torch.Tensor | None = None,
         alpha: float = 1.0, precision: int = PREC_TF32, bias: torch.Tensor | None | bool = True, M: int | None = None,
         a_rounded: bool = False, round_out: bool = False, rope: tuple | None = None, vt: tuple | None = None, unscaled: bool = False,
         ln_stats: torch.Tensor | None = None, stats_out: torch.Tensor | None = None, out_f32: torch.Tensor | None = None):
    """out[M,N] = act(alpha * x[M,K] @ W^T + bias) + residual.  x / out / residual are 2-D row-strided views.
    rope = (positions [M,2] int64, table from rope2d_table, ncols): RoPE-2D on output columns [0, ncols) in the epilogue.
    TF32 mode: the A operand must be round-to-nearest TF32 (a_rounded=True if its producer already did that);
    round_out=True stores the result rounded because it only feeds further TF32 GEMMs."""
    if precision == PREC_H3:
        b_ = wt.bias if bias is True else (None if bias in (False, None) else bias)
        if M is not None and M != x.shape[0]:
            x = x[:M]
        return _h3_linear([x], [wt], [out], act, [residual], alpha, [b_], rope, vt, round_out, unscaled,
                          ln_stats=None if ln_stats is None else [ln_stats], stats_out=None if stats_out is None else [stats_out],
                          outs_f32=None if out_f32 is None else [out_f32])[0]
    assert ln_stats is None and stats_out is None and out_f32 is None, "fused LayerNorm exists in h3 mode only"
    _chk_f32(x, out, residual)
    assert x.dim() == 2 and x.stride(1) == 1
    M = x.shape[0] if M is None else M
    K = x.shape[1]
    assert K in (wt.K, wt.w.shape[1]), (K, wt.K)
    if out is None:
        out = torch.empty(M, wt.N, device=x.device, dtype=torch.float32)
    assert out.stride(1) == 1
    b = wt.bias if bias is True else (None if bias in (False, None) else bias)
    ldw = wt.w.shape[1]
    if M <= SMALL_M and rope is None and vt is None and precision == PREC_TF32 and SKINNY:
        # the ~85 GEMMs per pair with at most 128 rows (the 100 Mask2Former queries): set-up latency on the tensor-core kernels
        a = act | (ACT_ROUND_TF32 if round_out else 0)
        with _Prof("gemm_skinny", 2.0 * M * wt.N * K, ("skinny", M, wt.N, K)):
            code = _lib.load().siu3r_gemm_skinny(M, wt.N, K, _p(x), x.stride(0), _p(wt.w), ldw, _p(out), out.stride(0), _p(b), _p(residual),
                                                 0 if residual is None else residual.stride(0), a, alpha, _stream())
        _lib.check(code, "gemm_skinny")
        return out
    if x.stride(0) % 4 != 0 or K % 4 != 0 or x.data_ptr() % 16 != 0 or (M <= SMALL_M and wt.N * K <= SMALL_NK and rope is None and precision == PREC_TF32):
        # odd shapes (K or the row pitch not a multiple of 4 floats: no TMA) -> fp32 FFMA kernel
        a = act | (ACT_ROUND_TF32 if (round_out and precision == PREC_TF32) else 0)
        code = _lib.load().siu3r_gemm_simt(M, wt.N, K, _p(x), x.stride(0), _p(wt.w), ldw, _p(out), out.stride(0), _p(b), _p(residual),
                                           0 if residual is None else residual.stride(0), a, alpha, _stream())
        _lib.check(code, "gemm_simt")
        return out
    x_lo = None
    if precision == PREC_FP32X3:
        assert wt.w_lo is not None
        xc = x if x.is_contiguous() else x.contiguous()
        x_hi, x_lo = split_tf32(xc)
        x = x_hi
    else:
        if not a_rounded:
            x = round_tf32(x)
        if round_out:
            act = act | ACT_ROUND_TF32
    if vt is not None:
        # vt = (tensor [rows, ld], col0, state): columns >= col0 of the result go to V^T (see siu3r_gemm_tc_rope_vt); state["ok"] tells the caller
        # whether that happened (False: plain GEMM was run, use the transpose pass)
        vt_t, vt_col0, state = vt
        state["ok"] = False
        if precision == PREC_TF32 and residual is None and alpha == 1.0 and x.stride(0) % 4 == 0 and K % 4 == 0 and x.data_ptr() % 16 == 0:
            pos, tab, ncols = rope if rope is not None else (None, None, 0)
            with _Prof("gemm_tc", 2.0 * M * wt.N * K, ("vt", M, wt.N, K)):
                code = _lib.load().siu3r_gemm_tc_rope_vt(M, wt.N, K, _p(x), x.stride(0), _p(wt.w), ldw, _p(out), out.stride(0), _p(b), act, _p(pos),
                                                         _p(tab), ncols, _p(vt_t), vt_t.stride(0), vt_col0, _stream())
            if code == 0:
                state["ok"] = True
                return out
            if code != -4:
                _lib.check(code, "gemm_tc_rope_vt")
    if rope is not None:
        pos, tab, ncols = rope
        assert residual is None and alpha == 1.0 and pos.dtype == torch.int64 and pos.is_contiguous() and pos.numel() == 2 * M
        with _Prof("gemm_tc", 2.0 * M * wt.N * K, ("rope", M, wt.N, K)):
            code = _lib.load().siu3r_gemm_tc_rope(M, wt.N, K, _p(x), _p(x_lo), x.stride(0), _p(wt.w), _p(wt.w_lo), ldw, _p(out), out.stride(0),
                                                  _p(b), act, precision, _p(pos), _p(tab), ncols, _stream())
        _lib.check(code, "gemm_tc_rope")
        return out
    with _Prof("gemm_tc", 2.0 * M * wt.N * K, ("lin", M, wt.N, K)):
        code = _lib.load().siu3r_gemm_tc(M, wt.N, K, _p(x), _p(x_lo), x.stride(0), _p(wt.w), _p(wt.w_lo), ldw, _p(out), out.stride(0), _p(b),
                                         _p(residual), 0 if residual is None else residual.stride(0), act, alpha, precision, _stream())
    _lib.check(code, "gemm_tc")
    return out


def gemm_group2(xs, wts, outs=None, act: int = ACT_NONE, residuals=None, precision: int = PREC_TF32, a_rounded: bool = False,
                round_out: bool = False, rope: tuple | None = None, vt: tuple | None = None, unscaled: bool = False, ln_stats=None, stats_out=None,
                outs_f32=None):
    """Two linear layers of the same shape class (same N, K, strides, epilogue; rows may differ) in ONE persistent launch:
    outs[g] = act(xs[g] @ wts[g]^T + bias_g) + residuals[g].  Falls back to two gemm() calls when the shape / precision is not
    eligible for the grouped kernel (3xTF32 mode, tiny N or K, mismatching strides)."""
    assert len(xs) == 2 and len(wts) == 2
    if precision == PREC_H3:
        v = None
        if vt is not None:   # vt = ([window Splits], [cols], col0, state)
            v = (vt[0], vt[2], vt[3])
        return _h3_linear(list(xs), list(wts), list(outs) if outs is not None else [None, None], act, list(residuals or [None, None]), 1.0,
                          [w.bias for w in wts], rope, v, round_out, unscaled, ln_stats=ln_stats, stats_out=stats_out, outs_f32=outs_f32)
    assert ln_stats is None and stats_out is None and outs_f32 is None
    if outs is None:
        outs = [torch.empty(x.shape[0], w.N, device=x.device, dtype=torch.float32) for x, w in zip(xs, wts)]
    residuals = residuals or [None, None]

    if vt is not None:
        vt[3]["ok"] = False   # vt = ([vt0, vt1] column windows of one V^T buffer, [cols0, cols1], col0, state)

    def fallback():
        for g in range(2):
            r = None if rope is None else (rope[0].view(-1, 2)[: xs[g].shape[0]], rope[1], rope[2])   # positions repeat per image
            gemm(xs[g], wts[g], out=outs[g], act=act, residual=residuals[g], precision=precision, a_rounded=a_rounded, round_out=round_out, rope=r)
        return outs

    w0, w1 = wts
    ok = (precision == PREC_TF32 and w0.N == w1.N and w0.w.shape[1] == w1.w.shape[1] and xs[0].shape[1] == xs[1].shape[1]
          and xs[0].stride(0) == xs[1].stride(0) and outs[0].stride(0) == outs[1].stride(0) and (w0.bias is None) == (w1.bias is None)
          and (residuals[0] is None) == (residuals[1] is None) and xs[0].shape[1] % 4 == 0 and xs[0].stride(0) % 4 == 0
          and all(x.data_ptr() % 16 == 0 and x.stride(1) == 1 for x in xs) and all(o.stride(1) == 1 for o in outs))
    if ok and residuals[0] is not None:
        ok = residuals[0].stride(0) == residuals[1].stride(0)
    if not ok:
        return fallback()
    _chk_f32(*xs, *outs, *[r for r in residuals if r is not None])
    if not a_rounded:
        xs = [round_tf32(x) for x in xs]
    a = act | (ACT_ROUND_TF32 if round_out else 0)
    K = xs[0].shape[1]
    Ms = (C.c_int * 2)(xs[0].shape[0], xs[1].shape[0])
    arr = lambda ts: (C.c_void_p * 2)(*[t.data_ptr() for t in ts])
    pos = tab = None
    ncols = 0
    if rope is not None:
        pos, tab, ncols = rope[0], rope[1], rope[2]
        assert residuals[0] is None and pos.dtype == torch.int64 and pos.is_contiguous() and pos.numel() >= 2 * max(Ms[0], Ms[1])
    work = 2.0 * (Ms[0] + Ms[1]) * w0.N * K
    vts = vcols = None
    vt_ld = vt_col0 = 0
    if vt is not None:
        vts, vcols, vt_ld, vt_col0 = arr(vt[0]), (C.c_int * 2)(*vt[1]), vt[0][0].stride(0), vt[2]
    with _Prof("gemm_tc", work, ("grp2", Ms[0] + Ms[1], w0.N, K)):
        code = _lib.load().siu3r_gemm_tc_group2(Ms, w0.N, K, arr(xs), xs[0].stride(0), arr([w0.w, w1.w]), w0.w.shape[1], arr(outs), outs[0].stride(0),
                                                None if w0.bias is None else arr([w0.bias, w1.bias]),
                                                None if residuals[0] is None else arr(residuals),
                                                0 if residuals[0] is None else residuals[0].stride(0), a, 1.0, _p(pos), _p(tab), ncols, vts, vcols,
                                                vt_ld, vt_col0, _stream())
    if code == -4:
        return fallback()
    _lib.check(code, "gemm_tc_group2")
    if vt is not None:
        vt[3]["ok"] = True
    return outs


def _ptr_arr(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _h3_linear(xs, wts, outs, act, residuals, alpha, biases, rope, vt, split_out, unscaled, ln_stats=None, stats_out=None, outs_f32=None,
               ln_eps=1e-6):
    """h3 mode: outs[g] = act(alpha * xs[g] @ wts[g]^T + bias_g) [RoPE] + residuals[g] for 1 or 2 same-shape problems in one launch
    (siu3r_gemm_h3).  xs: Split plane pairs or fp32 tensors (split here, one extra pass); outs[g]: fp32 tensor, Split, or None (allocated:
    Split if split_out else fp32).  vt = (Split [rows, ld] or [window Splits], col0, state): columns >= col0 go to V^T.
    Fused LayerNorm (siu3r_gemm_h3_ln): ln_stats[g] = int64 [M_g, 2] row statistics of the RAW rows xs[g] and wts[g] = Weight.fold_ln(...) -> the
    launch computes Linear(LayerNorm(x));  stats_out[g] = int64 [M_g, 2] (zeroed) accumulates the statistics of the rows written;  outs_f32[g]: with
    Split outs, an additional fp32 copy of the result (the residual stream)."""
    G = len(xs)
    lib = _lib.load()
    Ms = [x.shape[0] for x in xs]
    K = xs[0].shape[1]
    N = wts[0].N
    assert all(x.dim() == 2 and x.shape[1] == K for x in xs) and all(w.N == N and w.K == K and w.h3 is not None for w in wts), (K, N, [w.K for w in wts])
    if not all(isinstance(x, Split) for x in xs):
        kp = (K + 7) // 8 * 8
        parent = Split.empty(sum(Ms), kp, device=xs[0].device)
        r0, sl = 0, []
        for x, m in zip(xs, Ms):
            dst = parent[r0:r0 + m, :K]
            if isinstance(x, Split):   # mixed: re-pack under the common parent (rare)
                dst.t.copy_(x.t)
            else:
                assert x.stride(1) == 1
                split(x, dst)
            sl.append(dst)
            r0 += m
        xs = sl
    assert all(not x.unscaled and x.stride(1) == 1 and x.stride(0) % 8 == 0 and x.plane % 8 == 0 and x.data_ptr() % 16 == 0 for x in xs)
    assert len({(x.stride(0), x.plane) for x in xs}) == 1, "grouped h3 operands must share pitch and plane distance (slices of one buffer)"
    assert len({(w.h3.stride(0), w.h3.plane) for w in wts}) == 1
    want_split = split_out or any(isinstance(o, Split) for o in outs)
    dev = xs[0].device
    if all(o is None for o in outs):
        if want_split:
            parent = Split.empty(sum(Ms), N, device=dev, unscaled=unscaled)
            r0, outs = 0, []
            for m in Ms:
                outs.append(parent[r0:r0 + m])
                r0 += m
        else:
            outs = [torch.empty(m, N, device=dev, dtype=torch.float32) for m in Ms]
    assert all(o is not None for o in outs)
    is_split = isinstance(outs[0], Split)
    assert all(isinstance(o, Split) == is_split for o in outs) and (is_split or not want_split)
    Cf = Ch = None
    ldc = ldh = hpl = 0
    if is_split:
        assert all(o.stride(1) == 1 and o.unscaled == unscaled for o in outs) and len({(o.stride(0), o.plane) for o in outs}) == 1
        Ch, ldh, hpl = _ptr_arr(outs), outs[0].stride(0), outs[0].plane
        if outs_f32 is not None:
            _chk_f32(*outs_f32)
            assert vt is None and all(o.stride(1) == 1 and o.shape[0] == m for o, m in zip(outs_f32, Ms)) and len({o.stride(0) for o in outs_f32}) == 1
            Cf, ldc = _ptr_arr(outs_f32), outs_f32[0].stride(0)
    else:
        assert outs_f32 is None
        _chk_f32(*outs)
        assert all(o.stride(1) == 1 for o in outs) and len({o.stride(0) for o in outs}) == 1
        Cf, ldc = _ptr_arr(outs), outs[0].stride(0)
    has_b = biases[0] is not None
    assert all((b is not None) == has_b for b in biases)
    has_r = residuals[0] is not None
    assert all((r is not None) == has_r for r in residuals)
    if has_r:
        _chk_f32(*residuals)
        assert len({r.stride(0) for r in residuals}) == 1 and all(r.stride(1) == 1 for r in residuals)
    pos = tab = None
    ncols = 0
    if rope is not None:
        pos, tab, ncols = rope
        assert not has_r and pos.dtype == torch.int64 and pos.is_contiguous() and pos.numel() >= 2 * max(Ms)
    vts = vcols = None
    vt_ld = vt_pl = vt_col0 = 0
    if vt is not None:
        vt_t, vt_col0, state = vt
        wins = vt_t if isinstance(vt_t, (list, tuple)) else [vt_t]
        assert len(wins) == G and all(isinstance(w_, Split) and w_.unscaled == unscaled and w_.stride(1) == 1 for w_ in wins)
        assert len({(w_.stride(0), w_.plane) for w_ in wins}) == 1
        vts, vcols = _ptr_arr(wins), (C.c_int * G)(*[w_.shape[1] for w_ in wins])
        vt_ld, vt_pl = wins[0].stride(0), wins[0].plane
        state["ok"] = True
    Mh = (C.c_int * G)(*Ms)
    st_in = ln_s = st_out = None
    if ln_stats is not None:
        assert alpha == 1.0 and all(w.ln_s is not None for w in wts), "fused LayerNorm needs Weight.fold_ln() weights"
        assert all(s_.dtype == torch.int64 and s_.is_contiguous() and s_.shape == (m, 2) for s_, m in zip(ln_stats, Ms))
        st_in, ln_s = _ptr_arr(ln_stats), _ptr_arr([w.ln_s for w in wts])
    if stats_out is not None:
        assert all(s_.dtype == torch.int64 and s_.is_contiguous() and s_.shape == (m, 2) for s_, m in zip(stats_out, Ms))
        st_out = _ptr_arr(stats_out)
    with _Prof("gemm_h3", 2.0 * sum(Ms) * N * K, ("h3", sum(Ms), N, K)):
        code = lib.siu3r_gemm_h3_ln(G, Mh, N, K, _ptr_arr(xs), xs[0].stride(0), xs[0].plane, _ptr_arr([w.h3 for w in wts]), wts[0].h3.stride(0),
                                    wts[0].h3.plane, Cf, ldc, Ch, ldh, hpl, _ptr_arr(biases) if has_b else None,
                                    _ptr_arr(residuals) if has_r else None, residuals[0].stride(0) if has_r else 0, act & 3, alpha, _p(pos), _p(tab),
                                    ncols, vts, vcols, vt_ld, vt_pl, vt_col0, 1 if unscaled else 0, st_in, ln_s, ln_eps, st_out, _stream())
    _lib.check(code, "gemm_h3")
    return outs


def rope2d_table(maxpos: int, D: int = 64, base: float = 100.0, fwd: float = 1.0, device="cuda") -> torch.Tensor:
    """[maxpos, D/4, 2] (cos, sin) factors shared by every RoPE-fused projection (siu3r_gemm_tc_rope)."""
    tab = torch.empty(maxpos, D // 4, 2, device=device, dtype=torch.float32)
    _lib.check(_lib.load().siu3r_rope2d_table(_p(tab), maxpos, D, base, fwd, _stream()), "rope2d_table")
    return tab


def gemm_simt(x, w, bias=None, out=None, act=ACT_NONE, residual=None, alpha=1.0):
    """Plain fp32 FFMA GEMM (reference-grade; any shape).  w: [N, K] tensor."""
    _chk_f32(x, w, out, residual)
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    code = _lib.load().siu3r_gemm_simt(M, N, K, _p(x), x.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), _p(bias), _p(residual),
                                       0 if residual is None else residual.stride(0), act, alpha, _stream())
    _lib.check(code, "gemm_simt")
    return out


def conv2d_tc_supported(H: int, W: int, Cin: int) -> bool:
    return Cin % 32 == 0 and W % 16 == 0 and H % 8 == 0


def conv2d(x: torch.Tensor, wt: Weight, KH: int, KW: int, stride: int = 1, pad: int | tuple = 0, act: int = ACT_NONE,
           residual: torch.Tensor | None = None, out: torch.Tensor | None = None, precision: int = PREC_TF32, a_rounded: bool = False,
           round_out: bool = False) -> torch.Tensor:
    """NHWC convolution.  wt is [Cout, KH*KW*Cin] ((kh, kw, ci) fastest = ci).  Stride-1 convs with tensor-core friendly
    shapes run as implicit GEMM (4-D TMA); everything else as im2col + tensor-core GEMM."""
    if precision == PREC_H3:
        return _h3_conv2d(x, wt, KH, KW, stride, pad, act, residual, out, round_out)
    _chk_f32(x, residual, out)
    N, H, W, Cin = x.shape
    assert x.is_contiguous()
    Cout = wt.N
    pad_h, pad_w = pad if isinstance(pad, tuple) else (pad, pad)
    OH = (H + 2 * pad_h - KH) // stride + 1
    OW = (W + 2 * pad_w - KW) // stride + 1
    if out is None:
        out = torch.empty(N, OH, OW, Cout, device=x.device, dtype=torch.float32)
    if KH == 1 and KW == 1 and stride == 1 and pad_h == 0 and pad_w == 0:
        gemm(x.view(-1, Cin), wt, out=out.view(-1, Cout), act=act, residual=None if residual is None else residual.view(-1, Cout),
             precision=precision, a_rounded=a_rounded, round_out=round_out)
        return out
    if stride == 1 and KH > 1 and KW > 1 and KW * Cin <= 32 and conv2d_tc_supported(H, W, 32) and OH == H and OW == W:
        # few input channels (the RGB image): pack the KW taps of a filter row into one 32-wide vector per pixel (a 1 x KW
        # im2col, 32 floats per pixel instead of KH*KW*Cin) and run the KH x 1 remainder as implicit GEMM on the tensor cores
        rows = torch.empty(N, H, W, 32, device=x.device, dtype=torch.float32)
        rnd = 1 if precision == PREC_TF32 else 0
        _lib.check(_lib.load().siu3r_im2col_nhwc(_p(x), N, H, W, Cin, 1, KW, 1, 0, pad_w, _p(rows), 32, rnd, _stream()), "im2col(rows)")
        return conv2d(rows, wt.rowpacked(KH, KW, Cin), KH, 1, 1, (pad_h, 0), act, residual, out, precision, True, round_out)
    if stride == 1 and conv2d_tc_supported(H, W, Cin) and OH == H and OW == W:
        x_lo = None
        xx = x
        if precision == PREC_FP32X3:
            xx, x_lo = split_tf32(x)
        else:
            if not a_rounded:
                xx = round_tf32(x)
            if round_out:
                act = act | ACT_ROUND_TF32
        with _Prof("conv2d_tc", 2.0 * N * H * W * Cout * KH * KW * Cin, ("conv", N * H * W, Cout, KH * KW * Cin)):
            code = _lib.load().siu3r_conv2d_tc(N, H, W, Cin, Cout, KH, KW, pad_h, pad_w, _p(xx), _p(x_lo), _p(wt.w), _p(wt.w_lo), _p(out), Cout,
                                               _p(wt.bias), _p(residual), Cout, act, precision, _stream())
        _lib.check(code, "conv2d_tc")
        return out
    K = KH * KW * Cin
    ldo = wt.w.shape[1]
    cols = torch.empty(N * OH * OW, ldo, device=x.device, dtype=torch.float32)
    rnd = 1 if precision == PREC_TF32 else 0  # the column matrix only feeds the GEMM: round it on the way out
    code = _lib.load().siu3r_im2col_nhwc(_p(x), N, H, W, Cin, KH, KW, stride, pad_h, pad_w, _p(cols), ldo, rnd, _stream())
    _lib.check(code, "im2col")
    assert K <= ldo
    gemm(cols, wt, out=out.view(-1, Cout), act=act, residual=None if residual is None else residual.view(-1, Cout), precision=precision,
         a_rounded=True, round_out=round_out)
    return out


def im2col_h3(x: torch.Tensor, KH: int, KW: int, stride: int, pad_h: int, pad_w: int, ldo: int) -> Split:
    """NHWC fp32 image -> column matrix [(n, oh, ow), ldo] as a plane pair ((kh, kw, ci) order, zero pad columns)."""
    _chk_f32(x)
    N, H, W, Cin = x.shape
    assert x.is_contiguous() and ldo % 8 == 0
    OH, OW = (H + 2 * pad_h - KH) // stride + 1, (W + 2 * pad_w - KW) // stride + 1
    cols = Split.empty(N * OH * OW, ldo, device=x.device)
    _lib.check(_lib.load().siu3r_im2col_nhwc_h3(_p(x), N, H, W, Cin, KH, KW, stride, pad_h, pad_w, cols.data_ptr(), ldo, cols.plane, _stream()),
               "im2col_h3")
    return cols


def _h3_conv2d(x, wt: Weight, KH, KW, stride, pad, act, residual, out, split_out):
    """h3-mode NHWC convolution.  x: fp32 [N,H,W,Cin] or Split of that shape; out: fp32 tensor / Split / None."""
    N, H, W, Cin = x.shape
    Cout = wt.N
    pad_h, pad_w = pad if isinstance(pad, tuple) else (pad, pad)
    OH = (H + 2 * pad_h - KH) // stride + 1
    OW = (W + 2 * pad_w - KW) // stride + 1
    want_split = split_out or isinstance(out, Split)
    if out is None:
        out = Split.empty(N, OH, OW, Cout, device=x.device) if want_split else torch.empty(N, OH, OW, Cout, device=x.device, dtype=torch.float32)
    o2 = out.view(-1, Cout)
    r2 = None if residual is None else residual.view(-1, Cout)
    if KH == 1 and KW == 1 and stride == 1 and pad_h == 0 and pad_w == 0:
        assert x.is_contiguous()
        _h3_linear([x.view(-1, Cin)], [wt], [o2], act, [r2], 1.0, [wt.bias], None, None, want_split, False)
        return out
    same = stride == 1 and OH == H and OW == W and 2 * pad_h == KH - 1 and 2 * pad_w == KW - 1
    if same and KH > 1 and KW > 1 and KW * Cin <= 32 and W % 16 == 0 and not isinstance(x, Split):
        # few input channels (the RGB image): pack the KW taps of a filter row into one 32-wide vector per pixel and run the KH x 1 remainder
        # as implicit GEMM (the 64-wide channel block is half zero fill)
        rows = im2col_h3(x, 1, KW, 1, 0, pad_w, 32).view(N, H, W, 32)
        return _h3_conv2d(rows, wt.rowpacked(KH, KW, Cin), KH, 1, 1, (pad_h, 0), act, residual, out, split_out)
    if same and W % 16 == 0 and Cin % 8 == 0:
        xs = x if isinstance(x, Split) else split(x)
        assert xs.is_contiguous() and xs.data_ptr() % 16 == 0 and xs.plane % 8 == 0 and wt.h3 is not None and wt.K == KH * KW * Cin
        if residual is not None:
            _chk_f32(residual)
            assert residual.is_contiguous()
        is_split = isinstance(out, Split)
        assert out.is_contiguous()
        with _Prof("conv2d_h3", 2.0 * N * H * W * Cout * KH * KW * Cin, ("conv", N * H * W, Cout, KH * KW * Cin)):
            code = _lib.load().siu3r_conv2d_h3(N, H, W, Cin, Cout, KH, KW, pad_h, pad_w, xs.data_ptr(), Cin, xs.plane, wt.h3.data_ptr(), wt.h3.stride(0),
                                               wt.h3.plane, None if is_split else _p(out), Cout, out.data_ptr() if is_split else None, Cout,
                                               out.plane if is_split else 0, _p(wt.bias), _p(residual), Cout, act & 3, _stream())
        _lib.check(code, "conv2d_h3")
        return out
    # strided / odd shapes: im2col (plane pair written directly) + tensor-core GEMM; im2col reads fp32
    if isinstance(x, Split):
        x = x.view(-1, Cin).float().view(N, H, W, Cin)
    cols = im2col_h3(x, KH, KW, stride, pad_h, pad_w, wt.h3.shape[1])
    _h3_linear([cols[:, : wt.K]], [wt], [o2], act, [r2], 1.0, [wt.bias], None, None, want_split, False)
    return out


def conv_kxk_up2x(x: torch.Tensor, wt: Weight, KH: int, KW: int, low: torch.Tensor, act: int = ACT_NONE, round_out: bool = False,
                  out: torch.Tensor | None = None):
    """out = act(conv_{KH x KW, stride 1, same padding}(x) + bias) + bilinear_x2_align_corners(low), TF32, for a few-channel image x
    [N,H,W,Cin] with KW*Cin <= 32 (the RGB(+pad) image of the Gaussian-parameter head's input_merger) and low [N,H/2,W/2,Cout]:
    the horizontal taps are packed per pixel (1 x KW im2col, 32 floats), the vertical taps + the upsampled residual run in one persistent
    tensor-core launch per image.  Returns None when the shape is not eligible (caller uses resize_bilinear + conv2d)."""
    N, H, W, Cin = x.shape
    Cout = wt.N
    if KW * Cin > 32 or H % 2 or W % 2 or tuple(low.shape) != (N, H // 2, W // 2, Cout) or not low.is_contiguous():
        return None
    lib = _lib.load()
    rows = torch.empty(N, H, W, 32, device=x.device, dtype=torch.float32)
    _lib.check(lib.siu3r_im2col_nhwc(_p(x), N, H, W, Cin, 1, KW, 1, 0, KW // 2, _p(rows), 32, 1, _stream()), "im2col(rows)")
    wr = wt.rowpacked(KH, KW, Cin)
    if out is None:
        out = torch.empty(N, H, W, Cout, device=x.device, dtype=torch.float32)
    a = act | (ACT_ROUND_TF32 if round_out else 0)
    for n in range(N):
        with _Prof("conv2d_tc", 2.0 * H * W * Cout * KH * KW * Cin, ("rows_up2x", H * W, Cout, KH * 32)):
            code = lib.siu3r_conv_rows_up2x_tc(H, W, KH, KH // 2, Cout, _p(rows[n]), _p(wr.w), _p(wr.bias), _p(low[n]), _p(out[n]), Cout, a, _stream())
        if code == -4:
            return None
        _lib.check(code, "conv_rows_up2x_tc")
    return out


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, out: torch.Tensor | None = None, add: torch.Tensor | None = None,
              round_out: bool = False):
    _chk_f32(x, w, b, out, add)
    assert x.dim() == 2 and x.stride(1) == 1
    rows, Cc = x.shape
    if out is None:
        out = torch.empty(rows, Cc, device=x.device, dtype=torch.float32)
    code = _lib.load().siu3r_layernorm(_p(x), x.stride(0), _p(w), _p(b), _p(out), out.stride(0), rows, Cc, eps, _p(add),
                                       0 if add is None else add.stride(0), 1 if round_out else 0, _stream())
    _lib.check(code, "layernorm")
    return out


def layernorm_group2(xs, wbs, eps: float, outs, round_out: bool = False):
    """outs[g] = LayerNorm(xs[g]; wbs[g]) for two row blocks with different affine parameters in one launch."""
    _chk_f32(*xs, *outs, *wbs[0], *wbs[1])
    Cc = xs[0].shape[1]
    assert xs[1].shape[1] == Cc and xs[0].stride(0) == xs[1].stride(0) and outs[0].stride(0) == outs[1].stride(0)
    assert all(t.stride(1) == 1 for t in (*xs, *outs))
    code = _lib.load().siu3r_layernorm_group2(_p(xs[0]), _p(xs[1]), xs[0].stride(0), _p(wbs[0][0]), _p(wbs[0][1]), _p(wbs[1][0]), _p(wbs[1][1]),
                                              _p(outs[0]), _p(outs[1]), outs[0].stride(0), xs[0].shape[0], xs[1].shape[0], Cc, eps,
                                              1 if round_out else 0, _stream())
    _lib.check(code, "layernorm_group2")
    return outs


def layernorm_h3(xs, wbs, eps: float, outs=None, outs_f32=None):
    """h3 mode LayerNorm of one or two row blocks (different affine parameters) in one launch, result as plane pairs `outs` (allocated as
    slices of one buffer when None) and / or fp32 `outs_f32`.  Returns the list of plane pairs (or of fp32 tensors when only those)."""
    G = len(xs)
    assert G in (1, 2) and len(wbs) == G
    _chk_f32(*xs, *[t for wb in wbs for t in wb])
    Cc = xs[0].shape[1]
    assert all(x.dim() == 2 and x.shape[1] == Cc and x.stride(1) == 1 for x in xs) and len({x.stride(0) for x in xs}) == 1
    rows = [x.shape[0] for x in xs]
    if outs is None and outs_f32 is None:
        parent = Split.empty(sum(rows), Cc, device=xs[0].device)
        outs, r0 = [], 0
        for r in rows:
            outs.append(parent[r0:r0 + r])
            r0 += r
    ldy = ldh = pl = 0
    if outs is not None:
        assert all(isinstance(o, Split) and not o.unscaled and o.stride(1) == 1 for o in outs) and len({(o.stride(0), o.plane) for o in outs}) == 1
        ldh, pl = outs[0].stride(0), outs[0].plane
    if outs_f32 is not None:
        _chk_f32(*outs_f32)
        assert len({o.stride(0) for o in outs_f32}) == 1
        ldy = outs_f32[0].stride(0)
    g1 = G == 2
    code = _lib.load().siu3r_layernorm_h3(_p(xs[0]), _p(xs[1]) if g1 else None, xs[0].stride(0), _p(wbs[0][0]), _p(wbs[0][1]),
                                          _p(wbs[1][0]) if g1 else None, _p(wbs[1][1]) if g1 else None,
                                          _p(outs_f32[0]) if outs_f32 is not None else None, _p(outs_f32[1]) if (outs_f32 is not None and g1) else None,
                                          ldy, outs[0].data_ptr() if outs is not None else None, outs[1].data_ptr() if (outs is not None and g1) else None,
                                          ldh, pl, rows[0], rows[1] if g1 else 0, Cc, eps, _stream())
    _lib.check(code, "layernorm_h3")
    return outs if outs is not None else outs_f32


def flash_attn_h3(q: Split, q_col0: int, k: Split, k_col0: int, vt: Split, vt_batch_cols: int, B: int, H: int, Nq: int, Nk: int, scale: float,
                  out=None, split_out: bool = True, vt_b_split: int = 0, vt_extra: int = 0):
    """tcgen05 flash attention at fp32-grade accuracy (csrc/flash_h3.cu).  q [B*Nq, wq] / k [B*Nk, wk]: UNSCALED plane pairs, head h at columns
    col0 + 64 h; vt: unscaled V^T plane pair [H*64, ld] with image b at column b * vt_batch_cols (+ vt_extra for b >= vt_b_split > 0), or
    [(B*H)*64, ld] when vt_batch_cols == 0.  out: Split / fp32 [B*Nq, H*64] (allocated when None)."""
    assert q.unscaled and k.unscaled and vt.unscaled and q.dim() == 2 and k.dim() == 2 and vt.dim() == 2
    assert q.stride(1) == 1 and k.stride(1) == 1 and vt.stride(1) == 1
    dev = q.device
    if out is None:
        out = Split.empty(B * Nq, H * 64, device=dev) if split_out else torch.empty(B * Nq, H * 64, device=dev, dtype=torch.float32)
    is_split = isinstance(out, Split)
    assert out.stride(1) == 1 and tuple(out.shape) == (B * Nq, H * 64)
    with _Prof("flash_attn", 4.0 * B * H * Nq * Nk * 64):
        code = _lib.load().siu3r_flash_attn_h3(q.data_ptr(), Nq * q.stride(0), q.stride(0), q.plane, q.shape[1], q_col0, k.data_ptr(), Nk * k.stride(0),
                                               k.stride(0), k.plane, k.shape[1], k_col0, vt.data_ptr(), vt.stride(0), vt.plane, vt_batch_cols,
                                               vt_b_split, vt_extra, None if is_split else _p(out), Nq * out.stride(0), out.stride(0),
                                               out.data_ptr() if is_split else None, Nq * out.stride(0), out.stride(0), out.plane if is_split else 0,
                                               B, H, Nq, Nk, scale, _stream())
    _lib.check(code, "flash_attn_h3")
    return out


def transpose_v_h3(v: torch.Tensor, v_off: int, v_bs: int, v_ts: int, B: int, N: int, H: int) -> Split:
    """fp32 V inside a fused buffer -> unscaled V^T plane pair [(B*H)*64, roundup8(N)] (fallback when the projection could not write V^T)."""
    ld = (N + 7) // 8 * 8
    vt = Split.empty(B * H * 64, ld, device=v.device, unscaled=True)
    _lib.check(_lib.load().siu3r_transpose_v_h3(v.data_ptr() + 4 * v_off, v_bs, v_ts, B, N, H, vt.data_ptr(), ld, vt.plane, _stream()), "transpose_v_h3")
    return vt


def rope2d_(tokens_ptr_tensor: torch.Tensor, offset: int, positions: torch.Tensor, B: int, N: int, H: int, D: int, batch_stride: int,
            token_stride: int, base: float = 100.0, fwd: float = 1.0, nparts: int = 1, part_stride: int = 0, round_out: bool = False):
    """In-place 2-D RoPE on tokens[b,n,h,d] located at tokens.data_ptr() + 4*(offset + b*batch_stride + n*token_stride + h*D + d)."""
    assert positions.dtype == torch.int64 and positions.is_cuda and positions.is_contiguous()
    code = _lib.load().siu3r_rope2d(tokens_ptr_tensor.data_ptr() + 4 * offset, _p(positions), B, N, H, D, batch_stride, token_stride, base, fwd,
                                    nparts, part_stride, 1 if round_out else 0, _stream())
    _lib.check(code, "rope2d")


def flash_attn_d64(q: torch.Tensor, q_off: int, q_bs: int, q_ts: int, k: torch.Tensor, k_off: int, k_bs: int, k_ts: int, v: torch.Tensor,
                   v_off: int, v_bs: int, v_ts: int, out: torch.Tensor, B: int, H: int, Nq: int, Nk: int, scale: float, precision: int,
                   round_out: bool = False):
    """out [B, Nq, H*64] contiguous."""
    with _Prof("flash_attn", 4.0 * B * H * Nq * Nk * 64):
        code = _lib.load().siu3r_flash_attn_d64(q.data_ptr() + 4 * q_off, q_bs, q_ts, k.data_ptr() + 4 * k_off, k_bs, k_ts, v.data_ptr() + 4 * v_off,
                                                v_bs, v_ts, _p(out), Nq * H * 64, H * 64, B, H, Nq, Nk, scale, precision, 1 if round_out else 0, _stream())
    _lib.check(code, "flash_attn_d64")
    return out


def flash_attn_tc(q: torch.Tensor, q_off: int, q_bs: int, q_ts: int, q_width: int, k: torch.Tensor, k_off: int, k_bs: int, k_ts: int, k_width: int,
                  v: torch.Tensor, v_off: int, v_bs: int, v_ts: int, out: torch.Tensor, B: int, H: int, Nq: int, Nk: int, scale: float,
                  round_out: bool = False, vt: tuple | None = None):
    """tcgen05 flash attention (TF32).  q/k: element (b, n, h, d) at tensor.data_ptr() + 4*(b*bs + n*ts + off + h*64 + d), `width` = row
    width in floats; v likewise (transposed + rounded internally), unless vt = (V^T tensor [H*64, ld], image column stride) written by the
    projection GEMM itself is given.  out [B, Nq, H*64] contiguous."""
    lib = _lib.load()
    if vt is None:
        ld, bcols = (Nk + 3) // 4 * 4, 0
        vt_t = torch.empty(B * H * 64, ld, device=out.device, dtype=torch.float32)
        _lib.check(lib.siu3r_transpose_v(v.data_ptr() + 4 * v_off, v_bs, v_ts, B, Nk, H, _p(vt_t), ld, _stream()), "transpose_v")
    else:
        vt_t, bcols = vt
        ld = vt_t.stride(0)
    with _Prof("flash_attn", 4.0 * B * H * Nq * Nk * 64):
        code = lib.siu3r_flash_attn_tc(q.data_ptr(), q_bs, q_ts, q_width, q_off, k.data_ptr(), k_bs, k_ts, k_width, k_off, _p(vt_t), ld, bcols, _p(out),
                                       Nq * H * 64, H * 64, B, H, Nq, Nk, scale, 1 if round_out else 0, _stream())
    _lib.check(code, "flash_attn_tc")
    return out


def attn_small_d32(q, q_bs, q_ts, k, k_bs, k_ts, v, v_bs, v_ts, out, o_bs, o_ts, mask, B, H, Nq, Nk, scale, round_out=False):
    lib = _lib.load()
    nb = lib.siu3r_attn_small_d32_ws_bytes(B, H, Nq, Nk)
    ws = torch.empty(nb, dtype=torch.uint8, device=q.device)
    code = lib.siu3r_attn_small_d32(_p(q), q_bs, q_ts, _p(k), k_bs, k_ts, _p(v), v_bs, v_ts, _p(out), o_bs, o_ts, _p(mask), B, H, Nq, Nk,
                                    scale, 1 if round_out else 0, _p(ws), nb, _stream())
    _lib.check(code, "attn_small_d32")
    return out


def msdeform_attn(value, Lin, ow, ref, levels_hw, P, B, Lq, nH, hd, out, round_out=False):
    L = len(levels_hw)
    arr = (C.c_int * (2 * L))(*[int(v) for hw in levels_hw for v in hw])
    code = _lib.load().siu3r_msdeform_attn(_p(value), value.stride(-2) if value.dim() > 1 else nH * hd, Lin, _p(ow), ow.stride(-2), _p(ref), arr, L,
                                           P, B, Lq, nH, hd, _p(out), out.stride(-2), 1 if round_out else 0, _stream())
    _lib.check(code, "msdeform_attn")
    return out


def eltwise(op: int, a: torch.Tensor, b: torch.Tensor | None = None, out: torch.Tensor | None = None):
    _chk_f32(a, b, out)
    assert a.is_contiguous() and (b is None or b.is_contiguous())
    if out is None:
        out = torch.empty_like(a)
    _lib.check(_lib.load().siu3r_eltwise(op, _p(a), _p(b), _p(out), a.numel(), _stream()), "eltwise")
    return out


def eltwise_h3(op: int, a: torch.Tensor, b: torch.Tensor | None = None) -> Split:
    """eltwise (ops 0..6) whose result is stored as a plane pair of a's shape."""
    _chk_f32(a, b)
    assert a.is_contiguous() and (b is None or b.is_contiguous()) and a.numel() % 4 == 0
    out = Split.empty(*a.shape, device=a.device)
    _lib.check(_lib.load().siu3r_eltwise_h3(op, _p(a), _p(b), out.data_ptr(), out.plane, a.numel(), _stream()), "eltwise_h3")
    return out


def scale_(x: torch.Tensor, alpha: float):
    """x *= alpha in place."""
    _chk_f32(x)
    assert x.is_contiguous()
    _lib.check(_lib.load().siu3r_scale(_p(x), float(alpha), _p(x), x.numel(), _stream()), "scale")
    return x


def rows_affine(x, scale=None, shift=None, add=None, out=None, relu=False, rows=None, C_=None):
    """out[r,:C] = relu?(x[r,:C]*scale + shift + add[r,:C]) for 2-D row-strided views."""
    _chk_f32(x, scale, shift, add, out)
    rows = x.shape[0] if rows is None else rows
    Cc = x.shape[1] if C_ is None else C_
    if out is None:
        out = torch.empty(rows, Cc, device=x.device, dtype=torch.float32)
    code = _lib.load().siu3r_rows_affine(_p(x), x.stride(0), _p(scale), _p(shift), _p(add), 0 if add is None else add.stride(0), _p(out),
                                         out.stride(0), rows, Cc, 1 if relu else 0, _stream())
    _lib.check(code, "rows_affine")
    return out


def resize_bilinear(x: torch.Tensor, OH: int, OW: int, align_corners: bool, out: torch.Tensor | None = None, accumulate: bool = False,
                    ldx: int | None = None, round_out: bool = False):
    """x: [N,H,W,C] (pixel stride ldx, images densely packed H*W*ldx apart) -> out [N,OH,OW,C]."""
    _chk_f32(x, out)
    N, H, W, Cc = x.shape
    ldx = x.stride(2) if ldx is None else ldx
    if out is None:
        out = torch.empty(N, OH, OW, Cc, device=x.device, dtype=torch.float32)
    code = _lib.load().siu3r_resize_bilinear_nhwc(_p(x), N, H, W, Cc, ldx, _p(out), OH, OW, out.stride(2), 1 if align_corners else 0,
                                                  (1 if accumulate else 0) | (2 if round_out else 0), _stream())
    _lib.check(code, "resize_bilinear")
    return out


def resize_bilinear_h3(x: torch.Tensor, OH: int, OW: int, align_corners: bool) -> Split:
    """resize_bilinear whose result [N,OH,OW,C] is stored as a plane pair (it only feeds a convolution)."""
    _chk_f32(x)
    N, H, W, Cc = x.shape
    out = Split.empty(N, OH, OW, Cc, device=x.device)
    code = _lib.load().siu3r_resize_bilinear_nhwc_h3(_p(x), N, H, W, Cc, x.stride(2), out.data_ptr(), OH, OW, Cc, out.plane, 1 if align_corners else 0,
                                                     _stream())
    _lib.check(code, "resize_bilinear_h3")
    return out


def pixel_shuffle(g: torch.Tensor, N: int, H: int, W: int, Cc: int, s: int, add: torch.Tensor | None = None, out: torch.Tensor | None = None):
    if out is None:
        out = torch.empty(N, H * s, W * s, Cc, device=g.device, dtype=torch.float32)
    _lib.check(_lib.load().siu3r_pixel_shuffle_nhwc(_p(g), N, H, W, Cc, s, _p(add), _p(out), _stream()), "pixel_shuffle")
    return out


def nchw_to_nhwc(x: torch.Tensor, ldy: int | None = None):
    N, Cc, H, W = x.shape
    ldy = Cc if ldy is None else ldy
    y = torch.zeros(N, H, W, ldy, device=x.device, dtype=torch.float32) if ldy != Cc else torch.empty(N, H, W, ldy, device=x.device,
                                                                                                      dtype=torch.float32)
    _lib.check(_lib.load().siu3r_nchw_to_nhwc(_p(x.contiguous()), _p(y), N, Cc, H * W, ldy, _stream()), "nchw_to_nhwc")
    return y


def nhwc_to_nchw(x: torch.Tensor, Cc: int | None = None):
    N, H, W, ld = x.shape
    Cc = ld if Cc is None else Cc
    y = torch.empty(N, Cc, H, W, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().siu3r_nhwc_to_nchw(_p(x), ld, _p(y), N, Cc, H * W, _stream()), "nhwc_to_nchw")
    return y


def maxpool3x3s2(x: torch.Tensor):
    N, H, W, Cc = x.shape
    OH, OW = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    y = torch.empty(N, OH, OW, Cc, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().siu3r_maxpool3x3s2_nhwc(_p(x), N, H, W, Cc, _p(y), _stream()), "maxpool")
    return y


def dwconv3x3(x_ptr: int, ldx: int, bsx: int, N: int, H: int, W: int, Cc: int, w: torch.Tensor, b: torch.Tensor, y_ptr: int, ldy: int, bsy: int,
              gelu: bool):
    code = _lib.load().siu3r_dwconv3x3_nhwc(x_ptr, ldx, bsx, N, H, W, Cc, _p(w), _p(b), y_ptr, ldy, bsy, 1 if gelu else 0, _stream())
    _lib.check(code, "dwconv3x3")


def groupnorm(x: torch.Tensor, groups: int, w: torch.Tensor, b: torch.Tensor, eps: float, relu: bool, out: torch.Tensor | None = None):
    """x: [N, HW, C] contiguous."""
    N, HW, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    ws = torch.empty(N * groups * 2, device=x.device, dtype=torch.float64)
    _lib.check(_lib.load().siu3r_groupnorm_nhwc(_p(x), N, HW, Cc, groups, _p(w), _p(b), eps, 1 if relu else 0, _p(out), _p(ws), _stream()), "groupnorm")
    return out


def depth_exp(xyz: torch.Tensor, n: int, ldx: int):
    pts = torch.empty(n, 3, device=xyz.device, dtype=torch.float32)
    _lib.check(_lib.load().siu3r_depth_exp(_p(xyz), ldx, _p(pts), n, _stream()), "depth_exp")
    return pts


def gaussian_adapter(raw: torch.Tensor):
    """raw [G, 83] contiguous -> (cov [G,3,3], harmonics [G,3,25], opacities [G], scales [G,3], rotations [G,4])."""
    _chk_f32(raw)
    assert raw.is_contiguous() and raw.shape[-1] == 83
    G = raw.numel() // 83
    dev = raw.device
    cov = torch.empty(G, 3, 3, device=dev)
    harm = torch.empty(G, 3, 25, device=dev)
    opac = torch.empty(G, device=dev)
    scales = torch.empty(G, 3, device=dev)
    rots = torch.empty(G, 4, device=dev)
    _lib.check(_lib.load().siu3r_gaussian_adapter(_p(raw), G, _p(cov), _p(harm), _p(opac), _p(scales), _p(rots), _stream()), "gaussian_adapter")
    return cov, harm, opac, scales, rots


def attn_mask_from_logits(logits: torch.Tensor, B: int, T: int, Hm: int, Wm: int, Q: int, oh: int, ow: int):
    mask = torch.empty(B, Q, T * oh * ow, device=logits.device, dtype=torch.uint8)
    _lib.check(_lib.load().siu3r_attn_mask_from_logits(_p(logits), B, T, Hm, Wm, Q, oh, ow, _p(mask), _stream()), "attn_mask")
    return mask


def resize_select(x: torch.Tensor, idx: torch.Tensor, OH: int, OW: int):
    N, H, W, Cc = x.shape
    nsel = idx.numel()
    y = torch.empty(N, OH, OW, nsel, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().siu3r_resize_select(_p(x), N, H, W, Cc, _p(idx), nsel, _p(y), OH, OW, _stream()), "resize_select")
    return y


def argmax_area(probs: torch.Tensor, score: torch.Tensor, thr: float):
    nq = probs.shape[-1]
    npix = probs.numel() // nq
    labels = torch.empty(npix, device=probs.device, dtype=torch.int32)
    area = torch.empty(nq, device=probs.device, dtype=torch.int32)
    orig = torch.empty(nq, device=probs.device, dtype=torch.int32)
    _lib.check(_lib.load().siu3r_argmax_area(_p(probs), npix, nq, _p(score), thr, _p(labels), _p(area), _p(orig), _stream()), "argmax_area")
    return labels, area, orig


def label_lut(labels: torch.Tensor, seg_lut: torch.Tensor, sem_lut: torch.Tensor, sem_out=None, inst_out=None):
    npix = labels.numel()
    seg = torch.empty(npix, device=labels.device, dtype=torch.int32)
    ret_all = sem_out is None
    sem = torch.empty_like(seg) if sem_out is None else sem_out
    inst = torch.empty_like(seg) if inst_out is None else inst_out
    assert sem.is_contiguous() and inst.is_contiguous() and sem.numel() == npix and inst.numel() == npix
    _lib.check(_lib.load().siu3r_label_lut(_p(labels), npix, _p(seg_lut), _p(sem_lut), _p(seg), _p(sem), _p(inst), _stream()), "label_lut")
    return (seg, sem, inst) if ret_all else seg


def qc_logits(probs: torch.Tensor, keep: torch.Tensor, cls: torch.Tensor):
    nq = probs.shape[-1]
    npix = probs.numel() // nq
    nk, ncls = cls.shape
    out = torch.empty(npix, nk, ncls, device=probs.device, dtype=torch.float32)
    _lib.check(_lib.load().siu3r_qc_logits(_p(probs), npix, nq, _p(keep), nk, _p(cls), ncls, _p(out), _stream()), "qc_logits")
    return out


def raster_forward(means, cov, shs, opac, viewmatrix, projmatrix, campos, bg, tan_fovx, tan_fovy, H, W, sh_degree, sh_layout, dup_capacity=None,
                   count_touched=True, debug=False):
    """One camera.  means [G,3]; cov [G,6] or [G,3,3]; shs [G,M,3] (sh_layout 0) or [G,3,M] (1); returns dict."""
    lib = _lib.load()
    _chk_f32(means, cov, shs, opac, viewmatrix, projmatrix, campos, bg)
    G = means.shape[0]
    cov_stride = 6 if cov.shape[-1] == 6 else 9
    M = shs.shape[1] if sh_layout == 0 else shs.shape[2]
    dev = means.device
    if dup_capacity is None:
        dup_capacity = max(1 << 16, 8 * G)
    while True:
        ws_bytes = int(lib.siu3r_raster_workspace_bytes(G, H, W, dup_capacity))
        if ws_bytes < 0:
            raise RuntimeError("raster_workspace_bytes: invalid arguments")
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        color = torch.empty(3, H, W, device=dev)
        depth = torch.empty(H, W, device=dev)
        opacity = torch.empty(H, W, device=dev)
        radii = torch.empty(G, device=dev, dtype=torch.int32)
        n_touched = torch.empty(G, device=dev, dtype=torch.int32) if count_touched else None
        nren = C.c_int64(0)
        gx, gy = (W + 15) // 16, (H + 15) // 16
        dbg = {}
        if debug:
            dbg = dict(tiles=torch.zeros(G, device=dev, dtype=torch.int32), offsets=torch.zeros(G, device=dev, dtype=torch.int32),
                       keys=torch.zeros(dup_capacity, device=dev, dtype=torch.int64), values=torch.zeros(dup_capacity, device=dev, dtype=torch.int32),
                       ranges=torch.zeros(gx * gy, 2, device=dev, dtype=torch.int32))
        code = lib.siu3r_raster_forward(G, H, W, sh_degree, M, sh_layout, cov_stride, _p(means), _p(cov), _p(shs), _p(opac), _p(viewmatrix),
                                        _p(projmatrix), _p(campos), _p(bg), float(tan_fovx), float(tan_fovy), _p(color), _p(depth), _p(opacity),
                                        _p(radii), _p(n_touched), _p(ws), ws_bytes, dup_capacity, C.byref(nren), _p(dbg.get("tiles")),
                                        _p(dbg.get("offsets")), _p(dbg.get("keys")), _p(dbg.get("values")), _p(dbg.get("ranges")), _stream())
        if code == -2 and nren.value > dup_capacity:
            dup_capacity = int(nren.value * 1.05) + 1024  # exact count is known now: retry once with enough room
            continue
        _lib.check(code, "raster_forward")
        break
    res = dict(color=color, depth=depth, opacity=opacity, radii=radii, n_touched=n_touched, num_rendered=int(nren.value))
    res.update(dbg)
    return res


def raster_forward_nosync(means, cov, shs, opac, viewmatrix, projmatrix, campos, bg, tan_fovx, tan_fovy, H, W, sh_degree, sh_layout, status, ws=None,
                          dup_capacity=None, count_touched=False, out=None):
    """One camera without any host synchronisation (siu3r_raster_forward_nosync): `status` = 4 uint32 on the device that receive
    {duplicates, largest tile, flags, 0}; the caller checks flags != 0 later (see renderer.render_cuda).  `ws` = reusable workspace tensor."""
    lib = _lib.load()
    _chk_f32(means, cov, shs, opac, viewmatrix, projmatrix, campos, bg)
    G = means.shape[0]
    cov_stride = 6 if cov.shape[-1] == 6 else 9
    M = shs.shape[1] if sh_layout == 0 else shs.shape[2]
    dev = means.device
    if dup_capacity is None:
        dup_capacity = max(1 << 16, 8 * G)
    ws_bytes = int(lib.siu3r_raster_workspace_bytes(G, H, W, dup_capacity))
    if ws is None or ws.numel() < ws_bytes:
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    if out is None:
        out = dict(color=torch.empty(3, H, W, device=dev), depth=torch.empty(H, W, device=dev), opacity=torch.empty(H, W, device=dev),
                   radii=torch.empty(G, device=dev, dtype=torch.int32), n_touched=torch.empty(G, device=dev, dtype=torch.int32) if count_touched else None)
    assert status.dtype == torch.int32 and status.numel() >= 4 and status.is_cuda
    code = lib.siu3r_raster_forward_nosync(G, H, W, sh_degree, M, sh_layout, cov_stride, _p(means), _p(cov), _p(shs), _p(opac), _p(viewmatrix),
                                           _p(projmatrix), _p(campos), _p(bg), float(tan_fovx), float(tan_fovy), _p(out["color"]), _p(out["depth"]),
                                           _p(out["opacity"]), _p(out["radii"]), _p(out["n_touched"]), _p(ws), ws.numel(), dup_capacity, _p(status),
                                           _stream())
    _lib.check(code, "raster_forward_nosync")
    out["ws"], out["dup_capacity"] = ws, dup_capacity
    return out


def raster_features_forward_nosync(means, cov, opac, feats, viewmat, intr, near, far, H, W, status, ws=None, dup_capacity=None, out=None, alpha=None):
    """raster_features_forward without a host round trip (siu3r_raster_features_forward_nosync): `status` = 4 int32 on the device receiving
    {duplicates, largest tile, flags, 0}; flags != 0 means the frame was NOT rendered (the caller re-renders it through raster_features_forward).
    `ws` = reusable workspace.  -> dict(features [H,W,C], alpha, ws)."""
    lib = _lib.load()
    _chk_f32(means, cov, opac, feats, viewmat)
    assert means.is_contiguous() and cov.is_contiguous() and opac.is_contiguous() and feats.is_contiguous() and viewmat.is_contiguous()
    assert status.dtype == torch.int32 and status.numel() >= 4 and status.is_contiguous()
    G, Cc = means.shape[0], feats.shape[1]
    cov_stride = 6 if cov.shape[-1] == 6 else 9
    dev = means.device
    if dup_capacity is None:
        dup_capacity = max(1 << 16, 8 * G)
    ws_bytes = int(lib.siu3r_raster_workspace_bytes(G, H, W, dup_capacity))
    if ws is None or ws.numel() < ws_bytes:
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    if out is None:
        out = torch.empty(H, W, Cc, device=dev)
    fx, fy, cx, cy = (float(v) for v in intr)
    code = lib.siu3r_raster_features_forward_nosync(G, H, W, Cc, cov_stride, _p(means), _p(cov), _p(opac), _p(feats), _p(viewmat), fx, fy, cx, cy, float(near),
                                                    float(far), _p(out), _p(alpha), None, _p(ws), ws.numel(), dup_capacity, _p(status), _stream())
    _lib.check(code, "raster_features_forward_nosync")
    return dict(features=out, alpha=alpha, ws=ws)


def raster_features_forward(means, cov, opac, feats, viewmat, intr, near, far, H, W, dup_capacity=None, want_alpha=True, want_radii=False):
    """N-channel feature splatting of one camera (gsplat.rasterization semantics).  means [G,3]; cov [G,3,3] or [G,6]; opac [G];
    feats [G,C]; viewmat [4,4] world-to-camera (device); intr = (fx, fy, cx, cy) in pixels.  -> dict(features [H,W,C], alpha [H,W], ...)."""
    lib = _lib.load()
    _chk_f32(means, cov, opac, feats, viewmat)
    assert means.is_contiguous() and cov.is_contiguous() and opac.is_contiguous() and feats.is_contiguous() and viewmat.is_contiguous()
    G, Cc = means.shape[0], feats.shape[1]
    cov_stride = 6 if cov.shape[-1] == 6 else 9
    dev = means.device
    if dup_capacity is None:
        dup_capacity = max(1 << 16, 8 * G)
    intr_arr = (C.c_float * 4)(*[float(v) for v in intr])
    while True:
        ws_bytes = int(lib.siu3r_raster_workspace_bytes(G, H, W, dup_capacity))
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        out = torch.empty(H, W, Cc, device=dev)
        alpha = torch.empty(H, W, device=dev) if want_alpha else None
        radii = torch.empty(G, 2, device=dev, dtype=torch.int32) if want_radii else None
        nren = C.c_int64(0)
        code = lib.siu3r_raster_features_forward(G, H, W, Cc, cov_stride, _p(means), _p(cov), _p(opac), _p(feats), _p(viewmat), intr_arr, float(near),
                                                 float(far), _p(out), _p(alpha), _p(radii), _p(ws), ws_bytes, dup_capacity, C.byref(nren), _stream())
        if code == -2 and nren.value > dup_capacity:
            dup_capacity = int(nren.value * 1.05) + 1024
            continue
        _lib.check(code, "raster_features_forward")
        break
    return dict(features=out, alpha=alpha, radii=radii, num_rendered=int(nren.value))
