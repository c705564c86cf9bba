"""SIU3RModel -- the per-image-pair hot path on hand-written sm_100a kernels, behind the reference's call surface.

Mirrors /root/reference/src/models/model.py:31-389 (SIU3RModel.__init__/forward): same constructor config fields, same
state_dict key set (loaded unchanged, see weights.py), same return tuple.  Everything between the input tensors and the
returned tensors runs through the C ABI (ops.py); PyTorch only owns memory and the stream.

Layouts: activations are token-major / NHWC fp32.  The two views are batched view-major like the reference's
torch.cat((img1, img2), 0) (backbone_croco.py:174-176): row block i = v * B + b.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from types import SimpleNamespace

import torch

from . import ops
from .gaussians import Gaussians
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, ELT_ADD, ELT_RELU, ELT_SIGMOID
from .weights import Packer, reference_points, sine_pos_2d, sine_pos_3d

# label tables of the reference (src/utils/scannet_constant.py:1-35): 20 ScanNet classes, stuff = {wall, floor}
PANOPTIC_SEMANTIC2NAME = {1: "wall", 2: "floor", 3: "cabinet", 4: "bed", 5: "chair", 6: "sofa", 7: "table", 8: "door", 9: "window",
                          10: "bookshelf", 11: "picture", 12: "counter", 13: "desk", 14: "curtain", 15: "refrigerator", 16: "shower curtain",
                          17: "toilet", 18: "sink", 19: "bathtub", 20: "otherfurniture"}
STUFF_CLASSES = [0, 1]
# Scheduling of the two concurrent chains of a forward (A/B switches for measurements; results do not depend on them)
SEG_PRIORITY = os.environ.get("SIU3R_SEG_PRIORITY", "1") != "0"          # panoptic chain on a high-priority stream
FUSE_LN = os.environ.get("SIU3R_FUSE_LN", "1") != "0"                    # h3: LayerNorm of the ViT blocks fused into the adjacent GEMM epilogues
HEAD_CLUSTER_CAP = int(os.environ.get("SIU3R_HEAD_CLUSTER_CAP", "0"))    # CTA pairs (of 74) the decoder / head GEMMs may occupy next to it; 0 = all
#                                                                          (70 was measured neutral to slightly negative in steady state: 17.53 vs 17.40 ms)


@dataclass
class ModelCfg:
    """Subset of src/config.py:46-80 (ModelCfg / CrocoCfg / Mask2formerCfg / GaussianHeadCfg) that shapes the forward pass."""
    image_size: tuple = (256, 256)
    enc_depth: int = 24
    dec_depth: int = 12
    enc_embed_dim: int = 1024
    dec_embed_dim: int = 768
    enc_num_heads: int = 16
    dec_num_heads: int = 12
    patch_size: int = 16
    num_queries: int = 100
    seg_threshold: float = 0.5
    id2label: dict = field(default_factory=lambda: dict(PANOPTIC_SEMANTIC2NAME))
    label_ids_to_fuse: list = field(default_factory=lambda: list(STUFF_CLASSES))
    sh_degree: int = 4
    interaction_indexes: tuple = (5, 11, 17, 23)

    @classmethod
    def from_reference(cls, cfg) -> "ModelCfg":
        """Accepts the reference's nested ModelCfg (src/config.py:46-80: .croco, .mask2former, .gaussian_head, .image_size) -- the object
        SIU3RModel(cfg) is built from in pipeline.py:31 -- and checks that it describes the architecture this engine implements
        (ViT-L/16 encoder, 12-layer 768-wide decoder, 64-wide heads, RoPE100, SH degree 4); anything else is refused, not approximated."""
        c, m, g = cfg.croco, cfg.mask2former, cfg.gaussian_head
        ref = cls()
        got = dict(enc_depth=c.enc_depth, dec_depth=c.dec_depth, enc_embed_dim=c.enc_embed_dim, dec_embed_dim=c.dec_embed_dim,
                   enc_num_heads=c.enc_num_heads, dec_num_heads=c.dec_num_heads, patch_size=c.patch_size, sh_degree=g.sh_degree)
        bad = {k: v for k, v in got.items() if v != getattr(ref, k)}
        if getattr(c, "pos_embed", "RoPE100") != "RoPE100":
            bad["pos_embed"] = c.pos_embed
        if bad:
            raise ValueError(f"siu3r_b200 implements the published SIU3R architecture only; unsupported configuration values: {bad}")
        return cls(image_size=tuple(cfg.image_size), num_queries=m.num_queries, seg_threshold=m.seg_threshold, id2label=dict(m.id2label),
                   label_ids_to_fuse=list(m.label_ids_to_fuse), **got)


class SIU3RModel:
    def __init__(self, cfg: ModelCfg | None = None, precision: str = "tf32"):
        if cfg is not None and hasattr(cfg, "croco") and hasattr(cfg, "mask2former"):
            cfg = ModelCfg.from_reference(cfg)        # the reference's own nested config object
        self.cfg = cfg or ModelCfg()
        assert precision in ("tf32", "fp32x3", "h3")
        self.prec = {"tf32": ops.PREC_TF32, "fp32x3": ops.PREC_FP32X3, "h3": ops.PREC_H3}[precision]
        self.precision = precision
        self._sd = None
        self._ready = False
        self._cache = {}
        self.capture = None  # set to a dict to record stage-boundary tensors (tests)
        S0, S1 = self.cfg.image_size
        assert S0 % 32 == 0 and S1 % 32 == 0, "image size must be a multiple of 32 (patch 16, adapter stride 32)"
        if S0 > S1:
            # the reference's heads run portrait batches transposed (transpose_to_landscape, croco/misc.py:71-113); not implemented here
            raise NotImplementedError(f"portrait image_size {S0}x{S1}: only landscape / square sizes (H <= W) are implemented")

    # ---- nn.Module-like surface -------------------------------------------------------------------------------
    def load_state_dict(self, sd: dict, strict: bool = False):
        """nn.Module.load_state_dict semantics against the reference key set (state_shapes.json: the 1660 entries of SIU3RModel.state_dict()):
        returns (missing_keys, unexpected_keys); strict=True raises on either; a tensor of the wrong shape always raises, as torch does.
        The reference loads with strict=False (inference.py:119-121), so e.g. the lpips.* entries of a Lightning checkpoint are ignored."""
        from .synth import load_state_shapes
        expected = load_state_shapes()
        if any(k.startswith("model.") for k in sd) and not any(k in expected for k in sd):
            # a Lightning checkpoint's state_dict ("model." + key, inference.py:119-121): strip the prefix BEFORE the shape check
            sd = {(k[6:] if k.startswith("model.") else k): v for k, v in sd.items()}
        missing = [k for k in expected if k not in sd]
        unexpected = [k for k in sd if k not in expected]
        wrong = [f"{k}: checkpoint {list(sd[k].shape)} vs model {expected[k][0]}" for k in expected
                 if k in sd and hasattr(sd[k], "shape") and list(sd[k].shape) != list(expected[k][0])]
        if wrong:
            raise RuntimeError("Error(s) in loading state_dict for SIU3RModel: size mismatch for " + "; ".join(wrong[:8]) + (" ..." if len(wrong) > 8 else ""))
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for SIU3RModel: {len(missing)} missing key(s) {missing[:4]}, "
                               f"{len(unexpected)} unexpected key(s) {unexpected[:4]}")
        self._sd = sd
        self._ready = False
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def eval(self):
        return self

    def cuda(self, device=None):
        """Pack the weights onto `device` (default: the current CUDA device).  Every later call runs under torch.cuda.device(self.dev), so kernels
        are launched on that device's current stream whatever the caller's current device is."""
        if isinstance(device, torch.device):
            device = device.index
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        if self._ready and dev == getattr(self, "dev", None):
            return self
        if self._sd is None:
            raise RuntimeError("load_state_dict() first: this engine has no random init of its own")
        self.dev = dev
        self._cache, self._graphs = {}, {}
        self._streams, self._seg_stream = None, None
        with torch.cuda.device(self.dev):
            self._pack()
        return self

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    # ---- weight packing -----------------------------------------------------------------------------------------
    def _pack(self):
        P = Packer(self._sd, self.dev, self.prec)
        c = self.cfg
        w = SimpleNamespace()
        # backbone
        w.patch = P.conv("backbone.patch_embed.proj", pad_cin_to=4)
        w.intr_w, w.intr_b = P.vec("backbone.intrinsic_encoder.weight"), P.vec("backbone.intrinsic_encoder.bias")
        w.enc = []
        for i in range(c.enc_depth):
            p = f"backbone.enc_blocks.{i}."
            w.enc.append(SimpleNamespace(n1=(P.vec(p + "norm1.weight"), P.vec(p + "norm1.bias")), qkv=P.linear(p + "attn.qkv"),
                                         proj=P.linear(p + "attn.proj"), n2=(P.vec(p + "norm2.weight"), P.vec(p + "norm2.bias")),
                                         fc1=P.linear(p + "mlp.fc1"), fc2=P.linear(p + "mlp.fc2")))
        w.enc_norm = (P.vec("backbone.enc_norm.weight"), P.vec("backbone.enc_norm.bias"))
        w.dec_embed = P.linear("backbone.decoder_embed")
        w.dec = []
        for name in ("dec_blocks", "dec_blocks2"):
            blocks = []
            for i in range(c.dec_depth):
                p = f"backbone.{name}.{i}."
                ln = lambda n: (P.vec(p + n + ".weight"), P.vec(p + n + ".bias"))
                blocks.append(SimpleNamespace(n1=ln("norm1"), n2=ln("norm2"), n3=ln("norm3"), ny=ln("norm_y"), qkv=P.linear(p + "attn.qkv"),
                                              proj=P.linear(p + "attn.proj"), cq=P.linear(p + "cross_attn.projq"),
                                              ckv=P.linear_cat([p + "cross_attn.projk", p + "cross_attn.projv"]),
                                              cproj=P.linear(p + "cross_attn.proj"), fc1=P.linear(p + "mlp.fc1"), fc2=P.linear(p + "mlp.fc2")))
            w.dec.append(blocks)
        w.dec_norm = (P.vec("backbone.dec_norm.weight"), P.vec("backbone.dec_norm.bias"))
        if self.prec == ops.PREC_H3 and FUSE_LN:
            # LayerNorm folded into the consuming projection (ops.Weight.fold_ln): norm1 -> qkv, norm2 -> fc1 (encoder); norm1 -> qkv, norm_y -> k|v,
            # norm2 -> q, norm3 -> fc1 (decoder).  The un-folded weights stay: block 0 of the encoder and the un-fused reference path use them.
            for bk in w.enc:
                bk.qkv_ln, bk.fc1_ln = bk.qkv.fold_ln(*bk.n1), bk.fc1.fold_ln(*bk.n2)
            for blocks in w.dec:
                for bk in blocks:
                    bk.qkv_ln, bk.ckv_ln, bk.cq_ln, bk.fc1_ln = bk.qkv.fold_ln(*bk.n1), bk.ckv.fold_ln(*bk.ny), bk.cq.fold_ln(*bk.n2), bk.fc1.fold_ln(*bk.n3)
        # DPT heads
        w.heads = {}
        for hname in ("downstream_head1", "downstream_head2", "gaussian_param_head1", "gaussian_param_head2"):
            p = hname + ".dpt."
            h = SimpleNamespace()
            h.act_conv = [P.conv(p + f"act_postprocess.{i}.0") for i in range(4)]
            h.act_up0, h.s0 = P.conv_transpose(p + "act_postprocess.0.1")
            h.act_up1, h.s1 = P.conv_transpose(p + "act_postprocess.1.1")
            h.act_down3 = P.conv(p + "act_postprocess.3.1")
            h.layer_rn = [P.conv(p + f"scratch.layer_rn.{i}") for i in range(4)]
            h.refine = []
            for r in (1, 2, 3, 4):
                q = p + f"scratch.refinenet{r}."
                h.refine.append(SimpleNamespace(out_conv=P.conv(q + "out_conv"),
                                                r1=(P.conv(q + "resConfUnit1.conv1"), P.conv(q + "resConfUnit1.conv2")),
                                                r2=(P.conv(q + "resConfUnit2.conv1"), P.conv(q + "resConfUnit2.conv2"))))
            h.head0 = P.conv(p + "head.0")
            if hname.startswith("downstream"):
                h.head2, h.head4 = P.conv(p + "head.2"), P.conv(p + "head.4")
            else:
                h.head4 = P.conv(p + "head.4")
                h.merger = P.conv(p + "input_merger.0", pad_cin_to=4)
            w.heads[hname] = h
        # ViT adapter
        a = SimpleNamespace()
        a.stem = [P.conv("adapter.spm.stem.0", bn="adapter.spm.stem.1", pad_cin_to=4), P.conv("adapter.spm.stem.3", bn="adapter.spm.stem.4"),
                  P.conv("adapter.spm.stem.6", bn="adapter.spm.stem.7")]
        a.conv2 = P.conv("adapter.spm.conv2.0", bn="adapter.spm.conv2.1")
        a.conv3 = P.conv("adapter.spm.conv3.0", bn="adapter.spm.conv3.1")
        a.conv4 = P.conv("adapter.spm.conv4.0", bn="adapter.spm.conv4.1")
        lvl = P.t("adapter.level_embed")
        a.fc1 = P.conv("adapter.spm.fc1")
        a.fc = [P.conv("adapter.spm.fc2", extra_bias=lvl[0]), P.conv("adapter.spm.fc3", extra_bias=lvl[1]), P.conv("adapter.spm.fc4", extra_bias=lvl[2])]

        def extractor(p):
            ln = lambda n: (P.vec(p + n + ".weight"), P.vec(p + n + ".bias"))
            dw, db = P.dwconv(p + "ffn.dwconv.dwconv")
            return SimpleNamespace(qn=ln("query_norm"), fn=ln("feat_norm"), ffn_norm=ln("ffn_norm"), value=P.linear(p + "attn.value_proj"),
                                   ow=P.linear_cat([p + "attn.sampling_offsets", p + "attn.attention_weights"]), out=P.linear(p + "attn.output_proj"),
                                   fc1=P.linear(p + "ffn.fc1"), dw=dw, db=db, fc2=P.linear(p + "ffn.fc2"))
        a.inter = []
        for i in range(4):
            ex = [extractor(f"adapter.interactions.{i}.extractor.")]
            if i == 3:
                ex += [extractor(f"adapter.interactions.{i}.extra_extractors.{j}.") for j in range(2)]
            a.inter.append(ex)
        a.up, a.up_s = P.conv_transpose("adapter.up")
        a.bn = [tuple(t.contiguous().to(self.dev) for t in P.bn_scale_shift(f"adapter.norm{i}")) for i in (1, 2, 3, 4)]
        w.adapter = a
        # Mask2Former
        m = SimpleNamespace()
        pd = "mask2former.model.pixel_decoder."
        m.in_proj = [(P.conv(pd + f"input_projections.{i}.0"), P.vec(pd + f"input_projections.{i}.1.weight"), P.vec(pd + f"input_projections.{i}.1.bias"))
                     for i in range(3)]
        m.pd_level_embed = P.t(pd + "level_embed")
        m.enc = []
        for i in range(6):
            p = pd + f"encoder.layers.{i}."
            ln = lambda n: (P.vec(p + n + ".weight"), P.vec(p + n + ".bias"))
            m.enc.append(SimpleNamespace(value=P.linear(p + "self_attn.value_proj"),
                                         ow=P.linear_cat([p + "self_attn.sampling_offsets", p + "self_attn.attention_weights"]),
                                         out=P.linear(p + "self_attn.output_proj"), ln1=ln("self_attn_layer_norm"), fc1=P.linear(p + "fc1"),
                                         fc2=P.linear(p + "fc2"), ln2=ln("final_layer_norm")))
        m.lateral = (P.conv(pd + "adapter_1.0"), P.vec(pd + "adapter_1.1.weight"), P.vec(pd + "adapter_1.1.bias"))
        m.output = (P.conv(pd + "layer_1.0"), P.vec(pd + "layer_1.1.weight"), P.vec(pd + "layer_1.1.bias"))
        m.mask_proj = P.conv(pd + "mask_projection")
        tm = "mask2former.model.transformer_module."
        m.q_feat, m.q_pos = P.vec(tm + "queries_features.weight"), P.vec(tm + "queries_embedder.weight")
        m.tm_level_embed = P.t(tm + "level_embed.weight")
        m.dec = []
        E = 256
        for i in range(9):
            p = tm + f"decoder.layers.{i}."
            ln = lambda n: (P.vec(p + n + ".weight"), P.vec(p + n + ".bias"))
            m.dec.append(SimpleNamespace(
                cq=P.linear_rows(p + "cross_attn.in_proj_weight", p + "cross_attn.in_proj_bias", 0, E),
                ck=P.linear_rows(p + "cross_attn.in_proj_weight", p + "cross_attn.in_proj_bias", E, 2 * E),
                cv=P.linear_rows(p + "cross_attn.in_proj_weight", p + "cross_attn.in_proj_bias", 2 * E, 3 * E),
                cout=P.linear(p + "cross_attn.out_proj"), cln=ln("cross_attn_layer_norm"),
                sqk=P.linear_cat([p + "self_attn.q_proj", p + "self_attn.k_proj"]), sv=P.linear(p + "self_attn.v_proj"),
                sout=P.linear(p + "self_attn.out_proj"), sln=ln("self_attn_layer_norm"), fc1=P.linear(p + "fc1"), fc2=P.linear(p + "fc2"),
                fln=ln("final_layer_norm")))
        m.dec_ln = (P.vec(tm + "decoder.layernorm.weight"), P.vec(tm + "decoder.layernorm.bias"))
        m.mask_mlp = [P.linear(tm + f"decoder.mask_predictor.mask_embedder.{i}.0") for i in range(3)]
        m.cls = P.linear("mask2former.class_predictor")
        w.m2f = m
        self.w = w
        self._ready = True   # (the caller's state_dict stays referenced so that .cuda(other_device) can re-pack)

    # ---- shape-dependent constants (host-built once per image size) -------------------------------------------------
    def _consts(self, B: int, S0: int, S1: int, V: int = 2):
        key = (B, S0, S1, V)
        if key in self._cache:
            return self._cache[key]
        gh, gw = S0 // 16, S1 // 16
        ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
        pos = torch.stack([ys.flatten(), xs.flatten()], -1)
        pos = torch.cat([pos, torch.tensor([[gh, 0]])], 0)  # intrinsics token at (y_last + 1, 0): backbone_croco.py:148-150
        k = SimpleNamespace()
        k.pos_enc = pos[None].repeat(V * B, 1, 1).contiguous().to(self.dev)
        k.pos_dec = pos[None].repeat(B, 1, 1).contiguous().to(self.dev)
        k.rope_tab = ops.rope2d_table(int(pos.max()) + 1, 64, 100.0, 1.0, self.dev)   # RoPE factors for the fused projection epilogues
        ad_shapes = [(S0 // 8, S1 // 8), (S0 // 16, S1 // 16), (S0 // 32, S1 // 32)]
        k.ad_ref = reference_points(ad_shapes).to(self.dev)
        m = self.w.m2f
        lv = [(S0 // 32, S1 // 32), (S0 // 16, S1 // 16), (S0 // 8, S1 // 8)]  # pixel-decoder order: low -> high resolution
        k.m2f_shapes = lv
        k.m2f_ref = reference_points(lv).to(self.dev)
        pe = torch.cat([sine_pos_2d(h, w) + m.pd_level_embed[i][None] for i, (h, w) in enumerate(lv)], 0)
        k.m2f_pos = pe[None].repeat(V * B, 1, 1).reshape(-1, 256).contiguous().to(self.dev)
        k.tm_pos = [sine_pos_3d(V, h, w)[None].repeat(B, 1, 1).reshape(-1, 256).contiguous().to(self.dev) for (h, w) in lv]
        k.tm_lvl = [m.tm_level_embed[i].contiguous().to(self.dev) for i in range(3)]
        k.hidden0 = m.q_feat[None].repeat(B, 1, 1).view(-1, 256).contiguous()
        k.qpos = m.q_pos[None].repeat(B, 1, 1).view(-1, 256).contiguous()
        self._cache[key] = k
        return k

    def _par(self, fns):
        """Run independent branches on side streams (fork/join with events; captured as parallel graph branches).
        Every branch starts after all work enqueued so far and the caller's stream resumes after all of them."""
        if getattr(self, "serial", False):   # profiling: one stream, clean per-kernel timings
            return [fn() for fn in fns]
        if not getattr(self, "_streams", None):
            self._streams = [torch.cuda.Stream(device=self.dev) for _ in range(6)]
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        out = []
        used = []
        for i, fn in enumerate(fns):
            st = self._streams[i % len(self._streams)]
            if st not in used:
                st.wait_event(fork)
                used.append(st)
            with torch.cuda.stream(st):
                out.append(fn())
        for st in used:
            cur.wait_stream(st)
        return out

    def _mark(self, name):
        """Stage timing marks (tools/stage_times.py): external events so that they also become nodes of a captured graph."""
        if getattr(self, "marks", None) is not None:
            ev = torch.cuda.Event(enable_timing=True, external=True)
            ev.record()
            self.marks[name] = ev

    def _cap(self, name, t):
        if self.capture is not None:
            if isinstance(t, ops.Split):   # plane pair -> fp32 copy in the logical shape
                t = t.view(-1, t.shape[-1]).float().view(*t.shape)
            self.capture[name] = t

    # ---- building blocks ------------------------------------------------------------------------------------------
    # TF32 mode: tensors that ONLY feed GEMM / conv A operands are stored round-to-nearest by their producer (R = True);
    # `ar=` tells the consumer its A operand is already rounded, `ro=` asks a GEMM / conv to round its own output.
    # h3 mode (S = True): the same tensors are stored as fp16 (hi, lo) plane pairs (ops.Split) by their producer; consumers that are handed a
    # plain fp32 tensor split it themselves (one extra pass).  `ro=` therefore means "this result only feeds tensor-core operands".
    @property
    def R(self):
        return self.prec == ops.PREC_TF32

    @property
    def S(self):
        return self.prec == ops.PREC_H3

    @property
    def fuse_ln(self):
        """LayerNorm of the ViT blocks fused into the adjacent GEMM epilogues: the packed weights carry the gamma-folded projections."""
        return hasattr(self.w.enc[0], "qkv_ln")

    def _lin(self, x, wt, ar=False, ro=False, **kw):
        return ops.gemm(x, wt, precision=self.prec, a_rounded=ar and self.R, round_out=ro and (self.R or self.S), **kw)

    def _conv(self, x, wt, k, ar=False, ro=False, **kw):
        return ops.conv2d(x, wt, k, k, precision=self.prec, a_rounded=ar and self.R, round_out=ro and (self.R or self.S), **kw)

    def _ln(self, x, wb, eps, ro=True, out=None, **kw):
        if self.S and ro:
            return ops.layernorm_h3([x], [wb], eps, outs=None if out is None else [out])[0]
        return ops.layernorm(x, wb[0], wb[1], eps, round_out=ro and self.R, out=out, **kw)

    def _relu_op(self, x):
        """ReLU whose result only feeds a convolution."""
        if self.S:
            return ops.eltwise_h3(ELT_RELU, x)
        return ops.eltwise(ops.ELT_RELU_RN if self.R else ELT_RELU, x)

    def _add_op(self, a, b):
        """a + b whose result only feeds GEMM operands."""
        if self.S:
            return ops.eltwise_h3(ELT_ADD, a, b)
        return ops.eltwise(ops.ELT_ADD_RN if self.R else ELT_ADD, a, b)

    def _resize_op(self, x, OH, OW, align):
        """Bilinear resize whose result only feeds a convolution."""
        if self.S:
            return ops.resize_bilinear_h3(x, OH, OW, align)
        return ops.resize_bilinear(x, OH, OW, align, round_out=self.R)

    def _self_attn(self, h, blk, pos, Bn, N, C, nh, ln_stats=None):
        M = Bn * N
        if self.S:   # h3: q | k as an unscaled plane pair (RoPE in the epilogue), V^T plane pair written by the same launch, plane-pair output
            qkv = ops.Split.empty(M, 3 * C, device=self.dev, unscaled=True)
            vth = ops.Split.empty(C, (M + 7) // 8 * 8, device=self.dev, unscaled=True)
            # ln_stats: h is the RAW residual stream and norm1 rides in the projection's epilogue
            self._lin(h, blk.qkv if ln_stats is None else blk.qkv_ln, ro=True, out=qkv, rope=(pos, self._k.rope_tab, 2 * C), vt=(vth, 2 * C, {}),
                      unscaled=True, ln_stats=ln_stats)
            return ops.flash_attn_h3(qkv, 0, qkv, C, vth, N, Bn, nh, N, N, 0.125, split_out=True)
        vt = st = None
        if self.R:   # the V third of the projection is written as V^T by the GEMM epilogue (no transpose pass)
            vt, st = torch.empty(C, (M + 3) // 4 * 4, device=self.dev), {}
        qkv = self._lin(h, blk.qkv, ar=True, ro=True, rope=(pos, self._k.rope_tab, 2 * C),   # RoPE on q and k inside the epilogue
                        **({"vt": (vt, 2 * C, st)} if vt is not None else {}))
        a = torch.empty(M, C, device=self.dev)
        if self.R:   # TF32 mode: tcgen05 / TMEM flash attention
            ops.flash_attn_tc(qkv, 0, N * 3 * C, 3 * C, 3 * C, qkv, C, N * 3 * C, 3 * C, 3 * C, qkv, 2 * C, N * 3 * C, 3 * C, a, Bn, nh, N, N, 0.125,
                              round_out=True, vt=(vt, N) if st.get("ok") else None)
        else:        # 3xTF32 mode: mma.sync kernel with the register-level split
            ops.flash_attn_d64(qkv, 0, N * 3 * C, 3 * C, qkv, C, N * 3 * C, 3 * C, qkv, 2 * C, N * 3 * C, 3 * C, a, Bn, nh, N, N, 0.125, self.prec)
        return a

    def _encoder(self, x, pos, Bn, N):
        """24 x Block (croco/blocks.py:127-130).  The residual stream is updated in place, except that the outputs of the
        blocks the adapter reads (interaction_indexes) are frozen: the following block writes into a fresh buffer."""
        if self.fuse_ln:
            return self._encoder_fused(x, pos, Bn, N)
        keep = {}
        C = 1024
        frozen = False
        for i, blk in enumerate(self.w.enc):
            h = self._ln(x, blk.n1, 1e-6)
            a = self._self_attn(h, blk, pos, Bn, N, C, 16)
            if frozen:
                x = self._lin(a, blk.proj, ar=True, residual=x)  # new buffer; the kept tensor stays intact
                frozen = False
            else:
                self._lin(a, blk.proj, ar=True, residual=x, out=x)
            h = self._ln(x, blk.n2, 1e-6)
            f = self._lin(h, blk.fc1, ar=True, ro=True, act=ACT_GELU)
            self._lin(f, blk.fc2, ar=True, residual=x, out=x)
            if i in self.cfg.interaction_indexes:
                keep[i] = x
                frozen = True
                self._cap(f"enc{i}", x)
                ev = torch.cuda.Event()
                ev.record()
                self._keep_ev[i] = ev   # the adapter stream starts interaction #k as soon as its ViT block is done
        return x, keep

    def _encoder_fused(self, x, pos, Bn, N):
        """_encoder with norm1 / norm2 fused into the GEMMs around them (h3 mode): every residual-adding projection also emits the plane pair of
        the new residual stream and its row statistics; the next projection reads the RAW rows and applies the LayerNorm in its epilogue
        (siu3r_gemm_h3_ln).  One layernorm_kernel launch is left (block 0, whose input rows come from two different producers)."""
        keep = {}
        C, M = 1024, Bn * N
        nblk = len(self.w.enc)
        stats = torch.zeros(2 * nblk, M, 2, device=self.dev, dtype=torch.int64)   # [2i]: rows after proj of block i, [2i+1]: after its fc2
        xs = None
        frozen = False
        for i, blk in enumerate(self.w.enc):
            if i == 0:
                a = self._self_attn(self._ln(x, blk.n1, 1e-6), blk, pos, Bn, N, C, 16)
            else:
                a = self._self_attn(xs, blk, pos, Bn, N, C, 16, ln_stats=stats[2 * i - 1])
            xn = torch.empty(M, C, device=self.dev) if frozen else x      # a kept tensor stays intact: the next block writes a fresh buffer
            frozen = False
            xs = ops.Split.empty(M, C, device=self.dev)
            self._lin(a, blk.proj, residual=x, out=xs, out_f32=xn, stats_out=stats[2 * i])
            x = xn
            f = self._lin(xs, blk.fc1_ln, ro=True, act=ACT_GELU, ln_stats=stats[2 * i])
            if i + 1 < nblk:
                xs = ops.Split.empty(M, C, device=self.dev)
                self._lin(f, blk.fc2, residual=x, out=xs, out_f32=x, stats_out=stats[2 * i + 1])
            else:
                self._lin(f, blk.fc2, residual=x, out=x)      # enc_norm reads the fp32 stream
            if i in self.cfg.interaction_indexes:
                keep[i] = x
                frozen = True
                self._cap(f"enc{i}", x)
                ev = torch.cuda.Event()
                ev.record()
                self._keep_ev[i] = ev
        return x, keep

    # ---- DPT heads (heads/dpt_head.py:36-79, dpt_gs_head.py:121-171, dpt_block.py) ------------------------------------
    def _rcu(self, x, unit):
        r = self._relu_op(x)
        t = self._conv(r, unit[0], 3, ar=True, ro=True, pad=1, act=ACT_RELU)
        return self._conv(t, unit[1], 3, ar=True, pad=1, residual=x)

    def _fusion(self, rf, x0, x1=None, ro=False):
        out = x0
        if x1 is not None:
            res = self._rcu(x1, rf.r1)
            out = ops.eltwise(ELT_ADD, out, res)
        out = self._rcu(out, rf.r2)
        n, h, w_, c = out.shape
        out = self._resize_op(out, 2 * h, 2 * w_, True)
        return self._conv(out, rf.out_conv, 1, ar=True, ro=ro)

    def _dpt_trunk(self, hw, toks, B, N, gh, gw, p1_ro=True):
        """toks: 4 token tensors [B*N, C] (trailing intrinsics token per image is skipped) -> path_1 [B, 8gh, 8gw, 256]."""
        P = gh * gw
        layers = []
        for i in range(4):
            C_in = toks[i].shape[1]
            cw = hw.act_conv[i]
            # (h3: layer 3 goes through a strided conv, whose im2col reads fp32)
            o = ops.Split.empty(B, gh, gw, cw.N, device=self.dev) if (self.S and i != 3) else torch.empty(B, gh, gw, cw.N, device=self.dev)
            for b in range(B):
                self._lin(toks[i][b * N: b * N + P], cw, ro=not (self.S and i == 3), out=o[b].view(P, cw.N))
            del C_in
            layers.append(o)
        # act_postprocess tails
        g0 = self._lin(layers[0].view(B * P, -1), hw.act_up0, ar=True)
        layers[0] = ops.pixel_shuffle(g0, B, gh, gw, hw.act_up0.N // (hw.s0 * hw.s0), hw.s0)
        g1 = self._lin(layers[1].view(B * P, -1), hw.act_up1, ar=True)
        layers[1] = ops.pixel_shuffle(g1, B, gh, gw, hw.act_up1.N // (hw.s1 * hw.s1), hw.s1)
        layers[3] = self._conv(layers[3], hw.act_down3, 3, ro=True, stride=2, pad=1)
        layers = [self._conv(layers[i], hw.layer_rn[i], 3, ar=(i >= 2), pad=1) for i in range(4)]
        p4 = self._fusion(hw.refine[3], layers[3])
        p3 = self._fusion(hw.refine[2], p4, layers[2])
        p2 = self._fusion(hw.refine[1], p3, layers[1])
        p1 = self._fusion(hw.refine[0], p2, layers[0], ro=p1_ro)  # path_1 only feeds a conv (centre head) or a resize (GS head)
        return p1

    def _center_head(self, hw, toks, B, N, gh, gw, dst):
        """dst[i]: [S0*S1, 3] slice of Gaussians.means that receives the pts3d of batch entry i."""
        p1 = self._dpt_trunk(hw, toks, B, N, gh, gw)
        x = self._conv(p1, hw.head0, 3, ar=True, pad=1)
        n, h, w_, c = x.shape
        x = self._resize_op(x, 2 * h, 2 * w_, True)
        x = self._conv(x, hw.head2, 3, ar=True, ro=True, pad=1, act=ACT_RELU)
        S0, S1 = 2 * h, 2 * w_
        xyz = torch.empty(B * S0 * S1, 4, device=self.dev)
        self._lin(x.view(-1, x.shape[-1]), hw.head4, ar=True, out=xyz[:, :3])
        for b in range(B):  # pts3d written straight into Gaussians.means[b, v]
            ops._lib.check(ops._lib.load().siu3r_depth_exp(xyz[b * S0 * S1:].data_ptr(), 4, dst[b].data_ptr(), S0 * S1, ops._stream()),
                           "depth_exp")

    def _gs_head(self, hw, toks, img4, B, N, gh, gw):
        """-> raw Gaussian parameters [B, S*S, 83] (model.py:195-210)."""
        rows = None
        if self.S and img4.shape[2] % 16 == 0:
            # the 7x7 image conv runs as a 7x1 implicit GEMM over per-pixel packed filter rows (7 taps x 4 channels -> 32 wide); the packing only needs
            # the image, so it is enqueued before the trunk instead of sitting between the trunk and the full-resolution convs
            rows = ops.im2col_h3(img4, 1, 7, 1, 0, 3, 32).view(img4.shape[0], img4.shape[1], img4.shape[2], 32)
        p1 = self._dpt_trunk(hw, toks, B, N, gh, gw, p1_ro=not self.S)   # h3: the bilinear upsampling below reads fp32
        n, h, w_, c = p1.shape
        s = ops.conv_kxk_up2x(img4, hw.merger, 7, 7, p1, act=ACT_RELU, round_out=True) if self.R else None   # fused: no [S,S,256] upsampled map
        if s is None:
            up = ops.resize_bilinear(p1, 2 * h, 2 * w_, True)
            if rows is not None:
                s = ops.conv2d(rows, hw.merger.rowpacked(7, 7, 4), 7, 1, pad=(3, 0), act=ACT_RELU, residual=up, precision=self.prec, round_out=True)
            else:
                s = self._conv(img4, hw.merger, 7, ro=True, pad=3, act=ACT_RELU, residual=up)
        t = self._conv(s, hw.head0, 3, ar=True, ro=True, pad=1, act=ACT_RELU)
        raw = torch.empty(B, 4 * h * w_, 83, device=self.dev)
        self._lin(t.view(-1, 256), hw.head4, ar=True, out=raw.view(-1, 83))
        return raw

    # ---- ViT adapter (vit_adapter/vit_adapter.py:393-441) ---------------------------------------------------------------
    def _extractor(self, ex, c, featn_src, k, B, N, P, gh, gw):
        """c [B*Lq, 1024] updated in place; featn_src: [B*N, 1024] encoder block output (token rows incl. the intrinsics token)."""
        C = 1024
        Lq = c.shape[0] // B
        qn = self._ln(c, ex.qn, 1e-6)
        fn = ops.Split.empty(B * P, C, device=self.dev) if self.S else torch.empty(B * P, C, device=self.dev)
        for b in range(B):
            self._ln(featn_src[b * N: b * N + P], ex.fn, 1e-6, out=fn[b * P:(b + 1) * P])
        value = self._lin(fn, ex.value, ar=True)
        ow = self._lin(qn, ex.ow, ar=True)
        samp = torch.empty(B * Lq, C, device=self.dev)
        ops.msdeform_attn(value, P, ow, k.ad_ref, [(gh, gw)], 4, B, Lq, 16, 64, samp, round_out=self.R)
        self._lin(samp, ex.out, ar=True, residual=c, out=c)
        t = self._ln(c, ex.ffn_norm, 1e-6)
        t1 = self._lin(t, ex.fc1, ar=True)  # [B*Lq, 256]
        dw = torch.empty_like(t1)
        n = Lq // 21
        off = 0
        for (hh, ww, cnt) in ((2 * gh, 2 * gw, 16 * n), (gh, gw, 4 * n), (gh // 2, gw // 2, Lq - 20 * n)):
            ops.dwconv3x3(t1.data_ptr() + 4 * off * 256, 256, Lq * 256, B, hh, ww, 256, ex.dw, ex.db, dw.data_ptr() + 4 * off * 256, 256, Lq * 256, True)
            off += cnt
        self._lin(dw, ex.fc2, residual=c, out=c)

    def _adapter_stem(self, img4):
        """SpatialPriorModule convs (vit_adapter/blocks.py:200-262): depends on the image only."""
        a = self.w.adapter
        x = self._conv(img4, a.stem[0], 3, ro=True, stride=2, pad=1, act=ACT_RELU)
        x = self._conv(x, a.stem[1], 3, ar=True, ro=True, pad=1, act=ACT_RELU)
        f32 = not self.S    # h3: max-pool and the strided convs (im2col) read fp32, so their inputs stay fp32
        x = self._conv(x, a.stem[2], 3, ar=True, ro=f32, pad=1, act=ACT_RELU)
        c1 = ops.maxpool3x3s2(x)                                           # [Bn, S/4, S/4, 64] (max of rounded values stays rounded)
        c2 = self._conv(c1, a.conv2, 3, ro=f32, stride=2, pad=1, act=ACT_RELU)     # S/8, 128
        c3 = self._conv(c2, a.conv3, 3, ro=f32, stride=2, pad=1, act=ACT_RELU)     # S/16, 256
        c4 = self._conv(c3, a.conv4, 3, ro=f32, stride=2, pad=1, act=ACT_RELU)     # S/32, 256
        c1 = self._conv(c1, a.fc1, 1, ar=True)                             # [Bn, S/4, S/4, 1024]
        return c1, c2, c3, c4

    def _adapter(self, img4, feats, k, B, N, gh, gw, out_slots, V=2):
        """img4 [V*B,S0,S1,4] view-major (image j = v*B + b); feats: {block idx: [V*B*N, 1024]} -> f1..f4 of every image into
        out_slots[l][b*V+v].  All views go through every kernel as one batch."""
        a = self.w.adapter
        P = gh * gw
        Bn = V * B
        c1, c2, c3, c4 = self._adapter_stem(img4)
        n2, n3, n4 = 4 * P, P, P // 4
        Lq = n2 + n3 + n4
        c = torch.empty(Bn, Lq, 1024, device=self.dev)
        for b in range(Bn):  # fc2..4 (+ level embed folded into the bias) written straight into the concatenated query buffer
            self._lin(c2[b].view(n2, -1), a.fc[0], ar=True, out=c[b, :n2])
            self._lin(c3[b].view(n3, -1), a.fc[1], ar=True, out=c[b, n2:n2 + n3])
            self._lin(c4[b].view(n4, -1), a.fc[2], ar=True, out=c[b, n2 + n3:])
        c = c.view(Bn * Lq, 1024)
        for i, exs in enumerate(a.inter):
            idx = self.cfg.interaction_indexes[i]
            if idx in self._keep_ev:
                torch.cuda.current_stream().wait_event(self._keep_ev[idx])
            for ex in exs:
                self._extractor(ex, c, feats[idx], k, Bn, N, P, gh, gw)
        c = c.view(Bn, Lq, 1024)
        # c1 = up(c2) + c1
        c2d = torch.empty(Bn, n2, 1024, device=self.dev)
        for b in range(Bn):
            ops.rows_affine(c[b, :n2], out=c2d[b])
        g = self._lin(c2d.view(Bn * n2, 1024), a.up)
        c1 = ops.pixel_shuffle(g, Bn, 2 * gh, 2 * gw, 1024, a.up_s, add=c1)
        maps = [c1, c2d.view(Bn, 2 * gh, 2 * gw, 1024), None, None]
        c3d = torch.empty(Bn, gh, gw, 1024, device=self.dev)
        c4d = torch.empty(Bn, gh // 2, gw // 2, 1024, device=self.dev)
        for b in range(Bn):
            ops.rows_affine(c[b, n2:n2 + n3], out=c3d[b].view(n3, 1024))
            ops.rows_affine(c[b, n2 + n3:], out=c4d[b].view(n4, 1024))
        maps[2], maps[3] = c3d, c4d
        # + bilinear-resized ViT features (align_corners=False), then eval-mode BatchNorm
        for l, (scale_hw, idx) in enumerate(zip(((4 * gh, 4 * gw), (2 * gh, 2 * gw), (gh, gw), (gh // 2, gw // 2)), self.cfg.interaction_indexes)):
            src = feats[idx]
            for j in range(Bn):
                v, b = divmod(j, B)
                xb = src[j * N: j * N + P].view(1, gh, gw, 1024)
                if l == 2:
                    ops.eltwise(ELT_ADD, maps[l][j].view(P, 1024), xb.view(P, 1024).contiguous(), out=maps[l][j].view(P, 1024))
                else:
                    ops.resize_bilinear(xb, scale_hw[0], scale_hw[1], False, out=maps[l][j:j + 1], accumulate=True)
                sc, sh = a.bn[l]
                rows = scale_hw[0] * scale_hw[1]
                ops.rows_affine(maps[l][j].view(rows, 1024), scale=sc, shift=sh, out=out_slots[l][b * V + v].view(rows, 1024))

    # ---- Mask2Former (mask2former/video_seg_decoder.py:2072-2196, 1506-1575, 1204-1360) ---------------------------------
    def _m2f(self, feats, k, B, S0, S1, T=2):
        """feats: 4 maps [B*T, h, w, 1024] at strides 4/8/16/32 (frame index bt = b*T + t).  Returns class logits [B,100,21],
        mask logits pixel-major [B*T, S0/4, S1/4, 100]."""
        m = self.w.m2f
        E, Q = 256, self.cfg.num_queries
        BT = B * T
        lv = k.m2f_shapes
        Ltot = sum(h * w for h, w in lv)
        starts = [0, lv[0][0] * lv[0][1], lv[0][0] * lv[0][1] + lv[1][0] * lv[1][1]]
        x = torch.empty(BT, Ltot, E, device=self.dev)
        for i, f in enumerate((feats[3], feats[2], feats[1])):  # features[::-1][:3]
            cw, gw_, gb_ = m.in_proj[i]
            h, w_ = lv[i]
            e = self._conv(f, cw, 1)
            e = ops.groupnorm(e.view(BT, h * w_, E), 32, gw_, gb_, 1e-5, False)
            for bt in range(BT):
                ops.rows_affine(e[bt], out=x[bt, starts[i]:starts[i] + h * w_])
        x = x.view(BT * Ltot, E)
        # h3: LayerNorm outputs that are also residuals stay fp32 (their GEMM consumers split them)
        lnro = not self.S
        for li_, lyr in enumerate(m.enc):
            q = self._add_op(x, k.m2f_pos)
            value = self._lin(x, lyr.value, ar=li_ > 0)   # x is a (rounded) LayerNorm output from the second layer on
            ow = self._lin(q, lyr.ow, ar=True)
            samp = torch.empty(BT * Ltot, E, device=self.dev)
            ops.msdeform_attn(value, Ltot, ow, k.m2f_ref, lv, 4, BT, Ltot, 8, 32, samp, round_out=self.R)
            y = self._lin(samp, lyr.out, ar=True, residual=x)
            x = self._ln(y, lyr.ln1, 1e-5, ro=lnro)
            f1 = self._lin(x, lyr.fc1, ar=True, ro=True, act=ACT_RELU)
            y = self._lin(f1, lyr.fc2, ar=True, residual=x)
            x = self._ln(y, lyr.ln2, 1e-5, ro=lnro)
        x = x.view(BT, Ltot, E)
        # FPN level (stride 4)
        h4, w4 = S0 // 4, S1 // 4
        cw, gw_, gb_ = m.lateral
        cur = ops.groupnorm(self._conv(feats[0], cw, 1).view(BT, h4 * w4, E), 32, gw_, gb_, 1e-5, False).view(BT, h4, w4, E)
        h2, w2 = lv[2]
        for bt in range(BT):
            ops.resize_bilinear(x[bt, starts[2]:].view(1, h2, w2, E), h4, w4, False, out=cur[bt:bt + 1], accumulate=True)
        cw, gw_, gb_ = m.output
        o = self._conv(cur, cw, 3, pad=1)
        o = ops.groupnorm(o.view(BT, h4 * w4, E), 32, gw_, gb_, 1e-5, True).view(BT, h4, w4, E)
        mask_feat = self._conv(o, m.mask_proj, 1, ro=True)  # [BT, h4, w4, 256]; only ever the A operand of the mask GEMMs
        self._cap("m2f_mask_features", mask_feat)
        self._cap("m2f_tokens", x)
        # transformer module: keys of level i for batch b = frames (t) x pixels, + level embedding
        src, srcpos = [], []
        for i, (h, w_) in enumerate(lv):
            n = h * w_
            s = torch.empty(B, T * n, E, device=self.dev)
            for b in range(B):
                for t in range(T):
                    ops.rows_affine(x[b * T + t, starts[i]:starts[i] + n], shift=k.tm_lvl[i], out=s[b, t * n:(t + 1) * n])
            s = s.view(B * T * n, E)
            srcpos.append(self._add_op(s, k.tm_pos[i]))
            src.append(ops.split(s) if self.S else (ops.round_tf32(s) if self.R else s))   # keys / values of all 9 decoder layers: rounded / split once
        hidden, qpos = k.hidden0, k.qpos
        mf = mask_feat.view(B, T * h4 * w4, E)

        def predict(hid, target_hw):
            inter = self._ln(hid, m.dec_ln, 1e-5)
            e = self._lin(inter, m.mask_mlp[0], ar=True, ro=True, act=ACT_RELU)
            e = self._lin(e, m.mask_mlp[1], ar=True, ro=True, act=ACT_RELU)
            e = self._lin(e, m.mask_mlp[2], ar=True)  # [B*Q, 256]
            logits = torch.empty(B, T * h4 * w4, Q, device=self.dev)
            for b in range(B):  # einsum("bqc,btchw->bqthw") as a per-batch GEMM, pixel-major output
                wq = ops.Weight(e[b * Q:(b + 1) * Q], None, self.prec)
                self._lin(mf[b], wq, ar=True, out=logits[b])
            amask = None
            if target_hw is not None:
                amask = ops.attn_mask_from_logits(logits, B, T, h4, w4, Q, target_hw[0], target_hw[1])
            return inter, logits, amask

        inter, logits, amask = predict(hidden, lv[0])
        nh = 8
        for idx, lyr in enumerate(m.dec):
            li = idx % 3
            n = lv[li][0] * lv[li][1] * T
            # masked cross-attention (post-norm)
            qin = self._add_op(hidden, qpos)
            qh = self._lin(qin, lyr.cq, ar=True)
            kh = self._lin(srcpos[li], lyr.ck, ar=True)
            vh = self._lin(src[li], lyr.cv, ar=True)
            att = torch.empty(B * Q, E, device=self.dev)
            ops.attn_small_d32(qh, Q * E, E, kh, n * E, E, vh, n * E, E, att, Q * E, E, amask, B, nh, Q, n, 32 ** -0.5, round_out=self.R)
            y = self._lin(att, lyr.cout, ar=True, residual=hidden)
            hidden = self._ln(y, lyr.cln, 1e-5, ro=lnro)
            # query self-attention
            qin = self._add_op(hidden, qpos)
            qk = self._lin(qin, lyr.sqk, ar=True)          # [B*Q, 512] = [q | k]
            vv = self._lin(hidden, lyr.sv, ar=True)
            att = torch.empty(B * Q, E, device=self.dev)
            ops.attn_small_d32(qk, Q * 2 * E, 2 * E, qk[:, E:], Q * 2 * E, 2 * E, vv, Q * E, E, att, Q * E, E, None, B, nh, Q, Q, 32 ** -0.5,
                               round_out=self.R)
            y = self._lin(att, lyr.sout, ar=True, residual=hidden)
            hidden = self._ln(y, lyr.sln, 1e-5, ro=lnro)
            # FFN
            f1 = self._lin(hidden, lyr.fc1, ar=True, ro=True, act=ACT_RELU)
            y = self._lin(f1, lyr.fc2, ar=True, residual=hidden)
            hidden = self._ln(y, lyr.fln, 1e-5, ro=lnro)
            last = idx == len(m.dec) - 1
            inter, logits, amask = predict(hidden, None if last else lv[(idx + 1) % 3])
        cls = self._lin(inter, m.cls, ar=True)  # [B*Q, 21]
        return cls.view(B, Q, -1), logits.view(B * T, h4, w4, Q)

    # ---- panoptic post-process (image_processing_video_mask2former.py:1238-1481, model.py:231-312) ------------------------
    def _post_process(self, cls_logits, mask_logits, B, S0, S1, lift, T=2):
        cfg = self.cfg
        Q = cfg.num_queries
        num_labels = cls_logits.shape[-1] - 1
        h4, w4 = mask_logits.shape[1], mask_logits.shape[2]
        probs256 = ops.eltwise(ELT_SIGMOID, ops.resize_bilinear(mask_logits, 256, 256, False))  # [B*T,256,256,Q] (:1298-1312)
        # tiny host-side decision data: class probabilities of 100 queries
        cp = torch.softmax(cls_logits.detach().float().cpu(), dim=-1)
        scores_all, labels_all = cp.max(-1)
        fuse = set(cfg.label_ids_to_fuse)
        seg_masks, seg_infos, qc_list, qscore_list = [], [], [], []
        sem_all = torch.zeros(B, T, S0, S1, dtype=torch.int32, device=self.dev)
        inst_all = torch.zeros(B, T, S0, S1, dtype=torch.int32, device=self.dev)
        for b in range(B):
            keep = (labels_all[b] != num_labels) & (scores_all[b] > cfg.seg_threshold)
            kidx = torch.nonzero(keep).flatten()
            if kidx.numel() == 0:
                seg_masks.append(torch.zeros(T, S0, S1, device=self.dev) - 1)
                seg_infos.append([])
                qc = torch.zeros(T * S0 * S1, 1, num_labels + 1, device=self.dev)
                qc[:, 0, -1] = 1
                qc_list.append(qc)
                qscore_list.append([0.0])
                continue
            kscore = scores_all[b][kidx].contiguous()
            klabel = labels_all[b][kidx]
            sel = ops.resize_select(probs256[b * T:(b + 1) * T], kidx.to(torch.int32).to(self.dev), S0, S1)  # [T,S0,S1,qk]
            labels, area, orig = ops.argmax_area(sel, kscore.to(self.dev), 0.5)
            area_h, orig_h = area.cpu(), orig.cpu()
            segments, keep_q, keep_scores = [], [], []
            seg_lut = torch.zeros(kidx.numel(), dtype=torch.int32)
            sem_lut = torch.zeros(kidx.numel(), dtype=torch.int32)
            current, stuff_memory = 0, {}
            for j in range(kidx.numel()):
                pred_class = int(klabel[j])
                should_fuse = pred_class in fuse
                a_k, a_o = int(area_h[j]), int(orig_h[j])
                exists = a_k > 0 and a_o > 0
                if exists and not (torch.tensor(a_k) / torch.tensor(a_o)).item() > 0.8:
                    exists = False
                if not exists:
                    continue
                if pred_class in stuff_memory:
                    fuse_id = stuff_memory[pred_class]
                else:
                    current += 1
                    fuse_id = current
                sid = fuse_id if should_fuse else current
                score = round(float(kscore[j]), 6)
                segments.append({"id": sid, "label_id": pred_class, "was_fused": should_fuse, "score": score})
                seg_lut[j] = sid
                sem_lut[j] = pred_class + 1
                keep_q.append(j)
                keep_scores.append(score)
                if should_fuse and pred_class not in stuff_memory:
                    stuff_memory[pred_class] = current
            seg = ops.label_lut(labels, seg_lut.to(self.dev), sem_lut.to(self.dev), sem_out=sem_all[b], inst_out=inst_all[b])
            seg_masks.append(seg.view(T, S0, S1))
            seg_infos.append(segments)
            if keep_q:
                kq = torch.tensor(keep_q, dtype=torch.int32, device=self.dev)
                qc = ops.qc_logits(sel, kq, cp[b][kidx][keep_q].contiguous().to(self.dev)) if lift else None
            else:  # (:1468-1472) note the reference uses the *mask-logit* height/width here
                qc = torch.zeros(T, 1, num_labels + 1, h4, w4, device=self.dev)
                qc[:, 0, -1] = 1
                qc = qc.permute(0, 3, 4, 1, 2).reshape(T * h4 * w4, 1, num_labels + 1)
            qc_list.append(qc)
            qscore_list.append(keep_scores)
        return seg_masks, seg_infos, qc_list, qscore_list, sem_all.view(B, -1), inst_all.view(B, -1)

    # ---- forward --------------------------------------------------------------------------------------------------
    def enable_cuda_graph(self):
        """Replay the device part of forward() (everything up to the mask / class logits) from a CUDA graph: one capture per
        (batch, image size).  The returned tensors then alias static buffers that the next forward() overwrites."""
        self._use_graph = True
        self._graphs = {}

    def disable_cuda_graph(self):
        self._use_graph = False

    def _lin2(self, xs, wts, ar=False, ro=False, **kw):
        """Two same-shape linear layers (the two decoder streams) in one grouped launch."""
        return ops.gemm_group2(xs, wts, precision=self.prec, a_rounded=ar and self.R, round_out=ro and self.R, **kw)

    def _dec_layer(self, l, f, B, N, V):
        """One decoder layer for ALL views (backbone_croco.py:244-250 for V = 2, :514-531 for V > 2).  f: previous-layer tokens
        [V*B*N, 768], view-major.  Stream 0 = view 0 through dec_blocks[l], stream 1 = views 1..V-1 (one batch) through
        dec_blocks2[l]; every projection of the two streams is ONE grouped GEMM launch, self- and cross-attention run as one flash
        launch over all V*B images.  The cross-attention memory of an image of view i is the concatenation, in view order, of
        norm_y(tokens) of the same sample's other views (generate_ctx_views :500-506), keys rotated with their own positions;
        norm_y + the k|v projection are per token, so they run once per (stream weights, needed view).  Returns a new buffer."""
        if self.S:
            return self._dec_layer_h3(l, f, B, N, V)
        C, nh = 768, 12
        blks = (self.w.dec[0][l], self.w.dec[1][l])
        R0, R = B * N, V * B * N
        rows = ((0, R0), (R0, R))
        pos, tab = self._k.pos_enc, self._k.rope_tab        # positions are the same for every image
        split = lambda t: [t[a:b] for a, b in rows]
        # ---- self-attention ----
        ln2 = lambda xs, wbs, outs: ops.layernorm_group2(xs, wbs, 1e-6, outs, round_out=self.R)   # both streams' norms in one launch
        h = torch.empty(R, C, device=self.dev)
        ln2(split(f), [bk.n1 for bk in blks], split(h))
        def vt_windows():
            """V^T buffer [C, ld] for the grouped projections: stream g owns a 16-byte aligned column window; all images must end up at one
            uniform column stride (possible for B = 1 with a padded stride, or when B*N is a multiple of 4)."""
            if not self.R:
                return None
            if V == 2 and B == 1:
                stride = (N + 3) // 4 * 4
                buf = torch.empty(C, 2 * stride, device=self.dev)
                return buf, [buf[:, :stride], buf[:, stride:]], [stride, stride], stride
            if V == 2 and (B * N) % 4 == 0:
                buf = torch.empty(C, 2 * B * N, device=self.dev)
                return buf, [buf[:, :B * N], buf[:, B * N:]], [B * N, B * N], N
            return None

        qkv = torch.empty(R, 3 * C, device=self.dev)
        vw, st = vt_windows(), {}
        self._lin2(split(h), [bk.qkv for bk in blks], outs=split(qkv), ar=True, ro=True, rope=(pos, tab, 2 * C),
                   **({"vt": (vw[1], vw[2], 2 * C, st)} if vw else {}))
        att = torch.empty(R, C, device=self.dev)
        if self.R:
            ops.flash_attn_tc(qkv, 0, N * 3 * C, 3 * C, 3 * C, qkv, C, N * 3 * C, 3 * C, 3 * C, qkv, 2 * C, N * 3 * C, 3 * C, att, V * B, nh, N, N, 0.125,
                              round_out=True, vt=(vw[0], vw[3]) if st.get("ok") else None)
        else:
            ops.flash_attn_d64(qkv, 0, N * 3 * C, 3 * C, qkv, C, N * 3 * C, 3 * C, qkv, 2 * C, N * 3 * C, 3 * C, att, V * B, nh, N, N, 0.125, self.prec)
        x1 = torch.empty(R, C, device=self.dev)
        self._lin2(split(att), [bk.proj for bk in blks], outs=split(x1), ar=True, residuals=split(f))
        # ---- cross-attention memory: k|v of the other views ----
        if V == 2:
            # stream 0 (view 0) attends to view 1, stream 1 (view 1) to view 0: memory rows are already image-aligned
            yn = torch.empty(R, C, device=self.dev)
            ln2([f[R0:], f[:R0]], [bk.ny for bk in blks], split(yn))
            ctx = torch.empty(R, 2 * C, device=self.dev)
            vw2, st2 = vt_windows(), {}
            self._lin2(split(yn), [bk.ckv for bk in blks], outs=split(ctx), ar=True, ro=True, rope=(pos, tab, C),
                       **({"vt": (vw2[1], vw2[2], C, st2)} if vw2 else {}))
            Nk = N
        else:
            yn0 = torch.empty(R - R0, C, device=self.dev)     # views 1..V-1 under stream-0 weights
            yn1 = torch.empty(R, C, device=self.dev)          # all views under stream-1 weights
            ln2([f[R0:], f], [bk.ny for bk in blks], [yn0, yn1])
            kv0 = torch.empty(R - R0, 2 * C, device=self.dev)
            kv1 = torch.empty(R, 2 * C, device=self.dev)
            self._lin2([yn0, yn1], [bk.ckv for bk in blks], outs=[kv0, kv1], ar=True, ro=True, rope=(pos, tab, C))
            Nk = (V - 1) * N
            vw2, st2 = None, {}
            ctx = torch.empty(V * B, Nk, 2 * C, device=self.dev)
            for i in range(V):
                for b in range(B):
                    slot = 0
                    for j in range(V):
                        if j == i:
                            continue
                        src = kv0[((j - 1) * B + b) * N:][:N] if i == 0 else kv1[(j * B + b) * N:][:N]
                        ops.rows_affine(src, out=ctx[i * B + b, slot * N:(slot + 1) * N])
                        slot += 1
        h2 = torch.empty(R, C, device=self.dev)
        ln2(split(x1), [bk.n2 for bk in blks], split(h2))
        q = torch.empty(R, C, device=self.dev)
        self._lin2(split(h2), [bk.cq for bk in blks], outs=split(q), ar=True, ro=True, rope=(pos, tab, C))
        a2 = torch.empty(R, C, device=self.dev)
        if self.R:
            ops.flash_attn_tc(q, 0, N * C, C, C, ctx, 0, Nk * 2 * C, 2 * C, 2 * C, ctx, C, Nk * 2 * C, 2 * C, a2, V * B, nh, N, Nk, 0.125, round_out=True,
                              vt=(vw2[0], vw2[3]) if (vw2 and st2.get("ok")) else None)
        else:
            ops.flash_attn_d64(q, 0, N * C, C, ctx, 0, Nk * 2 * C, 2 * C, ctx, C, Nk * 2 * C, 2 * C, a2, V * B, nh, N, Nk, 0.125, self.prec)
        self._lin2(split(a2), [bk.cproj for bk in blks], outs=split(x1), ar=True, residuals=split(x1))
        # ---- MLP ----
        h3 = torch.empty(R, C, device=self.dev)
        ln2(split(x1), [bk.n3 for bk in blks], split(h3))
        m = torch.empty(R, 4 * C, device=self.dev)
        self._lin2(split(h3), [bk.fc1 for bk in blks], outs=split(m), ar=True, ro=True, act=ACT_GELU)
        self._lin2(split(m), [bk.fc2 for bk in blks], outs=split(x1), ar=True, residuals=split(x1))
        return x1

    def _dec_layer_h3_fused(self, l, f, B, N):
        """_dec_layer_h3 for V = 2 with norm1 / norm_y / norm2 / norm3 fused into the projections that consume them (siu3r_gemm_h3_ln): the
        residual-adding projections emit fp32 rows + plane pair + row statistics, the consumers read the raw plane pair."""
        C, nh, V = 768, 12, 2
        dev = self.dev
        blks = (self.w.dec[0][l], self.w.dec[1][l])
        R0, R = B * N, V * B * N
        rows = ((0, R0), (R0, R))
        pos, tab = self._k.pos_enc, self._k.rope_tab
        sp = lambda t: [t[a:b] for a, b in rows]
        W0 = (R0 + 7) // 8 * 8
        W1 = (R - R0 + 7) // 8 * 8
        st = self._dec_fused
        fs, s_in, sA, sB, sC = st["fs"], st["stats"][3 * l], st["stats"][3 * l + 1], st["stats"][3 * l + 2], st["stats"][3 * l + 3]

        def vt_buf():
            buf = ops.Split.empty(C, W0 + W1, device=dev, unscaled=True)
            return buf, [buf[:, :W0], buf[:, W0:]]
        # ---- self-attention: norm1 inside the qkv projection ----
        qkv = ops.Split.empty(R, 3 * C, device=dev, unscaled=True)
        vb, wins = vt_buf()
        self._lin2(sp(fs), [bk.qkv_ln for bk in blks], outs=sp(qkv), ro=True, rope=(pos, tab, 2 * C), vt=(wins, [W0, W1], 2 * C, {}), unscaled=True,
                   ln_stats=sp(s_in))
        att = ops.flash_attn_h3(qkv, 0, qkv, C, vb, N, V * B, nh, N, N, 0.125, split_out=True, vt_b_split=B, vt_extra=W0 - R0)
        x1 = torch.empty(R, C, device=dev)
        x1s = ops.Split.empty(R, C, device=dev)
        self._lin2(sp(att), [bk.proj for bk in blks], outs=sp(x1s), outs_f32=sp(x1), residuals=sp(f), stats_out=sp(sA))
        # ---- cross-attention memory: norm_y inside the k|v projection (stream 0 reads view 1's rows and vice versa) ----
        ctx = ops.Split.empty(R, 2 * C, device=dev, unscaled=True)
        vb2, wins2 = vt_buf()
        self._lin2([fs[R0:], fs[:R0]], [bk.ckv_ln for bk in blks], outs=sp(ctx), ro=True, rope=(pos, tab, C), vt=(wins2, [W0, W1], C, {}), unscaled=True,
                   ln_stats=[s_in[R0:], s_in[:R0]])
        q = ops.Split.empty(R, C, device=dev, unscaled=True)
        self._lin2(sp(x1s), [bk.cq_ln for bk in blks], outs=sp(q), ro=True, rope=(pos, tab, C), unscaled=True, ln_stats=sp(sA))
        a2 = ops.flash_attn_h3(q, 0, ctx, 0, vb2, N, V * B, nh, N, N, 0.125, split_out=True, vt_b_split=B, vt_extra=W0 - R0)
        x2s = ops.Split.empty(R, C, device=dev)
        self._lin2(sp(a2), [bk.cproj for bk in blks], outs=sp(x2s), outs_f32=sp(x1), residuals=sp(x1), stats_out=sp(sB))
        # ---- MLP: norm3 inside fc1 ----
        m = ops.Split.empty(R, 4 * C, device=dev)
        self._lin2(sp(x2s), [bk.fc1_ln for bk in blks], outs=sp(m), ro=True, act=ACT_GELU, ln_stats=sp(sB))
        if l + 1 < self.cfg.dec_depth:
            st["fs"] = ops.Split.empty(R, C, device=dev)
            self._lin2(sp(m), [bk.fc2 for bk in blks], outs=sp(st["fs"]), outs_f32=sp(x1), residuals=sp(x1), stats_out=sp(sC))
        else:
            self._lin2(sp(m), [bk.fc2 for bk in blks], outs=sp(x1), residuals=sp(x1))     # dec_norm reads the fp32 rows
        return x1

    def _dec_layer_h3(self, l, f, B, N, V):
        """_dec_layer in h3 mode: every GEMM operand is an fp16 plane pair written by its producer (LayerNorm, projection epilogue, attention
        epilogue); q / k / V^T of both attentions are unscaled plane pairs (flash_h3.cu)."""
        if getattr(self, "_dec_fused", None) is not None:
            return self._dec_layer_h3_fused(l, f, B, N)
        C, nh = 768, 12
        dev = self.dev
        blks = (self.w.dec[0][l], self.w.dec[1][l])
        R0, R = B * N, V * B * N
        rows = ((0, R0), (R0, R))
        pos, tab = self._k.pos_enc, self._k.rope_tab
        sp = lambda t: [t[a:b] for a, b in rows]
        W0 = (R0 + 7) // 8 * 8                 # V^T window of stream 0; stream 1 starts there (16-byte aligned)
        W1 = (R - R0 + 7) // 8 * 8

        def vt_buf():
            buf = ops.Split.empty(C, W0 + W1, device=dev, unscaled=True)
            return buf, [buf[:, :W0], buf[:, W0:]]
        # ---- self-attention ----
        h = ops.Split.empty(R, C, device=dev)
        ops.layernorm_h3(sp(f), [bk.n1 for bk in blks], 1e-6, outs=sp(h))
        qkv = ops.Split.empty(R, 3 * C, device=dev, unscaled=True)
        vb, wins = vt_buf()
        self._lin2(sp(h), [bk.qkv for bk in blks], outs=sp(qkv), ro=True, rope=(pos, tab, 2 * C), vt=(wins, [W0, W1], 2 * C, {}), unscaled=True)
        att = ops.flash_attn_h3(qkv, 0, qkv, C, vb, N, V * B, nh, N, N, 0.125, split_out=True, vt_b_split=B, vt_extra=W0 - R0)
        x1 = torch.empty(R, C, device=dev)
        self._lin2(sp(att), [bk.proj for bk in blks], outs=sp(x1), residuals=sp(f))
        # ---- cross-attention memory: k | v of the other views under the attending stream's weights ----
        if V == 2:
            yn = ops.Split.empty(R, C, device=dev)
            ops.layernorm_h3([f[R0:], f[:R0]], [bk.ny for bk in blks], 1e-6, outs=sp(yn))
            ctx = ops.Split.empty(R, 2 * C, device=dev, unscaled=True)
            vb2, wins2 = vt_buf()
            self._lin2(sp(yn), [bk.ckv for bk in blks], outs=sp(ctx), ro=True, rope=(pos, tab, C), vt=(wins2, [W0, W1], C, {}), unscaled=True)
            Nk = N
            kctx, vt2, vt_cols, vsplit, vextra = ctx, vb2, N, B, W0 - R0
        else:
            # V > 2: the memory of an image is the concatenation of V-1 views; it is assembled in fp32 (row copies), then split once
            yn0 = ops.Split.empty(R - R0, C, device=dev)
            yn1 = ops.Split.empty(R, C, device=dev)
            ops.layernorm_h3([f[R0:]], [blks[0].ny], 1e-6, outs=[yn0])
            ops.layernorm_h3([f], [blks[1].ny], 1e-6, outs=[yn1])
            kv0 = torch.empty(R - R0, 2 * C, device=dev)
            kv1 = torch.empty(R, 2 * C, device=dev)
            pos_all = pos.view(-1, 2)
            self._lin(yn0, blks[0].ckv, out=kv0, rope=(pos_all[:R - R0], tab, C))
            self._lin(yn1, blks[1].ckv, out=kv1, rope=(pos_all, tab, C))
            Nk = (V - 1) * N
            ctxf = torch.empty(V * B, Nk, 2 * C, device=dev)
            for i in range(V):
                for b in range(B):
                    slot = 0
                    for j in range(V):
                        if j == i:
                            continue
                        src = kv0[((j - 1) * B + b) * N:][:N] if i == 0 else kv1[(j * B + b) * N:][:N]
                        ops.rows_affine(src, out=ctxf[i * B + b, slot * N:(slot + 1) * N])
                        slot += 1
            ctxf = ctxf.view(V * B * Nk, 2 * C)
            kctx = ops.split(ctxf[:, :C], unscaled=True)
            vt2 = ops.transpose_v_h3(ctxf, C, Nk * 2 * C, 2 * C, V * B, Nk, nh)
            vt_cols, vsplit, vextra = 0, 0, 0
        h2 = ops.Split.empty(R, C, device=dev)
        ops.layernorm_h3(sp(x1), [bk.n2 for bk in blks], 1e-6, outs=sp(h2))
        q = ops.Split.empty(R, C, device=dev, unscaled=True)
        self._lin2(sp(h2), [bk.cq for bk in blks], outs=sp(q), ro=True, rope=(pos, tab, C), unscaled=True)
        a2 = ops.flash_attn_h3(q, 0, kctx, 0, vt2, vt_cols, V * B, nh, N, Nk, 0.125, split_out=True, vt_b_split=vsplit, vt_extra=vextra)
        self._lin2(sp(a2), [bk.cproj for bk in blks], outs=sp(x1), residuals=sp(x1))
        # ---- MLP ----
        h3_ = ops.Split.empty(R, C, device=dev)
        ops.layernorm_h3(sp(x1), [bk.n3 for bk in blks], 1e-6, outs=sp(h3_))
        m = ops.Split.empty(R, 4 * C, device=dev)
        self._lin2(sp(h3_), [bk.fc1 for bk in blks], outs=sp(m), ro=True, act=ACT_GELU)
        self._lin2(sp(m), [bk.fc2 for bk in blks], outs=sp(x1), residuals=sp(x1))
        return x1

    def _forward_device(self, imgs, Kin):
        try:
            return self._forward_device_impl(imgs, Kin)
        finally:
            ops._lib.load().siu3r_gemm_h3_cluster_cap(0)

    def _forward_device_impl(self, imgs, Kin):
        """All device work of SIU3RModel.forward / SIU3RMultiViewModel.forward up to (and excluding) the host-assisted panoptic
        post-process.  Images are batched view-major (image j = v*B + b) in both models."""
        B, V, _, S0, S1 = imgs.shape
        w, c = self.w, self.cfg
        k = self._k = self._consts(B, S0, S1, V)
        gh, gw = S0 // 16, S1 // 16
        P, N = gh * gw, gh * gw + 1
        Bn = V * B
        img4 = torch.zeros(Bn, S0, S1, 4, device=self.dev)  # view-major NHWC, 4th channel = 0
        lib = ops._lib.load()
        for v in range(V):
            for b in range(B):
                ops._lib.check(lib.siu3r_nchw_to_nhwc(imgs[b, v].data_ptr(), img4[v * B + b].data_ptr(), 1, 3, S0 * S1, 4, ops._stream()), "nchw_to_nhwc")
        # ---- encoder input: patch tokens + intrinsics token ----
        x = torch.empty(Bn, N, 1024, device=self.dev)
        if self.S:
            cols = ops.im2col_h3(img4, 16, 16, 16, 0, 0, 1024)
        else:
            cols = torch.empty(Bn * P, 1024, device=self.dev)
            ops._lib.check(lib.siu3r_im2col_nhwc(img4.data_ptr(), Bn, S0, S1, 4, 16, 16, 16, 0, 0, cols.data_ptr(), 1024, 1 if self.R else 0, ops._stream()),
                           "im2col")
        for i in range(Bn):
            self._lin(cols[i * P:(i + 1) * P], w.patch, ar=True, out=x[i, :P])
        x = x.view(Bn * N, 1024)
        Kflat = Kin.view(B, 9 * V)
        for v in range(V):  # intrinsics token = Linear(9 -> 1024) on the flattened K (backbone_croco.py:278-280, :546-549)
            ops.gemm_simt(Kflat[:, 9 * v: 9 * v + 9], w.intr_w, w.intr_b, out=x[v * B * N + P:: N][:B])
        # DAG: the panoptic chain (adapter -> Mask2Former) only needs the image and the four kept encoder outputs, so it runs
        # on its own stream next to the rest of the encoder, the decoder and the Gaussian heads (one graph branch when captured).
        self._keep_ev = {}
        serial = getattr(self, "serial", False)
        main = torch.cuda.current_stream()
        if not serial:
            if getattr(self, "_seg_stream", None) is None:
                # high priority: its tail (the Mask2Former decoder) is a chain of ~270 tiny dependent kernels that must not queue behind the heads
                self._seg_stream = torch.cuda.Stream(device=self.dev, priority=-1 if SEG_PRIORITY else 0)
            fork = torch.cuda.Event()
            fork.record(main)
        self._mark("start")
        x, keep = self._encoder(x, k.pos_enc, Bn, N)
        self._mark("encoder")
        shapes = [(S0 // 4, S1 // 4), (S0 // 8, S1 // 8), (S0 // 16, S1 // 16), (S0 // 32, S1 // 32)]
        ms = [torch.empty(B * V, h, w_, 1024, device=self.dev) for (h, w_) in shapes]

        def seg_chain():
            self._adapter(img4, keep, k, B, N, gh, gw, ms, V)
            self._cap("adapter_ms", ms)
            self._mark("adapter")
            r = self._m2f(ms, k, B, S0, S1, V)
            self._mark("m2f")
            return r

        if not serial:
            self._seg_stream.wait_event(fork)
            with torch.cuda.stream(self._seg_stream):
                cls_logits, mask_logits = seg_chain()
        # The persistent GEMMs of the decoder and the heads leave a few CTA pairs free for the high-priority panoptic chain running next to them
        # (a persistent kernel never retires a CTA early, so stream priority alone cannot get the chain's small kernels onto an SM).
        if not serial and self.S and HEAD_CLUSTER_CAP:
            lib.siu3r_gemm_h3_cluster_cap(HEAD_CLUSTER_CAP)     # reset by _forward_device
        feat = ops.layernorm(x, w.enc_norm[0], w.enc_norm[1], 1e-6)  # [V*B*N, 1024]
        self._cap("enc_norm", feat)
        # ---- decoder ----
        self._dec_fused = None
        if self.fuse_ln and V == 2:
            R = V * B * N
            self._dec_fused = {"stats": torch.zeros(1 + 3 * c.dec_depth, R, 2, device=self.dev, dtype=torch.int64),
                               "fs": ops.Split.empty(R, 768, device=self.dev)}
            f = torch.empty(R, 768, device=self.dev)
            self._lin(feat, w.dec_embed, out=self._dec_fused["fs"], out_f32=f, stats_out=self._dec_fused["stats"][0])
        else:
            f = self._lin(feat, w.dec_embed)  # [V*B*N, 768]
        # both decoder streams advance together: one grouped launch per projection, one flash launch per attention
        decs = [feat]
        for l in range(c.dec_depth):
            f = self._dec_layer(l, f, B, N, V)
            decs.append(f)
            self._cap(f"dec1_{l}", f[:B * N])
            self._cap(f"dec2_{l}", f[B * N:])
        decs[-1] = ops.layernorm(decs[-1], w.dec_norm[0], w.dec_norm[1], 1e-6)
        dec_first, dec_rest = [t[:B * N] for t in decs], [t[B * N:] for t in decs]
        self._mark("decoder")
        # ---- DPT heads: head1 on view 0, head2 on the other views (one batch); independent branches ----
        G1 = S0 * S1
        means = torch.empty(B, V, G1, 3, device=self.dev)
        hooks = [0, c.dec_depth * 2 // 4, c.dec_depth * 3 // 4, c.dec_depth]
        toks = [[dec_first[hk] for hk in hooks], [dec_rest[hk] for hk in hooks]]
        Br = (V - 1) * B
        dst = [[means[b, 0] for b in range(B)], [means[b, v] for v in range(1, V) for b in range(B)]]
        branches = [lambda: self._center_head(w.heads["downstream_head1"], toks[0], B, N, gh, gw, dst[0]),
                    lambda: self._gs_head(w.heads["gaussian_param_head1"], toks[0], img4[:B], B, N, gh, gw),
                    lambda: self._center_head(w.heads["downstream_head2"], toks[1], Br, N, gh, gw, dst[1]),
                    lambda: self._gs_head(w.heads["gaussian_param_head2"], toks[1], img4[B:], Br, N, gh, gw)]
        res = self._par(branches)
        raws = [res[1], res[3]]     # [B, G1, 83] (view 0), [(V-1)*B, G1, 83] (views 1.., view-major)
        self._cap("gs_raw", raws)
        cov = torch.empty(B, V * G1, 3, 3, device=self.dev)
        harm = torch.empty(B, V * G1, 3, 25, device=self.dev)
        opac = torch.empty(B, V * G1, device=self.dev)
        scales = torch.empty(B, V * G1, 3, device=self.dev)
        rots = torch.empty(B, V * G1, 4, device=self.dev)
        for v in range(V):
            for b in range(B):
                o = v * G1
                raw = raws[0][b] if v == 0 else raws[1][(v - 1) * B + b]
                ops._lib.check(lib.siu3r_gaussian_adapter(raw.data_ptr(), G1, cov[b, o:].data_ptr(), harm[b, o:].data_ptr(), opac[b, o:].data_ptr(),
                                                          scales[b, o:].data_ptr(), rots[b, o:].data_ptr(), ops._stream()), "gaussian_adapter")
        self._mark("heads")
        if serial:
            cls_logits, mask_logits = seg_chain()
        else:
            main.wait_stream(self._seg_stream)
        self._mark("end")
        return means.view(B, V * G1, 3), cov, harm, opac, scales, rots, cls_logits, mask_logits

    multiview = False   # SIU3RMultiViewModel sets True

    def _check_inputs(self, imgs):
        assert self._ready, "call load_state_dict(...).cuda() first"
        B, V, _, S0, S1 = imgs.shape
        if self.multiview:
            assert V >= 2, "SIU3RMultiViewModel needs at least two context views"
        else:
            assert V == 2, "two-view path; the V-view model is SIU3RMultiViewModel"
        assert S0 % 16 == 0 and S1 % 16 == 0, f"Input image size ({S0}x{S1}) is not a multiple of patch size (16)."
        assert (S0, S1) == tuple(self.cfg.image_size), "model was built for a different image_size (vit_adapter.py:328-329)"
        return B, V, S0, S1

    @torch.no_grad()
    def forward_async(self, context_views_images, context_views_intrinsics, slot: int = 0):
        """Enqueue the device part of forward() without waiting for it and return a handle for forward_finish().  With
        enable_cuda_graph() every `slot` has its own captured graph, static buffers and stream, so the forwards of consecutive pairs
        overlap on the GPU (serving.PairPipeline, bench.py): slot s may be re-submitted once its previous handle has been finished."""
        assert self._ready, "call load_state_dict(...).cuda() first"
        with torch.cuda.device(self.dev):
            return self._forward_async(context_views_images, context_views_intrinsics, slot)

    def _forward_async(self, context_views_images, context_views_intrinsics, slot):
        imgs = context_views_images
        B, V, S0, S1 = self._check_inputs(imgs)
        cur = torch.cuda.current_stream()
        if getattr(self, "_use_graph", False) and self.capture is None:
            key = (B, V, S0, S1, slot)
            if key not in self._graphs:
                si = torch.empty(B, V, 3, S0, S1, device=self.dev)
                sk = torch.empty(B, V, 3, 3, device=self.dev)
                si.copy_(imgs)
                sk.copy_(context_views_intrinsics)
                side = torch.cuda.Stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    self._forward_device(si, sk)  # warm-up: builds host tables, sets kernel attributes
                cur.wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                n0 = ops.launch_count()
                with torch.cuda.graph(graph):
                    outs = self._forward_device(si, sk)
                self._graphs[key] = (graph, si, sk, outs, ops.launch_count() - n0, torch.cuda.Stream(device=self.dev))  # kernels of ours per replay
            graph, si, sk, outs, nlaunch, st = self._graphs[key]
            st.wait_stream(cur)            # inputs produced / previous results of this slot consumed on the caller's stream
            with torch.cuda.stream(st):
                si.copy_(imgs, non_blocking=True)
                sk.copy_(context_views_intrinsics, non_blocking=True)
                graph.replay()
                done = torch.cuda.Event()
                done.record(st)
            ops._lib.load().siu3r_note_launch(nlaunch)
        else:
            imgs = imgs.to(self.dev, torch.float32).contiguous()
            Kin = context_views_intrinsics.to(self.dev, torch.float32).contiguous()
            outs = self._forward_device(imgs, Kin)
            done = None
        return (outs, done, (B, V, S0, S1))

    @torch.no_grad()
    def forward_finish(self, handle, enable_query_class_logit_lift=False):
        """Second half of forward(): waits for the device part of `handle` and runs the panoptic post-process (host decisions)."""
        with torch.cuda.device(self.dev):
            return self._forward_finish(handle, enable_query_class_logit_lift)

    def _forward_finish(self, handle, enable_query_class_logit_lift):
        outs, done, (B, V, S0, S1) = handle
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
        means, cov, harm, opac, scales, rots, cls_logits, mask_logits = outs
        gaussians = Gaussians(means=means, covariances=cov, harmonics=harm, opacities=opac, scales=scales, rotations=rots)
        h4, w4 = S0 // 4, S1 // 4
        seg_output = SimpleNamespace(class_queries_logits=cls_logits,
                                     masks_queries_logits=mask_logits.view(B, V, h4, w4, -1).permute(0, 4, 1, 2, 3))
        seg_masks, seg_infos, qc_list, qscores, sem, inst = self._post_process(cls_logits, mask_logits, B, S0, S1, enable_query_class_logit_lift, V)
        gaussians.semantic_labels = sem
        gaussians.instance_labels = inst
        if enable_query_class_logit_lift:
            gaussians.seg_query_class_logits = qc_list
            return gaussians, seg_output, seg_masks, seg_infos, qscores
        return gaussians, seg_output, seg_masks, seg_infos

    @torch.no_grad()
    def forward(self, context_views_images, context_views_intrinsics, mask_labels=None, class_labels=None, enable_query_class_logit_lift=False):
        assert mask_labels is None and class_labels is None, "training losses are out of scope (SURVEY.md section 8)"
        return self.forward_finish(self.forward_async(context_views_images, context_views_intrinsics), enable_query_class_logit_lift)


class SIU3RMultiViewModel(SIU3RModel):
    """V-view model (src/models/model_multi.py:28-392, backbone AsymmetricCroCoMulti backbone_croco.py:350-590): same weights
    and state_dict keys as SIU3RModel; view 0 is the reference view (dec_blocks / head1 / gaussian_param_head1), every other
    view goes through dec_blocks2 / head2 / gaussian_param_head2 and cross-attends to the V-1 other views."""
    multiview = True
