"""GPU parity tests of the h3 kernels (fp32-grade results on the fp16 tensor-core path: csrc/h3.cuh, gemm_h3.cu, flash_h3.cu) against plain
PyTorch references of the same ops (float64).  Tolerances are relative to the output scale.  The tensor core truncates its fp32 accumulator
on every MMA (tools/acc_probe.py), so the error of an output grows with the length of its MMA chain: measured ~2e-8 per chained MMA, i.e.
5e-6 at K = 4096 (256 MMAs), 4.5e-6 for a 3x3 conv over 768 channels (432), 6e-6 for attention over 1025 keys (204) and 1.6e-5 over 3075 keys.
End to end (tests/test_model_gpu.py, tests/test_fulltensor_gpu.py) that leaves the logits at 2e-5 rel and the Gaussians at 1.6e-4 abs,
5-6x inside the north-star tolerances.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
H3 = 4


@pytest.fixture(scope="module", autouse=True)
def _no_tf32_in_reference():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def test_split_roundtrip_and_range():
    from siu3r_b200 import ops
    x = rnd(777, 264, seed=1) * torch.logspace(-6, 3, 264, device=DEV)[None]
    x[0, :8] = torch.tensor([0.0, -0.0, 1e-9, -3e-8, 65504.0, 7e4, -1e6, 1.0], device=DEV)
    s = ops.split(x)
    y = s.float()
    ok = x.abs() <= 65504
    err = ((y - x).abs() / x.abs().clamp_min(1e-30))[ok & (x.abs() > 1e-4)]
    assert float(err.max()) < 2 ** -21, float(err.max())
    assert float((y - x).abs()[ok & (x.abs() <= 1e-4)].max()) < 1e-10      # tiny values: hi goes subnormal, the scaled lo plane keeps the rest
    assert torch.equal(y[0, 5:7], torch.tensor([65504.0, -65504.0], device=DEV))   # saturates instead of producing inf
    # strided destination inside a wider buffer
    big = ops.Split.empty(777, 300, device=DEV)
    ops.split(x, big[:, 16:280])
    assert torch.equal(big[:, 16:280].t, s.t)


@pytest.mark.parametrize("M,N,K", [(2050, 3072, 1024), (300, 768, 768), (100, 256, 2048), (128, 64, 32), (1025, 96, 1024), (77, 83, 256),
                                   (4096, 4096, 1024), (513, 21, 256), (2050, 1024, 4096), (5, 2048, 256), (100, 256, 260), (33, 40, 20)])
def test_gemm_h3_plain(M, N, K):
    from siu3r_b200 import ops
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    wt = ops.Weight(w, b, H3)
    ref = F.linear(x.double(), w.double(), b.double())
    y = ops.gemm(x, wt, precision=H3)
    assert y.dtype == torch.float32 and rel_err(y, ref) < 1e-5, (rel_err(y, ref), M, N, K)
    # operands that are already plane pairs, plane-pair result
    ys = ops.gemm(ops.split(x), wt, precision=H3, round_out=True)
    assert isinstance(ys, ops.Split) and rel_err(ys.float(), ref) < 1e-5
    # every legal token tile width gives the same answer (all tile / accumulator-buffering variants of the kernel)
    lib = ops._lib.load()
    try:
        for tw in (32, 48, 128, 144, 256):
            lib.siu3r_gemm_h3_force(tw)
            assert rel_err(ops.gemm(x, wt, precision=H3), ref) < 1e-5, tw
    finally:
        lib.siu3r_gemm_h3_force(0)


@pytest.mark.parametrize("M,N,K", [(2050, 128, 1024), (1000, 96, 256), (513, 21, 256), (70, 128, 64)])
def test_gemm_h3_half_m_mode_is_bit_identical(M, N, K):
    """N <= 128 runs M = 128 MMAs (64 weight rows per CTA, other TMEM layout); the k-order of the accumulation is the same, so the results
    must equal the M = 256 kernel's bit for bit -- linear and conv, fp32 and plane-pair outputs, every tile width."""
    from siu3r_b200 import ops
    lib = ops._lib.load()
    x, w, b = rnd(M, K, seed=21), rnd(N, K, seed=22, scale=K ** -0.5), rnd(N, seed=23)
    res = rnd(M, N, seed=24)
    wt = ops.Weight(w, b, H3)
    xs = ops.split(x)
    img = ops.split(rnd(2, 32, 48, 128, seed=25))
    wc = ops.Weight(rnd(N, 9 * 128, seed=26, scale=0.03), b, H3)
    rc = rnd(2, 32, 48, N, seed=27)
    outs = {}
    try:
        for mode in (0, 1):
            lib.siu3r_gemm_h3_set_mhalf(mode)
            for tw in (0, 32, 64, 128, 192, 256):
                lib.siu3r_gemm_h3_force(tw)
                a = ops.gemm(xs, wt, act=1, residual=res, precision=H3)
                s = ops.gemm(xs, wt, act=2, precision=H3, round_out=True).t.clone()
                c = ops.conv2d(img, wc, 3, 3, pad=1, act=2, residual=rc, precision=H3)
                outs[(mode, tw)] = (a, s, c)
    finally:
        lib.siu3r_gemm_h3_set_mhalf(1)
        lib.siu3r_gemm_h3_force(0)
    ref = F.gelu(F.linear(x.double(), w.double(), b.double())) + res.double()
    assert rel_err(outs[(1, 0)][0], ref) < 1e-5
    for tw in (0, 32, 64, 128, 192, 256):
        for i in range(3):
            assert torch.equal(outs[(0, tw)][i], outs[(1, tw)][i]), (tw, i)


@pytest.mark.parametrize("M,N,K", [(2050, 1024, 512), (1025, 768, 256), (1025, 96, 256), (130, 512, 64), (17, 256, 128), (257, 83, 192)])
def test_gemm_h3_narrow_last_tile_is_bit_identical(M, N, K):
    """The last token tile of a problem runs narrower MMAs (its remaining rows rounded up to 16 / 32 columns) instead of a full-width tile that is
    mostly padding; same k order, so every output bit must equal the uniform-tile kernel's -- fp32 / plane-pair / GELU / residual / grouped."""
    from siu3r_b200 import ops
    lib = ops._lib.load()
    x, w, b = rnd(M, K, seed=51), rnd(N, K, seed=52, scale=K ** -0.5), rnd(N, seed=53)
    res = rnd(M, N, seed=54)
    wt = ops.Weight(w, b, H3)
    xs = ops.split(x)
    h = M // 2
    outs = {}
    try:
        for mode in (0, 1):
            lib.siu3r_gemm_h3_set_remainder_tiles(mode)
            for tw in (0, 32, 64, 96, 128, 160, 256):
                lib.siu3r_gemm_h3_force(tw)
                a = ops.gemm(xs, wt, act=1, residual=res, precision=H3)
                s = ops.gemm(xs, wt, act=2, precision=H3, round_out=True).t.clone()
                g2 = ops.gemm_group2([xs[:h], xs[h:]], [wt, wt], precision=H3)
                outs[(mode, tw)] = (a, s, g2[0], g2[1])
    finally:
        lib.siu3r_gemm_h3_set_remainder_tiles(1)
        lib.siu3r_gemm_h3_force(0)
    ref = F.gelu(F.linear(x.double(), w.double(), b.double())) + res.double()
    assert rel_err(outs[(1, 0)][0], ref) < 1e-5
    for tw in (0, 32, 64, 96, 128, 160, 256):
        for i in range(4):
            assert torch.equal(outs[(0, tw)][i], outs[(1, tw)][i]), (tw, i)


@pytest.mark.parametrize("M,C,N", [(2050, 1024, 3072), (1025, 768, 768), (77, 256, 96)])
def test_gemm_h3_fused_layernorm(M, C, N):
    """LayerNorm fused across two GEMMs (siu3r_gemm_h3_ln): the producer (residual-adding projection) emits fp32 rows + plane pair + fixed-point row
    statistics, the consumer multiplies the RAW rows by gamma-folded weights and normalises in its epilogue.  Reference: the same chain in float64
    (croco/blocks.py:127-130: x = x + proj(a); y = fc(norm(x)))."""
    from siu3r_b200 import ops
    a = rnd(M, C, seed=31)
    x0 = rnd(M, C, seed=32) * 3.0 + 2.0 * rnd(M, 1, seed=33)          # residual stream with a per-row offset of ~2 sigma / 3
    wp, bp = rnd(C, C, seed=34, scale=C ** -0.5), rnd(C, seed=35)
    gam, bet = 1.0 + 0.3 * rnd(C, seed=36), 0.2 * rnd(C, seed=37)
    w2, b2 = rnd(N, C, seed=38, scale=C ** -0.5), rnd(N, seed=39)
    wproj, wfc = ops.Weight(wp, bp, H3), ops.Weight(w2, b2, H3)
    wfold = wfc.fold_ln(gam, bet)
    x_ref = x0.double() + F.linear(a.double(), wp.double(), bp.double())
    y_ref = F.gelu(F.linear(F.layer_norm(x_ref, (C,), gam.double(), bet.double(), 1e-6), w2.double(), b2.double()))
    # producer
    x_plain = ops.gemm(a, wproj, residual=x0, precision=H3)
    stats = torch.zeros(M, 2, device=DEV, dtype=torch.int64)
    xs, xf = ops.Split.empty(M, C, device=DEV), torch.empty(M, C, device=DEV)
    ops.gemm(a, wproj, residual=x0, out=xs, out_f32=xf, stats_out=stats, precision=H3)
    assert torch.equal(xf, x_plain)                                   # the extra outputs do not change the fp32 result
    assert torch.equal(xs.t, ops.split(xf).t)                         # the plane pair is the split of exactly those values
    s1, s2 = stats[:, 0].double() / 2 ** 32, stats[:, 1].double() / 2 ** 26
    assert float((s1 - xf.double().sum(1)).abs().max()) < 2e-6 * float(xf.abs().sum(1).max()) + 1e-5
    assert float(((s2 - (xf.double() ** 2).sum(1)).abs() / (xf.double() ** 2).sum(1)).max()) < 1e-6
    stats2 = torch.zeros_like(stats)
    ops.gemm(a, wproj, residual=x0, out=ops.Split.empty(M, C, device=DEV), out_f32=torch.empty(M, C, device=DEV), stats_out=stats2, precision=H3)
    assert torch.equal(stats, stats2)                                 # integer accumulation: bit-reproducible whatever the order of the atomics
    # consumer
    y = ops.gemm(xs, wfold, act=1, ln_stats=stats, precision=H3)
    assert rel_err(y, y_ref) < 1e-5, rel_err(y, y_ref)
    y_unfused = ops.gemm(ops.layernorm_h3([xf], [(gam, bet)], 1e-6)[0], wfc, act=1, precision=H3)
    assert rel_err(y, y_unfused.double()) < 1e-5
    # every tile width / the grouped launch
    lib = ops._lib.load()
    try:
        for tw in (32, 96, 128, 256):
            lib.siu3r_gemm_h3_force(tw)
            assert rel_err(ops.gemm(xs, wfold, act=1, ln_stats=stats, precision=H3), y_ref) < 1e-5, tw
            st = torch.zeros_like(stats)
            ops.gemm(a, wproj, residual=x0, out=ops.Split.empty(M, C, device=DEV), out_f32=torch.empty(M, C, device=DEV), stats_out=st, precision=H3)
            assert torch.equal(st, stats), tw
    finally:
        lib.siu3r_gemm_h3_force(0)
    h = M // 2
    ys = ops.gemm_group2([xs[:h], xs[h:]], [wfold, wfold], act=1, precision=H3, round_out=True, ln_stats=[stats[:h], stats[h:]])
    assert rel_err(torch.cat([ys[0].float(), ys[1].float()]), y_ref) < 1e-5


@pytest.mark.parametrize("mag", [1e-3, 1.0, 300.0])
def test_gemm_h3_fused_layernorm_row_magnitudes(mag):
    """The fixed-point row statistics (2^-32 resolution, int64 range) keep LayerNorm accurate for rows of magnitude 1e-3 as well as for rows with
    outliers of several hundred (the 'massive activations' of real ViT checkpoints)."""
    from siu3r_b200 import ops
    M, C, N = 515, 768, 256
    x0 = rnd(M, C, seed=41) * mag
    x0[:, 7] *= 40.0                                                  # one outlier channel
    zero_w = ops.Weight(torch.zeros(C, C, device=DEV), torch.zeros(C, device=DEV), H3)
    gam, bet = 1.0 + 0.3 * rnd(C, seed=42), 0.2 * rnd(C, seed=43)
    w2, b2 = rnd(N, C, seed=44, scale=C ** -0.5), rnd(N, seed=45)
    wfold = ops.Weight(w2, b2, H3).fold_ln(gam, bet)
    stats = torch.zeros(M, 2, device=DEV, dtype=torch.int64)
    xs, xf = ops.Split.empty(M, C, device=DEV), torch.empty(M, C, device=DEV)
    ops.gemm(torch.zeros(M, C, device=DEV), zero_w, residual=x0, out=xs, out_f32=xf, stats_out=stats, precision=H3)
    assert torch.equal(xf, x0)
    y = ops.gemm(xs, wfold, ln_stats=stats, precision=H3)
    ref = F.linear(F.layer_norm(x0.double(), (C,), gam.double(), bet.double(), 1e-6), w2.double(), b2.double())
    assert rel_err(y, ref) < 2e-5, (mag, rel_err(y, ref))


def test_model_fused_layernorm_matches_unfused(monkeypatch):
    """The ViT stages of the h3 forward with LayerNorm fused into the projections agree with the un-fused reference path (separate
    layernorm_kernel launches) far inside the parity tolerances."""
    from siu3r_b200 import model as model_mod, synth
    from siu3r_b200.model import ModelCfg, SIU3RModel
    S = 128
    sd = synth.make_state_dict()
    img, K = synth.pair_inputs(1, 2, S)
    caps = []
    for fuse in (True, False):
        monkeypatch.setattr(model_mod, "FUSE_LN", fuse)
        m = SIU3RModel(ModelCfg(image_size=(S, S)), precision="h3")
        m.load_state_dict(sd)
        m.cuda()
        assert m.fuse_ln == fuse
        m.capture = {}
        out = m(img.cuda(), K.cuda())
        caps.append((m.capture, out[0]))
    for name in ("enc_norm", "dec1_5", "dec2_11", "gs_raw"):
        a_, b_ = caps[0][0][name], caps[1][0][name]
        a_, b_ = (torch.cat([t.flatten() for t in a_]), torch.cat([t.flatten() for t in b_])) if isinstance(a_, (list, tuple)) else (a_, b_)
        assert rel_err(a_, b_.double()) < 2e-5, (name, rel_err(a_, b_.double()))
    assert float((caps[0][1].means - caps[1][1].means).abs().max()) < 1e-4


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("split_out", [False, True])
def test_gemm_h3_epilogue_strided(act, split_out):
    from siu3r_b200 import ops
    M, N, K = 777, 320, 512
    xbig = ops.split(rnd(M, K + 64, seed=4))
    x = xbig[:, :K]  # pitch K + 64
    xf = x.float()
    w, b = rnd(N, K, seed=5, scale=K ** -0.5), rnd(N, seed=6)
    res = rnd(M, N + 32, seed=7)[:, :N]
    wt = ops.Weight(w, b, H3)
    ref = 0.5 * F.linear(xf.double(), w.double()) + b.double()
    ref = F.gelu(ref) if act == 1 else (F.relu(ref) if act == 2 else ref)
    ref = ref + res.double()
    if split_out:
        big = ops.Split(torch.full((2, M, N + 16), 7.0, device=DEV, dtype=torch.float16))
        ops.gemm(x, wt, out=big[:, :N], act=act, residual=res, alpha=0.5, precision=H3)
        assert rel_err(big[:, :N].float(), ref) < 3e-6
        assert torch.all(big.t[:, :, N:] == 7.0)
    else:
        big = torch.full((M, N + 16), 7.0, device=DEV)
        ops.gemm(x, wt, out=big[:, :N], act=act, residual=res, alpha=0.5, precision=H3)
        assert rel_err(big[:, :N], ref) < 3e-6
        assert torch.all(big[:, N:] == 7.0)  # ldc padding untouched


def test_gemm_h3_inplace_residual_and_long_k():
    from siu3r_b200 import ops
    M, N, K = 1025, 1024, 4096
    x, w, b = rnd(M, K, seed=8), rnd(N, K, seed=9, scale=K ** -0.5), rnd(N, seed=10)
    r = rnd(M, N, seed=11)
    ref = F.linear(x.double(), w.double(), b.double()) + r.double()
    ops.gemm(x, ops.Weight(w, b, H3), out=r, residual=r, precision=H3)
    assert rel_err(r, ref) < 1e-5


def _positions(n_tok, Bn):
    g = int((n_tok - 1) ** 0.5)
    ys, xs = torch.meshgrid(torch.arange(g), torch.arange(g), indexing="ij")
    pos = torch.cat([torch.stack([ys.flatten(), xs.flatten()], -1), torch.tensor([[g, 0]])], 0)[None].repeat(Bn, 1, 1).contiguous().to(DEV)
    assert pos.shape[1] == n_tok
    return pos, g


@pytest.mark.parametrize("M0,M1,N,K,cols,col0", [(1025, 1025, 2304, 768, 1536, 1536), (257, 257, 2304, 768, 1536, 1536), (1025, 1025, 1536, 768, 768, 768),
                                                 (2050, 0, 3072, 1024, 2048, 2048), (1025, 3075, 768, 768, 768, 0)])
def test_gemm_h3_group_rope_vt(M0, M1, N, K, cols, col0):
    """One or two problems per launch with RoPE on the q / k columns and the V columns emitted as V^T (unscaled plane pairs), against
    fp32 GEMM -> rope2d (fp32 kernel, itself checked against the reference's curope) -> transpose."""
    from siu3r_b200 import ops
    Ms = [M0] + ([M1] if M1 else [])
    n_tok = 1025 if M0 % 1025 == 0 else 257
    pos, g = _positions(n_tok, max(Ms) // n_tok)
    tab = ops.rope2d_table(g + 1)
    xs_f = [rnd(m, K, seed=80 + i) for i, m in enumerate(Ms)]
    wts = [ops.Weight(rnd(N, K, seed=90 + i) / K ** 0.5, rnd(N, seed=95 + i), H3) for i in range(len(Ms))]
    parent = ops.Split.empty(sum(Ms), K, device=DEV)
    xs, r0 = [], 0
    for xf in xs_f:
        xs.append(ops.split(xf, parent[r0:r0 + xf.shape[0]]))
        r0 += xf.shape[0]
    outp = ops.Split.empty(sum(Ms), N, device=DEV, unscaled=True)
    outs, r0 = [], 0
    for m in Ms:
        outs.append(outp[r0:r0 + m])
        r0 += m
    vt = None
    if col0:
        strides = [(m + 7) // 8 * 8 for m in Ms]
        vbuf = ops.Split(torch.full((2, N - col0, sum(strides)), 9.0, device=DEV, dtype=torch.float16), unscaled=True)
        wins, c0 = [], 0
        for st in strides:
            wins.append(vbuf[:, c0:c0 + st])
            c0 += st
    state = {}
    if len(Ms) == 2:
        ops.gemm_group2(xs, wts, outs=outs, precision=H3, rope=(pos.view(-1, 2), tab, cols), vt=(wins, strides, col0, state) if col0 else None,
                        unscaled=True)
    else:
        ops.gemm(xs[0], wts[0], out=outs[0], precision=H3, rope=(pos.view(-1, 2), tab, cols), vt=(wins[0], col0, state) if col0 else None, unscaled=True)
    for i, m in enumerate(Ms):
        ref = F.linear(xs_f[i].double(), wts[i].w.double(), wts[i].bias.double()).float().contiguous()
        Bn = m // n_tok
        ops.rope2d_(ref, 0, pos, Bn, n_tok, cols // 64, 64, n_tok * N, N)
        ncol = col0 if col0 else N
        got = outs[i][:, :ncol].float()
        assert float((got - ref[:, :ncol]).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max())), i
        if col0:
            w_ = wins[i]
            gv = w_.float()        # [N - col0, stride]
            assert float((gv[:, :m] - ref[:, col0:].t()).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max())), i
            assert torch.all(w_.t[:, :, m:] == 0)   # pad columns zero-filled


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 256, 3), (1, 32, 64, 96, 256, 3), (1, 128, 128, 256, 256, 3), (1, 64, 64, 768, 256, 3),
                                   (1, 24, 32, 256, 128, 3), (3, 8, 16, 128, 83, 1), (1, 256, 256, 64, 64, 3), (2, 40, 48, 32, 256, 7)])
@pytest.mark.parametrize("split_out", [False, True])
def test_conv2d_h3(shape, split_out):
    from siu3r_b200 import ops
    n, h, w_, cin, cout, k = shape
    x = rnd(n, h, w_, cin, seed=12)
    w = rnd(cout, cin, k, k, seed=13, scale=(cin * k * k) ** -0.5)
    b = rnd(cout, seed=14)
    res = rnd(n, h, w_, cout, seed=15)
    wt = ops.Weight(w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous(), b, H3)
    ref = F.relu(F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=k // 2)).permute(0, 2, 3, 1) + res.double()
    y = ops.conv2d(x, wt, k, k, pad=k // 2, act=2, residual=res, precision=H3, round_out=split_out)
    if split_out:
        assert isinstance(y, ops.Split)
        y = y.view(-1, cout).float().view(n, h, w_, cout)
    assert rel_err(y, ref) < 1e-5, rel_err(y, ref)
    if cin >= 64:   # plane-pair input, every patch height
        lib = ops._lib.load()
        try:
            for tw in (32, 64, 128, 256):
                lib.siu3r_gemm_h3_force(tw)
                y2 = ops.conv2d(ops.split(x), wt, k, k, pad=k // 2, act=2, residual=res, precision=H3)
                assert rel_err(y2, ref) < 1e-5, tw
        finally:
            lib.siu3r_gemm_h3_force(0)


@pytest.mark.parametrize("cfg", [(1, 64, 64, 4, 64, 3, 2, 1), (2, 32, 32, 4, 256, 7, 1, 3), (1, 32, 32, 768, 768, 3, 2, 1), (1, 4, 4, 96, 256, 3, 1, 1),
                                 (2, 64, 64, 4, 1024, 16, 16, 0)])
def test_conv2d_h3_im2col_and_rowpacked(cfg):
    from siu3r_b200 import ops
    n, h, w_, cin, cout, k, stride, pad = cfg
    x = rnd(n, h, w_, cin, seed=16)
    w = rnd(cout, cin, k, k, seed=17, scale=(cin * k * k) ** -0.5)
    b = rnd(cout, seed=18)
    wt = ops.Weight(w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous(), b, H3)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    y = ops.conv2d(x, wt, k, k, stride=stride, pad=pad, precision=H3)
    assert rel_err(y, ref) < 1e-5, rel_err(y, ref)


@pytest.mark.parametrize("C_", [1024, 768, 256])
def test_layernorm_h3(C_):
    from siu3r_b200 import ops
    x0, x1 = rnd(1025, C_, seed=19) * 3 + 0.5, rnd(2050, C_, seed=20)
    wb = [(rnd(C_, seed=21), rnd(C_, seed=22)), (rnd(C_, seed=23), rnd(C_, seed=24))]
    outs = ops.layernorm_h3([x0, x1], wb, 1e-6)
    f32 = [torch.empty_like(x0), torch.empty_like(x1)]
    ops.layernorm_h3([x0, x1], wb, 1e-6, outs_f32=f32)
    for x, (w, b), o, f in zip((x0, x1), wb, outs, f32):
        ref = F.layer_norm(x.double(), (C_,), w.double(), b.double(), 1e-6)
        assert rel_err(o.float(), ref) < 2e-6
        assert rel_err(f, ref) < 2e-6
    one = ops.layernorm_h3([x0], wb[:1], 1e-6)[0]
    assert torch.equal(one.t, outs[0].t)


def test_eltwise_resize_im2col_h3():
    from siu3r_b200 import ops
    a, b = rnd(3, 20, 24, 64, seed=25), rnd(3, 20, 24, 64, seed=26)
    assert rel_err(ops.eltwise_h3(ops.ELT_RELU, a).view(-1, 64).float(), F.relu(a).view(-1, 64)) < 1e-6
    assert rel_err(ops.eltwise_h3(ops.ELT_ADD, a, b).view(-1, 64).float(), (a + b).view(-1, 64)) < 1e-6
    for align in (True, False):
        ref = ops.resize_bilinear(a, 40, 48, align)
        got = ops.resize_bilinear_h3(a, 40, 48, align)
        assert rel_err(got.view(-1, 64).float(), ref.view(-1, 64)) < 1e-6
    cols = ops.im2col_h3(a, 3, 3, 2, 1, 1, 9 * 64)
    ref = torch.empty(cols.shape[0], 9 * 64, device=DEV)
    ops._lib.check(ops._lib.load().siu3r_im2col_nhwc(a.data_ptr(), 3, 20, 24, 64, 3, 3, 2, 1, 1, ref.data_ptr(), 9 * 64, 0, ops._stream()), "im2col")
    assert rel_err(cols.float(), ref) < 1e-6


def _attn_ref(q, k, v, scale):
    s = torch.einsum("bhqd,bhkd->bhqk", q, k) * scale
    return torch.einsum("bhqk,bhkd->bhqd", s.softmax(-1), v)


def _unscaled(x2d):
    from siu3r_b200 import ops
    hi = x2d.clamp(-65504, 65504).half()
    lo = (x2d - hi.float()).half()
    return ops.Split(torch.stack([hi, lo]).contiguous(), unscaled=True)


@pytest.mark.parametrize("Nq,Nk,H", [(1025, 1025, 16), (257, 257, 12), (100, 3075, 12), (17, 17, 16), (128, 256, 2)])
@pytest.mark.parametrize("split_out", [False, True])
def test_flash_attn_h3(Nq, Nk, H, split_out):
    from siu3r_b200 import ops
    B, D = 2, 64
    q, k, v = rnd(B, Nq, H, D, seed=58), rnd(B, Nk, H, D, seed=59), rnd(B, Nk, H, D, seed=60)
    ref = _attn_ref(q.permute(0, 2, 1, 3).double(), k.permute(0, 2, 1, 3).double(), v.permute(0, 2, 1, 3).double(), D ** -0.5)
    ref = ref.permute(0, 2, 1, 3).reshape(B * Nq, H * D)
    vt = ops.transpose_v_h3(v, 0, Nk * H * D, H * D, B, Nk, H)
    errs = {}
    for swap in (0, 1):    # packing order of the fp16 P pairs inside a TMEM column: exactly one of the two is right
        ops._lib.load().siu3r_flash_h3_debug_swap(swap)
        out = ops.flash_attn_h3(_unscaled(q.view(B * Nq, H * D)), 0, _unscaled(k.view(B * Nk, H * D)), 0, vt, 0, B, H, Nq, Nk, D ** -0.5, split_out=split_out)
        errs[swap] = rel_err(out.float() if split_out else out, ref)
    ops._lib.load().siu3r_flash_h3_debug_swap(0)
    print("flash_h3 rel err (swap 0 / 1):", errs)
    assert errs[0] < (1e-5 if Nk <= 1025 else 3e-5), errs


def test_flash_attn_h3_inside_projection_output():
    """The engine's layout: q | k columns of an unscaled plane pair written by the fused qkv projection, V^T written by the same launch with
    images at an unaligned column pitch (1025 tokens per image) and a second window for the second problem of a grouped launch."""
    from siu3r_b200 import ops
    N, H, D, K = 1025, 12, 64, 768
    C = H * D
    pos, g = _positions(N, 1)
    tab = ops.rope2d_table(g + 1)
    xs_f = [rnd(N, K, seed=101), rnd(N, K, seed=102)]
    wts = [ops.Weight(rnd(3 * C, K, seed=103 + i) / K ** 0.5, rnd(3 * C, seed=105 + i), H3) for i in range(2)]
    qkv = ops.Split.empty(2 * N, 3 * C, device=DEV, unscaled=True)
    stride = (N + 7) // 8 * 8
    vbuf = ops.Split.empty(C, 2 * stride, device=DEV, unscaled=True)
    ops.gemm_group2(xs_f, wts, outs=[qkv[:N], qkv[N:]], precision=H3, rope=(pos.view(-1, 2), tab, 2 * C),
                    vt=([vbuf[:, :stride], vbuf[:, stride:]], [stride, stride], 2 * C, {}), unscaled=True)
    out = ops.flash_attn_h3(qkv, 0, qkv, C, vbuf, stride, 2, H, N, N, 0.125, split_out=False)
    for i in range(2):
        ref = F.linear(xs_f[i].double(), wts[i].w.double(), wts[i].bias.double()).float().contiguous()
        ops.rope2d_(ref, 0, pos, 1, N, 2 * C // 64, 64, N * 3 * C, 3 * C)
        q, k, v = [ref[:, j * C:(j + 1) * C].view(1, N, H, D).permute(0, 2, 1, 3).double() for j in range(3)]
        want = _attn_ref(q, k, v, 0.125).permute(0, 2, 1, 3).reshape(N, C)
        assert rel_err(out[i * N:(i + 1) * N], want) < 3e-5, i
    # images at an unaligned uniform pitch (one problem, two images): vt_batch_cols = N, second window offset unused
    x2 = rnd(2 * N, K, seed=110)
    qkv2 = ops.Split.empty(2 * N, 3 * C, device=DEV, unscaled=True)
    v2 = ops.Split.empty(C, (2 * N + 7) // 8 * 8, device=DEV, unscaled=True)
    pos2, _ = _positions(N, 2)
    ops.gemm(x2, wts[0], out=qkv2, precision=H3, rope=(pos2.view(-1, 2), tab, 2 * C), vt=(v2, 2 * C, {}), unscaled=True)
    out2 = ops.flash_attn_h3(qkv2, 0, qkv2, C, v2, N, 2, H, N, N, 0.125, split_out=False)
    ref = F.linear(x2.double(), wts[0].w.double(), wts[0].bias.double()).float().contiguous()
    ops.rope2d_(ref, 0, pos2, 2, N, 2 * C // 64, 64, N * 3 * C, 3 * C)
    for i in range(2):
        q, k, v = [ref[i * N:(i + 1) * N, j * C:(j + 1) * C].view(1, N, H, D).permute(0, 2, 1, 3).double() for j in range(3)]
        want = _attn_ref(q, k, v, 0.125).permute(0, 2, 1, 3).reshape(N, C)
        assert rel_err(out2[i * N:(i + 1) * N], want) < 3e-5, i
