// HBM-bound supporting kernels of the per-pair hot path (fp32, token-major / NHWC layouts).
// Each entry point cites the reference op it replaces (paths relative to /root/reference).
#include "common.cuh"
#include "h3.cuh"

namespace {

// optional fp16 plane-pair destination (h3 mode, see h3.cuh) next to / instead of the fp32 one: p == nullptr -> unused
struct H3Out { __half* p; int64_t ld; int64_t plane; };
__device__ __forceinline__ void h3_store4(const H3Out& o, int64_t off, float4 v) {
    uint2 hi, lo;
    h3_split2(v.x, v.y, hi.x, lo.x);
    h3_split2(v.z, v.w, hi.y, lo.y);
    *reinterpret_cast<uint2*>(o.p + off) = hi;
    *reinterpret_cast<uint2*>(o.p + o.plane + off) = lo;
}

__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim: one warp per row, row cached in registers, two-pass (mean, then centred variance)
// like ATen.  Replaces nn.LayerNorm in croco/blocks.py:119,123,176,180-184 (eps 1e-6, croco/croco.py:35),
// vit_adapter/vit_adapter.py:74-93, mask2former/video_seg_decoder.py:945-952,1738-1744 (eps 1e-5).
// ------------------------------------------------------------------------------------------------------------
// Second segment of a grouped launch (siu3r_layernorm_group2): rows >= rows0 read x1 / w1 / b1 and write y1 (row index rebased).
struct LnSeg2 { const float* x1; const float* w1; const float* b1; float* y1; int rows0; __half* yh1; };

template <int VEC_PER_LANE>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ y, int64_t ldy, int rows, int C,
                                                        float eps, const float* __restrict__ add, int64_t ldadd, int round_out, LnSeg2 g, H3Out yh) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_wait();                 // launched with programmatic stream serialization (common.cuh): the producer of x has completed after this
    pdl_launch_dependents();
    if (row >= rows) return;
    if (g.x1 && row >= g.rows0) { row -= g.rows0; x = g.x1; w = g.w1; b = g.b1; y = g.y1; yh.p = g.yh1; }
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * ldx);
    const int nvec = C >> 2;
    float4 v[VEC_PER_LANE];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_PER_LANE; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            v[i] = xr[idx];
            s += v[i].x + v[i].y + v[i].z + v[i].w;
        }
    }
    s = warp_sum(s);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_PER_LANE; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += a * a + bb * bb + c * c + d * d;
        }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / (float)C + eps);
    float4* yr = reinterpret_cast<float4*>(y + (int64_t)row * ldy);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    const float4* a4 = add ? reinterpret_cast<const float4*>(add + (int64_t)row * ldadd) : nullptr;
#pragma unroll
    for (int i = 0; i < VEC_PER_LANE; ++i) {
        const int idx = lane + i * 32;
        if (idx < nvec) {
            const float4 ww = w4[idx], bb = b4[idx];
            float4 o;
            o.x = (v[i].x - mean) * rstd * ww.x + bb.x;
            o.y = (v[i].y - mean) * rstd * ww.y + bb.y;
            o.z = (v[i].z - mean) * rstd * ww.z + bb.z;
            o.w = (v[i].w - mean) * rstd * ww.w + bb.w;
            if (a4) { const float4 aa = a4[idx]; o.x += aa.x; o.y += aa.y; o.z += aa.z; o.w += aa.w; }
            if (round_out) { o.x = rn_tf32(o.x); o.y = rn_tf32(o.y); o.z = rn_tf32(o.z); o.w = rn_tf32(o.w); }
            if (y) yr[idx] = o;
            if (yh.p) h3_store4(yh, (int64_t)row * yh.ld + idx * 4, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// 2-D rotary embedding, in place.  Replaces curope.rope_2d (croco/curope/curope.cpp:49-65, kernels.cu:17-82) as
// called from croco/blocks.py:101-103,158-160.  tokens[b][n][h][d] at b*batch_stride + n*token_stride + h*D + d.
// One thread per (token, head, quarter index d<D/4, x/y half): rotates the pair (d, d+D/4) of its half.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rope2d_kernel(float* __restrict__ tokens, const int64_t* __restrict__ pos, int B, int N, int H, int D,
                                                     int64_t batch_stride, int64_t token_stride, float base, float fwd, int nparts,
                                                     int64_t part_stride, int round_out) {
    const int Q = D >> 2;
    const int64_t total = (int64_t)B * N * nparts * H * 2 * Q;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int d = (int)(idx % Q);
    int64_t t = idx / Q;
    const int xh = (int)(t % 2); t /= 2;
    const int h = (int)(t % H); t /= H;
    const int part = (int)(t % nparts); t /= nparts;
    const int n = (int)(t % N);
    const int b = (int)(t / N);
    tokens += (int64_t)part * part_stride;
    const float inv_freq = fwd / powf(base, (float)d / (float)Q);
    const float f = (float)pos[((int64_t)b * N + n) * 2 + xh] * inv_freq;
    float s, c;
    sincosf(f, &s, &c);
    float* p = tokens + (int64_t)b * batch_stride + (int64_t)n * token_stride + (int64_t)h * D + xh * (D >> 1) + d;
    const float u = p[0], v = p[Q];
    const float o0 = u * c - v * s, o1 = v * c + u * s;
    p[0] = round_out ? rn_tf32(o0) : o0;   // q / k only feed the TF32 attention: round to nearest here (the tensor core truncates)
    p[Q] = round_out ? rn_tf32(o1) : o1;
}

// (cos, sin)(pos * fwd / base^(d/Q)) for pos < maxpos, d < Q: the same expressions as rope2d_kernel, evaluated once
__global__ void rope2d_table_kernel(float2* __restrict__ tab, int maxpos, int Q, float base, float fwd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= maxpos * Q) return;
    const int d = idx % Q, pos = idx / Q;
    const float inv_freq = fwd / powf(base, (float)d / (float)Q);
    const float f = (float)pos * inv_freq;
    float s, c;
    sincosf(f, &s, &c);
    tab[idx] = make_float2(c, s);
}

// x = hi + lo with hi = RN_tf32(x), lo = RN_tf32(x - hi): operands of the 3xTF32 tensor-core path
__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = x[i];
    float4 h, l;
    const float in[4] = {v.x, v.y, v.z, v.w};
    float ho[4], lo_[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        uint32_t hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[e]));
        ho[e] = __uint_as_float(hb);
        const float r = in[e] - ho[e];
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(r));
        lo_[e] = __uint_as_float(lb);
    }
    h = make_float4(ho[0], ho[1], ho[2], ho[3]);
    l = make_float4(lo_[0], lo_[1], lo_[2], lo_[3]);
    hi[i] = h;
    lo[i] = l;
}

// ------------------------------------------------------------------------------------------------------------
// Elementwise family (vectorised float4, n % 4 == 0 enforced by the host wrapper; tails handled scalar)
// ------------------------------------------------------------------------------------------------------------
enum EltOp { ELT_RELU = 0, ELT_ADD = 1, ELT_ADD_RELU = 2, ELT_GELU = 3, ELT_COPY = 4, ELT_SIGMOID = 5, ELT_CLAMP01 = 6, ELT_RELU_RN = 7, ELT_ROUND = 8, ELT_ADD_RN = 9 };

__device__ __forceinline__ float elt_apply(int op, float a, float b) {
    switch (op) {
        case ELT_RELU: return fmaxf(a, 0.f);
        case ELT_ADD: return a + b;
        case ELT_ADD_RELU: return fmaxf(a + b, 0.f);
        case ELT_GELU: return 0.5f * a * (1.0f + erff(a * 0.70710678118654752440f));
        case ELT_SIGMOID: return 1.0f / (1.0f + expf(-a));
        case ELT_CLAMP01: return fminf(fmaxf(a, 0.f), 1.f);
        case ELT_RELU_RN: return rn_tf32(fmaxf(a, 0.f));
        case ELT_ROUND: return rn_tf32(a);
        case ELT_ADD_RN: return rn_tf32(a + b);
        default: return a;
    }
}

__global__ void __launch_bounds__(256) eltwise_kernel(int op, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n,
                                                      H3Out oh) {
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 4 <= n) {
        const float4 va = *reinterpret_cast<const float4*>(a + i4);
        float4 vb = make_float4(0, 0, 0, 0);
        if (b) vb = *reinterpret_cast<const float4*>(b + i4);
        float4 o;
        o.x = elt_apply(op, va.x, vb.x); o.y = elt_apply(op, va.y, vb.y);
        o.z = elt_apply(op, va.z, vb.z); o.w = elt_apply(op, va.w, vb.w);
        if (out) *reinterpret_cast<float4*>(out + i4) = o;
        if (oh.p) h3_store4(oh, i4, o);
    } else {
        for (int64_t i = i4; i < n; ++i) {
            const float v = elt_apply(op, a[i], b ? b[i] : 0.f);
            if (out) out[i] = v;
            if (oh.p) h3_split(v, oh.p[i], oh.p[oh.plane + i]);
        }
    }
}

__global__ void __launch_bounds__(256) scale_kernel(const float* a, float alpha, float* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * alpha;
}

// rows x C strided copy / broadcast-add:  y[r, :] = x[r, :] (+ vec[:]) (+ y2[r, :])   with independent leading dims
__global__ void __launch_bounds__(256) rows_affine_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, const float* __restrict__ add, int64_t ldadd,
                                                          float* __restrict__ y, int64_t ldy, int64_t rows, int C, int relu) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = C >> 2;
    if (idx >= rows * c4) return;
    const int64_t r = idx / c4;
    const int c = (int)(idx % c4) * 4;
    float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    if (scale) { const float4 s = *reinterpret_cast<const float4*>(scale + c); v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w; }
    if (shift) { const float4 s = *reinterpret_cast<const float4*>(shift + c); v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w; }
    if (add) { const float4 s = *reinterpret_cast<const float4*>(add + r * ldadd + c); v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w; }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4*>(y + r * ldy + c) = v;
}

// ------------------------------------------------------------------------------------------------------------
// Bilinear resize, NHWC, align_corners True/False (ATen upsample_bilinear2d semantics), optional accumulate.
// Replaces F.interpolate(mode="bilinear") at heads/dpt_block.py:229-235,279-284, heads/dpt_gs_head.py:113,
// vit_adapter/vit_adapter.py:429-433, mask2former/video_seg_decoder.py:1461-1466,2173-2178.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void src_index(int o, int in_size, int out_size, bool align, float scale, int& i0, int& i1, float& l1) {
    float src;
    if (align) src = scale * (float)o;
    else {
        src = scale * ((float)o + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
    }
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ x, int N, int H, int W, int C, int64_t ldx,
                                                              float* __restrict__ y, int OH, int OW, int64_t ldy, int align, float sh,
                                                              float sw, int accumulate, H3Out yh) {
    const int c4 = C >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)N * OH * OW * c4;
    if (idx >= total) return;
    const int c = (int)(idx % c4) * 4;
    int64_t t = idx / c4;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH);
    const int n = (int)(t / OH);
    int h0, h1, w0, w1; float lh, lw;
    src_index(oh, H, OH, align != 0, sh, h0, h1, lh);
    src_index(ow, W, OW, align != 0, sw, w0, w1, lw);
    const float* base = x + (int64_t)n * H * W * ldx + c;
    const float4 v00 = *reinterpret_cast<const float4*>(base + ((int64_t)h0 * W + w0) * ldx);
    const float4 v01 = *reinterpret_cast<const float4*>(base + ((int64_t)h0 * W + w1) * ldx);
    const float4 v10 = *reinterpret_cast<const float4*>(base + ((int64_t)h1 * W + w0) * ldx);
    const float4 v11 = *reinterpret_cast<const float4*>(base + ((int64_t)h1 * W + w1) * ldx);
    const float h0l = 1.f - lh, w0l = 1.f - lw;
    float4 o;
    o.x = h0l * (w0l * v00.x + lw * v01.x) + lh * (w0l * v10.x + lw * v11.x);
    o.y = h0l * (w0l * v00.y + lw * v01.y) + lh * (w0l * v10.y + lw * v11.y);
    o.z = h0l * (w0l * v00.z + lw * v01.z) + lh * (w0l * v10.z + lw * v11.z);
    o.w = h0l * (w0l * v00.w + lw * v01.w) + lh * (w0l * v10.w + lw * v11.w);
    if (yh.p) { h3_store4(yh, (((int64_t)n * OH + oh) * OW + ow) * yh.ld + c, o); return; }
    float4* yp = reinterpret_cast<float4*>(y + (((int64_t)n * OH + oh) * OW + ow) * ldy + c);
    if (accumulate & 1) { const float4 p = *yp; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
    if (accumulate & 2) { o.x = rn_tf32(o.x); o.y = rn_tf32(o.y); o.z = rn_tf32(o.z); o.w = rn_tf32(o.w); }
    *yp = o;
}

// ConvTranspose2d(kernel = stride = s) scatter: g[(n,h,w), (dy,dx,co)] -> y[n, h*s+dy, w*s+dx, co] (+ add)
// (heads/dpt_block.py:422-451, vit_adapter/vit_adapter.py:356,425: the GEMM part runs on the tensor cores)
__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const float* __restrict__ g, int N, int H, int W, int C, int s,
                                                            const float* __restrict__ add, float* __restrict__ y) {
    const int c4 = C >> 2;
    const int OH = H * s, OW = W * s;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)N * OH * OW * c4;
    if (idx >= total) return;
    const int c = (int)(idx % c4) * 4;
    int64_t t = idx / c4;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH);
    const int n = (int)(t / OH);
    const int h = oh / s, dy = oh % s, w = ow / s, dx = ow % s;
    float4 v = *reinterpret_cast<const float4*>(g + (((int64_t)n * H + h) * W + w) * ((int64_t)s * s * C) + (int64_t)(dy * s + dx) * C + c);
    const int64_t o = (((int64_t)n * OH + oh) * OW + ow) * C + c;
    if (add) { const float4 a = *reinterpret_cast<const float4*>(add + o); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
    *reinterpret_cast<float4*>(y + o) = v;
}

// im2col for NHWC input: out[(n,oh,ow), (kh,kw,ci)] (row stride ldo >= KH*KW*C, pad columns zeroed by caller's memset)
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                                                     int pad_w, int OH, int OW, float* __restrict__ out, int64_t ldo, int round_out, H3Out oh) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int K = KH * KW * C;
    const int64_t total = (int64_t)N * OH * OW * ldo;
    if (idx >= total) return;
    const int k = (int)(idx % ldo);
    int64_t t = idx / ldo;
    float v = 0.f;
    if (k < K) {
        const int ow = (int)(t % OW);
        int64_t t2 = t / OW;
        const int oh = (int)(t2 % OH);
        const int n = (int)(t2 / OH);
        const int ci = k % C;
        const int tap = k / C;
        const int kh = tap / KW, kw = tap % KW;
        const int ih = oh * stride + kh - pad, iw = ow * stride + kw - pad_w;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[(((int64_t)n * H + ih) * W + iw) * C + ci];
    }
    if (oh.p) { h3_split(v, oh.p[idx], oh.p[oh.plane + idx]); return; }
    out[idx] = round_out ? rn_tf32(v) : v;
}

// NCHW <-> NHWC (images come in as [B,V,3,H,W], inference.py:117-118; outputs that the reference returns as NCHW)
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int C, int HW, int64_t ldy) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * C * HW) return;
    const int c = (int)(idx % C);
    int64_t t = idx / C;
    const int p = (int)(t % HW);
    const int n = (int)(t / HW);
    y[((int64_t)n * HW + p) * ldy + c] = x[((int64_t)n * C + c) * HW + p];
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int N, int C, int HW) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * C * HW) return;
    const int p = (int)(idx % HW);
    int64_t t = idx / HW;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    y[idx] = x[((int64_t)n * HW + p) * ldx + c];
}

// MaxPool2d(3, stride 2, pad 1), NHWC (vit_adapter/vit_adapter.py:220)
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const float* __restrict__ x, int N, int H, int W, int C, float* __restrict__ y, int OH, int OW) {
    const int c4 = C >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * OH * OW * c4) return;
    const int c = (int)(idx % c4) * 4;
    int64_t t = idx / c4;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH);
    const int n = (int)(t / OH);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh * 2 + kh - 1;
        if (ih < 0 || ih >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
            const int iw = ow * 2 + kw - 1;
            if (iw < 0 || iw >= W) continue;
            const float4 v = *reinterpret_cast<const float4*>(x + (((int64_t)n * H + ih) * W + iw) * C + c);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    *reinterpret_cast<float4*>(y + (((int64_t)n * OH + oh) * OW + ow) * C + c) = m;
}

// Depthwise 3x3 conv (stride 1, pad 1) + bias + optional exact GELU, NHWC with token leading dim
// (vit_adapter/vit_adapter.py:16-31 DWConv applied to each of the three token ranges, then act at :55)
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const float* __restrict__ x, int64_t ldx, int N, int H, int W, int C,
                                                        const float* __restrict__ w /*[3][3][C]*/, const float* __restrict__ b,
                                                        float* __restrict__ y, int64_t ldy, int64_t batch_stride_x, int64_t batch_stride_y, int gelu) {
    const int c4 = C >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * H * W * c4) return;
    const int c = (int)(idx % c4) * 4;
    int64_t t = idx / c4;
    const int ww = (int)(t % W); t /= W;
    const int hh = (int)(t % H);
    const int n = (int)(t / H);
    float4 acc = *reinterpret_cast<const float4*>(b + c);
    const float* xb = x + (int64_t)n * batch_stride_x;
    for (int kh = 0; kh < 3; ++kh) {
        const int ih = hh + kh - 1;
        if (ih < 0 || ih >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
            const int iw = ww + kw - 1;
            if (iw < 0 || iw >= W) continue;
            const float4 v = *reinterpret_cast<const float4*>(xb + ((int64_t)ih * W + iw) * ldx + c);
            const float4 k = *reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C + c);
            acc.x += v.x * k.x; acc.y += v.y * k.y; acc.z += v.z * k.z; acc.w += v.w * k.w;
        }
    }
    if (gelu) {
        acc.x = elt_apply(ELT_GELU, acc.x, 0); acc.y = elt_apply(ELT_GELU, acc.y, 0);
        acc.z = elt_apply(ELT_GELU, acc.z, 0); acc.w = elt_apply(ELT_GELU, acc.w, 0);
    }
    *reinterpret_cast<float4*>(y + (int64_t)n * batch_stride_y + ((int64_t)hh * W + ww) * ldy + c) = acc;
}

// GroupNorm over NHWC [N, HW, C] (mask2former/video_seg_decoder.py:2004,2036,2048: GroupNorm(32, 256), eps 1e-5, optional ReLU).
// Pass 1: per-(n, group) sum and sum of squares, accumulated in fp64 (block partials -> atomicAdd(double)), grid over pixels;
// pass 2: elementwise normalise.  (The former one-CTA-per-group kernel took 0.27 ms per call at 128x128.)
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const float* __restrict__ x, int HW, int C, int groups, double* __restrict__ stats) {
    // grid: (pixel chunks, N); each thread walks pixels of one channel-quad; C % 4 == 0 and (C/groups) % 4 == 0
    const int n = blockIdx.y;
    const int c4 = C >> 2;
    const int cq = threadIdx.x % c4;                 // channel quad handled by this thread (blockDim.x multiple of c4)
    const int prow = threadIdx.x / c4, pstep = blockDim.x / c4;
    const int cpg = C / groups;
    const int g = (cq * 4) / cpg;
    const int pix_per_block = (HW + gridDim.x - 1) / gridDim.x;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
    double s = 0.0, q = 0.0;
    const float* xb = x + (int64_t)n * HW * C + cq * 4;
    for (int p = p0 + prow; p < p1; p += pstep) {
        const float4 v = *reinterpret_cast<const float4*>(xb + (int64_t)p * C);
        s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
        q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    __shared__ double sh_s[32], sh_q[32];
    if (threadIdx.x < 32) { sh_s[threadIdx.x] = 0.0; sh_q[threadIdx.x] = 0.0; }
    __syncthreads();
    atomicAdd(&sh_s[g], s);
    atomicAdd(&sh_q[g], q);
    __syncthreads();
    if (threadIdx.x < groups) {
        atomicAdd(&stats[((int64_t)n * groups + threadIdx.x) * 2], sh_s[threadIdx.x]);
        atomicAdd(&stats[((int64_t)n * groups + threadIdx.x) * 2 + 1], sh_q[threadIdx.x]);
    }
}

__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float* __restrict__ x, int HW, int C, int groups, const double* __restrict__ stats,
                                                              const float* __restrict__ w, const float* __restrict__ b, float eps, int relu,
                                                              float* __restrict__ y, int64_t total4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c4 = C >> 2;
    const int c = (int)(i % c4) * 4;
    const int64_t pix = i / c4;
    const int n = (int)(pix / HW);
    const int cpg = C / groups;
    const int g = c / cpg;
    const double cnt = (double)HW * cpg;
    const double mean = stats[((int64_t)n * groups + g) * 2] / cnt;
    const double var = stats[((int64_t)n * groups + g) * 2 + 1] / cnt - mean * mean;
    const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)eps));
    float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    const float4 ww = *reinterpret_cast<const float4*>(w + c), bb = *reinterpret_cast<const float4*>(b + c);
    v.x = (v.x - mu) * rstd * ww.x + bb.x; v.y = (v.y - mu) * rstd * ww.y + bb.y;
    v.z = (v.z - mu) * rstd * ww.z + bb.z; v.w = (v.w - mu) * rstd * ww.w + bb.w;
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4*>(y + i * 4) = v;
}

inline unsigned grid_for(int64_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

extern "C" {

static int layernorm_impl(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy, int rows, int C,
                          float eps, const float* add, int64_t ldadd, int round_out, LnSeg2 g, void* stream_, H3Out yh = H3Out{nullptr, 0, 0}) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && w && b && (y || yh.p) && rows > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && (!y || ldy % 4 == 0));
    if (yh.p) SIU3R_REQUIRE(((uintptr_t)yh.p & 7) == 0 && yh.ld % 4 == 0 && yh.plane % 4 == 0);
    SIU3R_REQUIRE(C <= 4096);
    const int nvec = C / 4;
    const int vpl = ceil_div(nvec, 32);
    const int wpb = 8;
    dim3 grid(ceil_div(rows, wpb));
    if (vpl <= 2) siu3r_launch_pdl(layernorm_kernel<2>, grid, dim3(wpb * 32), 0, stream, x, ldx, w, b, y, ldy, rows, C, eps, add, ldadd, round_out, g, yh);
    else if (vpl <= 6) siu3r_launch_pdl(layernorm_kernel<6>, grid, dim3(wpb * 32), 0, stream, x, ldx, w, b, y, ldy, rows, C, eps, add, ldadd, round_out, g, yh);
    else if (vpl <= 8) siu3r_launch_pdl(layernorm_kernel<8>, grid, dim3(wpb * 32), 0, stream, x, ldx, w, b, y, ldy, rows, C, eps, add, ldadd, round_out, g, yh);
    else siu3r_launch_pdl(layernorm_kernel<32>, grid, dim3(wpb * 32), 0, stream, x, ldx, w, b, y, ldy, rows, C, eps, add, ldadd, round_out, g, yh);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}
int siu3r_layernorm(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy, int rows, int C, float eps,
                    const float* add, int64_t ldadd, int round_out, void* stream) {
    return layernorm_impl(x, ldx, w, b, y, ldy, rows, C, eps, add, ldadd, round_out, LnSeg2{nullptr, nullptr, nullptr, nullptr, 0, nullptr}, stream);
}

// Two LayerNorms with different affine parameters (and possibly different source / destination buffers) in one launch: the norm1 / norm2 /
// norm3 / norm_y pairs of the two decoder streams (croco/blocks.py:186-190 under dec_blocks and dec_blocks2).  Same C, pitch and eps.
int siu3r_layernorm_group2(const float* x0, const float* x1, int64_t ldx, const float* w0, const float* b0, const float* w1, const float* b1,
                           float* y0, float* y1, int64_t ldy, int rows0, int rows1, int C, float eps, int round_out, void* stream) {
    SIU3R_REQUIRE(x1 && w1 && b1 && y1 && rows0 > 0 && rows1 > 0);
    SIU3R_REQUIRE(((uintptr_t)x1 & 15) == 0 && ((uintptr_t)y1 & 15) == 0 && ((uintptr_t)w1 & 15) == 0 && ((uintptr_t)b1 & 15) == 0);
    return layernorm_impl(x0, ldx, w0, b0, y0, ldy, rows0 + rows1, C, eps, nullptr, 0, round_out, LnSeg2{x1, w1, b1, y1, rows0, nullptr}, stream);
}

// LayerNorm whose result goes to an fp16 plane pair yh (h3 mode: the row only feeds tensor-core operands) and / or to fp32 y (either may be
// null, not both).  x1 != null: second segment of a grouped launch as in siu3r_layernorm_group2 (rows >= rows0 use x1 / w1 / b1 / y1 / yh1).
int siu3r_layernorm_h3(const float* x0, const float* x1, int64_t ldx, const float* w0, const float* b0, const float* w1, const float* b1, float* y0,
                       float* y1, int64_t ldy, void* yh0, void* yh1, int64_t ldh, int64_t plane, int rows0, int rows1, int C, float eps,
                       void* stream) {
    SIU3R_REQUIRE(rows0 > 0 && rows1 >= 0 && (rows1 == 0 || (x1 && w1 && b1 && (y1 || yh1))));
    return layernorm_impl(x0, ldx, w0, b0, y0, ldy, rows0 + rows1, C, eps, nullptr, 0, 0,
                          LnSeg2{rows1 > 0 ? x1 : nullptr, w1, b1, y1, rows0, (__half*)yh1}, stream, H3Out{(__half*)yh0, ldh, plane});
}


int siu3r_rope2d(float* tokens, const int64_t* positions, int B, int N, int H, int D, int64_t batch_stride, int64_t token_stride,
                 float base, float fwd, int nparts, int64_t part_stride, int round_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(tokens && positions && B > 0 && N > 0 && H > 0);
    SIU3R_REQUIRE(D % 4 == 0);  // "token dim must be multiple of 4" (kernels.cu:94)
    SIU3R_REQUIRE(nparts >= 1);
    const int64_t total = (int64_t)B * N * nparts * H * 2 * (D / 4);
    rope2d_kernel<<<grid_for(total), 256, 0, stream>>>(tokens, positions, B, N, H, D, batch_stride, token_stride, base, fwd, nparts, part_stride, round_out);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_rope2d_table(float* tab, int maxpos, int D, float base, float fwd, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(tab && maxpos > 0 && D > 0 && D % 4 == 0);
    const int Q = D / 4;
    rope2d_table_kernel<<<(unsigned)ceil_div(maxpos * Q, 256), 256, 0, stream>>>((float2*)tab, maxpos, Q, base, fwd);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && hi && lo && n > 0 && n % 4 == 0);
    split_tf32_kernel<<<grid_for(n / 4), 256, 0, stream>>>((const float4*)x, (float4*)hi, (float4*)lo, n / 4);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// op: 0 relu(a), 1 a+b, 2 relu(a+b), 3 gelu(a), 4 copy, 5 sigmoid(a), 6 clamp(a, 0, 1), 7 RN_tf32(relu(a)), 8 RN_tf32(a), 9 RN_tf32(a+b)
int siu3r_eltwise(int op, const float* a, const float* b, float* out, int64_t n, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(a && out && n > 0 && op >= 0 && op <= 9);
    SIU3R_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)out & 15) == 0 && (!b || ((uintptr_t)b & 15) == 0));
    eltwise_kernel<<<grid_for(ceil_div_i64(n, 4)), 256, 0, stream>>>(op, a, b, out, n, H3Out{nullptr, 0, 0});
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// siu3r_eltwise whose result is stored as an fp16 plane pair (contiguous, lo plane `plane` elements after the hi plane); ops 0..6
int siu3r_eltwise_h3(int op, const float* a, const float* b, void* outh, int64_t plane, int64_t n, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(a && outh && n > 0 && op >= 0 && op <= 6 && plane >= n && plane % 4 == 0);
    SIU3R_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)outh & 7) == 0 && (!b || ((uintptr_t)b & 15) == 0));
    eltwise_kernel<<<grid_for(ceil_div_i64(n, 4)), 256, 0, stream>>>(op, a, b, nullptr, n, H3Out{(__half*)outh, 0, plane});
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// out = a * alpha (in place allowed): the x10 / x100 scene rescale of gaussian_renderer.py:43-46
int siu3r_scale(const float* a, float alpha, float* out, int64_t n, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(a && out && n > 0);
    scale_kernel<<<grid_for(n), 256, 0, stream>>>(a, alpha, out, n);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// y[r, 0:C] = relu?( x[r, 0:C] * scale[0:C] + shift[0:C] + add[r, 0:C] )   (scale / shift / add optional)
// Covers eval-mode BatchNorm (vit_adapter.py:357-360,437-440), level-embed adds (:387-391), strided copies.
int siu3r_rows_affine(const float* x, int64_t ldx, const float* scale, const float* shift, const float* add, int64_t ldadd, float* y,
                      int64_t ldy, int64_t rows, int C, int relu, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && y && rows > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && (!add || ldadd % 4 == 0));
    rows_affine_kernel<<<grid_for(rows * (C / 4)), 256, 0, stream>>>(x, ldx, scale, shift, add, ldadd, y, ldy, rows, C, relu);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

static int resize_impl(const float* x, int N, int H, int W, int C, int64_t ldx, float* y, int OH, int OW, int64_t ldy, int align_corners,
                       int accumulate, void* stream_, H3Out yh) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && (y || yh.p) && N > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0);
    float sh, sw;
    if (align_corners) {
        sh = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
        sw = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    } else {
        // F.interpolate(scale_factor=s) without recompute_scale_factor uses 1/s; with size= it uses in/out: identical for
        // the exact ratios on this path
        sh = (float)H / (float)OH;
        sw = (float)W / (float)OW;
    }
    resize_bilinear_kernel<<<grid_for((int64_t)N * OH * OW * (C / 4)), 256, 0, stream>>>(x, N, H, W, C, ldx, y, OH, OW, ldy, align_corners, sh,
                                                                                         sw, accumulate, yh);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}
int siu3r_resize_bilinear_nhwc(const float* x, int N, int H, int W, int C, int64_t ldx, float* y, int OH, int OW, int64_t ldy,
                               int align_corners, int accumulate, void* stream) {
    SIU3R_REQUIRE(y != nullptr);
    return resize_impl(x, N, H, W, C, ldx, y, OH, OW, ldy, align_corners, accumulate, stream, H3Out{nullptr, 0, 0});
}
// ... with the result stored as an fp16 plane pair [N, OH, OW, ldh] (h3 mode: the resized map only feeds a convolution)
int siu3r_resize_bilinear_nhwc_h3(const float* x, int N, int H, int W, int C, int64_t ldx, void* yh, int OH, int OW, int64_t ldh, int64_t plane,
                                  int align_corners, void* stream) {
    SIU3R_REQUIRE(yh && ((uintptr_t)yh & 7) == 0 && ldh % 4 == 0 && plane % 4 == 0);
    return resize_impl(x, N, H, W, C, ldx, nullptr, OH, OW, 4, align_corners, 0, stream, H3Out{(__half*)yh, ldh, plane});
}

int siu3r_pixel_shuffle_nhwc(const float* g, int N, int H, int W, int C, int s, const float* add, float* y, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(g && y && N > 0 && H > 0 && W > 0 && C % 4 == 0 && s >= 1);
    pixel_shuffle_kernel<<<grid_for((int64_t)N * H * s * W * s * (C / 4)), 256, 0, stream>>>(g, N, H, W, C, s, add, y);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

static int im2col_impl(const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int pad_w, float* out, int64_t ldo, int round_out,
                       void* stream_, H3Out oh) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && (out || oh.p) && N > 0 && stride >= 1 && ldo >= (int64_t)KH * KW * C);
    const int OH = (H + 2 * pad - KH) / stride + 1, OW = (W + 2 * pad_w - KW) / stride + 1;
    im2col_kernel<<<grid_for((int64_t)N * OH * OW * ldo), 256, 0, stream>>>(x, N, H, W, C, KH, KW, stride, pad, pad_w, OH, OW, out, ldo, round_out, oh);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}
int siu3r_im2col_nhwc(const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int pad_w, float* out, int64_t ldo, int round_out, void* stream) {
    SIU3R_REQUIRE(out != nullptr);
    return im2col_impl(x, N, H, W, C, KH, KW, stride, pad, pad_w, out, ldo, round_out, stream, H3Out{nullptr, 0, 0});
}
// ... with the column matrix stored as an fp16 plane pair [rows, ldo] (pad columns K..ldo-1 are zero)
int siu3r_im2col_nhwc_h3(const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int pad_w, void* outh, int64_t ldo, int64_t plane,
                         void* stream) {
    SIU3R_REQUIRE(outh != nullptr && plane > 0);
    return im2col_impl(x, N, H, W, C, KH, KW, stride, pad, pad_w, nullptr, ldo, 0, stream, H3Out{(__half*)outh, ldo, plane});
}

int siu3r_nchw_to_nhwc(const float* x, float* y, int N, int C, int HW, int64_t ldy, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && y && N > 0 && C > 0 && HW > 0 && ldy >= C);
    nchw_to_nhwc_kernel<<<grid_for((int64_t)N * C * HW), 256, 0, stream>>>(x, y, N, C, HW, ldy);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_nhwc_to_nchw(const float* x, int64_t ldx, float* y, int N, int C, int HW, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && y && N > 0 && C > 0 && HW > 0 && ldx >= C);
    nhwc_to_nchw_kernel<<<grid_for((int64_t)N * C * HW), 256, 0, stream>>>(x, ldx, y, N, C, HW);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_maxpool3x3s2_nhwc(const float* x, int N, int H, int W, int C, float* y, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && y && C % 4 == 0);
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    maxpool3x3s2_kernel<<<grid_for((int64_t)N * OH * OW * (C / 4)), 256, 0, stream>>>(x, N, H, W, C, y, OH, OW);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_dwconv3x3_nhwc(const float* x, int64_t ldx, int64_t batch_stride_x, int N, int H, int W, int C, const float* w, const float* b,
                         float* y, int64_t ldy, int64_t batch_stride_y, int gelu, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && y && w && b && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0);
    dwconv3x3_kernel<<<grid_for((int64_t)N * H * W * (C / 4)), 256, 0, stream>>>(x, ldx, N, H, W, C, w, b, y, ldy, batch_stride_x, batch_stride_y, gelu);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// stats_ws: device scratch of N*groups*2 doubles (zeroed here)
int siu3r_groupnorm_nhwc(const float* x, int N, int HW, int C, int groups, const float* w, const float* b, float eps, int relu, float* y,
                         double* stats_ws, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && y && w && b && stats_ws && groups > 0 && groups <= 32 && C % groups == 0);
    SIU3R_REQUIRE(C % 4 == 0 && (C / groups) % 4 == 0 && 256 % (C / 4) == 0);
    SIU3R_CUDA_CHECK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * N * groups, stream));
    const int chunks = max(1, min(256, HW / 64));
    groupnorm_stats_kernel<<<dim3(chunks, N), 256, 0, stream>>>(x, HW, C, groups, stats_ws);
    const int64_t total4 = (int64_t)N * HW * (C / 4);
    groupnorm_apply_kernel<<<grid_for(total4), 256, 0, stream>>>(x, HW, C, groups, stats_ws, w, b, eps, relu, y, total4);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(2);
    return SIU3R_OK;
}

}  // extern "C"
