// Output-side kernels of the hot path: depth post-process, Gaussian adapter, panoptic post-process, reference fp32 GEMM.
#include "common.cuh"

namespace {

// pts = xyz / max(||xyz||, 1e-8) * expm1(||xyz||)     (heads/postprocess.py:46-61, mode "exp", no bounds)
__global__ void __launch_bounds__(256) depth_exp_kernel(const float* __restrict__ xyz, int64_t ldx, float* __restrict__ pts, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xyz[i * ldx], y = xyz[i * ldx + 1], z = xyz[i * ldx + 2];
    const float d = sqrtf(x * x + y * y + z * z);
    const float dc = fmaxf(d, 1e-8f);
    const float e = expm1f(d);
    pts[i * 3] = x / dc * e;
    pts[i * 3 + 1] = y / dc * e;
    pts[i * 3 + 2] = z / dc * e;
}

// UnifiedGaussianAdapter.forward (gaussian_adapter.py:81-110): raw [G, 83] = (opacity 1, scales 3, rot 4 xyzw, sh 3x25)
// The CTA's raw block (128 x 83 floats, contiguous) is staged through shared memory so every HBM access is coalesced.
constexpr int GA_THREADS = 128;
constexpr int GA_RAW = 83;
constexpr int GA_DSH = 25;

__global__ void __launch_bounds__(GA_THREADS) gaussian_adapter_kernel(const float* __restrict__ raw, int64_t G, float* __restrict__ cov,
                                                                      float* __restrict__ harm, float* __restrict__ opac,
                                                                      float* __restrict__ scales, float* __restrict__ rots) {
    __shared__ float s_raw[GA_THREADS * GA_RAW];
    const int64_t base = (int64_t)blockIdx.x * GA_THREADS;
    const int nvalid = (int)min((int64_t)GA_THREADS, G - base);
    const int total = nvalid * GA_RAW;
    const float* src = raw + base * GA_RAW;
    for (int i = threadIdx.x; i < total; i += GA_THREADS) s_raw[i] = src[i];
    __syncthreads();
    // harmonics [G,3,25] = sh * mask(degree): flat, coalesced
    float* hdst = harm + base * 75;
    for (int i = threadIdx.x; i < nvalid * 75; i += GA_THREADS) {
        const int gq = i / 75, j = i % 75;
        const int k = j % GA_DSH;
        float mk = 1.0f;
        if (k >= 16) mk = 0.1f * 0.00390625f;       // 0.1 * 0.25^4
        else if (k >= 9) mk = 0.1f * 0.015625f;     // 0.1 * 0.25^3
        else if (k >= 4) mk = 0.1f * 0.0625f;       // 0.1 * 0.25^2
        else if (k >= 1) mk = 0.1f * 0.25f;         // 0.1 * 0.25
        hdst[i] = s_raw[gq * GA_RAW + 8 + j] * mk;
    }
    if (threadIdx.x >= nvalid) return;
    const float* r = s_raw + threadIdx.x * GA_RAW;
    const int64_t gi = base + threadIdx.x;
    opac[gi] = 1.0f / (1.0f + expf(-r[0]));
    float sc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float x = r[1 + k];
        const float sp = x > 20.f ? x : log1pf(expf(x));  // F.softplus (beta 1, threshold 20)
        sc[k] = fminf(0.001f * sp, 0.3f);
        scales[gi * 3 + k] = sc[k];
    }
    const float q0 = r[4], q1 = r[5], q2 = r[6], q3 = r[7];
    rots[gi * 4 + 0] = q0; rots[gi * 4 + 1] = q1; rots[gi * 4 + 2] = q2; rots[gi * 4 + 3] = q3;  // returned raw (:109)
    const float nrm = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3) + 1e-8f;
    const float i = q0 / nrm, j = q1 / nrm, k = q2 / nrm, w = q3 / nrm;  // xyzw
    const float two_s = 2.0f / ((i * i + j * j + k * k + w * w) + 1e-8f);
    float R[3][3];
    R[0][0] = 1.f - two_s * (j * j + k * k); R[0][1] = two_s * (i * j - k * w); R[0][2] = two_s * (i * k + j * w);
    R[1][0] = two_s * (i * j + k * w); R[1][1] = 1.f - two_s * (i * i + k * k); R[1][2] = two_s * (j * k - i * w);
    R[2][0] = two_s * (i * k - j * w); R[2][1] = two_s * (j * k + i * w); R[2][2] = 1.f - two_s * (i * i + j * j);
    // cov = R S S^T R^T
    float RS[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) RS[a][b] = R[a][b] * sc[b];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            // ((R S) S^T) R^T evaluated left to right like the reference's chained matmul
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc += (RS[a][c] * sc[c]) * R[b][c];
            cov[gi * 9 + a * 3 + b] = acc;
        }
}

// ---- panoptic post-process helpers (image_processing_video_mask2former.py:1238-1481) ----
// y[n,oh,ow,j] = bilinear(x[n,:,:,idx[j]]) (align_corners False): resize of the kept queries' mask probabilities (:1386-1391)
__global__ void __launch_bounds__(256) resize_select_kernel(const float* __restrict__ x, int N, int H, int W, int C, const int* __restrict__ idx,
                                                            int nsel, float* __restrict__ y, int OH, int OW, float sh, float sw) {
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t0 >= (int64_t)N * OH * OW * nsel) return;
    const int j = (int)(t0 % nsel);
    int64_t t = t0 / nsel;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH);
    const int n = (int)(t / OH);
    float srch = sh * ((float)oh + 0.5f) - 0.5f; if (srch < 0.f) srch = 0.f;
    float srcw = sw * ((float)ow + 0.5f) - 0.5f; if (srcw < 0.f) srcw = 0.f;
    int h0 = min((int)srch, H - 1), w0 = min((int)srcw, W - 1);
    const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
    const float lh = srch - (float)h0, lw = srcw - (float)w0;
    const int c = idx[j];
    const float* b = x + (int64_t)n * H * W * C + c;
    const float v00 = b[((int64_t)h0 * W + w0) * C], v01 = b[((int64_t)h0 * W + w1) * C];
    const float v10 = b[((int64_t)h1 * W + w0) * C], v11 = b[((int64_t)h1 * W + w1) * C];
    y[t0] = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
}

// per pixel: k* = argmax_k probs[p,k]*score[k] (first max wins); area[k*]++ ; orig[k] += (probs[p,k]*score[k] >= thr)
__global__ void __launch_bounds__(256) argmax_area_kernel(const float* __restrict__ probs, int64_t npix, int nq, const float* __restrict__ score,
                                                          float thr, int32_t* __restrict__ labels, int32_t* __restrict__ area,
                                                          int32_t* __restrict__ orig) {
    extern __shared__ int32_t s_cnt[];  // [2*nq]
    for (int i = threadIdx.x; i < 2 * nq; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npix) {
        const float* r = probs + p * nq;
        float best = -INFINITY; int bk = 0;
        for (int k = 0; k < nq; ++k) {
            const float v = r[k] * score[k];
            if (v > best) { best = v; bk = k; }
            if (v >= thr) atomicAdd(&s_cnt[nq + k], 1);
        }
        labels[p] = bk;
        atomicAdd(&s_cnt[bk], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nq; i += blockDim.x) {
        if (s_cnt[i]) atomicAdd(&area[i], s_cnt[i]);
        if (s_cnt[nq + i]) atomicAdd(&orig[i], s_cnt[nq + i]);
    }
}

// segmentation / semantic / instance maps from the per-pixel argmax and the host-decided per-query LUTs
__global__ void __launch_bounds__(256) label_lut_kernel(const int32_t* __restrict__ labels, int64_t npix, const int32_t* __restrict__ seg_lut,
                                                        const int32_t* __restrict__ sem_lut, int32_t* __restrict__ seg, int32_t* __restrict__ sem,
                                                        int32_t* __restrict__ inst) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int k = labels[p];
    const int s = seg_lut[k];
    seg[p] = s;
    sem[p] = sem_lut[k];
    inst[p] = s;
}

// out[p, j, c] = probs[p, keep[j]] * class_probs[j, c]      (:1463-1467 + model.py:261-263 "(n h w) q c")
__global__ void __launch_bounds__(256) qc_logits_kernel(const float* __restrict__ probs, int64_t npix, int nq, const int* __restrict__ keep, int nk,
                                                        const float* __restrict__ cls /*[nk, ncls]*/, int ncls, float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = nk * ncls;
    if (t >= npix * per) return;
    const int64_t p = t / per;
    const int jc = (int)(t % per);
    const int j = jc / ncls;
    out[t] = probs[p * nq + keep[j]] * cls[jc];
}

// mask -> boolean attention mask: m[b, q, t*h*w] = sigmoid(bilinear(mask_logits[b,t,:,:,q] -> (h,w))) < 0.5
// (mask2former/video_seg_decoder.py:1461-1478; logits stored pixel-major [B*T, Hm, Wm, Q])
__global__ void __launch_bounds__(256) attn_mask_kernel(const float* __restrict__ logits, int BT, int T, int Hm, int Wm, int Q, int oh_, int ow_,
                                                        float sh, float sw, uint8_t* __restrict__ mask) {
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t0 >= (int64_t)BT * oh_ * ow_ * Q) return;
    const int q = (int)(t0 % Q);
    int64_t t = t0 / Q;
    const int ow = (int)(t % ow_); t /= ow_;
    const int oh = (int)(t % oh_);
    const int bt = (int)(t / oh_);
    float srch = sh * ((float)oh + 0.5f) - 0.5f; if (srch < 0.f) srch = 0.f;
    float srcw = sw * ((float)ow + 0.5f) - 0.5f; if (srcw < 0.f) srcw = 0.f;
    const int h0 = min((int)srch, Hm - 1), w0 = min((int)srcw, Wm - 1);
    const int h1 = h0 + (h0 < Hm - 1 ? 1 : 0), w1 = w0 + (w0 < Wm - 1 ? 1 : 0);
    const float lh = srch - (float)h0, lw = srcw - (float)w0;
    const float* b = logits + (int64_t)bt * Hm * Wm * Q + q;
    const float v00 = b[((int64_t)h0 * Wm + w0) * Q], v01 = b[((int64_t)h0 * Wm + w1) * Q];
    const float v10 = b[((int64_t)h1 * Wm + w0) * Q], v11 = b[((int64_t)h1 * Wm + w1) * Q];
    const float v = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
    const float sg = 1.0f / (1.0f + expf(-v));
    const int bidx = bt / T, tt = bt % T;
    mask[(((int64_t)bidx * Q + q) * T + tt) * oh_ * ow_ + (int64_t)oh * ow_ + ow] = sg < 0.5f ? 1 : 0;
}

// Plain fp32 (FFMA) GEMM: C = act(alpha * A W^T + bias) + residual.  Shape-agnostic path for tiny / odd shapes: the K = 9 intrinsics
// encoder (backbone_croco.py:59) and the ~80 GEMMs per pair with M <= 128 rows (the 100 Mask2Former queries,
// video_seg_decoder.py:957-1025,1423-1480), whose cost on the tensor-core kernels is pure launch / TMEM / TMA set-up latency.
// 32 x 64 output tile per CTA, 2 x 4 outputs per thread, operands staged K-major-transposed in shared memory.
constexpr int SG_M = 32, SG_N = 64, SG_K = 16;
__global__ void __launch_bounds__(256) gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int64_t lda,
                                                        const float* __restrict__ W, int64_t ldw, float* __restrict__ C, int64_t ldc,
                                                        const float* __restrict__ bias, const float* __restrict__ res, int64_t ldr, int act,
                                                        float alpha) {
    __shared__ float sA[SG_K][SG_M + 4];
    __shared__ __align__(16) float sW[SG_K][SG_N + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * SG_M, n0 = blockIdx.x * SG_N;
    float acc[2][4] = {};
    const int ar = tid >> 3, ak = (tid & 7) * 2;        // A tile: 32 rows x 16 k, two k per thread
    const int wr = tid >> 2, wk = (tid & 3) * 4;        // W tile: 64 rows x 16 k, four k per thread
    for (int k0 = 0; k0 < K; k0 += SG_K) {
        {
            const int m = m0 + ar;
            const float* ap = A + (int64_t)m * lda + k0 + ak;
            sA[ak][ar] = (m < M && k0 + ak < K) ? ap[0] : 0.f;
            sA[ak + 1][ar] = (m < M && k0 + ak + 1 < K) ? ap[1] : 0.f;
            const int n = n0 + wr;
            const float* wp = W + (int64_t)n * ldw + k0 + wk;
#pragma unroll
            for (int j = 0; j < 4; ++j) sW[wk + j][wr] = (n < N && k0 + wk + j < K) ? wp[j] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SG_K; ++k) {
            const float a0 = sA[k][ty * 2], a1 = sA[k][ty * 2 + 1];
            const float4 w = *reinterpret_cast<const float4*>(&sW[k][tx * 4]);
            acc[0][0] = fmaf(a0, w.x, acc[0][0]); acc[0][1] = fmaf(a0, w.y, acc[0][1]); acc[0][2] = fmaf(a0, w.z, acc[0][2]); acc[0][3] = fmaf(a0, w.w, acc[0][3]);
            acc[1][0] = fmaf(a1, w.x, acc[1][0]); acc[1][1] = fmaf(a1, w.y, acc[1][1]); acc[1][2] = fmaf(a1, w.z, acc[1][2]); acc[1][3] = fmaf(a1, w.w, acc[1][3]);
        }
        __syncthreads();
    }
    const bool rnd = (act & 4) != 0;
    const int a = act & 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float x = acc[i][j] * alpha;
            if (bias) x += bias[n];
            if (a == 1) x = 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
            else if (a == 2) x = fmaxf(x, 0.f);
            if (res) x += res[(int64_t)m * ldr + n];
            if (rnd) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); x = __uint_as_float(r); }
            C[(int64_t)m * ldc + n] = x;
        }
    }
}

// Skinny GEMM for M <= 128 rows (the 100 Mask2Former queries: ~85 linear layers per pair with 256..2048 columns, video_seg_decoder.py:957-1025,
// 1423-1480).  On the tensor-core kernels these are pure set-up latency (TMEM allocation, barrier init, TMA descriptor fetch, one CTA row).
// Here: 32 x 32 output tile per CTA, the 8 warps split K 8 ways (K = 256 -> one 32-wide chunk per warp, all of its loads in flight at once),
// each lane accumulates a 4 x 8 register block, partial tiles are reduced through shared memory in a fixed order.  fp32 FFMA (exact products).
constexpr int SK_T = 32, SK_WARPS = 8;
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_kernel(int M, int N, int K, const float* __restrict__ A, int64_t lda,
                                                                    const float* __restrict__ W, int64_t ldw, float* __restrict__ C, int64_t ldc,
                                                                    const float* __restrict__ bias, const float* __restrict__ res, int64_t ldr,
                                                                    int act, float alpha) {
    extern __shared__ float sk_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sA = sk_smem + warp * (2 * SK_T * (SK_T + 1));     // [32 rows][33]
    float* sW = sA + SK_T * (SK_T + 1);                       // [32 cols][33]
    const int m0 = blockIdx.y * SK_T, n0 = blockIdx.x * SK_T;
    const int kper = ((K + SK_WARPS - 1) / SK_WARPS + 31) / 32 * 32;      // K slice of this warp, multiple of 32
    const int kbeg = warp * kper, kend = min(K, kbeg + kper);
    const int rg = lane >> 2, cg = lane & 3;                  // lane block: rows rg*4 .. +3, columns cg*8 .. +7
    float acc[4][8] = {};
    for (int k0 = kbeg; k0 < kend; k0 += 32) {
        const int k = k0 + lane;
#pragma unroll 8
        for (int r = 0; r < SK_T; ++r) {
            const int m = m0 + r, n = n0 + r;
            sA[r * (SK_T + 1) + lane] = (m < M && k < kend) ? __ldg(A + (int64_t)m * lda + k) : 0.f;
            sW[r * (SK_T + 1) + lane] = (n < N && k < kend) ? __ldg(W + (int64_t)n * ldw + k) : 0.f;
        }
        __syncwarp();
#pragma unroll 4
        for (int kk = 0; kk < 32; ++kk) {
            float a[4], w[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[(rg * 4 + i) * (SK_T + 1) + kk];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = sW[(cg * 8 + j) * (SK_T + 1) + kk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncwarp();
    }
    __syncthreads();                       // every warp is done with its staging tiles: reuse the shared memory for the partial tiles
    float* part = sk_smem + warp * (SK_T * SK_T);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) part[(rg * 4 + i) * SK_T + cg * 8 + j] = acc[i][j];
    __syncthreads();
    const bool rnd = (act & 4) != 0;
    const int a_ = act & 3;
    for (int e = threadIdx.x; e < SK_T * SK_T; e += SK_WARPS * 32) {
        const int r = e >> 5, c = e & 31;
        const int m = m0 + r, n = n0 + c;
        if (m >= M || n >= N) continue;
        float x = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < SK_WARPS; ++w8) x += sk_smem[w8 * (SK_T * SK_T) + e];
        x *= alpha;
        if (bias) x += bias[n];
        if (a_ == 1) x = 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
        else if (a_ == 2) x = fmaxf(x, 0.f);
        if (res) x += res[(int64_t)m * ldr + n];
        if (rnd) { uint32_t q; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q) : "f"(x)); x = __uint_as_float(q); }
        C[(int64_t)m * ldc + n] = x;
    }
}

inline unsigned grid_for(int64_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

// ---- PLY vertex records (src/utils/ply_export.py:30-97): one 32-bit word per thread, record-major so that the D2H copy of the
//      packed buffer IS the file body.  Field order: x y z | nx ny nz (0) | f_dc_0..2 | f_rest (harmonics[..., 1:] flattened
//      channel-major, omitted when dc_only) | opacity | log(scale_0..2) | rot w x y z (stored xyzw) | semantic_label i4 |
//      instance_label i4 | seg_query_class_logits (q*c floats).  HBM-bound: ~(95 + qc) words read, F words written per Gaussian.
__global__ void __launch_bounds__(256) ply_pack_kernel(const float* __restrict__ means, const float* __restrict__ scales,
                                                       const float* __restrict__ rot, const float* __restrict__ harm,
                                                       const float* __restrict__ opac, const int32_t* __restrict__ sem,
                                                       const int32_t* __restrict__ inst, const float* __restrict__ qc, int64_t G, int d_sh,
                                                       int n_rest, int has_labels, int qc_words, int F, uint32_t* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * F) return;
    const int64_t g = t / F;
    int f = (int)(t - g * F);
    uint32_t w;
    if (f < 3) w = __float_as_uint(means[3 * g + f]);
    else if (f < 6) w = 0u;
    else if (f < 9) w = __float_as_uint(harm[(3 * g + (f - 6)) * d_sh]);
    else {
        f -= 9;
        if (f < n_rest) {
            const int ch = f / (d_sh - 1), k = f - ch * (d_sh - 1) + 1;
            w = __float_as_uint(harm[(3 * g + ch) * d_sh + k]);
        } else {
            f -= n_rest;
            if (f == 0) w = __float_as_uint(opac[g]);
            else if (f < 4) w = __float_as_uint(logf(scales[3 * g + (f - 1)]));
            else if (f < 8) w = __float_as_uint(rot[4 * g + ((f - 4 + 3) & 3)]);   // (w, x, y, z) from (x, y, z, w)
            else {
                f -= 8;
                if (has_labels && f < 2) w = (uint32_t)(f == 0 ? sem[g] : inst[g]);
                else w = __float_as_uint(qc[g * qc_words + (f - (has_labels ? 2 : 0))]);
            }
        }
    }
    out[t] = w;
}

// Render record of the joint-scene all-gather: what the rasterizer consumes, 85 floats per Gaussian =
// means 3 | covariance upper triangle 6 (xx xy xz yy yz zz: the cov3D_precomp of cuda_splatting.py:107,115) | harmonics 75 | opacity 1.
// One thread per word, record-major output: one coalesced pass instead of a torch.cat of four tensors.
__global__ void __launch_bounds__(256) render_record_pack_kernel(const float* __restrict__ means, const float* __restrict__ cov,
                                                                 const float* __restrict__ harm, const float* __restrict__ opac, int64_t G,
                                                                 float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * 85) return;
    const int64_t g = t / 85;
    const int f = (int)(t - g * 85);
    float w;
    if (f < 3) w = means[3 * g + f];
    else if (f < 9) {
        const int k = f - 3;                       // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
        const int idx = k < 3 ? k : (k < 5 ? k + 1 : 8);
        w = cov[9 * g + idx];
    } else if (f < 84) w = harm[75 * g + (f - 9)];
    else w = opac[g];
    out[t] = w;
}
__global__ void __launch_bounds__(256) render_record_unpack_kernel(const float* __restrict__ rec, int64_t G, float* __restrict__ means,
                                                                   float* __restrict__ cov6, float* __restrict__ harm, float* __restrict__ opac) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * 85) return;
    const int64_t g = t / 85;
    const int f = (int)(t - g * 85);
    const float w = rec[t];
    if (f < 3) means[3 * g + f] = w;
    else if (f < 9) cov6[6 * g + (f - 3)] = w;
    else if (f < 84) harm[75 * g + (f - 9)] = w;
    else opac[g] = w;
}

extern "C" {

// Gaussians -> packed render records [G, 85] (see render_record_pack_kernel) and back (covariances come back as the [G, 6] upper triangle the
// rasterizer takes with cov_stride = 6).  The record is the payload of the one NCCL all-gather of the path (siu3r_b200/parallel.py).
int siu3r_render_record_pack(const float* means, const float* cov33, const float* harmonics, const float* opacities, int64_t G, float* out,
                             void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(means && cov33 && harmonics && opacities && out && G > 0);
    render_record_pack_kernel<<<grid_for(G * 85), 256, 0, stream>>>(means, cov33, harmonics, opacities, G, out);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}
int siu3r_render_record_unpack(const float* rec, int64_t G, float* means, float* cov6, float* harmonics, float* opacities, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(rec && means && cov6 && harmonics && opacities && G > 0);
    render_record_unpack_kernel<<<grid_for(G * 85), 256, 0, stream>>>(rec, G, means, cov6, harmonics, opacities);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_depth_exp(const float* xyz, int64_t ldx, float* pts, int64_t n, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(xyz && pts && n > 0 && ldx >= 3);
    depth_exp_kernel<<<grid_for(n), 256, 0, stream>>>(xyz, ldx, pts, n);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// raw [G,83] -> covariances [G,3,3], harmonics [G,3,25], opacities [G], scales [G,3], rotations [G,4] (raw quaternion)
int siu3r_gaussian_adapter(const float* raw, int64_t G, float* covariances, float* harmonics, float* opacities, float* scales,
                           float* rotations, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(raw && covariances && harmonics && opacities && scales && rotations && G > 0);
    gaussian_adapter_kernel<<<grid_for(G, GA_THREADS), GA_THREADS, 0, stream>>>(raw, G, covariances, harmonics, opacities, scales, rotations);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_resize_select(const float* x, int N, int H, int W, int C, const int* idx, int nsel, float* y, int OH, int OW, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && idx && y && nsel > 0);
    resize_select_kernel<<<grid_for((int64_t)N * OH * OW * nsel), 256, 0, stream>>>(x, N, H, W, C, idx, nsel, y, OH, OW, (float)H / (float)OH,
                                                                                   (float)W / (float)OW);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_argmax_area(const float* probs, int64_t npix, int nq, const float* score, float thr, int32_t* labels, int32_t* area, int32_t* orig,
                      void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(probs && score && labels && area && orig && nq > 0 && nq <= 1024);
    SIU3R_CUDA_CHECK(cudaMemsetAsync(area, 0, sizeof(int32_t) * nq, stream));
    SIU3R_CUDA_CHECK(cudaMemsetAsync(orig, 0, sizeof(int32_t) * nq, stream));
    argmax_area_kernel<<<grid_for(npix), 256, 2 * nq * sizeof(int32_t), stream>>>(probs, npix, nq, score, thr, labels, area, orig);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_label_lut(const int32_t* labels, int64_t npix, const int32_t* seg_lut, const int32_t* sem_lut, int32_t* seg, int32_t* sem, int32_t* inst,
                    void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(labels && seg_lut && sem_lut && seg && sem && inst);
    label_lut_kernel<<<grid_for(npix), 256, 0, stream>>>(labels, npix, seg_lut, sem_lut, seg, sem, inst);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_qc_logits(const float* probs, int64_t npix, int nq, const int* keep, int nk, const float* cls, int ncls, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(probs && keep && cls && out && nk > 0);
    qc_logits_kernel<<<grid_for(npix * nk * ncls), 256, 0, stream>>>(probs, npix, nq, keep, nk, cls, ncls, out);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_attn_mask_from_logits(const float* logits, int B, int T, int Hm, int Wm, int Q, int oh, int ow, uint8_t* mask, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(logits && mask);
    attn_mask_kernel<<<grid_for((int64_t)B * T * oh * ow * Q), 256, 0, stream>>>(logits, B * T, T, Hm, Wm, Q, oh, ow, (float)Hm / (float)oh,
                                                                               (float)Wm / (float)ow, mask);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int siu3r_gemm_simt(int M, int N, int K, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, const float* bias,
                    const float* residual, int64_t ldr, int act, float alpha, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(M > 0 && N > 0 && K > 0 && A && W && C);
    dim3 grid(ceil_div(N, SG_N), ceil_div(M, SG_M));
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(M, N, K, A, lda, W, ldw, C, ldc, bias, residual, ldr, act, alpha);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}


// Packed PLY vertex records for export_ply (src/utils/ply_export.py:30-97).  out: G records of F = siu3r_ply_record_words(...) 32-bit words.
int siu3r_ply_record_words(int d_sh, int dc_only, int has_labels, int qc_words) {
    return 9 + (dc_only ? 0 : 3 * (d_sh - 1)) + 8 + (has_labels ? 2 : 0) + qc_words;
}
int siu3r_ply_pack(const float* means, const float* scales, const float* rotations, const float* harmonics, const float* opacities,
                   const int32_t* semantic_labels, const int32_t* instance_labels, const float* qc_logits, int64_t G, int d_sh, int dc_only,
                   int qc_words, uint32_t* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(means && scales && rotations && harmonics && opacities && out && G > 0 && d_sh >= 1 && qc_words >= 0);
    SIU3R_REQUIRE((semantic_labels == nullptr) == (instance_labels == nullptr));
    SIU3R_REQUIRE(qc_words == 0 || qc_logits);
    const int has_labels = semantic_labels ? 1 : 0;
    const int n_rest = dc_only ? 0 : 3 * (d_sh - 1);
    const int F = siu3r_ply_record_words(d_sh, dc_only, has_labels, qc_words);
    ply_pack_kernel<<<grid_for(G * F), 256, 0, stream>>>(means, scales, rotations, harmonics, opacities, semantic_labels, instance_labels, qc_logits,
                                                         G, d_sh, n_rest, has_labels, qc_words, F, out);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}


// act: 0 none, 1 GELU(erf), 2 ReLU; +4 = store RN_tf32(result).  Any alignment / K.
int siu3r_gemm_skinny(int M, int N, int K, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, const float* bias,
                      const float* residual, int64_t ldr, int act, float alpha, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(M > 0 && N > 0 && K > 0 && A && W && C);
    constexpr int smem = SK_WARPS * 2 * SK_T * (SK_T + 1) * 4;   // 67.6 KB (staging) >= 32 KB (partials)
    static bool attr[64] = {false};
    if (siu3r_first_use_on_device(attr)) SIU3R_CUDA_CHECK(cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid(ceil_div(N, SK_T), ceil_div(M, SK_T));
    gemm_skinny_kernel<<<grid, SK_WARPS * 32, smem, stream>>>(M, N, K, A, lda, W, ldw, C, ldc, bias, residual, ldr, act, alpha);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

}  // extern "C"
