// fp32-grade GEMM / implicit-GEMM convolution on the fp16 tensor-core path ("h3" mode, see h3.cuh), sm_100a.
//
//   C[M,N] = epilogue( X[M,K] * W[N,K]^T )       X, W given as fp16 (hi, lo * 2^11) plane pairs, C as fp32 or as a plane pair
//
// Replaces, at the accuracy the north star asks for (Gaussians 1e-3 abs, logits 1e-4 rel) and at about the cost of the TF32 path,
// every nn.Linear / Conv2d on the per-pair hot path: croco/blocks.py:97,110,74-77,154-156,167, backbone_croco.py:87,
// heads/dpt_block.py:98-116,181-189,358-364,385-391, vit_adapter/blocks.py:118-121, mask2former/video_seg_decoder.py.
//
// ONE persistent, operand-swapped CTA-pair kernel (tcgen05 cta_group::2) runs all of them:
//   * weight rows on the 256-wide MMA-M axis (128 per CTA), tokens / pixels on the MMA-N axis with a host-chosen tile width tw
//     (multiple of 16, <= 256), so that (weight pairs x token tiles) fills whole rounds of the 74 resident CTA pairs;
//   * per 64-deep k-block each CTA stages W hi|lo (2 x 16 KB) and its half of the token tile hi|lo (2 x tw/2 rows x 128 B), one TMA
//     instruction per operand (the plane index is the outermost box dimension);  4 k-steps x 3 kind::f16 MMAs (M = 256, N = tw, K = 16):
//     x += W_lo X_hi, x += W_hi X_lo, hh += W_hi X_hi;
//   * two accumulators (hh, x) per tile in TMEM: tw <= 128 -> double-buffered tiles (4 x 128 columns: the epilogue of tile i runs under
//     the mainloop of tile i+1), tw > 128 -> one tile in flight (2 x 256 columns);
//   * linear mode: X is a [M, K] plane pair (3-D TMA box);  conv mode (stride 1, "same" output): the token tile is a 16 x tw/16 pixel
//     patch of one NHWC image, the k-block is (filter tap, 64-channel block) and the tap is a coordinate shift of a 5-D TMA box, the
//     halo zero-filled by TMA = the conv padding; channel counts that are no multiple of 64 ride on the same zero fill;
//   * epilogue: TMEM lane = weight row n, column = token m, so every global access of a warp is 32 consecutive n of one token row:
//     acc = hh + x * 2^-11, * alpha, + bias, GELU(erf) / ReLU, RoPE-2D (pairs = lanes l, l^16), + fp32 residual, then either an fp32
//     store or the (hi, lo) fp16 split of the result when its only consumers are further h3 tensor-core operands (or both: the fp32
//     residual stream plus its plane pair); the V columns of a fused qkv / k|v projection go out transposed (V^T plane pair) for the
//     attention kernel's P.V operand;
//   * LayerNorm fused on both sides (siu3r_gemm_h3_ln): a residual-adding projection accumulates fixed-point row statistics of the rows it
//     writes, the next projection multiplies the RAW rows by gamma-folded weights and normalises in its epilogue;
//   * layers with <= 128 output channels run M = 128 MMAs (64 weight rows per CTA); the last token tile of a linear problem is as wide as
//     its remaining rows; the tile width comes from a measured cost model (pick_tw); launched with programmatic dependent launch.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "h3.cuh"

namespace {
using namespace h3;

constexpr int BKH = 64;                        // fp16 elements per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int W_ROWS = 128;                    // weight rows per CTA
constexpr int W_PLANE_BYTES = W_ROWS * 128;    // 16 KB
constexpr int W_BYTES = 2 * W_PLANE_BYTES;     // hi + lo
constexpr int PIPE_BYTES = 200 * 1024;
constexpr int MAX_STAGES = 6;
constexpr int LN_BYTES = 2 * 256 * 8;         // fused LayerNorm: (mean, rstd) of the tokens of the tile in flight, double buffered
constexpr int SMEM_BYTES = PIPE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + LN_BYTES;
constexpr int EPI_WARPS = 16;                  // four warps per TMEM lane quarter (they split the token columns): the epilogue of a GELU / RoPE
                                               // tile costs more issue slots than 8 warps provide under one tile's mainloop
constexpr int THREADS = 64 + 32 * EPI_WARPS;   // + TMA producer warp + MMA issuer warp
constexpr int CLUSTERS = 74;                   // TPC pairs of a B200 (148 SMs)

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_MASK = 3 };

struct H3Params {
    int N, num_kb;             // weight rows, k-blocks
    int tw, nbuf, nstages;     // token tile width, accumulator sets in TMEM (1 or 2), pipeline depth
    int mhalf;                 // N <= 128: MMA M = 128 (64 weight rows per CTA) instead of 256 -- no tensor time is spent on padding rows
    int act; float alpha;
    int64_t ldc;               // fp32 output pitch
    int64_t ldh, plane_h;      // split output pitch / plane distance (elements)
    int64_t ldr;               // fp32 residual pitch
    const long long* rope_pos; const float* rope_tab; int rope_cols;
    int vt_col0; int64_t vt_ld, vt_plane;
    unsigned long long* dbg_ts; // tuning aid: per-tile clock64 stamps of CTA pair 0 (8 per tile: MMA wait / start / issued, epilogue wait / start / end)
    int dbg_mode;              // timing experiments only (results are garbage): 1 = TMA pipeline without MMAs, 2 = MMAs without TMA loads
    int order;                 // tile order: 0 = weight pair fastest (neighbouring CTA pairs share the token tile), 1 = token tile fastest (share the weights)
    int ttiles0, ttiles1;      // token tiles of problem 0 / 1
    float ln_inv_c, ln_eps;    // fused LayerNorm on the A operand (H3Problem::stats_in): 1 / row length, epsilon
    float lo_scale;            // factor on the lo plane of split outputs: 2^11 (GEMM operands) or 1 (attention operands, see flash_h3.cu)
    // conv mode
    int conv, H, W, Cin, KW, pad_h, pad_w, tiles_w, tiles_per_img, cblocks;
};
struct H3Problem {       // what differs between the problems of a grouped launch
    float* C;            // fp32 output or null
    __half* Ch;          // split output (hi plane) or null
    const float* bias;
    const float* residual;
    int M;               // rows (linear) / images (conv)
    __half* vt;          // V^T destination (hi plane) of this problem's rows, or null
    int vt_cols;         // columns of a V^T row this problem owns (>= M, multiple of 8): [M, vt_cols) is zero-filled
    // Fused LayerNorm (linear mode).  Row statistics are int64 fixed-point pairs (sum x * 2^32, sum x^2 * 2^26) per row: integer atomics add in any order
    // to the same bits, so the statistics -- and everything downstream -- are reproducible run to run.
    const long long* stats_in;   // consumer: the A operand is the RAW row x, the weights carry gamma; the epilogue applies rstd*(acc - mu*ln_s[n]) + bias
    const float* ln_s;           //           ln_s[n] = sum_k gamma_k W[n,k]  (bias[n] = sum_k beta_k W[n,k] + b[n] is folded by the host)
    long long* stats_out;        // producer: accumulates the statistics of the rows it writes (every column block adds its 32-column partial sums)
};
// sum x in 2^-32 steps (range +-2.1e9), sum x^2 in 2^-26 steps: range 1.4e11 covers 1024 columns of fp16-representable magnitudes (one 65504 outlier is
// 4.3e9), and the 1.5e-8 step is 1e-5 of the 1e-6 epsilon that floors the variance anyway (rows of magnitude 1e-3 keep rstd to 1.5e-5 relative).
constexpr float STATS_SCALE = 4294967296.0f, STATS_SCALE_SQ = 67108864.0f;
struct H3Group {
    H3Problem prob[2]; int tiles0;
    int tw_r[2];         // linear mode: width of each problem's LAST token tile (its remaining rows rounded up to the MMA-N granularity, <= tw): M = 2050 =
                         // 16 x 128 + 2 would otherwise spend a full 128-wide tile on two tokens (1025 = 8 x 128 + 1 in the decoder: one tile in nine)
};

struct Frag {            // where the 16 tokens of one epilogue fragment live: 16 consecutive output rows
    int64_t rb;          // first output row
    int nv;              // valid tokens (0..16)
};

template <int ACTK, bool ROPE, bool RES, bool SPLIT>
__device__ __forceinline__ void epi_chunk(float (&v)[16], const Frag& f, int lane, int n, bool n_ok, float bias, int axis, int mrow, int jmax,
                                          const H3Params& p, const H3Problem& pr) {
    float r[16];
    if (RES) {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = (j < f.nv && n_ok) ? __ldcg(pr.residual + (f.rb + j) * p.ldr + n) : 0.0f;
    }
    long long pos_l = 0;
    const float* tab = nullptr;
    if (ROPE) {
        pos_l = lane < jmax ? p.rope_pos[(int64_t)(mrow + lane) * 2 + axis] : 0;   // lane j holds the position of token j of the fragment
        tab = p.rope_tab + (lane & 15) * 2;
    }
    // Phase 1 (branch free: the warp stays converged, so the RoPE shuffles compile to plain SHFL): all arithmetic, results left in v[].
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float x = v[j] * p.alpha + bias;
        if (ACTK == ACT_GELU) x = gelu_erf(x);
        if (ACTK == ACT_RELU) x = fmaxf(x, 0.0f);
        if (ROPE) {
            const float y = __shfl_xor_sync(0xffffffffu, x, 16);
            const int pj = __shfl_sync(0xffffffffu, (int)pos_l, j);
            const float2 cs = __ldg(reinterpret_cast<const float2*>(tab + pj * 32));
            x = (lane & 16) ? x * cs.x + y * cs.y : x * cs.x - y * cs.y;
        }
        if (RES) x += r[j];
        v[j] = x;
    }
    // Phase 2: stores (one output row = 32 consecutive n of the warp: a 128-byte line in fp32, 64 bytes per plane as a plane pair).  Full fragments
    // (all but the last of a problem) take the unguarded path with running row pointers: the guarded form costs a branch and a 64-bit address
    // computation per store, a third of the instructions of a whole GELU + split fragment.
    if (f.nv <= 0 || !n_ok) return;
    if (SPLIT) {
        unsigned short* dh = reinterpret_cast<unsigned short*>(pr.Ch + f.rb * p.ldh + n);
        unsigned short* dl = dh + p.plane_h;
        if (f.nv == 16) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {      // two tokens per packed conversion (h3.cuh)
                uint32_t hi2, lo2;
                h3_split2_s(v[i], v[i + 1], p.lo_scale, hi2, lo2);
                dh[0] = (unsigned short)hi2; dl[0] = (unsigned short)lo2;
                dh += p.ldh; dl += p.ldh;
                dh[0] = (unsigned short)(hi2 >> 16); dl[0] = (unsigned short)(lo2 >> 16);
                dh += p.ldh; dl += p.ldh;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                uint32_t hi2, lo2;
                h3_split2_s(v[i], v[i + 1], p.lo_scale, hi2, lo2);
                if (i < f.nv) { dh[(int64_t)i * p.ldh] = (unsigned short)hi2; dl[(int64_t)i * p.ldh] = (unsigned short)lo2; }
                if (i + 1 < f.nv) { dh[(int64_t)(i + 1) * p.ldh] = (unsigned short)(hi2 >> 16); dl[(int64_t)(i + 1) * p.ldh] = (unsigned short)(lo2 >> 16); }
            }
        }
    } else {
        float* d = pr.C + f.rb * p.ldc + n;
        if (f.nv == 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { d[0] = v[i]; d += p.ldc; }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < f.nv) d[(int64_t)i * p.ldc] = v[i];
        }
    }
}

// V columns of a fused qkv / k|v projection: this lane's output column n is one row of V^T, the fragment's 16 tokens are 32 contiguous bytes
// of it per plane.  Tokens past M up to the padded pitch are zero-filled so that the attention kernel's TMA never stages uninitialised memory.
__device__ __forceinline__ void epi_chunk_vt(const float (&v)[16], int jmax, int n, bool n_ok, float bias, int mrow, const H3Params& p,
                                             const H3Problem& pr) {
    if (!n_ok) return;
    __half* dst = pr.vt + (int64_t)(n - p.vt_col0) * p.vt_ld + mrow;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        const float x0 = j < jmax ? v[j] * p.alpha + bias : 0.0f;
        const float x1 = j + 1 < jmax ? v[j + 1] * p.alpha + bias : 0.0f;
        h3_split2_s(x0, x1, p.lo_scale, hi[j >> 1], lo[j >> 1]);
    }
    // a fragment cut by the END OF THE PROBLEM (mrow + jmax == M) also zero-fills the pad columns [M, vt_cols) (fewer than 8 of them)
    const int lim = (mrow + jmax == pr.M) ? min(16, pr.vt_cols - mrow) : jmax;
#pragma unroll
    for (int j = 0; j < 16; j += 8)
        if (j < lim) {
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(hi[j >> 1], hi[(j >> 1) + 1], hi[(j >> 1) + 2], hi[(j >> 1) + 3]);
            *reinterpret_cast<uint4*>(dst + p.vt_plane + j) = make_uint4(lo[j >> 1], lo[(j >> 1) + 1], lo[(j >> 1) + 2], lo[(j >> 1) + 3]);
        }
}

// Per-tile, per-warp constants of the epilogue (kept in one struct so that the fragment loop can be a template on the epilogue variant).
struct EpiCtx {
    uint32_t tmem_hh, tmem_x, lane_off;
    int lane, c_lo, c_hi, nb, n, tok_off, axis, m_base, img, h0, w0;
    bool n_ok, ln, split;
    float bias, ln_s;
    const float2* ln_mr;
};

// The fragment loop of one warp and tile, instantiated per epilogue variant: the variant dispatch happens ONCE per tile, so the hot loop is one
// contiguous piece of code (with the switch inside the loop the executed path of a GELU + split fragment was scattered over 96 KB of SASS and a
// quarter of the epilogue warps' stall samples were instruction fetches).
template <int ACTK, bool ROPE, bool RES, bool SPLIT>
__device__ __forceinline__ void epi_frag_loop(const EpiCtx& c, const H3Params& p, const H3Problem& pr) {
    const int lane = c.lane, n = c.n, nb = c.nb, tok_off = c.tok_off, axis = c.axis, m_base = c.m_base, img = c.img, h0 = c.h0, w0 = c.w0;
    const bool n_ok = c.n_ok, ln = c.ln, split = c.split;
    const float bias = c.bias, ln_s = c.ln_s;
    const float2* ln_mr = c.ln_mr;
    const uint32_t tmem_hh = c.tmem_hh, tmem_x = c.tmem_x, lane_off = c.lane_off;
    const int c_lo = c.c_lo, c_hi = c.c_hi;
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi && nb < p.N; c0 += 16) {
        uint32_t a[16], b[16];
        tmem_ld16(tmem_hh + lane_off + (uint32_t)c0, a);
        tmem_ld16(tmem_x + lane_off + (uint32_t)c0, b);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(b[j]), H3_LO_INV, __uint_as_float(a[j]));
        Frag f;
        int mrow = m_base + tok_off + c0, jmax = 16;
        if (p.conv) {
            const int hrow = h0 + ((tok_off + c0) >> 4);
            f.rb = ((int64_t)img * p.H + hrow) * p.W + w0;
            f.nv = hrow < p.H ? 16 : 0;
        } else {
            if (jmax > pr.M - mrow) jmax = pr.M - mrow;
            if (jmax <= 0) break;
            f.rb = mrow;
            f.nv = jmax;
            if (ln) {   // acc -> rstd * (acc - mean * s_n); the (mean, rstd) pairs are warp-uniform shared-memory broadcasts
                const float2* mr = ln_mr + tok_off + c0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 t = mr[j];
                    v[j] = t.y * fmaf(-t.x, ln_s, v[j]);
                }
            }
            if (pr.vt != nullptr && nb >= p.vt_col0) {        // warp-uniform: a 32-column block never straddles vt_col0 (multiple of 64)
                epi_chunk_vt(v, jmax, n, n_ok, bias, mrow, p, pr);
                continue;
            }
        }
        epi_chunk<ACTK, ROPE, RES, SPLIT>(v, f, lane, n, n_ok, bias, axis, mrow, jmax, p, pr);
        // v[] now holds the final values of this fragment (16 tokens x this lane's column n)
        if (split && pr.C != nullptr && f.nv > 0 && n_ok) {     // dual output: the plane pair went out above, the fp32 copy (residual stream) here
            float* d = pr.C + f.rb * p.ldc + n;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < f.nv) d[(int64_t)i * p.ldc] = v[i];
        }
        if (pr.stats_out != nullptr) {
            // Row statistics for a LayerNorm fused into the NEXT GEMM: 32 values per lane (16 sums, 16 sums of squares) are reduced over the warp's 32
            // columns by a halving butterfly (31 shuffles); lane l ends up with statistic l >> 4 of token l & 15 and adds it, as a fixed-point
            // integer, to the row's accumulator.
            float w[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float x = (j < f.nv && n_ok) ? v[j] : 0.0f;
                w[j] = x; w[16 + j] = x * x;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                    const float send = up ? w[i] : w[i + off];
                    const float keep = up ? w[i + off] : w[i];
                    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            const int tj = lane & 15;
            if (tj < f.nv)
                atomicAdd(reinterpret_cast<unsigned long long*>(pr.stats_out) + (f.rb + tj) * 2 + (lane >> 4),
                          (unsigned long long)__float2ll_rn(w[0] * ((lane >> 4) ? STATS_SCALE_SQ : STATS_SCALE)));
        }
    }
}

// Epilogue of one warp: TMEM lanes [32q, 32q+32) = weight rows n, columns [c_lo, c_hi) = its share of the tile's tokens, 16 at a time.
// mhalf (M = 128 over the CTA pair): lanes [0, 64) hold this CTA's 64 weight rows for the first half of the tile's tokens, lanes [64, 128) the
// same rows for the second half (the accumulator is tw/2 columns wide); tok_off = token index of column 0 for this warp.
__device__ __forceinline__ void run_epilogue(uint32_t tmem_hh, uint32_t tmem_x, int lane, int q, int c_lo, int c_hi, int twt, int n_cta, int tt,
                                             const H3Params& p, const H3Problem& pr, uint64_t* bar, uint32_t parity, float2* ln_mr, int epi_tid,
                                             unsigned long long* ts_start) {
    const int nb = n_cta + (p.mhalf ? (q & 1) : q) * 32;           // warp-uniform first weight row
    const int tok_off = p.mhalf ? (q >> 1) * (twt >> 1) : 0;      // twt = width of THIS tile (the last token tile of a problem may be narrower)
    const int n = nb + lane;
    const bool n_ok = n < p.N;
    const float bias = (pr.bias && n_ok) ? __ldg(pr.bias + n) : 0.0f;
    const bool ln = pr.stats_in != nullptr;
    const float ln_s = (ln && n_ok) ? __ldg(pr.ln_s + n) : 0.0f;
    const int act = p.act & ACT_MASK;
    const bool rope = p.rope_pos != nullptr && nb < p.rope_cols;
    const bool res = pr.residual != nullptr;
    const bool split = pr.Ch != nullptr;
    const int axis = (nb >> 5) & 1;
    // warp-uniform variant id: branches are hoisted out of the 16-element inner loops
    const int variant = rope ? (split ? 1 : 0) : 2 + (act * 4 + (res ? 2 : 0) + (split ? 1 : 0));
    // tile coordinates
    int img = 0, h0 = 0, w0 = 0, m_base = tt * p.tw;
    if (p.conv) {
        img = tt / p.tiles_per_img;
        const int rem = tt - img * p.tiles_per_img;
        h0 = (rem / p.tiles_w) * (p.tw >> 4);
        w0 = (rem % p.tiles_w) * 16;
    }
    if (ln) {
        // Fused LayerNorm of the A operand: (mean, rstd) of the tile's tokens, once per CTA and tile, while the tile's MMAs are still running
        // (one token per epilogue thread; float64 for the E[x^2] - mean^2 cancellation).
        for (int t = epi_tid; t < p.tw; t += EPI_WARPS * 32) {
            float2 mr = make_float2(0.0f, 0.0f);
            if (m_base + t < pr.M) {
                const longlong2 st = __ldcg(reinterpret_cast<const longlong2*>(pr.stats_in) + (m_base + t));
                const double m = (double)st.x * (double)(p.ln_inv_c / STATS_SCALE);
                const double var = fmax((double)st.y * (double)(p.ln_inv_c / STATS_SCALE_SQ) - m * m, 0.0);
                mr = make_float2((float)m, rsqrtf((float)var + p.ln_eps));
            }
            ln_mr[t] = mr;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    }
    mbar_wait(bar, parity);
    tc_fence_after();
    if (ts_start) *ts_start = clock64();
    EpiCtx c;
    c.tmem_hh = tmem_hh; c.tmem_x = tmem_x; c.lane_off = (uint32_t)(q * 32) << 16;
    c.lane = lane; c.c_lo = c_lo; c.c_hi = c_hi; c.nb = nb; c.n = n; c.tok_off = tok_off; c.axis = axis; c.m_base = m_base; c.img = img; c.h0 = h0; c.w0 = w0;
    c.n_ok = n_ok; c.ln = ln; c.split = split; c.bias = bias; c.ln_s = ln_s; c.ln_mr = ln_mr;
    switch (variant) {
        case 0: epi_frag_loop<ACT_NONE, true, false, false>(c, p, pr); break;
        case 1: epi_frag_loop<ACT_NONE, true, false, true>(c, p, pr); break;
        case 2: epi_frag_loop<ACT_NONE, false, false, false>(c, p, pr); break;
        case 3: epi_frag_loop<ACT_NONE, false, false, true>(c, p, pr); break;
        case 4: epi_frag_loop<ACT_NONE, false, true, false>(c, p, pr); break;
        case 5: epi_frag_loop<ACT_NONE, false, true, true>(c, p, pr); break;
        case 6: epi_frag_loop<ACT_GELU, false, false, false>(c, p, pr); break;
        case 7: epi_frag_loop<ACT_GELU, false, false, true>(c, p, pr); break;
        case 8: epi_frag_loop<ACT_GELU, false, true, false>(c, p, pr); break;
        case 9: epi_frag_loop<ACT_GELU, false, true, true>(c, p, pr); break;
        case 10: epi_frag_loop<ACT_RELU, false, false, false>(c, p, pr); break;
        case 11: epi_frag_loop<ACT_RELU, false, false, true>(c, p, pr); break;
        case 12: epi_frag_loop<ACT_RELU, false, true, false>(c, p, pr); break;
        default: epi_frag_loop<ACT_RELU, false, true, true>(c, p, pr); break;
    }
}

// Barriers:  full[s] (leader) <- complete_tx of both CTAs' TMA loads;  empty[s] (both) <- tcgen05.commit multicast by the leader;
//            tfull[b] (both)  <- commit multicast after the last k-block of a tile;
//            tempty[b] (leader, 32 arrivals) <- one per epilogue warp of both CTAs once accumulator set b has been drained.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_h3_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmXr, const __grid_constant__ CUtensorMap tmXr1,
               const H3Params p, const H3Group grp, int w_pairs, int num_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + PIPE_BYTES);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float2* ln_smem = reinterpret_cast<float2*>(smem + PIPE_BYTES + 256);

    // warp index through a shuffle broadcast: tells the compiler it is warp-uniform (the role dispatch and every loop bound derived from it stay on
    // the uniform datapath, and the epilogue's shuffles compile without the WARPSYNC.COLLECTIVE fallback)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int tw = p.tw;
    const int xrows = tw >> 1;                                  // token rows staged by each CTA
    const int x_plane_bytes = xrows * 128;
    const int wrows = p.mhalf ? W_ROWS / 2 : W_ROWS;            // weight rows per CTA
    const int w_plane_bytes = wrows * 128, w_bytes = 2 * w_plane_bytes;
    const int stage_bytes = w_bytes + 2 * x_plane_bytes;
    const int nstages = p.nstages, nbuf = p.nbuf;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmW);
        prefetch_tmap(&tmX);
        if (grp.tiles0 < num_tiles) { prefetch_tmap(&tmW1); prefetch_tmap(&tmX1); }
        if (!p.conv) { prefetch_tmap(&tmXr); if (grp.tiles0 < num_tiles) prefetch_tmap(&tmXr1); }
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 2 * EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    // Programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on its outputs are needed (TMA loads of the
    // activations, residual reads) or overwritten.  All CTAs of this persistent grid are resident, so the next kernel may be scheduled as they retire.
    pdl_wait();
    pdl_launch_dependents();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t x_off = nbuf == 2 ? 128u : 256u;             // column distance hh -> x inside one accumulator set (mhalf: tw/2 <= 128 columns each)

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; whole warp walks the loop, one elected lane issues: see the MMA warp) =====================
        {
            int stage = 0; uint32_t phase = 0;
            for (int u = cluster_id; u < num_tiles; u += num_clusters) {
                const int g = u >= grp.tiles0;
                const int tl = g ? u - grp.tiles0 : u;
                const CUtensorMap* mw = g ? &tmW1 : &tmW;
                const int tcount = g ? p.ttiles1 : p.ttiles0;
                const int wi = p.order ? tl / tcount : tl % w_pairs;
                const int tt = p.order ? tl % tcount : tl / w_pairs;
                const int twt = (!p.conv && tt == tcount - 1) ? (g ? grp.tw_r[1] : grp.tw_r[0]) : tw;      // this tile's width (narrower last tile)
                const CUtensorMap* mx = twt != tw ? (g ? &tmXr1 : &tmXr) : (g ? &tmX1 : &tmX);
                const uint32_t tile_tx = 2u * (uint32_t)(w_bytes + twt * 128);
                const int n0 = wi * 2 * wrows + (int)rank * wrows;      // this CTA's weight rows
                int m0 = tt * tw + (int)rank * (twt >> 1), img = 0, h0 = 0, w0 = 0;       // this CTA's half of the token tile
                if (p.conv) {
                    img = tt / p.tiles_per_img;
                    const int rem = tt - img * p.tiles_per_img;
                    h0 = (rem / p.tiles_w) * (tw >> 4) + (int)rank * (tw >> 5);
                    w0 = (rem % p.tiles_w) * 16;
                }
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sW = smem + stage * stage_bytes;
                    uint8_t* sX = sW + w_bytes;
                    const uint32_t lead_full = mapa_to_cta(smem_u32(&full_bar[stage]), 0);
                    if (elect_one_sync()) {
                        if (p.dbg_mode == 2) {
                            if (leader) mbar_arrive(&full_bar[stage]);
                        } else {
                            if (leader) mbar_expect_tx(&full_bar[stage], tile_tx);
                            if (p.conv) {
                                const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
                                const int kh = tap / p.KW, kw = tap - kh * p.KW;
                                tma2_load_3d(mw, lead_full, sW, tap * p.Cin + cb * BKH, n0, 0);
                                tma2_load_5d(mx, lead_full, sX, cb * BKH, w0 + kw - p.pad_w, h0 + kh - p.pad_h, img, 0);
                            } else {
                                tma2_load_3d(mw, lead_full, sW, kb * BKH, n0, 0);
                                tma2_load_3d(mx, lead_full, sX, kb * BKH, m0, 0);
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        // The WHOLE warp walks the loop (warp-uniform control flow and operands: descriptors live in uniform registers); only the tcgen05
        // instructions themselves are issued by one elected lane.  Issuing from inside an `if (lane == 0)` region instead makes ptxas wrap every
        // MMA into an elect / R2UR-broadcast / branch loop (~20 dependent instructions, ~100 clocks per MMA: more than a 256 x 128 x 16 MMA takes).
        if (leader) {
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int u = cluster_id; u < num_tiles; u += num_clusters, ++it) {
                int twt = tw;                                            // narrower last token tile of a linear problem (H3Group::tw_r)
                if (!p.conv) {
                    const int g = u >= grp.tiles0;
                    const int tl = g ? u - grp.tiles0 : u;
                    const int tcount = g ? p.ttiles1 : p.ttiles0;
                    const int tt = p.order ? tl % tcount : tl / w_pairs;
                    if (tt == tcount - 1) twt = g ? grp.tw_r[1] : grp.tw_r[0];
                }
                const uint32_t idesc = make_idesc_f16(2 * wrows, twt);
                const uint32_t xl_off = (uint32_t)(w_bytes + (twt >> 1) * 128);      // lo plane of the token tile follows its hi plane
                const int buf = nbuf == 2 ? (it & 1) : 0;
                const uint32_t use = nbuf == 2 ? ((uint32_t)it >> 1) : (uint32_t)it;
                const bool ts_on = p.dbg_ts != nullptr && cluster_id == 0 && lane == 0;
                if (ts_on) p.dbg_ts[it * 8 + 0] = clock64();
                mbar_wait(&tempty_bar[buf], (use & 1u) ^ 1u);   // both CTAs' epilogues have drained this accumulator set
                tc_fence_after();
                if (ts_on) p.dbg_ts[it * 8 + 1] = clock64();
                const uint32_t acc_hh = tmem_base + (uint32_t)(buf * 256);
                const uint32_t acc_x = acc_hh + x_off;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sWh = smem_u32(smem + stage * stage_bytes);
                    const uint64_t dWh = make_smem_desc(sWh), dWl = make_smem_desc(sWh + (uint32_t)w_plane_bytes);
                    const uint64_t dXh = make_smem_desc(sWh + (uint32_t)w_bytes), dXl = make_smem_desc(sWh + xl_off);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < (p.dbg_mode == 1 ? 0 : 4); ++k) {
                            const uint64_t ko = (uint64_t)(k * 2);    // 16 fp16 = 32 bytes per k-step, in the descriptor's 16-byte units
                            const uint32_t acc = (kb | k) != 0;
                            umma2_f16(acc_x, dWl + ko, dXh + ko, idesc, acc);
                            umma2_f16(acc_x, dWh + ko, dXl + ko, idesc, 1u);
                            umma2_f16(acc_hh, dWh + ko, dXh + ko, idesc, acc);
                        }
                        umma2_commit_mc(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                if (elect_one_sync()) umma2_commit_mc(&tfull_bar[buf]);
                __syncwarp();
                if (ts_on) p.dbg_ts[it * 8 + 2] = clock64();
            }
        }
    } else {
        // ===================== epilogue (warps 2..17 of both CTAs) =====================
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int sub = (warp - 2) >> 2;                // which quarter of the tile's 16-column fragments (tw is a multiple of 16)
        const int nfrag = (p.mhalf ? tw >> 1 : tw) >> 4, per = (nfrag + 3) >> 2;   // 16-column fragments of this warp's lane quarter
        const int c_lo = min(nfrag, sub * per) * 16;
        const int c_hi = min(nfrag, (sub + 1) * per) * 16;
        const uint32_t lead_tempty0 = mapa_to_cta(smem_u32(&tempty_bar[0]), 0);
        int it = 0;
        for (int u = cluster_id; u < num_tiles; u += num_clusters, ++it) {
            const int buf = nbuf == 2 ? (it & 1) : 0;
            const uint32_t use = nbuf == 2 ? ((uint32_t)it >> 1) : (uint32_t)it;
            const int g = u >= grp.tiles0;
            const int tl = g ? u - grp.tiles0 : u;
            const int tcount = g ? p.ttiles1 : p.ttiles0;
            const int wi = p.order ? tl / tcount : tl % w_pairs;
            const int tt = p.order ? tl % tcount : tl / w_pairs;
            const int n_cta = wi * 2 * wrows + (int)rank * wrows;
            const int twt = (!p.conv && tt == tcount - 1) ? (g ? grp.tw_r[1] : grp.tw_r[0]) : tw;
            // (static member selection: a runtime index into the kernel-parameter struct would force a local copy of it)
            const H3Problem prob{g ? grp.prob[1].C : grp.prob[0].C, g ? grp.prob[1].Ch : grp.prob[0].Ch, g ? grp.prob[1].bias : grp.prob[0].bias,
                                 g ? grp.prob[1].residual : grp.prob[0].residual, g ? grp.prob[1].M : grp.prob[0].M,
                                 g ? grp.prob[1].vt : grp.prob[0].vt, g ? grp.prob[1].vt_cols : grp.prob[0].vt_cols,
                                 g ? grp.prob[1].stats_in : grp.prob[0].stats_in, g ? grp.prob[1].ln_s : grp.prob[0].ln_s,
                                 g ? grp.prob[1].stats_out : grp.prob[0].stats_out};
            const uint32_t acc_hh = tmem_base + (uint32_t)(buf * 256);
            const bool ts_on = p.dbg_ts != nullptr && cluster_id == 0 && leader && lane == 0 && (warp == 2 || warp == 17);
            unsigned long long* ts = ts_on ? p.dbg_ts + it * 8 + (warp == 2 ? 3 : 6) : nullptr;
            if (ts_on && warp == 2) ts[0] = clock64();
            run_epilogue(acc_hh, acc_hh + x_off, lane, q, c_lo, min(c_hi, p.mhalf ? twt >> 1 : twt), twt, n_cta, tt, p, prob, &tfull_bar[buf], use & 1u,
                         ln_smem + (it & 1) * 256,
                         (int)threadIdx.x - 64, (ts_on && warp == 2) ? ts + 1 : nullptr);
            if (ts_on) ts[warp == 2 ? 2 : 0] = clock64();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty0 + (uint32_t)(buf * 8));
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// generic producer of plane pairs: rows x cols fp32 (pitch ldx) -> (hi, lo) fp16 planes (pitch ldo, plane distance `plane`)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_h3_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols, __half* __restrict__ out,
                                                       int64_t ldo, int64_t plane, int vec, float lo_scale) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {   // 4 columns per thread (cols % 4 == 0, 16-byte aligned rows)
        const int c4 = cols >> 2;
        if (idx >= rows * c4) return;
        const int64_t r = idx / c4;
        const int c = (int)(idx - r * c4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
        uint2 hi, lo;
        h3_split2_s(v.x, v.y, lo_scale, hi.x, lo.x);
        h3_split2_s(v.z, v.w, lo_scale, hi.y, lo.y);
        __half* d = out + r * ldo + c;
        *reinterpret_cast<uint2*>(d) = hi;
        *reinterpret_cast<uint2*>(d + plane) = lo;
    } else {
        if (idx >= rows * cols) return;
        const int64_t r = idx / cols;
        const int c = (int)(idx - r * cols);
        __half hi, lo;
        h3_split_s(x[r * ldx + c], lo_scale, hi, lo);
        out[r * ldo + c] = hi;
        out[plane + r * ldo + c] = lo;
    }
}

// (hi, lo) planes -> fp32 (tests / debugging)
__global__ void __launch_bounds__(256) merge_h3_kernel(const __half* __restrict__ in, int64_t ldi, int64_t plane, int64_t rows, int cols,
                                                       float* __restrict__ y, int64_t ldy) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t r = idx / cols;
    const int c = (int)(idx - r * cols);
    y[r * ldy + c] = fmaf(__half2float(in[plane + r * ldi + c]), H3_LO_INV, __half2float(in[r * ldi + c]));
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
int g_dbg_mode = 0;
unsigned long long* g_dbg_ts = nullptr;
int g_order = 0;      // tuning aid: tile order (see H3Params::order)
int g_force_tw = 0;   // tuning aid (tools/gemm_sweep.py): > 0 = use this token tile width wherever it is legal
int g_no_rem = -1;     // tuning aid (SIU3R_H3_REM=0 in the environment, or the setter below): 1 = the last token tile of a problem is as wide as the others (siu3r_gemm_h3_set_remainder_tiles(0))
int g_cluster_cap = 0; // > 0: a launch uses at most this many CTA pairs (siu3r_gemm_h3_cluster_cap)
int g_mhalf = -1;     // M = 128 mode for N <= 128 (see H3Params::mhalf); SIU3R_H3_MHALF=0 turns it off (A/B measurements)
bool use_mhalf(int N) {
    if (g_mhalf < 0) { const char* e = getenv("SIU3R_H3_MHALF"); g_mhalf = (e && e[0] == '0') ? 0 : 1; }
    return g_mhalf && N <= 128;
}

// Token tile width for (M [+ M1]) tokens x N weight rows x K.  Cost model per CTA pair (clocks), fitted to tools/h3_bench.py twsweep / tiles:
//   mainloop of a tile = k-blocks x max(tensor time, operand bytes per CTA / its L2->SM share) + 700;  kind::f16 rate 8192 flop/clk/SM -> the 3 MMAs of a
//   k-step (256 x tw x 16) take 3 * tw/2 clocks, a k-block 6 * tw (3 * tw in the M = 128 mode); operands per CTA and k-block: 32 KB of weights +
//   tw/2 * 256 B of tokens at ~42 B/clk (chip-wide L2 cap of ~6300 B/clk);
//   epilogue of a tile E = ew x tw clocks, ew = per-token cost of the epilogue variant (17 plain ... 50 GELU + plane-pair split, measured with clock64
//   stamps);  double-buffered accumulators (tw <= 128 or M = 128 mode): E hides under the next mainloop and is exposed once; single-buffered
//   (tw > 128): every tile pays E plus the MMA-completion / barrier hand-off (~3700 clocks by the stamps; 7500 reproduces the measured sweep).  (The first version of this model charged 10 * tw for the
//   single buffer and put the encoder's fc1 on tw = 256: 49.9 us instead of 42.7.)
int pick_tw(int64_t M, int N, int K, int64_t M1, bool conv, int conv_h = 0, int conv_w = 0, int ew = 17) {
    const int w_pairs = ceil_div(N, 256);
    const int num_kb = ceil_div(K, BKH);
    const bool mh = use_mhalf(N);
    int best = 0; double best_t = 1e30;
    const int step = (conv || mh) ? 32 : 16;
    for (int tw = 32; tw <= 256; tw += step) {
        if (g_force_tw && g_force_tw % step == 0 && tw != g_force_tw) continue;
        int64_t T;
        if (conv) T = M * ceil_div(conv_h, tw / 16) * (conv_w / 16);
        else T = ceil_div_i64(M, tw) + (M1 > 0 ? ceil_div_i64(M1, tw) : 0);
        const int64_t tiles = (int64_t)w_pairs * T;
        const int64_t rounds = ceil_div_i64(tiles, CLUSTERS);
        const double kb = mh ? fmax(3.0 * tw, (16384.0 + 128.0 * tw) / 42.0) : fmax(6.0 * tw, (32768.0 + 128.0 * tw) / 42.0);
        const double mainloop = num_kb * kb + 700.0;
        const double E = (double)ew * tw * (mh ? 0.5 : 1.0);
        const bool dbuf = tw <= 128 || mh;
        const double t = dbuf ? (double)rounds * fmax(mainloop, E + 1000.0) + E + 2000.0 : (double)rounds * (mainloop + E + 7500.0);
        if (t < best_t * 0.999) { best_t = t; best = tw; }
    }
    return best;
}
// per-token epilogue cost (clocks per token of the tile width) of a launch's epilogue variant, for pick_tw
int epilogue_weight(int act, bool split, bool rope, bool residual, bool stats_out, bool dual, bool ln) {
    return 17 + (split ? 13 : 0) + (act == ACT_GELU ? 15 : act == ACT_RELU ? 2 : 0) + (rope ? 12 : 0) + (residual ? 8 : 0) + (stats_out ? 6 : 0) + (dual ? 6 : 0) +
           (ln ? 5 : 0);
}

int launch_h3(const CUtensorMap& w, const CUtensorMap& x, const CUtensorMap& w1, const CUtensorMap& x1, const CUtensorMap& xr, const CUtensorMap& xr1,
              H3Params& p, const H3Group& grp,
              int w_pairs, int num_tiles, cudaStream_t stream) {
    static int max_clusters[64] = {0};
    int dev = 0;
    SIU3R_CUDA_CHECK(cudaGetDevice(&dev));
    dev &= 63;
    if (max_clusters[dev] == 0) {   // per device: the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(gemm_h3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * CLUSTERS); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, gemm_h3_kernel, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 64; }
        max_clusters[dev] = n;
        if (getenv("SIU3R_GEMM_VERBOSE")) fprintf(stderr, "[siu3r_b200] gemm_h3: %d resident clusters, %d B smem\n", n, SMEM_BYTES);
    }
    p.mhalf = use_mhalf(p.N) ? 1 : 0;
    const int stage_bytes = (p.mhalf ? W_BYTES / 2 : W_BYTES) + p.tw * 128;
    p.nbuf = (p.tw <= 128 || p.mhalf) ? 2 : 1;
    p.nstages = PIPE_BYTES / stage_bytes > MAX_STAGES ? MAX_STAGES : PIPE_BYTES / stage_bytes;
    int clusters = num_tiles < max_clusters[dev] ? num_tiles : max_clusters[dev];
    if (g_cluster_cap > 0 && clusters > g_cluster_cap) clusters = g_cluster_cap;
    SIU3R_CUDA_CHECK(siu3r_launch_pdl(gemm_h3_kernel, dim3((unsigned)(2 * clusters)), dim3(THREADS), SMEM_BYTES, stream, w, x, w1, x1, xr, xr1, p, grp, w_pairs, num_tiles));
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int map_rows(CUtensorMap* m, const void* base, int64_t K, int64_t rows, int64_t ld, int64_t plane, int box_rows) {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)rows, 2};
    uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)plane * 2};
    uint32_t box[3] = {BKH, (uint32_t)box_rows, 2};
    return make_map_f16(m, base, 3, dims, str, box);
}

bool plane_ok(const void* base, int64_t ld, int64_t plane) { return base && ((uintptr_t)base & 15) == 0 && ld % 8 == 0 && plane % 8 == 0; }

}  // namespace

extern "C" {

// tuning aid: 0 = cost model, otherwise the token tile width to use (multiple of 16 / 32 for convs, <= 256)
void siu3r_gemm_h3_force(int tw) { g_force_tw = tw; }
// Scheduling aid for concurrent branches: launches issued while cap > 0 occupy at most `cap` of the 74 CTA pairs, so that a latency-bound chain of
// small kernels on a high-priority stream (the Mask2Former decoder next to the DPT heads) always finds free SMs; 0 = no cap (default).
void siu3r_gemm_h3_cluster_cap(int cap) { g_cluster_cap = cap > 0 ? cap : 0; }
// tuning aid: 1 = narrower last token tile per problem (default), 0 = uniform tiles
void siu3r_gemm_h3_set_remainder_tiles(int on) { g_no_rem = on ? 0 : 1; }
// tuning aid: 1 = M = 128 MMAs for N <= 128 (default), 0 = always M = 256
void siu3r_gemm_h3_set_mhalf(int on) { g_mhalf = on ? 1 : 0; }
void siu3r_gemm_h3_order(int order) { g_order = order ? 1 : 0; }
void siu3r_gemm_h3_debug_ts(void* dev_buf) { g_dbg_ts = (unsigned long long*)dev_buf; }   // tuning aid: see H3Params::dbg_ts (>= 8 * tiles-per-pair u64)
void siu3r_gemm_h3_debug(int mode) { g_dbg_mode = mode; }   // timing experiments: 1 = TMA only, 2 = MMA only (outputs are garbage)

// Host-only view of the tile planner: token tile width, number of 256 x tw tiles and rounds over the 74 resident CTA pairs.
int siu3r_gemm_h3_plan(int M, int N, int K, int M1, int* tw_out, int* tiles_out, int* rounds_out) {
    SIU3R_REQUIRE(M > 0 && N > 0 && K > 0 && M1 >= 0 && tw_out && tiles_out && rounds_out);
    const int tw = pick_tw(M, N, K, M1, false);
    *tw_out = tw;
    *tiles_out = ceil_div(N, 256) * (ceil_div(M, tw) + (M1 > 0 ? ceil_div(M1, tw) : 0));
    *rounds_out = ceil_div(*tiles_out, CLUSTERS);
    return SIU3R_OK;
}

// rows x cols fp32 (pitch ldx) -> fp16 plane pair (see h3.cuh) at `out` (pitch ldo, lo plane `plane` elements after the hi plane);
// unscaled_lo != 0: lo = fp16(x - hi) without the 2^11 factor (attention operands)
int siu3r_split_h3(const float* x, int64_t ldx, int64_t rows, int cols, void* out, int64_t ldo, int64_t plane, int unscaled_lo, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(x && out && rows > 0 && cols > 0 && ldx >= cols && ldo >= cols && plane > 0);
    const int vec = (cols % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && plane % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 7) == 0) ? 1 : 0;
    const int64_t n = vec ? rows * (cols / 4) : rows * cols;
    split_h3_kernel<<<(unsigned)ceil_div_i64(n, 256), 256, 0, stream>>>(x, ldx, rows, cols, (__half*)out, ldo, plane, vec, unscaled_lo ? 1.0f : H3_LO_SCALE);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// inverse of siu3r_split_h3 (tests, debugging): y = hi + lo * 2^-11
int siu3r_merge_h3(const void* in, int64_t ldi, int64_t plane, int64_t rows, int cols, float* y, int64_t ldy, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(in && y && rows > 0 && cols > 0);
    merge_h3_kernel<<<(unsigned)ceil_div_i64(rows * cols, 256), 256, 0, stream>>>((const __half*)in, ldi, plane, rows, cols, y, ldy);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// One or two linear layers of the same shape class in ONE persistent launch:
//   out_g = act(alpha * X_g W_g^T + bias_g) [RoPE-2D on columns < rope_cols] + residual_g,   g < ngroups (1 or 2)
// X_g: [M_g, K] plane pair (pitch lda, plane a_plane), W_g: [N, K] plane pair (pitch ldw, plane w_plane).  Exactly one of C_g (fp32, pitch ldc)
// and Ch_g (plane pair, pitch ldh, plane h_plane) receives the result -- the same kind for both groups.  vt_g != null: output columns
// >= vt_col0 go to the V^T plane pair vt_g[(n - vt_col0)][m] (pitch vt_ld, plane vt_plane, zero-filled up to vt_cols_g) instead.
// unscaled_lo != 0: the split outputs (Ch and V^T) carry lo = fp16(x - hi) without the 2^11 factor (operands of siu3r_flash_attn_h3).
// Pointer arrays are HOST arrays of device pointers.  This is how the two decoder streams of AsymmetricCroCo (dec_blocks / dec_blocks2:
// backbone_croco.py:244-250, :514-531) share the machine; ngroups = 1 is the plain nn.Linear replacement.
// Fused-LayerNorm extension of siu3r_gemm_h3 (same arguments, plus):
//   stats_out_g != null: the launch also accumulates the row statistics (sum * 2^32, sum of squares * 2^26, int64 pairs, ZEROED by the caller) of
//     the rows it writes -- the producer half of a LayerNorm fused into the GEMM that consumes these rows next;
//   stats_in_g != null (with ln_s_g): X_g holds the RAW rows, W_g = W * gamma, bias_g = W beta + b, ln_s_g[n] = sum_k gamma_k W[n,k]; the epilogue applies
//     rstd_m * (acc - mean_m * ln_s[n]) + bias[n] before activation / RoPE  ==  Linear(LayerNorm(x)) (croco/blocks.py:127-130,186-190).  alpha must be 1.
//   C_host and Ch_host may BOTH be given: the result goes out as fp32 (the residual stream) and as a plane pair (the next GEMM's operand).
int siu3r_gemm_h3_ln(int ngroups, const int* M_host, int N, int K, const void* const* X_host, int64_t lda, int64_t a_plane, const void* const* W_host,
                     int64_t ldw, int64_t w_plane, float* const* C_host, int64_t ldc, void* const* Ch_host, int64_t ldh, int64_t h_plane,
                     const float* const* bias_host, const float* const* residual_host, int64_t ldr, int act, float alpha, const int64_t* positions,
                     const float* rope_tab, int rope_cols, void* const* vt_host, const int* vt_cols_host, int64_t vt_ld, int64_t vt_plane, int vt_col0,
                     int unscaled_lo, const int64_t* const* stats_in_host, const float* const* ln_s_host, float ln_eps, int64_t* const* stats_out_host,
                     void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE((ngroups == 1 || ngroups == 2) && M_host && X_host && W_host && N > 0 && K > 0);
    SIU3R_REQUIRE(C_host != nullptr || Ch_host != nullptr);
    for (int g = 0; g < ngroups; ++g) {
        SIU3R_REQUIRE(M_host[g] > 0 && plane_ok(X_host[g], lda, a_plane) && plane_ok(W_host[g], ldw, w_plane));
        SIU3R_REQUIRE((!C_host || C_host[g] != nullptr) && (!Ch_host || (Ch_host[g] != nullptr && ((uintptr_t)Ch_host[g] & 1) == 0)));
        if (stats_in_host) SIU3R_REQUIRE(stats_in_host[g] && ((uintptr_t)stats_in_host[g] & 15) == 0 && ln_s_host && ln_s_host[g] && alpha == 1.0f);
        if (stats_out_host) SIU3R_REQUIRE(stats_out_host[g] && ((uintptr_t)stats_out_host[g] & 15) == 0);
    }
    SIU3R_REQUIRE(lda >= K && ldw >= K && (act & ~ACT_MASK) == 0);
    if (positions) SIU3R_REQUIRE(rope_tab && rope_cols > 0 && rope_cols % 64 == 0 && rope_cols <= N && ((uintptr_t)rope_tab & 15) == 0 && !residual_host);
    if (vt_host) {
        SIU3R_REQUIRE(vt_cols_host && vt_ld % 8 == 0 && vt_plane % 8 == 0 && vt_col0 % 64 == 0 && vt_col0 < N);
        SIU3R_REQUIRE(positions == nullptr || rope_cols <= vt_col0);
        for (int g = 0; g < ngroups; ++g)
            SIU3R_REQUIRE(vt_host[g] && ((uintptr_t)vt_host[g] & 15) == 0 && vt_cols_host[g] % 8 == 0 && vt_cols_host[g] >= M_host[g]);
    }
    const int M0 = M_host[0], M1 = ngroups == 2 ? M_host[1] : 0;
    const int tw = pick_tw(M0, N, K, M1, false, 0, 0, epilogue_weight(act, Ch_host != nullptr, positions != nullptr, residual_host != nullptr,
                                                                    stats_out_host != nullptr, C_host && Ch_host, stats_in_host != nullptr));
    SIU3R_REQUIRE(tw >= 32 && tw <= 256 && tw % (use_mhalf(N) ? 32 : 16) == 0);
    CUtensorMap mw[2], mx[2];
    for (int g = 0; g < ngroups; ++g) {
        int r = map_rows(&mw[g], W_host[g], K, N, ldw, w_plane, use_mhalf(N) ? W_ROWS / 2 : W_ROWS); if (r) return r;
        r = map_rows(&mx[g], X_host[g], K, M_host[g], lda, a_plane, tw / 2); if (r) return r;
    }
    if (ngroups == 1) { mw[1] = mw[0]; mx[1] = mx[0]; }
    H3Params p{};
    p.N = N; p.num_kb = ceil_div(K, BKH); p.tw = tw; p.act = act; p.alpha = alpha; p.ldc = ldc; p.ldh = ldh; p.plane_h = h_plane; p.ldr = ldr;
    p.rope_pos = (const long long*)positions; p.rope_tab = rope_tab; p.rope_cols = rope_cols;
    p.vt_col0 = vt_col0; p.vt_ld = vt_ld; p.vt_plane = vt_plane; p.conv = 0; p.lo_scale = unscaled_lo ? 1.0f : H3_LO_SCALE;
    p.ln_inv_c = 1.0f / (float)K; p.ln_eps = ln_eps;
    const int w_pairs = ceil_div(N, 256);
    H3Group grp{};
    for (int g = 0; g < 2; ++g) {
        const int s = g < ngroups ? g : 0;
        grp.prob[g] = H3Problem{C_host ? C_host[s] : nullptr, Ch_host ? (__half*)Ch_host[s] : nullptr, bias_host ? bias_host[s] : nullptr,
                                residual_host ? residual_host[s] : nullptr, M_host[s], vt_host ? (__half*)vt_host[s] : nullptr,
                                vt_cols_host ? vt_cols_host[s] : 0, stats_in_host ? (const long long*)stats_in_host[s] : nullptr,
                                ln_s_host ? ln_s_host[s] : nullptr, stats_out_host ? (long long*)stats_out_host[s] : nullptr};
    }
    grp.tiles0 = w_pairs * ceil_div(M0, tw);
    p.ttiles0 = ceil_div(M0, tw); p.ttiles1 = ngroups == 2 ? ceil_div(M1, tw) : 1; p.order = g_order; p.dbg_mode = g_dbg_mode; p.dbg_ts = g_dbg_ts;
    const int tiles = grp.tiles0 + (ngroups == 2 ? w_pairs * ceil_div(M1, tw) : 0);
    // narrower last token tile per problem (H3Group::tw_r): the remaining rows rounded up to the MMA-N granularity
    CUtensorMap mxr[2];
    const int gran = use_mhalf(N) ? 32 : 16;
    for (int g = 0; g < 2; ++g) {
        const int sg = g < ngroups ? g : 0;
        const int Mg = M_host[sg];
        const int rem = Mg - (ceil_div(Mg, tw) - 1) * tw;
        int twr = ceil_div(rem, gran) * gran;
        if (g_no_rem < 0) { const char* e = getenv("SIU3R_H3_REM"); g_no_rem = (e && e[0] == '0') ? 1 : 0; }
        if (twr > tw || g_no_rem) twr = tw;
        grp.tw_r[g] = twr;
        int r = map_rows(&mxr[g], X_host[sg], K, Mg, lda, a_plane, twr / 2); if (r) return r;
    }
    return launch_h3(mw[0], mx[0], mw[1], mx[1], mxr[0], mxr[1], p, grp, w_pairs, tiles, stream);
}

int siu3r_gemm_h3(int ngroups, const int* M_host, int N, int K, const void* const* X_host, int64_t lda, int64_t a_plane, const void* const* W_host,
                  int64_t ldw, int64_t w_plane, float* const* C_host, int64_t ldc, void* const* Ch_host, int64_t ldh, int64_t h_plane,
                  const float* const* bias_host, const float* const* residual_host, int64_t ldr, int act, float alpha, const int64_t* positions,
                  const float* rope_tab, int rope_cols, void* const* vt_host, const int* vt_cols_host, int64_t vt_ld, int64_t vt_plane, int vt_col0,
                  int unscaled_lo, void* stream_) {
    SIU3R_REQUIRE((C_host != nullptr) != (Ch_host != nullptr));
    return siu3r_gemm_h3_ln(ngroups, M_host, N, K, X_host, lda, a_plane, W_host, ldw, w_plane, C_host, ldc, Ch_host, ldh, h_plane, bias_host, residual_host,
                            ldr, act, alpha, positions, rope_tab, rope_cols, vt_host, vt_cols_host, vt_ld, vt_plane, vt_col0, unscaled_lo, nullptr, nullptr,
                            0.0f, nullptr, stream_);
}

// Stride-1 KH x KW convolution with "same"-size output, NHWC:  y[n,h,w,co] = act(sum x[n,h+kh-pad_h,w+kw-pad_w,ci] Wt[co,(kh,kw,ci)] + bias) + residual
// x: [Nimg,H,W,Cin] plane pair (pixel pitch ldx >= Cin, plane x_plane), Wt: [Cout, KH*KW*Cin] plane pair (pitch ldw), output fp32 y (pixel pitch
// ldc) or plane pair yh (pixel pitch ldh, plane h_plane); residual fp32 (pixel pitch ldr).  Requirements: W % 16 == 0, Cin % 8 == 0.
// Replaces nn.Conv2d(k, stride=1, padding=k//2) on the DPT / adapter / FPN paths (heads/dpt_block.py, vit_adapter/vit_adapter.py:200-262).
int siu3r_conv2d_h3(int Nimg, int H, int W, int Cin, int Cout, int KH, int KW, int pad_h, int pad_w, const void* x, int64_t ldx, int64_t x_plane,
                    const void* Wt, int64_t ldw, int64_t w_plane, float* y, int64_t ldc, void* yh, int64_t ldh, int64_t h_plane, const float* bias,
                    const float* residual, int64_t ldr, int act, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(Nimg > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && (y != nullptr) != (yh != nullptr));
    SIU3R_REQUIRE(plane_ok(x, ldx, x_plane) && plane_ok(Wt, ldw, w_plane) && ldx >= Cin && ldw >= (int64_t)KH * KW * Cin && (act & ~ACT_MASK) == 0);
    SIU3R_REQUIRE(2 * pad_h == KH - 1 && 2 * pad_w == KW - 1);
    if (W % 16 != 0 || Cin % 8 != 0) return SIU3R_ERR_UNSUPPORTED;
    const int cblocks = ceil_div(Cin, BKH);
    const int Keff = KH * KW * cblocks * BKH;     // k-blocks run over (tap, 64-channel block); a partial last block is zero-filled by TMA
    const int tw = pick_tw(Nimg, Cout, Keff, 0, true, H, W, epilogue_weight(act, yh != nullptr, false, residual != nullptr, false, false, false));
    SIU3R_REQUIRE(tw >= 32 && tw <= 256 && tw % 32 == 0);
    CUtensorMap mw, mx;
    int r = map_rows(&mw, Wt, (int64_t)KH * KW * Cin, Cout, ldw, w_plane, use_mhalf(Cout) ? W_ROWS / 2 : W_ROWS); if (r) return r;
    {
        uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg, 2};
        uint64_t str[4] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2, (uint64_t)x_plane * 2};
        uint32_t box[5] = {BKH, 16, (uint32_t)(tw / 32), 1, 2};
        r = make_map_f16(&mx, x, 5, dims, str, box); if (r) return r;
    }
    H3Params p{};
    p.N = Cout; p.num_kb = KH * KW * cblocks; p.tw = tw; p.act = act; p.alpha = 1.0f; p.ldc = ldc; p.ldh = ldh; p.plane_h = h_plane; p.ldr = ldr;
    p.lo_scale = H3_LO_SCALE; p.conv = 1; p.H = H; p.W = W; p.Cin = Cin; p.KW = KW; p.pad_h = pad_h; p.pad_w = pad_w; p.cblocks = cblocks;
    p.tiles_w = W / 16; p.tiles_per_img = p.tiles_w * ceil_div(H, tw / 16);
    const int w_pairs = ceil_div(Cout, 256);
    H3Group grp{};
    grp.prob[0] = H3Problem{y, (__half*)yh, bias, residual, Nimg, nullptr, 0};
    grp.prob[1] = grp.prob[0];
    grp.tiles0 = w_pairs * Nimg * p.tiles_per_img;
    p.ttiles0 = Nimg * p.tiles_per_img; p.ttiles1 = 1; p.order = g_order;
    grp.tw_r[0] = grp.tw_r[1] = tw;
    return launch_h3(mw, mx, mw, mx, mx, mx, p, grp, w_pairs, grp.tiles0, stream);
}

}  // extern "C"
