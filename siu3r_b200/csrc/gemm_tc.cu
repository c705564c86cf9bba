// tcgen05 (5th-gen tensor core) GEMM / implicit-GEMM convolution for sm_100a, fp32 storage, TF32 tensor-core math.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T )            (W is a torch nn.Linear weight: K-major, exactly the UMMA B operand)
//
// It is the workhorse behind every nn.Linear / 1x1 conv / 3x3 stride-1 conv on the hot path:
//   croco/blocks.py:97,110,74-77,154-156,167 (qkv / proj / fc1 / fc2 / projq,k,v), backbone_croco.py:87 (decoder_embed),
//   heads/dpt_block.py:98-116,181-189,358-364,385-391 (DPT 3x3 / 1x1 convs), vit_adapter/blocks.py:118-121,
//   mask2former/video_seg_decoder.py (all Linear / Conv2d 1x1 / 3x3).
//
// Structure (one 128 x BN output tile per CTA, 192 threads, warp-specialised):
//   warp 0   : TMA producer   - cp.async.bulk.tensor (2-D for a row-major A, 4-D NHWC box for the conv A operand: the
//              tile's 8x16 pixel patch shifted by the filter tap, halo zero-filled by TMA = conv padding for free),
//              128B-swizzled stages, mbarrier complete_tx.
//   warp 1   : MMA issuer     - one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) from shared-memory
//              descriptors, accumulator in TMEM; tcgen05.commit frees stages / signals the epilogue.
//   warps 2-5: epilogue       - tcgen05.ld 32x32b.x32 -> registers -> bias / GELU(erf) / ReLU / residual -> global.
//
// Precision modes:
//   NSPLIT = 1 : plain TF32 (what the reference itself runs on GPU: croco/croco.py:13 allow_tf32 = True).
//   NSPLIT = 3 : 3xTF32 split (A = A_hi + A_lo, W = W_hi + W_lo; hi*hi + lo*hi + hi*lo, fp32 accumulate) - ~fp32
//                accuracy on the tensor cores; used for the strict parity tolerances of BASELINE.json's north_star.
#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;          // fp32 elements per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 8;       // tf32
constexpr int NUM_THREADS = 192;
constexpr int CONV_TW = 16, CONV_TH = 8;  // spatial patch of one M tile in conv mode

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_MASK = 3, ACT_ROUND_TF32 = 4 /* flag: store RN_tf32(result) */ };

struct GemmParams {
    int M, N, num_kb;
    float* C; int64_t ldc;
    const float* bias;         // [N] or null
    const float* residual;     // [M, ldr] or null (added after activation)
    int64_t ldr;
    int act;
    float alpha;               // scales the accumulator before bias
    // conv mode
    int conv;                  // 0 = linear, 1 = conv (stride 1)
    int H, W, Cin, KH, KW, pad /* rows */, pad_w /* columns */;
    int tiles_w, tiles_h;      // tiles per image row / column
    // fused RoPE-2D on the output (q / k columns of a qkv projection; croco/blocks.py:101-103,158-160)
    const long long* rope_pos; // [M, 2] (y, x) positions or null
    const float* rope_tab;     // [maxpos][16] (cos, sin) pairs from siu3r_rope2d_table (head dim 64)
    int rope_cols;             // columns [0, rope_cols) are rotated (multiple of 64)
    long long* dbg;            // optional: per-CTA clock64 stamps [cta][8] (tools/gemm_probe.py), null in production
    // gemm_tc3: output columns >= vt_col0 (the V part of a qkv / k|v projection) are written TRANSPOSED into Vt[(n - vt_col0)][m] (row pitch
    // vt_ld) instead of C: the K-major V^T operand of the flash-attention P.V product, without a separate transpose pass
    int vt_col0; long long vt_ld;
    // gemm_tc3 "rows" mode: KH x 1 convolution over a flattened [H*W, 32] row-packed image (one k-block = one vertical tap):
    // k-block kb reads the X operand rows m + (kb - rows_pad) * rows_W at K offset 0 (TMA zero-fills above / below the image)
    int rows_W, rows_pad;
    // ... with the residual taken from a bilinear x2 (align_corners) upsampling of a [up_h, up_w, N] map computed in the epilogue
    int up_h, up_w;
    float up_sh, up_sw;
};

// ------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// one lane of the converged warp (issuing tcgen05 / TMA instructions from warp-uniform code keeps their operands on the uniform datapath; from
// inside an `if (lane == 0)` region ptxas wraps every such instruction into an elect / R2UR-broadcast / branch loop of ~20 dependent instructions)
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16 |
// SBO(1024 B>>4 = 64)<<32 | version(1)<<46 | layout SWIZZLE_128B(2)<<61
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::tf32: D fp32 (bit 4), A/B = TF32 (2 at bits 7 / 10), both K-major,
// N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------------------------
// Epilogue store of one 32-row x 32-column accumulator block held as "one row per lane" (tcgen05.ld 32x32b layout).
// Round-1 probe (tools/gemm_probe.py): writing rows straight from that layout (each lane a 16-byte piece of a different
// 128-byte line) ran at ~0.6 TB/s and dominated every GEMM.  The block is therefore transposed through a padded
// shared-memory tile (stride 33: conflict-free both ways) so that each store instruction writes one full 128-byte line;
// bias / activation / residual / rounding are applied in the transposed domain (bias = one value per lane, residual read
// coalesced).  No alignment requirement on C / residual / ldc remains.
// ------------------------------------------------------------------------------------------------------------
constexpr int EPI_LD = 36;                   // padded row stride (floats): 16-byte aligned rows, conflict-free v4 access both ways
constexpr int EPI_TR_FLOATS = 32 * EPI_LD;   // per epilogue warp

struct EpiRowMap {  // maps a tile row r (0..127) to its global row index (pixel index in conv mode) or -1 when masked
    int conv, m0, M, img, h0, w0, H, W, nimg;
    __device__ __forceinline__ int64_t row(int r) const {
        if (conv) {
            const int h = h0 + r / CONV_TW, w = w0 + r % CONV_TW;
            if (img >= nimg || h >= H || w >= W) return -1;
            return ((int64_t)img * H + h) * W + w;
        }
        const int m = m0 + r;
        return m < M ? (int64_t)m : -1;
    }
};

__device__ __forceinline__ void sts_v4(uint32_t saddr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t saddr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
    return r;
}
__device__ __forceinline__ float epi_act(float x, int act) {
    if (act == ACT_GELU) return gelu_erf(x);
    if (act == ACT_RELU) return fmaxf(x, 0.0f);
    return x;
}

__device__ __forceinline__ float4 epi_load_bias(const GemmParams& p, int col) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) {
        if (col < p.N) b.x = __ldg(p.bias + col);
        if (col + 1 < p.N) b.y = __ldg(p.bias + col + 1);
        if (col + 2 < p.N) b.z = __ldg(p.bias + col + 2);
        if (col + 3 < p.N) b.w = __ldg(p.bias + col + 3);
    }
    return b;
}

// Stage the warp's 32x32 block (row = lane) into its padded shared tile.
__device__ __forceinline__ void epi_stage(const uint32_t (&v)[32], uint32_t tr_saddr, int lane) {
#pragma unroll
    for (int j = 0; j < 32; j += 4)
        sts_v4(tr_saddr + (uint32_t)(lane * EPI_LD + j) * 4, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
               __uint_as_float(v[j + 3]));
}

// Read the staged block back row-wise (each instruction = 4 rows x 128 B), apply the epilogue, store full 128-byte lines.
__device__ __forceinline__ void epi_emit(uint32_t tr_saddr, int lane, int q, int nbase, float4 bias, const GemmParams& p, const EpiRowMap& rm) {
    const int act = p.act & ACT_MASK;
    const bool rnd = (p.act & ACT_ROUND_TF32) != 0;
    const int c4 = (lane & 7) * 4;     // this lane's 4 columns inside the block
    const int rsub = lane >> 3;        // row inside each group of 4 rows
    const int col = nbase + c4;
    const bool full4 = col + 4 <= p.N;
    const bool c_vec = ((p.ldc & 3) == 0) && ((((uintptr_t)p.C) & 15) == 0) && full4 && ((col & 3) == 0);
    const bool r_vec = p.residual && ((p.ldr & 3) == 0) && ((((uintptr_t)p.residual) & 15) == 0) && full4 && ((col & 3) == 0);
    // RoPE: a 32-column block is one (y or x) half of a 64-wide head; the pair (d, d+16) lives in lanes l and l^4 of the same row
    const bool rope = p.rope_pos != nullptr && nbase < p.rope_cols;   // warp-uniform
    const int axis = (nbase >> 5) & 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rsub + 4 * i;
        const int64_t grow = rm.row(q * 32 + r);
        float4 x = lds_v4(tr_saddr + (uint32_t)(r * EPI_LD + c4) * 4);
        x.x = epi_act(x.x * p.alpha + bias.x, act); x.y = epi_act(x.y * p.alpha + bias.y, act);
        x.z = epi_act(x.z * p.alpha + bias.z, act); x.w = epi_act(x.w * p.alpha + bias.w, act);
        if (rope) {
            float4 y;
            y.x = __shfl_xor_sync(0xffffffffu, x.x, 4); y.y = __shfl_xor_sync(0xffffffffu, x.y, 4);
            y.z = __shfl_xor_sync(0xffffffffu, x.z, 4); y.w = __shfl_xor_sync(0xffffffffu, x.w, 4);
            if (grow >= 0) {
                const long long pp = p.rope_pos[grow * 2 + axis];
                const float4* t = reinterpret_cast<const float4*>(p.rope_tab + (pp * 16 + (c4 & 15)) * 2);
                const float4 t0 = __ldg(t), t1 = __ldg(t + 1);   // (cos, sin) of d = c4..c4+3 (mod 16)
                if (c4 < 16) {   // first element of the pair: u*c - v*s
                    x.x = x.x * t0.x - y.x * t0.y; x.y = x.y * t0.z - y.y * t0.w;
                    x.z = x.z * t1.x - y.z * t1.y; x.w = x.w * t1.z - y.w * t1.w;
                } else {         // second: v*c + u*s
                    x.x = x.x * t0.x + y.x * t0.y; x.y = x.y * t0.z + y.y * t0.w;
                    x.z = x.z * t1.x + y.z * t1.y; x.w = x.w * t1.z + y.w * t1.w;
                }
            }
        }
        if (grow < 0 || col >= p.N) continue;
        if (p.residual) {
            const float* rp = p.residual + grow * p.ldr + col;
            if (r_vec) { const float4 rr = *reinterpret_cast<const float4*>(rp); x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w; }
            else { x.x += rp[0]; if (col + 1 < p.N) x.y += rp[1]; if (col + 2 < p.N) x.z += rp[2]; if (col + 3 < p.N) x.w += rp[3]; }
        }
        if (rnd) { x.x = rn_tf32(x.x); x.y = rn_tf32(x.y); x.z = rn_tf32(x.z); x.w = rn_tf32(x.w); }
        float* cp = p.C + grow * p.ldc + col;
        if (c_vec) *reinterpret_cast<float4*>(cp) = x;
        else { cp[0] = x.x; if (col + 1 < p.N) cp[1] = x.y; if (col + 2 < p.N) cp[2] = x.z; if (col + 3 < p.N) cp[3] = x.w; }
    }
}

// Whole-tile epilogue of one warp (its 32 TMEM lanes x NCOLS columns).  TF32 mode is software-pipelined: the TMEM load of
// chunk c+1 and the bias of chunk c+1 are in flight while chunk c is emitted.  3xTF32 sums the 4 accumulators per chunk.
template <int NCOLS, int NSPLIT>
__device__ __forceinline__ void run_epilogue(uint32_t tmem_base, uint32_t tr, int lane, int q, int n0, const GemmParams& p, const EpiRowMap& rm,
                                             uint64_t* acc_bar, uint32_t acc_parity = 0) {
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const int c4 = (lane & 7) * 4;
    float4 bias = epi_load_bias(p, n0 + c4);   // issued before the accumulator is ready: latency hidden behind the mainloop
    mbar_wait(acc_bar, acc_parity);
    tcgen05_fence_after();
    uint32_t v[32];
    if (NSPLIT == 1) tmem_ld_32x32b_x32(tq, v);
#pragma unroll 1
    for (int c0 = 0; c0 < NCOLS; c0 += 32) {
        const int nbase = n0 + c0;
        if (nbase >= p.N) break;
        const bool has_next = (c0 + 32 < NCOLS) && (nbase + 32 < p.N);
        const float4 bias_next = has_next ? epi_load_bias(p, nbase + 32 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (NSPLIT == 3) {
            tmem_ld_32x32b_x32(tq + (uint32_t)c0, v);
            tmem_ld_wait();
            const int nks = p.num_kb * (BK / UMMA_K);
#pragma unroll
            for (int a = 1; a < 4; ++a) {
                if (a < 3 && a >= nks) continue;  // hi*hi accumulator never written (K < 24)
                uint32_t t[32];
                tmem_ld_32x32b_x32(tq + (uint32_t)(a * NCOLS + c0), t);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t[j]));
            }
            epi_stage(v, tr, lane);
            __syncwarp();
        } else {
            tmem_ld_wait();
            epi_stage(v, tr, lane);
            __syncwarp();
            if (has_next) tmem_ld_32x32b_x32(tq + (uint32_t)(c0 + 32), v);  // overlaps with the stores below
        }
        epi_emit(tr, lane, q, nbase, bias, p, rm);
        __syncwarp();
        bias = bias_next;
    }
}

// DEEP: one CTA per SM with the whole shared memory as pipeline (TF32 mode).  Used when the grid has fewer CTAs than SMs anyway
// (the 3x3 convs of the DPT fusion stages at 16^2 .. 64^2: 8-64 CTAs x 72 k-blocks each): such launches are bound by the TMA round
// trip per stage, so twice the stages in flight is close to twice the speed.
template <int BN, int NSPLIT, bool DEEP = false>
struct Cfg {
    static constexpr int NOPER = NSPLIT == 1 ? 1 : 2;  // hi (+ lo) planes per operand
    static constexpr int A_BYTES = BM * BK * 4;
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = NOPER * (A_BYTES + B_BYTES);
    // TF32: <= 96 KB of stages so that TWO CTAs fit per SM (one CTA's epilogue overlaps the other's mainloop; 2 x 256 TMEM
    // columns).  3xTF32 needs all 512 TMEM columns -> one CTA per SM, deeper pipeline.
    static constexpr int BUDGET = (NSPLIT == 1 && !DEEP) ? 96 * 1024 : 208 * 1024;
    static constexpr int STAGES = BUDGET / STAGE_BYTES > 10 ? 10 : BUDGET / STAGE_BYTES;
    static constexpr int CTAS_PER_SM = (NSPLIT == 1 && !DEEP) ? 2 : 1;
    // the epilogue's transpose tiles (4 x 4.5 KB) alias the pipeline stages: the mainloop is over when the accumulator is ready
    static_assert(STAGES * STAGE_BYTES >= 4 * EPI_TR_FLOATS * 4, "stage memory must cover the epilogue tiles");
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    // 3xTF32: four accumulators (3 round-robin for hi*hi + 1 for the cross terms).  The tensor core truncates its fp32
    // accumulator on every MMA, so the error grows with the number of MMAs chained into ONE accumulator; spreading the
    // chain over several accumulators that are summed in registers (round-to-nearest) divides that error accordingly.
    static constexpr int NACC = NSPLIT == 1 ? 1 : 4;
    static constexpr int TMEM_COLS = NACC * BN < 32 ? 32 : NACC * BN;
};

template <int BN, int NSPLIT, bool DEEP = false>
__global__ void __launch_bounds__(NUM_THREADS, ((NSPLIT == 1 && !DEEP) ? 2 : 1))
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo, const GemmParams p) {
    using C_ = Cfg<BN, NSPLIT, DEEP>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C_::STAGES * C_::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C_::STAGES;
    uint64_t* acc_bar = empty_bar + C_::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (shuffle broadcast: provably warp-uniform)
    const int n0 = blockIdx.y * BN;
    long long* dbg = (p.dbg && blockIdx.y == 0 && blockIdx.x < 16) ? p.dbg + blockIdx.x * 8 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();
    // M-tile coordinates
    int m0 = blockIdx.x * BM, img = 0, h0 = 0, w0 = 0;
    if (p.conv) {
        const int per_img = p.tiles_w * p.tiles_h;
        img = blockIdx.x / per_img;
        const int t = blockIdx.x % per_img;
        h0 = (t / p.tiles_w) * CONV_TH;
        w0 = (t % p.tiles_w) * CONV_TW;
    }

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if (NSPLIT == 3) { prefetch_tmap(&tmAlo); prefetch_tmap(&tmBlo); }
        for (int s = 0; s < C_::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<C_::TMEM_COLS>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();

    if (warp == 0) {
        // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
        {
            const int cblocks = p.conv ? p.Cin / BK : 0;
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * C_::STAGE_BYTES;
                uint8_t* sB = sA + C_::NOPER * C_::A_BYTES;
                if (elect_one_sync()) {
                mbar_expect_tx(&full_bar[stage], C_::STAGE_BYTES);
                if (p.conv) {
                    const int tap = kb / cblocks, cb = kb - tap * cblocks;
                    const int kh = tap / p.KW, kw = tap - kh * p.KW;
                    tma_load_4d(&tmA, &full_bar[stage], sA, cb * BK, w0 + kw - p.pad_w, h0 + kh - p.pad, img);
                    if (NSPLIT == 3) tma_load_4d(&tmAlo, &full_bar[stage], sA + C_::A_BYTES, cb * BK, w0 + kw - p.pad_w, h0 + kh - p.pad, img);
                } else {
                    tma_load_2d(&tmA, &full_bar[stage], sA, kb * BK, m0);
                    if (NSPLIT == 3) tma_load_2d(&tmAlo, &full_bar[stage], sA + C_::A_BYTES, kb * BK, m0);
                }
                tma_load_2d(&tmB, &full_bar[stage], sB, kb * BK, n0);
                if (NSPLIT == 3) tma_load_2d(&tmBlo, &full_bar[stage], sB + C_::B_BYTES, kb * BK, n0);
                if (kb == 0 && dbg) dbg[2] = clock64();
                }
                __syncwarp();
                if (++stage == C_::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, one elected lane issues) =====================
        {
            constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t sA = smem_u32(smem + stage * C_::STAGE_BYTES);
                const uint32_t sB = sA + C_::NOPER * C_::A_BYTES;
                if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t koff = k * UMMA_K * 4;
                    const uint32_t first = (kb | k) != 0;
                    if (NSPLIT == 3) {
                        const int ks = kb * (BK / UMMA_K) + k;
                        const uint32_t acc_hh = tmem_base + (uint32_t)((ks % 3) * BN);   // hi*hi: round-robin over 3 accumulators
                        const uint32_t acc_x = tmem_base + (uint32_t)(3 * BN);           // lo*hi + hi*lo: 4th accumulator
                        umma_tf32(acc_x, make_smem_desc(sA + C_::A_BYTES + koff), make_smem_desc(sB + koff), idesc, first);
                        umma_tf32(acc_x, make_smem_desc(sA + koff), make_smem_desc(sB + C_::B_BYTES + koff), idesc, 1u);
                        umma_tf32(acc_hh, make_smem_desc(sA + koff), make_smem_desc(sB + koff), idesc, ks >= 3 ? 1u : 0u);
                    } else {
                        umma_tf32(tmem_base, make_smem_desc(sA + koff), make_smem_desc(sB + koff), idesc, first);
                    }
                }
                umma_commit(&empty_bar[stage]);  // stage reusable once these MMAs have read it
                }
                __syncwarp();
                if (++stage == C_::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(acc_bar);  // accumulator complete
            __syncwarp();
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;             // TMEM lane quarter this warp may access
        const uint32_t tr = smem_u32(smem) + (uint32_t)(q * EPI_TR_FLOATS * 4);   // aliases stage memory (free once acc_bar fires)
        EpiRowMap rm{p.conv, m0, p.M, img, h0, w0, p.H, p.W, p.conv ? (int)(p.M / (BM * p.tiles_w * p.tiles_h)) : 0};
        run_epilogue<BN, NSPLIT>(tmem_base, tr, lane, q, n0, p, rm, acc_bar);
        if (dbg && warp == 2 && lane == 0) { dbg[5] = 0; dbg[6] = clock64(); }
    }
    __syncwarp();  // lane 0 of the producer / MMA warps rejoins its warp before the CTA-wide barrier (bar.sync counts whole warps)
    tcgen05_fence_before();
    __syncthreads();
    if (dbg && threadIdx.x == 0) dbg[7] = clock64();
    if (warp == 1) tmem_dealloc<C_::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// 2-CTA variant (tcgen05 cta_group::2, cluster of 2 CTAs = one TPC): the pair computes a 256 x 256 tile with ONE
// tcgen05.mma (M=256, N=256, K=8) per k-step, issued by the leader CTA.  Each CTA stages its own 128 rows of A and only
// HALF of the B tile (128 of the 256 weight rows) -> 32 KB per k-block per SM instead of 48 KB for the same math, which is
// what matters with fp32 operands: round-1 ncu showed the 1-CTA kernel starved on operand ingest (tensor pipe 17 % active).
//   * TMA loads of both CTAs signal the LEADER's full barrier (mbarrier address mapped into the leader with mapa);
//   * the leader's tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs;
//   * TMEM (256 fp32 columns) is allocated with cta_group::2 by the same warp id in both CTAs.
// TF32 mode only (the 3xTF32 mode needs 4 accumulators = 1024 columns and stays on the 1-CTA kernel).
// ------------------------------------------------------------------------------------------------------------
constexpr int TC2_BN = 256;            // N extent of the pair tile; each CTA stages TC2_BN / 2 rows of W
constexpr int TC2_STAGE_BYTES = BM * BK * 4 + (TC2_BN / 2) * BK * 4;   // 32 KB
constexpr int TC2_STAGES = 3;   // 96 KB: two CTAs (of two different pairs) per SM
constexpr int TC2_SMEM_BYTES = TC2_STAGES * TC2_STAGE_BYTES + 1024 + 256;   // epilogue tiles alias the stages

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_to_cta(uint32_t local_saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {  // arrive on `bar` (same smem offset) in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 2)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TC2_STAGES * TC2_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + TC2_STAGES;
    uint64_t* acc_bar = empty_bar + TC2_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    constexpr int A_BYTES = BM * BK * 4;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (shuffle broadcast: provably warp-uniform)
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n0 = blockIdx.y * TC2_BN;
    int m0 = blockIdx.x * BM, img = 0, h0 = 0, w0 = 0;   // blockIdx.x = 2 * pair + rank: consecutive 128-row tiles
    if (p.conv) {
        const int per_img = p.tiles_w * p.tiles_h;
        img = blockIdx.x / per_img;
        const int t = blockIdx.x % per_img;
        h0 = (t / p.tiles_w) * CONV_TH;
        w0 = (t % p.tiles_w) * CONV_TW;
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC2_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of both CTAs are initialised before any remote complete_tx / multicast arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {
            const int cblocks = p.conv ? p.Cin / BK : 0;
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * TC2_STAGE_BYTES;
                uint8_t* sB = sA + A_BYTES;
                const uint32_t lead_full = mapa_to_cta(smem_u32(&full_bar[stage]), 0);
                if (elect_one_sync()) {
                if (leader) mbar_expect_tx(&full_bar[stage], 2 * TC2_STAGE_BYTES);  // bytes of both CTAs land on the leader's barrier
                if (p.conv) {
                    const int tap = kb / cblocks, cb = kb - tap * cblocks;
                    const int kh = tap / p.KW, kw = tap - kh * p.KW;
                    tma2_load_4d(&tmA, lead_full, sA, cb * BK, w0 + kw - p.pad_w, h0 + kh - p.pad, img);
                } else {
                    tma2_load_2d(&tmA, lead_full, sA, kb * BK, m0);
                }
                tma2_load_2d(&tmB, lead_full, sB, kb * BK, n0 + (int)rank * (TC2_BN / 2));
                }
                __syncwarp();
                if (++stage == TC2_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            constexpr uint32_t idesc = make_idesc_tf32(2 * BM, TC2_BN);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t sA = smem_u32(smem + stage * TC2_STAGE_BYTES);
                const uint32_t sB = sA + A_BYTES;
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 4;
                        umma2_tf32(tmem_base, make_smem_desc(sA + koff), make_smem_desc(sB + koff), idesc, (kb | k) != 0);
                    }
                    umma2_commit_mc(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == TC2_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma2_commit_mc(acc_bar);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const uint32_t tr = smem_u32(smem) + (uint32_t)(q * EPI_TR_FLOATS * 4);   // aliases stage memory
        EpiRowMap rm{p.conv, m0, p.M, img, h0, w0, p.H, p.W, p.M /* image count in conv mode */};
        run_epilogue<TC2_BN, 1>(tmem_base, tr, lane, q, n0, p, rm, acc_bar);
    }
    __syncwarp();
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer may still be reading TMEM / the leader's MMAs may still read this CTA's smem
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC2_BN) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// Persistent, operand-swapped 2-CTA kernel (linear mode, TF32):  C^T tile = W[256 rows] * X[tw tokens]^T.
//
// Why swapped.  The transformer GEMMs of one image pair have M = 2050 or 1025 token rows (1024 patches + the intrinsics
// token per image) against N = 768..4096 weight rows.  With tokens on the 128-row MMA-M axis the last M tile holds 1-2
// rows (6-20 % of the machine wasted) and the tile count (e.g. 9 x 12 pair tiles on 74 TPC pairs) quantises badly.  The
// weight rows are multiples of 256, so THEY go on the MMA-M axis (256 per CTA pair, nothing wasted) and the tokens on the
// MMA-N axis, whose tile width tw may be any multiple of 16 up to 256: the host picks tw so that (weight pair rows x
// token tiles) fills whole rounds of the 74 resident clusters (2050 tokens x 3072 rows: tw = 176 -> 144 tiles = 1.95 rounds).
// A second gain: TMEM lane = weight row n, TMEM column = token m, so one tcgen05.ld row-per-lane fragment already has 32
// consecutive n for a fixed m across the warp -> every global store / residual load is a full 128-byte line of C without
// the shared-memory transpose of the M-major kernels; bias is one register per thread, RoPE pairs (d, d+16) are lanes
// l and l^16.
//
// Persistent: one cluster (CTA pair) per TPC walks over its tiles; two TMEM accumulators (2 x 256 fp32 columns) let the
// epilogue of tile i run under the mainloop of tile i+1 and the 6-stage TMA ring stays full across tile boundaries.
//   full[s]   (leader)      <- complete_tx of both CTAs' TMA loads
//   empty[s]  (both CTAs)   <- tcgen05.commit multicast by the leader
//   tfull[b]  (both CTAs)   <- tcgen05.commit multicast after the last k-block of a tile
//   tempty[b] (leader, 16)  <- one arrive per epilogue warp of both CTAs once accumulator b has been drained
// ------------------------------------------------------------------------------------------------------------
struct Tc3 {
    static constexpr int W_BYTES = BM * BK * 4;             // 128 weight rows per CTA
    static constexpr int X_BYTES_MAX = 128 * BK * 4;        // up to 128 token rows per CTA (tw <= 256)
    static constexpr int STAGE_BYTES = W_BYTES + X_BYTES_MAX;
    static constexpr int STAGES = 6;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr int TMEM_COLS = 512;
    static constexpr int CLUSTERS = 74;                     // TPC pairs of a B200 (148 SMs)
};

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// Epilogue of one warp: TMEM lanes [32q, 32q+32) = weight rows n, columns [c_lo, c_hi) = its share of the tile's tokens.
// One 32-column fragment = 32 tokens x (this lane's n): every store / residual load below is one 128-byte line per warp.
constexpr int TC3_EPI_WARPS = 8;                       // two warps per TMEM lane quarter (they split the token columns)
constexpr int TC3_THREADS = 64 + 32 * TC3_EPI_WARPS;   // + TMA producer warp + MMA issuer warp

struct Tc3Problem {      // what differs between the problems of a grouped launch (same N, K, leading dimensions, epilogue)
    float* C;
    const float* bias;
    const float* residual;
    int M;
    float* vt;           // V^T destination of this problem's rows (null: everything goes to C); 16-byte aligned
    int vt_cols;         // columns of a V^T row this problem owns (>= M, multiple of 4): [M, vt_cols) is zero-filled
};
struct Tc3Group {
    Tc3Problem prob[2];
    int tiles0;          // tiles of problem 0 (tile t >= tiles0 belongs to problem 1)
    int nsplit;          // split-K factor (work unit u = split * num_tiles + tile; splits of a tile are serialised through `flags`)
    int kb_per_split;    // k-blocks per split
    int* flags;          // [num_tiles][2 * TC3_EPI_WARPS] progress words, zero between launches (null when nsplit == 1)
};

template <int ACTK, bool ROPE, bool RES, bool RND>
__device__ __forceinline__ void epi_chunk_swapped(const uint32_t (&v)[32], int jmax, int lane, int n, bool n_ok, float bias, int mrow, int axis,
                                                  const GemmParams& p, const Tc3Problem& pr, int64_t ldr) {
    float r[32];
    const float* rp = RES ? pr.residual + (int64_t)mrow * ldr + n : nullptr;
    if (RES) {
        // all loads in flight before the math; .cg (L2) because with split-K this is C as written by another SM a moment ago
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = (j < jmax && n_ok) ? __ldcg(rp + (int64_t)j * ldr) : 0.0f;
    }
    long long pos_l = 0;
    const float* tab = nullptr;
    if (ROPE) {
        pos_l = lane < jmax ? p.rope_pos[(int64_t)(mrow + lane) * 2 + axis] : 0;   // lane j holds the position of token j of the fragment
        tab = p.rope_tab + (lane & 15) * 2;
    }
    float* cp = pr.C + (int64_t)mrow * p.ldc + n;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(v[j]) * p.alpha + bias;
        if (ACTK == ACT_GELU) x = gelu_erf(x);
        if (ACTK == ACT_RELU) x = fmaxf(x, 0.0f);
        if (ROPE) {
            const float y = __shfl_xor_sync(0xffffffffu, x, 16);
            const int pj = __shfl_sync(0xffffffffu, (int)pos_l, j);
            const float2 cs = __ldg(reinterpret_cast<const float2*>(tab + pj * 32));
            x = (lane & 16) ? x * cs.x + y * cs.y : x * cs.x - y * cs.y;
        }
        if (RES) x += r[j];
        if (RND) x = rn_tf32(x);
        if (j < jmax && n_ok) cp[(int64_t)j * p.ldc] = x;
    }
}

// Epilogue fragment of the fused "7x7 image conv + ReLU + bilinear x2 upsampled trunk" step of the Gaussian-parameter head
// (heads/dpt_gs_head.py:162-164 after :155-160): x = act(acc + bias) + bilinear(low)[pixel m, channel n], align_corners = True, same source
// index arithmetic as siu3r_resize_bilinear_nhwc.  The 4 taps of a pixel are 128-byte lines across the warp (lane = channel).
// V columns of a fused qkv / k|v projection: this lane's output column n is one row of V^T, the fragment's 32 tokens are 128 contiguous
// bytes of it (8 x 16-byte stores; m and vt_ld are multiples of 4).  Tokens past M up to the padded pitch are zero-filled so that the
// attention kernel's TMA never stages uninitialised memory.  Values are stored round-to-nearest TF32 like siu3r_transpose_v does.
__device__ __forceinline__ void epi_chunk_swapped_vt(const uint32_t (&v)[32], int jmax, int n, bool n_ok, float bias, int mrow, const GemmParams& p,
                                                     const Tc3Problem& pr) {
    if (!n_ok) return;
    float* dst = pr.vt + (int64_t)(n - p.vt_col0) * p.vt_ld + mrow;
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = j < jmax ? rn_tf32(__uint_as_float(v[j]) * p.alpha + bias) : 0.0f;
    // a fragment cut by the END OF THE PROBLEM (mrow + jmax == M) also zero-fills the pad columns [M, vt_cols); one cut by the tile edge (tw is a
    // multiple of 16, not of 32) must stop there: the next columns belong to the neighbouring tile
    const int lim = (mrow + jmax == pr.M) ? min(32, pr.vt_cols - mrow) : jmax;
#pragma unroll
    for (int j = 0; j < 32; j += 4)
        if (j < lim) *reinterpret_cast<float4*>(dst + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
}

template <bool RND>
__device__ __forceinline__ void epi_chunk_swapped_up2x(const uint32_t (&v)[32], int jmax, int n, bool n_ok, float bias, int mrow, const GemmParams& p,
                                                       const Tc3Problem& pr) {
    const int W = 2 * p.up_w;
    const int act = p.act & ACT_MASK;
    int y = mrow / W, x = mrow - y * W;
    const float* __restrict__ low = pr.residual + n;
    float* __restrict__ cp = pr.C + (int64_t)mrow * p.ldc + n;
#pragma unroll
    for (int h = 0; h < 2; ++h) {            // 16 pixels at a time: all 64 tap loads are issued before any of them is consumed
        float t00[16], t01[16], t10[16], t11[16], wy[16], wx[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float sy = p.up_sh * (float)y, sx = p.up_sw * (float)x;
            int y0 = (int)sy, x0 = (int)sx;
            if (y0 > p.up_h - 1) y0 = p.up_h - 1;
            if (x0 > p.up_w - 1) x0 = p.up_w - 1;
            const int y1 = y0 + (y0 < p.up_h - 1 ? 1 : 0), x1 = x0 + (x0 < p.up_w - 1 ? 1 : 0);
            wy[j] = sy - (float)y0; wx[j] = sx - (float)x0;
            const bool ok = (h * 16 + j) < jmax && n_ok;
            const float* r0 = low + (int64_t)y0 * p.up_w * p.N;
            const float* r1 = low + (int64_t)y1 * p.up_w * p.N;
            t00[j] = ok ? __ldg(r0 + (int64_t)x0 * p.N) : 0.f; t01[j] = ok ? __ldg(r0 + (int64_t)x1 * p.N) : 0.f;
            t10[j] = ok ? __ldg(r1 + (int64_t)x0 * p.N) : 0.f; t11[j] = ok ? __ldg(r1 + (int64_t)x1 * p.N) : 0.f;
            if (++x == W) { x = 0; ++y; }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int jj = h * 16 + j;
            const float up = (1.f - wy[j]) * ((1.f - wx[j]) * t00[j] + wx[j] * t01[j]) + wy[j] * ((1.f - wx[j]) * t10[j] + wx[j] * t11[j]);
            float r = epi_act(__uint_as_float(v[jj]) * p.alpha + bias, act) + up;
            if (RND) r = rn_tf32(r);
            if (jj < jmax && n_ok) cp[(int64_t)jj * p.ldc] = r;
        }
    }
}

// split / nsplit / flag: split-K.  Split 0 writes C = alpha*acc + bias + residual, split s > 0 waits until the SAME warp slot of split s-1 has
// published its part of the tile (flag == s), then accumulates C += alpha*acc (the last split applies the TF32 rounding) and publishes s+1
// (the last split resets the word to 0 for the next launch).  One writer per flag word and phase: plain release / acquire, no atomics, and a
// fixed summation order -> bit-reproducible results.
template <bool UP2X>
__device__ __forceinline__ void run_epilogue_swapped(uint32_t tmem_acc, int lane, int q, int c_lo, int c_hi, int n_cta, int m_base,
                                                     const GemmParams& p, const Tc3Problem& pr_in, uint64_t* bar, uint32_t parity, int split,
                                                     int nsplit, int* flag) {
    const int nb = n_cta + q * 32;           // warp-uniform first weight row
    const int n = nb + lane;
    const bool n_ok = n < p.N;
    // split s > 0 accumulates into C: "residual" = C itself (pitch ldc), no bias
    const Tc3Problem pr{pr_in.C, split > 0 ? nullptr : pr_in.bias, split > 0 ? (const float*)pr_in.C : pr_in.residual, pr_in.M, pr_in.vt,
                        pr_in.vt_cols};
    const int64_t ldr = split > 0 ? p.ldc : p.ldr;
    const float bias = (pr.bias && n_ok) ? __ldg(pr.bias + n) : 0.0f;
    const int act = p.act & ACT_MASK;
    const bool rnd = (p.act & ACT_ROUND_TF32) != 0 && split == nsplit - 1;
    const bool rope = p.rope_pos != nullptr && nb < p.rope_cols;
    const bool res = pr.residual != nullptr;
    const int axis = (nb >> 5) & 1;
    // warp-uniform variant id: branches are hoisted out of the 32-element inner loops
    const int variant = rope ? (rnd ? 1 : 0) : 2 + (act * 4 + (res ? 2 : 0) + (rnd ? 1 : 0));
    mbar_wait(bar, parity);
    tcgen05_fence_after();
    if (split > 0) {   // the previous split's partial sums of this warp's block must be visible
        if (lane == 0) {
            int f;
            do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(flag) : "memory"); } while (f != split);
        }
        __syncwarp();
    }
    const uint32_t tq = tmem_acc + ((uint32_t)(q * 32) << 16);
    uint32_t v[32];
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_hi && nb < p.N; c0 += 32) {
        tmem_ld_32x32b_x32(tq + (uint32_t)c0, v);
        tmem_ld_wait();
        const int mrow = m_base + c0;
        int jmax = c_hi - c0;
        if (jmax > 32) jmax = 32;
        if (jmax > pr.M - mrow) jmax = pr.M - mrow;
        if (jmax <= 0) break;
        if (!UP2X && pr.vt != nullptr && nb >= p.vt_col0) {        // warp-uniform: a 32-column block never straddles vt_col0 (multiple of 64)
            epi_chunk_swapped_vt(v, jmax, n, n_ok, bias, mrow, p, pr);
            continue;
        }
        if (UP2X) {
            if (rnd) epi_chunk_swapped_up2x<true>(v, jmax, n, n_ok, bias, mrow, p, pr);
            else epi_chunk_swapped_up2x<false>(v, jmax, n, n_ok, bias, mrow, p, pr);
            continue;
        }
        switch (variant) {
            case 0: epi_chunk_swapped<ACT_NONE, true, false, false>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 1: epi_chunk_swapped<ACT_NONE, true, false, true>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 2: epi_chunk_swapped<ACT_NONE, false, false, false>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 3: epi_chunk_swapped<ACT_NONE, false, false, true>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 4: epi_chunk_swapped<ACT_NONE, false, true, false>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 5: epi_chunk_swapped<ACT_NONE, false, true, true>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 6: epi_chunk_swapped<ACT_GELU, false, false, false>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 7: epi_chunk_swapped<ACT_GELU, false, false, true>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 8: epi_chunk_swapped<ACT_GELU, false, true, false>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 9: epi_chunk_swapped<ACT_GELU, false, true, true>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 10: epi_chunk_swapped<ACT_RELU, false, false, false>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 11: epi_chunk_swapped<ACT_RELU, false, false, true>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            case 12: epi_chunk_swapped<ACT_RELU, false, true, false>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
            default: epi_chunk_swapped<ACT_RELU, false, true, true>(v, jmax, lane, n, n_ok, bias, mrow, axis, p, pr, ldr); break;
        }
    }
    if (nsplit > 1) {
        __syncwarp();
        if (lane == 0) {
            const int nf = split == nsplit - 1 ? 0 : split + 1;
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(nf) : "memory");
        }
    }
}

template <bool UP2X>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC3_THREADS, 1)   // 168 registers: 10 warps are allocated as 12
gemm_tc3_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmX1, const GemmParams p, const Tc3Group grp, int w_pairs, int num_tiles, int tw) {
    using C_ = Tc3;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C_::STAGES * C_::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C_::STAGES;
    uint64_t* tfull_bar = empty_bar + C_::STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (shuffle broadcast: provably warp-uniform)
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int num_units = num_tiles * grp.nsplit;               // split-major: every split-0 unit precedes the split-1 units
    const int xrows = tw >> 1;                                  // token rows staged by each CTA
    const uint32_t stage_tx = 2u * (uint32_t)(C_::W_BYTES + xrows * BK * 4);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmW);
        prefetch_tmap(&tmX);
        if (grp.tiles0 < num_tiles) { prefetch_tmap(&tmW1); prefetch_tmap(&tmX1); }
        for (int s = 0; s < C_::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 2 * TC3_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C_::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs; whole warp, one elected lane issues) =====================
        {
            int stage = 0; uint32_t phase = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters) {
                const int split = u / num_tiles, t = u - split * num_tiles;
                const int g = t >= grp.tiles0;
                const int tl = g ? t - grp.tiles0 : t;
                const CUtensorMap* mw = g ? &tmW1 : &tmW;
                const CUtensorMap* mx = g ? &tmX1 : &tmX;
                const int n0 = (tl % w_pairs) * 2 * BM + (int)rank * BM;       // this CTA's 128 weight rows
                const int m0 = (tl / w_pairs) * tw + (int)rank * xrows;        // this CTA's half of the token tile
                const int kb1 = min(p.num_kb, (split + 1) * grp.kb_per_split);
                for (int kb = split * grp.kb_per_split; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sW = smem + stage * C_::STAGE_BYTES;
                    uint8_t* sX = sW + C_::W_BYTES;
                    const uint32_t lead_full = mapa_to_cta(smem_u32(&full_bar[stage]), 0);
                    if (elect_one_sync()) {
                        if (leader) mbar_expect_tx(&full_bar[stage], stage_tx);
                        tma2_load_2d(mw, lead_full, sW, kb * BK, n0);
                        if (UP2X) tma2_load_2d(mx, lead_full, sX, 0, m0 + (kb - p.rows_pad) * p.rows_W);
                        else tma2_load_2d(mx, lead_full, sX, kb * BK, m0);
                    }
                    __syncwarp();
                    if (++stage == C_::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA; whole warp, one elected lane issues) =====================
        if (leader) {
            const uint32_t idesc = make_idesc_tf32(2 * BM, tw);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
                const int buf = it & 1;
                mbar_wait(&tempty_bar[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);   // both CTAs' epilogues have drained this accumulator
                tcgen05_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
                const int split = u / num_tiles;
                const int kb0 = split * grp.kb_per_split, kb1 = min(p.num_kb, kb0 + grp.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sW = smem_u32(smem + stage * C_::STAGE_BYTES);
                    const uint32_t sX = sW + C_::W_BYTES;
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint32_t koff = k * UMMA_K * 4;
                            umma2_tf32(acc, make_smem_desc(sW + koff), make_smem_desc(sX + koff), idesc, ((kb - kb0) | k) != 0);
                        }
                        umma2_commit_mc(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == C_::STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one_sync()) umma2_commit_mc(&tfull_bar[buf]);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 2..9 of both CTAs) =====================
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;               // which half of the tile's 32-column fragments
        const int nfrag = (tw + 31) >> 5;
        const int c_lo = half == 0 ? 0 : ((nfrag + 1) >> 1) * 32;
        const int c_hi = half == 0 ? min(tw, ((nfrag + 1) >> 1) * 32) : tw;
        const uint32_t lead_tempty0 = mapa_to_cta(smem_u32(&tempty_bar[0]), 0);
        int it = 0;
        for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
            const int buf = it & 1;
            const int split = u / num_tiles, t = u - split * num_tiles;
            const int g = t >= grp.tiles0;
            const int tl = g ? t - grp.tiles0 : t;
            const int n_cta = (tl % w_pairs) * 2 * BM + (int)rank * BM;
            const int m_base = (tl / w_pairs) * tw;
            int* flag = grp.flags ? grp.flags + ((size_t)t * 2 + rank) * TC3_EPI_WARPS + (warp - 2) : nullptr;
            // (static member selection: a runtime index into the kernel-parameter struct would force a local copy of it)
            const Tc3Problem prob{g ? grp.prob[1].C : grp.prob[0].C, g ? grp.prob[1].bias : grp.prob[0].bias,
                                  g ? grp.prob[1].residual : grp.prob[0].residual, g ? grp.prob[1].M : grp.prob[0].M,
                                  g ? grp.prob[1].vt : grp.prob[0].vt, g ? grp.prob[1].vt_cols : grp.prob[0].vt_cols};
            run_epilogue_swapped<UP2X>(tmem_base + (uint32_t)(buf * 256), lane, q, c_lo, c_hi, n_cta, m_base, p, prob, &tfull_bar[buf],
                                       ((uint32_t)it >> 1) & 1u, split, grp.nsplit, flag);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty0 + (uint32_t)(buf * 8));
        }
    }
    __syncwarp();
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C_::TMEM_COLS) : "memory");
}

template <bool UP2X = false>
int launch_tc3(const CUtensorMap& w, const CUtensorMap& x, const CUtensorMap& w1, const CUtensorMap& x1, const GemmParams& p, const Tc3Group& grp,
               int w_pairs, int num_tiles, int tw, cudaStream_t stream) {
    using C_ = Tc3;
    static int max_clusters_dev[64] = {0};
    int dev_ = 0;
    SIU3R_CUDA_CHECK(cudaGetDevice(&dev_));
    int& max_clusters = max_clusters_dev[dev_ & 63];
    if (max_clusters == 0) {
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc3_kernel<UP2X>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * C_::CLUSTERS); cfg.blockDim = dim3(TC3_THREADS); cfg.dynamicSmemBytes = C_::SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, gemm_tc3_kernel<UP2X>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 64; }
        max_clusters = n;
        if (getenv("SIU3R_GEMM_VERBOSE")) fprintf(stderr, "[siu3r_b200] gemm_tc3: %d resident clusters, %d stages, %d B smem\n", n, C_::STAGES, C_::SMEM_BYTES);
    }
    const int units = num_tiles * grp.nsplit;
    const int clusters = units < max_clusters ? units : max_clusters;
    gemm_tc3_kernel<UP2X><<<dim3((unsigned)(2 * clusters)), TC3_THREADS, C_::SMEM_BYTES, stream>>>(w, x, w1, x1, p, grp, w_pairs, num_tiles, tw);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

int launch_tc2(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, dim3 grid, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    if (siu3r_first_use_on_device(attr_set)) SIU3R_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES));
    gemm_tc2_kernel<<<grid, NUM_THREADS, TC2_SMEM_BYTES, stream>>>(a, b, p);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

bool tc2_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SIU3R_DISABLE_TC2"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

// ------------------------------------------------------------------------------------------------------------
// Host side: tensor-map construction (driver entry point resolved at run time: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    });
    return fn;
}

int make_map(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { fprintf(stderr, "[siu3r_b200] cuTensorMapEncodeTiled unavailable\n"); return SIU3R_ERR_CUDA; }
    cuuint64_t d[5]; cuuint64_t s[5]; cuuint32_t b[5]; cuuint32_t e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[siu3r_b200] cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu box %u %u)\n", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return SIU3R_ERR_INVALID;
    }
    return SIU3R_OK;
}

template <int BN, int NSPLIT, bool DEEP = false>
int launch(const CUtensorMap& a, const CUtensorMap& alo, const CUtensorMap& b, const CUtensorMap& blo, const GemmParams& p, dim3 grid,
           cudaStream_t stream) {
    using C_ = Cfg<BN, NSPLIT, DEEP>;
    static bool attr_set[64] = {false};
    if (siu3r_first_use_on_device(attr_set))
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, NSPLIT, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES));
    gemm_tc_kernel<BN, NSPLIT, DEEP><<<grid, NUM_THREADS, C_::SMEM_BYTES, stream>>>(a, alo, b, blo, p);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

long long* g_gemm_dbg = nullptr;

// Tile selection (TF32 mode).  148 SMs x 2 resident CTAs: prefer the 2-CTA 256x256 pair tile when it still fills the
// machine, otherwise fall back to smaller 1-CTA tiles so that small problems (M = 1025 / 2050 rows) launch enough CTAs.
constexpr int kFillCtas = 222;  // ~1.5 CTAs per SM
constexpr int kFlagSlotInts = 1 << 15, kFlagSlots = 32;   // split-K progress words: 32 slots x 128 KB

bool use_tc2(int N, int64_t mtiles) {
    if (!tc2_enabled() || N % TC2_BN != 0 || mtiles < 2) return false;
    return 2 * ceil_div_i64(mtiles, 2) * (N / TC2_BN) >= kFillCtas;
}

// forced kernel choice for tools/bench_ops.py sweeps: SIU3R_GEMM_FORCE = tc3 | tc2 | tc1 (unset = heuristic)
int g_force = -1;
int forced_kernel() {
    int& v = g_force;
    if (v < 0) {
        const char* e = getenv("SIU3R_GEMM_FORCE");
        v = 0;
        if (e) {
            if (!strcmp(e, "tc3")) v = 1; else if (!strcmp(e, "tc2")) v = 3; else if (!strcmp(e, "tc1")) v = 4;
        }
    }
    return v;
}

// Persistent swapped pair kernel: returns the token tile width tw (multiple of 16, <= 256) or 0 = use the one-tile kernels.
// Cost model per cluster (clocks): rounds x (k-blocks x max(tensor time, operand bytes per CTA / L2->SM share) + tile overhead);
// TF32 tensor rate 4096 flop/clk/SM -> a 128 x tw x 32 k-block takes 2*tw clocks; the chip-wide L2->SM cap (~6300 B/clk, measured
// on the 3x3 convs) gives each SM ~42 B/clk.
int pick_tc3(int M, int N, int K, int M1 = 0, bool allow_split = false, int* nsplit_out = nullptr) {
    if (nsplit_out) *nsplit_out = 1;
    const int f = forced_kernel();
    if (f == 3 || f == 4 || !tc2_enabled()) return 0;
    if (f == 0 && (N < 256 || K < 128 || M + M1 < 256)) return 0;
    const int tw_env = f >= 16 ? f : 0;   // siu3r_gemm_force(tw): this token tile width (sweeps)
    // Split-K is OFF unless SIU3R_TC3_SPLITK is set (1 = cost model decides, n >= 2 = force n where legal).  Its progress words live in a ring of
    // 32 slots handed out per launch; two forwards that overlap on the GPU (the two graph slots of forward_async) can be handed the same slot for
    // launches that run at the same time, which corrupts the words and leaves a warp spinning.  A second hazard is co-residency: split s spins
    // on split s-1 of its tile, which may belong to a CTA pair that is not resident yet when another persistent kernel (a parallel graph branch or
    // the other slot) holds part of the SMs and itself waits the same way.  Until the words are owned per graph instance and a waiter only ever
    // depends on work units with a lower index (dispatched first), the 1 % it gains on the K = 3072 / 4096 layers is not worth that risk.
    static int split_env = -1;
    if (split_env < 0) { const char* e = getenv("SIU3R_TC3_SPLITK"); const int v = e ? atoi(e) : 0; split_env = v <= 0 ? 1 : (v == 1 ? 0 : v + 1); }
    const int w_pairs = ceil_div(N, 256);
    const int num_kb = ceil_div(K, BK);
    int best = 0, best_s = 1; double best_t = 1e30;
    const int smax = (allow_split && nsplit_out && split_env != 1) ? 4 : 1;
    for (int S = 1; S <= smax; ++S) {
        if (split_env > 1 && smax > 1 && S != split_env - 1 && (split_env - 1) <= 4) continue;
        const int kbs = ceil_div(num_kb, S);
        if (S > 1 && (kbs < 8 || (S - 1) * kbs >= num_kb)) continue;      // every split needs work; short splits are all prologue
        for (int tw = 32; tw <= 256; tw += 16) {
            if (tw_env && tw != tw_env) continue;
            const int T = ceil_div(M, tw) + (M1 > 0 ? ceil_div(M1, tw) : 0);
            const int64_t tiles = (int64_t)w_pairs * T;
            if (S > 1 && tiles * 2 * TC3_EPI_WARPS > kFlagSlotInts) continue;
            const int64_t rounds = ceil_div_i64(tiles * S, Tc3::CLUSTERS);
            const double kb = fmax(2.0 * tw, (16384.0 + 64.0 * tw) / 42.0);
            // split-K: the splits' epilogues are serialised (flag hop + membar + C read-back), measured ~8 us on top of the model
            const double t = (double)rounds * (kbs * kb + 700.0 + 6.0 * tw) + (S > 1 ? 16000.0 : 0.0);
            if (t < best_t * 0.999) { best_t = t; best = tw; best_s = S; }
        }
    }
    if (nsplit_out) *nsplit_out = best_s;
    return best;
}

// Split-K progress flags: one device buffer for the life of the library, used as a ring of slots so that launches that may run
// concurrently (parallel graph branches) never share words; every kernel leaves its words at zero.
int* g_flag_ring = nullptr;
unsigned g_flag_next = 0;
int* next_flag_slot(cudaStream_t stream) {
    if (!g_flag_ring) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
        if (cudaMalloc(&g_flag_ring, sizeof(int) * kFlagSlotInts * kFlagSlots) != cudaSuccess) { cudaGetLastError(); g_flag_ring = nullptr; return nullptr; }
        cudaMemset(g_flag_ring, 0, sizeof(int) * kFlagSlotInts * kFlagSlots);
    }
    return g_flag_ring + (size_t)(g_flag_next++ % kFlagSlots) * kFlagSlotInts;
}

int pick_bn(int N, int64_t mtiles) {
    if (N <= 64) return 64;
    return mtiles * ceil_div(N, 128) >= 148 ? 128 : 64;   // BN = 64 when 128-wide tiles would leave SMs idle
}

}  // namespace

extern "C" {

// debugging aid (not part of the product ABI): device buffer [16][8] of clock64 stamps written by the 1-CTA linear kernel
void siu3r_gemm_debug_set(long long* dev_buf) { g_gemm_dbg = dev_buf; }
// tuning aid: 0 = heuristic, 1 = persistent swapped pair kernel wherever it is legal (>= 16: with that token tile width),
// 3 = one-tile pair kernel, 4 = 1-CTA kernels only
void siu3r_gemm_force(int kernel) { g_force = kernel; }
// Host-only view of the tile planner (no CUDA call): token tile width chosen for the persistent kernel (0 = the one-tile kernels run this shape),
// split-K factor, number of 256 x tw tiles and rounds over the 74 resident CTA pairs.  M1 > 0 = second problem of a grouped launch.
int siu3r_gemm_plan(int M, int N, int K, int M1, int allow_split, int* tw_out, int* nsplit_out, int* tiles_out, int* rounds_out) {
    SIU3R_REQUIRE(M > 0 && N > 0 && K > 0 && M1 >= 0 && tw_out && nsplit_out && tiles_out && rounds_out);
    int ns = 1;
    const int tw = pick_tc3(M, N, K, M1, allow_split != 0, &ns);
    *tw_out = tw; *nsplit_out = ns;
    const int tiles = tw ? ceil_div(N, 256) * (ceil_div(M, tw) + (M1 > 0 ? ceil_div(M1, tw) : 0)) : 0;
    *tiles_out = tiles;
    *rounds_out = tw ? ceil_div(tiles * ns, Tc3::CLUSTERS) : 0;
    return SIU3R_OK;
}

// C[M,N] (ldc) = act(alpha * A[M,K] (lda) @ W[N,K]^T (ldw) + bias[N]) + residual[M,N] (ldr)
// fp32 storage; precision 1 = TF32, 3 = 3xTF32 (needs the *_lo planes: x = hi + lo with hi = tf32-rounded x).
// Requirements: K % 4 == 0 handled by zero-filled TMA only if lda/ldw are multiples of 4 floats and all bases 16-byte aligned.
// act: 0 none, 1 GELU(erf), 2 ReLU; +4 = store the result rounded to nearest TF32 (output only feeds TF32 GEMMs).  Replaces torch.nn.functional.linear (+ fused bias/activation/residual).
static int gemm_tc_impl(int M, int N, int K, const float* A, const float* A_lo, int64_t lda, const float* Wt, const float* W_lo, int64_t ldw,
                        float* C, int64_t ldc, const float* bias, const float* residual, int64_t ldr, int act, float alpha, int precision,
                        const int64_t* rope_pos, const float* rope_tab, int rope_cols, void* stream_, float* vt = nullptr, int64_t vt_ld = 0,
                        int vt_col0 = 0) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rope_pos) SIU3R_REQUIRE(rope_tab && rope_cols > 0 && rope_cols % 64 == 0 && rope_cols <= N && ((uintptr_t)rope_tab & 15) == 0);
    SIU3R_REQUIRE(M > 0 && N > 0 && K > 0 && A && Wt && C);
    SIU3R_REQUIRE(precision == 1 || precision == 3);
    SIU3R_REQUIRE(precision == 1 || (A_lo && W_lo));
    SIU3R_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && lda >= K && ldw >= K);
    SIU3R_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)Wt & 15) == 0);
    const int64_t mtiles = ceil_div_i64(M, BM);
    int nsplit = 1;
    const bool can_split = (act & ACT_MASK) == ACT_NONE && rope_pos == nullptr;
    int tw = precision == 1 ? pick_tc3(M, N, K, 0, can_split, &nsplit) : 0;
    int* fl = nullptr;
    if (tw && nsplit > 1) {
        fl = next_flag_slot(stream);
        if (!fl) { nsplit = 1; tw = pick_tc3(M, N, K); }     // no flag buffer yet and the stream is capturing: plain schedule
    }
    if (vt && (!tw || nsplit > 1)) return SIU3R_ERR_UNSUPPORTED;   // V^T emission exists in the persistent kernel's epilogue only
    if (tw) {
        // persistent swapped pair kernel: weights on the MMA-M axis, tokens on the MMA-N axis
        CUtensorMap mw, mx;
        uint64_t dimsW[2] = {(uint64_t)K, (uint64_t)N}; uint64_t strW[1] = {(uint64_t)ldw * 4}; uint32_t boxW[2] = {BK, BM};
        int r = make_map(&mw, Wt, 2, dimsW, strW, boxW); if (r) return r;
        uint64_t dimsX[2] = {(uint64_t)K, (uint64_t)M}; uint64_t strX[1] = {(uint64_t)lda * 4}; uint32_t boxX[2] = {BK, (uint32_t)(tw / 2)};
        r = make_map(&mx, A, 2, dimsX, strX, boxX); if (r) return r;
        GemmParams p{};
        p.M = M; p.N = N; p.num_kb = ceil_div(K, BK); p.C = C; p.ldc = ldc; p.bias = bias; p.residual = residual; p.ldr = ldr;
        p.act = act; p.alpha = alpha; p.conv = 0;
        p.rope_pos = (const long long*)rope_pos; p.rope_tab = rope_tab; p.rope_cols = rope_cols;
        p.vt_col0 = vt_col0; p.vt_ld = vt_ld;
        const int w_pairs = ceil_div(N, 256);
        Tc3Group grp{};
        grp.prob[0] = Tc3Problem{C, bias, residual, M, vt, (int)vt_ld};
        grp.prob[1] = grp.prob[0];
        grp.tiles0 = w_pairs * ceil_div(M, tw);
        grp.nsplit = 1; grp.kb_per_split = p.num_kb; grp.flags = nullptr;
        if (nsplit > 1) { grp.nsplit = nsplit; grp.kb_per_split = ceil_div(p.num_kb, nsplit); grp.flags = fl; }
        return launch_tc3(mw, mx, mw, mx, p, grp, w_pairs, grp.tiles0, tw, stream);
    }
    if (precision == 1 && forced_kernel() != 4 && use_tc2(N, mtiles)) {
        // 2-CTA path: pairs of consecutive 128-row tiles share one 256 x 256 MMA
        CUtensorMap ma, mb;
        uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M}; uint64_t strA[1] = {(uint64_t)lda * 4}; uint32_t boxA[2] = {BK, BM};
        int r = make_map(&ma, A, 2, dimsA, strA, boxA); if (r) return r;
        uint64_t dimsB[2] = {(uint64_t)K, (uint64_t)N}; uint64_t strB[1] = {(uint64_t)ldw * 4}; uint32_t boxB[2] = {BK, TC2_BN / 2};
        r = make_map(&mb, Wt, 2, dimsB, strB, boxB); if (r) return r;
        GemmParams p{};
        p.M = M; p.N = N; p.num_kb = ceil_div(K, BK); p.C = C; p.ldc = ldc; p.bias = bias; p.residual = residual; p.ldr = ldr;
        p.act = act; p.alpha = alpha; p.conv = 0;
        p.rope_pos = (const long long*)rope_pos; p.rope_tab = rope_tab; p.rope_cols = rope_cols;
        dim3 grid((unsigned)(2 * ceil_div_i64(mtiles, 2)), (unsigned)(N / TC2_BN));
        return launch_tc2(ma, mb, p, grid, stream);
    }
    int bn = pick_bn(N, mtiles);
    CUtensorMap ma, malo, mb, mblo;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}; uint64_t str[1] = {(uint64_t)lda * 4}; uint32_t box[2] = {BK, BM};
        int r = make_map(&ma, A, 2, dims, str, box); if (r) return r;
        malo = ma;
        if (precision == 3) { r = make_map(&malo, A_lo, 2, dims, str, box); if (r) return r; }
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}; uint64_t str[1] = {(uint64_t)ldw * 4}; uint32_t box[2] = {BK, (uint32_t)bn};
        int r = make_map(&mb, Wt, 2, dims, str, box); if (r) return r;
        mblo = mb;
        if (precision == 3) { r = make_map(&mblo, W_lo, 2, dims, str, box); if (r) return r; }
    }
    GemmParams p{};
    p.M = M; p.N = N; p.num_kb = ceil_div(K, BK); p.C = C; p.ldc = ldc; p.bias = bias; p.residual = residual; p.ldr = ldr;
    p.act = act; p.alpha = alpha; p.conv = 0; p.dbg = g_gemm_dbg;
    p.rope_pos = (const long long*)rope_pos; p.rope_tab = rope_tab; p.rope_cols = rope_cols;
    dim3 grid((unsigned)mtiles, (unsigned)ceil_div(N, bn));
    if (precision == 1) {
        const bool deep = (int64_t)grid.x * grid.y <= 148 && p.num_kb >= 16;   // under-filled, long K loop: latency-bound
        if (bn == 256) return launch<256, 1>(ma, malo, mb, mblo, p, grid, stream);
        if (bn == 128) return deep ? launch<128, 1, true>(ma, malo, mb, mblo, p, grid, stream) : launch<128, 1>(ma, malo, mb, mblo, p, grid, stream);
        return deep ? launch<64, 1, true>(ma, malo, mb, mblo, p, grid, stream) : launch<64, 1>(ma, malo, mb, mblo, p, grid, stream);
    }
    if (bn == 128) return launch<128, 3>(ma, malo, mb, mblo, p, grid, stream);
    return launch<64, 3>(ma, malo, mb, mblo, p, grid, stream);
}

int siu3r_gemm_tc(int M, int N, int K, const float* A, const float* A_lo, int64_t lda, const float* Wt, const float* W_lo, int64_t ldw,
                  float* C, int64_t ldc, const float* bias, const float* residual, int64_t ldr, int act, float alpha, int precision,
                  void* stream) {
    return gemm_tc_impl(M, N, K, A, A_lo, lda, Wt, W_lo, ldw, C, ldc, bias, residual, ldr, act, alpha, precision, nullptr, nullptr, 0, stream);
}

// Two independent linear layers of the same shape class in ONE persistent launch (TF32):  C_g = act(alpha * A_g W_g^T + bias_g) + R_g,
// g = 0, 1, with M_g rows each and common N, K, leading dimensions and epilogue.  This is how the two decoder streams of
// AsymmetricCroCo (dec_blocks on view 1, dec_blocks2 on view 2: backbone_croco.py:244-250, :514-531) share the machine: each
// stream alone has M = 1025 rows, i.e. too few tiles to fill 148 SMs.  Pointer arrays are HOST arrays of device pointers.
// Returns SIU3R_ERR_UNSUPPORTED when the shape is not eligible for the persistent kernel (caller then issues two siu3r_gemm_tc).
int siu3r_gemm_tc_group2(const int* M_host, int N, int K, const float* const* A_host, int64_t lda, const float* const* W_host, int64_t ldw,
                         float* const* C_host, int64_t ldc, const float* const* bias_host, const float* const* residual_host, int64_t ldr,
                         int act, float alpha, const int64_t* positions, const float* rope_tab, int rope_cols, float* const* vt_host,
                         const int* vt_cols_host, int64_t vt_ld, int vt_col0, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(M_host && A_host && W_host && C_host && N > 0 && K > 0 && M_host[0] > 0 && M_host[1] > 0);
    SIU3R_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && lda >= K && ldw >= K);
    for (int g = 0; g < 2; ++g) SIU3R_REQUIRE(A_host[g] && W_host[g] && C_host[g] && ((uintptr_t)A_host[g] & 15) == 0 && ((uintptr_t)W_host[g] & 15) == 0);
    if (positions) SIU3R_REQUIRE(rope_tab && rope_cols > 0 && rope_cols % 64 == 0 && rope_cols <= N && ((uintptr_t)rope_tab & 15) == 0 && !residual_host);
    int nsplit = 1;
    const bool can_split = (act & ACT_MASK) == ACT_NONE && positions == nullptr;
    int tw = pick_tc3(M_host[0], N, K, M_host[1], can_split, &nsplit);
    if (tw == 0) return SIU3R_ERR_UNSUPPORTED;
    if (vt_host) {
        SIU3R_REQUIRE(vt_cols_host && vt_ld % 4 == 0 && vt_col0 % 64 == 0 && vt_col0 < N && (((uintptr_t)vt_host[0] | (uintptr_t)vt_host[1]) & 15) == 0);
        SIU3R_REQUIRE(vt_cols_host[0] % 4 == 0 && vt_cols_host[1] % 4 == 0 && vt_cols_host[0] >= M_host[0] && vt_cols_host[1] >= M_host[1]);
        if (nsplit > 1) { nsplit = 1; tw = pick_tc3(M_host[0], N, K, M_host[1]); }
    }
    int* fl = nsplit > 1 ? next_flag_slot(stream) : nullptr;
    if (nsplit > 1 && !fl) { nsplit = 1; tw = pick_tc3(M_host[0], N, K, M_host[1]); }
    CUtensorMap mw[2], mx[2];
    for (int g = 0; g < 2; ++g) {
        uint64_t dimsW[2] = {(uint64_t)K, (uint64_t)N}; uint64_t strW[1] = {(uint64_t)ldw * 4}; uint32_t boxW[2] = {BK, BM};
        int r = make_map(&mw[g], W_host[g], 2, dimsW, strW, boxW); if (r) return r;
        uint64_t dimsX[2] = {(uint64_t)K, (uint64_t)M_host[g]}; uint64_t strX[1] = {(uint64_t)lda * 4}; uint32_t boxX[2] = {BK, (uint32_t)(tw / 2)};
        r = make_map(&mx[g], A_host[g], 2, dimsX, strX, boxX); if (r) return r;
    }
    GemmParams p{};
    p.M = M_host[0]; p.N = N; p.num_kb = ceil_div(K, BK); p.C = C_host[0]; p.ldc = ldc; p.ldr = ldr; p.act = act; p.alpha = alpha; p.conv = 0;
    p.rope_pos = (const long long*)positions; p.rope_tab = rope_tab; p.rope_cols = rope_cols;
    p.vt_col0 = vt_col0; p.vt_ld = vt_ld;
    const int w_pairs = ceil_div(N, 256);
    Tc3Group grp{};
    for (int g = 0; g < 2; ++g)
        grp.prob[g] = Tc3Problem{C_host[g], bias_host ? bias_host[g] : nullptr, residual_host ? residual_host[g] : nullptr, M_host[g],
                                 vt_host ? vt_host[g] : nullptr, vt_cols_host ? vt_cols_host[g] : 0};
    grp.tiles0 = w_pairs * ceil_div(M_host[0], tw);
    grp.nsplit = nsplit; grp.kb_per_split = ceil_div(p.num_kb, nsplit); grp.flags = fl;
    const int tiles = grp.tiles0 + w_pairs * ceil_div(M_host[1], tw);
    return launch_tc3(mw[0], mx[0], mw[1], mx[1], p, grp, w_pairs, tiles, tw, stream);
}

// KH x 1 convolution over a row-packed image + activation + bilinear x2 (align_corners) upsampled residual, one image per call:
//   out[(y, x), co] = act( sum_kh rows[(y + kh - pad, x), :] . Wt[co, kh*32 : kh*32+32] + bias[co] ) + up2x(low)[(y, x), co]
// rows [H*W, 32] is siu3r_im2col_nhwc(KH = 1, KW) of the image (the KW horizontal taps of <= 32/KW channels packed per pixel), Wt
// [Cout, KH*32] the matching row-packed filter, low [H/2, W/2, Cout].  This is the input_merger step of the Gaussian-parameter head
// (heads/dpt_gs_head.py:113-119,155-164: F.interpolate(path_1, x2, bilinear, align_corners=True) + ReLU(Conv7x7(image))) in ONE
// persistent tensor-core launch: the [H, W, Cout] upsampled map is never materialised.  TF32 only; -4 when the shape is not eligible.
int siu3r_conv_rows_up2x_tc(int H, int W, int KH, int pad, int Cout, const float* rows, const float* Wt, const float* bias, const float* low,
                            float* out, int64_t ldc, int act, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(H > 0 && W > 0 && KH > 0 && pad >= 0 && Cout > 0 && rows && Wt && low && out && H % 2 == 0 && W % 2 == 0);
    SIU3R_REQUIRE(((uintptr_t)rows & 15) == 0 && ((uintptr_t)Wt & 15) == 0);
    const int M = H * W, K = KH * BK;
    const int tw = pick_tc3(M, Cout, K);
    if (tw == 0 || Cout % 32 != 0) return SIU3R_ERR_UNSUPPORTED;
    CUtensorMap mw, mx;
    uint64_t dimsW[2] = {(uint64_t)K, (uint64_t)Cout}; uint64_t strW[1] = {(uint64_t)K * 4}; uint32_t boxW[2] = {BK, BM};
    int r = make_map(&mw, Wt, 2, dimsW, strW, boxW); if (r) return r;
    uint64_t dimsX[2] = {(uint64_t)BK, (uint64_t)M}; uint64_t strX[1] = {(uint64_t)BK * 4}; uint32_t boxX[2] = {BK, (uint32_t)(tw / 2)};
    r = make_map(&mx, rows, 2, dimsX, strX, boxX); if (r) return r;
    GemmParams p{};
    p.M = M; p.N = Cout; p.num_kb = KH; p.C = out; p.ldc = ldc; p.bias = bias; p.act = act; p.alpha = 1.0f; p.conv = 0;
    p.rows_W = W; p.rows_pad = pad; p.up_h = H / 2; p.up_w = W / 2;
    p.up_sh = H > 1 ? (float)(H / 2 - 1) / (float)(H - 1) : 0.f;
    p.up_sw = W > 1 ? (float)(W / 2 - 1) / (float)(W - 1) : 0.f;
    const int w_pairs = ceil_div(Cout, 256);
    Tc3Group grp{};
    grp.prob[0] = Tc3Problem{out, bias, low, M, nullptr, 0};
    grp.prob[1] = grp.prob[0];
    grp.tiles0 = w_pairs * ceil_div(M, tw);
    grp.nsplit = 1; grp.kb_per_split = p.num_kb; grp.flags = nullptr;
    return launch_tc3<true>(mw, mx, mw, mx, p, grp, w_pairs, grp.tiles0, tw, stream);
}

// nn.Linear followed by RoPE-2D on output columns [0, rope_cols) (head dim 64), i.e. the qkv / q / kv projections of
// croco/blocks.py:97-103,154-160 with curope.rope_2d folded into the epilogue.  positions [M, 2] int64 (row m = token m),
// rope_tab from siu3r_rope2d_table.  Rotation happens after the bias and before the optional TF32 rounding.
int siu3r_gemm_tc_rope(int M, int N, int K, const float* A, const float* A_lo, int64_t lda, const float* Wt, const float* W_lo, int64_t ldw,
                       float* C, int64_t ldc, const float* bias, int act, int precision, const int64_t* positions, const float* rope_tab,
                       int rope_cols, void* stream) {
    SIU3R_REQUIRE(positions && rope_tab);
    return gemm_tc_impl(M, N, K, A, A_lo, lda, Wt, W_lo, ldw, C, ldc, bias, nullptr, 0, act, 1.0f, precision, positions, rope_tab, rope_cols, stream);
}

// siu3r_gemm_tc_rope whose output columns >= vt_col0 (the V third of a qkv projection) are written as V^T [(n - vt_col0)][m] (pitch vt_ld,
// multiple of 4, >= M; columns [M, vt_ld) zero-filled) instead of C -- the operand layout siu3r_flash_attn_tc consumes with
// vt_batch_cols = tokens per image.  TF32 only; -4 when the shape is not run by the persistent kernel (use siu3r_transpose_v then).
int siu3r_gemm_tc_rope_vt(int M, int N, int K, const float* A, int64_t lda, const float* Wt, int64_t ldw, float* C, int64_t ldc, const float* bias,
                          int act, const int64_t* positions, const float* rope_tab, int rope_cols, float* vt, int64_t vt_ld, int vt_col0,
                          void* stream) {
    SIU3R_REQUIRE(vt && vt_ld % 4 == 0 && vt_ld >= M && vt_col0 % 64 == 0 && vt_col0 < N && ((uintptr_t)vt & 15) == 0);
    SIU3R_REQUIRE(positions == nullptr || rope_cols <= vt_col0);
    return gemm_tc_impl(M, N, K, A, nullptr, lda, Wt, nullptr, ldw, C, ldc, bias, nullptr, 0, act, 1.0f, 1, positions, rope_tab, rope_cols, stream,
                        vt, vt_ld, vt_col0);
}

// Stride-1 KHxKW convolution, NHWC fp32:  y[n,h,w,co] = act(sum x[n,h+kh-pad,w+kw-pad,ci] * Wt[co,kh,kw,ci] + bias) + residual
// x [Nimg,H,W,Cin], Wt [Cout, KH*KW*Cin] (repacked from torch's [Cout,Cin,KH,KW]), y [Nimg,H,W,ldc>=Cout].
// Requirements: Cin % 32 == 0, W % 16 == 0, H % 8 == 0.  Replaces nn.Conv2d(k, stride=1, padding=pad) on the DPT / FPN paths.
int siu3r_conv2d_tc(int Nimg, int H, int W, int Cin, int Cout, int KH, int KW, int pad_h, int pad_w, const float* x, const float* x_lo,
                    const float* Wt, const float* W_lo, float* y, int64_t ldc, const float* bias, const float* residual, int64_t ldr,
                    int act, int precision, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(Nimg > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && x && Wt && y);
    SIU3R_REQUIRE(precision == 1 || precision == 3);
    SIU3R_REQUIRE(precision == 1 || (x_lo && W_lo));
    if (Cin % BK != 0 || W % CONV_TW != 0 || H % CONV_TH != 0) return SIU3R_ERR_UNSUPPORTED;
    SIU3R_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)Wt & 15) == 0);
    const int tiles_w = W / CONV_TW, tiles_h = H / CONV_TH;
    const int64_t mtiles = (int64_t)Nimg * tiles_w * tiles_h;
    const int Ktot = KH * KW * Cin;
    if (precision == 1 && use_tc2(Cout, mtiles)) {
        CUtensorMap ma, mb;
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t str[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
        uint32_t box[4] = {BK, CONV_TW, CONV_TH, 1};
        int r = make_map(&ma, x, 4, dims, str, box); if (r) return r;
        uint64_t dimsB[2] = {(uint64_t)Ktot, (uint64_t)Cout}; uint64_t strB[1] = {(uint64_t)Ktot * 4}; uint32_t boxB[2] = {BK, TC2_BN / 2};
        r = make_map(&mb, Wt, 2, dimsB, strB, boxB); if (r) return r;
        GemmParams p{};
        p.M = Nimg; /* image count: rows of the padded last pair are masked with it */ p.N = Cout; p.num_kb = KH * KW * (Cin / BK); p.C = y;
        p.ldc = ldc; p.bias = bias; p.residual = residual; p.ldr = ldr; p.act = act; p.alpha = 1.0f; p.conv = 1; p.H = H; p.W = W; p.Cin = Cin;
        p.KH = KH; p.KW = KW; p.pad = pad_h; p.pad_w = pad_w; p.tiles_w = tiles_w; p.tiles_h = tiles_h;
        dim3 grid((unsigned)(2 * ceil_div_i64(mtiles, 2)), (unsigned)(Cout / TC2_BN));
        return launch_tc2(ma, mb, p, grid, stream);
    }
    int bn = pick_bn(Cout, mtiles);
    CUtensorMap ma, malo, mb, mblo;
    {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t str[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
        uint32_t box[4] = {BK, CONV_TW, CONV_TH, 1};
        int r = make_map(&ma, x, 4, dims, str, box); if (r) return r;
        malo = ma;
        if (precision == 3) { r = make_map(&malo, x_lo, 4, dims, str, box); if (r) return r; }
    }
    {
        uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)Cout}; uint64_t str[1] = {(uint64_t)Ktot * 4}; uint32_t box[2] = {BK, (uint32_t)bn};
        int r = make_map(&mb, Wt, 2, dims, str, box); if (r) return r;
        mblo = mb;
        if (precision == 3) { r = make_map(&mblo, W_lo, 2, dims, str, box); if (r) return r; }
    }
    GemmParams p{};
    p.M = (int)(mtiles * BM); p.N = Cout; p.num_kb = KH * KW * (Cin / BK); p.C = y; p.ldc = ldc; p.bias = bias; p.residual = residual;
    p.ldr = ldr; p.act = act; p.alpha = 1.0f; p.conv = 1; p.H = H; p.W = W; p.Cin = Cin; p.KH = KH; p.KW = KW; p.pad = pad_h; p.pad_w = pad_w;
    p.tiles_w = tiles_w; p.tiles_h = tiles_h;
    dim3 grid((unsigned)mtiles, (unsigned)ceil_div(Cout, bn));
    if (precision == 1) {
        const bool deep = (int64_t)grid.x * grid.y <= 148 && p.num_kb >= 16;   // under-filled, long K loop: latency-bound
        if (bn == 256) return launch<256, 1>(ma, malo, mb, mblo, p, grid, stream);
        if (bn == 128) return deep ? launch<128, 1, true>(ma, malo, mb, mblo, p, grid, stream) : launch<128, 1>(ma, malo, mb, mblo, p, grid, stream);
        return deep ? launch<64, 1, true>(ma, malo, mb, mblo, p, grid, stream) : launch<64, 1>(ma, malo, mb, mblo, p, grid, stream);
    }
    if (bn == 128) return launch<128, 3>(ma, malo, mb, mblo, p, grid, stream);
    return launch<64, 3>(ma, malo, mb, mblo, p, grid, stream);
}

}  // extern "C"
