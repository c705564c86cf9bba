// Flash attention on the 5th-gen tensor cores (tcgen05 + TMEM), head dim 64, fp32 storage, TF32 math.
//
//   O[b, n, h*64 + d] = softmax_k( Q K^T * scale ) V        replaces croco/blocks.py:105-109 (Attention) and :162-166
//                                                            (CrossAttention) without materialising the N x N matrix.
//
// One CTA = 128 queries of one (batch, head); keys are visited in tiles of 64, ONE pass (online softmax with lazy rescaling).
//   warp 0   : TMA producer   Q tile once; per key tile K [64 keys x 64 d] and V^T [64 d x 64 keys] (K-major, 128B-swizzled), double-buffered
//   warp 1   : MMA issuer     S = Q K^T   : tcgen05.mma kind::tf32 M=128 N=64 K=8 x 8    -> TMEM S/P buffer t % 3 (64 columns)
//                             O += P V    : M=128 N=64 K=8 x 8, A = P read straight from TMEM -> TMEM columns [192,256)
//   warps 2-5: softmax        one query row per thread (tcgen05.ld 32x32b gives exactly that: no shuffles); P = exp2(S*scale*log2e - m_ref)
//                             written back over S in place with tcgen05.st (no shared-memory round trip); finally O / l -> global
//                             through a shared-memory transpose.
// 96 KB of shared memory + 256 TMEM columns per CTA -> two CTAs per SM.  See the comment above flash_tc_kernel for the rescaling rule.
// V must be supplied transposed ([b*H + h][d][key], row pitch vt_ld) so that every UMMA operand is K-major:
// siu3r_transpose_v produces it (and rounds it to TF32) from the fused qkv / kv buffer.
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace {

constexpr int FT_BM = 128, FT_BN = 64, FT_D = 64, FT_BK = 32;
constexpr int FT_THREADS = 192;
constexpr int FT_Q_BYTES = 2 * FT_BM * FT_BK * 4;        // 32 KB : 2 k-blocks [128 q x 32 d]
constexpr int FT_K_BYTES = 2 * FT_BN * FT_BK * 4;        // 16 KB per buffer : 2 k-blocks [64 keys x 32 d], double-buffered
constexpr int FT_V_BYTES = (FT_BN / FT_BK) * FT_D * FT_BK * 4;   // 16 KB per buffer : 2 k-blocks [64 d x 32 keys], double-buffered
constexpr int FT_SMEM = FT_Q_BYTES + 2 * FT_K_BYTES + 2 * FT_V_BYTES + 1024 + 256;   // 97.3 KB -> two CTAs per SM
constexpr int FT_TMEM_COLS = 256;                        // S/P 0..2: [0,64) [64,128) [128,192)   O: [192,256)
static_assert(2 * FT_K_BYTES >= 4 * 32 * 36 * 4, "the epilogue transpose tiles alias the K buffers");

struct FlashParams {
    int B, H, Nq, Nk;
    int q_col0, k_col0;          // column of head 0 inside the Q / K row
    float scale_log2e;
    float* O; int64_t o_bs, o_ts;
    int round_out;
    long long vt_batch_cols;     // 0: V^T rows indexed by (b, h, d); > 0: rows (h, d), image b at column offset b * vt_batch_cols
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {   // one lane of the converged warp (see gemm_tc.cu)
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t saddr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
    return r;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (P, one query row per lane, one key per 32-bit column) stays in tensor memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// barrier indices (double-buffered ones: index + buffer)
// K / V tiles are double-buffered, the S/P accumulator is TRIPLE-buffered: P(t) aliases S(t), so with two buffers S(t+1) could only be
// issued after P(t-1) V(t-1) had completed, i.e. the softmax warps waited one full MMA round trip per tile (ncu: 15 % of all stall samples on the
// S-ready wait).  With three, S(t+1) only needs P(t-2) V(t-2), which finished a whole tile earlier.
enum { B_Q = 0, B_KFULL = 1, B_KEMPTY = 3, B_SFULL = 5, B_PFULL = 8, B_VFULL = 11, B_VFREE = 13, B_SFREE = 15, B_OFULL = 18, B_COUNT };

// Single pass, online softmax with LAZY rescaling: the running reference maximum m_ref (log2 domain) of a row is only raised when the
// tile maximum exceeds it by more than 8, so P = exp2(s - m_ref) <= 256 stays exact in fp32 / TF32 and the O accumulator in TMEM
// is read-modify-written only when some row of the warp actually moves its reference (first tile aside, almost never).
// softmax is shift invariant, so the result is the exact softmax(QK^T)V -- O and l carry the same factor.
//   S(t) = Q K_t^T        -> TMEM buffer t&1 (64 fp32 columns)
//   P(t) overwrites S(t) in place (thread = query row = TMEM lane), RN-TF32, and is the A operand of O += P(t) V_t straight from TMEM
//   K / V^T tiles of 64 keys by TMA, both double-buffered; 96 KB of shared memory and 256 TMEM columns -> two CTAs per SM, so one
//   CTA's exponentials (MUFU-bound: 64 per row and tile) overlap the other's MMAs.
__global__ void __launch_bounds__(FT_THREADS, 2)
flash_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmVt,
                const FlashParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + FT_Q_BYTES;            // 2 buffers
    uint8_t* sV = sK + 2 * FT_K_BYTES;        // 2 buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * FT_V_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (shuffle broadcast: provably warp-uniform)
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * FT_BM;
    // TMA needs a 16-byte aligned start in the innermost (key) dimension of V^T.  When image b's keys start at an unaligned column
    // (vt_batch_cols = tokens per image, e.g. 1025) the key tiling is shifted down by kshift = start & 3 keys: tile t covers keys
    // [t*64 - kshift, ...), the (at most 3) phantom keys in front of key 0 are masked like the padding behind key Nk-1.
    const int kshift = p.vt_batch_cols ? (int)(((long long)b * p.vt_batch_cols) & 3) : 0;
    const int ntiles = (p.Nk + kshift + FT_BN - 1) / FT_BN;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmVt) : "memory");
        mbar_init(&bars[B_Q], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[B_KFULL + i], 1); mbar_init(&bars[B_KEMPTY + i], 1);
            mbar_init(&bars[B_VFULL + i], 1); mbar_init(&bars[B_VFREE + i], 1);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(&bars[B_SFULL + i], 1); mbar_init(&bars[B_PFULL + i], 4);     // one arrival per softmax warp
            mbar_init(&bars[B_SFREE + i], 1);
        }
        mbar_init(&bars[B_OFULL], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(FT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_O = tmem_base + 3 * FT_BN;

    if (warp == 0) {
        // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
        if (elect_one_sync()) {
            mbar_expect_tx(&bars[B_Q], FT_Q_BYTES);
            tma_load_3d(&tmQ, &bars[B_Q], sQ, p.q_col0 + h * FT_D, q0, b);
            tma_load_3d(&tmQ, &bars[B_Q], sQ + FT_BM * FT_BK * 4, p.q_col0 + h * FT_D + FT_BK, q0, b);
        }
        __syncwarp();
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            const uint32_t use = (uint32_t)t >> 1;
            mbar_wait(&bars[B_KEMPTY + buf], (use & 1) ^ 1);
            uint8_t* dK = sK + buf * FT_K_BYTES;
            if (elect_one_sync()) {
                mbar_expect_tx(&bars[B_KFULL + buf], FT_K_BYTES);
                tma_load_3d(&tmK, &bars[B_KFULL + buf], dK, p.k_col0 + h * FT_D, t * FT_BN - kshift, b);
                tma_load_3d(&tmK, &bars[B_KFULL + buf], dK + FT_BN * FT_BK * 4, p.k_col0 + h * FT_D + FT_BK, t * FT_BN - kshift, b);
            }
            __syncwarp();
            mbar_wait(&bars[B_VFREE + buf], (use & 1) ^ 1);      // P V of tile t-2 has read this V buffer
            uint8_t* dV = sV + buf * FT_V_BYTES;
            if (elect_one_sync()) {
                mbar_expect_tx(&bars[B_VFULL + buf], FT_V_BYTES);
#pragma unroll
                for (int j = 0; j < FT_BN / FT_BK; ++j)
                    tma_load_2d(&tmVt, &bars[B_VFULL + buf], dV + j * FT_D * FT_BK * 4,
                                (int)(p.vt_batch_cols ? b * p.vt_batch_cols : 0) + t * FT_BN - kshift + j * FT_BK,
                                (p.vt_batch_cols ? h : b * p.H + h) * FT_D);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: S(t+1) is issued before P(t) V(t); whole warp, one elected lane issues =====================
        constexpr uint32_t idesc_s = make_idesc_tf32(FT_BM, FT_BN);
        constexpr uint32_t idesc_o = make_idesc_tf32(FT_BM, FT_D);
        mbar_wait(&bars[B_Q], 0);
        auto issue_s = [&](int t) {
            const int buf = t & 1, sb = t % 3;
            const uint32_t use = (uint32_t)t >> 1;
            mbar_wait(&bars[B_KFULL + buf], use & 1);
            mbar_wait(&bars[B_SFREE + sb], (((uint32_t)t / 3) & 1) ^ 1);   // P(t-3) (aliasing this S buffer) has been consumed
            tc_fence_after();
            const uint32_t kb = smem_u32(sK + buf * FT_K_BYTES);
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < FT_D / 8; ++ks) {
                    const uint32_t blk = (ks >> 2), koff = (ks & 3) * 32;
                    umma_tf32(tmem_base + (uint32_t)(sb * FT_BN), make_smem_desc(smem_u32(sQ) + blk * FT_BM * FT_BK * 4 + koff),
                              make_smem_desc(kb + blk * FT_BN * FT_BK * 4 + koff), idesc_s, ks != 0);
                }
                umma_commit(&bars[B_KEMPTY + buf]);   // K buffer reusable
                umma_commit(&bars[B_SFULL + sb]);     // S ready
            }
            __syncwarp();
        };
        issue_s(0);
        for (int t = 0; t < ntiles; ++t) {
            if (t + 1 < ntiles) issue_s(t + 1);
            const int buf = t & 1, sb = t % 3;
            const uint32_t use = (uint32_t)t >> 1;
            mbar_wait(&bars[B_VFULL + buf], use & 1);
            mbar_wait(&bars[B_PFULL + sb], ((uint32_t)t / 3) & 1);   // P written (and, if needed, O rescaled) by the softmax warps
            tc_fence_after();
            const uint32_t vb = smem_u32(sV + buf * FT_V_BYTES);
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < FT_BN / 8; ++ks) {
                    const uint32_t blk = (ks >> 2), koff = (ks & 3) * 32;
                    umma_tf32_ts(tmem_O, tmem_base + (uint32_t)(sb * FT_BN + ks * 8), make_smem_desc(vb + blk * FT_D * FT_BK * 4 + koff), idesc_o,
                                 (t | ks) != 0);
                }
                umma_commit(&bars[B_VFREE + buf]);    // V buffer ...
                umma_commit(&bars[B_SFREE + sb]);     // ... and the S/P buffer reusable once these MMAs are done
            }
            __syncwarp();
        }
        if (elect_one_sync()) umma_commit(&bars[B_OFULL]);
        __syncwarp();
    } else {
        // ===================== softmax / epilogue warps: thread = one query row =====================
        const int qd = warp & 3;                 // TMEM lane quarter
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        float m_ref = -INFINITY, l = 0.f;        // m_ref in the log2 domain (already multiplied by scale * log2 e)
        for (int t = 0; t < ntiles; ++t) {
            const int sb = t % 3;
            const uint32_t tS = tmem_base + (uint32_t)(sb * FT_BN) + lane_off;
            mbar_wait(&bars[B_SFULL + sb], ((uint32_t)t / 3) & 1);
            tc_fence_after();
            uint32_t v0[32], v1[32];
            tmem_ld32(tS, v0);
            tmem_ld32(tS + 32, v1);
            tmem_ld_wait();
            const int kvalid = p.Nk - (t * FT_BN - kshift);    // columns >= kvalid are padding (only ever on the last tile)
            const int klo = t == 0 ? kshift : 0;               // columns < klo are the phantom keys of a shifted tiling (first tile only)
            if (kvalid < FT_BN || klo > 0) {                   // -inf scores: they drop out of the maximum and exponentiate to exactly 0
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j >= kvalid || j < klo) v0[j] = 0xff800000u;
                    if (32 + j >= kvalid) v1[j] = 0xff800000u;
                }
            }
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
            for (int j = 0; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
            const float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
            const float m_new = fmaxf(m_ref, mt * p.scale_log2e);
            const bool need = m_new > m_ref + 8.0f;          // also true on the first tile (m_ref = -inf)
            if (__any_sync(0xffffffffu, need)) {
                const float factor = need ? ex2_approx(m_ref - m_new) : 1.0f;   // 0 on the first tile (l = 0, O not yet written)
                if (t > 0) {
                    // O is about to be rescaled: P(t-1) V(t-1) (and every earlier MMA into O) must have completed
                    mbar_wait(&bars[B_SFREE + ((t - 1) % 3)], ((uint32_t)(t - 1) / 3) & 1);
                    tc_fence_after();
                    const uint32_t tO = tmem_O + lane_off;
#pragma unroll 1
                    for (int c0 = 0; c0 < FT_D; c0 += 32) {
                        uint32_t o[32];
                        tmem_ld32(tO + c0, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * factor);
                        tmem_st32(tO + c0, o);
                    }
                    tmem_st_wait();
                }
                l *= factor;
                if (need) m_ref = m_new;
            }
            // P = exp2(s * scale*log2e - m_ref) in (0, 256].  The tensor core truncates its fp32 A operand to TF32, so adding half a
            // TF32 ulp (0x1000) to the bit pattern here makes that truncation a round-to-nearest (what cvt.rna.tf32 does, minus its
            // inf / NaN guard: P is finite).  l sums the unrounded values: the difference is an unbiased 2^-12 relative noise.
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float e0 = ex2_approx(__uint_as_float(v0[j]) * p.scale_log2e - m_ref);
                const float e1 = ex2_approx(__uint_as_float(v1[j]) * p.scale_log2e - m_ref);
                ls[j & 3] += e0 + e1;
                v0[j] = __float_as_uint(e0) + 0x1000u;
                v1[j] = __float_as_uint(e1) + 0x1000u;
            }
            l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
            tmem_st32(tS, v0);          // P(t) over S(t): same lane, same columns
            tmem_st32(tS + 32, v1);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_PFULL + sb]);
        }
        // ---- epilogue: O / l -> global (transposed through the now idle K buffers for 128-byte coalesced stores) ----
        mbar_wait(&bars[B_OFULL], 0);
        tc_fence_after();
        const float inv = 1.f / l;
        const uint32_t tO = tmem_O + lane_off;
        const uint32_t tr = smem_u32(sK) + (uint32_t)qd * (32 * 36 * 4);
        float* Ob = p.O + (int64_t)b * p.o_bs + (int64_t)h * FT_D;
#pragma unroll 1
        for (int c0 = 0; c0 < FT_D; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tO + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float a0 = __uint_as_float(v[j]) * inv, a1 = __uint_as_float(v[j + 1]) * inv, a2 = __uint_as_float(v[j + 2]) * inv,
                      a3 = __uint_as_float(v[j + 3]) * inv;
                if (p.round_out) { a0 = rn_tf32(a0); a1 = rn_tf32(a1); a2 = rn_tf32(a2); a3 = rn_tf32(a3); }
                sts_v4(tr + (uint32_t)(lane * 36 + j) * 4, a0, a1, a2, a3);
            }
            __syncwarp();
            const int c4 = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = rsub + 4 * i;
                const int qrow = q0 + qd * 32 + rr;
                const float4 x = lds_v4(tr + (uint32_t)(rr * 36 + c4) * 4);
                if (qrow < p.Nq) *reinterpret_cast<float4*>(Ob + (int64_t)qrow * p.o_ts + c0 + c4) = x;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(FT_TMEM_COLS) : "memory");
}

// V [b][n][h*64 + d] (inside a fused buffer) -> V^T [(b*H + h)*64 + d][n] (row pitch ld), rounded to nearest TF32.
__global__ void __launch_bounds__(256) transpose_v_kernel(const float* __restrict__ V, int64_t v_bs, int64_t v_ts, int N, int H,
                                                         float* __restrict__ Vt, int64_t ld) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;   // c = h*64 + d
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int n = n0 + i;
        tile[i][tx] = n < N ? V[(int64_t)b * v_bs + (int64_t)n * v_ts + c0 + tx] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, n = n0 + tx;
        if (n < ld) Vt[((int64_t)b * H * 64 + c) * ld + n] = rn_tf32(tile[tx][i]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    });
    return fn;
}
int make_map(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SIU3R_ERR_CUDA;
    cuuint64_t d[5]; cuuint64_t s[5]; cuuint32_t bx[5]; cuuint32_t e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, d, s, bx, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[siu3r_b200] flash_tc: cuTensorMapEncodeTiled failed: %d\n", (int)r);
        return SIU3R_ERR_INVALID;
    }
    return SIU3R_OK;
}

}  // namespace

extern "C" {

// V^T producer for siu3r_flash_attn_tc: Vt [(b*H + h)*64 + d][ld] with ld >= N, ld % 4 == 0; columns >= N are zero-filled up to ld.
int siu3r_transpose_v(const float* V, int64_t v_bs, int64_t v_ts, int B, int N, int H, float* Vt, int64_t ld, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(V && Vt && B > 0 && N > 0 && H > 0 && ld >= N && ld % 4 == 0);
    dim3 grid(ceil_div((int)ld, 32), H * 2, B);
    transpose_v_kernel<<<grid, 256, 0, stream>>>(V, v_bs, v_ts, N, H, Vt, ld);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// Q: rows [B][Nq] of width q_width floats (token pitch q_ts, batch pitch q_bs), head h at columns q_col0 + 64 h .. ; K likewise.
// Vt from siu3r_transpose_v.  O [B][Nq][H*64] (pitches o_bs / o_ts).  Q / K should already be RN-TF32 (the tensor core truncates).
int siu3r_flash_attn_tc(const float* Q, int64_t q_bs, int64_t q_ts, int q_width, int q_col0, const float* K, int64_t k_bs, int64_t k_ts,
                        int k_width, int k_col0, const float* Vt, int64_t vt_ld, int64_t vt_batch_cols, float* O, int64_t o_bs, int64_t o_ts, int B,
                        int H, int Nq, int Nk, float scale, int round_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(Q && K && Vt && O && B > 0 && H > 0 && Nq > 0 && Nk > 0);
    SIU3R_REQUIRE(q_ts % 4 == 0 && k_ts % 4 == 0 && q_bs % 4 == 0 && k_bs % 4 == 0 && vt_ld % 4 == 0 && o_ts % 4 == 0 && o_bs % 4 == 0);
    SIU3R_REQUIRE(((uintptr_t)Q & 15) == 0 && ((uintptr_t)K & 15) == 0 && ((uintptr_t)Vt & 15) == 0 && ((uintptr_t)O & 15) == 0);
    SIU3R_REQUIRE(q_col0 % 4 == 0 && k_col0 % 4 == 0 && q_width >= q_col0 + H * 64 && k_width >= k_col0 + H * 64 && vt_ld >= Nk && vt_batch_cols >= 0);
    CUtensorMap mq, mk, mv;
    {
        uint64_t dims[3] = {(uint64_t)q_width, (uint64_t)Nq, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)q_ts * 4, (uint64_t)q_bs * 4};
        uint32_t box[3] = {FT_BK, FT_BM, 1};
        int r = make_map(&mq, Q, 3, dims, str, box); if (r) return r;
    }
    {
        uint64_t dims[3] = {(uint64_t)k_width, (uint64_t)Nk, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)k_ts * 4, (uint64_t)k_bs * 4};
        uint32_t box[3] = {FT_BK, FT_BN, 1};
        int r = make_map(&mk, K, 3, dims, str, box); if (r) return r;
    }
    {
        uint64_t dims[2] = {(uint64_t)vt_ld, (uint64_t)(vt_batch_cols ? 1 : B) * H * 64};
        uint64_t str[1] = {(uint64_t)vt_ld * 4};
        uint32_t box[2] = {FT_BK, FT_D};
        int r = make_map(&mv, Vt, 2, dims, str, box); if (r) return r;
    }
    FlashParams p{};
    p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk; p.q_col0 = q_col0; p.k_col0 = k_col0;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.O = O; p.o_bs = o_bs; p.o_ts = o_ts; p.round_out = round_out; p.vt_batch_cols = vt_batch_cols;
    static bool attr[64] = {false};
    if (siu3r_first_use_on_device(attr))
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
    dim3 grid(ceil_div(Nq, FT_BM), H, B);
    flash_tc_kernel<<<grid, FT_THREADS, FT_SMEM, stream>>>(mq, mk, mv, p);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

}  // extern "C"
