// Flash attention on the 5th-gen tensor cores (tcgen05 + TMEM), head dim 64, fp32 storage, TF32 math.
//
//   O[b, n, h*64 + d] = softmax_k( Q K^T * scale ) V        replaces croco/blocks.py:105-109 (Attention) and :162-166
//                                                            (CrossAttention) without materialising the N x N matrix.
//
// One CTA = 128 queries of one (batch, head); keys are visited in tiles of 128.
//   warp 0   : TMA producer   Q tile once; per key tile K [128 keys x 64 d] and V^T [64 d x 128 keys] (both K-major, 128B-swizzled)
//   warp 1   : MMA issuer     S = Q K^T   : tcgen05.mma kind::tf32 M=128 N=128 K=8 x 8   -> TMEM columns [0,128)
//                             O += P V    : M=128 N=64 K=8 x 16 (A = P from shared memory)  -> TMEM columns [128,192)
//   warps 2-5: softmax        one query row per thread (tcgen05.ld 32x32b gives exactly that: no shuffles);
//                             P = exp2((S - m) * scale*log2e) rounded to TF32 and written into the swizzled K-major A-operand
//                             layout in shared memory; finally O / l -> global through a shared-memory transpose.
// Two passes over the key tiles: pass 1 only computes the row maxima m (S is recomputed in pass 2), so the O accumulator in
// TMEM never needs rescaling -- QK^T is cheap on the tensor cores, a TMEM read-modify-write of O per tile is not.
// V must be supplied transposed ([b*H + h][d][key], row pitch vt_ld) so that every UMMA operand is K-major:
// siu3r_transpose_v produces it (and rounds it to TF32) from the fused qkv / kv buffer.
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace {

constexpr int FT_BM = 128, FT_BN = 128, FT_D = 64, FT_BK = 32;
constexpr int FT_THREADS = 192;
constexpr int FT_Q_BYTES = 2 * FT_BM * FT_BK * 4;        // 32 KB : 2 k-blocks [128 x 32]
constexpr int FT_K_BYTES = 2 * FT_BN * FT_BK * 4;        // 32 KB per buffer, double-buffered
constexpr int FT_V_BYTES = 4 * FT_D * FT_BK * 4;         // 32 KB : 4 k-blocks [64 d x 32 keys]
constexpr int FT_P_BYTES = 4 * FT_BM * FT_BK * 4;        // 64 KB : 4 k-blocks [128 q x 32 keys]
constexpr int FT_SMEM = FT_Q_BYTES + 2 * FT_K_BYTES + FT_V_BYTES + FT_P_BYTES + 1024 + 256;
constexpr int FT_TMEM_COLS = 512;                        // S0: [0,128)  S1: [128,256)  O: [256,320)

struct FlashParams {
    int B, H, Nq, Nk;
    int q_col0, k_col0;          // column of head 0 inside the Q / K row
    float scale_log2e;
    float* O; int64_t o_bs, o_ts;
    int round_out;
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

// barrier indices (K / S barriers are double-buffered: index + buffer)
enum { B_Q = 0, B_KFULL = 1, B_KEMPTY = 3, B_SFULL = 5, B_SEMPTY = 7, B_VFULL = 9, B_VEMPTY, B_PFULL, B_OFULL, B_COUNT };

__global__ void __launch_bounds__(FT_THREADS, 1)
flash_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmVt,
                const FlashParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + FT_Q_BYTES;            // 2 buffers
    uint8_t* sV = sK + 2 * FT_K_BYTES;
    uint8_t* sP = sV + FT_V_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + FT_P_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * FT_BM;
    const int ntiles = (p.Nk + FT_BN - 1) / FT_BN;
    const int niter = 2 * ntiles;             // pass 1 (row maxima) then pass 2 (P, O)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmVt) : "memory");
        mbar_init(&bars[B_Q], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[B_KFULL + i], 1); mbar_init(&bars[B_KEMPTY + i], 1);
            mbar_init(&bars[B_SFULL + i], 1); mbar_init(&bars[B_SEMPTY + i], 4);   // one arrival per softmax warp
        }
        mbar_init(&bars[B_VFULL], 1); mbar_init(&bars[B_VEMPTY], 1);
        mbar_init(&bars[B_PFULL], 4);
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
    const uint32_t tmem_O = tmem_base + 256;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(&bars[B_Q], FT_Q_BYTES);
            tma_load_3d(&tmQ, &bars[B_Q], sQ, p.q_col0 + h * FT_D, q0, b);
            tma_load_3d(&tmQ, &bars[B_Q], sQ + FT_BM * FT_BK * 4, p.q_col0 + h * FT_D + FT_BK, q0, b);
            uint32_t vcnt = 0;
            for (int it = 0; it < niter; ++it) {
                const int t = it < ntiles ? it : it - ntiles;
                const int buf = it & 1;
                const uint32_t use = (uint32_t)it >> 1;
                mbar_wait(&bars[B_KEMPTY + buf], (use & 1) ^ 1);
                uint8_t* dK = sK + buf * FT_K_BYTES;
                mbar_expect_tx(&bars[B_KFULL + buf], FT_K_BYTES);
                tma_load_3d(&tmK, &bars[B_KFULL + buf], dK, p.k_col0 + h * FT_D, t * FT_BN, b);
                tma_load_3d(&tmK, &bars[B_KFULL + buf], dK + FT_BN * FT_BK * 4, p.k_col0 + h * FT_D + FT_BK, t * FT_BN, b);
                if (it >= ntiles) {
                    mbar_wait(&bars[B_VEMPTY], (vcnt & 1) ^ 1);
                    mbar_expect_tx(&bars[B_VFULL], FT_V_BYTES);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        tma_load_2d(&tmVt, &bars[B_VFULL], sV + j * FT_D * FT_BK * 4, t * FT_BN + j * FT_BK, (b * p.H + h) * FT_D);
                    ++vcnt;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (S of tile it+1 is issued before the P V of tile it) =====================
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_tf32(FT_BM, FT_BN);
            constexpr uint32_t idesc_o = make_idesc_tf32(FT_BM, FT_D);
            mbar_wait(&bars[B_Q], 0);
            auto issue_s = [&](int it) {
                const int buf = it & 1;
                const uint32_t use = (uint32_t)it >> 1;
                mbar_wait(&bars[B_KFULL + buf], use & 1);
                mbar_wait(&bars[B_SEMPTY + buf], (use & 1) ^ 1);   // softmax finished reading the previous S in this buffer
                tc_fence_after();
                const uint32_t kb = smem_u32(sK + buf * FT_K_BYTES);
#pragma unroll
                for (int ks = 0; ks < FT_D / 8; ++ks) {
                    const uint32_t blk = (ks >> 2), koff = (ks & 3) * 32;
                    umma_tf32(tmem_base + (uint32_t)buf * 128, make_smem_desc(smem_u32(sQ) + blk * FT_BM * FT_BK * 4 + koff),
                              make_smem_desc(kb + blk * FT_BN * FT_BK * 4 + koff), idesc_s, ks != 0);
                }
                umma_commit(&bars[B_KEMPTY + buf]);   // K buffer reusable
                umma_commit(&bars[B_SFULL + buf]);    // S ready
            };
            issue_s(0);
            uint32_t vcnt = 0;
            for (int it = 0; it < niter; ++it) {
                if (it + 1 < niter) issue_s(it + 1);
                if (it >= ntiles) {
                    mbar_wait(&bars[B_VFULL], vcnt & 1);
                    mbar_wait(&bars[B_PFULL], vcnt & 1);   // P written (and fenced) by the softmax warps
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < FT_BN / 8; ++ks) {
                        const uint32_t blk = (ks >> 2), koff = (ks & 3) * 32;
                        umma_tf32(tmem_O, make_smem_desc(smem_u32(sP) + blk * FT_BM * FT_BK * 4 + koff),
                                  make_smem_desc(smem_u32(sV) + blk * FT_D * FT_BK * 4 + koff), idesc_o, (vcnt | ks) != 0);
                    }
                    umma_commit(&bars[B_VEMPTY]);   // V tile and P buffer reusable once these MMAs are done
                    ++vcnt;
                }
            }
            umma_commit(&bars[B_OFULL]);
        }
    } else {
        // ===================== softmax / epilogue warps: thread = one query row =====================
        const int qd = warp & 3;                 // TMEM lane quarter
        const int r = qd * 32 + lane;            // row in the tile
        float m = -INFINITY, l = 0.f;
        uint32_t vcnt = 0;
        for (int it = 0; it < niter; ++it) {
            const int t = it < ntiles ? it : it - ntiles;
            const int buf = it & 1;
            const uint32_t tS = tmem_base + (uint32_t)buf * 128 + ((uint32_t)(qd * 32) << 16);
            mbar_wait(&bars[B_SFULL + buf], ((uint32_t)it >> 1) & 1);
            tc_fence_after();
            const int kvalid = p.Nk - t * FT_BN;   // columns >= kvalid are padding
            if (it < ntiles) {
#pragma unroll 1
                for (int c0 = 0; c0 < FT_BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tS + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < kvalid) m = fmaxf(m, __uint_as_float(v[j]));
                }
            } else {
                const float ms = m * p.scale_log2e;
#pragma unroll 1
                for (int c0 = 0; c0 < FT_BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tS + c0, v);
                    tmem_ld_wait();
                    float pv[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float e = (c0 + j < kvalid) ? exp2f(__uint_as_float(v[j]) * p.scale_log2e - ms) : 0.f;
                        pv[j] = rn_tf32(e);
                        l += pv[j];
                    }
                    // the P buffer is shared by consecutive tiles: wait for the previous tile's P V only now, after the first exps
                    if (c0 == 0 && vcnt > 0) mbar_wait(&bars[B_VEMPTY], (vcnt - 1) & 1);
                    // k-block c0/32 of P: row r, 128 B per row, 16-byte chunk c stored at chunk (c ^ (r & 7))
                    const uint32_t rowaddr = smem_u32(sP) + (uint32_t)(c0 / 32) * FT_BM * FT_BK * 4 + (uint32_t)r * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        sts_v4(rowaddr + (uint32_t)((c ^ (r & 7)) * 16), pv[4 * c], pv[4 * c + 1], pv[4 * c + 2], pv[4 * c + 3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
                ++vcnt;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (it >= ntiles) mbar_arrive(&bars[B_PFULL]);
                mbar_arrive(&bars[B_SEMPTY + buf]);
            }
        }
        // ---- epilogue: O / l -> global (transposed through the now idle P buffer for 128-byte coalesced stores) ----
        mbar_wait(&bars[B_OFULL], 0);
        tc_fence_after();
        const float inv = 1.f / l;
        const uint32_t tO = tmem_O + ((uint32_t)(qd * 32) << 16);
        const uint32_t tr = smem_u32(sP) + (uint32_t)qd * (32 * 36 * 4);
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
                        int k_width, int k_col0, const float* Vt, int64_t vt_ld, float* O, int64_t o_bs, int64_t o_ts, int B, int H, int Nq,
                        int Nk, float scale, int round_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(Q && K && Vt && O && B > 0 && H > 0 && Nq > 0 && Nk > 0);
    SIU3R_REQUIRE(q_ts % 4 == 0 && k_ts % 4 == 0 && q_bs % 4 == 0 && k_bs % 4 == 0 && vt_ld % 4 == 0 && o_ts % 4 == 0 && o_bs % 4 == 0);
    SIU3R_REQUIRE(((uintptr_t)Q & 15) == 0 && ((uintptr_t)K & 15) == 0 && ((uintptr_t)Vt & 15) == 0 && ((uintptr_t)O & 15) == 0);
    SIU3R_REQUIRE(q_col0 % 4 == 0 && k_col0 % 4 == 0 && q_width >= q_col0 + H * 64 && k_width >= k_col0 + H * 64 && vt_ld >= Nk);
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
        uint64_t dims[2] = {(uint64_t)vt_ld, (uint64_t)B * H * 64};
        uint64_t str[1] = {(uint64_t)vt_ld * 4};
        uint32_t box[2] = {FT_BK, FT_D};
        int r = make_map(&mv, Vt, 2, dims, str, box); if (r) return r;
    }
    FlashParams p{};
    p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk; p.q_col0 = q_col0; p.k_col0 = k_col0;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.O = O; p.o_bs = o_bs; p.o_ts = o_ts; p.round_out = round_out;
    static bool attr = false;
    if (!attr) {
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        attr = true;
    }
    dim3 grid(ceil_div(Nq, FT_BM), H, B);
    flash_tc_kernel<<<grid, FT_THREADS, FT_SMEM, stream>>>(mq, mk, mv, p);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

}  // extern "C"
