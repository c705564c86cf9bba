// Flash attention on the 5th-gen tensor cores (tcgen05 + TMEM), head dim 64, fp32-grade accuracy on the fp16 MMA path ("h3", see h3.cuh).
//
//   O[b, n, h*64 + d] = softmax_k( Q K^T * scale ) V        replaces croco/blocks.py:105-109 (Attention) and :162-166
//                                                            (CrossAttention) without materialising the N x N matrix.
//
// Q, K and V^T arrive as fp16 (hi, lo) plane pairs with an UNSCALED lo plane (lo = fp16(x - hi); the projection GEMM that produces them is
// launched with lo_scale = 1): all three partial products of a split operand pair then add up in ONE accumulator,
//     S = Q_hi K_hi^T + Q_lo K_hi^T + Q_hi K_lo^T,      O += P_hi V_hi + P_lo V_hi + P_hi V_lo,
// which keeps the kernel at 256 TMEM columns (two CTAs per SM).  q, k, v are O(1) activations, so the unscaled residues (~2^-11 |x|) stay
// far above the fp16 subnormal step (6e-8) that bounds their absolute error.
//
// One CTA = 128 queries of one (batch, head); keys in tiles of 64, ONE pass (online softmax with lazy rescaling):
//   warp 0   : TMA producer   Q tile once (hi|lo: one 4-D box); per key tile K [64 keys x 64 d] and V^T [64 d x 64 keys] hi|lo, double-buffered
//   warp 1   : MMA issuer     S(t) : 4 k-steps x 3 kind::f16 MMAs (M=128, N=64, K=16)          -> TMEM S/P buffer t % 3 (64 columns)
//                             O += P V : 4 k-steps x 3 MMAs, A = P_hi / P_lo read straight from TMEM (two fp16 per 32-bit column)
//   warps 2-5: softmax        one query row per thread; P = exp2(S*scale*log2e - m_ref) in fp32, split into fp16 hi / lo and written back
//                             over S in place with tcgen05.st (P_hi -> columns [0,32), P_lo -> [32,64) of the buffer); finally O / l -> global
//                             (fp32 or a scaled-lo plane pair for the output projection) through a shared-memory transpose.
#include <cuda.h>
#include <cuda_fp16.h>

#include "h3.cuh"

namespace {
using namespace h3;

constexpr int FH_BM = 128, FH_BN = 64, FH_D = 64;
constexpr int FH_THREADS = 192;
constexpr int FH_Q_BYTES = 2 * FH_BM * 128;              // 32 KB : hi | lo planes of [128 q x 64 d] fp16
constexpr int FH_K_BYTES = 2 * FH_BN * 128;              // 16 KB per buffer : hi | lo of [64 keys x 64 d]
constexpr int FH_V_BYTES = 2 * FH_D * 128;               // 16 KB per buffer : hi | lo of [64 d x 64 keys]
constexpr int FH_SMEM = FH_Q_BYTES + 2 * FH_K_BYTES + 2 * FH_V_BYTES + 1024 + 256;   // 97.3 KB -> two CTAs per SM
constexpr int FH_TMEM_COLS = 256;                        // S/P 0..2: [0,64) [64,128) [128,192)   O: [192,256)
static_assert(2 * FH_K_BYTES >= 4 * 32 * 36 * 4, "the epilogue transpose tiles alias the K buffers");

struct FlashH3Params {
    int B, H, Nq, Nk;
    int q_col0, k_col0;          // column of head 0 inside the Q / K row
    float scale_log2e;
    float* O; int64_t o_bs, o_ts;                 // fp32 output, or
    __half* Oh; int64_t oh_bs, oh_ts, oh_plane;   // plane-pair output (scaled lo)
    long long vt_batch_cols;     // 0: V^T rows indexed by (b, h, d); > 0: rows (h, d), image b at column offset b * vt_batch_cols ...
    int vt_b_split; long long vt_extra;   // ... + vt_extra for images b >= vt_b_split (second window of a grouped projection)
    int swap_halves;             // debugging aid: pack (odd, even) instead of (even, odd) keys per TMEM column
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t saddr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
    return r;
}
// (p0, p1) -> packed fp16 hi pair and packed fp16 residue pair (unscaled), first element in the low half
__device__ __forceinline__ void split_pair(float p0, float p1, uint32_t& hi2, uint32_t& lo2) {
    hi2 = h3_pack2(p0, p1);          // packed conversions: FMA pipe, not the XU pipe the exponentials run on (h3.cuh)
    const float2 f = h3_unpack2(hi2);
    lo2 = h3_pack2(p0 - f.x, p1 - f.y);
}

// barrier indices (double-buffered ones: index + buffer).  K / V tiles are double-buffered, the S/P accumulator is TRIPLE-buffered: P(t)
// aliases S(t), so with two buffers S(t+1) could only be issued after P(t-1) V(t-1) had completed.
enum { B_Q = 0, B_KFULL = 1, B_KEMPTY = 3, B_SFULL = 5, B_PFULL = 8, B_VFULL = 11, B_VFREE = 13, B_SFREE = 15, B_OFULL = 18, B_COUNT };

// Lazy rescaling: the running reference maximum m_ref (log2 domain) of a row is only raised when the tile maximum exceeds it by more than 8,
// so P = exp2(s - m_ref) <= 256 stays well inside fp16 and the O accumulator in TMEM is read-modify-written only when some row of the warp
// actually moves its reference (first tile aside, almost never).  softmax is shift invariant: O and l carry the same factor.
__global__ void __launch_bounds__(FH_THREADS, 2)
flash_h3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmVt,
                const FlashH3Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + FH_Q_BYTES;            // 2 buffers
    uint8_t* sV = sK + 2 * FH_K_BYTES;        // 2 buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * FH_V_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (shuffle broadcast: provably warp-uniform)
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * FH_BM;
    // The innermost (key) coordinate of a V^T box must be 16-byte aligned.  When image b's keys start at an unaligned column (vt_batch_cols =
    // tokens per image, e.g. 1025) the key tiling is shifted down by kshift = start & 7 keys: tile t covers keys [t*64 - kshift, ...), the (at
    // most 7) phantom keys in front of key 0 are masked like the padding behind key Nk-1.
    const long long vcol0 = p.vt_batch_cols ? (long long)b * p.vt_batch_cols + (b >= p.vt_b_split ? p.vt_extra : 0) : 0;
    const int kshift = (int)(vcol0 & 7);
    const int ntiles = (p.Nk + kshift + FH_BN - 1) / FH_BN;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmQ);
        prefetch_tmap(&tmK);
        prefetch_tmap(&tmVt);
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
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(FH_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                 // programmatic dependent launch (common.cuh): the prologue above overlapped the tail of the kernel that produced Q/K/V^T
    pdl_launch_dependents();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_O = tmem_base + 3 * FH_BN;

    if (warp == 0) {
        // ===================== TMA producer (whole warp walks the loop, one elected lane issues: uniform-datapath operands) =====================
        if (elect_one_sync()) {
            mbar_expect_tx(&bars[B_Q], FH_Q_BYTES);
            tma_load_4d(&tmQ, &bars[B_Q], sQ, p.q_col0 + h * FH_D, q0, b, 0);
        }
        __syncwarp();
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            const uint32_t use = (uint32_t)t >> 1;
            mbar_wait(&bars[B_KEMPTY + buf], (use & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&bars[B_KFULL + buf], FH_K_BYTES);
                tma_load_4d(&tmK, &bars[B_KFULL + buf], sK + buf * FH_K_BYTES, p.k_col0 + h * FH_D, t * FH_BN - kshift, b, 0);
            }
            __syncwarp();
            mbar_wait(&bars[B_VFREE + buf], (use & 1) ^ 1);      // P V of tile t-2 has read this V buffer
            if (elect_one_sync()) {
                mbar_expect_tx(&bars[B_VFULL + buf], FH_V_BYTES);
                tma_load_3d(&tmVt, &bars[B_VFULL + buf], sV + buf * FH_V_BYTES, (int)vcol0 + t * FH_BN - kshift,
                            (p.vt_batch_cols ? h : b * p.H + h) * FH_D, 0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: S(t+1) is issued before P(t) V(t); whole warp, one elected lane issues =====================
        constexpr uint32_t idesc = make_idesc_f16(FH_BM, FH_BN);   // S: N = 64 keys; O: N = 64 = head dim
        mbar_wait(&bars[B_Q], 0);
        const uint64_t dQh = make_smem_desc(smem_u32(sQ)), dQl = make_smem_desc(smem_u32(sQ) + FH_BM * 128);
        auto issue_s = [&](int t) {
            const int buf = t & 1, sb = t % 3;
            const uint32_t use = (uint32_t)t >> 1;
            mbar_wait(&bars[B_KFULL + buf], use & 1);
            mbar_wait(&bars[B_SFREE + sb], (((uint32_t)t / 3) & 1) ^ 1);   // P(t-3) (aliasing this S buffer) has been consumed
            tc_fence_after();
            const uint32_t kh = smem_u32(sK + buf * FH_K_BYTES);
            const uint64_t dKh = make_smem_desc(kh), dKl = make_smem_desc(kh + FH_BN * 128);
            const uint32_t tS = tmem_base + (uint32_t)(sb * FH_BN);
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < FH_D / 16; ++ks) {
                    const uint64_t ko = (uint64_t)(ks * 2);     // 32 bytes per k-step in the descriptor's 16-byte units
                    umma_f16(tS, dQl + ko, dKh + ko, idesc, ks != 0);
                    umma_f16(tS, dQh + ko, dKl + ko, idesc, 1u);
                    umma_f16(tS, dQh + ko, dKh + ko, idesc, 1u);
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
            const uint32_t vh = smem_u32(sV + buf * FH_V_BYTES);
            const uint64_t dVh = make_smem_desc(vh), dVl = make_smem_desc(vh + FH_D * 128);
            const uint32_t tPh = tmem_base + (uint32_t)(sb * FH_BN), tPl = tPh + 32;
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < FH_BN / 16; ++ks) {
                    const uint64_t ko = (uint64_t)(ks * 2);
                    umma_f16_ts(tmem_O, tPl + ks * 8, dVh + ko, idesc, (t | ks) != 0);
                    umma_f16_ts(tmem_O, tPh + ks * 8, dVl + ko, idesc, 1u);
                    umma_f16_ts(tmem_O, tPh + ks * 8, dVh + ko, idesc, 1u);
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
            const uint32_t tS = tmem_base + (uint32_t)(sb * FH_BN) + lane_off;
            mbar_wait(&bars[B_SFULL + sb], ((uint32_t)t / 3) & 1);
            tc_fence_after();
            uint32_t v0[32], v1[32];
            tmem_ld32(tS, v0);
            tmem_ld32(tS + 32, v1);
            tmem_ld_wait();
            const int kvalid = p.Nk - (t * FH_BN - kshift);    // columns >= kvalid are padding (only ever on the last tile)
            const int klo = t == 0 ? kshift : 0;               // columns < klo are the phantom keys of a shifted tiling (first tile only)
            if (kvalid < FH_BN || klo > 0) {                   // -inf scores: they drop out of the maximum and exponentiate to exactly 0
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
                    for (int c0 = 0; c0 < FH_D; c0 += 32) {
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
            // P = exp2(s * scale*log2e - m_ref) in [0, 256]; keys (2c, 2c+1) -> column c of the P_hi block, their fp16 residues -> column c of
            // the P_lo block.  l sums the unrounded values.
            float ls[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t ph[32], pl[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float e0 = ex2_approx(__uint_as_float(v0[j]) * p.scale_log2e - m_ref);
                const float e1 = ex2_approx(__uint_as_float(v0[j + 1]) * p.scale_log2e - m_ref);
                const float f0 = ex2_approx(__uint_as_float(v1[j]) * p.scale_log2e - m_ref);
                const float f1 = ex2_approx(__uint_as_float(v1[j + 1]) * p.scale_log2e - m_ref);
                ls[(j >> 1) & 3] += (e0 + e1) + (f0 + f1);
                if (p.swap_halves) { split_pair(e1, e0, ph[j >> 1], pl[j >> 1]); split_pair(f1, f0, ph[16 + (j >> 1)], pl[16 + (j >> 1)]); }
                else { split_pair(e0, e1, ph[j >> 1], pl[j >> 1]); split_pair(f0, f1, ph[16 + (j >> 1)], pl[16 + (j >> 1)]); }
            }
            l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
            tmem_st32(tS, ph);          // P_hi(t) | P_lo(t) over S(t): same lane, same 64 columns
            tmem_st32(tS + 32, pl);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_PFULL + sb]);
        }
        // ---- epilogue: O / l -> global (transposed through the now idle K buffers for coalesced stores) ----
        mbar_wait(&bars[B_OFULL], 0);
        tc_fence_after();
        const float inv = 1.f / l;
        const uint32_t tO = tmem_O + lane_off;
        const uint32_t tr = smem_u32(sK) + (uint32_t)qd * (32 * 36 * 4);
#pragma unroll 1
        for (int c0 = 0; c0 < FH_D; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tO + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                sts_v4(tr + (uint32_t)(lane * 36 + j) * 4, __uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv, __uint_as_float(v[j + 2]) * inv,
                       __uint_as_float(v[j + 3]) * inv);
            __syncwarp();
            const int c4 = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = rsub + 4 * i;
                const int qrow = q0 + qd * 32 + rr;
                const float4 x = lds_v4(tr + (uint32_t)(rr * 36 + c4) * 4);
                if (qrow < p.Nq) {
                    const int col = h * FH_D + c0 + c4;
                    if (p.Oh) {
                        uint2 hi, lo;
                        h3_split2(x.x, x.y, hi.x, lo.x);
                        h3_split2(x.z, x.w, hi.y, lo.y);
                        __half* d = p.Oh + (int64_t)b * p.oh_bs + (int64_t)qrow * p.oh_ts + col;
                        *reinterpret_cast<uint2*>(d) = hi;
                        *reinterpret_cast<uint2*>(d + p.oh_plane) = lo;
                    } else {
                        *reinterpret_cast<float4*>(p.O + (int64_t)b * p.o_bs + (int64_t)qrow * p.o_ts + col) = x;
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(FH_TMEM_COLS) : "memory");
}

int g_swap_halves = 0;

}  // namespace

extern "C" {

// debugging aid (not part of the product ABI): packing order of the two fp16 P values of one TMEM column
void siu3r_flash_h3_debug_swap(int swap) { g_swap_halves = swap; }

// Q: rows [B][Nq] of q_width fp16 columns (token pitch q_ts, batch pitch q_bs, UNSCALED lo plane q_plane elements after the hi plane), head h at
// columns q_col0 + 64 h ..; K likewise.  Vt: plane pair [rows][vt_ld] with rows = (b*H + h)*64 + d (vt_batch_cols = 0) or h*64 + d with image b
// at column offset b * vt_batch_cols (+ vt_extra for images b >= vt_b_split > 0: the second window of a grouped projection), unscaled lo.  Exactly one of O (fp32 [B][Nq][H*64], pitches o_bs / o_ts) and Oh (plane pair with the
// standard 2^11-scaled lo, pitches oh_bs / oh_ts, plane oh_plane) receives the result.
int siu3r_flash_attn_h3(const void* Q, int64_t q_bs, int64_t q_ts, int64_t q_plane, int q_width, int q_col0, const void* K, int64_t k_bs,
                        int64_t k_ts, int64_t k_plane, int k_width, int k_col0, const void* Vt, int64_t vt_ld, int64_t vt_plane,
                        int64_t vt_batch_cols, int vt_b_split, int64_t vt_extra, float* O, int64_t o_bs, int64_t o_ts, void* Oh, int64_t oh_bs, int64_t oh_ts, int64_t oh_plane, int B,
                        int H, int Nq, int Nk, float scale, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(Q && K && Vt && (O != nullptr) != (Oh != nullptr) && B > 0 && H > 0 && Nq > 0 && Nk > 0);
    SIU3R_REQUIRE(q_ts % 8 == 0 && k_ts % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 && q_plane % 8 == 0 && k_plane % 8 == 0 && vt_ld % 8 == 0 &&
                  vt_plane % 8 == 0);
    SIU3R_REQUIRE(((uintptr_t)Q & 15) == 0 && ((uintptr_t)K & 15) == 0 && ((uintptr_t)Vt & 15) == 0);
    SIU3R_REQUIRE(q_col0 % 8 == 0 && k_col0 % 8 == 0 && q_width >= q_col0 + H * 64 && k_width >= k_col0 + H * 64 && vt_ld >= Nk && vt_batch_cols >= 0);
    if (O) SIU3R_REQUIRE(((uintptr_t)O & 15) == 0 && o_ts % 4 == 0 && o_bs % 4 == 0);
    if (Oh) SIU3R_REQUIRE(((uintptr_t)Oh & 7) == 0 && oh_ts % 4 == 0 && oh_bs % 4 == 0 && oh_plane % 4 == 0);
    CUtensorMap mq, mk, mv;
    {
        uint64_t dims[4] = {(uint64_t)q_width, (uint64_t)Nq, (uint64_t)B, 2};
        uint64_t str[3] = {(uint64_t)q_ts * 2, (uint64_t)q_bs * 2, (uint64_t)q_plane * 2};
        uint32_t box[4] = {FH_D, FH_BM, 1, 2};
        int r = make_map_f16(&mq, Q, 4, dims, str, box); if (r) return r;
    }
    {
        uint64_t dims[4] = {(uint64_t)k_width, (uint64_t)Nk, (uint64_t)B, 2};
        uint64_t str[3] = {(uint64_t)k_ts * 2, (uint64_t)k_bs * 2, (uint64_t)k_plane * 2};
        uint32_t box[4] = {FH_D, FH_BN, 1, 2};
        int r = make_map_f16(&mk, K, 4, dims, str, box); if (r) return r;
    }
    {
        uint64_t dims[3] = {(uint64_t)vt_ld, (uint64_t)(vt_batch_cols ? 1 : B) * H * 64, 2};
        uint64_t str[2] = {(uint64_t)vt_ld * 2, (uint64_t)vt_plane * 2};
        uint32_t box[3] = {FH_BN, FH_D, 2};
        int r = make_map_f16(&mv, Vt, 3, dims, str, box); if (r) return r;
    }
    FlashH3Params p{};
    p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk; p.q_col0 = q_col0; p.k_col0 = k_col0;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.O = O; p.o_bs = o_bs; p.o_ts = o_ts; p.Oh = (__half*)Oh; p.oh_bs = oh_bs; p.oh_ts = oh_ts; p.oh_plane = oh_plane;
    p.vt_batch_cols = vt_batch_cols; p.vt_b_split = vt_b_split > 0 ? vt_b_split : 0x7fffffff; p.vt_extra = vt_extra; p.swap_halves = g_swap_halves;
    SIU3R_CUDA_CHECK(cudaFuncSetAttribute(flash_h3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_SMEM));   // per device, cheap
    dim3 grid(ceil_div(Nq, FH_BM), H, B);
    SIU3R_CUDA_CHECK(siu3r_launch_pdl(flash_h3_kernel, grid, dim3(FH_THREADS), FH_SMEM, stream, mq, mk, mv, p));
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// V [b][n][h*64 + d] (fp32, inside a fused buffer) -> V^T plane pair [(b*H + h)*64 + d][n] (pitch ld, unscaled lo); columns >= N are zero-filled
// up to ld.  Only used when the projection GEMM could not write V^T itself.
static __global__ void __launch_bounds__(256) transpose_v_h3_kernel(const float* __restrict__ V, int64_t v_bs, int64_t v_ts, int N, int H,
                                                            __half* __restrict__ Vt, int64_t ld, int64_t plane) {
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
        if (n < ld) {
            const float x = tile[tx][i];
            const __half hi = __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
            const int64_t o = ((int64_t)b * H * 64 + c) * ld + n;
            Vt[o] = hi;
            Vt[plane + o] = __float2half_rn(x - __half2float(hi));
        }
    }
}

int siu3r_transpose_v_h3(const float* V, int64_t v_bs, int64_t v_ts, int B, int N, int H, void* Vt, int64_t ld, int64_t plane, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(V && Vt && B > 0 && N > 0 && H > 0 && ld >= N && ld % 8 == 0 && plane % 8 == 0);
    dim3 grid(ceil_div((int)ld, 32), H * 2, B);
    transpose_v_h3_kernel<<<grid, 256, 0, stream>>>(V, v_bs, v_ts, N, H, (__half*)Vt, ld, plane);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

}  // extern "C"
