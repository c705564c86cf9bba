// Attention kernels of the hot path.
//  (1) flash attention, head dim 64, fp32 in/out, TF32 tensor-core math (mma.sync m16n8k8; register-level 3xTF32 split
//      for the strict-parity mode).  Replaces the materialised softmax(QK^T * s)V of croco/blocks.py:105-109,162-166.
//      (A tcgen05/TMEM version is the planned successor; this one already removes the N x N HBM round trip.)
//  (2) small masked attention (head dim 32, ~100 queries) for the Mask2Former decoder:
//      mask2former/video_seg_decoder.py:975-983 (nn.MultiheadAttention with boolean attn_mask, incl. the
//      "fully masked row -> unmasked" rule of :1306-1308) and :994-999 (query self-attention).
//  (3) multi-scale deformable attention sampling: vit_adapter/blocks.py:171-213,217-267 and
//      mask2former/video_seg_decoder.py:1679-1720 (softmax over levels*points, bilinear zero-padded gathers).
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------
// (1) flash attention D = 64
// ------------------------------------------------------------------------------------------------------------
constexpr int FA_BM = 64;        // queries per CTA (4 warps x 16 rows)
constexpr int FA_BN = 64;        // keys per tile
constexpr int FA_D = 64;
constexpr int FA_LD = FA_D + 4;  // padded smem row (floats): conflict-free fragment reads for both K and V patterns
constexpr int FA_THREADS = 128;
constexpr int FA_SMEM = 2 /*K,V*/ * 2 /*stages*/ * FA_BN * FA_LD * 4;

__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NSPLIT>
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = f2tf32(x);
    if (NSPLIT == 3) lo = f2tf32(x - __uint_as_float(hi));
}

template <int NSPLIT>
__global__ void __launch_bounds__(FA_THREADS) flash_attn_d64_kernel(const float* __restrict__ Q, int64_t q_bs, int64_t q_ts,
                                                                    const float* __restrict__ K, int64_t k_bs, int64_t k_ts,
                                                                    const float* __restrict__ V, int64_t v_bs, int64_t v_ts,
                                                                    float* __restrict__ O, int64_t o_bs, int64_t o_ts, int Nq, int Nk,
                                                                    float scale, int round_out) {
    extern __shared__ __align__(16) float fa_smem[];
    float* sK = fa_smem;                          // [2][FA_BN][FA_LD]
    float* sV = fa_smem + 2 * FA_BN * FA_LD;      // [2][FA_BN][FA_LD]
    const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* Qb = Q + (int64_t)b * q_bs + (int64_t)h * FA_D;
    const float* Kb = K + (int64_t)b * k_bs + (int64_t)h * FA_D;
    const float* Vb = V + (int64_t)b * v_bs + (int64_t)h * FA_D;
    const int row0 = qt * FA_BM + warp * 16 + g, row1 = row0 + 8;

    // Q fragments for the 8 k-steps over d (hi / lo planes)
    uint32_t qh[8][4], ql[NSPLIT == 3 ? 8 : 1][4];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const float a0 = row0 < Nq ? Qb[(int64_t)row0 * q_ts + ks * 8 + t] : 0.f;
        const float a1 = row1 < Nq ? Qb[(int64_t)row1 * q_ts + ks * 8 + t] : 0.f;
        const float a2 = row0 < Nq ? Qb[(int64_t)row0 * q_ts + ks * 8 + t + 4] : 0.f;
        const float a3 = row1 < Nq ? Qb[(int64_t)row1 * q_ts + ks * 8 + t + 4] : 0.f;
        uint32_t l0 = 0, l1 = 0, l2 = 0, l3 = 0;
        split<NSPLIT>(a0, qh[ks][0], l0); split<NSPLIT>(a1, qh[ks][1], l1);
        split<NSPLIT>(a2, qh[ks][2], l2); split<NSPLIT>(a3, qh[ks][3], l3);
        if constexpr (NSPLIT == 3) { ql[ks][0] = l0; ql[ks][1] = l1; ql[ks][2] = l2; ql[ks][3] = l3; }
    }

    float o_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0s = 0.f, l1s = 0.f;

    const int ntiles = (Nk + FA_BN - 1) / FA_BN;
    auto load_tile = [&](int kt, int buf) {
        float* dk = sK + buf * FA_BN * FA_LD;
        float* dv = sV + buf * FA_BN * FA_LD;
#pragma unroll
        for (int i = 0; i < (FA_BN * FA_D / 4) / FA_THREADS; ++i) {
            const int c = threadIdx.x + i * FA_THREADS;
            const int r = c >> 4, cc = (c & 15) * 4;
            const int key = kt * FA_BN + r;
            const bool ok = key < Nk;
            const int64_t kk = ok ? key : 0;
            cp_async16(dk + r * FA_LD + cc, Kb + kk * k_ts + cc, ok);
            cp_async16(dv + r * FA_LD + cc, Vb + kk * v_ts + cc, ok);
        }
        cp_async_commit();
    };
    load_tile(0, 0);

    for (int kt = 0; kt < ntiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ntiles) { load_tile(kt + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float* tk = sK + buf * FA_BN * FA_LD;
        const float* tv = sV + buf * FA_BN * FA_LD;

        // ---- S = Q K^T (16 x 64 per warp) ----
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float k0 = tk[(nt * 8 + g) * FA_LD + ks * 8 + t];
                const float k1 = tk[(nt * 8 + g) * FA_LD + ks * 8 + t + 4];
                uint32_t b0h, b0l = 0, b1h, b1l = 0;
                split<NSPLIT>(k0, b0h, b0l); split<NSPLIT>(k1, b1h, b1l);
                if constexpr (NSPLIT == 3) {
                    mma_tf32(s[nt], ql[ks], b0h, b1h);
                    mma_tf32(s[nt], qh[ks], b0l, b1l);
                }
                mma_tf32(s[nt], qh[ks], b0h, b1h);
            }
        }
        // ---- scale, mask the key tail, online softmax ----
        const int kbase = kt * FA_BN;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = kbase + nt * 8 + 2 * t;
            s[nt][0] = c < Nk ? s[nt][0] * scale : -INFINITY;
            s[nt][1] = c + 1 < Nk ? s[nt][1] * scale : -INFINITY;
            s[nt][2] = c < Nk ? s[nt][2] * scale : -INFINITY;
            s[nt][3] = c + 1 < Nk ? s[nt][3] * scale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float corr0 = NSPLIT == 3 ? expf(m0 - mn0) : __expf(m0 - mn0);
        const float corr1 = NSPLIT == 3 ? expf(m1 - mn1) : __expf(m1 - mn1);
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            if (NSPLIT == 3) {
                s[nt][0] = expf(s[nt][0] - mn0); s[nt][1] = expf(s[nt][1] - mn0);
                s[nt][2] = expf(s[nt][2] - mn1); s[nt][3] = expf(s[nt][3] - mn1);
            } else {
                s[nt][0] = __expf(s[nt][0] - mn0); s[nt][1] = __expf(s[nt][1] - mn0);
                s[nt][2] = __expf(s[nt][2] - mn1); s[nt][3] = __expf(s[nt][3] - mn1);
            }
            rs0 += s[nt][0] + s[nt][1];
            rs1 += s[nt][2] + s[nt][3];
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0s = l0s * corr0 + rs0; l1s = l1s * corr1 + rs1;
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int i = 0; i < 8; ++i) { o_acc[i][0] *= corr0; o_acc[i][1] *= corr0; o_acc[i][2] *= corr1; o_acc[i][3] *= corr1; }

        // ---- O += P V.  The C-fragment of S is reused as the A-fragment of P by permuting the k index:
        //      k-slot t <-> key 2t, k-slot t+4 <-> key 2t+1 of each 8-key group (V rows are read in the same order). ----
        // 3xTF32: the tile's P V product goes into a fresh accumulator that is then added in registers (round-to-nearest), so
        // the tensor core's truncating accumulate never chains over more than one tile.
        float pv[NSPLIT == 3 ? 8 : 1][4];
        if constexpr (NSPLIT == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { pv[i][0] = pv[i][1] = pv[i][2] = pv[i][3] = 0.f; }
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            uint32_t ph[4], pl[4] = {0, 0, 0, 0};
            split<NSPLIT>(s[kk][0], ph[0], pl[0]);  // (row g,   key 2t)
            split<NSPLIT>(s[kk][2], ph[1], pl[1]);  // (row g+8, key 2t)
            split<NSPLIT>(s[kk][1], ph[2], pl[2]);  // (row g,   key 2t+1)
            split<NSPLIT>(s[kk][3], ph[3], pl[3]);  // (row g+8, key 2t+1)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float v0 = tv[(kk * 8 + 2 * t) * FA_LD + nt * 8 + g];
                const float v1 = tv[(kk * 8 + 2 * t + 1) * FA_LD + nt * 8 + g];
                uint32_t b0h, b0l = 0, b1h, b1l = 0;
                split<NSPLIT>(v0, b0h, b0l); split<NSPLIT>(v1, b1h, b1l);
                if constexpr (NSPLIT == 3) {
                    mma_tf32(pv[nt], pl, b0h, b1h);
                    mma_tf32(pv[nt], ph, b0l, b1l);
                    mma_tf32(pv[nt], ph, b0h, b1h);
                } else {
                    mma_tf32(o_acc[nt], ph, b0h, b1h);
                }
            }
        }
        if constexpr (NSPLIT == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { o_acc[i][0] += pv[i][0]; o_acc[i][1] += pv[i][1]; o_acc[i][2] += pv[i][2]; o_acc[i][3] += pv[i][3]; }
        }
        __syncthreads();  // tile `buf` is overwritten by the prefetch issued in the next iteration
    }
    const float inv0 = 1.f / l0s, inv1 = 1.f / l1s;
    float* Ob = O + (int64_t)b * o_bs + (int64_t)h * FA_D;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * t;
        float r0 = o_acc[nt][0] * inv0, r1 = o_acc[nt][1] * inv0, r2 = o_acc[nt][2] * inv1, r3 = o_acc[nt][3] * inv1;
        if (round_out) {  // output feeds a TF32 GEMM only: round to nearest here instead of letting the tensor core truncate
            r0 = __uint_as_float(f2tf32(r0)); r1 = __uint_as_float(f2tf32(r1)); r2 = __uint_as_float(f2tf32(r2)); r3 = __uint_as_float(f2tf32(r3));
        }
        if (row0 < Nq) *reinterpret_cast<float2*>(Ob + (int64_t)row0 * o_ts + c) = make_float2(r0, r1);
        if (row1 < Nq) *reinterpret_cast<float2*>(Ob + (int64_t)row1 * o_ts + c) = make_float2(r2, r3);
    }
}

// ------------------------------------------------------------------------------------------------------------
// (2) small attention, head dim 32, few queries (Mask2Former's 100 object queries against up to 2*64*64 pixel keys).
//     Exact fp32 FFMA (the boolean attention masks downstream are threshold decisions: no TF32 here).
//     Split over the keys ("flash-decoding"): CTA (b, h, split) stages its <= 128 keys of K and V in shared memory ONCE
//     and every thread owns one query, so K/V are read from L2 once per head instead of once per (head, query);
//     the per-split (max, sum, acc[32]) partials are merged by a second tiny kernel.  Fully masked rows attend everywhere
//     (video_seg_decoder.py:1306-1308): a row-flag pre-pass decides that before the split kernel runs.
// ------------------------------------------------------------------------------------------------------------
constexpr int AS_KEYS = 128;     // keys per split (K and V tiles: 2 x 16 KB of shared memory)
constexpr int AS_QT = 128;       // queries per CTA = threads
constexpr int AS_REC = 34;       // partial record: m, l, acc[32]

__global__ void __launch_bounds__(256) attn_small_rowflags_kernel(const uint8_t* __restrict__ mask, int Nk, uint8_t* __restrict__ all_masked) {
    // one CTA per (b, q) row: all_masked = 1 iff every key is masked
    const uint8_t* row = mask + (int64_t)blockIdx.x * Nk;
    int any = 0;
    if ((Nk & 15) == 0 && (((uintptr_t)row) & 15) == 0) {
        const uint4* r4 = reinterpret_cast<const uint4*>(row);
        for (int i = threadIdx.x; i < (Nk >> 4); i += blockDim.x) {
            const uint4 v = r4[i];   // a byte is 0 (= attend) iff (x - 0x01010101) & ~x & 0x80808080 is non-zero
            any |= (((v.x - 0x01010101u) & ~v.x) | ((v.y - 0x01010101u) & ~v.y) | ((v.z - 0x01010101u) & ~v.z) | ((v.w - 0x01010101u) & ~v.w)) & 0x80808080u;
        }
    } else {
        for (int i = threadIdx.x; i < Nk; i += blockDim.x) any |= (row[i] == 0);
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) all_masked[blockIdx.x] = any ? 0 : 1;
}

__global__ void __launch_bounds__(AS_QT) attn_small_split_kernel(const float* __restrict__ Q, int64_t q_bs, int64_t q_ts,
                                                                 const float* __restrict__ K, int64_t k_bs, int64_t k_ts,
                                                                 const float* __restrict__ V, int64_t v_bs, int64_t v_ts,
                                                                 float* __restrict__ O, int64_t o_bs, int64_t o_ts,
                                                                 const uint8_t* __restrict__ mask, const uint8_t* __restrict__ all_masked,
                                                                 float* __restrict__ part, int H, int Nq, int Nk, int keys_per_split,
                                                                 int nsplit, float scale, int round_out) {
    __shared__ __align__(16) float sK[AS_KEYS * 32];
    __shared__ __align__(16) float sV[AS_KEYS * 32];
    const int split = blockIdx.x, bh = blockIdx.y, qt = blockIdx.z;
    const int b = bh / H, h = bh % H;
    const int key0 = split * keys_per_split;
    const int nkeys = min(keys_per_split, Nk - key0);
    // ---- stage K / V of this split (coalesced: 8 threads per 128-byte key row) ----
    {
        const float* Kb = K + (int64_t)b * k_bs + h * 32;
        const float* Vb = V + (int64_t)b * v_bs + h * 32;
        for (int i = threadIdx.x; i < nkeys * 8; i += AS_QT) {
            const int key = i >> 3, c = (i & 7) * 4;
            *reinterpret_cast<float4*>(sK + key * 32 + c) = *reinterpret_cast<const float4*>(Kb + (int64_t)(key0 + key) * k_ts + c);
            *reinterpret_cast<float4*>(sV + key * 32 + c) = *reinterpret_cast<const float4*>(Vb + (int64_t)(key0 + key) * v_ts + c);
        }
    }
    __syncthreads();
    const int q = qt * AS_QT + threadIdx.x;
    if (q >= Nq) return;
    float qr[32];
    {
        const float* qp = Q + (int64_t)b * q_bs + (int64_t)q * q_ts + h * 32;
#pragma unroll
        for (int d = 0; d < 32; d += 4) {
            const float4 v = *reinterpret_cast<const float4*>(qp + d);
            qr[d] = v.x; qr[d + 1] = v.y; qr[d + 2] = v.z; qr[d + 3] = v.w;
        }
    }
    const uint8_t* mrow = nullptr;
    if (mask && !all_masked[b * Nq + q]) mrow = mask + ((int64_t)b * Nq + q) * Nk + key0;
    float m = -INFINITY, l = 0.f, acc[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] = 0.f;
    for (int c0 = 0; c0 < nkeys; c0 += 32) {
        const int cn = min(32, nkeys - c0);
        float sc[32];
        float cmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float dot = -INFINITY;
            if (j < cn && !(mrow && mrow[c0 + j])) {
                const float* kp = sK + (c0 + j) * 32;
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;   // four independent chains (latency, not throughput, bounds this loop)
#pragma unroll
                for (int d = 0; d < 32; d += 4) {
                    const float4 kv = *reinterpret_cast<const float4*>(kp + d);   // broadcast read
                    d0 += qr[d] * kv.x; d1 += qr[d + 1] * kv.y; d2 += qr[d + 2] * kv.z; d3 += qr[d + 3] * kv.w;
                }
                dot = ((d0 + d1) + (d2 + d3)) * scale;
            }
            sc[j] = dot;
            cmax = fmaxf(cmax, dot);
        }
        if (cmax == -INFINITY) continue;   // every key of this chunk masked for this query
        const float mn = fmaxf(m, cmax);
        const float corr = expf(m - mn);   // m = -inf -> 0
        l *= corr;
#pragma unroll
        for (int d = 0; d < 32; ++d) acc[d] *= corr;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (sc[j] == -INFINITY) continue;
            const float pw = expf(sc[j] - mn);
            l += pw;
            const float* vp = sV + (c0 + j) * 32;
#pragma unroll
            for (int d = 0; d < 32; d += 4) {
                const float4 vv = *reinterpret_cast<const float4*>(vp + d);
                acc[d] += pw * vv.x; acc[d + 1] += pw * vv.y; acc[d + 2] += pw * vv.z; acc[d + 3] += pw * vv.w;
            }
        }
        m = mn;
    }
    if (nsplit == 1) {
        float* op = O + (int64_t)b * o_bs + (int64_t)q * o_ts + h * 32;
        const float inv = 1.0f / l;
#pragma unroll
        for (int d = 0; d < 32; ++d) { acc[d] *= inv; if (round_out) acc[d] = __uint_as_float(f2tf32(acc[d])); }
#pragma unroll
        for (int d = 0; d < 32; d += 4) *reinterpret_cast<float4*>(op + d) = make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]);
        return;
    }
    float* rec = part + (((int64_t)bh * nsplit + split) * Nq + q) * AS_REC;
    rec[0] = m; rec[1] = l;
#pragma unroll
    for (int d = 0; d < 32; ++d) rec[2 + d] = acc[d];
}

// one warp per (b, h, q): lane = channel.  The split weights exp(m_s - M) are computed once (lane s) and broadcast, so the
// partial-record loads of all splits are independent and pipeline.
__global__ void __launch_bounds__(128) attn_small_merge_kernel(const float* __restrict__ part, float* __restrict__ O, int64_t o_bs, int64_t o_ts,
                                                               int H, int Nq, int nsplit, int total, int round_out) {
    const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= total) return;
    const int q = w % Nq, bh = w / Nq;
    const int b = bh / H, h = bh % H;
    const float* rec = part + ((int64_t)bh * nsplit * Nq + q) * AS_REC;
    const int64_t step = (int64_t)Nq * AS_REC;
    float L = 0.f, o = 0.f;
    float M = -INFINITY;
    for (int s = lane; s < nsplit; s += 32) M = fmaxf(M, rec[s * step]);
    M = warp_max(M);
    for (int s0 = 0; s0 < nsplit; s0 += 32) {
        const int s = s0 + lane;
        float g = 0.f, lg = 0.f;
        if (s < nsplit) {
            const float ms = rec[s * step];
            if (ms != -INFINITY) { g = expf(ms - M); lg = rec[s * step + 1] * g; }
        }
        L += lg;
        const int n = min(32, nsplit - s0);
#pragma unroll 8
        for (int j = 0; j < n; ++j) o += rec[(s0 + j) * step + 2 + lane] * __shfl_sync(0xffffffffu, g, j);
    }
    L = warp_sum(L);
    float r = o / L;
    if (round_out) r = __uint_as_float(f2tf32(r));
    O[(int64_t)b * o_bs + (int64_t)q * o_ts + h * 32 + lane] = r;
}

// ------------------------------------------------------------------------------------------------------------
// (3) multi-scale deformable attention sampling.  One warp per (batch, query, head); lanes span the head channels.
//     ow: [B*Lq, ldow] rows hold [offsets (nH*L*P*2) | attention logits (nH*L*P)] as produced by one fused GEMM.
// ------------------------------------------------------------------------------------------------------------
constexpr int MSDA_MAX_LP = 16;
struct MsdaLevels { int H[4]; int W[4]; int start[4]; };

template <int CPL /*channels per lane*/>
__global__ void __launch_bounds__(256) msdeform_kernel(const float* __restrict__ value, int64_t ldv, int Lin, const float* __restrict__ ow,
                                                      int64_t ldow, const float* __restrict__ ref /*[Lq,2] (x,y)*/, MsdaLevels lv, int B, int Lq,
                                                      int nH, int L, int P, float* __restrict__ out, int64_t ldo) {
    const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= B * Lq * nH) return;
    const int h = wid % nH;
    const int q = (wid / nH) % Lq;
    const int b = wid / (nH * Lq);
    const int hd = CPL * 32;
    const int LP = L * P;
    const float* row = ow + ((int64_t)b * Lq + q) * ldow;
    const float* offs = row + (int64_t)h * LP * 2;
    const float* logit = row + (int64_t)nH * LP * 2 + (int64_t)h * LP;
    float wgt[MSDA_MAX_LP];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MSDA_MAX_LP; ++i) if (i < LP) { wgt[i] = logit[i]; mx = fmaxf(mx, wgt[i]); }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MSDA_MAX_LP; ++i) if (i < LP) { wgt[i] = expf(wgt[i] - mx); sum += wgt[i]; }
    const float inv = 1.f / sum;
    const float rx = ref[2 * q], ry = ref[2 * q + 1];
    float acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
    const float* vb = value + (int64_t)b * Lin * ldv + h * hd + lane * CPL;
#pragma unroll
    for (int i = 0; i < MSDA_MAX_LP; ++i) {
        if (i >= LP) break;
        const int l = i / P;
        const int Hl = lv.H[l], Wl = lv.W[l];
        const float locx = rx + offs[2 * i] / (float)Wl, locy = ry + offs[2 * i + 1] / (float)Hl;
        const float gx = 2.f * locx - 1.f, gy = 2.f * locy - 1.f;
        const float px = ((gx + 1.f) * (float)Wl - 1.f) * 0.5f, py = ((gy + 1.f) * (float)Hl - 1.f) * 0.5f;
        const float fx = floorf(px), fy = floorf(py);
        const int x0 = (int)fx, y0 = (int)fy;
        const float ax = px - fx, ay = py - fy;
        const float w = wgt[i] * inv;
        const float* vl = vb + (int64_t)lv.start[l] * ldv;
#pragma unroll
        for (int cy = 0; cy < 2; ++cy) {
            const int yy = y0 + cy;
            if (yy < 0 || yy >= Hl) continue;
            const float wy = cy ? ay : 1.f - ay;
#pragma unroll
            for (int cx = 0; cx < 2; ++cx) {
                const int xx = x0 + cx;
                if (xx < 0 || xx >= Wl) continue;
                const float wx = cx ? ax : 1.f - ax;
                const float* p = vl + ((int64_t)yy * Wl + xx) * ldv;
                const float ww = w * wx * wy;
                if (CPL == 2) {
                    const float2 v2 = *reinterpret_cast<const float2*>(p);
                    acc[0] += ww * v2.x; acc[CPL - 1] += ww * v2.y;
                } else {
                    acc[0] += ww * p[0];
                }
            }
        }
    }
    float* op = out + ((int64_t)b * Lq + q) * ldo + h * hd + lane * CPL;
    if (CPL == 2) *reinterpret_cast<float2*>(op) = make_float2(acc[0], acc[CPL - 1]);
    else op[0] = acc[0];
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// (3b) deformable attention, vectorised: a warp serves one query and 32/(HD/4) heads at once (lane group = head, each lane
//      4 channels -> every gather is a 16-byte load and one warp instruction moves 512 bytes).  The scalar work (softmax
//      of the L*P logits, sampling coordinates, the 4 corner indices / bilinear weights) is computed ONCE per
//      (head, point) by one lane and handed over through a small per-warp shared record, instead of redundantly by all
//      32 lanes; out-of-range corners become weight 0 at a clamped address, so all loads of a point issue back to back.
// ------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256) msdeform_vec_kernel(const float* __restrict__ value, int64_t ldv, int Lin, const float* __restrict__ ow,
                                                          int64_t ldow, const float* __restrict__ ref, MsdaLevels lv, int B, int Lq, int nH, int L,
                                                          int P, float* __restrict__ out, int64_t ldo, int round_out) {
    constexpr int LPH = HD / 4;        // lanes per head
    constexpr int HPW = 32 / LPH;      // heads per warp
    __shared__ __align__(16) float s_rec[8][HPW * MSDA_MAX_LP][8];   // per warp: [item][w00 w01 w10 w11 | i00 i01 i10 i11]
    const int wslot = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nHG = nH / HPW;
    const int wid = blockIdx.x * 8 + wslot;
    if (wid >= B * Lq * nHG) return;
    const int hg = wid % nHG;
    const int q = (wid / nHG) % Lq;
    const int b = wid / (nHG * Lq);
    const int LP = L * P;
    const float* row = ow + ((int64_t)b * Lq + q) * ldow;
    const float rx = ref[2 * q], ry = ref[2 * q + 1];
    // ---- phase 1: one lane per (head, point) ----
    for (int id = lane; id < HPW * LP; id += 32) {
        const int hh = id / LP, i = id - hh * LP;
        const int h = hg * HPW + hh;
        const float* offs = row + (int64_t)h * LP * 2;
        const float* logit = row + (int64_t)nH * LP * 2 + (int64_t)h * LP;
        float mx = -INFINITY;
        for (int j = 0; j < LP; ++j) mx = fmaxf(mx, logit[j]);
        float sum = 0.f;
        for (int j = 0; j < LP; ++j) sum += expf(logit[j] - mx);
        const float inv = 1.f / sum;
        const float w = expf(logit[i] - mx) * inv;
        const int l = i / P;
        const int Hl = lv.H[l], Wl = lv.W[l], st = lv.start[l];
        const float locx = rx + offs[2 * i] / (float)Wl, locy = ry + offs[2 * i + 1] / (float)Hl;
        const float gx = 2.f * locx - 1.f, gy = 2.f * locy - 1.f;
        const float px = ((gx + 1.f) * (float)Wl - 1.f) * 0.5f, py = ((gy + 1.f) * (float)Hl - 1.f) * 0.5f;
        const float fx = floorf(px), fy = floorf(py);
        const int x0 = (int)fx, y0 = (int)fy;
        const float ax = px - fx, ay = py - fy;
        float* rec = s_rec[wslot][id];
#pragma unroll
        for (int cy = 0; cy < 2; ++cy)
#pragma unroll
            for (int cx = 0; cx < 2; ++cx) {
                const int yy = y0 + cy, xx = x0 + cx;
                const bool ok = yy >= 0 && yy < Hl && xx >= 0 && xx < Wl;
                const float wy = cy ? ay : 1.f - ay, wx = cx ? ax : 1.f - ax;
                rec[cy * 2 + cx] = ok ? w * wx * wy : 0.f;
                rec[4 + cy * 2 + cx] = __int_as_float(ok ? st + yy * Wl + xx : st);
            }
    }
    __syncwarp();
    // ---- phase 2: lane group = head, lane = 4 channels ----
    const int hh = lane / LPH;
    const int h = hg * HPW + hh;
    const float* vb = value + (int64_t)b * Lin * ldv + h * HD + (lane % LPH) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* recs = s_rec[wslot][hh * LP];
#pragma unroll 4
    for (int i = 0; i < LP; ++i) {
        const float4 wv = *reinterpret_cast<const float4*>(recs + i * 8);
        const float4 iv = *reinterpret_cast<const float4*>(recs + i * 8 + 4);
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(vb + (int64_t)__float_as_int(iv.x) * ldv));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(vb + (int64_t)__float_as_int(iv.y) * ldv));
        const float4 v2 = __ldg(reinterpret_cast<const float4*>(vb + (int64_t)__float_as_int(iv.z) * ldv));
        const float4 v3 = __ldg(reinterpret_cast<const float4*>(vb + (int64_t)__float_as_int(iv.w) * ldv));
        acc.x += wv.x * v0.x; acc.y += wv.x * v0.y; acc.z += wv.x * v0.z; acc.w += wv.x * v0.w;
        acc.x += wv.y * v1.x; acc.y += wv.y * v1.y; acc.z += wv.y * v1.z; acc.w += wv.y * v1.w;
        acc.x += wv.z * v2.x; acc.y += wv.z * v2.y; acc.z += wv.z * v2.z; acc.w += wv.z * v2.w;
        acc.x += wv.w * v3.x; acc.y += wv.w * v3.y; acc.z += wv.w * v3.z; acc.w += wv.w * v3.w;
    }
    if (round_out) {   // the sampled values only feed the output projection (a TF32 GEMM): round to nearest here
        acc.x = __uint_as_float(f2tf32(acc.x)); acc.y = __uint_as_float(f2tf32(acc.y));
        acc.z = __uint_as_float(f2tf32(acc.z)); acc.w = __uint_as_float(f2tf32(acc.w));
    }
    *reinterpret_cast<float4*>(out + ((int64_t)b * Lq + q) * ldo + h * HD + (lane % LPH) * 4) = acc;
}

extern "C" {

// out[b, n, h*64 + d] = softmax(Q K^T * scale) V per (batch, head).  Element (b, n, h, d) of X lives at
// X + b*x_bs + n*x_ts + h*64 + d (so q/k/v can point into a fused qkv buffer).  precision: 1 = TF32, 3 = 3xTF32.
int siu3r_flash_attn_d64(const float* Q, int64_t q_bs, int64_t q_ts, const float* K, int64_t k_bs, int64_t k_ts, const float* V,
                         int64_t v_bs, int64_t v_ts, float* O, int64_t o_bs, int64_t o_ts, int B, int H, int Nq, int Nk, float scale,
                         int precision, int round_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(Q && K && V && O && B > 0 && H > 0 && Nq > 0 && Nk > 0);
    SIU3R_REQUIRE(precision == 1 || precision == 3);
    SIU3R_REQUIRE(k_ts % 4 == 0 && v_ts % 4 == 0 && k_bs % 4 == 0 && v_bs % 4 == 0 && o_ts % 2 == 0 && o_bs % 2 == 0);
    SIU3R_REQUIRE(((uintptr_t)K & 15) == 0 && ((uintptr_t)V & 15) == 0 && ((uintptr_t)O & 7) == 0);
    static bool attr[64] = {false};
    if (siu3r_first_use_on_device(attr)) {
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_d64_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(flash_attn_d64_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    }
    dim3 grid(ceil_div(Nq, FA_BM), H, B);
    if (precision == 1)
        flash_attn_d64_kernel<1><<<grid, FA_THREADS, FA_SMEM, stream>>>(Q, q_bs, q_ts, K, k_bs, k_ts, V, v_bs, v_ts, O, o_bs, o_ts, Nq, Nk, scale, round_out);
    else
        flash_attn_d64_kernel<3><<<grid, FA_THREADS, FA_SMEM, stream>>>(Q, q_bs, q_ts, K, k_bs, k_ts, V, v_bs, v_ts, O, o_bs, o_ts, Nq, Nk, scale, round_out);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

static void attn_small_plan(int B, int H, int Nq, int Nk, int* keys_per_split, int* nsplit) {
    // enough CTAs for ~2 per SM: keys per split in [32, AS_KEYS], multiple of 32
    const int64_t groups = (int64_t)B * H * ceil_div(Nq, AS_QT);
    int ks = (int)ceil_div_i64((int64_t)Nk * groups, 296);
    ks = ceil_div(ks, 32) * 32;
    if (ks < 32) ks = 32;
    if (ks > AS_KEYS) ks = AS_KEYS;
    *nsplit = ceil_div(Nk, ks);
    *keys_per_split = ks;
}

// bytes of scratch siu3r_attn_small_d32 needs for this problem (row flags + split partials)
int64_t siu3r_attn_small_d32_ws_bytes(int B, int H, int Nq, int Nk) {
    int ks, ns;
    attn_small_plan(B, H, Nq, Nk, &ks, &ns);
    const int64_t flags = ((int64_t)B * Nq + 255) / 256 * 256;
    return flags + (ns > 1 ? (int64_t)B * H * ns * Nq * AS_REC * 4 : 0);
}

// Head dim 32; mask (optional) uint8 [B, Nq, Nk], non-zero = key not attended; rows that are fully masked attend everywhere.
// workspace: >= siu3r_attn_small_d32_ws_bytes(B, H, Nq, Nk) bytes of device memory, 256-byte aligned.
int siu3r_attn_small_d32(const float* Q, int64_t q_bs, int64_t q_ts, const float* K, int64_t k_bs, int64_t k_ts, const float* V, int64_t v_bs,
                         int64_t v_ts, float* O, int64_t o_bs, int64_t o_ts, const uint8_t* mask, int B, int H, int Nq, int Nk, float scale,
                         int round_out, void* workspace, int64_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(Q && K && V && O && B > 0 && H > 0 && Nq > 0 && Nk > 0);
    SIU3R_REQUIRE(q_ts % 4 == 0 && k_ts % 4 == 0 && v_ts % 4 == 0 && q_bs % 4 == 0 && k_bs % 4 == 0 && v_bs % 4 == 0 && o_ts % 4 == 0 && o_bs % 4 == 0);
    SIU3R_REQUIRE((((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) & 15) == 0);
    SIU3R_REQUIRE(workspace && workspace_bytes >= siu3r_attn_small_d32_ws_bytes(B, H, Nq, Nk) && ((uintptr_t)workspace & 255) == 0);
    int ks, ns;
    attn_small_plan(B, H, Nq, Nk, &ks, &ns);
    uint8_t* flags = (uint8_t*)workspace;
    float* part = (float*)((uint8_t*)workspace + ((int64_t)B * Nq + 255) / 256 * 256);
    int launches = 1;
    if (mask) {
        attn_small_rowflags_kernel<<<B * Nq, 256, 0, stream>>>(mask, Nk, flags);
        SIU3R_LAUNCH_CHECK();
        ++launches;
    }
    dim3 grid((unsigned)ns, (unsigned)(B * H), (unsigned)ceil_div(Nq, AS_QT));
    attn_small_split_kernel<<<grid, AS_QT, 0, stream>>>(Q, q_bs, q_ts, K, k_bs, k_ts, V, v_bs, v_ts, O, o_bs, o_ts, mask, flags, part, H, Nq, Nk, ks,
                                                        ns, scale, round_out);
    SIU3R_LAUNCH_CHECK();
    if (ns > 1) {
        const int total = B * H * Nq;
        attn_small_merge_kernel<<<ceil_div(total, 4), 128, 0, stream>>>(part, O, o_bs, o_ts, H, Nq, ns, total, round_out);
        SIU3R_LAUNCH_CHECK();
        ++launches;
    }
    siu3r_note_launch(launches);
    return SIU3R_OK;
}

// value [B, Lin, nH*hd] (ldv), ow [B*Lq, ldow] = [sampling offsets | attention logits], ref [Lq, 2] normalised (x, y),
// levels: level_hw [L][2] = (H_l, W_l) (host ints); out [B*Lq, ldo].  hd in {32, 64}; L*P <= 16.
int siu3r_msdeform_attn(const float* value, int64_t ldv, int Lin, const float* ow, int64_t ldow, const float* ref, const int* level_hw, int L,
                        int P, int B, int Lq, int nH, int hd, float* out, int64_t ldo, int round_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(value && ow && ref && level_hw && out && L >= 1 && L <= 4 && P >= 1 && L * P <= MSDA_MAX_LP);
    SIU3R_REQUIRE(hd == 32 || hd == 64);
    MsdaLevels lv{};
    int start = 0;
    for (int l = 0; l < L; ++l) { lv.H[l] = level_hw[2 * l]; lv.W[l] = level_hw[2 * l + 1]; lv.start[l] = start; start += lv.H[l] * lv.W[l]; }
    SIU3R_REQUIRE(start == Lin);
    SIU3R_REQUIRE(!round_out || (nH % (32 / (hd / 4)) == 0));   // rounding lives in the vectorised kernel
    const bool vec = (ldv % 4 == 0) && (ldo % 4 == 0) && (((uintptr_t)value | (uintptr_t)out) & 15) == 0 && nH % (32 / (hd / 4)) == 0;
    if (vec) {
        const int hpw = 32 / (hd / 4);
        const int nw = B * Lq * (nH / hpw);
        if (hd == 64) msdeform_vec_kernel<64><<<ceil_div(nw, 8), 256, 0, stream>>>(value, ldv, Lin, ow, ldow, ref, lv, B, Lq, nH, L, P, out, ldo, round_out);
        else msdeform_vec_kernel<32><<<ceil_div(nw, 8), 256, 0, stream>>>(value, ldv, Lin, ow, ldow, ref, lv, B, Lq, nH, L, P, out, ldo, round_out);
        SIU3R_LAUNCH_CHECK();
        siu3r_note_launch(1);
        return SIU3R_OK;
    }
    const int warps = B * Lq * nH;
    if (hd == 64)
        msdeform_kernel<2><<<ceil_div(warps, 8), 256, 0, stream>>>(value, ldv, Lin, ow, ldow, ref, lv, B, Lq, nH, L, P, out, ldo);
    else
        msdeform_kernel<1><<<ceil_div(warps, 8), 256, 0, stream>>>(value, ldv, Lin, ow, ldow, ref, lv, B, Lq, nH, L, P, out, ldo);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

}  // extern "C"
