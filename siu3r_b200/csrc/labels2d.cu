// 2-D semantic / instance label maps from the rendered query-class logits -- the step that follows the N-channel rasterisation in the
// reference's validation / test step (src/pipeline.py:132-193) and in its viewer (viewer.py:404-446).  SURVEY.md section 8(f) row 1.
//
// HBM-bound: every logit is read once (4*V*Q*C*H*W bytes), 16 bytes of labels are written per pixel.  The rasteriser emits the logits
// channel-last ([v, h, w, q*c]; the reference's einops rearrange to "n q c h w" is a view of the same memory), so one WARP owns one
// pixel: for a fixed query its lanes read the C consecutive class logits (one or two 128-byte lines), the max / argmax over queries
// is a per-lane scan and the max / argmax over classes a 5-step shuffle reduction.  Arbitrary element strides are accepted, so the
// contiguous [v, q, c, h, w] layout works too (correct, not coalesced).
#include <limits.h>

#include "common.cuh"
#include "labels2d_core.h"

namespace {

constexpr int L2D_WARPS = 8;

__global__ void __launch_bounds__(L2D_WARPS * 32)
labels2d_kernel(const float* __restrict__ logits, int64_t npix, int HW, int W, int Q, int C, int64_t sv, int64_t sq, int64_t sc, int64_t sh,
                int64_t sw, float threshold, L2dFuse fuse, int64_t* __restrict__ sem_id, int64_t* __restrict__ ins_id,
                int* __restrict__ first_pix) {
    const int lane = threadIdx.x & 31;
    const int64_t pix = (int64_t)blockIdx.x * L2D_WARPS + (threadIdx.x >> 5);
    if (pix >= npix) return;                              // warp-uniform
    const int v = (int)(pix / HW);
    const int r = (int)(pix - (int64_t)v * HW);
    const int y = r / W, x = r - y * W;
    const float* px = logits + v * sv + y * sh + x * sw;
    float val;
    int j, bq;
    l2d_lane_scan(px, sq, sc, Q, C, lane, val, j, bq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, val, o);
        const int oj = __shfl_xor_sync(0xffffffffu, j, o);
        const int oq = __shfl_xor_sync(0xffffffffu, bq, o);
        if (l2d_better(ov, oj, val, j)) { val = ov; j = oj; bq = oq; }
    }
    if (lane == 0) {
        int q_idx;
        const int sem = l2d_finish(val, j, bq, threshold, &q_idx);
        sem_id[pix] = sem;
        ins_id[pix] = l2d_fuse(sem, q_idx, fuse);
        if (q_idx > 0) atomicMin(&first_pix[q_idx - 1], (int)pix);     // first pixel (v, h, w order) that a query owns BEFORE fusing (:166-170)
    }
}

// label of the first pixel each query owns (-1: the query owns no pixel and is dropped from seg_infos, :168-169)
__global__ void labels2d_first_kernel(const int* __restrict__ first_pix, const int64_t* __restrict__ sem_id, int Q, int64_t npix, int* __restrict__ first_sem) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int p = first_pix[q];
    first_sem[q] = (p >= 0 && p < npix) ? (int)sem_id[p] : -1;
}

}  // namespace

extern "C" {

int siu3r_labels_from_qc_logits(const float* logits, int V, int Q, int C, int H, int W, int64_t sv, int64_t sq, int64_t sc, int64_t sh,
                                int64_t sw, float threshold, const int* fuse_sem, const int* fuse_ins, int n_fuse, int64_t* sem_id,
                                int64_t* ins_id, int32_t* first_pix, int32_t* first_sem, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(logits && sem_id && ins_id && first_pix && first_sem);
    SIU3R_REQUIRE(V > 0 && Q > 0 && C > 0 && H > 0 && W > 0 && n_fuse >= 0 && n_fuse <= L2D_MAX_FUSE && (n_fuse == 0 || (fuse_sem && fuse_ins)));
    const int64_t npix = (int64_t)V * H * W;
    SIU3R_REQUIRE(npix < 0x7f7f7f7f);                     // pixel indices live in int32 words whose "none" value is 0x7f7f7f7f
    L2dFuse fuse{};
    fuse.n = n_fuse;
    for (int i = 0; i < n_fuse; ++i) { fuse.sem[i] = fuse_sem[i]; fuse.ins[i] = fuse_ins[i]; }
    SIU3R_CUDA_CHECK(cudaMemsetAsync(first_pix, 0x7f, sizeof(int32_t) * Q, stream));
    labels2d_kernel<<<(unsigned)ceil_div_i64(npix, L2D_WARPS), L2D_WARPS * 32, 0, stream>>>(logits, npix, H * W, W, Q, C, sv, sq, sc, sh, sw, threshold,
                                                                                          fuse, sem_id, ins_id, first_pix);
    SIU3R_LAUNCH_CHECK();
    labels2d_first_kernel<<<ceil_div(Q, 128), 128, 0, stream>>>(first_pix, sem_id, Q, npix, first_sem);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(2);
    return SIU3R_OK;
}

}  // extern "C"
