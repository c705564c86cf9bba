// Per-pixel arithmetic of the 2-D label extraction from rendered query-class logits (src/pipeline.py:132-193; same code in
// viewer.py:404-446).  Plain C++ so that the device kernel (labels2d.cu) and the host-compiled check of tests/ run the very same
// functions; the kernel only adds the distribution of a pixel's classes over the lanes of a warp.
//
// For one pixel of one rendered view, with logits L[q][c] (q surviving queries, c = classes + void, void LAST):
//   c_logit[c], q_index[c] = max / argmax over q                       (pipeline.py:143)
//   rotate the class axis so that void comes FIRST: j = 0 <-> c = C-1, j >= 1 <-> c = j-1   (:145-150)
//   sem_logit, sem_id = max / argmax over j                            (:151)
//   q_idx = q_index[sem_id] + 1                                        (:161)
//   sem_logit < threshold (0.3) -> sem_id = 0;  sem_id == 0 -> q_idx = 0   (:162-164)
//   stuff classes: pixels with sem_id == stuff + 1 get instance id num_queries + stuff + 1   (:182-186), passed as (sem -> ins) pairs
// Ties resolve to the first index, as torch.max does on the CPU.
#pragma once
#include <limits.h>
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define L2D_HD __host__ __device__ __forceinline__
#else
#define L2D_HD inline
#endif

#define L2D_MAX_FUSE 8

// stuff fusing as (semantic id -> instance id) pairs: pipeline.py:182-186 uses (stuff + 1 -> num_queries + stuff + 1) for stuff in
// label_ids_to_fuse; the viewer (viewer.py:433-434) hard-codes (1 -> 102), (2 -> 103)
struct L2dFuse {
    int n;
    int sem[L2D_MAX_FUSE];
    int ins[L2D_MAX_FUSE];
};

// class index in the logits for position j of the rotated (void-first) axis
L2D_HD int l2d_class_of(int j, int C) { return j == 0 ? C - 1 : j - 1; }

// max / first argmax over the queries of one class; px -> L[0][0] of the pixel, element strides sq, sc
L2D_HD void l2d_scan_queries(const float* px, int64_t sq, int64_t sc, int Q, int c, float& best, int& best_q) {
    const float* p = px + (int64_t)c * sc;
    best = p[0];
    best_q = 0;
    for (int q = 1; q < Q; ++q) {
        const float v = p[(int64_t)q * sq];
        if (v > best) { best = v; best_q = q; }
    }
}

// does candidate (vb, jb) beat (va, ja) in the argmax over the rotated class axis?  (greater value, or equal value at a lower position)
L2D_HD bool l2d_better(float vb, int jb, float va, int ja) { return vb > va || (vb == va && jb < ja); }

// One lane's share of a pixel: positions j = lane, lane + 32, ... of the rotated class axis.  A lane without a class returns
// (-inf, INT_MAX, 0), which loses against every real candidate (and ties only with other empty lanes).
L2D_HD void l2d_lane_scan(const float* px, int64_t sq, int64_t sc, int Q, int C, int lane, float& val, int& j_best, int& q_best) {
    val = -INFINITY;
    j_best = INT_MAX;
    q_best = 0;
    for (int j = lane; j < C; j += 32) {
        float b;
        int bq;
        l2d_scan_queries(px, sq, sc, Q, l2d_class_of(j, C), b, bq);
        if (l2d_better(b, j, val, j_best)) { val = b; j_best = j; q_best = bq; }
    }
}

// threshold + void handling; returns the semantic id, *q_idx = instance id before stuff fusing (0 = none)
L2D_HD int l2d_finish(float sem_logit, int j, int best_q, float threshold, int* q_idx) {
    int sem = j;
    if (sem_logit < threshold) sem = 0;
    *q_idx = sem == 0 ? 0 : best_q + 1;
    return sem;
}

// instance id after stuff fusing
L2D_HD int l2d_fuse(int sem, int q_idx, const L2dFuse& f) {
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 0; i < L2D_MAX_FUSE; ++i)       // fixed trip count: the ids stay in registers / constant bank
        if (i < f.n && sem == f.sem[i]) q_idx = f.ins[i];
    return q_idx;
}
