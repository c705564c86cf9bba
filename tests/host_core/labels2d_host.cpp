// TEST INFRASTRUCTURE: runs siu3r_b200/csrc/labels2d_core.h -- the per-pixel functions the CUDA kernel is made of -- on the host, with the
// kernel's distribution of a pixel over 32 lanes emulated by a loop, so that the arithmetic can be checked against the reference goldens
// on a box without a GPU (tests/test_oracle_cpu.py).  Built by the test with g++; never part of libsiu3r_b200.so.
#include "labels2d_core.h"

extern "C" int labels2d_host(const float* logits, int V, int Q, int C, int H, int W, int64_t sv, int64_t sq, int64_t sc, int64_t sh, int64_t sw,
                             float threshold, const int* fuse_sem, const int* fuse_ins, int n_fuse, int64_t* sem_id, int64_t* ins_id,
                             int32_t* first_sem) {
    if (n_fuse > L2D_MAX_FUSE) return -1;
    L2dFuse fuse{};
    fuse.n = n_fuse;
    for (int i = 0; i < n_fuse; ++i) { fuse.sem[i] = fuse_sem[i]; fuse.ins[i] = fuse_ins[i]; }
    for (int q = 0; q < Q; ++q) first_sem[q] = -1;
    int64_t pix = 0;
    for (int v = 0; v < V; ++v)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x, ++pix) {
                const float* px = logits + v * sv + y * sh + x * sw;
                float val[32];
                int j[32], bq[32];
                for (int lane = 0; lane < 32; ++lane) l2d_lane_scan(px, sq, sc, Q, C, lane, val[lane], j[lane], bq[lane]);
                for (int o = 16; o > 0; o >>= 1) {          // the kernel's xor butterfly, all lanes updated from the previous step's values
                    float nv[32];
                    int nj[32], nq[32];
                    for (int lane = 0; lane < 32; ++lane) {
                        const int p = lane ^ o;
                        const bool take = l2d_better(val[p], j[p], val[lane], j[lane]);
                        nv[lane] = take ? val[p] : val[lane];
                        nj[lane] = take ? j[p] : j[lane];
                        nq[lane] = take ? bq[p] : bq[lane];
                    }
                    for (int lane = 0; lane < 32; ++lane) { val[lane] = nv[lane]; j[lane] = nj[lane]; bq[lane] = nq[lane]; }
                }
                for (int lane = 1; lane < 32; ++lane)
                    if (j[lane] != j[0] || bq[lane] != bq[0]) return -2;     // the butterfly must leave every lane with the same winner
                int q_idx;
                const int sem = l2d_finish(val[0], j[0], bq[0], threshold, &q_idx);
                sem_id[pix] = sem;
                ins_id[pix] = l2d_fuse(sem, q_idx, fuse);
                if (q_idx > 0 && first_sem[q_idx - 1] < 0) first_sem[q_idx - 1] = sem;     // pixels are visited in (v, h, w) order
            }
    return 0;
}
