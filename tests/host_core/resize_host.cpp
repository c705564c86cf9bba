// TEST INFRASTRUCTURE: the two passes of siu3r_b200/csrc/resize.cu run on the host through the SAME per-sample functions
// (siu3r_b200/csrc/resize_core.h) and the same index arithmetic, so that tables + integer passes can be checked against PIL on a box without
// a GPU (tests/test_io.py).  Built by the test with g++; never part of libsiu3r_b200.so.
#include <stdlib.h>

#include "resize_core.h"

extern "C" int resize_host(const uint8_t* src, int H, int W, int64_t src_pitch, const int32_t* bounds_x, const int32_t* kx, int ksize_x, int out_w,
                           const int32_t* bounds_y, const int32_t* ky, int ksize_y, int out_h, int crop_x, int crop_y, int cw, int ch, int row0,
                           int rows, float* out) {
    if (row0 < 0 || rows <= 0 || row0 + rows > H) return -1;
    uint8_t* tmp = (uint8_t*)malloc((size_t)rows * cw * 3);
    for (int64_t idx = 0; idx < (int64_t)rows * cw * 3; ++idx) {              // resize_h_kernel
        const int c = (int)(idx % 3);
        const int ox = (int)((idx / 3) % cw);
        const int r = (int)(idx / (3 * (int64_t)cw));
        const int xx = crop_x + ox;
        tmp[idx] = (xx >= 0 && xx < out_w) ? rsz_horizontal(src, src_pitch, row0 + r, xx, c, bounds_x, kx, ksize_x) : (uint8_t)0;
    }
    for (int64_t idx = 0; idx < (int64_t)3 * ch * cw; ++idx) {                // resize_v_kernel
        const int ox = (int)(idx % cw);
        const int oy = (int)((idx / cw) % ch);
        const int c = (int)(idx / ((int64_t)cw * ch));
        const int yy = crop_y + oy, xx = crop_x + ox;
        const bool inside = yy >= 0 && yy < out_h && xx >= 0 && xx < out_w;
        const uint8_t u = inside ? rsz_vertical(tmp, (int64_t)cw * 3, row0, yy, ox, c, bounds_y, ky, ksize_y) : (uint8_t)0;
        out[idx] = (float)u / 255.0f;
    }
    free(tmp);
    return 0;
}
