// Image ingest on the GPU: LANCZOS resize (Pillow's fixed-point algorithm, bit-exact) + centre crop + /255 of inference.py:13-38, for raw
// 8-bit RGB frames uploaded as they were decoded.  SURVEY.md section 8(f) row 3.  Byte / integer work, HBM- (really L2-) bound: a
// 1296x968 frame is 3.8 MB, the two passes read it once and write 0.4 MB + 0.8 MB.
#include "common.cuh"
#include "resize_core.h"

namespace {

// tmp[(y - row0), ox, c] = horizontal sample of source row y at output column crop_x + ox
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int row0, int rows, int crop_x, int cw, int out_w, const int32_t* __restrict__ bounds,
                const int32_t* __restrict__ kk, int ksize, uint8_t* __restrict__ tmp) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)rows * cw * 3) return;
    const int c = (int)(idx % 3);
    const int ox = (int)((idx / 3) % cw);
    const int r = (int)(idx / (3 * (int64_t)cw));
    const int xx = crop_x + ox;
    // a crop window that sticks out of the resized image is filled with black, as PIL's Image.crop does (square inputs whose resized
    // side comes out as size - 1 by float rounding hit this: inference.py:27-33 with int(W * (256 / H)))
    tmp[idx] = (xx >= 0 && xx < out_w) ? rsz_horizontal(src, src_pitch, row0 + r, xx, c, bounds, kk, ksize) : (uint8_t)0;
}

// out[c, oy, ox] = vertical sample / 255 (planar float32: the tensor preprocess_image returns)
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t* __restrict__ tmp, int row0, int crop_x, int crop_y, int ch, int cw, int out_w, int out_h, const int32_t* __restrict__ bounds,
                const int32_t* __restrict__ kk, int ksize, float* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)3 * ch * cw) return;
    const int ox = (int)(idx % cw);
    const int oy = (int)((idx / cw) % ch);
    const int c = (int)(idx / ((int64_t)cw * ch));
    const int yy = crop_y + oy, xx = crop_x + ox;
    const bool inside = yy >= 0 && yy < out_h && xx >= 0 && xx < out_w;
    const uint8_t u = inside ? rsz_vertical(tmp, (int64_t)cw * 3, row0, yy, ox, c, bounds, kk, ksize) : (uint8_t)0;
    out[idx] = (float)u / 255.0f;
}

}  // namespace

extern "C" {

int siu3r_resize_lanczos_u8(const uint8_t* src, int H, int W, int64_t src_pitch, const int32_t* bounds_x, const int32_t* kx, int ksize_x,
                            int out_w, const int32_t* bounds_y, const int32_t* ky, int ksize_y, int out_h, int crop_x, int crop_y, int cw,
                            int ch, int row0, int rows, uint8_t* tmp, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(src && bounds_x && kx && bounds_y && ky && tmp && out);
    SIU3R_REQUIRE(H > 0 && W > 0 && src_pitch >= (int64_t)W * 3 && ksize_x > 0 && ksize_y > 0 && out_w > 0 && out_h > 0);
    SIU3R_REQUIRE(cw > 0 && ch > 0 && crop_x + cw > 0 && crop_x < out_w && crop_y + ch > 0 && crop_y < out_h);   // the window overlaps the image
    SIU3R_REQUIRE(row0 >= 0 && rows > 0 && row0 + rows <= H);
    const int64_t n_h = (int64_t)rows * cw * 3, n_v = (int64_t)3 * ch * cw;
    resize_h_kernel<<<(unsigned)ceil_div_i64(n_h, 256), 256, 0, stream>>>(src, src_pitch, row0, rows, crop_x, cw, out_w, bounds_x, kx, ksize_x, tmp);
    SIU3R_LAUNCH_CHECK();
    resize_v_kernel<<<(unsigned)ceil_div_i64(n_v, 256), 256, 0, stream>>>(tmp, row0, crop_x, crop_y, ch, cw, out_w, out_h, bounds_y, ky, ksize_y, out);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(2);
    return SIU3R_OK;
}

}  // extern "C"
