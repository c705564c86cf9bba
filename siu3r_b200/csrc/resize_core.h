// Integer arithmetic of the image ingest (reference: inference.py:13-38 preprocess_image = PIL `Image.resize(..., LANCZOS)` + centre crop + /255).
// The resampling itself is Pillow's (third-party, libImaging/Resample.c; Pillow 12.2.0 in this image): two separable passes over 8-bit
// samples -- horizontal first, rounded back to 8 bits, then vertical -- each output sample a fixed-point dot product
//     clip8((2^21 + sum_x src[x] * k[x]) >> 22),   k = round(coefficient * 2^22)
// with per-output-sample windows (first tap, tap count).  The coefficient tables are built on the host (siu3r_b200/io.py: lanczos_tables, double
// precision like Pillow's precompute_coeffs); the functions below are the per-sample work, shared by the CUDA kernels (resize.cu) and the
// host-compiled check in tests/.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RSZ_HD __host__ __device__ __forceinline__
#else
#define RSZ_HD inline
#endif

#define RSZ_PRECISION_BITS 22

RSZ_HD uint8_t rsz_clip8(int32_t ss) {
    ss >>= RSZ_PRECISION_BITS;              // arithmetic shift (negative sums clip to 0)
    return (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
}

RSZ_HD uint8_t rsz_sample(const uint8_t* src, int64_t stride, const int32_t* k, int taps) {
    int32_t ss = 1 << (RSZ_PRECISION_BITS - 1);
    for (int x = 0; x < taps; ++x) ss += (int32_t)src[(int64_t)x * stride] * k[x];
    return rsz_clip8(ss);
}

// horizontal pass: sample (row y, output column xx, channel c) of an interleaved RGB image with `pitch` bytes per row
RSZ_HD uint8_t rsz_horizontal(const uint8_t* src, int64_t pitch, int y, int xx, int c, const int32_t* bounds, const int32_t* kk, int ksize) {
    const int xmin = bounds[2 * xx], taps = bounds[2 * xx + 1];
    return rsz_sample(src + (int64_t)y * pitch + (int64_t)xmin * 3 + c, 3, kk + (int64_t)xx * ksize, taps);
}

// vertical pass over the horizontally resampled rows [row0, ...) held in tmp (interleaved RGB, `pitch` bytes per row)
RSZ_HD uint8_t rsz_vertical(const uint8_t* tmp, int64_t pitch, int row0, int yy, int x, int c, const int32_t* bounds, const int32_t* kk, int ksize) {
    const int ymin = bounds[2 * yy], taps = bounds[2 * yy + 1];
    return rsz_sample(tmp + (int64_t)(ymin - row0) * pitch + (int64_t)x * 3 + c, pitch, kk + (int64_t)yy * ksize, taps);
}
