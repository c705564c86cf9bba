// 3D Gaussian splatting rasterizer, forward pass, written from scratch for sm_100a.
//
// Drop-in for what the reference reaches through
//   /root/reference/src/models/cuda_splatting.py:90-118  (GaussianRasterizer(settings)(means3D, ..., cov3D_precomp))
// i.e. the third-party `diff-gaussian-rasterization-w-pose @ 43e21bf` forward.
//
// Pipeline (all on the caller's stream):
//   preprocess (cull / project / EWA covariance / SH->RGB / tile rect)   HBM-bound, 340 B read + 52 B written per Gaussian
//   3-kernel inclusive scan of tiles_touched
//   duplicate_with_keys                                                   12 B written per duplicate
//   radix sort of (tile<<32 | depth bits, gaussian id)                    (cub::DeviceRadixSort for now)
//   identify_tile_ranges                                                  8 B read per duplicate
//   render: one 16x16 CTA per tile, 256-Gaussian shared-memory batches    44 B staged per duplicate, 20 B written per pixel
//
// Integer outputs (radii, tiles_touched, offsets, sorted key/value list, tile ranges) are bit-exact against
// oracle/raster_ref.c: every float op that feeds them is an explicit round-to-nearest intrinsic (no FMA contraction),
// in the oracle's operation order.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

constexpr int TILE_X = 16;
constexpr int TILE_Y = 16;
constexpr int PRE_THREADS = 128;
constexpr int MAX_SH_FLOATS = 75;  // 25 coefficients x 3 channels

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define DIV(a, b) __fdiv_rn((a), (b))

__constant__ float c_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
__constant__ float c_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                 -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};

struct Camera {
    float view[16];
    float proj[16];
    float campos[3];
    float bg[3];
};

// x' = m[0]x + m[4]y + m[8]z + m[12], left-to-right, no contraction
__device__ __forceinline__ float row_dot(const float* m, int r, float x, float y, float z) {
    return ADD(ADD(ADD(MUL(m[r], x), MUL(m[4 + r], y)), MUL(m[8 + r], z)), m[12 + r]);
}

__device__ __forceinline__ float ndc2pix(float v, int S) { return MUL(SUB(MUL(ADD(v, 1.0f), (float)S), 1.0f), 0.5f); }

struct PreOut {
    float* depths;
    float2* xy;
    float4* conic_o;
    float* rgb;
    uint32_t* tiles;
    ushort4* rects;
};

// One thread per Gaussian; the CTA's SH block (PRE_THREADS x sh_floats, contiguous in HBM) is staged through shared
// memory with coalesced 128-bit loads, then each thread reads its own record (odd stride -> conflict-free).
__global__ void __launch_bounds__(PRE_THREADS) preprocess_kernel(
    int G, int H, int W, int gx, int gy, int deg, int sh_coeffs, int sh_layout, int cov_stride,
    const float* __restrict__ means, const float* __restrict__ cov, const float* __restrict__ shs,
    const float* __restrict__ opac, const float* __restrict__ cam_dev, float tanx, float tany, float fx, float fy,
    PreOut o, int32_t* __restrict__ radii) {
    __shared__ float s_sh[PRE_THREADS * MAX_SH_FLOATS];
    __shared__ Camera s_cam;
    const int tid = threadIdx.x;
    const int base = blockIdx.x * PRE_THREADS;
    const int shf = sh_coeffs * 3;
    if (tid < (int)(sizeof(Camera) / sizeof(float))) reinterpret_cast<float*>(&s_cam)[tid] = cam_dev[tid];
    {
        const int nvalid = min(PRE_THREADS, G - base);
        const size_t total = (size_t)nvalid * shf;
        const float* src = shs + (size_t)base * shf;
        if ((((uintptr_t)src) & 15) == 0) {
            const size_t n4 = total / 4;
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(s_sh);
            for (size_t i = tid; i < n4; i += PRE_THREADS) d4[i] = __ldg(s4 + i);
            for (size_t i = n4 * 4 + tid; i < total; i += PRE_THREADS) s_sh[i] = __ldg(src + i);
        } else {
            for (size_t i = tid; i < total; i += PRE_THREADS) s_sh[i] = __ldg(src + i);
        }
    }
    __syncthreads();
    const int i = base + tid;
    if (i >= G) return;
    radii[i] = 0;          // culled Gaussians keep these zeros (the kernel zeroes its own outputs: no memset launches in front of it)
    o.tiles[i] = 0;

    const float px = means[3 * (size_t)i], py = means[3 * (size_t)i + 1], pz = means[3 * (size_t)i + 2];
    const float* vm = s_cam.view;
    const float* pm = s_cam.proj;
    float tx = row_dot(vm, 0, px, py, pz), ty = row_dot(vm, 1, px, py, pz);
    const float tz = row_dot(vm, 2, px, py, pz);
    if (tz <= 0.2f) return;
    const float hx = row_dot(pm, 0, px, py, pz), hy = row_dot(pm, 1, px, py, pz);
    const float hw = row_dot(pm, 3, px, py, pz);
    const float pw = DIV(1.0f, ADD(hw, 0.0000001f));
    const float ndx = MUL(hx, pw), ndy = MUL(hy, pw);

    // --- EWA covariance (oracle/raster_ref.c: cov2d) ---
    float c3[6];
    {
        const float* c = cov + (size_t)i * cov_stride;
        if (cov_stride == 6) {
#pragma unroll
            for (int k = 0; k < 6; ++k) c3[k] = c[k];
        } else {  // full 3x3 row-major (Gaussians.covariances): upper triangle xx,xy,xz,yy,yz,zz
            c3[0] = c[0]; c3[1] = c[1]; c3[2] = c[2]; c3[3] = c[4]; c3[4] = c[5]; c3[5] = c[8];
        }
    }
    const float limx = MUL(1.3f, tanx), limy = MUL(1.3f, tany);
    const float txtz = DIV(tx, tz), tytz = DIV(ty, tz);
    tx = MUL(fminf(limx, fmaxf(-limx, txtz)), tz);
    ty = MUL(fminf(limy, fmaxf(-limy, tytz)), tz);
    const float j00 = DIV(fx, tz), j02 = DIV(-MUL(fx, tx), MUL(tz, tz));
    const float j11 = DIV(fy, tz), j12 = DIV(-MUL(fy, ty), MUL(tz, tz));
    float M0[3], M1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float r0 = vm[4 * c + 0], r1 = vm[4 * c + 1], r2 = vm[4 * c + 2];  // R[k][c] = vm[4c+k]
        M0[c] = ADD(ADD(MUL(j00, r0), MUL(0.0f, r1)), MUL(j02, r2));
        M1[c] = ADD(ADD(MUL(0.0f, r0), MUL(j11, r1)), MUL(j12, r2));
    }
    const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float MV0[3], MV1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        MV0[c] = ADD(ADD(MUL(M0[0], V[0][c]), MUL(M0[1], V[1][c])), MUL(M0[2], V[2][c]));
        MV1[c] = ADD(ADD(MUL(M1[0], V[0][c]), MUL(M1[1], V[1][c])), MUL(M1[2], V[2][c]));
    }
    const float ca = ADD(ADD(ADD(MUL(MV0[0], M0[0]), MUL(MV0[1], M0[1])), MUL(MV0[2], M0[2])), 0.3f);
    const float cb = ADD(ADD(MUL(MV1[0], M0[0]), MUL(MV1[1], M0[1])), MUL(MV1[2], M0[2]));
    const float cc = ADD(ADD(ADD(MUL(MV1[0], M1[0]), MUL(MV1[1], M1[1])), MUL(MV1[2], M1[2])), 0.3f);

    const float det = SUB(MUL(ca, cc), MUL(cb, cb));
    if (det == 0.0f) return;
    const float det_inv = DIV(1.0f, det);
    const float mid = MUL(0.5f, ADD(ca, cc));
    const float disc = __fsqrt_rn(fmaxf(0.1f, SUB(MUL(mid, mid), det)));
    const float lambda1 = ADD(mid, disc), lambda2 = SUB(mid, disc);
    const float my_radius = ceilf(MUL(3.0f, __fsqrt_rn(fmaxf(lambda1, lambda2))));
    const float pix = ndc2pix(ndx, W), piy = ndc2pix(ndy, H);
    const int rminx = min(gx, max(0, (int)DIV(SUB(pix, my_radius), (float)TILE_X)));
    const int rminy = min(gy, max(0, (int)DIV(SUB(piy, my_radius), (float)TILE_Y)));
    const int rmaxx = min(gx, max(0, (int)DIV(SUB(ADD(ADD(pix, my_radius), (float)TILE_X), 1.0f), (float)TILE_X)));
    const int rmaxy = min(gy, max(0, (int)DIV(SUB(ADD(ADD(piy, my_radius), (float)TILE_Y), 1.0f), (float)TILE_Y)));
    if ((rmaxx - rminx) * (rmaxy - rminy) == 0) return;

    // --- SH -> RGB (degrees 0..3 of the 25 stored coefficients; float tolerance, contraction allowed) ---
    {
        const float* sh = s_sh + tid * shf;
        const int sk = sh_layout == 0 ? 3 : 1;          // stride between coefficients
        const int sc = sh_layout == 0 ? 1 : sh_coeffs;  // stride between channels
        const float dx = px - s_cam.campos[0], dy = py - s_cam.campos[1], dz = pz - s_cam.campos[2];
        const float len = sqrtf(dx * dx + dy * dy + dz * dz);
        const float x = dx / len, y = dy / len, z = dz / len;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float* s = sh + ch * sc;
            float r = 0.28209479177387814f * s[0];
            if (deg > 0) {
                const float C1 = 0.4886025119029199f;
                r = r - C1 * y * s[1 * sk] + C1 * z * s[2 * sk] - C1 * x * s[3 * sk];
                if (deg > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    r = r + c_SH_C2[0] * xy * s[4 * sk] + c_SH_C2[1] * yz * s[5 * sk] +
                        c_SH_C2[2] * (2.0f * zz - xx - yy) * s[6 * sk] + c_SH_C2[3] * xz * s[7 * sk] +
                        c_SH_C2[4] * (xx - yy) * s[8 * sk];
                    if (deg > 2) {
                        r = r + c_SH_C3[0] * y * (3.0f * xx - yy) * s[9 * sk] + c_SH_C3[1] * xy * z * s[10 * sk] +
                            c_SH_C3[2] * y * (4.0f * zz - xx - yy) * s[11 * sk] +
                            c_SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * s[12 * sk] +
                            c_SH_C3[4] * x * (4.0f * zz - xx - yy) * s[13 * sk] + c_SH_C3[5] * z * (xx - yy) * s[14 * sk] +
                            c_SH_C3[6] * x * (xx - 3.0f * yy) * s[15 * sk];
                    }
                }
            }
            r += 0.5f;
            o.rgb[3 * (size_t)i + ch] = fmaxf(r, 0.0f);
        }
    }
    o.depths[i] = tz;
    radii[i] = (int32_t)my_radius;
    o.xy[i] = make_float2(pix, piy);
    o.conic_o[i] = make_float4(MUL(cc, det_inv), MUL(-cb, det_inv), MUL(ca, det_inv), opac[i]);
    o.tiles[i] = (uint32_t)((rmaxy - rminy) * (rmaxx - rminx));
    o.rects[i] = make_ushort4((unsigned short)rminx, (unsigned short)rminy, (unsigned short)rmaxx, (unsigned short)rmaxy);
}

// ---------------------------------------------------------------------------------------------------------------
// Inclusive scan of tiles_touched (uint32), 3 kernels: per-CTA scan + CTA totals, scan of totals, add CTA prefix.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_inclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += n;
        }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = w;
    }
    __syncthreads();
    total = s_warp[SCAN_THREADS / 32 - 1];
    const uint32_t prefix = warp > 0 ? s_warp[warp - 1] : 0;
    __syncthreads();
    return v + prefix;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_local_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                  uint32_t* __restrict__ block_sums, int n) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        sum += v[k];
        v[k] = sum;
    }
    uint32_t total;
    const uint32_t incl = block_inclusive_scan(sum, s_warp, total);
    const uint32_t excl = incl - sum;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] = v[k] + excl;
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the CTA totals in place; writes the grand total to *total_out
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(uint32_t* __restrict__ block_sums, int nblocks,
                                                                 uint32_t* __restrict__ total_out) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    uint32_t carry = 0;
    for (int start = 0; start < nblocks; start += SCAN_THREADS) {
        const int i = start + threadIdx.x;
        const uint32_t v = i < nblocks ? block_sums[i] : 0u;
        uint32_t total;
        const uint32_t incl = block_inclusive_scan(v, s_warp, total);
        if (i < nblocks) block_sums[i] = carry + incl - v;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ block_sums, int n) {
    const uint32_t add = block_sums[blockIdx.x];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) duplicate_with_keys_kernel(int G, int gx, const int32_t* __restrict__ radii,
                                                                 const uint32_t* __restrict__ offsets, const float* __restrict__ depths,
                                                                 const ushort4* __restrict__ rects, uint64_t* __restrict__ keys,
                                                                 uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    if (radii[i] <= 0) return;
    uint32_t off = i == 0 ? 0u : offsets[i - 1];
    const ushort4 r = rects[i];
    const uint64_t dbits = (uint64_t)__float_as_uint(depths[i]);
    for (uint32_t y = r.y; y < r.w; ++y)
        for (uint32_t x = r.x; x < r.z; ++x) {
            keys[off] = ((uint64_t)(y * (uint32_t)gx + x) << 32) | dbits;
            vals[off] = (uint32_t)i;
            ++off;
        }
}

__global__ void __launch_bounds__(256) identify_tile_ranges_kernel(uint32_t D, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    const uint32_t t = (uint32_t)(keys[j] >> 32);
    if (j == 0) ranges[t].x = 0;
    else {
        const uint32_t tp = (uint32_t)(keys[j - 1] >> 32);
        if (t != tp) {
            ranges[tp].y = j;
            ranges[t].x = j;
        }
    }
    if (j == D - 1) ranges[t].y = D;
}

// ---------------------------------------------------------------------------------------------------------------
// Binned fast path (replaces: scan over Gaussians -> duplicateWithKeys -> 6-pass global radix sort -> identifyTileRanges).
// The global sort only ever orders records WITHIN a tile (the tile id is the high key word), so it is done per tile in shared memory:
//   preprocess counts the duplicates of every tile (atomics on <= 8160 counters);
//   tile_scan_kernel  : exclusive scan of the tile counts -> ranges[t] = (begin, end), total D, largest tile;
//   bin_scatter_kernel: every Gaussian appends (depth bits << 32 | id) to the segments of its tiles (arbitrary order inside a segment);
//   tile_sort_kernel  : one CTA per tile loads its segment into shared memory, bitonic-sorts the 64-bit words and writes the ids back.
// (depth, id) is exactly the order the stable radix sort of (tile << 32 | depth) produces from records emitted in id order, so the
// sorted list -- and everything downstream -- is bit-identical to the reference pipeline (tests/test_ops_gpu.py).  Tiles with more than
// TS_CAP records fall back to the global sort for the whole frame.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TS_CAP = 8192;          // records per tile that fit the shared-memory sort (64 KB)
constexpr int TS_THREADS = 256;
constexpr int TS_FAST = 2048;         // largest tile for which the binned path is used (measured: above it the global radix sort is faster)

__global__ void __launch_bounds__(1024) tile_scan_kernel(const uint32_t* __restrict__ counts, int ntiles, uint2* __restrict__ ranges,
                                                         uint32_t* __restrict__ cursors, uint32_t* __restrict__ total_max, uint32_t capacity,
                                                         uint32_t* __restrict__ status) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry, s_max;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry = 0; s_max = 0; }
    __syncthreads();
    uint32_t mymax = 0;
    for (int start = 0; start < ntiles; start += 1024) {
        const int i = start + threadIdx.x;
        const uint32_t c = i < ntiles ? counts[i] : 0u;
        mymax = max(mymax, c);
        uint32_t v = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += n; }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += n; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + (warp > 0 ? s_warp[warp - 1] : 0u) + v - c;
        if (i < ntiles) { ranges[i] = make_uint2(excl, excl + c); cursors[i] = excl; }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_warp[31];
        __syncthreads();
    }
    atomicMax(&s_max, mymax);
    __syncthreads();
    if (threadIdx.x == 0) {
        total_max[0] = s_carry; total_max[1] = s_max;
        if (status) {   // no-sync protocol: the later kernels of the frame read the flags and skip their work; the host looks at them once per batch of cameras
            status[0] = s_carry; status[1] = s_max;
            status[2] = (s_carry > capacity ? 1u : 0u) | (s_max > (uint32_t)TS_CAP ? 2u : 0u);
        }
    }
}

// Counting and scattering go through a per-CTA shared-memory histogram of the tiles: a CTA walks BIN_CHUNK Gaussians, so a tile that is hit k
// times by the chunk costs one global atomic instead of k (500k splats @512^2: ~9x fewer, and far less same-address serialisation).
constexpr int BIN_THREADS = 256, BIN_CHUNK = 4096;

__global__ void __launch_bounds__(BIN_THREADS) bin_count_kernel(int G, int gx, int ntiles, const int32_t* __restrict__ radii,
                                                               const ushort4* __restrict__ rects, uint32_t* __restrict__ tile_counts) {
    extern __shared__ uint32_t s_cnt[];
    for (int t = threadIdx.x; t < ntiles; t += BIN_THREADS) s_cnt[t] = 0;
    __syncthreads();
    const int g0 = blockIdx.x * BIN_CHUNK, g1 = min(G, g0 + BIN_CHUNK);
    for (int i = g0 + threadIdx.x; i < g1; i += BIN_THREADS) {
        if (radii[i] <= 0) continue;
        const ushort4 r = rects[i];
        for (uint32_t y = r.y; y < r.w; ++y)
            for (uint32_t x = r.x; x < r.z; ++x) atomicAdd(&s_cnt[y * (uint32_t)gx + x], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ntiles; t += BIN_THREADS) {
        const uint32_t c = s_cnt[t];
        if (c) atomicAdd(&tile_counts[t], c);
    }
}

__global__ void __launch_bounds__(BIN_THREADS) bin_scatter_kernel(int G, int gx, int ntiles, const int32_t* __restrict__ radii,
                                                                 const float* __restrict__ depths, const ushort4* __restrict__ rects,
                                                                 uint32_t* __restrict__ cursors, uint64_t* __restrict__ list,
                                                                 const uint32_t* __restrict__ status) {
    if (status && status[2]) return;
    extern __shared__ uint32_t s_bin[];          // [ntiles] count, then local cursor | [ntiles] base of this CTA's run inside the tile segment
    uint32_t* s_cnt = s_bin;
    uint32_t* s_base = s_bin + ntiles;
    for (int t = threadIdx.x; t < ntiles; t += BIN_THREADS) s_cnt[t] = 0;
    __syncthreads();
    const int g0 = blockIdx.x * BIN_CHUNK, g1 = min(G, g0 + BIN_CHUNK);
    for (int i = g0 + threadIdx.x; i < g1; i += BIN_THREADS) {
        if (radii[i] <= 0) continue;
        const ushort4 r = rects[i];
        for (uint32_t y = r.y; y < r.w; ++y)
            for (uint32_t x = r.x; x < r.z; ++x) atomicAdd(&s_cnt[y * (uint32_t)gx + x], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ntiles; t += BIN_THREADS) {
        const uint32_t c = s_cnt[t];
        s_base[t] = c ? atomicAdd(&cursors[t], c) : 0u;   // reserve a contiguous run for this CTA
        s_cnt[t] = 0;
    }
    __syncthreads();
    for (int i = g0 + threadIdx.x; i < g1; i += BIN_THREADS) {
        if (radii[i] <= 0) continue;
        const ushort4 r = rects[i];
        const uint64_t word = ((uint64_t)__float_as_uint(depths[i]) << 32) | (uint32_t)i;
        for (uint32_t y = r.y; y < r.w; ++y)
            for (uint32_t x = r.x; x < r.z; ++x) {
                const uint32_t t = y * (uint32_t)gx + x;
                list[s_base[t] + atomicAdd(&s_cnt[t], 1u)] = word;
            }
    }
}

// One CTA per tile: bitonic sort of the tile's (depth bits << 32 | id) words in shared memory.  Every thread keeps TS_ITEMS consecutive
// elements, so the sub-passes with stride < TS_ITEMS stay in registers-by-way-of-private-smem without barriers.
__global__ void __launch_bounds__(TS_THREADS) tile_sort_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ list,
                                                               uint32_t* __restrict__ sorted_ids, uint64_t* __restrict__ sorted_words) {
    extern __shared__ uint64_t s_w[];
    const uint2 rg = ranges[blockIdx.x];
    const int n = (int)(rg.y - rg.x);
    if (n <= 0 || n > TS_CAP) return;
    int P = 1;
    while (P < n) P <<= 1;
    for (int i = threadIdx.x; i < P; i += TS_THREADS) s_w[i] = i < n ? list[rg.x + i] : ~0ull;
    __syncthreads();
    const int half = P >> 1;
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int jm = j - 1;
            for (int t = threadIdx.x; t < half; t += TS_THREADS) {
                const int lo = ((t & ~jm) << 1) | (t & jm), hi = lo + j;     // j is a power of two
                const bool up = (lo & k) == 0;
                const uint64_t a = s_w[lo], b = s_w[hi];
                if ((a > b) == up) { s_w[lo] = b; s_w[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += TS_THREADS) {
        const uint64_t w = s_w[i];
        sorted_ids[rg.x + i] = (uint32_t)w;
        if (sorted_words) sorted_words[rg.x + i] = w;
    }
}

// Register-resident variant: every thread keeps ITEMS consecutive elements of the (padded) tile list, P = 256 * ITEMS.  Bitonic sub-stages
// with stride j < ITEMS are compare-exchanges inside a thread, ITEMS <= j < 32 ITEMS are warp shuffles (partner lane = lane ^ j/ITEMS, same
// register), only j >= 32 ITEMS goes through shared memory (element-major layout s[r * 256 + tid]: conflict free): for P = 2048 that is 6 of
// the 66 sub-stages and 12 CTA barriers instead of 66.  Handles the tiles with n_lo < n <= n_hi; one launch per size class.
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

template <int ITEMS>
__device__ __forceinline__ void tile_sort_reg_body(const uint2 rg, const int n, const uint64_t* __restrict__ list, uint32_t* __restrict__ sorted_ids,
                                                   uint64_t* s_w) {
    constexpr int P = TS_THREADS * ITEMS;
    const int tid = threadIdx.x;
    const int e0 = tid * ITEMS;
    uint64_t v[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) v[r] = (e0 + r) < n ? list[rg.x + e0 + r] : ~0ull;
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j >= 32 * ITEMS; j >>= 1) {            // partner in another warp
            const int m = j / ITEMS;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) s_w[r * TS_THREADS + tid] = v[r];
            __syncthreads();
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                const uint64_t o = s_w[r * TS_THREADS + (tid ^ m)];
                const int e = e0 + r;
                const bool keep_min = ((e & j) == 0) == ((e & k) == 0);
                v[r] = keep_min ? (o < v[r] ? o : v[r]) : (o > v[r] ? o : v[r]);
            }
            __syncthreads();
        }
        for (int j = min(k >> 1, 16 * ITEMS); j >= ITEMS; j >>= 1) {   // partner lane in the same warp
            const int m = j / ITEMS;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                const uint64_t o = shfl_xor_u64(v[r], m);
                const int e = e0 + r;
                const bool keep_min = ((e & j) == 0) == ((e & k) == 0);
                v[r] = keep_min ? (o < v[r] ? o : v[r]) : (o > v[r] ? o : v[r]);
            }
        }
#pragma unroll
        for (int J = ITEMS / 2; J >= 1; J >>= 1) {                   // partner register in the same thread
            if (J <= (k >> 1)) {
#pragma unroll
                for (int r = 0; r < ITEMS; ++r) {
                    if ((r & J) == 0) {
                        const bool up = ((e0 + r) & k) == 0;
                        const uint64_t a = v[r], b = v[r | J];
                        if ((a > b) == up) { v[r] = b; v[r | J] = a; }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ITEMS; ++r)
        if (e0 + r < n) sorted_ids[rg.x + e0 + r] = (uint32_t)v[r];
}

// Two size classes per launch (A keys per thread for n_lo < n <= n_mid, B for n_mid < n <= n_hi): the no-sync path cannot know the largest tile on
// the host, so it launches every class over all tiles: the two small classes share a launch (the two large ones would
// need 190 registers together and stay separate).
template <int A, int B>
__global__ void __launch_bounds__(TS_THREADS) tile_sort_reg_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ list,
                                                                   uint32_t* __restrict__ sorted_ids, int n_lo, int n_mid, int n_hi,
                                                                   const uint32_t* __restrict__ status) {
    extern __shared__ uint64_t s_w[];
    if (status && status[2]) return;
    const uint2 rg = ranges[blockIdx.x];
    const int n = (int)(rg.y - rg.x);
    if (n <= n_lo || n > n_hi) return;
    if (n <= n_mid) tile_sort_reg_body<A>(rg, n, list, sorted_ids, s_w);
    else tile_sort_reg_body<B>(rg, n, list, sorted_ids, s_w);
}

// ---------------------------------------------------------------------------------------------------------------
// Blend: one CTA (256 threads) per 16x16 tile, front-to-back over the tile's sorted list in batches of 256 records that
// are staged once in shared memory (id, xy, conic+opacity, rgb+depth = 44 B) and then broadcast-read by the pixels.
//
// Sub-tile culling.  A warp owns an 8x4 pixel block of the tile.  A record can only change a pixel where
//     power = -(a dx^2 + c dy^2)/2 - b dx dy  satisfies  power <= 0  and  o * exp(power) >= 1/255,
// i.e. inside the ellipse  q(d) <= tau = ln(255 o).  The staging thread turns that into a conservative bounding box
// (half extents sqrt(2 tau Sigma_xx), sqrt(2 tau Sigma_yy) with Sigma = conic^-1, widened by 0.02 in tau, 0.1 % and
// 0.01 px against rounding of the per-pixel arithmetic and of ex2.approx); every warp then compacts, in list order, the
// records whose box meets its block and blends only those.  Skipped records would have taken the `power > 0` or
// `alpha < 1/255` exit at all 32 pixels, so colour, depth, opacity and n_touched are bit-identical to the unculled loop,
// while a pixel-aligned splat (1-3 px) is evaluated by 1-3 warps instead of 8.
// ---------------------------------------------------------------------------------------------------------------
constexpr int RB = TILE_X * TILE_Y;   // records per batch = threads per CTA
constexpr int SUB_W = 8, SUB_H = 4;   // pixel block of one warp

template <bool COUNT_TOUCHED>
__global__ void __launch_bounds__(RB) render_kernel(int W, int H, int gx, const uint2* __restrict__ ranges,
                                                    const uint32_t* __restrict__ point_list, const float2* __restrict__ xy,
                                                    const float4* __restrict__ conic_o, const float* __restrict__ rgb,
                                                    const float* __restrict__ depths, const float* __restrict__ cam_dev,
                                                    float* __restrict__ out_color, float* __restrict__ out_depth,
                                                    float* __restrict__ out_opacity, int32_t* __restrict__ n_touched, int cull,
                                                    const uint32_t* __restrict__ status) {
    if (status && status[2]) return;   // no-sync protocol: capacity overflow flagged by tile_scan_kernel (the caller retries with more room)
    __shared__ uint32_t s_id[RB];
    __shared__ float2 s_xy[RB];
    __shared__ float4 s_co[RB];
    __shared__ float4 s_rgbd[RB];
    __shared__ float4 s_box[RB];                 // xmin, xmax, ymin, ymax of the contributing region
    __shared__ uint8_t s_list[RB / 32][RB];      // per warp: batch slots that meet its pixel block, in list order
    __shared__ int s_cnt[COUNT_TOUCHED ? RB : 1];

    const int tile_x = blockIdx.x, tile_y = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bx0 = tile_x * TILE_X + (warp & 1) * SUB_W, by0 = tile_y * TILE_Y + (warp >> 1) * SUB_H;
    const int pxi = bx0 + (lane & 7), pyi = by0 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const float pfx = (float)pxi, pfy = (float)pyi;
    const float wx0 = (float)bx0, wx1 = (float)(bx0 + SUB_W - 1), wy0 = (float)by0, wy1 = (float)(by0 + SUB_H - 1);
    const uint2 range = ranges[tile_y * gx + tile_x];
    const int rounds = (int)((range.y - range.x + RB - 1) / RB);
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dz = 0.f;
    uint8_t* my_list = s_list[warp];

    for (int r = 0; r < rounds; ++r, todo -= RB) {
        const int num_done = __syncthreads_count(done);
        if (num_done == RB) break;
        const uint32_t progress = range.x + (uint32_t)r * RB + tid;
        if (progress < range.y) {
            const uint32_t id = point_list[progress];
            const float2 p = xy[id];
            const float4 co = conic_o[id];
            s_id[tid] = id;
            s_xy[tid] = p;
            s_co[tid] = co;
            s_rgbd[tid] = make_float4(rgb[3 * (size_t)id], rgb[3 * (size_t)id + 1], rgb[3 * (size_t)id + 2], depths[id]);
            const float det = co.x * co.z - co.y * co.y;
            float ex = INFINITY, ey = INFINITY;   // not a proper ellipse: never cull
            if (cull && det > 0.0f && co.x > 0.0f && co.z > 0.0f && co.w == co.w) {   // NaN opacity: fminf(0.99, NaN) blends -> keep
                const float tau = __logf(255.0f * co.w) + 0.02f;
                if (tau < 0.0f || !(co.w > 0.0f)) {
                    ex = ey = -INFINITY;          // o < 1/255: alpha can never reach 1/255 -> empty box
                } else {
                    const float inv = 1.0f / det;
                    ex = sqrtf(2.0f * tau * co.z * inv) * 1.001f + 0.01f;
                    ey = sqrtf(2.0f * tau * co.x * inv) * 1.001f + 0.01f;
                }
            }
            s_box[tid] = make_float4(p.x - ex, p.x + ex, p.y - ey, p.y + ey);
        }
        if (COUNT_TOUCHED) s_cnt[tid] = 0;
        __syncthreads();
        const int nb = min(RB, todo);
        // ---- per-warp compaction of the batch (order preserving) ----
        int nl = 0;
        if (!__all_sync(0xffffffffu, done)) {
#pragma unroll
            for (int k = 0; k < RB / 32; ++k) {
                const int j = k * 32 + lane;
                bool hit = false;
                if (j < nb) {
                    const float4 b = s_box[j];
                    hit = b.y >= wx0 && b.x <= wx1 && b.w >= wy0 && b.z <= wy1;
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) my_list[nl + __popc(m & ((1u << lane) - 1u))] = (uint8_t)j;
                nl += __popc(m);
            }
            __syncwarp();
        }
        for (int i = 0; i < nl; ++i) {
            const int j = my_list[i];
            bool touch = false;
            if (!done) {
                const float2 p = s_xy[j];
                const float4 co = s_co[j];
                const float dx = p.x - pfx, dy = p.y - pfy;
                const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                if (power <= 0.0f) {
                    const float alpha = fminf(0.99f, co.w * __expf(power));
                    if (alpha >= 1.0f / 255.0f) {
                        const float test_T = T * (1.0f - alpha);
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
                            const float4 cd = s_rgbd[j];
                            const float w = alpha * T;
                            C0 += cd.x * w;
                            C1 += cd.y * w;
                            C2 += cd.z * w;
                            Dz += cd.w * w;
                            touch = test_T > 0.5f;
                            T = test_T;
                        }
                    }
                }
            }
            if (COUNT_TOUCHED) {
                const unsigned m = __ballot_sync(0xffffffffu, touch);
                if (m != 0 && lane == 0) atomicAdd(&s_cnt[j], __popc(m));
            }
        }
        if (COUNT_TOUCHED) {
            __syncthreads();
            if (progress < range.y && s_cnt[tid] > 0) atomicAdd(&n_touched[s_id[tid]], s_cnt[tid]);
        }
    }
    if (inside) {
        const size_t pid = (size_t)pyi * W + pxi;
        const size_t hw = (size_t)H * W;
        const float* bg = cam_dev + 35;  // Camera::bg
        out_color[pid] = C0 + T * bg[0];
        out_color[hw + pid] = C1 + T * bg[1];
        out_color[2 * hw + pid] = C2 + T * bg[2];
        out_depth[pid] = Dz;
        out_opacity[pid] = 1.0f - T;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// N-channel ("feature") splatting: what the reference obtains from gsplat.rasterization(means, covars, opacities, colors[N,C],
// viewmats, Ks, width, height, near_plane, far_plane) at src/models/gaussian_renderer.py:92-106 to render the per-Gaussian
// query x class logits.  gsplat (1.5.2, un-vendored) is NOT available offline: this follows its published classic-mode algorithm
// (projection with the 0.3-px blur, opacity-aware 3.33-sigma extents, 16x16 tiles, alpha = min(0.999, o exp(-sigma)), alpha < 1/255
// skipped, stop before T <= 1e-4, pixel centres at +0.5) -- PARITY UNPINNED, see oracle/gsplat_ref.py.  Binning, scan, sort and tile
// ranges are shared with the colour rasterizer above.
// ---------------------------------------------------------------------------------------------------------------
struct FeatCam { float V[16]; float fx, fy, cx, cy, near_plane, far_plane; };

__global__ void __launch_bounds__(128) preprocess_feat_kernel(int G, int H, int W, int gx, int gy, int cov_stride, const float* __restrict__ means,
                                                              const float* __restrict__ cov, const float* __restrict__ opac,
                                                              const float* __restrict__ cam_dev, float* __restrict__ depths,
                                                              float2* __restrict__ xy, float4* __restrict__ conic_o, uint32_t* __restrict__ tiles,
                                                              ushort4* __restrict__ rects, int32_t* __restrict__ radii, int32_t* __restrict__ radii_xy) {
    __shared__ FeatCam s_cam;
    if (threadIdx.x < (int)(sizeof(FeatCam) / sizeof(float))) reinterpret_cast<float*>(&s_cam)[threadIdx.x] = cam_dev[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G) return;
    radii[i] = 0;               // culled Gaussians keep these zeros
    tiles[i] = 0;
    if (radii_xy) { radii_xy[2 * i] = 0; radii_xy[2 * i + 1] = 0; }
    const float* V = s_cam.V;   // world-to-camera, row-major
    const float px = means[3 * (size_t)i], py = means[3 * (size_t)i + 1], pz = means[3 * (size_t)i + 2];
    const float x = ADD(ADD(ADD(MUL(V[0], px), MUL(V[1], py)), MUL(V[2], pz)), V[3]);
    const float y = ADD(ADD(ADD(MUL(V[4], px), MUL(V[5], py)), MUL(V[6], pz)), V[7]);
    const float z = ADD(ADD(ADD(MUL(V[8], px), MUL(V[9], py)), MUL(V[10], pz)), V[11]);
    if (z < s_cam.near_plane || z > s_cam.far_plane) return;
    float S[3][3];
    {
        const float* c = cov + (size_t)i * cov_stride;
        if (cov_stride == 6) { S[0][0] = c[0]; S[0][1] = S[1][0] = c[1]; S[0][2] = S[2][0] = c[2]; S[1][1] = c[3]; S[1][2] = S[2][1] = c[4]; S[2][2] = c[5]; }
        else { for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) S[r][q] = c[3 * r + q]; }
    }
    float RS[3][3], Cc[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) RS[r][q] = ADD(ADD(MUL(V[4 * r], S[0][q]), MUL(V[4 * r + 1], S[1][q])), MUL(V[4 * r + 2], S[2][q]));
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) Cc[r][q] = ADD(ADD(MUL(RS[r][0], V[4 * q]), MUL(RS[r][1], V[4 * q + 1])), MUL(RS[r][2], V[4 * q + 2]));
    const float fx = s_cam.fx, fy = s_cam.fy, cx = s_cam.cx, cy = s_cam.cy;
    const float tanx = DIV(MUL(0.5f, (float)W), fx), tany = DIV(MUL(0.5f, (float)H), fy);
    const float lxp = ADD(DIV(SUB((float)W, cx), fx), MUL(0.3f, tanx)), lxn = ADD(DIV(cx, fx), MUL(0.3f, tanx));
    const float lyp = ADD(DIV(SUB((float)H, cy), fy), MUL(0.3f, tany)), lyn = ADD(DIV(cy, fy), MUL(0.3f, tany));
    const float rz = DIV(1.0f, z), rz2 = MUL(rz, rz);
    const float tx = MUL(z, fminf(lxp, fmaxf(-lxn, MUL(x, rz)))), ty = MUL(z, fminf(lyp, fmaxf(-lyn, MUL(y, rz))));
    const float ja = MUL(fx, rz), jb = -MUL(MUL(fx, tx), rz2), jc = MUL(fy, rz), jd = -MUL(MUL(fy, ty), rz2);
    float JC0[3], JC1[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) { JC0[q] = ADD(MUL(ja, Cc[0][q]), MUL(jb, Cc[2][q])); JC1[q] = ADD(MUL(jc, Cc[1][q]), MUL(jd, Cc[2][q])); }
    float c00 = ADD(MUL(JC0[0], ja), MUL(JC0[2], jb));
    const float c01 = ADD(MUL(JC0[1], jc), MUL(JC0[2], jd));
    float c11 = ADD(MUL(JC1[1], jc), MUL(JC1[2], jd));
    const float mx = ADD(MUL(MUL(fx, x), rz), cx), my = ADD(MUL(MUL(fy, y), rz), cy);
    c00 = ADD(c00, 0.3f); c11 = ADD(c11, 0.3f);
    const float det = SUB(MUL(c00, c11), MUL(c01, c01));
    if (!(det > 0.0f)) return;
    const float o = opac[i];
    const float thr = 1.0f / 255.0f;
    if (o < thr) return;
    const float ext = fminf(3.33f, __fsqrt_rn(MUL(2.0f, logf(DIV(o, thr)))));
    const float bb = MUL(0.5f, ADD(c00, c11));
    const float v1 = ADD(bb, __fsqrt_rn(fmaxf(0.01f, SUB(MUL(bb, bb), det))));
    const float r1 = MUL(ext, __fsqrt_rn(v1));
    const float rx = ceilf(fminf(MUL(ext, __fsqrt_rn(c00)), r1)), ry = ceilf(fminf(MUL(ext, __fsqrt_rn(c11)), r1));
    if (rx <= 0.0f && ry <= 0.0f) return;
    if (ADD(mx, rx) <= 0.0f || SUB(mx, rx) >= (float)W || ADD(my, ry) <= 0.0f || SUB(my, ry) >= (float)H) return;
    const float tsx = DIV(mx, 16.0f), tsy = DIV(my, 16.0f), trx = DIV(rx, 16.0f), try_ = DIV(ry, 16.0f);
    const int rminx = min(gx, max(0, (int)floorf(SUB(tsx, trx)))), rmaxx = min(gx, max(0, (int)ceilf(ADD(tsx, trx))));
    const int rminy = min(gy, max(0, (int)floorf(SUB(tsy, try_)))), rmaxy = min(gy, max(0, (int)ceilf(ADD(tsy, try_))));
    const int nt = (rmaxx - rminx) * (rmaxy - rminy);
    if (nt <= 0) return;
    const float inv = DIV(1.0f, det);
    depths[i] = z;
    xy[i] = make_float2(mx, my);
    conic_o[i] = make_float4(MUL(c11, inv), MUL(-c01, inv), MUL(c00, inv), o);
    tiles[i] = (uint32_t)nt;
    rects[i] = make_ushort4((unsigned short)rminx, (unsigned short)rminy, (unsigned short)rmaxx, (unsigned short)rmaxy);
    radii[i] = (int32_t)fmaxf(rx, ry);
    if (radii_xy) { radii_xy[2 * (size_t)i] = (int32_t)rx; radii_xy[2 * (size_t)i + 1] = (int32_t)ry; }
}

// Blend of one 32-channel slice of the features: grid (tiles_x, tiles_y, ceil(C / 32)); same 256-record batches and exact per-warp
// (8x4 pixel) culling as render_kernel; the slice of every staged record's feature row sits in shared memory (32 KB per batch).
constexpr int FCH = 32;
__global__ void __launch_bounds__(RB) render_feat_kernel(int W, int H, int gx, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                                         const float2* __restrict__ xy, const float4* __restrict__ conic_o,
                                                         const float* __restrict__ feats, int C, float* __restrict__ out, float* __restrict__ out_alpha,
                                                         const uint32_t* __restrict__ status) {
    if (status && status[2]) return;   // no-sync protocol: the frame overflowed, the host re-renders it
    __shared__ float2 s_xy[RB];
    __shared__ float4 s_co[RB];
    __shared__ float4 s_box[RB];
    __shared__ uint8_t s_list[RB / 32][RB];
    __shared__ __align__(16) float s_f[RB][FCH];
    const int tile_x = blockIdx.x, tile_y = blockIdx.y, c0 = blockIdx.z * FCH;
    const int nch = min(FCH, C - c0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bx0 = tile_x * TILE_X + (warp & 1) * SUB_W, by0 = tile_y * TILE_Y + (warp >> 1) * SUB_H;
    const int pxi = bx0 + (lane & 7), pyi = by0 + (lane >> 3);
    const bool inside = pxi < W && pyi < H;
    const float pfx = (float)pxi + 0.5f, pfy = (float)pyi + 0.5f;
    const float wx0 = (float)bx0 + 0.5f, wx1 = (float)(bx0 + SUB_W - 1) + 0.5f, wy0 = (float)by0 + 0.5f, wy1 = (float)(by0 + SUB_H - 1) + 0.5f;
    const uint2 range = ranges[tile_y * gx + tile_x];
    const int rounds = (int)((range.y - range.x + RB - 1) / RB);
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    float T = 1.0f;
    float acc[FCH];
#pragma unroll
    for (int k = 0; k < FCH; ++k) acc[k] = 0.f;
    uint8_t* my_list = s_list[warp];
    const bool vec = (C % 4 == 0) && nch == FCH && ((((uintptr_t)feats) & 15) == 0);

    for (int r = 0; r < rounds; ++r, todo -= RB) {
        const int num_done = __syncthreads_count(done);
        if (num_done == RB) break;
        const uint32_t progress = range.x + (uint32_t)r * RB + tid;
        if (progress < range.y) {
            const uint32_t id = point_list[progress];
            const float2 p = xy[id];
            const float4 co = conic_o[id];
            s_xy[tid] = p;
            s_co[tid] = co;
            const float* f = feats + (size_t)id * C + c0;
            if (vec) {
#pragma unroll
                for (int k = 0; k < FCH; k += 4) *reinterpret_cast<float4*>(&s_f[tid][k]) = __ldg(reinterpret_cast<const float4*>(f + k));
            } else {
                for (int k = 0; k < FCH; ++k) s_f[tid][k] = k < nch ? __ldg(f + k) : 0.f;
            }
            const float det = co.x * co.z - co.y * co.y;
            float ex = INFINITY, ey = INFINITY;
            if (det > 0.0f && co.x > 0.0f && co.z > 0.0f && co.w == co.w) {
                const float tau = __logf(255.0f * co.w) + 0.02f;
                if (tau < 0.0f || !(co.w > 0.0f)) {
                    ex = ey = -INFINITY;
                } else {
                    const float inv = 1.0f / det;
                    ex = sqrtf(2.0f * tau * co.z * inv) * 1.001f + 0.01f;
                    ey = sqrtf(2.0f * tau * co.x * inv) * 1.001f + 0.01f;
                }
            }
            s_box[tid] = make_float4(p.x - ex, p.x + ex, p.y - ey, p.y + ey);
        }
        __syncthreads();
        const int nb = min(RB, todo);
        int nl = 0;
        if (!__all_sync(0xffffffffu, done)) {
#pragma unroll
            for (int k = 0; k < RB / 32; ++k) {
                const int j = k * 32 + lane;
                bool hit = false;
                if (j < nb) {
                    const float4 b = s_box[j];
                    hit = b.y >= wx0 && b.x <= wx1 && b.w >= wy0 && b.z <= wy1;
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) my_list[nl + __popc(m & ((1u << lane) - 1u))] = (uint8_t)j;
                nl += __popc(m);
            }
            __syncwarp();
        }
        for (int i = 0; i < nl; ++i) {
            const int j = my_list[i];
            if (done) continue;
            const float2 p = s_xy[j];
            const float4 co = s_co[j];
            const float dx = p.x - pfx, dy = p.y - pfy;
            const float sigma = 0.5f * (co.x * dx * dx + co.z * dy * dy) + co.y * dx * dy;
            const float alpha = fminf(0.999f, co.w * __expf(-sigma));
            if (sigma < 0.0f || alpha < 1.0f / 255.0f) continue;
            const float next_T = T * (1.0f - alpha);
            if (next_T <= 1e-4f) { done = true; continue; }
            const float vis = alpha * T;
            const float4* f4 = reinterpret_cast<const float4*>(s_f[j]);
#pragma unroll
            for (int k = 0; k < FCH / 4; ++k) {
                const float4 f = f4[k];
                acc[4 * k] += f.x * vis; acc[4 * k + 1] += f.y * vis; acc[4 * k + 2] += f.z * vis; acc[4 * k + 3] += f.w * vis;
            }
            T = next_T;
        }
    }
    if (inside) {
        float* o = out + ((size_t)pyi * W + pxi) * C + c0;
        if (vec) {
#pragma unroll
            for (int k = 0; k < FCH; k += 4) *reinterpret_cast<float4*>(o + k) = make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < FCH; ++k)
                if (k < nch) o[k] = acc[k];
        }
        if (blockIdx.z == 0 && out_alpha) out_alpha[(size_t)pyi * W + pxi] = 1.0f - T;
    }
}

// camera block of the feature rasterizer: world-to-camera matrix (device) + six scalars passed by value
__global__ void pack_feat_camera_kernel(const float* __restrict__ viewmat, float fx, float fy, float cx, float cy, float near_plane, float far_plane,
                                        float* __restrict__ cam) {
    const int t = threadIdx.x;
    if (t < 16) cam[t] = viewmat[t];
    if (t == 16) cam[16] = fx;
    if (t == 17) cam[17] = fy;
    if (t == 18) cam[18] = cx;
    if (t == 19) cam[19] = cy;
    if (t == 20) cam[20] = near_plane;
    if (t == 21) cam[21] = far_plane;
}

__global__ void pack_camera_kernel(const float* __restrict__ view, const float* __restrict__ proj, const float* __restrict__ campos,
                                   const float* __restrict__ bg, float* __restrict__ cam) {
    const int t = threadIdx.x;
    if (t < 16) cam[t] = view[t];
    else if (t < 32) cam[t] = proj[t - 16];
    else if (t < 35) cam[t] = campos[t - 32];
    else if (t < 38) cam[t] = bg[t - 35];
}

struct Workspace {
    float* depths; float2* xy; float4* conic_o; float* rgb; uint32_t* tiles; ushort4* rects; uint32_t* offsets;
    uint32_t* block_sums; uint32_t* total; float* cam; uint2* ranges; uint32_t* tile_counts; uint32_t* tile_cursors;
    uint64_t* keys; uint64_t* keys_sorted; uint32_t* vals; uint32_t* vals_sorted; void* cub_temp; size_t cub_bytes;
    size_t bytes;
};

int higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return (int)msb;
}

size_t cub_temp_bytes(int64_t cap, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)cap, 0, end_bit);
    return bytes;
}

Workspace carve(void* base, int G, int H, int W, int64_t cap) {
    Workspace w{};
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t n) { void* p = b ? b + off : nullptr; off = align_up(off + n, 256); return p; };
    const int gx = ceil_div(W, TILE_X), gy = ceil_div(H, TILE_Y);
    w.depths = (float*)take(sizeof(float) * G);
    w.xy = (float2*)take(sizeof(float2) * G);
    w.conic_o = (float4*)take(sizeof(float4) * G);
    w.rgb = (float*)take(sizeof(float) * 3 * G);
    w.tiles = (uint32_t*)take(sizeof(uint32_t) * G);
    w.rects = (ushort4*)take(sizeof(ushort4) * G);
    w.offsets = (uint32_t*)take(sizeof(uint32_t) * G);
    w.block_sums = (uint32_t*)take(sizeof(uint32_t) * (ceil_div(G, SCAN_TILE) + 1));
    w.total = (uint32_t*)take(256);
    w.cam = (float*)take(256);
    w.ranges = (uint2*)take(sizeof(uint2) * gx * gy);
    w.tile_counts = (uint32_t*)take(sizeof(uint32_t) * gx * gy);
    w.tile_cursors = (uint32_t*)take(sizeof(uint32_t) * gx * gy);
    w.keys = (uint64_t*)take(sizeof(uint64_t) * cap);
    w.keys_sorted = (uint64_t*)take(sizeof(uint64_t) * cap);
    w.vals = (uint32_t*)take(sizeof(uint32_t) * cap);
    w.vals_sorted = (uint32_t*)take(sizeof(uint32_t) * cap);
    w.cub_bytes = cub_temp_bytes(cap, 32 + higher_msb((uint32_t)(gx * gy)));
    w.cub_temp = take(w.cub_bytes);
    w.bytes = off;
    return w;
}

int g_regsort = 1; // testing aid: 0 = the round-1 shared-memory bitonic tile sort (tiles <= 2048 records) instead of the register-resident one

// Per-tile sorts of one frame: one launch per size class that can occur (max_tile = largest tile if the host knows it, else TS_CAP).
int launch_tile_sorts(int ntiles, const uint2* ranges, const uint64_t* list, uint32_t* sorted_ids, int max_tile, const uint32_t* status,
                      cudaStream_t stream) {
    static bool attr[64] = {false};
    int dev = 0;
    SIU3R_CUDA_CHECK(cudaGetDevice(&dev));
    if (!attr[dev & 63]) {
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(tile_sort_reg_kernel<32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_THREADS * 32 * 8));
        attr[dev & 63] = true;
    }
    // size classes: 2 / 8 / 16 / 32 keys per thread = up to 512 / 2048 / 4096 / 8192 records (a 1 M-Gaussian 512^2 frame averages 2170 records per tile,
    // which a 32-key-only upper class sorted as 8192)
    tile_sort_reg_kernel<2, 8><<<ntiles, TS_THREADS, TS_THREADS * 8 * 8, stream>>>(ranges, list, sorted_ids, 0, 512, 2048, status);
    siu3r_note_launch(1);
    if (max_tile > 2048) { tile_sort_reg_kernel<16, 16><<<ntiles, TS_THREADS, TS_THREADS * 16 * 8, stream>>>(ranges, list, sorted_ids, 2048, 4096, 4096, status); siu3r_note_launch(1); }
    if (max_tile > 4096) { tile_sort_reg_kernel<32, 32><<<ntiles, TS_THREADS, TS_THREADS * 32 * 8, stream>>>(ranges, list, sorted_ids, 4096, TS_CAP, TS_CAP, status); siu3r_note_launch(1); }
    SIU3R_LAUNCH_CHECK();
    return SIU3R_OK;
}

int g_binned = 1; // 0: always use the reference-shaped global radix sort (testing aid, see siu3r_raster_set_binning)
int g_cull = 1;   // testing aid: 0 blends every record of the tile at every pixel (the unculled reference loop)

}  // namespace

// The sizing entry points read the duplicate count back (cudaStreamSynchronize): not possible on a capturing stream -> use the *_nosync variants.
static bool stream_is_capturing(cudaStream_t stream) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) { cudaGetLastError(); return false; }
    return st != cudaStreamCaptureStatusNone;
}

extern "C" {

void siu3r_raster_set_culling(int enabled) { g_cull = enabled ? 1 : 0; }
void siu3r_raster_set_binning(int enabled) { g_binned = enabled ? 1 : 0; }
void siu3r_raster_set_regsort(int enabled) { g_regsort = enabled ? 1 : 0; }

// Bytes of device scratch siu3r_raster_forward needs for G Gaussians, an HxW image and at most `dup_capacity`
// (tile, Gaussian) duplicates.  Mirrors the resize-callback buffers of the reference rasterizer (geomBuffer,
// binningBuffer, imageBuffer): here the caller (PyTorch) owns the allocation.
int64_t siu3r_raster_workspace_bytes(int G, int H, int W, int64_t dup_capacity) {
    if (G <= 0 || H <= 0 || W <= 0 || dup_capacity <= 0) return SIU3R_ERR_INVALID;
    if (dup_capacity >= (1ll << 31)) return SIU3R_ERR_INVALID;
    Workspace w = carve(nullptr, G, H, W, dup_capacity);
    return (int64_t)w.bytes;
}

// Forward rasterization of one camera.  All pointers are device pointers unless noted.
//   means3D [G,3]; cov: [G,6] (xx,xy,xz,yy,yz,zz; cov_stride=6) or [G,3,3] (cov_stride=9);
//   shs: sh_layout 0 = [G,sh_coeffs,3] (what render_cuda passes), 1 = [G,3,sh_coeffs] (Gaussians.harmonics as stored);
//   opacities [G]; viewmatrix/projmatrix [16] as torch lays out view_matrix[i] / full_projection[i]
//   (cuda_splatting.py:74-77); campos [3]; bg [3].
// Outputs: out_color [3,H,W], out_depth [H,W], out_opacity [H,W], radii [G] int32, n_touched [G] int32 (may be null).
// debug_* (may be null): tiles_touched [G], offsets [G], sorted keys/values [>= D], ranges [tiles*2].
// num_rendered_host (host pointer, may be null) receives D.  Synchronises the stream once (to read D), like the
// reference implementation does.
int siu3r_raster_forward(int G, int H, int W, int sh_degree, int sh_coeffs, int sh_layout, int cov_stride,
                         const float* means3D, const float* cov, const float* shs, const float* opacities,
                         const float* viewmatrix, const float* projmatrix, const float* campos, const float* bg,
                         float tan_fovx, float tan_fovy, float* out_color, float* out_depth, float* out_opacity,
                         int32_t* radii, int32_t* n_touched, void* workspace, int64_t workspace_bytes,
                         int64_t dup_capacity, int64_t* num_rendered_host, uint32_t* debug_tiles_touched,
                         uint32_t* debug_offsets, uint64_t* debug_keys, uint32_t* debug_values, uint32_t* debug_ranges,
                         void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(G > 0 && H > 0 && W > 0);
    SIU3R_REQUIRE(sh_coeffs >= 1 && sh_coeffs * 3 <= MAX_SH_FLOATS);
    SIU3R_REQUIRE((sh_degree + 1) * (sh_degree + 1) <= sh_coeffs);
    SIU3R_REQUIRE(sh_layout == 0 || sh_layout == 1);
    SIU3R_REQUIRE(cov_stride == 6 || cov_stride == 9);
    SIU3R_REQUIRE(means3D && cov && shs && opacities && viewmatrix && projmatrix && campos && bg);
    SIU3R_REQUIRE(out_color && out_depth && out_opacity && radii && workspace);
    SIU3R_REQUIRE(dup_capacity > 0 && dup_capacity < (1ll << 31));
    if (stream_is_capturing(stream)) return SIU3R_ERR_UNSUPPORTED;
    const int gx = ceil_div(W, TILE_X), gy = ceil_div(H, TILE_Y);
    SIU3R_REQUIRE(gx < 65536 && gy < 65536);
    Workspace w = carve(workspace, G, H, W, dup_capacity);
    if ((int64_t)w.bytes > workspace_bytes) return SIU3R_ERR_CAPACITY;
    const int deg = sh_degree > 3 ? 3 : sh_degree;  // the reference kernel evaluates SH bands 0..3 only
    const float focal_y = (float)H / (2.0f * tan_fovy), focal_x = (float)W / (2.0f * tan_fovx);

    // The binned fast path produces the same sorted list and ranges; the debug exports of the reference pipeline's intermediate
    // arrays (per-Gaussian offsets, 64-bit keys) only exist on the global-sort path.
    const bool want_debug = debug_tiles_touched || debug_offsets || debug_keys || debug_values || debug_ranges;
    bool binned = g_binned && !want_debug;

    SIU3R_CUDA_CHECK(cudaMemsetAsync(w.ranges, 0, sizeof(uint2) * gx * gy, stream));
    if (binned) SIU3R_CUDA_CHECK(cudaMemsetAsync(w.tile_counts, 0, sizeof(uint32_t) * gx * gy, stream));
    if (n_touched) SIU3R_CUDA_CHECK(cudaMemsetAsync(n_touched, 0, sizeof(int32_t) * G, stream));

    pack_camera_kernel<<<1, 64, 0, stream>>>(viewmatrix, projmatrix, campos, bg, w.cam);
    PreOut po{w.depths, w.xy, w.conic_o, w.rgb, w.tiles, w.rects};
    preprocess_kernel<<<ceil_div(G, PRE_THREADS), PRE_THREADS, 0, stream>>>(G, H, W, gx, gy, deg, sh_coeffs, sh_layout, cov_stride,
                                                                            means3D, cov, shs, opacities, w.cam, tan_fovx, tan_fovy,
                                                                            focal_x, focal_y, po, radii);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(2);
    const uint64_t* sorted_keys = w.keys_sorted;
    const uint32_t* sorted_vals = w.vals_sorted;
    int64_t D = 0;
    const int ntiles = gx * gy;
    if (binned && (size_t)ntiles * 8 > 200 * 1024) binned = false;        // per-CTA tile histograms must fit shared memory (<= 25600 tiles)
    if (binned) {
        static bool attr_bin[64] = {false};
        if (siu3r_first_use_on_device(attr_bin)) {
            SIU3R_CUDA_CHECK(cudaFuncSetAttribute(bin_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            SIU3R_CUDA_CHECK(cudaFuncSetAttribute(bin_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        bin_count_kernel<<<ceil_div(G, BIN_CHUNK), BIN_THREADS, (size_t)ntiles * 4, stream>>>(G, gx, ntiles, radii, w.rects, w.tile_counts);
        tile_scan_kernel<<<1, 1024, 0, stream>>>(w.tile_counts, ntiles, w.ranges, w.tile_cursors, w.total, (uint32_t)dup_capacity, nullptr);
        SIU3R_LAUNCH_CHECK();
        siu3r_note_launch(2);
        uint32_t tm[2] = {0, 0};
        SIU3R_CUDA_CHECK(cudaMemcpyAsync(tm, w.total, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        SIU3R_CUDA_CHECK(cudaStreamSynchronize(stream));
        D = (int64_t)tm[0];
        if (num_rendered_host) *num_rendered_host = D;
        if (D > dup_capacity) return SIU3R_ERR_CAPACITY;
        if (tm[1] > (uint32_t)(g_regsort ? TS_CAP : TS_FAST)) {
            binned = false;                                   // big tiles: the O(n log^2 n) shared-memory sort loses to the global radix sort
            SIU3R_CUDA_CHECK(cudaMemsetAsync(w.ranges, 0, sizeof(uint2) * gx * gy, stream));
        } else if (D > 0) {
            static bool attr[64] = {false};
            if (siu3r_first_use_on_device(attr)) SIU3R_CUDA_CHECK(cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_CAP * 8));
            bin_scatter_kernel<<<ceil_div(G, BIN_CHUNK), BIN_THREADS, (size_t)ntiles * 8, stream>>>(G, gx, ntiles, radii, w.depths, w.rects,
                                                                                                  w.tile_cursors, w.keys, nullptr);
            if (g_regsort) {
                int r = launch_tile_sorts(gx * gy, w.ranges, w.keys, w.vals_sorted, (int)tm[1], nullptr, stream); if (r) return r;
            } else {
                int P = 1;
                while (P < (int)tm[1]) P <<= 1;                   // shared memory for the largest tile of this frame
                tile_sort_kernel<<<gx * gy, TS_THREADS, (size_t)P * 8, stream>>>(w.ranges, w.keys, w.vals_sorted, nullptr);
                SIU3R_LAUNCH_CHECK();
                siu3r_note_launch(1);
            }
            siu3r_note_launch(1);
        }
    }
    if (!binned) {
        const int nsb = ceil_div(G, SCAN_TILE);
        scan_local_kernel<<<nsb, SCAN_THREADS, 0, stream>>>(w.tiles, w.offsets, w.block_sums, G);
        scan_sums_kernel<<<1, SCAN_THREADS, 0, stream>>>(w.block_sums, nsb, w.total);
        scan_add_kernel<<<nsb, SCAN_THREADS, 0, stream>>>(w.offsets, w.block_sums, G);
        SIU3R_LAUNCH_CHECK();
        siu3r_note_launch(3);
        uint32_t D32 = 0;
        SIU3R_CUDA_CHECK(cudaMemcpyAsync(&D32, w.total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        SIU3R_CUDA_CHECK(cudaStreamSynchronize(stream));
        D = (int64_t)D32;
        if (num_rendered_host) *num_rendered_host = D;
        if (debug_tiles_touched) SIU3R_CUDA_CHECK(cudaMemcpyAsync(debug_tiles_touched, w.tiles, sizeof(uint32_t) * G, cudaMemcpyDeviceToDevice, stream));
        if (debug_offsets) SIU3R_CUDA_CHECK(cudaMemcpyAsync(debug_offsets, w.offsets, sizeof(uint32_t) * G, cudaMemcpyDeviceToDevice, stream));
        if (D > dup_capacity) return SIU3R_ERR_CAPACITY;
        if (D > 0) {
            duplicate_with_keys_kernel<<<ceil_div(G, 256), 256, 0, stream>>>(G, gx, radii, w.offsets, w.depths, w.rects, w.keys, w.vals);
            const int end_bit = 32 + higher_msb((uint32_t)(gx * gy));
            size_t tb = w.cub_bytes;
            SIU3R_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.keys, w.keys_sorted, w.vals, w.vals_sorted, (int)D, 0,
                                                             end_bit, stream));
            identify_tile_ranges_kernel<<<(unsigned)ceil_div_i64(D, 256), 256, 0, stream>>>((uint32_t)D, sorted_keys, w.ranges);
            SIU3R_LAUNCH_CHECK();
            siu3r_note_launch(2);
        }
    }
    dim3 grid(gx, gy), block(RB);
    if (n_touched)
        render_kernel<true><<<grid, block, 0, stream>>>(W, H, gx, w.ranges, sorted_vals, w.xy, w.conic_o, w.rgb, w.depths, w.cam,
                                                        out_color, out_depth, out_opacity, n_touched, g_cull, nullptr);
    else
        render_kernel<false><<<grid, block, 0, stream>>>(W, H, gx, w.ranges, sorted_vals, w.xy, w.conic_o, w.rgb, w.depths, w.cam,
                                                         out_color, out_depth, out_opacity, nullptr, g_cull, nullptr);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    if (D > 0) {
        if (debug_keys) SIU3R_CUDA_CHECK(cudaMemcpyAsync(debug_keys, sorted_keys, sizeof(uint64_t) * D, cudaMemcpyDeviceToDevice, stream));
        if (debug_values) SIU3R_CUDA_CHECK(cudaMemcpyAsync(debug_values, sorted_vals, sizeof(uint32_t) * D, cudaMemcpyDeviceToDevice, stream));
    }
    if (debug_ranges) SIU3R_CUDA_CHECK(cudaMemcpyAsync(debug_ranges, w.ranges, sizeof(uint2) * gx * gy, cudaMemcpyDeviceToDevice, stream));
    return SIU3R_OK;
}

// siu3r_raster_forward WITHOUT the host synchronisation (and without the debug exports): the duplicate count stays on the device.  All
// kernels of the frame are enqueued unconditionally; if the frame needs more than dup_capacity duplicates, or one tile holds more than 8192
// records, tile_scan_kernel raises a flag in status_dev[2] (bit 0 / bit 1) and the later kernels return at once, leaving the outputs undefined.
// status_dev [4] uint32 (device): {duplicates D, largest tile, flags, 0}; the caller reads it whenever convenient -- e.g. once for all the
// cameras of a SplattingCUDA.forward call -- and re-renders a flagged camera through siu3r_raster_forward.  Capturable in a CUDA graph.
int siu3r_raster_forward_nosync(int G, int H, int W, int sh_degree, int sh_coeffs, int sh_layout, int cov_stride, const float* means3D,
                                const float* cov, const float* shs, const float* opacities, const float* viewmatrix, const float* projmatrix,
                                const float* campos, const float* bg, float tan_fovx, float tan_fovy, float* out_color, float* out_depth,
                                float* out_opacity, int32_t* radii, int32_t* n_touched, void* workspace, int64_t workspace_bytes,
                                int64_t dup_capacity, uint32_t* status_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(G > 0 && H > 0 && W > 0 && status_dev);
    SIU3R_REQUIRE(sh_coeffs >= 1 && sh_coeffs * 3 <= MAX_SH_FLOATS && (sh_degree + 1) * (sh_degree + 1) <= sh_coeffs);
    SIU3R_REQUIRE((sh_layout == 0 || sh_layout == 1) && (cov_stride == 6 || cov_stride == 9));
    SIU3R_REQUIRE(means3D && cov && shs && opacities && viewmatrix && projmatrix && campos && bg);
    SIU3R_REQUIRE(out_color && out_depth && out_opacity && radii && workspace && dup_capacity > 0 && dup_capacity < (1ll << 31));
    const int gx = ceil_div(W, TILE_X), gy = ceil_div(H, TILE_Y), ntiles = gx * gy;
    SIU3R_REQUIRE(gx < 65536 && gy < 65536);
    if ((size_t)ntiles * 8 > 200 * 1024) return SIU3R_ERR_UNSUPPORTED;   // per-CTA tile histograms must fit shared memory
    Workspace w = carve(workspace, G, H, W, dup_capacity);
    if ((int64_t)w.bytes > workspace_bytes) return SIU3R_ERR_CAPACITY;
    const int deg = sh_degree > 3 ? 3 : sh_degree;
    const float focal_y = (float)H / (2.0f * tan_fovy), focal_x = (float)W / (2.0f * tan_fovx);
    static bool attr_bin[64] = {false};
    int dev = 0;
    SIU3R_CUDA_CHECK(cudaGetDevice(&dev));
    if (!attr_bin[dev & 63]) {
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(bin_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(bin_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_bin[dev & 63] = true;
    }
    SIU3R_CUDA_CHECK(cudaMemsetAsync(w.tile_counts, 0, sizeof(uint32_t) * ntiles, stream));
    if (n_touched) SIU3R_CUDA_CHECK(cudaMemsetAsync(n_touched, 0, sizeof(int32_t) * G, stream));
    pack_camera_kernel<<<1, 64, 0, stream>>>(viewmatrix, projmatrix, campos, bg, w.cam);
    PreOut po{w.depths, w.xy, w.conic_o, w.rgb, w.tiles, w.rects};
    preprocess_kernel<<<ceil_div(G, PRE_THREADS), PRE_THREADS, 0, stream>>>(G, H, W, gx, gy, deg, sh_coeffs, sh_layout, cov_stride, means3D, cov, shs,
                                                                            opacities, w.cam, tan_fovx, tan_fovy, focal_x, focal_y, po, radii);
    bin_count_kernel<<<ceil_div(G, BIN_CHUNK), BIN_THREADS, (size_t)ntiles * 4, stream>>>(G, gx, ntiles, radii, w.rects, w.tile_counts);
    tile_scan_kernel<<<1, 1024, 0, stream>>>(w.tile_counts, ntiles, w.ranges, w.tile_cursors, w.total, (uint32_t)dup_capacity, status_dev);
    bin_scatter_kernel<<<ceil_div(G, BIN_CHUNK), BIN_THREADS, (size_t)ntiles * 8, stream>>>(G, gx, ntiles, radii, w.depths, w.rects, w.tile_cursors,
                                                                                          w.keys, status_dev);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(5);
    int r = launch_tile_sorts(ntiles, w.ranges, w.keys, w.vals_sorted, TS_CAP, status_dev, stream); if (r) return r;
    dim3 grid(gx, gy), block(RB);
    if (n_touched)
        render_kernel<true><<<grid, block, 0, stream>>>(W, H, gx, w.ranges, w.vals_sorted, w.xy, w.conic_o, w.rgb, w.depths, w.cam, out_color, out_depth,
                                                        out_opacity, n_touched, g_cull, status_dev);
    else
        render_kernel<false><<<grid, block, 0, stream>>>(W, H, gx, w.ranges, w.vals_sorted, w.xy, w.conic_o, w.rgb, w.depths, w.cam, out_color, out_depth,
                                                         out_opacity, nullptr, g_cull, status_dev);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// N-channel feature rasterization of one camera (gsplat.rasterization semantics, see render_feat_kernel).  viewmat: world-to-camera
// 4x4 row-major; intr_host = (fx, fy, cx, cy) in pixels; features [G, C]; out_features [H, W, C]; out_alpha [H, W] (may be null);
// radii_xy [G, 2] int32 (may be null).  Workspace / capacity protocol as siu3r_raster_forward.
int siu3r_raster_features_forward(int G, int H, int W, int C, int cov_stride, const float* means3D, const float* cov, const float* opacities,
                                  const float* features, const float* viewmat, const float* intr_host, float near_plane, float far_plane,
                                  float* out_features, float* out_alpha, int32_t* radii_xy, void* workspace, int64_t workspace_bytes,
                                  int64_t dup_capacity, int64_t* num_rendered_host, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(G > 0 && H > 0 && W > 0 && C > 0 && (cov_stride == 6 || cov_stride == 9));
    SIU3R_REQUIRE(means3D && cov && opacities && features && viewmat && intr_host && out_features && workspace);
    SIU3R_REQUIRE(dup_capacity > 0 && dup_capacity < (1ll << 31));
    if (stream_is_capturing(stream)) return SIU3R_ERR_UNSUPPORTED;
    const int gx = ceil_div(W, TILE_X), gy = ceil_div(H, TILE_Y);
    SIU3R_REQUIRE(gx < 65536 && gy < 65536 && ceil_div(C, FCH) < 65536);
    Workspace w = carve(workspace, G, H, W, dup_capacity);
    if ((int64_t)w.bytes > workspace_bytes) return SIU3R_ERR_CAPACITY;
    int32_t* radii = reinterpret_cast<int32_t*>(w.rgb);   // the colour slot of the shared workspace is free on this path
    SIU3R_CUDA_CHECK(cudaMemsetAsync(w.ranges, 0, sizeof(uint2) * gx * gy, stream));
    // camera block: V (device) + 6 host scalars -> one small device buffer
    SIU3R_CUDA_CHECK(cudaMemcpyAsync(w.cam, viewmat, 16 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    const float scal[6] = {intr_host[0], intr_host[1], intr_host[2], intr_host[3], near_plane, far_plane};
    SIU3R_CUDA_CHECK(cudaMemcpyAsync(w.cam + 16, scal, sizeof(scal), cudaMemcpyHostToDevice, stream));
    preprocess_feat_kernel<<<ceil_div(G, 128), 128, 0, stream>>>(G, H, W, gx, gy, cov_stride, means3D, cov, opacities, w.cam, w.depths, w.xy,
                                                                w.conic_o, w.tiles, w.rects, radii, radii_xy);
    const int nsb = ceil_div(G, SCAN_TILE);
    scan_local_kernel<<<nsb, SCAN_THREADS, 0, stream>>>(w.tiles, w.offsets, w.block_sums, G);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, stream>>>(w.block_sums, nsb, w.total);
    scan_add_kernel<<<nsb, SCAN_THREADS, 0, stream>>>(w.offsets, w.block_sums, G);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(4);
    uint32_t D32 = 0;
    SIU3R_CUDA_CHECK(cudaMemcpyAsync(&D32, w.total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    SIU3R_CUDA_CHECK(cudaStreamSynchronize(stream));   // also covers the pageable `scal` upload above
    const int64_t D = (int64_t)D32;
    if (num_rendered_host) *num_rendered_host = D;
    if (D > dup_capacity) return SIU3R_ERR_CAPACITY;
    if (D > 0) {
        duplicate_with_keys_kernel<<<ceil_div(G, 256), 256, 0, stream>>>(G, gx, radii, w.offsets, w.depths, w.rects, w.keys, w.vals);
        const int end_bit = 32 + higher_msb((uint32_t)(gx * gy));
        size_t tb = w.cub_bytes;
        SIU3R_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.keys, w.keys_sorted, w.vals, w.vals_sorted, (int)D, 0, end_bit, stream));
        identify_tile_ranges_kernel<<<(unsigned)ceil_div_i64(D, 256), 256, 0, stream>>>((uint32_t)D, w.keys_sorted, w.ranges);
        SIU3R_LAUNCH_CHECK();
        siu3r_note_launch(2);
    }
    dim3 grid(gx, gy, ceil_div(C, FCH));
    render_feat_kernel<<<grid, RB, 0, stream>>>(W, H, gx, w.ranges, w.vals_sorted, w.xy, w.conic_o, features, C, out_features, out_alpha, nullptr);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

// The same frame without any host synchronisation, on the binned path of siu3r_raster_forward_nosync (per-tile counts -> scan -> scatter of
// (depth bits << 32 | id) words -> per-tile register sort: the order the global radix sort of (tile << 32 | depth) yields, so the outputs are
// bit-identical).  status_dev: 4 device words {duplicates, largest tile, flags, 0}; flags != 0 (capacity exceeded / a tile above 8192 records): nothing
// was rendered and the caller re-renders through siu3r_raster_features_forward.  intr = (fx, fy, cx, cy) are passed by value.
int siu3r_raster_features_forward_nosync(int G, int H, int W, int C, int cov_stride, const float* means3D, const float* cov, const float* opacities,
                                         const float* features, const float* viewmat, float fx, float fy, float cx, float cy, float near_plane,
                                         float far_plane, float* out_features, float* out_alpha, int32_t* radii_xy, void* workspace,
                                         int64_t workspace_bytes, int64_t dup_capacity, uint32_t* status_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    SIU3R_REQUIRE(G > 0 && H > 0 && W > 0 && C > 0 && (cov_stride == 6 || cov_stride == 9) && status_dev);
    SIU3R_REQUIRE(means3D && cov && opacities && features && viewmat && out_features && workspace);
    SIU3R_REQUIRE(dup_capacity > 0 && dup_capacity < (1ll << 31));
    const int gx = ceil_div(W, TILE_X), gy = ceil_div(H, TILE_Y), ntiles = gx * gy;
    SIU3R_REQUIRE(gx < 65536 && gy < 65536 && ceil_div(C, FCH) < 65536);
    if ((size_t)ntiles * 8 > 200 * 1024) return SIU3R_ERR_UNSUPPORTED;
    Workspace w = carve(workspace, G, H, W, dup_capacity);
    if ((int64_t)w.bytes > workspace_bytes) return SIU3R_ERR_CAPACITY;
    static bool attr_bin[64] = {false};
    int dev = 0;
    SIU3R_CUDA_CHECK(cudaGetDevice(&dev));
    if (!attr_bin[dev & 63]) {
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(bin_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        SIU3R_CUDA_CHECK(cudaFuncSetAttribute(bin_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_bin[dev & 63] = true;
    }
    int32_t* radii = reinterpret_cast<int32_t*>(w.rgb);
    SIU3R_CUDA_CHECK(cudaMemsetAsync(w.tile_counts, 0, sizeof(uint32_t) * ntiles, stream));
    pack_feat_camera_kernel<<<1, 32, 0, stream>>>(viewmat, fx, fy, cx, cy, near_plane, far_plane, w.cam);
    preprocess_feat_kernel<<<ceil_div(G, 128), 128, 0, stream>>>(G, H, W, gx, gy, cov_stride, means3D, cov, opacities, w.cam, w.depths, w.xy,
                                                                w.conic_o, w.tiles, w.rects, radii, radii_xy);
    bin_count_kernel<<<ceil_div(G, BIN_CHUNK), BIN_THREADS, (size_t)ntiles * 4, stream>>>(G, gx, ntiles, radii, w.rects, w.tile_counts);
    tile_scan_kernel<<<1, 1024, 0, stream>>>(w.tile_counts, ntiles, w.ranges, w.tile_cursors, w.total, (uint32_t)dup_capacity, status_dev);
    bin_scatter_kernel<<<ceil_div(G, BIN_CHUNK), BIN_THREADS, (size_t)ntiles * 8, stream>>>(G, gx, ntiles, radii, w.depths, w.rects, w.tile_cursors,
                                                                                          w.keys, status_dev);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(5);
    int r = launch_tile_sorts(ntiles, w.ranges, w.keys, w.vals_sorted, TS_CAP, status_dev, stream); if (r) return r;
    dim3 grid(gx, gy, ceil_div(C, FCH));
    render_feat_kernel<<<grid, RB, 0, stream>>>(W, H, gx, w.ranges, w.vals_sorted, w.xy, w.conic_o, features, C, out_features, out_alpha, status_dev);
    SIU3R_LAUNCH_CHECK();
    siu3r_note_launch(1);
    return SIU3R_OK;
}

}  // extern "C"
