/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 *
 * CPU restatement (plain C, fp32, IEEE op-by-op: build with -ffp-contract=off) of the forward
 * pass of the third-party 3DGS rasterizer the reference calls at
 *   /root/reference/src/models/cuda_splatting.py:90-118   (GaussianRasterizer(settings)(...))
 * i.e. `diff-gaussian-rasterization` = rmurai0610/diff-gaussian-rasterization-w-pose
 * @ 43e21bff91cd24986ee3dd52fe0bb06952e50ec7 (uv.lock:439-441 of the reference).
 *
 * PARITY UNPINNED: that package is NOT under /root/reference (un-vendored git dependency, CUDA-only,
 * no network), the reference ships no tests / golden vectors for it, so this file restates the
 * published algorithm of the graphdeco-inria lineage (cuda_rasterizer/{forward.cu,rasterizer_impl.cu,
 * auxiliary.h}) as summarised in SURVEY.md Appendix D, driven exactly as render_cuda drives it
 * (cuda_splatting.py:63-118: transposed matrices, cov3D_precomp order xx,xy,xz,yy,yz,zz,
 * sh_degree 4 with 25 coefficients of which the kernel evaluates degrees 0..3).
 *
 * Steps: preprocess -> inclusive scan -> duplicateWithKeys -> stable radix sort on
 * (tile<<32 | depth bits) -> identifyTileRanges -> per-tile front-to-back blend.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLOCK_X 16
#define BLOCK_Y 16

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

static inline float ndc2pix(float v, int S) { return ((v + 1.0f) * (float)S - 1.0f) * 0.5f; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* matrices are stored as the reference passes them: element [r][c] of the torch tensor at m[4*r+c];
 * the kernel reads "column-major", i.e. x' = m[0]x + m[4]y + m[8]z + m[12]. */
static inline void xform4x3(const float* p, const float* m, float* o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static inline void xform4x4(const float* p, const float* m, float* o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
    o[3] = m[3] * p[0] + m[7] * p[1] + m[11] * p[2] + m[15];
}

/* EWA projection of the 3-D covariance (computeCov2D): returns (a, b, c) of the 2x2 covariance */
static void cov2d(const float* mean, float fx, float fy, float tanx, float tany, const float* c3,
                  const float* vm, float* out) {
    float t[3];
    xform4x3(mean, vm, t);
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf_(limx, fmaxf_(-limx, txtz)) * t[2];
    t[1] = fminf_(limy, fmaxf_(-limy, tytz)) * t[2];
    /* J rows (maths): [fx/tz, 0, -fx*tx/tz^2], [0, fy/tz, -fy*ty/tz^2] */
    const float j00 = fx / t[2], j02 = -(fx * t[0]) / (t[2] * t[2]);
    const float j11 = fy / t[2], j12 = -(fy * t[1]) / (t[2] * t[2]);
    /* R = rotation part of world->view (maths row i, col j) = vm[4*j + i] */
    float R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = vm[4 * j + i];
    /* M = J * R  (2x3), row r: sum_k J[r][k] * R[k][c], accumulated k = 0,1,2 */
    float M[2][3];
    for (int c = 0; c < 3; ++c) {
        M[0][c] = j00 * R[0][c] + 0.0f * R[1][c] + j02 * R[2][c];
        M[1][c] = 0.0f * R[0][c] + j11 * R[1][c] + j12 * R[2][c];
    }
    const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    /* cov = M V M^T ; first MV = M*V (2x3), then (MV) M^T */
    float MV[2][3];
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) MV[r][c] = M[r][0] * V[0][c] + M[r][1] * V[1][c] + M[r][2] * V[2][c];
    float a = MV[0][0] * M[0][0] + MV[0][1] * M[0][1] + MV[0][2] * M[0][2];
    /* upstream returns cov[0][1] of the column-major glm result = maths entry (1,0) */
    float b = MV[1][0] * M[0][0] + MV[1][1] * M[0][1] + MV[1][2] * M[0][2];
    float c = MV[1][0] * M[1][0] + MV[1][1] * M[1][1] + MV[1][2] * M[1][2];
    out[0] = a + 0.3f;
    out[1] = b;
    out[2] = c + 0.3f;
}

static void sh_to_rgb(const float* mean, const float* campos, const float* sh /* [M][3] */, int deg, float* rgb) {
    float dir[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
    float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    float x = dir[0] / len, y = dir[1] / len, z = dir[2] / len;
    for (int ch = 0; ch < 3; ++ch) {
        float r = SH_C0 * sh[0 * 3 + ch];
        if (deg > 0) {
            r = r - SH_C1 * y * sh[1 * 3 + ch] + SH_C1 * z * sh[2 * 3 + ch] - SH_C1 * x * sh[3 * 3 + ch];
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * sh[4 * 3 + ch] + SH_C2[1] * yz * sh[5 * 3 + ch] +
                    SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + ch] + SH_C2[3] * xz * sh[7 * 3 + ch] +
                    SH_C2[4] * (xx - yy) * sh[8 * 3 + ch];
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + ch] + SH_C3[1] * xy * z * sh[10 * 3 + ch] +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + ch] +
                        SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + ch] +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + ch] + SH_C3[5] * z * (xx - yy) * sh[14 * 3 + ch] +
                        SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + ch];
                }
            }
        }
        r += 0.5f;
        rgb[ch] = r < 0.0f ? 0.0f : r;
    }
}

typedef struct {
    int64_t num_rendered; /* D = number of (tile, gaussian) duplicates */
} raster_stats;

/*
 * Full forward.  Optional debug outputs may be NULL.  keys/values must hold `dup_capacity` entries
 * when non-NULL; returns -2 if D exceeds dup_capacity.  Returns 0 on success.
 */
int siu3r_oracle_rasterize(int G, int H, int W, int sh_degree, int sh_coeffs, const float* means3D, const float* cov3D,
                           const float* shs, const float* opacities, const float* viewmatrix, const float* projmatrix,
                           const float* campos, float tan_fovx, float tan_fovy, const float* bg, float* out_color,
                           float* out_depth, float* out_opacity, int32_t* radii, int32_t* n_touched,
                           uint32_t* tiles_touched_out, uint32_t* offsets_out, uint64_t* keys_out, uint32_t* values_out,
                           int64_t dup_capacity, uint32_t* ranges_out, int64_t* num_rendered_out) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const float focal_y = (float)H / (2.0f * tan_fovy), focal_x = (float)W / (2.0f * tan_fovx);
    const int deg = sh_degree > 3 ? 3 : sh_degree; /* upstream evaluates at most degree 3 */

    float* depths = (float*)calloc((size_t)G, sizeof(float));
    float* xy = (float*)calloc((size_t)G * 2, sizeof(float));
    float* conic_o = (float*)calloc((size_t)G * 4, sizeof(float));
    float* rgb = (float*)calloc((size_t)G * 3, sizeof(float));
    uint32_t* tiles = (uint32_t*)calloc((size_t)G, sizeof(uint32_t));
    uint32_t* rects = (uint32_t*)calloc((size_t)G * 4, sizeof(uint32_t));
    uint32_t* offs = (uint32_t*)calloc((size_t)G, sizeof(uint32_t));
    memset(radii, 0, (size_t)G * sizeof(int32_t));
    memset(n_touched, 0, (size_t)G * sizeof(int32_t));

    /* (1) preprocess */
    for (int i = 0; i < G; ++i) {
        const float* p = means3D + 3 * (size_t)i;
        float pv[3];
        xform4x3(p, viewmatrix, pv);
        if (pv[2] <= 0.2f) continue;
        float ph[4];
        xform4x4(p, projmatrix, ph);
        const float pw = 1.0f / (ph[3] + 0.0000001f);
        const float px = ph[0] * pw, py = ph[1] * pw;
        float cov[3];
        cov2d(p, focal_x, focal_y, tan_fovx, tan_fovy, cov3D + 6 * (size_t)i, viewmatrix, cov);
        const float det = cov[0] * cov[2] - cov[1] * cov[1];
        if (det == 0.0f) continue;
        const float det_inv = 1.0f / det;
        const float mid = 0.5f * (cov[0] + cov[2]);
        const float disc = sqrtf(fmaxf_(0.1f, mid * mid - det));
        const float lambda1 = mid + disc, lambda2 = mid - disc;
        const float my_radius = ceilf(3.0f * sqrtf(fmaxf_(lambda1, lambda2)));
        const float pix = ndc2pix(px, W), piy = ndc2pix(py, H);
        const int rminx = imin(gx, imax(0, (int)((pix - my_radius) / BLOCK_X)));
        const int rminy = imin(gy, imax(0, (int)((piy - my_radius) / BLOCK_Y)));
        const int rmaxx = imin(gx, imax(0, (int)((pix + my_radius + BLOCK_X - 1) / BLOCK_X)));
        const int rmaxy = imin(gy, imax(0, (int)((piy + my_radius + BLOCK_Y - 1) / BLOCK_Y)));
        if ((rmaxx - rminx) * (rmaxy - rminy) == 0) continue;
        sh_to_rgb(p, campos, shs + (size_t)i * sh_coeffs * 3, deg, rgb + 3 * (size_t)i);
        depths[i] = pv[2];
        radii[i] = (int32_t)my_radius;
        xy[2 * i] = pix;
        xy[2 * i + 1] = piy;
        conic_o[4 * i + 0] = cov[2] * det_inv;
        conic_o[4 * i + 1] = -cov[1] * det_inv;
        conic_o[4 * i + 2] = cov[0] * det_inv;
        conic_o[4 * i + 3] = opacities[i];
        tiles[i] = (uint32_t)((rmaxy - rminy) * (rmaxx - rminx));
        rects[4 * i + 0] = rminx; rects[4 * i + 1] = rminy; rects[4 * i + 2] = rmaxx; rects[4 * i + 3] = rmaxy;
    }

    /* (2) inclusive scan */
    uint64_t run = 0;
    for (int i = 0; i < G; ++i) { run += tiles[i]; offs[i] = (uint32_t)run; }
    const int64_t D = (int64_t)run;
    if (num_rendered_out) *num_rendered_out = D;
    if (tiles_touched_out) memcpy(tiles_touched_out, tiles, (size_t)G * sizeof(uint32_t));
    if (offsets_out) memcpy(offsets_out, offs, (size_t)G * sizeof(uint32_t));
    if ((keys_out || values_out) && D > dup_capacity) {
        free(depths); free(xy); free(conic_o); free(rgb); free(tiles); free(rects); free(offs);
        return -2;
    }

    /* (3) duplicateWithKeys */
    uint64_t* keys = (uint64_t*)malloc((size_t)(D > 0 ? D : 1) * sizeof(uint64_t));
    uint32_t* vals = (uint32_t*)malloc((size_t)(D > 0 ? D : 1) * sizeof(uint32_t));
    for (int i = 0; i < G; ++i) {
        if (radii[i] <= 0) continue;
        uint64_t off = (i == 0) ? 0 : offs[i - 1];
        uint32_t dbits;
        memcpy(&dbits, &depths[i], 4);
        for (uint32_t y = rects[4 * i + 1]; y < rects[4 * i + 3]; ++y)
            for (uint32_t x = rects[4 * i + 0]; x < rects[4 * i + 2]; ++x) {
                uint64_t key = (uint64_t)(y * (uint32_t)gx + x);
                key <<= 32;
                key |= dbits;
                keys[off] = key;
                vals[off] = (uint32_t)i;
                ++off;
            }
    }

    /* (4) stable LSD radix sort, 8-bit digits over all 64 key bits (superset of the bits CUB sorts:
     *     bits above 32+msb(tiles) are zero, so the order is identical) */
    {
        uint64_t* k2 = (uint64_t*)malloc((size_t)(D > 0 ? D : 1) * sizeof(uint64_t));
        uint32_t* v2 = (uint32_t*)malloc((size_t)(D > 0 ? D : 1) * sizeof(uint32_t));
        for (int pass = 0; pass < 8; ++pass) {
            size_t cnt[257];
            memset(cnt, 0, sizeof(cnt));
            const int sh = pass * 8;
            for (int64_t j = 0; j < D; ++j) cnt[((keys[j] >> sh) & 0xFF) + 1]++;
            for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
            for (int64_t j = 0; j < D; ++j) {
                size_t dst = cnt[(keys[j] >> sh) & 0xFF]++;
                k2[dst] = keys[j];
                v2[dst] = vals[j];
            }
            uint64_t* tk = keys; keys = k2; k2 = tk;
            uint32_t* tv = vals; vals = v2; v2 = tv;
        }
        free(k2); free(v2);
    }
    if (keys_out) memcpy(keys_out, keys, (size_t)D * sizeof(uint64_t));
    if (values_out) memcpy(values_out, vals, (size_t)D * sizeof(uint32_t));

    /* (5) tile ranges */
    uint32_t* ranges = (uint32_t*)calloc((size_t)gx * gy * 2, sizeof(uint32_t));
    for (int64_t j = 0; j < D; ++j) {
        uint32_t t = (uint32_t)(keys[j] >> 32);
        if (j == 0) ranges[2 * t] = 0;
        else {
            uint32_t tp = (uint32_t)(keys[j - 1] >> 32);
            if (t != tp) { ranges[2 * tp + 1] = (uint32_t)j; ranges[2 * t] = (uint32_t)j; }
        }
        if (j == D - 1) ranges[2 * t + 1] = (uint32_t)D;
    }
    if (ranges_out) memcpy(ranges_out, ranges, (size_t)gx * gy * 2 * sizeof(uint32_t));

    /* (6) blend, one pixel at a time, front to back */
    for (int ty = 0; ty < gy; ++ty)
        for (int tx = 0; tx < gx; ++tx) {
            const uint32_t r0 = ranges[2 * (ty * gx + tx)], r1 = ranges[2 * (ty * gx + tx) + 1];
            for (int ly = 0; ly < BLOCK_Y; ++ly)
                for (int lx = 0; lx < BLOCK_X; ++lx) {
                    const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                    if (pxi >= W || pyi >= H) continue;
                    const float pfx = (float)pxi, pfy = (float)pyi;
                    float T = 1.0f, C[3] = {0, 0, 0}, Dz = 0.0f;
                    for (uint32_t j = r0; j < r1; ++j) {
                        const uint32_t id = vals[j];
                        const float dx = xy[2 * id] - pfx, dy = xy[2 * id + 1] - pfy;
                        const float* co = conic_o + 4 * (size_t)id;
                        const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        if (power > 0.0f) continue;
                        const float alpha = fminf_(0.99f, co[3] * expf(power));
                        if (alpha < 1.0f / 255.0f) continue;
                        const float test_T = T * (1.0f - alpha);
                        if (test_T < 0.0001f) break;
                        for (int ch = 0; ch < 3; ++ch) C[ch] += rgb[3 * (size_t)id + ch] * alpha * T;
                        Dz += depths[id] * alpha * T;
                        if (test_T > 0.5f) n_touched[id] += 1;
                        T = test_T;
                    }
                    const size_t pid = (size_t)pyi * W + pxi;
                    for (int ch = 0; ch < 3; ++ch) out_color[(size_t)ch * H * W + pid] = C[ch] + T * bg[ch];
                    out_depth[pid] = Dz;
                    out_opacity[pid] = 1.0f - T;
                }
        }

    free(depths); free(xy); free(conic_o); free(rgb); free(tiles); free(rects); free(offs);
    free(keys); free(vals); free(ranges);
    return 0;
}

/* ----------------------------------------------------------------------------------------------
 * cuRoPE2D CPU restatement (reference: src/models/croco/curope/curope.cpp:11-47 rope_2d_cpu).
 * tokens [B,N,H,D] (in place), pos [B,N,2] int64.  Pinned in tests/test_oracle_cpu.py: the CPU-order variant is bit-identical to the
 * reference's own curope.cpp compiled here (oracle/_ref/curope_ref.so, oracle/build_ref.py); both are within 2e-5 of its PyTorch
 * fallback (croco/pos_embed.py:126-179).
 * -------------------------------------------------------------------------------------------- */
/* cpu_order = 0: angle = p * (fwd / base^(d/Q))   -- the reference's CUDA kernel (kernels.cu:44-53: inv_freq staged per block, then pos * inv_freq)
 * cpu_order = 1: angle = (fwd * p) / base^(d/Q)   -- the reference's CPU function (curope.cpp:36); the two differ by one rounding of the angle */
static void rope2d_impl(float* tokens, const int64_t* pos, int B, int N, int Hh, int D, float base, float fwd, int cpu_order) {
    const int Q = D / 4;
    for (int b = 0; b < B; ++b)
        for (int x = 0; x < 2; ++x)
            for (int n = 0; n < N; ++n) {
                const int64_t p = pos[((size_t)b * N + n) * 2 + x];
                for (int h = 0; h < Hh; ++h) {
                    float* t = tokens + (((size_t)b * N + n) * Hh + h) * D + x * (D / 2);
                    for (int d = 0; d < Q; ++d) {
                        const float pw = powf(base, (float)d / (float)Q);
                        float f;
                        if (cpu_order) {
                            f = (fwd * (float)(int)p) / pw;
                        } else {
                            const float inv_freq = fwd / pw;
                            f = (float)p * inv_freq;
                        }
                        const float c = cosf(f), s = sinf(f);
                        const float u = t[d], v = t[d + Q];
                        t[d] = u * c - v * s;
                        t[d + Q] = v * c + u * s;
                    }
                }
            }
}
void siu3r_oracle_rope2d(float* tokens, const int64_t* pos, int B, int N, int Hh, int D, float base, float fwd) {
    rope2d_impl(tokens, pos, B, N, Hh, D, base, fwd, 0);
}
/* bit-identical to the reference's curope.cpp compiled in this container (oracle/_ref/curope_ref.so): tests/test_oracle_cpu.py */
void siu3r_oracle_rope2d_cpu(float* tokens, const int64_t* pos, int B, int N, int Hh, int D, float base, float fwd) {
    rope2d_impl(tokens, pos, B, N, Hh, D, base, fwd, 1);
}
