/*
 * siu3r_b200 C ABI -- the drop-in boundary of the B200-native SIU3R hot path (libsiu3r_b200.so).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch) owns all memory;
 *   - fp32 storage everywhere, token-major / NHWC layouts ([rows, C] with an explicit leading dimension in floats);
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and never synchronise, except
 *     siu3r_raster_forward (one sync to read the duplicate count, like the reference rasterizer);
 *   - return value: 0 = ok, -1 invalid argument, -2 capacity/workspace too small, -3 CUDA error, -4 unsupported config
 *     (message on stderr); no exceptions cross the ABI, no global state besides the launch counter;
 *   - `precision`: 1 = TF32 tensor-core math (the reference's own GPU mode: src/models/croco/croco.py:13),
 *                  3 = 3xTF32 split (fp32-grade accuracy on the tensor cores; needs *_lo planes from siu3r_split_tf32).
 *   - `round_out` / act bit 4 / eltwise ops 7-8: the producer stores RN_tf32(value) because the tensor only feeds TF32
 *     GEMM A operands (the tensor core itself would truncate, a -3.5e-4 relative bias per GEMM).
 *
 * Every entry point cites the reference interface it replaces (paths relative to /root/reference).
 */
#ifndef SIU3R_B200_H
#define SIU3R_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- bookkeeping --------------------------------------------------------------------------------------------- */
int siu3r_abi_version(void);
void siu3r_note_launch(int n);
long long siu3r_launch_count(void);      /* kernels launched by this library since the last reset */
void siu3r_reset_launch_count(void);
/* Programmatic dependent launch of the h3 GEMM / attention / LayerNorm kernels (prologue of kernel n+1 overlaps the tail of kernel n; results
 * are unchanged).  Default on; SIU3R_PDL=0 in the environment or siu3r_set_pdl(0) turns it off (A/B measurements). */
int siu3r_pdl_enabled(void);
void siu3r_set_pdl(int on);

/* ---- 3D Gaussian splatting rasterizer (forward) ---------------------------------------------------------------
 * Replaces diff_gaussian_rasterization._C.rasterize_gaussians as driven by GaussianRasterizer(settings)(...) at
 * src/models/cuda_splatting.py:90-118 (means3D, shs, opacities, cov3D_precomp; sh_degree = isqrt(25)-1 = 4).
 * Returns the reference 5-tuple (image, radii, depth, opacity, n_touched) of cuda_splatting.py:109. */
int64_t siu3r_raster_workspace_bytes(int G, int H, int W, int64_t dup_capacity);
int siu3r_raster_forward(int G, int H, int W, int sh_degree, int sh_coeffs, int sh_layout, int cov_stride,
                         const float* means3D, const float* cov, const float* shs, const float* opacities,
                         const float* viewmatrix, const float* projmatrix, const float* campos, const float* bg,
                         float tan_fovx, float tan_fovy, float* out_color, float* out_depth, float* out_opacity,
                         int32_t* radii, int32_t* n_touched, void* workspace, int64_t workspace_bytes,
                         int64_t dup_capacity, int64_t* num_rendered_host, uint32_t* debug_tiles_touched,
                         uint32_t* debug_offsets, uint64_t* debug_keys, uint32_t* debug_values, uint32_t* debug_ranges,
                         void* stream);

/* The same frame without the host synchronisation and without debug exports: the duplicate count stays on the device; status_dev [4] uint32 =
 * {duplicates D, largest tile, flags, 0}, flags bit 0 = D > dup_capacity, bit 1 = a tile holds more than 8192 records -- in either case the later
 * kernels return at once and the outputs are undefined: the caller reads status_dev when convenient (once for all cameras of a
 * SplattingCUDA.forward) and re-renders flagged cameras through siu3r_raster_forward.  Capturable in a CUDA graph. */
int siu3r_raster_forward_nosync(int G, int H, int W, int sh_degree, int sh_coeffs, int sh_layout, int cov_stride, const float* means3D,
                                const float* cov, const float* shs, const float* opacities, const float* viewmatrix, const float* projmatrix,
                                const float* campos, const float* bg, float tan_fovx, float tan_fovy, float* out_color, float* out_depth,
                                float* out_opacity, int32_t* radii, int32_t* n_touched, void* workspace, int64_t workspace_bytes,
                                int64_t dup_capacity, uint32_t* status_dev, void* stream);
/* testing aid: 0 = the shared-memory bitonic tile sort of round 1 (tiles <= 2048 records) instead of the register-resident one */
void siu3r_raster_set_regsort(int enabled);

/* N-channel feature splatting: replaces gsplat.rasterization(means, quats=None, scales=None, covars, opacities, colors[N,C], viewmats,
 * Ks, width, height, sh_degree=None, near_plane, far_plane) as called at src/models/gaussian_renderer.py:92-106 (one camera per call).
 * viewmat = world-to-camera 4x4 row-major (device); intr_host = (fx, fy, cx, cy) in pixels (host); out_features [H,W,C], out_alpha [H,W]. */
int siu3r_raster_features_forward(int G, int H, int W, int C, int cov_stride, const float* means3D, const float* cov, const float* opacities,
                                  const float* features, const float* viewmat, const float* intr_host, float near_plane, float far_plane,
                                  float* out_features, float* out_alpha, int32_t* radii_xy, void* workspace, int64_t workspace_bytes,
                                  int64_t dup_capacity, int64_t* num_rendered_host, void* stream);
/* The same frame without host synchronisation (binned per-tile sort, as siu3r_raster_forward_nosync): status = 4 device words {duplicates, largest
 * tile, flags, 0}; flags != 0 -> nothing rendered, re-render through siu3r_raster_features_forward.  Outputs are bit-identical to the call above. */
int siu3r_raster_features_forward_nosync(int G, int H, int W, int C, int cov_stride, const float* means3D, const float* cov, const float* opacities,
                                         const float* features, const float* viewmat, float fx, float fy, float cx, float cy, float near_plane,
                                         float far_plane, float* out_features, float* out_alpha, int32_t* radii_xy, void* workspace,
                                         int64_t workspace_bytes, int64_t dup_capacity, uint32_t* status, void* stream);
/* testing aid: 0 disables the blend kernel's exact sub-tile culling (every pixel then evaluates every record of its tile, the
 * literal loop of the reference rasterizer); results must be bit-identical either way (tests/test_ops_gpu.py) */
void siu3r_raster_set_culling(int enabled);
/* testing aid: 0 = always bin / sort the duplicates the way the reference pipeline does (prefix scan over Gaussians, duplicateWithKeys, global
 * 64-bit radix sort, identifyTileRanges) instead of the per-tile binned sort; both give the identical sorted list */
void siu3r_raster_set_binning(int enabled);

/* ---- dense contractions on the tcgen05 tensor cores ------------------------------------------------------------
 * siu3r_gemm_tc   : torch.nn.functional.linear (+bias, +GELU/ReLU, +residual): croco/blocks.py:74-77,97,110,154-156,167;
 *                   backbone_croco.py:87; vit_adapter/blocks.py:118-121; every nn.Linear / 1x1 Conv2d of
 *                   mask2former/video_seg_decoder.py and heads/dpt_block.py.
 * siu3r_conv2d_tc : nn.Conv2d(k, stride 1, padding) as implicit GEMM over NHWC: heads/dpt_block.py:35-70,98-116,358-364,
 *                   385-387; mask2former/video_seg_decoder.py:2040-2047. */
int siu3r_gemm_tc(int M, int N, int K, const float* A, const float* A_lo, int64_t lda, const float* Wt, const float* W_lo,
                  int64_t ldw, float* C, int64_t ldc, const float* bias, const float* residual, int64_t ldr, int act,
                  float alpha, int precision, void* stream);
/* nn.Linear + curope.rope_2d on output columns [0, rope_cols) in one kernel (the q / k columns of the qkv, q and kv
 * projections, croco/blocks.py:97-103,154-160); positions [M,2] int64, rope_tab from siu3r_rope2d_table (head dim 64) */
int siu3r_gemm_tc_rope(int M, int N, int K, const float* A, const float* A_lo, int64_t lda, const float* Wt, const float* W_lo,
                       int64_t ldw, float* C, int64_t ldc, const float* bias, int act, int precision, const int64_t* positions,
                       const float* rope_tab, int rope_cols, void* stream);
/* two same-shape linear layers with different operands in one persistent launch (TF32 only): the two decoder streams of
 * AsymmetricCroCo (dec_blocks / dec_blocks2, backbone_croco.py:244-250, 514-531).  *_host = host arrays of 2 device pointers (bias_host /
 * residual_host may be null); M_host[2] rows per problem.  -4 when the shape is not eligible (issue two siu3r_gemm_tc instead). */
int siu3r_gemm_tc_group2(const int* M_host, int N, int K, const float* const* A_host, int64_t lda, const float* const* W_host, int64_t ldw,
                         float* const* C_host, int64_t ldc, const float* const* bias_host, const float* const* residual_host, int64_t ldr,
                         int act, float alpha, const int64_t* positions, const float* rope_tab, int rope_cols, float* const* vt_host,
                         const int* vt_cols_host, int64_t vt_ld, int vt_col0, void* stream);
/* siu3r_gemm_tc_rope with the output columns >= vt_col0 (the V third of a qkv projection) written as V^T [(n - vt_col0)][m] (pitch vt_ld)
 * instead of C: the operand siu3r_flash_attn_tc consumes with vt_batch_cols = tokens per image (no siu3r_transpose_v pass).  In the grouped
 * call vt_host[g] / vt_cols_host[g] give each problem's (16-byte aligned) column window inside the V^T rows. */
int siu3r_gemm_tc_rope_vt(int M, int N, int K, const float* A, int64_t lda, const float* Wt, int64_t ldw, float* C, int64_t ldc, const float* bias,
                          int act, const int64_t* positions, const float* rope_tab, int rope_cols, float* vt, int64_t vt_ld, int vt_col0,
                          void* stream);
int siu3r_conv2d_tc(int Nimg, int H, int W, int Cin, int Cout, int KH, int KW, int pad_h, int pad_w, const float* x, const float* x_lo,
                    const float* Wt, const float* W_lo, float* y, int64_t ldc, const float* bias, const float* residual,
                    int64_t ldr, int act, int precision, void* stream);
/* input_merger step of the Gaussian-parameter head (heads/dpt_gs_head.py:113-119,155-164) fused: KH x 1 conv over the row-packed image
 * (siu3r_im2col_nhwc with KH = 1) + bias + act + bilinear x2 (align_corners=True) upsampling of `low` [H/2, W/2, Cout] as residual; one image */
int siu3r_conv_rows_up2x_tc(int H, int W, int KH, int pad, int Cout, const float* rows, const float* Wt, const float* bias, const float* low,
                            float* out, int64_t ldc, int act, void* stream);
/* plain fp32 FFMA GEMM (tiny / odd shapes such as the K = 9 intrinsics encoder, backbone_croco.py:59,278) */
int siu3r_gemm_simt(int M, int N, int K, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                    const float* bias, const float* residual, int64_t ldr, int act, float alpha, void* stream);
/* skinny fp32 GEMM (M <= ~128 rows: the 100 Mask2Former queries, video_seg_decoder.py:957-1025,1423-1480): 32x32 tiles, 8-way in-CTA split-K */
int siu3r_gemm_skinny(int M, int N, int K, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, const float* bias,
                      const float* residual, int64_t ldr, int act, float alpha, void* stream);
int siu3r_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
void siu3r_gemm_debug_set(long long* dev_buf);
/* host-only: what the tile planner of the persistent kernel decides for a shape (tw = 0: one-tile kernels) */
int siu3r_gemm_plan(int M, int N, int K, int M1, int allow_split, int* tw_out, int* nsplit_out, int* tiles_out, int* rounds_out);
void siu3r_gemm_force(int kernel);               /* tuning aid: 0 heuristic, 1 persistent swapped pair kernel (>= 16: that token tile width), 3 one-tile pair, 4 1-CTA */   /* profiling aid: per-CTA clock64 stamps of the 1-CTA linear kernel */

/* ---- transformer pieces ----------------------------------------------------------------------------------------
 * siu3r_rope2d replaces curope.rope_2d(tokens, positions, base, fwd) (croco/curope/curope.cpp:49-65,
 * kernels.cu:17-82; called from croco/curope/curope2d.py:20,27 <- croco/blocks.py:101-103,158-160): in place, tokens
 * [B,N,H,D] with arbitrary batch/token strides (so q and k can be rotated inside the fused qkv buffer), positions
 * [B,N,2] int64 (y, x); nparts > 1 rotates several tensors that share the positions in one launch (q and k
 * of the fused qkv buffer: part_stride = C).  Error contract: D % 4 != 0 -> -1 ("token dim must be multiple of 4", kernels.cu:94). */
int siu3r_rope2d(float* tokens, const int64_t* positions, int B, int N, int H, int D, int64_t batch_stride,
                 int64_t token_stride, float base, float fwd, int nparts, int64_t part_stride, int round_out, void* stream);
/* tab[maxpos][D/4] (cos, sin) float pairs of pos * fwd / base^(d/(D/4)): the per-position factors of kernels.cu:36-48 */
int siu3r_rope2d_table(float* tab, int maxpos, int D, float base, float fwd, void* stream);
/* nn.LayerNorm over the last dim (+ optional fused add of `add` rows): croco/blocks.py:119-125,176-184 */
int siu3r_layernorm(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy, int rows, int C,
                    float eps, const float* add, int64_t ldadd, int round_out, void* stream);
/* two LayerNorms with different affine parameters / buffers in one launch (the two decoder streams: norm1, norm2, norm3, norm_y of
 * dec_blocks[l] and dec_blocks2[l], croco/blocks.py:186-190); same C, pitches and eps; segment g has rows_g rows */
int siu3r_layernorm_group2(const float* x0, const float* x1, int64_t ldx, const float* w0, const float* b0, const float* w1, const float* b1,
                           float* y0, float* y1, int64_t ldy, int rows0, int rows1, int C, float eps, int round_out, void* stream);
/* softmax(Q K^T * scale) V, head dim 64: croco/blocks.py:105-109 (Attention), :162-166 (CrossAttention) */
int siu3r_flash_attn_d64(const float* Q, int64_t q_bs, int64_t q_ts, const float* K, int64_t k_bs, int64_t k_ts,
                         const float* V, int64_t v_bs, int64_t v_ts, float* O, int64_t o_bs, int64_t o_ts, int B, int H,
                         int Nq, int Nk, float scale, int precision, int round_out, void* stream);
/* tcgen05 / TMEM flash attention (TF32 mode) and its V^T producer: same contract as siu3r_flash_attn_d64 */
int siu3r_transpose_v(const float* V, int64_t v_bs, int64_t v_ts, int B, int N, int H, float* Vt, int64_t ld, void* stream);
int siu3r_flash_attn_tc(const float* Q, int64_t q_bs, int64_t q_ts, int q_width, int q_col0, const float* K, int64_t k_bs,
                        int64_t k_ts, int k_width, int k_col0, const float* Vt, int64_t vt_ld, int64_t vt_batch_cols, float* O, int64_t o_bs,
                        int64_t o_ts, int B, int H, int Nq, int Nk, float scale, int round_out, void* stream);
/* vt_batch_cols = 0: Vt rows are (b*H + h)*64 + d, columns = keys of image b (siu3r_transpose_v layout);
 * vt_batch_cols > 0: Vt rows are h*64 + d, image b's keys start at column b * vt_batch_cols (layout written by siu3r_gemm_tc_rope_vt) */
/* masked / plain attention with head dim 32: mask2former/video_seg_decoder.py:975-983,994-999,1306-1308.
 * Keys are split over CTAs (K/V staged once per head in shared memory); workspace = scratch for the row flags and the
 * per-split softmax partials, at least siu3r_attn_small_d32_ws_bytes(B, H, Nq, Nk) bytes, 256-byte aligned. */
int64_t siu3r_attn_small_d32_ws_bytes(int B, int H, int Nq, int Nk);
int siu3r_attn_small_d32(const float* Q, int64_t q_bs, int64_t q_ts, const float* K, int64_t k_bs, int64_t k_ts,
                         const float* V, int64_t v_bs, int64_t v_ts, float* O, int64_t o_bs, int64_t o_ts,
                         const uint8_t* mask, int B, int H, int Nq, int Nk, float scale, int round_out, void* workspace,
                         int64_t workspace_bytes, void* stream);
/* multi_scale_deformable_attention (vit_adapter/blocks.py:171-213,217-267; video_seg_decoder.py:1679-1720) */
int siu3r_msdeform_attn(const float* value, int64_t ldv, int Lin, const float* ow, int64_t ldow, const float* ref,
                        const int* level_hw_host, int L, int P, int B, int Lq, int nH, int hd, float* out, int64_t ldo,
                        int round_out, void* stream);

/* ---- HBM-bound elementwise / resampling ------------------------------------------------------------------------ */
int siu3r_eltwise(int op, const float* a, const float* b, float* out, int64_t n, void* stream);
/* in-place scene rescale of SplattingCUDA.forward (gaussian_renderer.py:43-46) */
int siu3r_scale(const float* a, float alpha, float* out, int64_t n, void* stream);
int siu3r_rows_affine(const float* x, int64_t ldx, const float* scale, const float* shift, const float* add,
                      int64_t ldadd, float* y, int64_t ldy, int64_t rows, int C, int relu, void* stream);
/* F.interpolate(mode="bilinear"): heads/dpt_block.py:229-235,279-284; vit_adapter.py:429-433; video_seg_decoder.py:2173 */
int siu3r_resize_bilinear_nhwc(const float* x, int N, int H, int W, int C, int64_t ldx, float* y, int OH, int OW,
                               int64_t ldy, int align_corners, int accumulate, void* stream);
/* scatter half of nn.ConvTranspose2d(kernel = stride): heads/dpt_block.py:422-451; vit_adapter.py:356,425 */
int siu3r_pixel_shuffle_nhwc(const float* g, int N, int H, int W, int C, int s, const float* add, float* y, void* stream);
/* out[(n,oh,ow), (kh,kw,ci)] patches (row stride ldo, padding columns zeroed); pad_h / pad_w = zero padding per axis.
 * KH = 1 packs the KW horizontal taps of a few-channel input (the RGB image) into one 32-wide "channel" vector so that a
 * KHxKW conv becomes a KHx1 implicit-GEMM conv (heads/dpt_gs_head.py input_merger 7x7, Cin = 3). */
int siu3r_im2col_nhwc(const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad_h, int pad_w,
                      float* out, int64_t ldo, int round_out, void* stream);
int siu3r_nchw_to_nhwc(const float* x, float* y, int N, int C, int HW, int64_t ldy, void* stream);
int siu3r_nhwc_to_nchw(const float* x, int64_t ldx, float* y, int N, int C, int HW, void* stream);
int siu3r_maxpool3x3s2_nhwc(const float* x, int N, int H, int W, int C, float* y, void* stream);   /* vit_adapter.py:220 */
int siu3r_dwconv3x3_nhwc(const float* x, int64_t ldx, int64_t batch_stride_x, int N, int H, int W, int C, const float* w,
                         const float* b, float* y, int64_t ldy, int64_t batch_stride_y, int gelu, void* stream); /* :16-31 */
int siu3r_groupnorm_nhwc(const float* x, int N, int HW, int C, int groups, const float* w, const float* b, float eps,
                         int relu, float* y, double* stats_ws, void* stream);                         /* video_seg_decoder.py:2004,2036,2048 */

/* ---- heads / outputs ------------------------------------------------------------------------------------------- */
int siu3r_depth_exp(const float* xyz, int64_t ldx, float* pts, int64_t n, void* stream);   /* heads/postprocess.py:46-61 */
/* UnifiedGaussianAdapter.forward (gaussian_adapter.py:81-110) */
int siu3r_gaussian_adapter(const float* raw, int64_t G, float* covariances, float* harmonics, float* opacities,
                           float* scales, float* rotations, void* stream);
/* VideoMask2FormerMaskPredictor attention mask (video_seg_decoder.py:1461-1478) */
int siu3r_attn_mask_from_logits(const float* logits, int B, int T, int Hm, int Wm, int Q, int oh, int ow, uint8_t* mask,
                                void* stream);
/* post_process_panoptic_segmentation device stages (image_processing_video_mask2former.py:1386-1467; model.py:258-294) */
int siu3r_resize_select(const float* x, int N, int H, int W, int C, const int* idx, int nsel, float* y, int OH, int OW,
                        void* stream);
int siu3r_argmax_area(const float* probs, int64_t npix, int nq, const float* score, float thr, int32_t* labels,
                      int32_t* area, int32_t* orig, void* stream);
int siu3r_label_lut(const int32_t* labels, int64_t npix, const int32_t* seg_lut, const int32_t* sem_lut, int32_t* seg,
                    int32_t* sem, int32_t* inst, void* stream);
int siu3r_qc_logits(const float* probs, int64_t npix, int nq, const int* keep, int nk, const float* cls, int ncls,
                    float* out, void* stream);

/* ---- render record of the joint-scene all-gather (no precedent in the single-GPU reference) ----
 * 85 floats per Gaussian: means 3 | covariance upper triangle 6 (the cov3D_precomp gather of cuda_splatting.py:107,115) | harmonics 75 | opacity. */
int siu3r_render_record_pack(const float* means, const float* cov33, const float* harmonics, const float* opacities, int64_t G, float* out,
                             void* stream);
int siu3r_render_record_unpack(const float* rec, int64_t G, float* means, float* cov6, float* harmonics, float* opacities, void* stream);

/* ---- output wire format ------------------------------------------------------------------------------------------
 * Packed little-endian PLY vertex records of export_ply (src/utils/ply_export.py:12-97; field order :12-27, log(scales) :74,
 * rotations re-ordered xyzw -> wxyz :52-53, i4 labels :62-64, flattened seg_query_class_logits :65-71): the device buffer copied to
 * the host is the body of the file written by inference.py:137-150. */
int siu3r_ply_record_words(int d_sh, int dc_only, int has_labels, int qc_words);
int siu3r_ply_pack(const float* means, const float* scales, const float* rotations, const float* harmonics, const float* opacities,
                   const int32_t* semantic_labels, const int32_t* instance_labels, const float* qc_logits, int64_t G, int d_sh, int dc_only,
                   int qc_words, uint32_t* out, void* stream);

/* ---- 2-D labels from rendered query-class logits ---------------------------------------------------------------
 * Replaces the tensor code of src/pipeline.py:132-164,182-186 (validation / test step; duplicated in viewer.py:404-446) for ONE sample:
 * logits = render_qc_logits [V, Q, C, H, W] given by element strides (the rasteriser's memory is [V, H, W, Q*C]); C = classes + void,
 * void last.  Per pixel: max over queries, class axis rotated void-first, max over classes, sem_logit < threshold -> 0, instance id =
 * argmax query + 1 (0 where sem = 0); pixels whose semantic id equals fuse_sem[i] get instance id fuse_ins[i] (pipeline.py:182-186:
 * stuff + 1 -> num_queries + stuff + 1; viewer.py:433-434: 1 -> 102, 2 -> 103), n_fuse <= 8, host arrays.  sem_id / ins_id: [V, H, W] int64 (torch.max
 * indices).  first_sem [Q] (device, int32): semantic id of the first pixel (v, h, w order) a query owns before fusing, -1 = none --
 * what pipeline.py:166-180 reads with q_sem_ids[0]; first_pix [Q] int32 is workspace. */
int siu3r_labels_from_qc_logits(const float* logits, int V, int Q, int C, int H, int W, int64_t sv, int64_t sq, int64_t sc, int64_t sh,
                                int64_t sw, float threshold, const int* fuse_sem, const int* fuse_ins, int n_fuse, int64_t* sem_id,
                                int64_t* ins_id, int32_t* first_pix, int32_t* first_sem, void* stream);

/* ---- image ingest ---------------------------------------------------------------------------------------------------
 * preprocess_image (inference.py:13-38): PIL Image.resize((out_w, out_h), LANCZOS) of an 8-bit RGB frame [H, W, 3] (src_pitch bytes per
 * row), crop of the window (crop_x, crop_y, cw, ch) of the resized image, /255 -> out [3, ch, cw] float32.  The resampling is Pillow's
 * fixed-point algorithm (libImaging/Resample.c; two 8-bit passes, horizontal first), bit-exact.  bounds_* [out][2] = (first tap, tap count),
 * k* [out][ksize] = round(coefficient * 2^22): device tables built by the host exactly as Pillow's precompute_coeffs /
 * normalize_coeffs_8bpc do (siu3r_b200/io.py: lanczos_tables).  Only source rows [row0, row0 + rows) -- those the cropped output rows
 * touch -- go through the horizontal pass; tmp = rows * cw * 3 bytes of workspace.  Parts of the crop window outside the resized image
 * come out black (0.0), as PIL's Image.crop fills them. */
int siu3r_resize_lanczos_u8(const uint8_t* src, int H, int W, int64_t src_pitch, const int32_t* bounds_x, const int32_t* kx, int ksize_x,
                            int out_w, const int32_t* bounds_y, const int32_t* ky, int ksize_y, int out_h, int crop_x, int crop_y, int cw,
                            int ch, int row0, int rows, uint8_t* tmp, float* out, void* stream);

/* ---- "h3" mode: fp32-grade results on the fp16 tensor-core path ----------------------------------------------------
 * The mode that meets the north-star tolerances (Gaussians 1e-3 abs, segmentation logits 1e-4 rel) at about the cost of the
 * TF32 mode.  A tensor that only feeds tensor-core operands is stored as an fp16 PLANE PAIR: hi = fp16(x) and
 * lo = fp16((x - hi) * 2^11), the lo plane `plane` elements after the hi plane, both with the same row pitch (pointers 16-byte
 * aligned, pitches / plane distances multiples of 8 elements).  Products are evaluated as hi.hi + (lo.hi + hi.lo) * 2^-11 with
 * kind::f16 tcgen05 MMAs and fp32 accumulation (22 significand bits per operand, like a 3xTF32 split).  The attention kernel takes
 * plane pairs with an UNSCALED lo plane (lo = fp16(x - hi)); the projection GEMM writes them when asked to (unscaled_lo).
 * These entry points replace the same reference ops as their TF32 counterparts above (nn.Linear / Conv2d / Attention /
 * CrossAttention / LayerNorm / F.interpolate in croco/blocks.py, heads/dpt_block.py, vit_adapter/, mask2former/). */
int siu3r_split_h3(const float* x, int64_t ldx, int64_t rows, int cols, void* out, int64_t ldo, int64_t plane, int unscaled_lo, void* stream);
int siu3r_merge_h3(const void* in, int64_t ldi, int64_t plane, int64_t rows, int cols, float* y, int64_t ldy, void* stream);
/* one (ngroups = 1) or two same-shape linear layers in one persistent launch; *_host = host arrays of ngroups device pointers */
int siu3r_gemm_h3(int ngroups, const int* M_host, int N, int K, const void* const* X_host, int64_t lda, int64_t a_plane, const void* const* W_host,
                  int64_t ldw, int64_t w_plane, float* const* C_host, int64_t ldc, void* const* Ch_host, int64_t ldh, int64_t h_plane,
                  const float* const* bias_host, const float* const* residual_host, int64_t ldr, int act, float alpha, const int64_t* positions,
                  const float* rope_tab, int rope_cols, void* const* vt_host, const int* vt_cols_host, int64_t vt_ld, int64_t vt_plane, int vt_col0,
                  int unscaled_lo, void* stream);
/* siu3r_gemm_h3 with LayerNorm fused on either side (croco/blocks.py:127-130,186-190: x + f(norm(x)) blocks):
 *   stats_out_host[g] (int64 [M_g][2], zeroed by the caller): the launch adds the statistics (sum * 2^32, sum of squares * 2^26: fixed point) of the rows it writes;
 *   stats_in_host[g] + ln_s_host[g]: X_g are RAW rows, W_g = W*gamma, bias_g = W beta + b, ln_s[n] = sum_k gamma_k W[n,k]; computes Linear(LayerNorm(x));
 *   C_host and Ch_host may both be given (fp32 residual stream + plane pair for the next GEMM). */
int siu3r_gemm_h3_ln(int ngroups, const int* M_host, int N, int K, const void* const* X_host, int64_t lda, int64_t a_plane, const void* const* W_host,
                     int64_t ldw, int64_t w_plane, float* const* C_host, int64_t ldc, void* const* Ch_host, int64_t ldh, int64_t h_plane,
                     const float* const* bias_host, const float* const* residual_host, int64_t ldr, int act, float alpha, const int64_t* positions,
                     const float* rope_tab, int rope_cols, void* const* vt_host, const int* vt_cols_host, int64_t vt_ld, int64_t vt_plane, int vt_col0,
                     int unscaled_lo, const int64_t* const* stats_in_host, const float* const* ln_s_host, float ln_eps, int64_t* const* stats_out_host,
                     void* stream);
int siu3r_conv2d_h3(int Nimg, int H, int W, int Cin, int Cout, int KH, int KW, int pad_h, int pad_w, const void* x, int64_t ldx, int64_t x_plane,
                    const void* Wt, int64_t ldw, int64_t w_plane, float* y, int64_t ldc, void* yh, int64_t ldh, int64_t h_plane, const float* bias,
                    const float* residual, int64_t ldr, int act, void* stream);
int siu3r_flash_attn_h3(const void* Q, int64_t q_bs, int64_t q_ts, int64_t q_plane, int q_width, int q_col0, const void* K, int64_t k_bs,
                        int64_t k_ts, int64_t k_plane, int k_width, int k_col0, const void* Vt, int64_t vt_ld, int64_t vt_plane,
                        int64_t vt_batch_cols, int vt_b_split, int64_t vt_extra, float* O, int64_t o_bs, int64_t o_ts, void* Oh, int64_t oh_bs,
                        int64_t oh_ts, int64_t oh_plane, int B, int H, int Nq, int Nk, float scale, void* stream);
int siu3r_transpose_v_h3(const float* V, int64_t v_bs, int64_t v_ts, int B, int N, int H, void* Vt, int64_t ld, int64_t plane, void* stream);
int siu3r_layernorm_h3(const float* x0, const float* x1, int64_t ldx, const float* w0, const float* b0, const float* w1, const float* b1, float* y0,
                       float* y1, int64_t ldy, void* yh0, void* yh1, int64_t ldh, int64_t plane, int rows0, int rows1, int C, float eps,
                       void* stream);
int siu3r_eltwise_h3(int op, const float* a, const float* b, void* outh, int64_t plane, int64_t n, void* stream);
int siu3r_resize_bilinear_nhwc_h3(const float* x, int N, int H, int W, int C, int64_t ldx, void* yh, int OH, int OW, int64_t ldh, int64_t plane,
                                  int align_corners, void* stream);
int siu3r_im2col_nhwc_h3(const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int pad_w, void* outh, int64_t ldo,
                         int64_t plane, void* stream);
/* tuning / debugging aids */
void siu3r_gemm_h3_force(int tw);
void siu3r_gemm_h3_set_mhalf(int on);
void siu3r_gemm_h3_set_remainder_tiles(int on);
void siu3r_gemm_h3_cluster_cap(int cap);
void siu3r_gemm_h3_order(int order);   /* 0 = neighbouring CTA pairs share the token tile, 1 = they share the weight rows */
int siu3r_gemm_h3_plan(int M, int N, int K, int M1, int* tw_out, int* tiles_out, int* rounds_out);
void siu3r_flash_h3_debug_swap(int swap);
void siu3r_gemm_h3_debug_ts(void* dev_buf);   /* tuning aid: per-tile clock64 stamps of CTA pair 0 into dev_buf (8 u64 per tile); NULL = off */
void siu3r_gemm_h3_debug(int mode);   /* timing experiments only: 1 = operand pipeline without MMAs, 2 = MMAs without operand loads (garbage results) */

#ifdef __cplusplus
}
#endif
#endif /* SIU3R_B200_H */
