/* gaussctrl_b200 – C ABI of the B200-native GaussCtrl hot path (libgaussctrl_b200.so).
 *
 * The reference (ActiveVisionLab/gaussctrl) has no FFI of its own: its hot path calls two third-party Python
 * operator sets (gsplat 0.1.3, diffusers 0.26.0).  Each entry point below replaces the native work behind one
 * of those Python call sites; the citation says which (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `h_` (host);
 *   - the caller owns every buffer, including workspaces sized by the *_workspace_bytes queries; nothing is
 *     allocated, freed or synchronised inside (exceptions are documented per function);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it (CUDA-graph capturable);
 *   - return value: 0 = ok, negative = error (GCB_ERR_*); `gcb_last_error()` returns a thread-local message;
 *   - fp16 tensors are IEEE binary16 ("half"); activations are channels-last: images [B,H,W,C], tokens [B,N,C].
 */
#ifndef GAUSSCTRL_B200_H
#define GAUSSCTRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCB_VERSION 100

#define GCB_ERR_INVALID (-1)     /* bad argument / unsupported shape */
#define GCB_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define GCB_ERR_UNSUPPORTED (-3) /* valid request, kernel variant not built */
#define GCB_ERR_WORKSPACE (-4)   /* workspace too small */

int gcb_version(void);
const char* gcb_last_error(void);
int gcb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ======================================================================================================
 * A. Diffusion side – replaces the cuDNN/cuBLAS/eager kernels diffusers launches from
 *    `self.pipe(...)` (gaussctrl/gc_pipeline.py:142-145 and :209-219) and `self.pipe.vae.encode`
 *    (gc_pipeline.py:244).
 * ====================================================================================================== */

/* Epilogue activation of gcb_conv2d_nhwc_fwd */
#define GCB_ACT_NONE 0
#define GCB_ACT_SILU 1
#define GCB_ACT_GEGLU 2 /* weights packed [value-half | gate-half] per N tile: see gcb_geglu_pack_rows */

/* GEMM implementation selector (debug / bisecting; the product default is GCB_GEMM_TCGEN05) */
#define GCB_GEMM_TCGEN05 0 /* TMA + tcgen05.mma + TMEM accumulators */
#define GCB_GEMM_MMA_SYNC 1 /* legacy mma.sync path kept for bring-up comparison */
#define GCB_GEMM_TCGEN05_DIRECT 2 /* tcgen05 main loop with the round-1 epilogue (per-thread row stores): A/B only */
#define GCB_GEMM_TCGEN05_PERSISTENT 3 /* force the persistent schedule (one CTA per SM, two TMEM accumulators): A/B only */
#define GCB_GEMM_TCGEN05_ONE_TILE 4 /* force the one-tile-per-CTA schedule: A/B only */

/* Implicit-GEMM convolution / linear layer, fp16 in/out, fp32 accumulate.
 *   x        [B,H,W,Cin]                 (Linear: B=1,H=1,W=M rows)
 *   w        [Cout, ksize*ksize*Cin]     tap-major, channel-minor ("OHWI")
 *   bias     [Cout] or NULL
 *   rowvec   [B, rowvec_ld] or NULL      per-image vector added to every pixel (ResnetBlock2D time_emb_proj)
 *   residual [B*H*W, Cout] or NULL       added after bias (+ rowvec)
 *   y        [B*H*W, Cout]               (GCB_ACT_GEGLU: [B*H*W, Cout/2])
 *   ksize    1 or 3 (stride 1, padding ksize/2).  Stride-2 convs go through gcb_im2col_nhwc + ksize=1.
 * Replaces: torch.nn.Conv2d / Linear inside UNet2DConditionModel, ControlNetModel, AutoencoderKL. */
int gcb_conv2d_nhwc_fwd(const void* x, const void* w, const void* bias, const void* rowvec, int rowvec_ld,
                        const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int act,
                        int impl, void* stream);

/* Direct convolution for tiny channel counts (conv_in 4->320, conv_out 320->4, ControlNet cond embedding,
 * VAE conv_in/conv_out).  Any Cin/Cout, ksize 1|3, stride 1|2, pad_lo/pad_hi (VAE downsample pads (0,1)). */
int gcb_conv2d_direct_nhwc_fwd(const void* x, const void* w, const void* bias, const void* residual, void* y, int B,
                               int H, int W, int Cin, int Cout, int ksize, int stride, int pad_lo, int pad_hi, int act,
                               void* stream);

/* im2col for 3x3 stride-2 convolutions (Downsample2D): x [B,H,W,C] -> col [B*Ho*Wo, 9*C]. */
int gcb_im2col3x3_s2_nhwc(const void* x, void* col, int B, int H, int W, int C, int pad_lo, int pad_hi, void* stream);
/* 3x3 / stride 1 / pad 1 patches of a 4-channel tensor x [B,H,W,4] -> col [B*H*W, 40] fp16 ((kh,kw,c) order, columns
 * 36..39 zero): conv_in of UNet2DConditionModel / ControlNetModel (latents, 4 channels) as a K = 40 tensor-core GEMM. */
int gcb_im2col3x3_c4_nhwc(const void* x, void* col, int B, int H, int W, void* stream);

/* Row permutation that turns a diffusers GEGLU projection [8C, C] into the tile-interleaved layout
 * GCB_ACT_GEGLU expects.  h_perm (host, int32[Cout]) receives source-row indices; tile_n is the N tile used. */
int gcb_geglu_tile_n(int Cout);
int gcb_geglu_pack_rows(int Cout, int32_t* h_perm);

/* GroupNorm (+ optional SiLU) over channels-last input, optionally over the channel-concatenation of two
 * tensors (UNet skip connections): y[B,HW,C1+C2] = act(GN(cat(x1, x2))).  fp32 statistics.
 * Replaces: torch.nn.GroupNorm + F.silu in ResnetBlock2D / Transformer2DModel.norm / conv_norm_out. */
size_t gcb_groupnorm_workspace_bytes(int B, int groups);
int gcb_groupnorm_nhwc_fwd(const void* x1, const void* x2, const void* gamma, const void* beta, void* y, int B, int HW,
                           int C1, int C2, int groups, float eps, int silu, void* workspace, size_t workspace_bytes,
                           void* stream);

/* LayerNorm over the last dim: y[M,C]. Replaces torch.nn.LayerNorm in BasicTransformerBlock. */
int gcb_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, int M, int C, float eps,
                      void* stream);

/* Multi-source attention = the fused form of CrossViewAttnProcessor (gaussctrl/utils.py:44-133, compute_attn :25-37):
 *     out[b] = sum_s weight[s] * softmax(q[b] K_s^T * scale) V_s          (independent softmax per source, per head)
 * q [B,Nq,heads*d]; K/V sources live in two buffers: (k,v) [Bkv,Nk,heads*d] and (k2,v2) [Bkv2,Nk,heads*d].
 * src_index (device int32 [B*n_src]): for query row b and source s, the batch row to read; values >= 0 index (k,v),
 * values < 0 index (k2,v2) at row -(value+1).  h_src_weight: host float[n_src]; zero-weight sources are skipped.
 * The reference's 5 passes are n_src=5, weights {c, (1-c)/4 x4}, src rows {b, ref0..ref3 of b's CFG half}
 * (utils.py:88-117); text cross-attention and the vanilla AttnProcessor are n_src=1.
 * ld_* are row strides in elements (>= heads*d) so q/k/v may be slices of a fused QKV projection.
 * v_head_stride: elements between consecutive heads inside a V row (both V buffers).  Normally d.  If it is >= d+8,
 * column d of every head must hold 1.0 (the caller's V projection writes it: zero weight rows + bias 1): the tcgen05
 * kernel then gets the softmax row sums out of the P V product itself instead of adding them up in registers, and at
 * head dim 40 evaluates a quarter of its exponentials as packed-half polynomials on the FMA pipes (no fp32
 * probabilities to sum up any more).  Results of the two layouts agree to the fp16 rounding of the probabilities. */
#define GCB_ATTN_AUTO 0     /* tcgen05 kernel where it is built for the shape, else the mma.sync kernel */
#define GCB_ATTN_TCGEN05 1
#define GCB_ATTN_MMA_SYNC 2
int gcb_attn_multi_fwd(const void* q, int ld_q, const void* k, const void* v, int ld_kv, const void* k2,
                       const void* v2, int ld_kv2, void* out, int ld_out, int B, int Nq, int Nk, int heads, int d,
                       int v_head_stride, int n_src, const int32_t* src_index, const float* h_src_weight, float scale,
                       int impl, void* stream);

/* Row softmax with scale, fp16 in/out, fp32 math (VAE mid-block attention, 1 head of dim 512). */
int gcb_softmax_rows_fwd(const void* x, void* y, int rows, int cols, float scale, void* stream);

/* Elementwise / data-movement helpers of the denoising loop */
int gcb_silu_fwd(const void* x, void* y, long long n, void* stream);
int gcb_add_fwd(const void* a, const void* b, void* y, long long n, float alpha, float beta, void* stream);
int gcb_geglu_fwd(const void* x, void* y, int M, int C, void* stream); /* x [M,2C] -> y [M,C] = x[:, :C]*gelu(x[:, C:]) */
int gcb_upsample_nearest2x_nhwc(const void* x, void* y, int B, int H, int W, int C, void* stream);
/* timesteps: DEVICE float[B] (so one captured CUDA graph serves every DDIM step) */
int gcb_timestep_embedding(const float* timesteps, int B, int dim, void* y /* fp16 [B,dim] */, void* stream);
int gcb_nchw_to_nhwc_f16(const void* x, void* y, int B, int C, int H, int W, void* stream);
int gcb_nhwc_to_nchw_f16(const void* x, void* y, int B, int C, int H, int W, void* stream);
int gcb_transpose_f16(const void* x, void* y, int batch, int rows, int cols, void* stream);

/* Classifier-free-guidance combine + DDIM step (eta = 0), replaces
 * `noise_pred_uncond + g*(noise_pred_text - noise_pred_uncond)` and `DDIMScheduler.step` inside pipe() (:209-219);
 * with eps_cond == NULL it is the plain (inverse) DDIM update used by DDIMInverseScheduler (:141-145).
 *   x' = sqrt(a_prev) * (x - sqrt(1-a_t) * eps) / sqrt(a_t) + sqrt(1-a_prev) * eps       (fp32 math, fp16 storage)
 * coef: DEVICE float[4] = {sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)} (graph-replayable per step). */
int gcb_cfg_ddim_step(const void* eps_uncond, const void* eps_cond, const void* x, void* x_out, long long n,
                      float guidance, const float* coef, void* stream);

/* (x/2+0.5).clamp(0,1) on the decoded image + optional mask composite edited*m + unedited*(1-m)
 * (gc_pipeline.py:223-234); img [B,H,W,3] fp16 NHWC -> out [B,H,W,3] fp32. mask [B,H,W] fp32 or NULL. */
int gcb_postprocess_composite(const void* img, const float* mask, const void* unedited_f16, float* out, int B, int H,
                              int W, void* stream);

/* disparity = (1/(depth+1e-5)) / max(...) replicated to 3 channels (gc_pipeline.py:248-266); depth fp32 [B,H,W]
 * -> fp16 [B,H,W,3]; the max is per image.  workspace: B floats. */
int gcb_depth_to_disparity(const float* depth, void* disp_f16, float* workspace, int B, int HW, int round_f16_first,
                           void* stream);

/* ======================================================================================================
 * B. Rasterisation side – replaces gsplat's CUDA extension behind
 *    project_gaussians (gc_model.py:140-154), spherical_harmonics (:166), rasterize_gaussians (:174-186,:191-202).
 * ====================================================================================================== */

/* gsplat.project_gaussians forward. h_viewmat: host float[16] row-major world->camera (only rows 0..2 used);
 * h_projmat: host float[16] full projection.  Outputs: xys [N,2], depths [N], radii [N] i32, conics [N,3],
 * num_tiles_hit [N] i32, cov3d [N,6] (cov3d may be NULL).  fp32, bit-exact vs the oracle (no FMA contraction). */
int gcb_project_gaussians_fwd(const float* means3d, const float* scales, float glob_scale, const float* quats,
                              const float* h_viewmat, const float* h_projmat, float fx, float fy, float cx, float cy,
                              int img_h, int img_w, int tile_bx, int tile_by, float clip_thresh, int N, float* xys,
                              float* depths, int32_t* radii, float* conics, int32_t* num_tiles_hit, float* cov3d,
                              void* stream);

/* gsplat.spherical_harmonics forward: viewdirs [N,3] (unit), coeffs [N,K,3], colors [N,3]. */
int gcb_sh_fwd(int degree, int K, const float* viewdirs, const float* coeffs, float* colors, int N, void* stream);
/* Fused eval-path front end of GaussCtrlModel.get_outputs (gc_model.py:138-167): exp(scales), quats / |quats|,
 * projection, view direction from h_cam_origin (host float[3]), SH colour (+0.5, clamp >= 0) and sigmoid(opacity) in
 * ONE pass over the 236 B/Gaussian splatfacto parameter record.  Outputs as gcb_project_gaussians_fwd (no cov3d) plus
 * rgbd [N,4] = (r, g, b, depth) - the 4-channel colour of the fused rgb+depth composite - and opac [N]. */
int gcb_project_sh_fused_fwd(const float* means3d, const float* log_scales, const float* quats,
                             const float* features_dc, const float* features_rest, const float* opacity_logits,
                             const float* h_viewmat, const float* h_projmat, const float* h_cam_origin, float fx,
                             float fy, float cx, float cy, int img_h, int img_w, int tile_bx, int tile_by,
                             int sh_degree, int N, float* xys, float* depths, int32_t* radii, float* conics,
                             int32_t* num_tiles_hit, float* rgbd, float* opac, void* stream);

/* Inclusive prefix sum (torch.cumsum inside gsplat.rasterize_gaussians); workspace via gcb_scan_workspace_bytes. */
size_t gcb_scan_workspace_bytes(int N);
int gcb_cumsum_i32(const int32_t* in, int32_t* out, int N, void* workspace, size_t workspace_bytes, void* stream);

/* Tile binning (gsplat cumsum + map_gaussian_to_intersects + torch.sort + get_tile_bin_edges, inside each
 * rasterize_gaussians call, gc_model.py:174-186 and :191-202), re-designed and WITHOUT the host round trip gsplat has
 * for the intersection count: order the N Gaussians by depth once (stable 8-bit radix passes, ties by id), emit their
 * tile intersections in that order, one stable radix pass by tile id (two above 2048 tiles).  The result equals a STABLE
 * sort of (tile_id << 32 | depth bits) keys.
 *   isect_capacity  room (in intersections) of gaussian_ids / isect_keys and of the workspace: the caller sizes it;
 *   gaussian_ids    [isect_capacity] i32: the first M entries are the sorted Gaussian ids;
 *   tile_bins       [tiles,2] i32 (start, end) of every tile's run (start == end for an empty tile);
 *   isect_count     device int32[2]: [0] = M (the true count), [1] = 1 when M > isect_capacity - nothing is written
 *                   out of bounds then, the result is truncated and the caller re-runs with a larger capacity;
 *   isect_keys      optional [isect_capacity] i64 = the sorted 64-bit keys (NULL to skip).
 * Every kernel reads M from `isect_count` on the device; the call never synchronises. */
size_t gcb_bin_gaussians_workspace_bytes(int N, long long isect_capacity, int tile_bx, int tile_by);
int gcb_bin_gaussians(const float* xys, const float* depths, const int32_t* radii, const int32_t* num_tiles_hit, int N,
                      int tile_bx, int tile_by, long long isect_capacity, int32_t* gaussian_ids, int32_t* tile_bins,
                      int32_t* isect_count, int64_t* isect_keys, void* workspace, size_t workspace_bytes, void* stream);

/* Per-tile front-to-back alpha compositing (gsplat rasterize_forward).  colors [N,C] with C in {1,3,4};
 * background host float[C].  Outputs: out_img [H,W,C], final_T [H,W], final_idx [H,W] i32.
 * C=4 with colors = (r,g,b,depth) is the fused rgb+depth pass that replaces the reference's TWO rasterize calls
 * (gc_model.py:174-202).  radii [N] i32 (project_gaussians' output, may be NULL): lets every warp skip Gaussians
 * that cannot reach alpha >= 1/255 anywhere in its 8x4 pixel sub-block; the image is bit-identical with or without it. */
int gcb_rasterize_fwd(const float* xys, const float* conics, const float* colors, const float* opacities,
                      const int32_t* gaussian_ids, const int32_t* tile_bins, const int32_t* radii, int img_h, int img_w,
                      int C, const float* h_background, float* out_img, float* final_T, int32_t* final_idx, void* stream);
/* The eval branch of GaussCtrlModel.get_outputs in ONE composite (gc_model.py:174-204 runs two rasterize calls and
 * three elementwise passes): rgbd [N,4] = (r,g,b,depth) per Gaussian, d_background3 DEVICE float[3] (no host read);
 * out_rgb [H,W,3] = min(rgb + T*bg, 1), out_alpha [H,W] = 1 - T, out_depth [H,W] = depth / alpha (1000 where alpha == 0). */
int gcb_rasterize_rgbd_fwd(const float* xys, const float* conics, const float* rgbd, const float* opacities,
                           const int32_t* gaussian_ids, const int32_t* tile_bins, const int32_t* radii, int img_h,
                           int img_w, const float* d_background3, float* out_rgb, float* out_depth, float* out_alpha,
                           void* stream);

/* V eval-mode renders of one scene in one call - the loop of render_reverse over the training views
 * (gc_pipeline.py:126-133 -> gc_model.py:57-206 per view): for each view fused project+SH, binning, fused rgb+depth
 * composite, stream-ordered on `stream`, sharing one set of intermediates in `workspace`.  Host arrays: h_viewmats /
 * h_projmats [V,16] row-major 4x4, h_cam_origins [V,3], h_intrinsics [V,4] = (fx, fy, cx, cy).  Outputs: out_rgb
 * [V,H,W,3], out_depth [V,H,W], out_alpha [V,H,W]; isect_counts device int32 [V,2] = (M, overflow) per view (see
 * gcb_bin_gaussians).  Never synchronises. */
size_t gcb_render_eval_batch_workspace_bytes(int N, long long isect_capacity, int img_h, int img_w);
int gcb_render_eval_batch(const float* means3d, const float* log_scales, const float* quats, const float* features_dc,
                          const float* features_rest, const float* opacity_logits, int N, int sh_degree, int V,
                          const float* h_viewmats, const float* h_projmats, const float* h_cam_origins,
                          const float* h_intrinsics, int img_h, int img_w, const float* d_background3,
                          long long isect_capacity, float* out_rgb, float* out_depth, float* out_alpha,
                          int32_t* isect_counts, void* workspace, size_t workspace_bytes, void* stream);

/* ---- reference-K/V exchange over NVLink peer memory (the one collective of the path; replaces the
 *      per-layer `ncclAllGather` a multi-GPU port of gc_pipeline.py:206-219 would issue).  One process per GPU; the
 *      host side exchanges the 64-byte IPC handles once (e.g. torch.distributed.all_gather_object).  All calls enqueue
 *      plain kernels on `stream` (CUDA-graph capturable, no host synchronisation). ---- */
typedef struct gcb_handle gcb_handle_t;
/* Bytes at the head of every arena reserved for flags / epochs: payload regions start at or after this offset. */
size_t gcb_handle_control_bytes(void);
/* Allocates this rank's arena (cudaMalloc on the current device, control block zeroed). */
int gcb_handle_create(int world, int rank, size_t arena_bytes, gcb_handle_t** out);
int gcb_handle_destroy(gcb_handle_t* handle);
/* Local arena base (device pointer): the gathered blocks are read in place from here. */
void* gcb_handle_arena(gcb_handle_t* handle);
int gcb_handle_ipc_export(gcb_handle_t* handle, unsigned char* out64);
int gcb_handle_ipc_open(gcb_handle_t* handle, int peer, const unsigned char* in64);
/* All-gather: local_src (bytes_per_rank bytes, 16-byte aligned) lands at arena_offset + rank * bytes_per_rank in EVERY
 * rank's arena; returns (in stream order) once every peer's block has arrived in the local arena.  `slot` (0..127)
 * names the flag set; use one slot per concurrently live buffer. */
int gcb_allgather_ref_kv(gcb_handle_t* handle, size_t arena_offset, const void* local_src, size_t bytes_per_rank, int slot,
                         void* stream);
/* The K/V projection and its exchange as ONE kernel: y = x w^T (+ bias), x [M,Cin] = this rank's rows, w [Cout,Cin]
 * (Cout % 64 == 0); the tcgen05 GEMM's epilogue stores every output tile into all ranks' arenas (TMA bulk stores over
 * NVLink to the peers) at arena_offset + rank * M * Cout * 2 - the layout gcb_allgather_ref_kv produces - and the flag
 * exchange follows.  Only output columns >= peer_col_min (a multiple of 64) are sent to the peers: with w = [to_q;to_k;to_v]
 * and peer_col_min = C the Q third stays local (only its own rank reads it).  At most 8 ranks. */
int gcb_linear_allgather_fwd(gcb_handle_t* handle, const void* x, const void* w, const void* bias, int M, int Cin, int Cout,
                             int peer_col_min, size_t arena_offset, int slot, void* stream);
/* Cross-rank barrier in stream order (guards the reuse of gathered blocks that peers still read). */
int gcb_peer_barrier(gcb_handle_t* handle, int slot, void* stream);
/* Synchronous: *out != 0 when a wait timed out (a peer never signalled); the gathered data is then invalid. */
int gcb_handle_error(gcb_handle_t* handle, int* out);

/* ---- backward (the 3DGS fine-tune step after the edit: gc_trainer.py:257-301 -> loss.backward() through gsplat's
 *      autograd Functions).  Exact derivatives of the forward kernels above. ---- */

/* gsplat rasterize_backward: v_out [H,W,C] and optional v_out_alpha [H,W] (NULL = no alpha gradient) ->
 * v_xy [N,2], v_conic [N,3], v_colors [N,C], v_opacity [N]; the caller zero-initialises them (accumulated with one
 * atomicAdd per warp and Gaussian).  v_conic is the true gradient w.r.t. the stored conic (a, b, c). */
int gcb_rasterize_bwd(const float* xys, const float* conics, const float* colors, const float* opacities,
                      const int32_t* gaussian_ids, const int32_t* tile_bins, int img_h, int img_w, int C,
                      const float* h_background, const float* final_T, const int32_t* final_idx, const float* v_out,
                      const float* v_out_alpha, float* v_xy, float* v_conic, float* v_colors, float* v_opacity,
                      void* stream);

/* gsplat project_gaussians backward: (v_xy, v_depth, v_conic) -> v_means3d [N,3], v_scales [N,3], v_quats [N,4]
 * (zero for Gaussians with radii == 0). */
int gcb_project_gaussians_bwd(const float* means3d, const float* scales, float glob_scale, const float* quats,
                              const float* h_viewmat, const float* h_projmat, float fx, float fy, float cx, float cy,
                              int img_h, int img_w, const int32_t* radii, const float* v_xy, const float* v_depth,
                              const float* v_conic, int N, float* v_means3d, float* v_scales, float* v_quats,
                              void* stream);

/* spherical_harmonics backward: v_coeffs [N,K,3] = basis_k(viewdir) * v_colors (zero beyond the active degree). */
int gcb_sh_bwd(int degree, int K, const float* viewdirs, const float* v_colors, float* v_coeffs, int N, void* stream);

/* get_outputs epilogue (gc_model.py:188,203-204): rgb = min(rgb,1); depth = depth/alpha where alpha>0 else 1000.
 * in: img4 [H,W,4] (r,g,b,depth-accum), final_T [H,W]; out: rgb [H,W,3], depth [H,W,1], alpha [H,W,1]. */
int gcb_raster_finalize(const float* img4, const float* final_T, float* rgb, float* depth, float* alpha, int HW,
                        void* stream);

/* ======================================================================================================
 * C. 3DGS fine-tune step that follows the edit – replaces the eager torch ops behind
 *    `loss_dict = model.get_loss_dict(...)`, `loss.backward()` and `optimizers.optimizer_scaler_step_some(...)`
 *    in GaussCtrlTrainer.train_iteration (gaussctrl/gc_trainer.py:257-301; get_train_loss_dict
 *    gaussctrl/gc_pipeline.py:276-287; Adam groups gaussctrl/gc_config.py:57-89).
 * ====================================================================================================== */

/* nerfstudio SplatfactoModel.get_loss_dict main_loss and its gradient in one call:
 *     main_loss = (1-ssim_lambda) * mean|gt-pred| + ssim_lambda * (1 - SSIM(gt, pred))
 * SSIM as pytorch_msssim (11-tap Gaussian sigma 1.5, valid blur, K=(0.01,0.03), data_range 1, mean over map).
 * pred, gt [H,W,C] fp32 channels-last (H, W >= 11).  loss_out: DEVICE float[3] = (main_loss, L1, ssim);
 * v_pred [H,W,C] = d main_loss / d pred.  Deterministic (no atomics).  No host sync. */
size_t gcb_l1_ssim_workspace_bytes(int H, int W, int C);
int gcb_l1_ssim_loss_fwd_bwd(const float* pred, const float* gt, int H, int W, int C, float ssim_lambda,
                             float* loss_out, float* v_pred, void* workspace, size_t workspace_bytes, void* stream);

/* torch.optim.Adam step (no weight decay, no amsgrad) over n_tensors fp32 tensors in one launch per 8 tensors.
 * h_* are HOST arrays of n_tensors entries: device pointers of param / grad / exp_avg / exp_avg_sq (16-byte
 * aligned), element counts, and the learning rate of each tensor (one nerfstudio optimizer group per tensor).
 * step counts from 1 (bias corrections 1-beta^step are computed on the host in double, as torch does). */
int gcb_adam_step(int n_tensors, void* const* h_params, const void* const* h_grads, void* const* h_exp_avg,
                  void* const* h_exp_avg_sq, const long long* h_numel, const double* h_lr, double beta1, double beta2,
                  double eps, int step, void* stream);

/* ======================================================================================================
 * D. Prompt side – the CLIP text encoder diffusers runs inside `self.pipe(prompt=...,
 *    negative_prompt=...)` (gaussctrl/gc_pipeline.py:142-145, :209-219 -> encode_prompt -> CLIPTextModel).
 *    Projections / LayerNorm reuse gcb_conv2d_nhwc_fwd / gcb_layernorm_fwd; these are the remaining pieces.
 * ====================================================================================================== */

/* out[b,t,:] = tok_emb[ids[b,t],:] + pos_emb[t,:]   (fp16 tables [vocab,C] / [T,C]; ids DEVICE int32 [B,T],
 * clamped to the table). */
int gcb_embed_tokens_f16(const int32_t* ids, const void* tok_emb, const void* pos_emb, void* out, int B, int T, int C,
                         int vocab, void* stream);

/* y = x * sigmoid(1.702 x) (CLIP "quick_gelu"), fp16, n a multiple of 8. */
int gcb_quick_gelu_fwd(const void* x, void* y, long long n, void* stream);

/* Causal (lower-triangular) self-attention for short sequences: q,k,v [B,T,heads*d] slices with row stride ld_qkv
 * (a fused q|k|v projection), out [B,T,heads*d] with row stride ld_out; T <= 128, d = 64; fp32 softmax. */
int gcb_attn_causal_fwd(const void* q, const void* k, const void* v, int ld_qkv, void* out, int ld_out, int B, int T,
                        int heads, int d, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAUSSCTRL_B200_H */
