/*
 * gg_b200.h -- C ABI of the B200-native Gaussian-splatting rasterizer (libgg_b200.so).
 *
 * Drop-in boundary for the one hot path of leejaehot/GaussianGrasper: the gsplat==0.1.0
 * operators its nerfstudio Gaussian model calls (reference requirements.txt:78;
 * nerfstudio/models/gaussian_splatting.py:46-50 imports, :699-784 calls).  gsplat's own FFI is a
 * torch C++ extension (gsplat.cuda: project_gaussians_forward, compute_sh_forward,
 * map_gaussian_to_intersects, get_tile_bin_edges, rasterize_forward, nd_rasterize_forward and
 * their backward twins); each entry point below names the binding it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; fp32 / int32 / int64 as typed;
 *     row-major, contiguous; no torch types cross this boundary;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises the device;
 *   - return 0 on success, <0 for an argument error, >0 for a cudaError_t; the message is
 *     available from gg_last_error_string() (thread-local); the library never throws or aborts;
 *   - tiles are 16x16 pixels (gaussian_splatting.py:677-682); pixel (row i, col j) is sampled at
 *     (x=j, y=i); quaternions are (w,x,y,z).
 */
#ifndef GG_B200_H
#define GG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library state ---------------------------------------------------------------------- */
#define GG_ABI_VERSION 200 /* bumped whenever a signature below changes; the bindings refuse another value */
int gg_version(void); /* returns GG_ABI_VERSION of the build */
const char* gg_last_error_string(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long gg_launch_count(void);
/* 0 iff the current CUDA device is compute capability 10.x */
int gg_check_device(void);

/* FP32 FMA throughput probe (bench.py's practical roof for the blend kernels): launches
 * `blocks` x 256 threads x `iters` x 8 FMAs; *flops (host pointer, nullable) receives the count. */
int gg_bench_fma(int blocks, int iters, float* out, double* flops, void* stream);

/* ---- projection: replaces gsplat.cuda.project_gaussians_forward / _backward ---------------
 * (ProjectGaussians.apply, gaussian_splatting.py:699-713).  viewmat: 12 floats (rows 0..2 of
 * the 4x4 world->camera matrix), fullmat: 16 floats (projmat @ viewmat).  Outputs are written
 * for every Gaussian; culled Gaussians get zeros. */
int gg_project_fwd(int n, const float* means, const float* scales, float glob_scale, const float* quats,
                   const float* viewmat, const float* fullmat, float fx, float fy, float cx, float cy, int img_h,
                   int img_w, int tiles_x, int tiles_y, float clip_thresh, float* cov3d /*[n,6]*/,
                   float* xys /*[n,2]*/, float* depths /*[n]*/, int32_t* radii /*[n]*/, float* conics /*[n,3]*/,
                   int32_t* num_tiles_hit /*[n]*/, void* stream);
int gg_project_bwd(int n, const float* means, const float* scales, float glob_scale, const float* quats,
                   const float* viewmat, const float* fullmat, float fx, float fy, float cx, float cy, int img_h,
                   int img_w, const int32_t* radii, const float* conics, const float* v_xys,
                   const float* v_depths /*nullable*/, const float* v_conics, float* v_means /*[n,3]*/,
                   float* v_scales /*[n,3]*/, float* v_quats /*[n,4]*/, void* stream);
/* multi-view variants: viewmats [V,12], fullmats [V,16], intrins [V,4]=(fx,fy,cx,cy) or NULL to
 * use the scalar arguments for every view; per-view outputs at [view*n + i].  The backward sums
 * the contributions of all V views (accumulate != 0 adds to the existing v_* contents). */
int gg_project_fwd_views(int n, int n_views, const float* means, const float* scales, float glob_scale,
                         const float* quats, const float* viewmats, const float* fullmats, const float* intrins,
                         float fx, float fy, float cx, float cy, int img_h, int img_w, int tiles_x, int tiles_y,
                         float clip_thresh, float* cov3d, float* xys, float* depths, int32_t* radii, float* conics,
                         int32_t* num_tiles_hit, void* stream);
int gg_project_bwd_views(int n, int n_views, const float* means, const float* scales, float glob_scale,
                         const float* quats, const float* viewmats, const float* fullmats, const float* intrins,
                         float fx, float fy, float cx, float cy, int img_h, int img_w, const int32_t* radii,
                         const float* conics, const float* v_xys, const float* v_depths, const float* v_conics,
                         int accumulate, float* v_means, float* v_scales, float* v_quats, void* stream);

/* ---- spherical harmonics: replaces gsplat.cuda.compute_sh_forward / _backward --------------
 * (SphericalHarmonics.apply, gaussian_splatting.py:730).  coeffs [n, (degree+1)^2, 3]. */
int gg_sh_fwd(int n, int degree, int degrees_to_use, const float* dirs /*[n,3]*/, const float* coeffs,
              float* colors /*[n,3]*/, void* stream);
int gg_sh_bwd(int n, int degree, int degrees_to_use, const float* dirs, const float* v_colors /*[n,3]*/,
              float* v_coeffs, void* stream);

/* ---- binning: replaces gsplat.utils.compute_cumulative_intersects / bin_and_sort_gaussians --
 * (torch.cumsum + gsplat.cuda.map_gaussian_to_intersects + torch.sort + gather +
 * gsplat.cuda.get_tile_bin_edges, run inside every rasterize forward). */
size_t gg_cumsum_workspace_bytes(long long n);
/* inclusive int32 prefix sum; *total_dev (nullable) receives the grand total */
int gg_cumsum(long long n, const int32_t* in, int32_t* out, int32_t* total_dev, void* workspace,
              size_t workspace_bytes, void* stream);
/* key = ((view*tiles + tile) << 32) | float_bits(depth); id = Gaussian index inside its view */
int gg_map_to_intersects(int n, int n_views, const float* xys, const float* depths, const int32_t* radii,
                         const int32_t* cum_tiles_hit, int tiles_x, int tiles_y, int64_t* keys, int32_t* ids,
                         void* stream);
/* same, with the pixel centres read from packed geo records ([V*n,8], x and y first) */
int gg_map_to_intersects_geo(int n, int n_views, const float* geo, const float* depths, const int32_t* radii,
                             const int32_t* cum_tiles_hit, int tiles_x, int tiles_y, int64_t* keys, int32_t* ids,
                             void* stream);
size_t gg_sort_workspace_bytes(long long m);
/* stable LSD radix sort of (key, value) pairs on the low `key_bits` bits.  The workspace must be
 * zero-filled once when it is allocated and is then owned by the library between calls (it
 * carries the look-back status words); one workspace per concurrent stream. */
int gg_sort_pairs(long long m, int key_bits, const int64_t* keys_in, const int32_t* vals_in, int64_t* keys_out,
                  int32_t* vals_out, void* workspace, size_t workspace_bytes, void* stream);
/* tile_ranges [num_tiles,2] = (start,end) of each tile's run in the sorted keys, (0,0) if empty */
int gg_tile_ranges(long long m, const int64_t* keys_sorted, long long num_tiles, int32_t* tile_ranges,
                   void* stream);

/* ---- depth-first binning (internal fast path; same sorted ids / tile ranges as the entry points
 * above, with 2-3 M-sized sort passes instead of 6-7): sort the V*n Gaussians by (view, depth),
 * emit their tile entries in that order with key = view*tiles + tile, stable-sort by that key. */
int gg_depth_keys(long long n, int n_views, const float* depths, int64_t* keys /*[V*n]*/, int32_t* rows /*[V*n]*/,
                  void* stream);
int gg_gather_counts(long long total, const int32_t* order, const int32_t* num_tiles_hit, int32_t* counts,
                     void* stream);
int gg_emit_tiles_sorted(int n, int n_views, const int32_t* order, const float* xys, int xy_stride /*2 or 8*/,
                         const int32_t* radii, const int32_t* cum_sorted, int tiles_x, int tiles_y, int64_t* keys,
                         int32_t* ids, void* stream);
int gg_tile_ranges_lowkey(long long m, const int64_t* keys_sorted, long long num_tiles, int32_t* tile_ranges,
                          void* stream);

/* The same depth-first binning as two host calls (every kernel above enqueued from C).
 *   gg_bin_begin : depth keys -> sort by (view, depth) -> per-Gaussian tile counts in that order ->
 *                  inclusive scan; the total M is copied to *m_host (pinned host memory) and an event is
 *                  recorded behind the copy.  Work enqueued after this call overlaps the read-back.
 *   gg_bin_wait  : blocks the host until *m_host is valid (one outstanding gg_bin_begin per device).
 *   gg_bin_finish: emission in depth order -> stable sort by tile -> tile ranges -> tile order (m >= 1).
 * `scratch` buffers are caller-owned, 256-byte aligned and ZERO-FILLED when allocated (the sort's look-back
 * words live in them); begin_scratch must stay untouched between gg_bin_begin and gg_bin_finish. */
size_t gg_bin_begin_scratch_bytes(long long total /* n * n_views */);
size_t gg_bin_finish_scratch_bytes(long long m);
int gg_bin_begin(int n, int n_views, const float* depths /*[V*n]*/, const int32_t* num_tiles_hit /*[V*n]*/,
                 void* scratch, size_t scratch_bytes, int32_t* m_host, void* stream);
int gg_bin_wait(void);
int gg_bin_finish(int n, int n_views, long long m, const float* xys, int xy_stride /*2 or 8*/, const int32_t* radii,
                  int tiles_x, int tiles_y, const void* begin_scratch, void* scratch, size_t scratch_bytes,
                  int32_t* ids_sorted /*[m]*/, int32_t* tile_ranges /*[V*tiles, 2]*/, int32_t* tile_order /*[V*tiles]*/,
                  void* stream);

/* Tile-first binning with no host read (the default of the fused path; same ids_sorted / tile_ranges as every
 * formulation above, bit for bit): per-tile counts -> scan (tile_ranges, the intersection count M and the
 * capacity check stay on the device) -> entries scattered into their tile's segment -> every tile sorted by
 * (depth bits, id) in shared memory.  `capacity` = entries ids_sorted (and the scratch) can hold; info
 * (device, nullable: kept in the scratch) / info_host (host, nullable; valid once the stream has
 * passed this call) receive {M, overflow, longest tile list, 0}.  A 16-byte aligned pinned info_host is written by
 * the scan kernel itself through its device mapping (no copy on the stream); any other host pointer gets a
 * stream-ordered cudaMemcpyAsync.  If M > capacity nothing is binned: overflow = 1,
 * every tile range is (0,0) (the blend kernels then render the background only) and the caller repeats the call
 * with capacity >= M.  capacity = 0 only counts.  The scratch needs no initialisation.
 * longest_hint: the longest tile list an earlier call reported (info[2]) or 0 if unknown; it only decides
 * which sort kernels are worth launching (lists beyond the launched classes take the slower general kernel). */
size_t gg_bin_tiles_scratch_bytes(int n_views, long long tiles_per_view, long long capacity);
int gg_bin_tiles(int n, int n_views, const float* xys, int xy_stride /*2 or 8*/, const float* depths /*[V*n]*/,
                 const int32_t* radii /*[V*n]*/, int tiles_x, int tiles_y, long long capacity, void* scratch,
                 size_t scratch_bytes, int32_t* ids_sorted /*[capacity]*/, int32_t* tile_ranges /*[V*tiles, 2]*/,
                 int32_t* tile_order /*[V*tiles]*/, int32_t* info /*[4] device, nullable*/,
                 int32_t* info_host /*[4] pinned, nullable*/, int longest_hint, void* stream);

/* visiting order of the tiles for the blend kernels: tile ids by descending list length (only
 * scheduling depends on it, never results); tile_order [num_tiles] int32 */
size_t gg_tile_order_workspace_bytes(void);
int gg_tile_order(long long num_tiles, const int32_t* tile_ranges, int32_t* tile_order, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ---- blending: replaces gsplat.cuda.rasterize_forward / nd_rasterize_forward and backward ---
 * (RasterizeGaussians / NDRasterizeGaussians, gaussian_splatting.py:735,747,759,773).
 * geo is the packed per-Gaussian record {x, y, A/2, B, C/2, opacity, tau, 0} built by gg_pack_geo.
 * colors rows are color_stride floats apart; `channels` <= gg_blend_max_channels() per launch
 * (split wider feature maps into channel ranges by offsetting colors/out/bg).  colors_per_view:
 * 0 = one [n, stride] table shared by all views, 1 = [V*n, stride]. */
int gg_blend_max_channels(void);
int gg_pack_geo(long long n, int n_views, const float* xys, const float* conics, const float* opac,
                int opac_per_view, float* geo /*[V*n,8]*/, void* stream);
int gg_blend_fwd(int n_views, long long n, int channels, int color_stride, int colors_per_view, int out_stride,
                 int img_h, int img_w, int tiles_x, int tiles_y, const int32_t* ids_sorted,
                 const int32_t* tile_ranges, const int32_t* tile_order /*nullable*/, const float* geo,
                 const float* colors, const float* bg,
                 float* out /*[V,H,W,out_stride]*/, float* final_T /*[V,H,W]*/, int32_t* final_idx /*[V,H,W]*/,
                 unsigned long long* pair_counter /*nullable*/, uint32_t* hit_words /*nullable*/, void* stream);
/* Workload counters of a binned batch (bench.py's roofline accounting): stats[0] += pixel-Gaussian pairs visited
 * up to each pixel's stop (K of SURVEY 8d), stats[1] += pairs blended.  stats: 2 device uint64, zeroed by the caller. */
int gg_blend_pair_stats(int n_views, long long n, int img_h, int img_w, int tiles_x, int tiles_y,
                        const int32_t* ids_sorted, const int32_t* tile_ranges, const float* geo /*[V*n,8]*/,
                        unsigned long long* stats, void* stream);
/* hit_words: table of gg_blend_hit_words(m, V*tiles, channels) uint32, ZERO-FILLED by the caller before
 * gg_blend_fwd, in which the forward records per (tile batch, 8x4-pixel warp) the entries that
 * contributed; handed to gg_blend_bwd (same channel count) it spares the backward the culling and
 * the alpha test of every other entry.  Results are identical with or without it. */
size_t gg_blend_hit_words(long long m, long long num_tiles, int channels);
/* v_geo [V*n,8] and v_colors (indexed like colors) are accumulated into: zero them first */
int gg_blend_bwd(int n_views, long long n, int channels, int color_stride, int colors_per_view, int out_stride,
                 int img_h, int img_w, int tiles_x, int tiles_y, const int32_t* ids_sorted,
                 const int32_t* tile_ranges, const int32_t* tile_order /*nullable*/, const float* geo,
                 const float* colors, const float* bg,
                 const float* final_T, const int32_t* final_idx, const float* v_out,
                 const uint32_t* hit_words /*nullable*/, float* v_geo, float* v_colors, void* stream);
/* v_geo -> v_xys [V*n,2], v_conics [V*n,3], v_opac [n] (summed over views; nullable) */
int gg_unpack_vgeo(long long n, int n_views, const float* v_geo, float* v_xys, float* v_conics, float* v_opac,
                   int accumulate_opac, void* stream);

/* ---- fused multi-view preparation (no upstream twin: it fuses what the reference model does in
 * ~12 separate launches per view, gaussian_splatting.py:699-731, :605-619, :742-780) ------------
 * Raw model parameters in: log-scales, un-normalised wxyz quats, opacity logits, SH coefficients
 * [n,(degree+1)^2,3], features [n,D].  Out, per (view, Gaussian): the packed geo record, the
 * channel row chan[V*n, cp] = {r,g,b (SH+0.5 clamped), depth, normal(3), feature(D), 0-pad},
 * depths, radii, num_tiles_hit.  cp is a multiple of 4, >= 7 + D.  scales_out [n,3] / quats_out
 * [n,4] (nullable) receive the activated inputs of the projection (used by the parity tests).
 * phase: 0 = everything; 1 = geometry only (geo, depths, radii, num_tiles_hit); 2 = channel rows
 * only, reusing phase 1's radii / depths -- the host reads the intersection count in between. */
int gg_prepare_views(int n, int n_views, int feat_dim, int cp, int degree, int degrees_to_use, const float* means,
                     const float* log_scales, const float* quats, const float* opacity_logit,
                     const float* sh_coeffs, const float* features, const float* viewmats, const float* fullmats,
                     const float* intrins, const float* positions, int img_h, int img_w, int tiles_x, int tiles_y,
                     float clip_thresh, float* geo, float* chan, float* depths, int32_t* radii,
                     int32_t* num_tiles_hit, float* scales_out, float* quats_out, int phase, void* stream);
/* exact backward, summed over the views: v_geo [V*n,8] = {v_x, v_y, v_A, v_B, v_C, v_opacity,.,.}
 * and v_chan [V*n,cp] in; gradients of the raw parameters out (overwritten). */
int gg_prepare_views_bwd(int n, int n_views, int feat_dim, int cp, int degree, int degrees_to_use,
                         const float* means, const float* log_scales, const float* quats,
                         const float* opacity_logit, const float* features, const float* viewmats,
                         const float* fullmats, const float* intrins, const float* positions, int img_h, int img_w,
                         const float* geo, const float* chan, const int32_t* radii, const float* v_geo,
                         const float* v_chan, float* v_means, float* v_log_scales, float* v_quats,
                         float* v_opacity_logit, float* v_sh_coeffs /*nullable iff v_rgb_views*/, float* v_features,
                         float* v_rgb_views /*nullable*/, void* stream);

/* Factored SH gradient for view-sharded training.  The gradient of the [nb,3] coefficients of a Gaussian is
 * sum over views of Y(dir_v) (outer) v_rgb_v, so ranks exchange only the clamp-masked colour gradient
 * v_rgb_views [V*N, 3] (what gg_prepare_views_bwd writes when v_rgb_views != NULL, skipping v_sh_coeffs)
 * together with the camera centres, and every rank rebuilds the full sum here:
 * positions [n_views,3], v_rgb_views [n_views*n,3] over ALL views of all ranks -> v_sh_coeffs [n,nb,3]. */
int gg_sh_grad_from_views(int n, int n_views, int degree, int degrees_to_use, const float* means,
                          const float* positions, const float* v_rgb_views, float* v_sh_coeffs, void* stream);

/* Gradient exchange of a view-sharded training step as one kernel over symmetric memory: all-gather of
 * the per-view colour-gradient slots (the SH gradient's factors, see gg_sh_grad_from_views) and a two-shot
 * all-reduce of the other leaf gradients, through the NVSwitch with multimem.st / multimem.ld_reduce when the
 * buffers have a multicast mapping (bucket_multicast / rgb_multicast non-NULL), else over the peers' mapped
 * addresses.  bucket_peers / rgb_peers: HOST arrays of `world` device addresses (this rank's own included) of the
 * same symmetric buffers; bucket_floats = size of the bucket; rgb_slot_floats = size of ONE rank's slot of the
 * [world, slot] gather buffer (both multiples of 4).  The caller places a cross-rank barrier before the launch
 * (every rank has written its bucket and its slot) and after it (every rank's slice has landed everywhere).
 * parts: 1 = the gather only, 2 = the reduction only, 3 = both (the two halves can run as two concurrent launches
 * so that the SH rebuild, which needs only the gathered factors, overlaps the reduction).
 * Replaces ncclAllGather + ncclAllReduce of distributed.FactoredExchange. */
int gg_nvls_exchange(int rank, int world, float* bucket_multicast /*nullable*/, float* const* bucket_peers,
                     long long bucket_floats, float* rgb_multicast /*nullable*/, float* const* rgb_peers,
                     long long rgb_slot_floats, int parts, void* stream);

/* ---- next rows: the streaming steps directly behind the backward -----------------
 * gg_adam_step: fused torch.optim.Adam (no amsgrad / weight decay) over a flat gradient buffer that
 * holds up to 8 ordered, disjoint segments (padding between them is skipped), each with its own parameter
 * tensor and learning rate (replaces the
 * reference's per-group optimizers, method_configs.py:618-664).  params/offsets/counts/lrs/steps/modes are HOST
 * arrays of n_segments entries; steps[s] is segment s's own 1-based update count (bias corrections), as every
 * reference group has its own torch.optim.Adam.  modes (nullable = all GG_ADAM_STEP) implement the trainer's
 * gradient accumulation (engine/trainer.py:466-481, method_configs.py:611: xyz / color / feature every 10 steps):
 * ACC_FIRST accum = grad (the step after zero_grad), ACC accum += grad, ACC_STEP update with accum + grad,
 * STEP update with grad, SKIP leave the segment alone.  accum_flat is laid out like grad_flat. */
#define GG_ADAM_STEP 0
#define GG_ADAM_ACC_FIRST 1
#define GG_ADAM_ACC 2
#define GG_ADAM_ACC_STEP 3
#define GG_ADAM_SKIP 4
int gg_adam_step(int n_segments, float* const* params, const long long* offsets, const long long* counts,
                 const float* lrs, const int* steps, const int* modes /*nullable*/, const float* grad_flat,
                 float* accum_flat /*nullable unless a mode accumulates*/, float* exp_avg_flat, float* exp_avg_sq_flat,
                 float beta1, float beta2, float eps, void* stream);
/* gg_densify_stats: GaussianSplattingModel.after_train (gaussian_splatting.py:373-393) from the
 * blend gradient table: xys_grad_norm += |d loss / d xy|, vis_counts += 1, max_2dsize =
 * max(., radius / max(H, W)) for the Gaussians visible in each view; first_call != 0 initialises
 * (norms of every Gaussian, counts of one) the way the model does on its first step. */
int gg_densify_stats(long long n, int n_views, const float* v_geo /*[V*n,8]*/, const int32_t* radii /*[V*n]*/,
                     int img_h, int img_w, int first_call, float* xys_grad_norm, float* vis_counts,
                     float* max_2dsize, void* stream);

/* ---- refinement on the device: densify (split / duplicate) + cull and the Adam-state surgery of
 * GaussianSplattingModel.refinement_after (nerfstudio/models/gaussian_splatting.py:396-546, 333-371).
 * gg_refine_plan decides per Gaussian (split, duplicate, and which of {itself, its split children, its duplicate}
 * survive the cull), scans the counts and copies totals = {kept originals, kept split parents, kept duplicates,
 * all split parents} to totals_host (pinned; valid once the stream is synchronised).  The output set is
 * [kept originals | kept split children, sample-major | kept duplicates] -- the reference's order -- with
 * n_out = totals[0] + n_split_samples * totals[1] + totals[2] rows.  gg_refine_apply gathers every per-Gaussian
 * array (parameters and Adam moments) into its new [n_out, row] array in one launch; `kinds` says how children
 * are formed: COPY (parent's row), MOMENT (zeros), MEANS (split children: mean + R(q/|q|)(exp(scale) * z), z =
 * samples[s * totals[3] + rank among all split parents]), LOG_SCALES (split parents and their children:
 * log(exp(s) / 1.6)).  means / log_scales / quats are the OLD parameter arrays.
 * samples == NULL: z comes from the counter-based generator Philox4x32-10 with counter (parent row, sample index,
 * step, 0) and key (seed low, seed high), Box-Muller on the four output words -- a pure function of its arguments,
 * so the ranks of a view-sharded run (identical parameters, all-reduced statistics) create identical children
 * without exchanging anything (the reference draws torch.randn per process, :491).  gg_philox_normals writes the
 * same numbers for given parent rows ([n_samples, count, 3]; parents == NULL means rows 0..count-1);
 * gg_philox4x32_10_host is the raw block function on the host (known-answer tests). */
typedef struct {
    float max_dim;              /* max(W, H) of the last rendered image (:415) */
    float densify_grad_thresh;  /* config.densify_grad_thresh */
    float densify_size_thresh;  /* config.densify_size_thresh */
    float split_screen_size;    /* config.split_screen_size (used when split_by_screen) */
    float cull_alpha_thresh;    /* config.cull_alpha_thresh */
    float cull_scale_thresh;    /* config.cull_scale_thresh (used when cull_by_scale) */
    float cull_screen_size;     /* config.cull_screen_size (used when cull_by_screen) */
    int do_densify;             /* step < stop_split_at and past the post-reset window (:404-408) */
    int split_by_screen;        /* step < stop_screen_size_at (:421) */
    int do_cull;                /* :456 */
    int cull_by_scale;          /* step > refine_every * reset_alpha_every (:471) */
    int cull_by_screen;         /* ... and step < stop_screen_size_at (:475) */
} gg_refine_config;
#define GG_REFINE_MAX_ARRAYS 24
#define GG_REFINE_COPY 0
#define GG_REFINE_MOMENT 1
#define GG_REFINE_MEANS 2
#define GG_REFINE_LOG_SCALES 3
size_t gg_refine_workspace_bytes(int n);
int gg_refine_plan(int n, const float* xys_grad_norm, const float* vis_counts, const float* max_2dsize /*nullable*/,
                   const float* log_scales, const float* opacity_logit, const gg_refine_config* cfg, void* workspace,
                   size_t workspace_bytes, int32_t* totals_host /*[4]*/, void* stream);
int gg_refine_apply(int n, int n_split_samples, const int32_t* totals /*[4] host*/, const void* plan_workspace,
                    int n_arrays, const float* const* src, float* const* dst, const int* row_floats, const int* kinds,
                    const float* means, const float* log_scales, const float* quats, const float* samples /*nullable*/,
                    unsigned long long seed, unsigned int step, void* scratch /* 9 B per output row + 512 */,
                    size_t scratch_bytes, void* stream);
int gg_philox_normals(long long count, const int32_t* parents /*nullable*/, int n_samples, unsigned long long seed,
                      unsigned int step, float* out, void* stream);
void gg_philox4x32_10_host(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

/* k-nearest-neighbour scale initialisation of the model (gaussian_splatting.py:259-263, :315-331): dist3 [n,3] =
 * ascending distances to the three nearest OTHER points (exact, brute force), log_scales [n,3] = log of their mean on
 * all three axes.  Either output may be NULL. */
int gg_knn3_scales(int n, const float* means, float* dist3 /*nullable*/, float* log_scales /*nullable*/, void* stream);

/* ---- per-pixel loss with its gradient in one pass (the L1 term and its masked variant,
 * gaussian_splatting.py:853-866; kind 2 = mean squared error).  pred/target/grad are n contiguous floats
 * ([..., channels]); mask (nullable) has one byte per pixel (n / channels), 0 = pixel ignored (zero loss and
 * gradient).  The mean runs over all n elements, or, when valid_pixels (device int32: number of unmasked
 * pixels) is given, over the valid pixels' elements only (:882).  loss[0] = weight * mean(...),
 * grad = d loss / d pred.  workspace: gg_pixel_loss_workspace_bytes() bytes, ZERO-FILLED when allocated. */
size_t gg_pixel_loss_workspace_bytes(void);
int gg_pixel_loss(long long n, int channels, const float* pred, const float* target, const uint8_t* mask /*nullable*/,
                  const int32_t* valid_pixels /*nullable, device*/, int kind /*1 = L1, 2 = L2*/, float weight, float* grad,
                  float* loss, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the other terms of get_loss_dict (gaussian_splatting.py:876-925), value and gradient in one pass each.
 * workspace (all three): gg_loss_workspace_bytes() bytes, ZERO-FILLED when allocated.
 * gg_geom_loss: loss[0] = w_depth * mean_mask |depth - gt_depth| (:880), loss[1] = w_normal * (1/2 MSE(normal, gt) +
 *   1/2 (1 - mean cos(normal, gt))) over the masked pixels (:879, cosine_similarity_loss :113-118).  image is the
 *   blended [n_pixels, stride] picture of render_views (depth = channel 3, normal = channels 4..6); gt_normal
 *   [n_pixels, 3]; mask one byte per pixel; valid_pixels = device int32 count of set mask bytes.  grad
 *   [n_pixels, grad_stride] receives channels 3..6 (written, or added when accumulate != 0).
 * gg_cosine_rows_loss: loss = sum_k w_k (1 - cos(a_k, b_k)), rows a_k = a + (a_index ? a_index[k] : k) * a_stride
 *   (dim floats), same for b; weight nullable (= 1).  The contrastive feature loss over sampled pixel pairs
 *   (:905-912: a = b = the feature channels of the image, index lists = the pairs, w = 1 / (segments * pairs of the
 *   segment)) and up_loss (:913-914: a = MLP outputs, b = sampled CLIP features).  grad_a / grad_b (nullable) are
 *   ADDED to, with atomics when the operand is indexed (a pixel can be sampled twice).
 * gg_param_regs: loss[0] = w_sh * mean ||sh[:, 1:, :]||_2 over (Gaussian, colour) (:920), loss[1] = w_scale * 0.1 *
 *   mean(max(max_i s_i / min_i s_i, max_gauss_ratio) - max_gauss_ratio), s = exp(log_scales) (:921-923); the
 *   gradients are ADDED to v_sh_coeffs / v_log_scales (nullable). */
size_t gg_loss_workspace_bytes(void);
int gg_geom_loss(long long n_pixels, int stride, const float* image, const float* gt_depth, const float* gt_normal,
                 const uint8_t* mask, const int32_t* valid_pixels, float w_depth, float w_normal, float* grad,
                 int grad_stride, int accumulate, float* loss /*[2]*/, void* workspace, size_t workspace_bytes, void* stream);
int gg_cosine_rows_loss(long long n_rows, int dim, const float* a, long long a_stride, const int64_t* a_index /*nullable*/,
                        const float* b, long long b_stride, const int64_t* b_index /*nullable*/,
                        const float* weight /*nullable*/, float* grad_a /*nullable*/, long long grad_a_stride,
                        float* grad_b /*nullable*/, long long grad_b_stride, float* loss /*[1]*/, int accumulate_loss,
                        void* workspace, size_t workspace_bytes, void* stream);
int gg_param_regs(long long n, int num_bases, const float* sh_coeffs, const float* log_scales, float max_gauss_ratio,
                  float w_sh, float w_scale, float* v_sh_coeffs /*nullable*/, float* v_log_scales /*nullable*/,
                  float* loss /*[2]*/, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the CLIP up-projection MLP(32 -> 128 -> 512) of the reference model applied to a whole feature map
 * (gaussian_splatting.py:198-213, :294; base_pipeline.py:408 in render.sh's flow): one fused tcgen05 / TMEM kernel,
 * 3xTF32 split with fp32 accumulation (fp32-equivalent results), hidden activations kept on the SM, output written
 * once.  gg_mlp_pack_weights splits W1 [128,32] and W2 [512,128] (torch.nn.Linear layout) into their TF32 halves in
 * the kernel's shared-memory layout (gg_mlp_packed_floats() floats; once per weight set).  gg_mlp_up: x rows of 32
 * floats, x_stride floats apart (e.g. the feature channels of render_views' image) -> y [n_rows, 512]. */
size_t gg_mlp_packed_floats(void);
int gg_mlp_pack_weights(const float* w1, const float* w2, float* packed, void* stream);
int gg_mlp_up(long long n_rows, const float* x, long long x_stride, const float* packed, const float* b1, const float* b2,
              float* y, void* stream);

/* ---- SSIM loss with its gradient: weight * (1 - SSIM) with pytorch_msssim's defaults as the
 * reference configures them (gaussian_splatting.py:284, :885: 11x11 Gaussian window, sigma 1.5, data_range 1, no
 * padding, mean over batch, channels and positions).  pred / target / grad are channel-last [n_img, H, W, stride]
 * (strides in floats per pixel); the first `channels` channels are compared.  accumulate != 0 adds to *loss and to
 * grad instead of overwriting them (main_loss = (1-l) * L1 + l * (1 - SSIM), :931). */
size_t gg_ssim_workspace_bytes(int n_img, int img_h, int img_w, int channels);
int gg_ssim_loss(int n_img, int img_h, int img_w, int channels, const float* pred, int pred_stride, const float* target,
                 int target_stride, float weight, float* grad, int grad_stride, int accumulate, float* loss,
                 void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GG_B200_H */
