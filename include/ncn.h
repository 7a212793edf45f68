/*
 * ncn.h - C ABI of libncn.so: the B200-native (sm_100a) replacement for the
 * per-step NeRF hot path of nikola3794/normal-clustering-nerf.
 *
 * Boundary.  The reference reaches its native code through three python
 * extension surfaces: the `vren` pybind module (models/csrc/binding.cpp:330-350,
 * 15 functions), tiny-cuda-nn's `Encoding`/`Network` (models/ngp_mt.py:70-155)
 * and faiss' `Kmeans` (losses.py:86-92).  Every entry point below names the
 * reference interface it replaces (file:line relative to the reference root).
 *
 * Conventions
 *   - plain C: device pointers + sizes + a stream; no torch / C++ types.
 *   - all pointers are DEVICE pointers unless the name ends in `_host`.
 *   - tensors are dense row-major ("C-contiguous"), exactly the layouts the
 *     reference's CHECK_INPUT (models/csrc/include/utils.h:4-6) enforces.
 *   - the library owns no tensor memory: outputs and scratch are caller
 *     allocated (sizes are given by the ncn_*_workspace_bytes helpers).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     never synchronises the device, and is re-entrant per stream.
 *   - return value: 0 = ok, <0 = argument error (NCN_E_*), >0 = cudaError_t.
 *     ncn_error_string() renders either.
 *   - integer / index outputs are bit-exact with the reference kernels; fp32
 *     marching outputs are bit-exact as well (same rounding sequence, see
 *     DESIGN.md); compositing / field outputs are within the tolerances stated
 *     in DESIGN.md and tests/.
 */
#ifndef NCN_H_
#define NCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCN_VERSION 100 /* 0.1.0 */

typedef void* ncn_stream_t; /* cudaStream_t */

enum {
  NCN_OK = 0,
  NCN_E_NULL = -1,      /* a required pointer is NULL */
  NCN_E_SIZE = -2,      /* a size / count argument is out of range */
  NCN_E_CONFIG = -3,    /* unsupported configuration (e.g. grid_size > 1024) */
  NCN_E_ALIGN = -4,     /* pointer alignment requirement violated */
  NCN_E_NCCL = -5,      /* NCCL failure (see ncn_comm_last_error) */
  NCN_E_UNSUPPORTED = -6
};

int ncn_version(void);
const char* ncn_error_string(int code);
/* number of SMs / compute capability of the current device (for grid sizing / checks) */
int ncn_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* developer aid: slots[slot] = %globaltimer (ns) when `stream` reaches this point (works inside a captured graph) */
int ncn_debug_stamp(uint64_t* slots, int slot, ncn_stream_t stream);
/* measurement aid: node census of a captured CUDA graph (cudaGraph_t as void*): counts4_host = [kernel, memcpy, memset, other] */
int ncn_graph_node_counts(void* cuda_graph, int* counts4_host);
/* Sample-arena overflow guard of a sync-free step.  The reference allocates its sample arrays exactly after a host sync
 * (raymarching.cu:302-305: counter -> torch::zeros({total_samples,...})); a step that keeps the count on the device marches into
 * a fixed `capacity`.  counter (2) i32 as written by ncn_march_train; state (3) i32 = [this step overflowed, number of overflowed
 * steps so far, largest sample count seen]; when counter[0] > capacity, poison[0] (the first gradient element, may be NULL) is set
 * to NaN so that the optimizer's non-finite path skips the update everywhere (and zeroes the gradient). */
int ncn_step_guard(const int32_t* counter, int64_t capacity, int32_t* state, float* poison, ncn_stream_t stream);

/* ------------------------------------------------------------------------- */
/* (1) ray / volume intersection          replaces vren.ray_aabb_intersect,  */
/*     vren.ray_sphere_intersect  (binding.cpp:12-41, intersection.cu:25-197) */
/* ------------------------------------------------------------------------- */
/* rays_o, rays_d (R,3) f32; centers, half_sizes (V,3) f32;
 * out: hit_cnt (R) i32, hits_t (R,max_hits,2) f32, hits_idx (R,max_hits) i64.
 * Unfilled slots are -1; slots are ordered by ascending t1 exactly like the
 * reference's torch::sort over hits_t[...,0] (-1 slots therefore sort first). */
int ncn_ray_aabb_intersect(const float* rays_o, const float* rays_d,
                           const float* centers, const float* half_sizes,
                           int64_t n_rays, int64_t n_boxes, int max_hits,
                           int32_t* hit_cnt, float* hits_t, int64_t* hits_idx,
                           ncn_stream_t stream);
/* radii (V) f32 */
int ncn_ray_sphere_intersect(const float* rays_o, const float* rays_d,
                             const float* centers, const float* radii,
                             int64_t n_rays, int64_t n_spheres, int max_hits,
                             int32_t* hit_cnt, float* hits_t, int64_t* hits_idx,
                             ncn_stream_t stream);
/* Fused fast path used by render(): one box, max_hits = 1, plus the near clamp
 * of models/rendering.py:28 (t1 in [0,near) -> near).  hits_t (R,1,2). */
int ncn_ray_aabb_near(const float* rays_o, const float* rays_d,
                      const float* center, const float* half_size, float near_distance,
                      int64_t n_rays, float* hits_t, ncn_stream_t stream);

/* ------------------------------------------------------------------------- */
/* (2) occupancy grid: Morton codes and bit packing                           */
/*     replaces vren.morton3D / morton3D_invert / packbits                    */
/*     (binding.cpp:44-75, raymarching.cu:35-161)                             */
/* ------------------------------------------------------------------------- */
int ncn_morton3d(const int32_t* coords /*(n,3)*/, int64_t n, int32_t* indices /*(n)*/,
                 ncn_stream_t stream);
int ncn_morton3d_invert(const int32_t* indices /*(n)*/, int64_t n, int32_t* coords /*(n,3)*/,
                        ncn_stream_t stream);
/* bit i of byte b = density_grid[8b+i] > threshold (LSB first); n_bytes = C*G^3/8 */
int ncn_packbits(const float* density_grid, int64_t n_bytes, float threshold,
                 uint8_t* density_bitfield, ncn_stream_t stream);
/* Fused update of models/ngp_mt.py:360-367 minus the mean: grid = grid<0 ? grid :
 * max(grid*decay, tmp).  Also accumulates sum/count of cells > 0 into
 * stats[0] (f32 sum) / stats[1] (f32 count) which the caller zeroed. */
int ncn_density_grid_update(float* density_grid, const float* density_tmp, int64_t n_cells,
                            float decay, float* stats, ncn_stream_t stream);
/* Sampling half of the occupancy-grid update (models/ngp_mt.py:254-271 sample_uniform_and_occupied_cells + the jittered
 * cell centres of :345-357) in one launch: m uniformly drawn cells followed by m cells drawn uniformly among the occupied
 * ones (occ_csum = inclusive cumsum (G^3) i32 of density_grid > threshold; upper_bound = searchsorted(right=True), clamped
 * to G^3-1) -> indices (2m) i32 Morton order, xyz (2m,3) f32 = cell centre * (s - s/G) + U(-1,1) * s/G.
 * Random numbers come from a counter hash of *seed_dev (device i64; ncn_grid_scatter_density advances it), so a captured
 * CUDA graph draws new cells on every replay.  The random stream is not the reference's torch stream (no parity there). */
int ncn_grid_sample_cells(const int32_t* occ_csum, int grid_size, int64_t m, float s, const int64_t* seed_dev,
                          int32_t* indices, float* xyz, ncn_stream_t stream);
/* density_tmp[indices[i]] = exp(h[i*h_stride]) for i < n (TruncExp forward of the density head output, f16);
 * seed_dev (may be NULL) is incremented once. */
int ncn_grid_scatter_density(const void* h_f16, int h_stride, const int32_t* indices, int64_t n, float* density_tmp,
                             int64_t* seed_dev, ncn_stream_t stream);
/* packbits with threshold = min(stats[0]/stats[1], density_threshold) read on device:
 * removes the .item() host sync of models/ngp_mt.py:365. */
int ncn_packbits_auto(const float* density_grid, int64_t n_bytes, const float* stats,
                      float density_threshold, uint8_t* density_bitfield, ncn_stream_t stream);

/* ------------------------------------------------------------------------- */
/* (3) ray marching        replaces vren.raymarching_train / raymarching_test */
/*     (binding.cpp:78-131, raymarching.cu:166-453)                           */
/* ------------------------------------------------------------------------- */
/* Scratch needed by ncn_march_train for n_rays rays. */
size_t ncn_march_train_workspace_bytes(int64_t n_rays, int max_samples);
/*
 * Train-time march.  One DDA pass per ray (the reference does two), a
 * scan of the per-ray counts and a coalesced warp-per-ray expansion:
 * ray r owns rays_a row r = [r, start_idx, N_samples] with start_idx the
 * exclusive prefix sum of N_samples in ray order - a deterministic instance of
 * the layouts the reference's atomics can produce (raymarching.cu:237-241).
 *
 * hits_t (R,2) [t1,t2]; noise (R) f32 in [0,1) or NULL (=0); bitfield u8.
 * Samples are written only while start_idx+k < capacity; counter[0] always
 * receives the true total so the caller can detect overflow and retry with a
 * larger arena; counter[1] = n_rays (as in the reference).
 * out: rays_a (R,3) i64; xyzs, dirs (capacity,3) f32; deltas, ts (capacity) f32;
 *      counter (2) i32.  Entries >= total are left untouched.
 */
int ncn_march_train(const float* rays_o, const float* rays_d, const float* hits_t,
                    const uint8_t* density_bitfield, int cascades, float scale,
                    float exp_step_factor, const float* noise, int grid_size,
                    int max_samples, int64_t n_rays, int64_t capacity,
                    int64_t* rays_a, float* xyzs, float* dirs, float* deltas, float* ts,
                    int32_t* counter, void* workspace, size_t workspace_bytes,
                    ncn_stream_t stream);
/* The same march split at the point where the reference surface must learn the
 * sample count on the host (custom_functions.py:91-96): _count runs the DDA and the
 * scan (rays_a and counter are final afterwards), the caller reads counter[0],
 * allocates exact-size arrays and calls _expand with the same workspace. */
int ncn_march_train_count(const float* rays_o, const float* rays_d, const float* hits_t,
                          const uint8_t* density_bitfield, int cascades, float scale,
                          float exp_step_factor, const float* noise, int grid_size,
                          int max_samples, int64_t n_rays, int64_t* rays_a, int32_t* counter,
                          void* workspace, size_t workspace_bytes, ncn_stream_t stream);
int ncn_march_train_expand(const float* rays_o, const float* rays_d, const int64_t* rays_a,
                           float exp_step_factor, float scale, int grid_size, int max_samples,
                           int64_t n_rays, int64_t capacity, float* xyzs, float* dirs,
                           float* deltas, float* ts, const void* workspace,
                           size_t workspace_bytes, ncn_stream_t stream);
/*
 * Test-time march: up to n_samples occupied steps per alive ray, resuming from
 * hits_t[r][0] which is advanced in place.  Reproduces the reference's
 * calc_dt(..., cascades) argument quirk (raymarching.cu:370,399).
 * alive_indices (A) i64; out xyzs, dirs (A,n_samples,3), deltas, ts (A,n_samples)
 * (padding slots are zero-filled by this call), n_eff (A) i32.
 */
int ncn_march_test(const float* rays_o, const float* rays_d, float* hits_t,
                   const int64_t* alive_indices, const uint8_t* density_bitfield,
                   int cascades, float scale, float exp_step_factor, int grid_size,
                   int max_samples, int n_samples, int64_t n_alive,
                   float* xyzs, float* dirs, float* deltas, float* ts, int32_t* n_eff,
                   ncn_stream_t stream);

/* ------------------------------------------------------------------------- */
/* (4) volume compositing   replaces vren.composite_train_fw/_multi_fw,       */
/*     composite_train_bw/_multi_bw, composite_test_fw/_multi_fw              */
/*     (binding.cpp:134-298, volumerendering.cu:16-585)                       */
/* ------------------------------------------------------------------------- */
/* developer A/B knob: lanes per ray of the training compositing kernels (4, 8, 16 [default]: sub-warp kernels for C = 3/6/9;
 * 32: one warp per ray).  Returns the previous value.  Results are identical for T / ws / total_samples in every mode. */
int ncn_set_composite_width(int lanes_per_ray);
/* sigmas (N), raws (N,C), deltas, ts (N) f32, rays_a (R,3) i64.
 * out: total_samples (R) i64, opacity, depth (R), rend (R,C), ws (N); all are
 * fully written for the rays_a rows given (rays with 0 samples get zeros; ws
 * beyond an early stop = 0).  `capacity` is the length of the sample arrays:
 * a ray whose [start,start+N) range runs past it is clipped (sync-free arena
 * mode, see ncn_march_train); pass N for exact-size arrays.  C <= 64. */
int ncn_composite_train_fw(const float* sigmas, const float* raws, const float* deltas,
                           const float* ts, const int64_t* rays_a, float T_threshold,
                           int64_t n_rays, int64_t capacity, int n_channels,
                           int64_t* total_samples, float* opacity, float* depth, float* rend,
                           float* ws, ncn_stream_t stream);
/* ncn_composite_train_fw + ncn_photometric_loss in ONE launch (C = 3, 6 or 9; NCN_E_UNSUPPORTED otherwise): the lane holding a
 * ray's final sums also evaluates rgb = rend[:3] + bg (1 - opacity) (models/rendering.py:231-241), the squared error against
 * target_rgb (losses.py:353) and the opacity entropy term (losses.py:357-361) and writes their gradients; arguments as in the two
 * separate calls.  sums[0] += sum sq err, sums[1] += sum entropy. */
int ncn_composite_train_fw_photometric(const float* sigmas, const float* raws, const float* deltas, const float* ts,
                                       const int64_t* rays_a, float T_threshold, int64_t n_rays, int64_t capacity,
                                       int n_channels, int64_t* total_samples, float* opacity, float* depth, float* rend,
                                       float* ws, const float* target_rgb, const float* bg_rgb_host, float opacity_w,
                                       float grad_scale, float* rgb_out, float* sums, float* dL_drend, float* dL_dopacity,
                                       ncn_stream_t stream);
/* the same with a target colour on the first n_gt_rays rays only (--random_tr_poses, losses.py:265-297: the rays of the generated
 * poses have no ground truth): the squared error is a mean over 3 n_gt_rays values and rays >= n_gt_rays get a zero colour
 * gradient; the opacity term stays a mean over all n_rays.  target_rgb (n_gt_rays,3); may be NULL when n_gt_rays == 0. */
int ncn_composite_train_fw_photometric_gt(const float* sigmas, const float* raws, const float* deltas, const float* ts,
                                          const int64_t* rays_a, float T_threshold, int64_t n_rays, int64_t capacity,
                                          int n_channels, int64_t* total_samples, float* opacity, float* depth, float* rend,
                                          float* ws, const float* target_rgb, int64_t n_gt_rays, const float* bg_rgb_host,
                                          float opacity_w, float grad_scale, float* rgb_out, float* sums, float* dL_drend,
                                          float* dL_dopacity, ncn_stream_t stream);
/* out: dL_dsigmas (N), dL_draws (N,C) fully written for the samples the rays_a rows
 * cover.  dL_dopacity / dL_ddepth / dL_drend / dL_dws may each be NULL (= zeros).
 * Either output pointer (not both) may be NULL to skip it: dL_draws depends on dL_drend only, so a caller
 * can start the colour-head backward before the depth / opacity gradients exist. */
int ncn_composite_train_bw(const float* dL_dopacity, const float* dL_ddepth,
                           const float* dL_drend, const float* dL_dws,
                           const float* sigmas, const float* raws, const float* ws,
                           const float* deltas, const float* ts, const int64_t* rays_a,
                           const float* opacity, const float* depth, const float* rend,
                           float T_threshold, int64_t n_rays, int64_t capacity, int n_channels,
                           float* dL_dsigmas, float* dL_draws, ncn_stream_t stream);
/* sigmas, deltas, ts (A,S); raws (A,S,C); in/out alive_indices (A) i64,
 * opacity, depth (R), rend (R,C); n_eff (A) i32.  hits_t is accepted by the
 * reference signature but unused by its kernel (volumerendering.cu:504-550). */
int ncn_composite_test_fw(const float* sigmas, const float* raws, const float* deltas,
                          const float* ts, int64_t* alive_indices, float T_threshold,
                          const int32_t* n_eff, int64_t n_alive, int n_samples, int n_channels,
                          float* opacity, float* depth, float* rend, ncn_stream_t stream);

/* ------------------------------------------------------------------------- */
/* (5) distortion loss      replaces vren.distortion_loss_fw / _bw            */
/*     (binding.cpp:301-327, losses.cu:7-172)                                 */
/* ------------------------------------------------------------------------- */
int ncn_distortion_fw(const float* ws, const float* deltas, const float* ts,
                      const int64_t* rays_a, int64_t n_rays, int64_t n_samples,
                      float* loss /*(R)*/, float* ws_inclusive /*(N)*/, float* wts_inclusive /*(N)*/,
                      ncn_stream_t stream);
int ncn_distortion_bw(const float* dL_dloss, const float* ws_inclusive, const float* wts_inclusive,
                      const float* ws, const float* deltas, const float* ts,
                      const int64_t* rays_a, int64_t n_rays, int64_t n_samples,
                      float* dL_dws /*(N)*/, ncn_stream_t stream);

/* segmented (CSR) row sum: replaces torch_scatter.segment_csr as used by
 * RayMarcher.backward (models/custom_functions.py:102-112).
 * src (N,D) f32, indptr (S+1) i64 -> out (S,D). */
int ncn_segment_csr_sum(const float* src, const int64_t* indptr, int64_t n_segments, int dim,
                        float* out, ncn_stream_t stream);

/* ------------------------------------------------------------------------- */
/* (6) multiresolution hash grid   replaces tcnn.Encoding (Grid/Hash/Linear)  */
/*     call site models/ngp_mt.py:70-82; semantics: SURVEY.md Appendix B      */
/* ------------------------------------------------------------------------- */
#define NCN_GRID_MAX_LEVELS 32
typedef struct ncn_grid_desc {
  int32_t n_levels;             /* L (16)           */
  int32_t n_features;           /* F (2; 1,2,4,8)   */
  int32_t log2_hashmap_size;    /* log2 T (19)      */
  int32_t base_resolution;      /* N_min (16)       */
  float per_level_scale;        /* b                */
  /* derived by ncn_grid_desc_init: */
  float level_scale[NCN_GRID_MAX_LEVELS];     /* exp2f(l*log2f(b))*N_min - 1  */
  uint32_t level_res[NCN_GRID_MAX_LEVELS];    /* ceilf(scale)+1               */
  uint32_t level_size[NCN_GRID_MAX_LEVELS];   /* entries in the level         */
  uint32_t level_offset[NCN_GRID_MAX_LEVELS + 1]; /* entry offset, [L] = total */
} ncn_grid_desc;
/* host-only helper: fill the derived fields; returns total #params = entries*F */
int64_t ncn_grid_desc_init(ncn_grid_desc* desc);
/* x (N,3) f32 in [0,1]; table fp16 (entries,F); out (N, L*F) fp16 (row major).
 * xform_host: NULL, or 6 HOST floats (lo[3], size[3]): the kernel encodes (x - lo) / size
 * (the normalisation of models/ngp_mt.py:166 fused in).
 * n_dev: NULL, or a DEVICE int32 holding the live row count (rows processed = min(n, *n_dev));
 * lets a sync-free caller size its arrays by capacity (see ncn_march_train). */
int ncn_grid_fwd(const ncn_grid_desc* desc_host, const float* x, const void* table_f16,
                 int64_t n, void* out_f16, const float* xform_host, const int32_t* n_dev,
                 ncn_stream_t stream);
/* dL_dy (N,L*F) fp16 -> grad_table fp32 (entries*F), ACCUMULATED (caller zeroes). */
int ncn_grid_bwd(const ncn_grid_desc* desc_host, const float* x, const void* dL_dy_f16,
                 int64_t n, float* grad_table_f32, float grad_scale, const float* xform_host,
                 const int32_t* n_dev, ncn_stream_t stream);
/* ncn_grid_bwd restricted to levels [level_begin, level_end) (F = 2 tables; the gradient regions of consecutive levels are
 * contiguous, offsets in the desc), with at most ctas_per_sm (1..8) CTAs per SM - a data-parallel caller all-reduces the
 * first range while the second is still being computed. */
int ncn_grid_bwd_levels(const ncn_grid_desc* desc_host, const float* x, const void* dL_dy_f16, int64_t n, float* grad_table,
                        float grad_scale, const float* xform_host, const int32_t* n_dev, int level_begin, int level_end,
                        int ctas_per_sm, ncn_stream_t stream);
/* ncn_grid_bwd into an fp16 gradient table (entries x __half2, F = 2 only; ACCUMULATED, caller zeroes): the accumulation
 * type tiny-cuda-nn itself uses for F > 1 (grid.h, `grad_t`): run sums are still formed in fp32 registers, the reductions are
 * packed red.global.add.noftz.f16x2 / .v2.f16x2 (4 / 8 bytes instead of 8 / 16).  The caller's loss scale must keep the sums
 * inside fp16 range.  A selectable mode, measured by tools/sweep_hashgrid.py; the training step uses the fp32 table. */
int ncn_grid_bwd_f16(const ncn_grid_desc* desc_host, const float* x, const void* dL_dy_f16, int64_t n, void* grad_table_f16,
                     float grad_scale, const float* xform_host, const int32_t* n_dev, ncn_stream_t stream);
/* developer A/B knob of the F = 2 table backward: CTAs per SM its kernel is compiled for, 5 (default) or 6; returns the old value */
int ncn_set_grid_bwd_occupancy(int ctas_per_sm);
/* 1 (default): ncn_grid_bwd merges same-entry contributions of consecutive samples inside a warp before the
 * scatter; 0: one reduction per corner.  Returns the old value. */
int ncn_set_grid_bwd_merge(int on);
/* 1 (default): ncn_grid_fwd (L=16, F=2) evaluates one level for 32 consecutive samples per warp (coherent gathers);
 * 0: one thread per (sample, level).  Returns the old value. */
int ncn_set_grid_fwd_coherent(int on);
/* dL_dx (N,3) f32 = d out / d x contracted with dL_dy. */
int ncn_grid_bwd_input(const ncn_grid_desc* desc_host, const float* x, const void* table_f16,
                       const void* dL_dy_f16, int64_t n, float* dL_dx, ncn_stream_t stream);
/* double backward of the input gradient (tcnn bwd_bwd_input): given dL_ddLdx (N,3)
 * accumulates into grad_table (if non-NULL) and writes dL_ddLdy (N,L*F) f16 (if non-NULL).
 * (Linear interpolation => the second derivative wrt x itself is zero.) */
int ncn_grid_bwd_bwd_input(const ncn_grid_desc* desc_host, const float* x, const void* table_f16,
                           const float* dL_ddLdx, const void* dL_dy_f16, int64_t n,
                           float* grad_table_f32, void* dL_ddLdy_f16, ncn_stream_t stream);

/* ------------------------------------------------------------------------- */
/* (7) fully fused MLPs   replaces tcnn.Network (FullyFusedMLP, width 64)     */
/*     call sites models/ngp_mt.py:83-155                                     */
/* ------------------------------------------------------------------------- */
enum { NCN_ACT_NONE = 0, NCN_ACT_RELU = 1, NCN_ACT_SIGMOID = 2, NCN_ACT_EXP = 3 };
typedef struct ncn_mlp_desc {
  int32_t n_in;          /* logical input width (padded up to x16 with 1.0) */
  int32_t n_out;         /* logical output width (padded up to x16)         */
  int32_t n_hidden;      /* number of hidden layers (1 or 2..)              */
  int32_t width;         /* 64                                              */
  int32_t activation;    /* hidden: NCN_ACT_RELU                            */
  int32_t out_activation;/* NCN_ACT_NONE / NCN_ACT_SIGMOID                  */
} ncn_mlp_desc;
int64_t ncn_mlp_n_params(const ncn_mlp_desc* d);
size_t ncn_mlp_bwd_workspace_bytes(const ncn_mlp_desc* d, int64_t n);
size_t ncn_mlp_acts_bytes(const ncn_mlp_desc* d, int64_t n);
/* x (N, n_in_pad) f16, weights f16 (tcnn layout: consecutive (out,in) row-major
 * matrices: (64,in_pad), (n_hidden-1) x (64,64), (out_pad,64)); out (N, n_out_pad) f16.
 * If `acts` != NULL (ncn_mlp_acts_bytes(d, n) bytes, 128 B aligned) the post-activation hidden
 * states are kept for the backward pass - an opaque buffer between ncn_mlp_fwd / ncn_field_fwd
 * and ncn_mlp_bwd*: per layer ceil(n/128) tiles of 128 rows, each tile stored as the
 * [feature/8][row][8] f16 panel the tcgen05 backward multiplies from, so that a tile is one
 * contiguous 16 KB bulk copy (element (r, f) of layer l at
 * l*ceil128(n)*64 + (r/128)*8192 + ((f/8)*128 + r%128)*8 + f%8).
 * Padded dims are multiples of 16 and <= 64.
 * n_dev: NULL or a DEVICE int32 live row count (rows = min(n, *n_dev)); the layer stride
 * stays that of n. */
int ncn_mlp_fwd(const ncn_mlp_desc* d, const void* x_f16, const void* w_f16, int64_t n,
                void* out_f16, void* acts_f16, const int32_t* n_dev, ncn_stream_t stream);
/* dL_dout (N,n_out_pad) f16 -> grad_w f32 += grad_scale * dL/dW (ACCUMULATED; may be
 * NULL), dL_dx (N,n_in_pad) f16 or NULL (in the units of dL_dout, not scaled).
 * scratch: ncn_mlp_bwd_workspace_bytes(d, n) bytes, 16 B aligned. */
int ncn_mlp_bwd(const ncn_mlp_desc* d, const void* x_f16, const void* w_f16,
                const void* out_f16, const void* acts_f16, const void* dL_dout_f16, int64_t n,
                float* grad_w_f32, void* dL_dx_f16, float grad_scale, void* scratch,
                size_t scratch_bytes, const int32_t* n_dev, ncn_stream_t stream);

/* ncn_mlp_bwd whose dL/dout rows are assembled on the fly (tcgen05 implementation, n_out_pad == 16):
 *   mode 1 (colour head): dL/dout[:, j] = d_raws[:, c_off + j] * scale for j < n_ch, 0 otherwise   (= ncn_field_head_dout)
 *   mode 2 (density trunk): dL/dh = dx_rgb[:, 3:19] (+ dx_extra) + e0 * d_sigmas * exp(clamp(h[:,0],-15,15)) * scale
 *                           (= ncn_field_bwd_h; dx_extra = dL/dh of a further head that reads h, e.g. sem_net, ngp_mt.py:217-224) */
typedef struct ncn_mlp_bwd_src {
  int32_t mode;
  const float* d_raws;    /* (N, c_total) f32 */
  int32_t c_total, c_off, n_ch;
  const void* dx_rgb;     /* (N, 32) f16 */
  const float* d_sigmas;  /* (N) f32 */
  const void* h;          /* (N, 16) f16 */
  float scale;
  int32_t perm;           /* bit 0: this net's input columns are in the fused-forward order [h(16) | d(3) | 1(13)] (colour head):
                             W0 columns and dW0 are re-indexed accordingly and dL/dx comes out in that order;
                             bit 1 (mode 2): dx_rgb is in that order, i.e. dL/dh = dx_rgb[:, 0:16] */
  const void* dx_extra;   /* mode 2: NULL or (N, 16) f16 added to dL/dh (net without output activation only) */
} ncn_mlp_bwd_src;
int ncn_mlp_bwd_src_fused(const ncn_mlp_desc* d, const ncn_mlp_bwd_src* src, const void* x_f16, const void* w_f16,
                          const void* out_f16, const void* acts_f16, int64_t n, float* grad_w_f32, void* dL_dx_f16,
                          float grad_scale, void* scratch, size_t scratch_bytes, const int32_t* n_dev,
                          ncn_stream_t stream);

/* Fused field forward for the RGB+density configuration (models/ngp_mt.py:157-229): hash-grid gather -> sigma net ->
 * TruncExp -> [h | d/|d| | 1] -> rgb net (sigmoid), one kernel.  x, dirs (N,3) f32; L=16, F=2 grid; w_sigma (32->64->16),
 * w_rgb (32->64->64->16) in the tcnn layout.  Outputs: sigmas (N) f32, raws[:, 0:3] (row stride c_total) f32, and - each
 * optional (NULL) - what the backward needs: feat (N,32), h (N,16), sig_acts (ncn_mlp_acts_bytes), x_rgb (N,32) in the order
 * [h | d | 1] (see ncn_mlp_bwd_src.perm), rgb_acts (ncn_mlp_acts_bytes), rgb_out (N,16), all f16. */
int ncn_field_fwd(const ncn_grid_desc* desc_host, const float* x, const float* dirs, const void* table_f16,
                  const void* w_sigma_f16, const void* w_rgb_f16, int64_t n, const int32_t* n_dev,
                  const float* xform_host, float* sigmas, float* raws, int c_total, void* feat_f16, void* h_f16,
                  void* sig_acts_f16, void* x_rgb_f16, void* rgb_acts_f16, void* rgb_out_f16, ncn_stream_t stream);
/* The same without the encoder: density trunk + colour head in ONE launch on precomputed features feat (N,32) f16
 * (replaces ncn_mlp_fwd(sigma) -> ncn_field_prepare_rgb -> ncn_mlp_fwd(rgb) -> ncn_field_head_out). */
int ncn_field_mlp_fwd(const void* feat_f16, const float* dirs, const void* w_sigma_f16, const void* w_rgb_f16, int64_t n,
                      const int32_t* n_dev, float* sigmas, float* raws, int c_total, void* h_f16, void* sig_acts_f16,
                      void* x_rgb_f16, void* rgb_acts_f16, void* rgb_out_f16, ncn_stream_t stream);

/* The extra heads on h (models/ngp_mt.py:104-140, 217-224: sem_net / norm_net, 16 -> 64 -> 64 -> n_out <= 16, ReLU hidden,
 * no output activation), evaluated together in ONE launch: h (N,16) f16 is read once, outputs are written straight into
 * raws[:, c_off : c_off + n_ch] (f32, row stride c_total; the channel layout of rendering.py:203-208).  Head a / head b:
 * w NULL = absent; acts (ncn_mlp_acts_bytes of the net) and out (N,16) f16 are optional (NULL when no backward follows).
 * Replaces 2 x (ncn_mlp_fwd + ncn_field_head_out). */
int ncn_field_heads_fwd(const void* h_f16, int64_t n, const int32_t* n_dev, float* raws, int c_total,
                        const void* w_a_f16, int c_off_a, int n_ch_a, void* acts_a_f16, void* out_a_f16,
                        const void* w_b_f16, int c_off_b, int n_ch_b, void* acts_b_f16, void* out_b_f16,
                        ncn_stream_t stream);

/* Selects the ncn_mlp_bwd implementation: 1 (default) = every GEMM (dgrad and wgrad) on tcgen05 with TMEM
 * accumulators, 128-row tiles fed by bulk copies (shapes without an instantiation fall back to 0);
 * 0 = warp-MMA dgrad in registers + split-K wgrad kernels.  Returns the old value. */
int ncn_set_mlp_bwd_impl(int impl);
/* implementation of ncn_field_mlp_fwd: 1 = tcgen05 / TMEM accumulators, one 128-sample tile per CTA iteration (default);
 * 0 = warp-level mma.sync.  Returns the previous value (developer A/B knob; results agree to fp16 rounding of the activations). */
int ncn_set_field_fwd_impl(int impl);
/* Programmatic dependent launch along the training step's serial kernel chain (encoder -> field MLPs -> compositing -> cluster
 * chain -> ... -> table backward): 1 (default) = each of those kernels is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, sets up (weights, barriers, TMEM, first loads of tensors that older kernels
 * produced) while its stream predecessor drains and executes griddepcontrol.wait before it touches the predecessor's output;
 * 0 = plain stream order.  Results are identical.  Returns the previous value. */
int ncn_set_pdl(int on);
/* ncn_march_train* on the constant-step path (cascades == 1, exp_step_factor == 0): 1 (default) = four lanes per ray, each
 * marching a quarter of the candidate sequence (bit-identical output); 0 = one lane per ray; 2 = four lanes with every
 * segment re-marched from its predecessor's landing point (test mode for the repair path).  Returns the old value. */
int ncn_set_march_segments(int mode);

/* Elementwise glue of the NGPMT field (models/ngp_mt.py:157-229, rendering.py:203-212) between the
 * encoder / MLP kernels; every function takes the device-side live row count n_dev (may be NULL).
 *   prepare_rgb: x_rgb (N,32) f16 = [d/||d|| (3), h (16), 1.0 x13], sigmas (N) f32 = exp(h[:,0])
 *   head_out   : raws[:, c_offset:c_offset+n_ch] (row stride c_total, f32) = out_f16[:, :n_ch]
 *   head_dout  : dout_f16 (N,out_pad) = [dL_draws[:, c_offset:+n_ch] * scale, 0...]
 *   bwd_h      : dh (N,16) f16 = dx_rgb[:, 3:19] (+ dx_a + dx_b) + e0 * dL_dsigmas * exp(clamp(h0,-15,15)) * scale */
int ncn_field_prepare_rgb(const float* dirs, const void* h_f16, int64_t n, const int32_t* n_dev,
                          void* x_rgb_f16, float* sigmas, ncn_stream_t stream);
int ncn_field_head_out(const void* out_f16, int out_pad, int64_t n, const int32_t* n_dev,
                       float* raws, int c_total, int c_offset, int n_ch, ncn_stream_t stream);
int ncn_field_head_dout(const float* dL_draws, int c_total, int c_offset, int n_ch, float scale,
                        int64_t n, const int32_t* n_dev, void* dout_f16, int out_pad,
                        ncn_stream_t stream);
/* rays from (image, pixel) indices: rays_d = directions[pix] @ poses[img][:, :3]^T, rays_o = poses[img][:, 3]
 * (NeRFSystem.forward gather + get_rays: train_nerf.py:167-182, datasets/ray_utils.py:46-71).
 * poses (P,3,4) f32, directions (HW,3) f32, img_idx / pix_idx (n) i64 -> rays_o, rays_d (n,3) f32. */
int ncn_rays_from_pixels(const float* poses, const float* directions, const int64_t* img_idx,
                         const int64_t* pix_idx, int64_t n, float* rays_o, float* rays_d,
                         ncn_stream_t stream);
int ncn_field_bwd_h(const void* dx_rgb_f16, const float* dL_dsigmas, const void* h_f16,
                    const void* dx_a_f16, const void* dx_b_f16, float scale, int64_t n,
                    const int32_t* n_dev, void* dh_f16, ncn_stream_t stream);

/* ------------------------------------------------------------------------- */
/* (8) normals from rendered depth + Manhattan clustering loss                */
/*     replaces datasets/hypersim_src/utils.py:505-541, faiss.Kmeans          */
/*     (losses.py:86-92) and losses.py:97-166, 441-509                        */
/* ------------------------------------------------------------------------- */
/* P = origin + dir*depth; n = normalize(cross(P2-P1, P3-P1)), eps 1e-12.
 * origin/dir (R,3), depth (R), idx1/2/3 (M) i64 -> normals (M,3). */
int ncn_normals_from_depth_fw(const float* origin, const float* dir, const float* depth,
                              const int64_t* idx1, const int64_t* idx2, const int64_t* idx3,
                              int64_t n_tri, float* normals, ncn_stream_t stream);
/* dL_dnormals (M,3) -> dL_ddepth (R) ACCUMULATED via atomics (caller zeroes). */
int ncn_normals_from_depth_bw(const float* origin, const float* dir, const float* depth,
                              const int64_t* idx1, const int64_t* idx2, const int64_t* idx3,
                              const float* dL_dnormals, int64_t n_tri, float* dL_ddepth,
                              ncn_stream_t stream);

/* Training-batch sampling on the device: the index half of BaseDataset.__getitem__ (datasets/base.py:94-173)
 * and the target gather (:175-183).  strategy 0 = all_images_triang_patch, 1 = same_image_triang_patch (patch_size^2 rays per
 * patch: pix = corner_INDEX + dy*W + dx - the reference adds the offsets to the index into valid_idx['patch_corners'],
 * base.py:164-166, reproduced), 2 = all_images_triang, 3 = same_image_triang (3 rays per triangle: x1, x1 - W, x1 - 1).
 * seed_dev: device int64 read for this batch and advanced by one (graph replays draw fresh batches).  Rays beyond the last
 * whole patch / triangle get (0, 0).  The numpy MT19937 stream is not reproduced; index arithmetic and distributions are.
 * gather_pixels: out[r, :] = table[img_idx[r], pix_idx[r], :] with rows of words_per_pixel 4-byte words (rgb f32 x3, labels ...). */
int ncn_sample_ray_batch(int strategy, int64_t* seed_dev, int n_rays, int n_poses, int height, int width, int patch_size,
                         int64_t* img_idx, int64_t* pix_idx, ncn_stream_t stream);
/* the same with the triangle expansion of datasets/base.py:130-141 (`triang_max_expand`): x1 moves max_expand rows down, x2
 * max_expand rows up, x3 max_expand pixels left, each only when it stays inside the image / its row (triangle strategies) */
int ncn_sample_ray_batch_ex(int strategy, int64_t* seed_dev, int n_rays, int n_poses, int height, int width, int patch_size,
                            int max_expand, int64_t* img_idx, int64_t* pix_idx, ncn_stream_t stream);
/* random_tr_poses (datasets/base.py:106-126, 148-159 `rnd_img_idxs`; train_nerf.py:169-172): fills rows [n_gt_rays, 2 n_gt_rays) of
 * a batch whose first n_gt_rays rows ncn_sample_ray_batch[_ex] has drawn: pix_idx repeats the first half ("same pixels for random
 * camera poses"), img_idx = pose_offset + a draw in [0, n_random_poses) per patch / triangle (strategies 0, 2) or one draw for the
 * whole batch (1, 3).  pose_offset = where the generated poses start in the pose table handed to ncn_rays_from_pixels (the
 * reference concatenates them behind the training poses' rows, train_nerf.py:170).  Reads seed_dev (does not advance it). */
int ncn_sample_random_pose_half(int strategy, const int64_t* seed_dev, int n_gt_rays, int n_random_poses, int64_t pose_offset,
                                int patch_size, int64_t* img_idx, int64_t* pix_idx, ncn_stream_t stream);
int ncn_gather_pixels(const void* table, const int64_t* img_idx, const int64_t* pix_idx, int64_t n, int64_t pixels_per_image,
                      int words_per_pixel, void* out, ncn_stream_t stream);

/* Normals of rendered depth IMAGES for the evaluation loop (datasets/hypersim_src/utils.py:544-611,
 * _extract_normals_from_depth_batch): depth (B,H,W) f32, ray_dirs_cc (H*W,3) f32 camera-frame directions, poses
 * (B, pose_rows = 3 or 4, 4) f32 row-major camera-to-world -> normals (B,H,W,3) f32 in the world frame;
 * (0,0,0) on the one-pixel border and where the pixel's own depth is 0 / NaN / Inf. */
int ncn_normals_from_depth_image(const float* depth, const float* ray_dirs_cc, const float* poses, int pose_rows,
                                 int n_images, int height, int width, float* normals, ncn_stream_t stream);

typedef struct ncn_kmeans_params {
  int32_t k;                      /* 20  (losses.py:436) */
  int32_t niter;                  /* 20  (losses.py:437) */
  int32_t seed;                   /* 1234 (faiss default) */
  int32_t max_points_per_centroid;/* 256 (faiss default) */
  int32_t spherical;              /* 1 */
} ncn_kmeans_params;
size_t ncn_kmeans_workspace_bytes(int64_t n_points_max, int k);
/* Spherical k-means on the valid rows of x (M,3) (rows that are all-zero / NaN / Inf
 * are skipped, exactly the filter of losses.py:427-430).  Single-CTA, no host sync.
 * out: centroids (k,3) f32, assign (M) i32 (-1 for skipped rows), n_valid (1) i32. */
int ncn_kmeans_spherical(const float* x, int64_t n_points, const ncn_kmeans_params* p,
                         float* centroids, int32_t* assign, int32_t* n_valid,
                         void* workspace, size_t workspace_bytes, ncn_stream_t stream);
/* the per-triangle half of ncn_cluster_tail on its own (after ncn_cluster_chain): dL/dnormals of w[0] L_ort + w[1] L_dot + w[2] L_L1
 * (weights read from device memory) and, through the normals, dL/ddepth (atomically accumulated) */
int ncn_cluster_bw_depth(const float* normals, const int32_t* labels, int64_t n_points, const float* stats,
                         const float* weights_dev, float* dL_dnormals, const float* origin, const float* dir,
                         const float* depth, const int64_t* idx1, const int64_t* idx2, const int64_t* idx3,
                         float* dL_ddepth, ncn_stream_t stream);
/* ncn_normals_from_depth_fw -> ncn_kmeans_spherical -> ncn_cluster_select -> ncn_cluster_loss_fw as ONE thread-block-cluster
 * launch (3 <= k <= 32, else NCN_E_UNSUPPORTED): the normals are computed by the cluster's CTAs as a prologue, and the selection
 * (losses.py:97-166), the cluster statistics and the three loss terms (losses.py:441-478) run as an epilogue while the points and
 * centroids are still resident in the cluster's shared memory (member counts and fixed-point sums folded over distributed shared
 * memory).  Same outputs as the four calls: normals (M,3), centroids (k,3), assign (M), n_valid (1), labels (M), sel (3),
 * losses (3), stats (32); integer outputs identical, the float sums of the loss terms differ by fp32 summation order only. */
int ncn_cluster_chain(const float* origin, const float* dir, const float* depth, const int64_t* idx1, const int64_t* idx2,
                      const int64_t* idx3, int64_t n_tri, const ncn_kmeans_params* p, float t_similar, float* normals,
                      float* centroids, int32_t* assign, int32_t* n_valid, int32_t* labels, int32_t* sel, float* losses,
                      float* stats, void* workspace, size_t workspace_bytes, ncn_stream_t stream);
/* Orthogonal-triple selection + merge + opposite labelling (losses.py:97-166).
 * labels (M) i32 in {-3..3} (0 = unused / skipped), sel (3) i32 = (c1,c2,c3). */
int ncn_cluster_select(const float* centroids, const int32_t* assign, int64_t n_points, int k,
                       float t_similar, int32_t* labels, int32_t* sel, ncn_stream_t stream);
/* Loss terms of losses.py:441-478 and their gradient.
 * out losses (3) f32 = [ort_dot, centr_dot, centr_L1] (unweighted), NaN if a cluster
 * is empty (caller applies the reference's validity filter);
 * stats (32) f32: per cluster k at [8k..8k+7] = count, c_k (3), |mean|, sum sign(n-c_k) (3);
 * [24..26] = the three losses, [27] = 1 if all three clusters are non-empty. */
int ncn_cluster_loss_fw(const float* normals, const int32_t* labels, int64_t n_points,
                        float* losses, float* stats, ncn_stream_t stream);
/* weights (3) f32 = dL/d[ort_dot, centr_dot, centr_L1] -> dL_dnormals (M,3) fully written */
int ncn_cluster_loss_bw(const float* normals, const int32_t* labels, int64_t n_points,
                        const float* stats, const float* weights_dev, float* dL_dnormals,
                        ncn_stream_t stream);
/* ncn_cluster_select -> ncn_cluster_loss_fw -> ncn_cluster_loss_bw -> ncn_normals_from_depth_bw in TWO launches (the two
 * single-CTA stages share one, the two per-triangle stages the other); same arguments and results as the four calls,
 * dL_ddepth accumulated into a caller-zeroed buffer: the step between the k-means result and the depth gradient of
 * losses.py:441-509. */
int ncn_cluster_tail(const float* centroids, const int32_t* assign, int64_t n_points, int k, float t_similar,
                     int32_t* labels, int32_t* sel, const float* normals, float* losses, float* stats,
                     const float* weights_dev, float* dL_dnormals, const float* origin, const float* dir,
                     const float* depth, const int64_t* idx1, const int64_t* idx2, const int64_t* idx3,
                     float* dL_ddepth, ncn_stream_t stream);

/* Photometric terms fused with the background composite (rendering.py:231-241,
 * losses.py:347-361): rgb = rend[:, :3] + bg*(1-opacity); sums[0] += sum((rgb-target)^2),
 * sums[1] += sum(-(o+1e-10)log(o+1e-10)) (caller zeroes sums; means = /3R and /R).
 * Writes (not accumulates) dL_drend (R,C) and dL_dopacity (R) of
 *   grad_scale * ( mean((rgb-target)^2) + opacity_w * mean(entropy) );  either may be NULL.
 * bg_rgb_host: 3 floats on the HOST.  rgb_out (R,3) may be NULL. */
int ncn_photometric_loss(const float* rend, const float* opacity, const float* target_rgb,
                         int64_t n_rays, int n_channels, const float* bg_rgb_host,
                         float opacity_w, float grad_scale, float* rgb_out, float* sums,
                         float* dL_drend, float* dL_dopacity, ncn_stream_t stream);

/* ncn_photometric_loss with a target colour on the first n_gt_rays rays only (see ncn_composite_train_fw_photometric_gt) */
int ncn_photometric_loss_gt(const float* rend, const float* opacity, const float* target_rgb,
                            int64_t n_rays, int64_t n_gt_rays, int n_channels, const float* bg_rgb_host,
                            float opacity_w, float grad_scale, float* rgb_out, float* sums,
                            float* dL_drend, float* dL_dopacity, ncn_stream_t stream);

/* Semantic cross-entropy on the rendered logits (losses.py:226-242, 569-573: nn.CrossEntropyLoss(ignore_index=-1)
 * applied to (sem_pred, target - 1), mean over the non-void rays): logits = rend[:, c_off : c_off + n_cls] (row stride
 * c_total), labels (R) i64 in [0, n_cls] with 0 = void (ignored).  sums[0] += sum over valid rays of -log softmax[label-1],
 * sums[1] = number of valid rays (caller zeroes sums[0]; loss = sums[0] / sums[1], NaN when no ray is valid - the
 * reference then drops the term).  Writes (not accumulates) dL_drend[:, c_off : c_off + n_cls] =
 * grad_scale * (softmax - onehot) / n_valid for valid rays, 0 otherwise; other columns are left untouched.  n_cls <= 64. */
int ncn_semantic_ce_loss(const float* rend, int c_total, int c_off, int n_cls, const int64_t* labels,
                         int64_t n_rays, float grad_scale, float* sums, float* dL_drend,
                         ncn_stream_t stream);

/* ------------------------------------------------------------------------- */
/* (9) optimizer + data-parallel all-reduce  (the step either side of the path)*/
/*     replaces apex FusedAdam (train_nerf.py:262-285) and torch DDP          */
/*     (train_nerf.py:949-952)                                                */
/* ------------------------------------------------------------------------- */
/* Adam (apex FusedAdam adam_w_mode=True semantics: decoupled weight decay), fp32
 * master params; also refreshes the fp16 copy used by the kernels and zeroes the
 * gradient, in one pass.  grad is divided by *grad_div_dev if non-NULL
 * (loss-scale * world-size) and the step is skipped when *skip_dev != 0.
 * lr_bc_dev: NULL, or 3 DEVICE floats (lr, 1-beta1^t, 1-beta2^t) that override lr / step, so a
 * captured CUDA graph can be replayed while the host advances the schedule. */
int ncn_adam_step(float* param, float* grad, float* m, float* v, void* param_f16,
                  int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int step, const float* grad_div_dev, const int32_t* skip_dev,
                  const float* clip_coef_dev, const float* lr_bc_dev, ncn_stream_t stream);
/* The same update for SEVERAL parameter groups of one flat buffer in ONE launch (the reference's two FusedAdam groups,
 * train_nerf.py:264-274: hash table wd 0, MLPs wd 1e-6).  Group q covers [start[q], start[q+1]) (start[0] = 0, every
 * start a multiple of 4).  max_norm > 0 folds clip_grad_norm_(max_norm) in: sumsq_dev then holds the squared norm of the
 * unscaled gradient (ncn_grad_sumsq) and the coefficient min(1, max_norm/(sqrt(sumsq)+1e-6)) is derived in the kernel.
 * lr and the bias corrections always come from lr_bc_dev (3 device floats). */
#define NCN_ADAM_MAX_GROUPS 4
typedef struct ncn_adam_groups {
  int32_t n_groups;
  int32_t reserved;
  int64_t start[NCN_ADAM_MAX_GROUPS];
  float weight_decay[NCN_ADAM_MAX_GROUPS];
  float max_norm;
  float reserved2;
} ncn_adam_groups;
int ncn_adam_step_groups(float* param, float* grad, float* m, float* v, void* param_f16, int64_t n,
                         const ncn_adam_groups* groups, float beta1, float beta2, float eps,
                         const float* grad_div_dev, const int32_t* skip_dev, const float* sumsq_dev,
                         const float* lr_bc_dev, ncn_stream_t stream);
/* sum of squares of grad/(div) into out[0] (ACCUMULATED), and non-finite flag into flag[0].  Deterministic (fixed
 * summation order: bit-identical on every rank of a data-parallel job); not re-entrant across streams of one device. */
int ncn_grad_sumsq(const float* grad, int64_t n, const float* grad_div_dev,
                   float* out, int32_t* flag, ncn_stream_t stream);
/* coef[0] = min(1, max_norm / (sqrt(sumsq[0]) + 1e-6))  (torch clip_grad_norm_), on device */
int ncn_clip_coef(const float* sumsq_dev, float max_norm, float* coef_dev, ncn_stream_t stream);

typedef struct ncn_comm ncn_comm;
/* NCCL unique id plumbing: rank 0 calls ncn_comm_unique_id (128 bytes), shares the
 * bytes out of band (torch.distributed broadcast), every rank calls ncn_comm_init. */
int ncn_comm_unique_id(void* id128_host);
int ncn_comm_init(ncn_comm** comm, const void* id128_host, int world_size, int rank);
int ncn_comm_allreduce_sum_f32(ncn_comm* comm, float* buf, int64_t n, ncn_stream_t stream);
int ncn_comm_destroy(ncn_comm* comm);
const char* ncn_comm_last_error(void);

/* The same exchange step WITHOUT a collective library: gradient reduction, ||g||^2, clip, Adam and the refresh of the fp16
 * working parameters as two kernels over NVLink peer memory (one process per GPU, buffers shared by CUDA IPC).
 * Rank r owns the r-th 1/W slice of the flat parameter vector: it sums that slice out of every rank's gradient buffer
 * (P2P loads, fixed rank order), applies Adam to it (its p / m / v slices are the only ones kept current on this rank) and
 * stores the fp16 result into EVERY rank's fp16 parameter buffer - the buffer the forward kernels read.  Equivalent to
 * ncn_comm_allreduce_sum_f32 + ncn_grad_sumsq + ncn_adam_step_groups on every rank (train_nerf.py:949-955) up to the
 * summation order of the W gradient terms; the fp32 master outside a rank's own slice is stale by design.
 *   create : cudaMalloc's the gradient buffer (n f32, zeroed), the fp16 parameter buffer (n f16) and a sync block;
 *            n_params % 4 == 0, world <= 8.  The backward must accumulate into ncn_peer_grad(), the forward must read
 *            ncn_peer_p16().
 *   handles: 3 cudaIpcMemHandle_t (192 bytes) to ship to the other ranks out of band; connect: all ranks' handles
 *            (world x 192 bytes, rank order).  world == 1 needs neither.
 *   step   : asynchronous on `stream`, graph-capturable; groups / lr_bc_dev / skip_dev / grad_div_dev as in
 *            ncn_adam_step_groups; sumsq_out_dev (optional) receives the squared norm of the averaged gradient.
 *            Every rank must call it the same number of times.  Cross-GPU waits are bounded (20 s by default,
 *            ncn_peer_set_timeout) and FATAL for the exchange instead of hanging: the error word is latched, the step that
 *            timed out and every later one are skipped on this rank (no Adam, nothing published, gradient zeroed), and a
 *            time-out while waiting for the peers' gradients posts NaN as this rank's partial norm so the live peers skip the
 *            step too.  ncn_peer_poll reads the word from mapped host memory WITHOUT synchronising (0 = ok, 1 + phase);
 *            ncn_peer_error is the synchronising read. */
typedef struct ncn_peer ncn_peer;
int ncn_peer_create(ncn_peer** out, int rank, int world, int64_t n_params);
float* ncn_peer_grad(ncn_peer* p);
void* ncn_peer_p16(ncn_peer* p);
int ncn_peer_handles(ncn_peer* p, void* handles192_host);
int ncn_peer_connect(ncn_peer* p, const void* all_handles_host);
void ncn_peer_shard(int64_t n_params, int rank, int world, int64_t* lo, int64_t* hi);
int ncn_peer_step(ncn_peer* p, float* param, float* m, float* v, const ncn_adam_groups* groups, float beta1,
                  float beta2, float eps, const float* grad_div_dev, const int32_t* skip_dev,
                  const float* lr_bc_dev, float* sumsq_out_dev, ncn_stream_t stream);
/* on = 1: ncn_peer_step no longer zeroes the gradient buffer outside its own slice; the caller zeroes the WHOLE buffer after the step
 * (any stream, any time before the next backward) - takes the 4 B/param memset off the exchange's critical path */
int ncn_peer_set_external_zero(ncn_peer* p, int on);
/* Overlap of the exchange with the tail of the backward pass.  set_cut(cut): [cut, n) of the flat vector is the EARLY range - the
 * part of the gradient the backward completes first (the caller orders its launches accordingly: here the fine hash-grid levels
 * and the MLPs, then the coarse levels); every rank then owns the r-th 1/W slice of [0, cut) AND of [cut, n)
 * (ncn_peer_segments: {lo, hi, lo_early, hi_early} of rank q).  ncn_peer_early, launched on any stream once the early range of
 * THIS rank's gradient is complete, reduces this rank's early slice out of every peer's buffer behind its own flag phase while
 * the rest of the backward is still running everywhere; ncn_peer_step then only pulls the late slice.  With a cut set, every
 * ncn_peer_step must be preceded by exactly one ncn_peer_early since the previous step (or by none at all while the gradient
 * is still zero), on every rank.  cut == n (default): no early range, ncn_peer_early is refused. */
int ncn_peer_set_cut(ncn_peer* p, int64_t cut);
int ncn_peer_segments(ncn_peer* p, int rank_q, int64_t* seg4_out);
void ncn_peer_segments_of(int64_t n_params, int64_t cut, int rank, int world, int64_t* seg4_out);   /* host arithmetic only */
int ncn_peer_early(ncn_peer* p, const float* grad_div_dev, ncn_stream_t stream);
int ncn_peer_debug_times_early(ncn_peer* p, unsigned long long* times4_host);
int ncn_peer_error(ncn_peer* p, unsigned int* error_host);
unsigned int ncn_peer_poll(ncn_peer* p);
/* developer A/B knobs of the reduce kernel: 16-byte loads in flight per peer and thread (1 [default], 2, 4; returns the old value);
 * peers loaded per batch (2, 4 [default], 8) and CTAs per SM (1, 2 [default]) */
int ncn_peer_set_loads(int loads_per_peer);
int ncn_peer_set_shape(int peers_per_batch, int ctas_per_sm);
int ncn_peer_set_early_loads(int loads_per_peer);      /* the same knob for ncn_peer_early; 0 (default) = by world size (4 / 2 / 1 for <= 2 / <= 4 / 8 ranks) */
/* developer timeline of the last step (synchronises): 8 x ns = [K1 start, wait-0 over, K1 reduced, K2 start, wait-1 over, Adam+publish done, wait-2 over, -] */
int ncn_peer_debug_times(ncn_peer* p, unsigned long long* times8_host);
int ncn_peer_set_timeout(double seconds);
int ncn_peer_destroy(ncn_peer* p);

#ifdef __cplusplus
}
#endif
#endif /* NCN_H_ */
