/*
 * sgnerf_b200.h -- C ABI of libsgnerf_b200.so: SG-NeRF's per-ray render hot path on B200 (sm_100a).
 *
 * The reference has no native library; its boundary for this path is (paths relative to the reference):
 *   - PyCUDA kernel handles returned by lighting_fast_querier.build_cuda()
 *     (models/neural_points/query_point_indices_worldcoords.py:134-697) and called with raw
 *     torch data_ptr()s through `Holder` (:21-28, :719-938);
 *   - the torch modules NeuralPoints.forward (models/neural_points/neural_points.py:942-988),
 *     PointAggregator.forward (models/aggregators/point_aggregators.py:868-959) and
 *     ray_march / alpha_ray_march (models/rendering/diff_ray_marching.py:509-573).
 * Each entry point below names the reference interface it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.
 *   - every pointer is a DEVICE pointer owned by the caller unless marked [host]; the library never
 *     allocates or frees user-visible memory.  Scratch comes from a caller-supplied workspace whose size
 *     is queried first (sgn_*_workspace_bytes).  Workspaces must be 256-byte aligned.
 *   - every launch entry takes a cudaStream_t (as void*) and is asynchronous; nothing synchronises.
 *   - return value: 0 on success, negative SGN_E_* on error; sgn_last_error() gives a thread-local text.
 *   - batch size B is 1 throughout (the reference only ever runs B = 1).
 *   - there is no CPU fallback: without a CUDA device every launch entry returns SGN_E_CUDA.
 */
#ifndef SGNERF_B200_H
#define SGNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGN_OK 0
#define SGN_E_INVALID (-1)   /* bad argument / unsupported configuration */
#define SGN_E_CUDA (-2)      /* CUDA runtime error (text in sgn_last_error) */
#define SGN_E_WORKSPACE (-3) /* workspace too small or misaligned */

#define SGN_MAX_K 32

const char* sgn_last_error(void);
int sgn_version(void);
/* Number of kernel launches this library has issued so far in the process (host counter, not thread-safe). */
uint64_t sgn_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Occupancy grid ("occ vox").  Replaces claim_occ / map_coor2occ / fill_occ2pnts and build_occ_vox
 * (query_point_indices_worldcoords.py:265-410, :706-778).  The reference rebuilds it on every
 * query_points() call; here it is built once per point-cloud version and queried many times.
 * Slot numbering and per-voxel point lists follow the reference's sequential thread order
 * (first visitor by point index owns the slot; lists are in point-index order; the `voxel_idx > 0`
 * guard of :395 leaves slot 0 empty; overflow beyond max_o / P uses the same cuRAND XORWOW reservoir
 * with seeds `index + 2*seconds`).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float origin[3];      /* d_coord_shift = ranges[:3]                                   (:793) */
    float vsize[3];       /* scaled voxel size = vsize * vscale                           (:73)  */
    int32_t dim[3];       /* scaled_vdim                                                  (:86)  */
    int32_t query_size[3];/* occupancy dilation box (map_coor2occ's kernel_size argument) (:797) */
    int32_t P;            /* max points per voxel                                                */
    int32_t max_o;        /* max occupied voxels                                                 */
    uint64_t seconds_claim; /* time.time() at :715 (reservoir seed when > max_o voxels)          */
    uint64_t seconds_fill;  /* time.time() at :751 (reservoir seed when > P points in a voxel)   */
} SgnGridCfg;

typedef struct SgnGrid SgnGrid; /* opaque, host-side; holds device pointers into `persistent` */

/* persistent: lives as long as the grid; scratch: only during sgn_grid_build. */
int sgn_grid_workspace_bytes(int64_t N, const SgnGridCfg* cfg, size_t* persistent_bytes, size_t* scratch_bytes);
int sgn_grid_build(const float* xyz /*[N,3]*/, int64_t N, int64_t actual_n, const SgnGridCfg* cfg,
                   void* persistent, size_t persistent_bytes, void* scratch, size_t scratch_bytes,
                   SgnGrid** out, void* stream);
/* The same with options.  SGN_GRID_NO_NEIGHBOUR_LISTS: skip the per-voxel neighbour lists of the K-NN kernel (about two thirds of the
 * build time; the kernel then walks the compact brick index itself, a little slower per frame) -- for clouds that are edited between
 * frames, where the grid is rebuilt more often than it is queried. */
#define SGN_GRID_NO_NEIGHBOUR_LISTS 1
int sgn_grid_build_flags(const float* xyz /*[N,3]*/, int64_t N, int64_t actual_n, const SgnGridCfg* cfg, void* persistent,
                         size_t persistent_bytes, void* scratch, size_t scratch_bytes, int flags, SgnGrid** out, void* stream);
int sgn_grid_destroy(SgnGrid* g);
/* Device pointers of the built structures, for tests and tooling (all int32 unless noted):
 *   0 cell_slot [X*Y*Z] (= coor_2_occ), 1 occ_bits uint32[ceil(X*Y*Z/32)] (= coor_occ as a bitmask),
 *   2 slot_coor [max_o*3] (= occ_2_coor), 3 slot_count [max_o] (= occ_numpnts, uncapped),
 *   4 slot_start [max_o+1] (offset of the slot's list in cand), 5 cand float4[(x,y,z,bits(pidx))],
 *   6 counters [4]: {n_claimed (= occ_idx), n_candidates, 0, 0}, 7 brick mask of the march (1 bit per 8^3 voxels),
 *   8 knn_brick uint32[4 per 4^3-voxel brick] = (mask lo, mask hi, rank of the first listed voxel, 0),
 *   9 knn_list int32[2 per listed voxel] = (first candidate, count).                               */
int sgn_grid_buffer(const SgnGrid* g, int which, void** ptr, int64_t* n_elements);

/* ------------------------------------------------------------------------------------------------
 * Ray march + K-NN.  Replaces near-far ray positions (diff_ray_marching.py:387), mask_raypos,
 * the cumsum glue, get_shadingloc[_with_semantic] and query_neigh_along_ray_layered[_semantic_guidance]
 * (query_point_indices_worldcoords.py:413-681, :811-938) for all R rays in one pass, uncompacted:
 * row r of the outputs belongs to input ray r.  The reference's two ray compactions (:838, :949)
 * are pure row selections by `ray_mask` and are left to the caller.
 *   t            middle_point_ts: [D] shared by all rays (t_per_ray = 0) or [R,D] (t_per_ray = 1)
 *   ray_label    [R] or NULL.  Non-NULL selects the semantic-guidance kernel and needs pt_label [N],
 *                pt_label_prob_bits [N,20] (the int32 tensor the reference passes as float*, :916).
 * Outputs
 *   sample_pidx  int32 [R,SR,K]  (-1 = empty), slot order identical to the sequential reference
 *   sample_loc_w f32   [R,SR,3]  (unused slots are (0,0,0), :835)
 *   sample_mask  int32 [R,SR]    (= sample_loc_mask)
 *   sample_label int32 [R,SR]    scratch, only touched (and required) with semantic guidance
 *   ray_mask     int8  [R]       1 iff the ray has at least one neighbour (final mask of :948)
 * ---------------------------------------------------------------------------------------------- */
int sgn_query(const SgnGrid* g, const float* campos /*[3]*/, const float* raydir /*[R,3]*/, const float* t,
              int t_per_ray, int64_t R, int D, int SR, int K, int kernel_size0, float radius2,
              const int32_t* ray_label, const int32_t* pt_label, const int32_t* pt_label_prob_bits,
              uint64_t seconds_query, int32_t* sample_pidx, float* sample_loc_w, int32_t* sample_mask,
              int32_t* sample_label, int8_t* ray_mask, void* stream);

/* The frame variant: identical, except that the sample_pidx rows of slots with sample_mask == 0 are left UNWRITTEN (sgn_query fills them
 * with -1; at SR 200 they are nine tenths of a 4 GB array).  For consumers that take the mask: sgn_agg_forward_frame_masked. */
int sgn_query_frame(const SgnGrid* g, const float* campos /*[3]*/, const float* raydir /*[R,3]*/, const float* t,
              int t_per_ray, int64_t R, int D, int SR, int K, int kernel_size0, float radius2,
              const int32_t* ray_label, const int32_t* pt_label, const int32_t* pt_label_prob_bits,
              uint64_t seconds_query, int32_t* sample_pidx, float* sample_loc_w, int32_t* sample_mask,
              int32_t* sample_label, int8_t* ray_mask, void* stream);

/* Tuning / test aid: sgn_query has two march kernels with identical results -- a thread-per-ray brick walk (frames) and a
 * warp-per-ray kernel (small ray counts, e.g. a training patch); mode 0 picks by ray count, 1 / 2 force one of them. */
int sgn_query_march_mode(int mode);

/* Strict-compat materialisation of NeuralPoints.forward's gathered tensors
 * (neural_points.py:956-972): out[j, :] = table[max(pidx[j],0), :] for n_rows index entries. */
int sgn_gather_rows(const float* table /*[N,C]*/, int C, const int32_t* pidx, int64_t n_rows, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Aggregation.  Replaces NeuralPoints.forward's gathers (neural_points.py:956-988, w2pers :838-850),
 * PointAggregator.forward / linear / viewmlp (point_aggregators.py:868-959, :494-502, :561-786),
 * positional_encoding (helpers/networks.py:175-192) for the canonical branch (agg_dist_pers=20,
 * agg_distance_kernel=linear, agg_intrp_order=2, apply_pnt_mask=1, agg_weight_norm=1, Rw2c=identity).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t feat_dim;        /* point_features_dim (32)                       */
    int32_t num_feat_freqs;  /* 3 */
    int32_t dist_xyz_freq;   /* 5 */
    int32_t num_viewdir_freqs; /* 4 */
    int32_t width;           /* shading_feature_num (256)                     */
    int32_t n_block1;        /* shading_feature_mlp_layer1 (2)                */
    int32_t n_block2_bpnet;  /* shading_feature_mlp_layer2_bpnet (0/1)        */
    int32_t label_dim;       /* 96 when the label embedding is concatenated, else 0 */
    int32_t n_block3;        /* shading_feature_mlp_layer3 (2)                */
    int32_t n_color;         /* shading_color_mlp_layer (4)                   */
    int32_t act_super;       /* 1: softplus(x-1) and sigmoid*1.002-0.001      */
    float leaky_slope;       /* 0.01                                          */
} SgnAggCfg;

typedef struct {
    const float* xyz;        /* [N,3]                                  neural_points.xyz            */
    const float* embedding;  /* [N,feat_dim]                           neural_points.points_embeding */
    const float* color;      /* [N,3]                                  neural_points.points_color    */
    const float* dir;        /* [N,3]                                  neural_points.points_dir      */
    const float* conf;       /* [N]                                    neural_points.points_conf     */
    const float* label_emb;  /* [N,label_dim] or NULL                  bpnet_points_embedding        */
    int64_t N;
} SgnPointTables;

typedef struct {             /* gradient accumulators (+=); any may be NULL */
    float* embedding; float* color; float* dir; float* conf;
} SgnPointGrads;

/* Number of Linear layers for a config and their (in,out) sizes in state_dict order:
 * block1.*, block2_bpnet.*, block3.*, alpha_branch.0, color_branch.* ; weights are [out,in] row-major
 * exactly as torch.nn.Linear stores them.  `weights`/`biases` arguments below are [host] arrays of
 * device pointers in this order. */
int sgn_agg_num_layers(const SgnAggCfg* cfg);
int sgn_agg_layer_shape(const SgnAggCfg* cfg, int layer, int* in_features, int* out_features);

#define SGN_PRECISION_FP32 0   /* fp32 SIMT, strict parity mode                                   */
#define SGN_PRECISION_BF16 1   /* bf16 tcgen05/TMEM tensor-core tiles, fp32 accumulate (forward only) */
#define SGN_PRECISION_TF32 2   /* layer-wise path of FP32 with its GEMMs on tcgen05 kind::tf32 (fp32 storage and accumulation;
                                  the arithmetic of the reference's cuBLAS default, torch 1.10 allow_tf32) -- forward + backward */

/* save_for_backward != 0 keeps the per-layer activations in the workspace for sgn_agg_backward.
 * N = number of points in the tables (the bf16 path keeps one 448-byte operand row per point in the workspace). */
int sgn_agg_workspace_bytes(const SgnAggCfg* cfg, int64_t N, int64_t R, int SR, int K, int precision,
                            int save_for_backward, size_t* bytes);

/*   pidx [R,SR,K], loc_w [R,SR,3] from sgn_query; raydir [R,3]; campos [3]; camrotc2w [3,3] row-major.
 *   decoded [R,SR,4] (sigma,r,g,b; zero where !ray_valid), ray_valid uint8 [R,SR],
 *   loc_pers [R,SR,3] (= w2pers(loc_w)) or NULL, weight [R,SR,K] or NULL, conf_coef [R,SR,K] or NULL. */
int sgn_agg_forward(const SgnAggCfg* cfg, const float* const* weights /*[host]*/, const float* const* biases /*[host]*/,
                    const SgnPointTables* tables, const int32_t* pidx, const float* loc_w, const float* raydir,
                    const float* campos, const float* camrotc2w, int64_t R, int SR, int K, int precision,
                    int save_for_backward, float* decoded, uint8_t* ray_valid, float* loc_pers, float* weight,
                    float* conf_coef, void* workspace, size_t workspace_bytes, void* stream);

/* Inference cache of the bf16 path.  The first per-neighbour layer (and block2_bpnet.0) is linear in its input and most of that input
 * depends on the point only, so its point part is one table per point (bf16 [N,256]) -- a function of the embedding tables and the layer
 * weights alone.  Like the occupancy grid it can be built once per (point cloud, weights) version and passed to every frame's
 * sgn_agg_forward_cached; with point_cache = NULL (what sgn_agg_forward does) the tables are rebuilt inside the call's workspace.
 * The caller must rebuild the cache whenever embeddings or aggregator weights change. */
int sgn_agg_point_cache_bytes(const SgnAggCfg* cfg, int64_t N, size_t* bytes);
int sgn_agg_point_cache_build(const SgnAggCfg* cfg, const float* const* weights /*[host]*/, const SgnPointTables* tables, void* cache,
                              size_t cache_bytes, void* stream);
/* Incremental update after point edits: the cache rows of the n_rows points listed in `rows` (device int32; entries outside [0, N) are
 * skipped) are recomputed from the current embedding tables with the weights the cache was built with (they are kept, packed, inside
 * the cache).  The tables must have the same N as at build time. */
int sgn_agg_point_cache_update(const SgnAggCfg* cfg, const SgnPointTables* tables, void* cache, size_t cache_bytes, const int32_t* rows,
                               int64_t n_rows, void* stream);
int sgn_agg_forward_cached(const SgnAggCfg* cfg, const float* const* weights /*[host]*/, const float* const* biases /*[host]*/,
                           const SgnPointTables* tables, const int32_t* pidx, const float* loc_w, const float* raydir,
                           const float* campos, const float* camrotc2w, int64_t R, int SR, int K, int precision,
                           int save_for_backward, float* decoded, uint8_t* ray_valid, float* loc_pers, float* weight,
                           float* conf_coef, void* workspace, size_t workspace_bytes, const void* point_cache, void* stream);

/* The frame variant: sgn_agg_forward_cached plus loc_depth [R,SR] (or NULL) = the camera depth of every sample (loc_pers[..., 2]) as a
 * dense array -- the only part of loc_pers the frame tail (sgn_render_composite_depth) reads; loc_pers itself may then be NULL. */
int sgn_agg_forward_frame(const SgnAggCfg* cfg, const float* const* weights /*[host]*/, const float* const* biases /*[host]*/,
                          const SgnPointTables* tables, const int32_t* pidx, const float* loc_w, const float* raydir,
                          const float* campos, const float* camrotc2w, int64_t R, int SR, int K, int precision,
                          int save_for_backward, float* decoded, uint8_t* ray_valid, float* loc_pers, float* loc_depth, float* weight,
                          float* conf_coef, void* workspace, size_t workspace_bytes, const void* point_cache, void* stream);

/* The same with sgn_query's sample_mask [R,SR] (int32; 0 = the slot holds no sample): the all -1 sample_pidx rows of such slots -- two
 * thirds of a 640x480 / SR 24 frame, nine tenths at SR 200 -- are not read.  Identical results. */
int sgn_agg_forward_frame_masked(const SgnAggCfg* cfg, const float* const* weights /*[host]*/, const float* const* biases /*[host]*/,
                                 const SgnPointTables* tables, const int32_t* pidx, const int32_t* sample_mask, const float* loc_w,
                                 const float* raydir, const float* campos, const float* camrotc2w, int64_t R, int SR, int K, int precision,
                                 int save_for_backward, float* decoded, uint8_t* ray_valid, float* loc_pers, float* loc_depth, float* weight,
                                 float* conf_coef, void* workspace, size_t workspace_bytes, const void* point_cache, void* stream);

/* Backward of sgn_agg_forward (autograd of the reference path).  Needs the workspace
 * of a forward call made with save_for_backward = 1 and the same arguments.  d_weights/d_biases are
 * [host] arrays of device pointers (accumulated, +=); d_conf_coef may be NULL. */
int sgn_agg_backward(const SgnAggCfg* cfg, const float* const* weights, const float* const* biases,
                     const SgnPointTables* tables, const int32_t* pidx, const float* loc_w, const float* raydir,
                     const float* campos, const float* camrotc2w, int64_t R, int SR, int K,
                     const float* d_decoded /*[R,SR,4]*/, const float* d_conf_coef /*[R,SR,K]*/,
                     float* const* d_weights, float* const* d_biases, const SgnPointGrads* d_tables,
                     void* workspace, size_t workspace_bytes, void* stream);
/* The same with the arithmetic of the GEMMs chosen: SGN_PRECISION_FP32 (what sgn_agg_backward does) or SGN_PRECISION_TF32. */
int sgn_agg_backward_prec(const SgnAggCfg* cfg, const float* const* weights, const float* const* biases,
                          const SgnPointTables* tables, const int32_t* pidx, const float* loc_w, const float* raydir,
                          const float* campos, const float* camrotc2w, int64_t R, int SR, int K, int precision,
                          const float* d_decoded /*[R,SR,4]*/, const float* d_conf_coef /*[R,SR,K]*/,
                          float* const* d_weights, float* const* d_biases, const SgnPointGrads* d_tables,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Measurement aid (bench.py's roofline): when enabled, the bf16 path brackets its dominant kernel, agg_tuple_tc_kernel, with
 * CUDA events on the launching stream (this serialises the host with the previous call's kernel; leave it off otherwise).
 * sgn_agg_kernel_timing_read returns the summed duration and the number of timed launches since it was enabled. */
int sgn_agg_kernel_timing(int enable);
int sgn_agg_kernel_timing_read(float* total_ms, int* launches);

/* ------------------------------------------------------------------------------------------------
 * Compositing.  Replaces ray_march / alpha_ray_march (diff_ray_marching.py:509-573) with
 * render_func = radiance and blend_func = alpha (blend = 0) or alpha2 (blend = 1)
 * (diff_render_func.py:36-49), and the step-size glue of
 * models/neural_points_volumetric_model.py:569-577 (sgn_ray_dist).
 * ---------------------------------------------------------------------------------------------- */
int sgn_ray_dist(const float* loc_pers /*[R,SR,3]*/, const uint8_t* ray_valid /*[R,SR]*/, float vsize_z,
                 int raydist_mode_unit, int64_t R, int SR, float* ray_dist /*[R,SR]*/, void* stream);

/* decoded [R,SR,4]; ray_dist [R,SR]; valid uint8 [R,SR]; bg [3] or NULL.
 * Outputs (any may be NULL): ray_color [R,3], opacity [R,SR], acc_transmission [R,SR], blend_weight [R,SR],
 * bg_transmission [R]. */
int sgn_composite_forward(const float* decoded, const float* ray_dist, const uint8_t* valid, const float* bg,
                          int blend, int64_t R, int SR, float* ray_color, float* opacity, float* acc_transmission,
                          float* blend_weight, float* bg_transmission, void* stream);

/* Gradients w.r.t. decoded given cotangents of ray_color [R,3], opacity [R,SR], blend_weight [R,SR],
 * bg_transmission [R] (any may be NULL).  d_decoded [R,SR,4] is overwritten. */
int sgn_composite_backward(const float* decoded, const float* ray_dist, const uint8_t* valid, const float* bg,
                           int blend, int64_t R, int SR, const float* d_ray_color, const float* d_opacity,
                           const float* d_blend_weight, const float* d_bg_transmission, float* d_decoded,
                           void* stream);

/* fill_invalid (models/neural_points_volumetric_model.py:158-195) for uncompacted rows:
 * rows with ray_mask == 0 get bg colour / opacity 0 / is_background 1. */
int sgn_fill_invalid(const int8_t* ray_mask, const float* bg /*[3]*/, int64_t R, int SR, float* ray_color /*[R,3]*/,
                     float* opacity /*[R,SR]*/, float* bg_transmission /*[R]*/, void* stream);

/* Inference frame tail in one kernel: sgn_ray_dist -> sgn_composite_forward -> sgn_fill_invalid with the same results
 * (neural_points_volumetric_model.py:569-577, diff_ray_marching.py:509-555, neural_points_volumetric_model.py:158-195) and none
 * of the intermediate tensors.  ray_mask may be NULL (no fill); ray_color / opacity / bg_transmission / depth may be NULL.
 * depth [R] = sum_i w_i z_i / (sum_i w_i + 1e-6) with w = opacity * acc_transmission and z the samples' camera depth
 * (`coarse_depth`, neural_points_volumetric_model.py:620-624); 0 for rays that missed. */
int sgn_render_composite(const float* decoded /*[R,SR,4]*/, const float* loc_pers /*[R,SR,3]*/, const uint8_t* ray_valid /*[R,SR]*/,
                         const int8_t* ray_mask /*[R]*/, float vsize_z, int raydist_mode_unit, const float* bg /*[3]*/, int blend,
                         int64_t R, int SR, float* ray_color /*[R,3]*/, float* opacity /*[R,SR]*/, float* bg_transmission /*[R]*/,
                         float* depth /*[R]*/, void* stream);

/* The same from the dense depth array of sgn_agg_forward_frame (loc_depth [R,SR] = loc_pers[..., 2]): identical results, a third of the
 * position bytes. */
int sgn_render_composite_depth(const float* decoded /*[R,SR,4]*/, const float* loc_depth /*[R,SR]*/, const uint8_t* ray_valid /*[R,SR]*/,
                               const int8_t* ray_mask /*[R]*/, float vsize_z, int raydist_mode_unit, const float* bg /*[3]*/, int blend,
                               int64_t R, int SR, float* ray_color /*[R,3]*/, float* opacity /*[R,SR]*/, float* bg_transmission /*[R]*/,
                               float* depth /*[R]*/, void* stream);

/* `prob == 1` outputs of NeuralPointsRayMarching.forward (neural_points_volumetric_model.py:633-656), the inputs of probe_hole / point
 * growing (run/train_ft.py:425-540): per ray the first sample of largest opacity, its world position, the distance to its nearest
 * gathered neighbour (all K slots; invalid ones hold point 0 as clamp(pidx, 0) gathers it) and the weight * conf_coefficient sums of
 * its neighbours' colour / dir / conf / embedding.  Rows are per input ray (uncompacted); rays with ray_mask <= 0 get zeros
 * (ray_mask may be NULL).  Outputs [R], [R,3], [R], [R,3], [R,3], [R], [R,feat_dim]; any may be NULL. */
int sgn_probe_outputs(const float* opacity /*[R,SR]*/, const float* sample_loc_w /*[R,SR,3]*/, const int32_t* sample_pidx /*[R,SR,K]*/,
                      const float* weight /*[R,SR,K]*/, const float* conf_coef /*[R,SR,K]*/, const int8_t* ray_mask /*[R]*/,
                      const SgnPointTables* tables, int feat_dim, int64_t R, int SR, int K, float* ray_max_shading_opacity,
                      float* ray_max_sample_loc_w, float* ray_max_far_dist, float* shading_avg_color, float* shading_avg_dir,
                      float* shading_avg_conf, float* shading_avg_embedding, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Perspective-frustum querier (--wcoord_query 0).  Replaces get_occ_vox / near_vox_full / insert_vox_points /
 * query_neigh_along_ray_layered (NN > 0) / query_rand_along_ray (NN <= 0) and the torch glue of query_grid_point_index
 * (models/neural_points/query_point_indices.py:263-782) for all R rays of one camera in one call, uncompacted: row r of the outputs
 * belongs to input ray r (the reference keeps the rays with ray_mask > 0, :688).  The grid is in the camera's perspective coordinates
 * and is rebuilt per call.  Not reproduced: the reference's int8 overflow (> 127 selected voxels in one pixel column, :696) and a max_o
 * smaller than a column's selected-voxel count (lists are built for every voxel that holds a point).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float shift[3];          /* ranges[:3]                                                         (:604) */
    float vsize[3];          /* scaled_vsize = vsize * vscale                                      (:66)  */
    int32_t dim[3];          /* scaled_vdim                                                        (:65)  */
    int32_t vscale[3];
    int32_t kernel_size[3];
    int32_t query_size[3];
    float ray_vsize[3];      /* scaled_vsize / vscale                                              (:711) */
    int32_t P, SR, K, NN, inverse;
    float radius2, depth2;   /* radius_limit^2, depth_limit^2                                      (:749-750) */
    uint64_t seconds_insert; /* time.time() at :713 (P-cap reservoir, seed = index + seconds)              */
    uint64_t seconds_query;  /* time.time() at :737 (query_rand_along_ray)                                  */
} SgnPersCfg;
int sgn_pers_query_bytes(int64_t N, int64_t R, const SgnPersCfg* cfg, size_t* bytes);
/* xyz_pers [N,3] = (x/z, y/z, z) of every point (NeuralPoints.w2pers); pixel_idx int32 [R,2].
 * Outputs: sample_pidx int32 [R,SR,K] (-1 = empty), sample_loc f32 [R,SR,3] (perspective coordinates, :452-463), ray_mask int8 [R]. */
int sgn_pers_query(const float* xyz_pers, int64_t N, const int32_t* pixel_idx, int64_t R, const SgnPersCfg* cfg, void* workspace,
                   size_t workspace_bytes, int32_t* sample_pidx, float* sample_loc, int8_t* ray_mask, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Voxel-grid helpers next to the render path.
 * ---------------------------------------------------------------------------------------------- */
/* Point-cloud initialisation, construct_vox_points_closest (models/mvs/mvs_utils.py:536-561; run/train_ft.py:141, :715): one point per
 * occupied voxel of the grid floor((xyz - space_min) / vox_size).  Voxels come out in the lexicographic order of their coordinates
 * (what torch.unique(dim=0) returns): centroid [V,3] = mean of the voxel's points, grid_idx int32 [V,3], min_idx int64 [V] = index of the
 * point closest to the centroid (first one on ties), *count = V (device).  Outputs must hold N entries.  space_min / vox_size are
 * [host] float[3]. */
int sgn_voxel_downsample_bytes(int64_t N, size_t* bytes);
int sgn_voxel_downsample(const float* xyz /*[N,3]*/, int64_t N, const float* space_min /*[host]*/, const float* vox_size /*[host]*/,
                         int vox_res, void* workspace, size_t workspace_bytes, float* centroid, int32_t* grid_idx, int64_t* min_idx,
                         int32_t* count, void* stream);
/* NeuralPoints.query_vox_grid (models/neural_points/neural_points.py:814-826, the NN < 0 query): out int64 [n_samples, 8] = indices of
 * the 8 corners of the grid cell each sample falls in (full_grid_idx int32 [(grid_res+1)^3], -1 = no grid point), all -1 unless the
 * cell lies inside the grid and its 8 corners exist.  space_min is a [host] float[3]. */
int sgn_query_vox_grid(const float* sample_loc_w /*[n_samples,3]*/, int64_t n_samples, const int32_t* full_grid_idx, int grid_res,
                       const float* space_min /*[host]*/, float grid_vox_sz, int64_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step glue.  Replaces BaseRenderingModel.compute_losses for the canonical loss items
 * (models/base_rendering_model.py:543-641: `ray_masked_coarse_raycolor` colour MSE over the rays that hit the cloud + 1e-6, and the
 * zero-one regulariser mean(log v + log(1 - v)), v = clamp(conf_coefficient, zero_epsilon, 1 - zero_epsilon)) together with its
 * autograd backward, on UNCOMPACTED rows (ray_mask marks the hits), and torch.optim.Adam on the point tables
 * (models/mvs_points_volumetric_model.py:99-109).
 * ---------------------------------------------------------------------------------------------- */
/* *count += number of rays with ray_mask > 0 (device scalar; all-reduce it over ranks before the next call when ranks share a step). */
int sgn_loss_hit_count(const int8_t* ray_mask /*[R]*/, int64_t R, float* count, void* stream);
/* loss = color_weight * sum_hit |ray_color - gt|^2 / (3 n) + conf_weight * sum_hit (log v + log(1 - v)) / (n SR K) + const_term with
 * n = max(*hit_count, 1); *loss is overwritten; d_ray_color [R,3] and d_conf_coef [R,SR,K] (either may be NULL) receive d loss / d input
 * (zero for rays that missed and where the clamp is active).  conf_coef may be NULL (no regulariser). */
int sgn_loss_forward_backward(const float* ray_color /*[R,3]*/, const float* gt /*[R,3]*/, const int8_t* ray_mask /*[R]*/,
                              const float* conf_coef /*[R,SR,K]*/, int64_t R, int SR, int K, const float* hit_count, float color_weight,
                              float conf_weight, float zero_eps, float const_term, float* loss, float* d_ray_color, float* d_conf_coef,
                              void* stream);
/* *step += 1 (the optimiser's step counter lives on the device so that a captured step needs no host value). */
int sgn_adam_step_count(float* step, void* stream);
/* One Adam update (betas, eps, no weight decay, bias correction with t = *step, torch.optim.Adam's arithmetic) of the rows of a [N,C]
 * table whose gradient is non-zero now or was at any earlier step (`active` [N] bytes, maintained by the call; NULL = all rows): rows
 * that never received a gradient have zero moments and are exactly where dense Adam leaves them.  grad is multiplied by grad_scale
 * first; zero_grad != 0 clears the gradient rows it consumed (the accumulator then never needs a dense memset). */
int sgn_adam_rows(float* param, float* grad, float* exp_avg, float* exp_avg_sq, uint8_t* active, int64_t N, int C, float lr, float beta1,
                  float beta2, float eps, const float* step, float grad_scale, int zero_grad, void* stream);

/* The same for up to 8 tables [N, C_k] (C_k <= 32) that share their rows -- the point tables -- in one pass: a row is active when any
 * table has a non-zero gradient in it.  params / grads / exp_avg / exp_avg_sq and C are [host] arrays of n_tables entries. */
int sgn_adam_rows_multi(int n_tables, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                        const int32_t* C /*[host]*/, uint8_t* active, int64_t N, float lr, float beta1, float beta2, float eps,
                        const float* step, float grad_scale, int zero_grad, void* stream);

/* Dense Adam over n_tensors small tensors (the MLP's weights and biases) in one launch; same arithmetic as sgn_adam_rows with every
 * element active.  params / grads / exp_avg / exp_avg_sq / numel are [host] arrays. */
int sgn_adam_dense_multi(int n_tensors, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                         const int64_t* numel /*[host]*/, float lr, float beta1, float beta2, float eps, const float* step, float grad_scale,
                         int zero_grad, void* stream);

/* The same update driven by a list of the active rows, for steps that touch a small part of the cloud (a 56x56 patch touches ~5 % of 1M
 * points): nothing is read for the other rows.  sgn_adam_mark_rows sets touched[r] = 1 for every r >= 0 of `rows` (the step's
 * sample_pidx: a superset of the rows that receive a gradient); with several ranks `touched` is summed with the gradients.
 * sgn_adam_rows_list then (1) clears `touched`, appends the touched rows that have a non-zero gradient and were not active yet to
 * active_list (int32 [N]) / active_count (int32 [1], device) and sets their `active` flag, (2) updates the listed rows exactly as
 * sgn_adam_rows_multi does.  Same results as sgn_adam_rows_multi (and as dense Adam) bit for bit. */
int sgn_adam_mark_rows(const int32_t* rows, int64_t n, float* touched, void* stream);
int sgn_adam_rows_list(int n_tables, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                       const int32_t* C /*[host]*/, uint8_t* active, int32_t* active_list, int32_t* active_count, float* touched, int64_t N,
                       float lr, float beta1, float beta2, float eps, const float* step, float grad_scale, int zero_grad, void* stream);

/* Exchange of the touched rows only (data-parallel training, replaces the dense all-reduce of the point tables' gradients -- the flatten /
 * all-reduce / un-flatten of torch DDP in the reference's multi-GPU scripts).  With `touched` already summed over the ranks,
 * sgn_rows_union writes the ascending list of rows with touched != 0 and their number (int32 [1], device) -- identical on every rank;
 * sgn_rows_pack copies those rows of n_tables tables [N, C_k] into packed [count, stride] (unpack = 0) or back (unpack = 1).
 * The caller all-reduces packed[:count * stride]. */
int sgn_rows_union_bytes(int64_t N, size_t* bytes);
int sgn_rows_union(const float* touched, int64_t N, int32_t* list, int32_t* count, void* workspace, size_t workspace_bytes, void* stream);
int sgn_rows_pack(int n_tables, float* const* tables, const int32_t* C /*[host]*/, const int32_t* list, const int32_t* count, int64_t N,
                  float* packed, int stride, int unpack, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGNERF_B200_H */
