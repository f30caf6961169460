// agg_api.cu -- C-ABI entry points of the aggregator; dispatch on precision.
#include <stdlib.h>

#include "agg_common.cuh"

using namespace sgn;

int sgn_agg_fp32_workspace_bytes(const AggPlan& P, int64_t R, int SR, int K, int save, size_t* bytes);
int sgn_agg_fp32_forward(const AggPlan& P, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                         const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                         int64_t R, int SR, int K, int save, float* decoded, uint8_t* ray_valid, float* loc_pers, float* loc_depth, float* weight,
                         float* conf_coef, void* workspace, size_t workspace_bytes, bool tc, cudaStream_t st);
int sgn_agg_fp32_backward(const AggPlan& P, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                          const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                          int64_t R, int SR, int K, const float* d_decoded, const float* d_conf_coef, float* const* d_weights,
                          float* const* d_biases, const SgnPointGrads* d_tables, void* workspace, size_t workspace_bytes, bool tc, cudaStream_t st);
int sgn_agg_tc_workspace_bytes(const AggPlan& P, int64_t N, int64_t R, int SR, int K, size_t* bytes);
int sgn_agg_tc_forward(const AggPlan& P, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                       const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                       int64_t R, int SR, int K, float* decoded, uint8_t* ray_valid, float* loc_pers, float* loc_depth, float* weight, float* conf_coef,
                       void* workspace, size_t workspace_bytes, const void* point_cache, cudaStream_t st);
int sgn_agg_tc_point_cache_bytes(const AggPlan& P, int64_t N, size_t* bytes);
int sgn_agg_tc_point_cache_build(const AggPlan& P, const float* const* weights, const SgnPointTables* tables, void* cache, size_t cache_bytes, cudaStream_t st);
int sgn_agg_tc_point_cache_update(const AggPlan& P, const SgnPointTables* tables, void* cache, size_t cache_bytes, const int32_t* rows, int64_t n_rows,
                                  cudaStream_t st);

static int check_common(const SgnAggCfg* cfg, AggPlan* P, int64_t R, int SR, int K)
{
    int rc = make_plan(cfg, P);
    if (rc) return rc;
    SGN_CHECK_ARG(R >= 0 && SR > 0 && K > 0 && K <= SGN_MAX_K, "aggregator: bad R/SR/K");
    SGN_CHECK_ARG(R * (int64_t)SR * K < (1ll << 31), "aggregator: R*SR*K exceeds int32 indexing; split the rays");
    return SGN_OK;
}

extern "C" int sgn_agg_workspace_bytes(const SgnAggCfg* cfg, int64_t N, int64_t R, int SR, int K, int precision, int save_for_backward, size_t* bytes)
{
    AggPlan P;
    int rc = check_common(cfg, &P, R, SR, K);
    if (rc) return rc;
    SGN_CHECK_ARG(bytes != nullptr, "sgn_agg_workspace_bytes: bytes is NULL");
    if (precision == SGN_PRECISION_FP32 || precision == SGN_PRECISION_TF32) return sgn_agg_fp32_workspace_bytes(P, R, SR, K, save_for_backward, bytes);
    SGN_CHECK_ARG(precision == SGN_PRECISION_BF16, "aggregator: unknown precision %d", precision);
    SGN_CHECK_ARG(!save_for_backward, "aggregator: the bf16 tensor-core path is forward-only; train with SGN_PRECISION_FP32 or SGN_PRECISION_TF32");
    SGN_CHECK_ARG(N >= 0, "sgn_agg_workspace_bytes: bad N");
    return sgn_agg_tc_workspace_bytes(P, N, R, SR, K, bytes);
}

extern "C" int sgn_agg_forward_frame(const SgnAggCfg* cfg, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                                     const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                                     int64_t R, int SR, int K, int precision, int save_for_backward, float* decoded, uint8_t* ray_valid,
                                     float* loc_pers, float* loc_depth, float* weight, float* conf_coef, void* workspace, size_t workspace_bytes,
                                     const void* point_cache, void* stream)
{
    AggPlan P;
    int rc = check_common(cfg, &P, R, SR, K);
    if (rc) return rc;
    SGN_CHECK_ARG(weights && biases && tables && pidx && loc_w && raydir && campos && camrotc2w && decoded && ray_valid, "sgn_agg_forward: NULL argument");
    SGN_CHECK_ARG(tables->xyz && tables->embedding && tables->color && tables->dir, "sgn_agg_forward: xyz/embedding/color/dir tables are required");
    SGN_CHECK_ARG(P.dims.LD == 0 || tables->label_emb, "sgn_agg_forward: label embedding table missing");
    if (R == 0) return SGN_OK;
    if (precision == SGN_PRECISION_FP32 || precision == SGN_PRECISION_TF32)
        return sgn_agg_fp32_forward(P, weights, biases, tables, pidx, loc_w, raydir, campos, camrotc2w, R, SR, K, save_for_backward, decoded,
                                    ray_valid, loc_pers, loc_depth, weight, conf_coef, workspace, workspace_bytes, precision == SGN_PRECISION_TF32,
                                    (cudaStream_t)stream);
    SGN_CHECK_ARG(precision == SGN_PRECISION_BF16, "aggregator: unknown precision %d", precision);
    SGN_CHECK_ARG(!save_for_backward, "aggregator: the bf16 tensor-core path is forward-only; train with SGN_PRECISION_FP32 or SGN_PRECISION_TF32");
    return sgn_agg_tc_forward(P, weights, biases, tables, pidx, loc_w, raydir, campos, camrotc2w, R, SR, K, decoded, ray_valid, loc_pers, loc_depth,
                              weight, conf_coef, workspace, workspace_bytes, point_cache, (cudaStream_t)stream);
}

thread_local const int32_t* sgn::g_agg_sample_mask = nullptr;

extern "C" int sgn_agg_forward_frame_masked(const SgnAggCfg* cfg, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                                            const int32_t* pidx, const int32_t* sample_mask, const float* loc_w, const float* raydir, const float* campos,
                                            const float* camrotc2w, int64_t R, int SR, int K, int precision, int save_for_backward, float* decoded,
                                            uint8_t* ray_valid, float* loc_pers, float* loc_depth, float* weight, float* conf_coef, void* workspace,
                                            size_t workspace_bytes, const void* point_cache, void* stream)
{
    g_agg_sample_mask = sample_mask;
    const int rc = sgn_agg_forward_frame(cfg, weights, biases, tables, pidx, loc_w, raydir, campos, camrotc2w, R, SR, K, precision, save_for_backward, decoded,
                                         ray_valid, loc_pers, loc_depth, weight, conf_coef, workspace, workspace_bytes, point_cache, stream);
    g_agg_sample_mask = nullptr;
    return rc;
}

extern "C" int sgn_agg_forward_cached(const SgnAggCfg* cfg, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                               const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                               int64_t R, int SR, int K, int precision, int save_for_backward, float* decoded, uint8_t* ray_valid,
                               float* loc_pers, float* weight, float* conf_coef, void* workspace, size_t workspace_bytes, const void* point_cache,
                                      void* stream)
{
    return sgn_agg_forward_frame(cfg, weights, biases, tables, pidx, loc_w, raydir, campos, camrotc2w, R, SR, K, precision, save_for_backward,
                                 decoded, ray_valid, loc_pers, nullptr, weight, conf_coef, workspace, workspace_bytes, point_cache, stream);
}

extern "C" int sgn_agg_forward(const SgnAggCfg* cfg, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                               const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                               int64_t R, int SR, int K, int precision, int save_for_backward, float* decoded, uint8_t* ray_valid,
                               float* loc_pers, float* weight, float* conf_coef, void* workspace, size_t workspace_bytes, void* stream)
{
    return sgn_agg_forward_cached(cfg, weights, biases, tables, pidx, loc_w, raydir, campos, camrotc2w, R, SR, K, precision, save_for_backward,
                                  decoded, ray_valid, loc_pers, weight, conf_coef, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int sgn_agg_point_cache_bytes(const SgnAggCfg* cfg, int64_t N, size_t* bytes)
{
    AggPlan P;
    int rc = make_plan(cfg, &P);
    if (rc) return rc;
    SGN_CHECK_ARG(bytes != nullptr && N >= 0, "sgn_agg_point_cache_bytes: bad argument");
    return sgn_agg_tc_point_cache_bytes(P, N, bytes);
}

extern "C" int sgn_agg_point_cache_build(const SgnAggCfg* cfg, const float* const* weights, const SgnPointTables* tables, void* cache,
                                         size_t cache_bytes, void* stream)
{
    AggPlan P;
    int rc = make_plan(cfg, &P);
    if (rc) return rc;
    SGN_CHECK_ARG(weights && tables && cache && tables->embedding, "sgn_agg_point_cache_build: NULL argument");
    SGN_CHECK_ARG(P.dims.LD == 0 || tables->label_emb, "sgn_agg_point_cache_build: label embedding table missing");
    return sgn_agg_tc_point_cache_build(P, weights, tables, cache, cache_bytes, (cudaStream_t)stream);
}

extern "C" int sgn_agg_point_cache_update(const SgnAggCfg* cfg, const SgnPointTables* tables, void* cache, size_t cache_bytes, const int32_t* rows,
                                          int64_t n_rows, void* stream)
{
    AggPlan P;
    int rc = make_plan(cfg, &P);
    if (rc) return rc;
    SGN_CHECK_ARG(tables && cache && tables->embedding && (rows || n_rows == 0) && n_rows >= 0, "sgn_agg_point_cache_update: bad argument");
    SGN_CHECK_ARG(P.dims.LD == 0 || tables->label_emb, "sgn_agg_point_cache_update: label embedding table missing");
    return sgn_agg_tc_point_cache_update(P, tables, cache, cache_bytes, rows, n_rows, (cudaStream_t)stream);
}

extern "C" int sgn_agg_backward_prec(const SgnAggCfg* cfg, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                                     const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                                     int64_t R, int SR, int K, int precision, const float* d_decoded, const float* d_conf_coef,
                                     float* const* d_weights, float* const* d_biases, const SgnPointGrads* d_tables, void* workspace,
                                     size_t workspace_bytes, void* stream)
{
    AggPlan P;
    int rc = check_common(cfg, &P, R, SR, K);
    if (rc) return rc;
    SGN_CHECK_ARG(weights && tables && pidx && raydir && d_decoded, "sgn_agg_backward: NULL argument");
    SGN_CHECK_ARG(precision == SGN_PRECISION_FP32 || precision == SGN_PRECISION_TF32, "sgn_agg_backward: precision must be FP32 or TF32");
    if (R == 0) return SGN_OK;
    return sgn_agg_fp32_backward(P, weights, biases, tables, pidx, loc_w, raydir, campos, camrotc2w, R, SR, K, d_decoded, d_conf_coef,
                                 d_weights, d_biases, d_tables, workspace, workspace_bytes, precision == SGN_PRECISION_TF32, (cudaStream_t)stream);
}

extern "C" int sgn_agg_backward(const SgnAggCfg* cfg, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                                const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                                int64_t R, int SR, int K, const float* d_decoded, const float* d_conf_coef, float* const* d_weights,
                                float* const* d_biases, const SgnPointGrads* d_tables, void* workspace, size_t workspace_bytes, void* stream)
{
    return sgn_agg_backward_prec(cfg, weights, biases, tables, pidx, loc_w, raydir, campos, camrotc2w, R, SR, K, SGN_PRECISION_FP32, d_decoded,
                                 d_conf_coef, d_weights, d_biases, d_tables, workspace, workspace_bytes, stream);
}
