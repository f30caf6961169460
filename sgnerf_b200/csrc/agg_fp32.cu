// agg_fp32.cu -- neighbour aggregation, fp32 strict-parity path (forward + backward), layer-wise.
//
// Reference: NeuralPoints.forward gathers + w2pers (models/neural_points/neural_points.py:956-988, :838-850),
// PointAggregator.forward / linear / viewmlp / gradiant_clamp / raw2out_* (models/aggregators/
// point_aggregators.py:868-959, :494-502, :561-786, :863-865, :298-309), positional_encoding
// (models/helpers/networks.py:175-192), and the autograd backward of all of it.
//
// Work decomposition (T = valid (sample,neighbour) tuples, S = samples with >= 1 neighbour; both known
// only on the device, every kernel reads them from the workspace and the host never synchronises):
//   prepare  per sample : loc_pers, inverse-distance weights, conf clamp, counts
//   scan     compaction offsets (tuple <- sample,slot ; compact sample <- sample)
//   gather   per tuple  : X0 = [emb | PE(emb) | PE(dists)], label embedding, [colour | dir - view | dir.view]
//   layers   GEMMs with fused bias + LeakyReLU, concat inputs as a second operand pair: fp32 SIMT (gemm_simt.cuh, SGN_PRECISION_FP32)
//            or tcgen05 kind::tf32 (gemm_tc.cuh, SGN_PRECISION_TF32: what training uses by default)
//   alpha    per tuple  : raw sigma ; ksum per sample : sum_k w conf (sigma, h) ; colour MLP ; rgb
// The fused bf16 tcgen05 inference path lives in agg_tc.cu; this file is the layer-wise one (strict fp32 or TF32 GEMMs), the only
// one with a backward.
#include <algorithm>
#include <map>
#include <mutex>

#include "agg_kernels.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace sgn {

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------

// out[p, c] += sum_m w_p[m] X[m, c] (w = A[:, p], or 1 when A is NULL): bias gradients and the wgrads of the 1- / 3-output layers
static int launch_skinny(const float* A, int lda, int np, const float* X, int ld, int ncols, const int32_t* m_ptr, int m_max, float* out, int ldo,
                         cudaStream_t st)
{
    if (!out || m_max <= 0) return SGN_OK;
    dim3 grid(cdiv(m_max, SKINNY_ROWS), cdiv(ncols, 256));
    if (np == 1) launch(skinny_tn_kernel<1>, grid, 256, 0, st, A, lda, X, ld, ncols, m_ptr, m_max, out, ldo);
    else launch(skinny_tn_kernel<3>, grid, 256, 0, st, A, lda, X, ld, ncols, m_ptr, m_max, out, ldo);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

// GEMM dispatch: tc = TF32 tensor-core kernels (gemm_tc.cuh) where the operand shapes allow, fp32 SIMT otherwise.
// g.colsum (column sums of the result = the bias gradient of the layer below) is fused into the tensor-core epilogue and is a
// separate pass over C on the SIMT path.
// The dgrad epilogues multiply by leaky'(activation): a saving forward also stores the activations' sign bits (by the tensor-core GEMM's
// epilogue, or by mask_from_act_kernel after a SIMT GEMM) and the tensor-core dgrads read those instead (1/32 of the bytes).  The masks
// exist whatever arithmetic the forward ran in, so forward and backward may be called with different precisions.
static inline bool use_sign_masks(const AggDims& d) { return d.W % 32 == 0 && d.WC % 32 == 0; }

// grid of the warp-per-item kernels (8 warps per block): they stride over the items, whose number is only known on the device
static inline int item_grid(int64_t max_items) { return (int)std::min<int64_t>(cdiv(max_items, 8), 148 * 8); }
// the same, sized to what is resident at once (one full wave: a second, partly filled wave of a strided loop only adds a tail)
template <typename Kern>
static int item_grid(Kern kernel, int64_t max_items, size_t smem = 0)
{
    static std::mutex mu;
    static std::map<const void*, int> per_sm;
    int occ;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = per_sm.find((const void*)kernel);
        if (it == per_sm.end()) {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, smem) != cudaSuccess || n < 1) n = 4;
            it = per_sm.emplace((const void*)kernel, n).first;
        }
        occ = it->second;
    }
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::min<int64_t>(cdiv(max_items, 8), (int64_t)sms * occ);
}

// sign bits of a stored activation [M, N] (N a multiple of 32): the fallback producer of GemmNN::mask_out for layers the SIMT GEMM ran
static __global__ void mask_from_act_kernel(const float* __restrict__ act, int ld, int nwords, const int32_t* __restrict__ m_ptr, int m_max, uint32_t* __restrict__ mask)
{
    const int M = min(*m_ptr, m_max);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)M * nwords; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / nwords;
        const int w = (int)(i - m * nwords);
        const float* a = act + m * ld + 32 * w;
        uint32_t word = 0u;
        for (int b = 0; b < 32; b++) word |= (a[b] > 0.f ? 1u : 0u) << b;
        mask[i] = word;
    }
}

static int run_gemm_nn(const GemmNN& g, bool tc, cudaStream_t st)
{
    if (tc && gemm_tc_nn_ok(g)) return launch_gemm_tc_nn(g, st);
    if (g.mask_out) {              // the SIMT kernel does not write sign bits: derive them from the activation it stored
        int rc = launch_gemm_nn(g, st);
        if (rc) return rc;
        launch(mask_from_act_kernel, 148 * 4, 256, 0, st, g.C, g.ldc, g.N / 32, g.m_ptr, g.m_max, g.mask_out);
        SGN_LAUNCH_CHECK();
        return SGN_OK;
    }
    int rc = launch_gemm_nn(g, st);
    if (rc || !g.colsum) return rc;
    return launch_skinny(nullptr, 0, 1, g.C, g.ldc, g.N, g.m_ptr, g.m_max, g.colsum, 0, st);
}
static int run_gemm_tn(const GemmTN& t, bool tc, cudaStream_t st)
{
    if (tc && gemm_tc_tn_ok(t)) return launch_gemm_tc_tn(t, st);
    return launch_gemm_tn(t, st);
}

static int pack_weights(const AggPlan& P, const float* const* weights, AggWs& ws, cudaStream_t st)
{
    if (P.n_layers <= 16) {
        PackJobs J = {};
        for (int l = 0; l < P.n_layers; l++) {
            const LayerInfo& L = P.layers[l];
            J.W[l] = weights[l]; J.Wt[l] = ws.Wt[l]; J.Wp[l] = ws.Wp[l]; J.N[l] = L.out; J.Kin[l] = L.in; J.Npad[l] = L.npad; J.Kpad[l] = L.kpad;
        }
        launch(pack_weights_kernel, dim3(64, P.n_layers), 256, 0, st, J);
    } else {
        for (int l = 0; l < P.n_layers; l++) {
            const LayerInfo& L = P.layers[l];
            const int n = L.npad * L.kpad;
            launch(pack_weight_kernel, cdiv(n, 256), 256, 0, st, weights[l], L.out, L.in, L.npad, L.kpad, ws.Wt[l], ws.Wp[l]);
        }
    }
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

// forward for one chunk of rays
static int agg_forward_chunk(const AggPlan& P, const float* const* weights, const float* const* biases, const AggIn& in, int64_t Rc,
                             int SR, int K, float* decoded, uint8_t* ray_valid, float* loc_pers_out, float* loc_depth, float* weight_out, float* conf_out,
                             AggWs& ws, bool save, bool tc, cudaStream_t st)
{
    const AggDims& d = P.dims;
    const int64_t S = Rc * SR;
    const int Tm = (int)(S * K), Sm = (int)S;
    float* loc_pers = loc_pers_out ? loc_pers_out : ws.loc_pers;
    SGN_CUDA(cudaMemsetAsync(decoded, 0, sizeof(float) * 4 * (size_t)S, st));
    launch(agg_prepare_kernel, cdiv(S, 128), 128, 0, st, in, S, K, loc_pers, loc_depth, ws.wc, ws.weight_n, weight_out, conf_out, ray_valid, ws.nvalid, ws.svalid);
    int rc;
    // compaction offsets of the tuples (scan of nvalid) and of the samples with a neighbour (scan of nvalid > 0), one pass
    if ((rc = exclusive_scan_pair_i32(ws.nvalid, ws.tuple_start, ws.sample_cidx, S, ws.partials, st))) return rc;
    const int32_t* T_ptr = ws.tuple_start + S;
    const int32_t* S_ptr = ws.sample_cidx + S;
    launch(agg_index_kernel, cdiv(S, 128), 128, 0, st, in.pidx, S, K, ws.tuple_start, ws.sample_cidx, ws.nvalid, ws.tuple_src, ws.csample, ws.tuple_pt, ws.tuple_cs);
    launch(agg_gather_kernel, item_grid(agg_gather_kernel, Tm, (size_t)8 * d.k0pad * sizeof(float)), 256, (size_t)8 * d.k0pad * sizeof(float), st, in, d, K, SR, T_ptr, Tm, ws.tuple_src, ws.tuple_pt, loc_pers, ws.X0, d.LD > 0 ? ws.L : nullptr, ws.E7);
    SGN_LAUNCH_CHECK();

    // per-tuple layers
    const float* cur = ws.X0;
    int cur_ld = d.k0pad, cur_k = d.k0pad;
    for (int t = 0; t < P.n_tuple_layers; t++) {
        const LayerInfo& L = P.layers[t];
        float* out = ws.H[save ? t : (t & 1)];
        GemmNN g = {};
        g.A1 = cur; g.lda1 = cur_ld; g.B1 = ws.Wt[t]; g.ldb1 = L.npad; g.K1 = cur_k;
        g.Bt1 = ws.Wp[t]; g.ldbt1 = L.kpad;
        if (L.extra == EXTRA_LABEL) { g.A2 = ws.L; g.lda2 = d.LD; g.B2 = ws.Wt[t] + (size_t)cur_k * L.npad; g.ldb2 = L.npad; g.K2 = d.LD; }
        if (L.extra == EXTRA_COLORDIR) { g.A2 = ws.E7; g.lda2 = 8; g.B2 = ws.Wt[t] + (size_t)cur_k * L.npad; g.ldb2 = L.npad; g.K2 = 8; }
        if (L.extra != EXTRA_NONE) { g.Bt2 = ws.Wp[t] + cur_k; g.ldbt2 = L.kpad; }
        g.C = out; g.ldc = d.W; g.N = d.W; g.m_ptr = T_ptr; g.m_max = Tm; g.bias = biases[t]; g.epi = EPI_BIAS_LEAKY; g.slope = d.slope;
        if (save && use_sign_masks(d)) { g.mask_out = ws.HM[t]; g.ldmask = d.W / 32; }
        if ((rc = run_gemm_nn(g, tc, st))) return rc;
        cur = out; cur_ld = d.W; cur_k = d.W;
    }
    const float* Hlast = cur;
    const int la = P.alpha_layer;
    // raw alpha of every tuple + the K-sums of every sample, one pass over H
    if (K <= 8)
        launch(agg_ksum_kernel<8>, item_grid(agg_ksum_kernel<8>, Sm), 256, 0, st, in, d, K, SR, S_ptr, Sm, ws.csample, ws.tuple_start, ws.nvalid, ws.wc, Hlast, weights[la], biases[la],
               ws.araw, ws.C0, ws.sigma);
    else
        launch(agg_ksum_kernel<SGN_MAX_K>, item_grid(Sm), 256, 0, st, in, d, K, SR, S_ptr, Sm, ws.csample, ws.tuple_start, ws.nvalid, ws.wc, Hlast, weights[la],
               biases[la], ws.araw, ws.C0, ws.sigma);
    SGN_LAUNCH_CHECK();

    // colour MLP
    cur = ws.C0; cur_ld = d.kc0pad; cur_k = d.kc0pad;
    for (int c = 0; c < P.n_color_hidden; c++) {
        const int l = P.color_layer0 + c;
        const LayerInfo& L = P.layers[l];
        float* out = ws.CH[save ? c : (c & 1)];
        GemmNN g = {};
        g.A1 = cur; g.lda1 = cur_ld; g.B1 = ws.Wt[l]; g.ldb1 = L.npad; g.K1 = cur_k;
        g.Bt1 = ws.Wp[l]; g.ldbt1 = L.kpad;
        g.C = out; g.ldc = d.WC; g.N = d.WC; g.m_ptr = S_ptr; g.m_max = Sm; g.bias = biases[l]; g.epi = EPI_BIAS_LEAKY; g.slope = d.slope;
        if (save && use_sign_masks(d)) { g.mask_out = ws.CM[c]; g.ldmask = d.WC / 32; }
        if ((rc = run_gemm_nn(g, tc, st))) return rc;
        cur = out; cur_ld = d.WC; cur_k = d.WC;
    }
    const int ll = P.n_layers - 1;
    launch(agg_rgb_kernel, item_grid(Sm), 256, 0, st, d, S_ptr, Sm, ws.csample, cur, cur_ld, weights[ll], biases[ll], ws.sigma, decoded, ws.sig);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_agg_num_layers(const SgnAggCfg* cfg)
{
    AggPlan P;
    if (make_plan(cfg, &P)) return SGN_E_INVALID;
    return P.n_layers;
}

extern "C" int sgn_agg_layer_shape(const SgnAggCfg* cfg, int layer, int* in_features, int* out_features)
{
    AggPlan P;
    int rc = make_plan(cfg, &P);
    if (rc) return rc;
    SGN_CHECK_ARG(layer >= 0 && layer < P.n_layers, "sgn_agg_layer_shape: layer %d out of range", layer);
    if (in_features) *in_features = P.layers[layer].in;
    if (out_features) *out_features = P.layers[layer].out;
    return SGN_OK;
}

int sgn_agg_fp32_workspace_bytes(const AggPlan& P, int64_t R, int SR, int K, int save, size_t* bytes)
{
    AggWs ws;
    const int64_t Rc = save ? R : (R < AGG_FP32_CHUNK ? R : AGG_FP32_CHUNK);
    *bytes = carve_ws(P, Rc, SR, K, save != 0, nullptr, 0, &ws);
    return SGN_OK;
}

int sgn_agg_fp32_forward(const AggPlan& P, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                         const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                         int64_t R, int SR, int K, int save, float* decoded, uint8_t* ray_valid, float* loc_pers, float* loc_depth, float* weight,
                         float* conf_coef, void* workspace, size_t workspace_bytes, bool tc, cudaStream_t st)
{
    const int64_t chunk = save ? R : (R < AGG_FP32_CHUNK ? R : AGG_FP32_CHUNK);
    AggWs ws;
    size_t need = carve_ws(P, chunk, SR, K, save != 0, workspace, workspace_bytes, &ws);
    if (need > workspace_bytes || ((uintptr_t)workspace & 255)) {
        set_error("sgn_agg_forward: workspace too small or misaligned (need %zu bytes, got %zu)", need, workspace_bytes);
        return SGN_E_WORKSPACE;
    }
    int rc = pack_weights(P, weights, ws, st);
    if (rc) return rc;
    for (int64_t r0 = 0; r0 < R; r0 += chunk) {
        const int64_t Rc = R - r0 < chunk ? R - r0 : chunk;
        AggIn in;
        in.tab = *tables;
        in.pidx = pidx + r0 * SR * K; in.loc_w = loc_w + r0 * SR * 3; in.raydir = raydir + r0 * 3;
        in.smask = g_agg_sample_mask ? g_agg_sample_mask + r0 * SR : nullptr;
        in.campos = campos; in.camrot = camrotc2w;
        rc = agg_forward_chunk(P, weights, biases, in, Rc, SR, K, decoded + r0 * SR * 4, ray_valid + r0 * SR,
                               loc_pers ? loc_pers + r0 * SR * 3 : nullptr, loc_depth ? loc_depth + r0 * SR : nullptr, weight ? weight + r0 * SR * K : nullptr,
                               conf_coef ? conf_coef + r0 * SR * K : nullptr, ws, save != 0, tc, st);
        if (rc) return rc;
    }
    return SGN_OK;
}

int sgn_agg_fp32_backward(const AggPlan& P, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                          const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                          int64_t R, int SR, int K, const float* d_decoded, const float* d_conf_coef, float* const* d_weights,
                          float* const* d_biases, const SgnPointGrads* d_tables, void* workspace, size_t workspace_bytes, bool tc, cudaStream_t st)
{
    (void)biases; (void)loc_w;
    const AggDims& d = P.dims;
    AggWs ws;
    size_t need = carve_ws(P, R, SR, K, true, workspace, workspace_bytes, &ws);
    if (need > workspace_bytes || ((uintptr_t)workspace & 255)) {
        set_error("sgn_agg_backward: workspace too small or misaligned (need %zu bytes, got %zu)", need, workspace_bytes);
        return SGN_E_WORKSPACE;
    }
    const int64_t S = R * SR;
    const int Tm = (int)(S * K), Sm = (int)S;
    const int32_t* T_ptr = ws.tuple_start + S;
    const int32_t* S_ptr = ws.sample_cidx + S;
    AggIn in;
    in.tab = *tables; in.pidx = pidx; in.loc_w = loc_w; in.raydir = raydir; in.campos = campos; in.camrot = camrotc2w; in.smask = nullptr;
    SgnPointGrads g = {};
    if (d_tables) g = *d_tables;
    int rc;
    auto skinny = [&](const float* A, int lda, int np, const float* X, int ld, int ncols, const int32_t* m_ptr, int m_max, float* out, int ldo) -> int {
        return launch_skinny(A, lda, np, X, ld, ncols, m_ptr, m_max, out, ldo, st);
    };
    auto bias_grad = [&](const float* dZ, int ld, int ncols, const int32_t* m_ptr, int m_max, float* out) -> int {
        return skinny(nullptr, 0, 1, dZ, ld, ncols, m_ptr, m_max, out, 0);
    };
    auto wgrad = [&](const float* dZ, int ldz, int Pn, const float* act, int lda, int Q, float* out, int ldc, const int32_t* m_ptr, int m_max) -> int {
        if (!out) return SGN_OK;
        if (Pn == 1 || Pn == 3) return skinny(dZ, ldz, Pn, act, lda, Q, m_ptr, m_max, out, ldc);   // alpha_branch.0, color_branch.6
        GemmTN t = {};
        t.A = dZ; t.lda = ldz; t.P = Pn; t.B = act; t.ldb = lda; t.Q = Q; t.C = out; t.ldc = ldc; t.m_ptr = m_ptr; t.m_max = m_max;
        return run_gemm_tn(t, tc, st);
    };

    // ---- colour branch ----
    const int ll = P.n_layers - 1;
    const float* Clast = P.n_color_hidden > 0 ? ws.CH[P.n_color_hidden - 1] : ws.C0;
    const int Clast_ld = P.n_color_hidden > 0 ? d.WC : d.kc0pad;
    const int Clast_k = P.n_color_hidden > 0 ? d.WC : d.kc0;
    launch(agg_rgb_bwd_kernel, cdiv(Sm, 256), 256, 0, st, d, S_ptr, Sm, ws.csample, d_decoded, ws.sig, ws.d_raw);
    SGN_LAUNCH_CHECK();
    if ((rc = wgrad(ws.d_raw, 8, 3, Clast, Clast_ld, Clast_k, d_weights ? d_weights[ll] : nullptr, P.layers[ll].in, S_ptr, Sm))) return rc;
    if ((rc = bias_grad(ws.d_raw, 8, 3, S_ptr, Sm, d_biases ? d_biases[ll] : nullptr))) return rc;
    // dgrad through the last colour Linear: [S,8] x Wp[8, kpad]
    float* dcur = ws.dC[0];
    int dcur_ld;
    {
        GemmNN q = {};
        q.A1 = ws.d_raw; q.lda1 = 8; q.B1 = ws.Wp[ll]; q.ldb1 = P.layers[ll].kpad; q.K1 = 8;
        q.Bt1 = ws.Wt[ll]; q.ldbt1 = P.layers[ll].npad;
        const bool to_c0 = P.n_color_hidden == 0;
        q.C = dcur; q.ldc = to_c0 ? d.W : d.WC; q.N = to_c0 ? d.W : d.WC; q.m_ptr = S_ptr; q.m_max = Sm;
        q.epi = to_c0 ? EPI_NONE : EPI_MUL_DLEAKY; q.aux = Clast; q.ldaux = Clast_ld; q.slope = d.slope;
        if (!to_c0 && use_sign_masks(d)) { q.mask_in = ws.CM[P.n_color_hidden - 1]; q.ldmask = d.WC / 32; }
        if (!to_c0 && d_biases) q.colsum = d_biases[P.color_layer0 + P.n_color_hidden - 1];     // bias gradient of the layer that produced Clast
        if ((rc = run_gemm_nn(q, tc, st))) return rc;
        dcur_ld = q.ldc;
    }
    for (int c = P.n_color_hidden - 1; c >= 0; c--) {
        const int l = P.color_layer0 + c;
        const float* a_in = c > 0 ? ws.CH[c - 1] : ws.C0;
        const int a_ld = c > 0 ? d.WC : d.kc0pad, a_k = c > 0 ? d.WC : d.kc0;
        if ((rc = wgrad(dcur, dcur_ld, d.WC, a_in, a_ld, a_k, d_weights ? d_weights[l] : nullptr, P.layers[l].in, S_ptr, Sm))) return rc;
        float* dnext = ws.dC[(P.n_color_hidden - c) & 1];
        GemmNN q = {};
        q.A1 = dcur; q.lda1 = dcur_ld; q.B1 = ws.Wp[l]; q.ldb1 = P.layers[l].kpad; q.K1 = d.WC;
        q.Bt1 = ws.Wt[l]; q.ldbt1 = P.layers[l].npad;
        q.C = dnext; q.m_ptr = S_ptr; q.m_max = Sm; q.slope = d.slope;
        if (c > 0) {
            q.ldc = d.WC; q.N = d.WC; q.epi = EPI_MUL_DLEAKY; q.aux = a_in; q.ldaux = a_ld; q.colsum = d_biases ? d_biases[l - 1] : nullptr;
            if (use_sign_masks(d)) { q.mask_in = ws.CM[c - 1]; q.ldmask = d.WC / 32; }
        }
        else { q.ldc = d.W; q.N = d.W; q.epi = EPI_NONE; }      // only dF = dC0[:, :W] is needed (view encoding has no gradient)
        if ((rc = run_gemm_nn(q, tc, st))) return rc;
        dcur = dnext; dcur_ld = q.ldc;
    }
    const float* dF = dcur;   // [S_v, W]

    // ---- ksum + alpha ----
    const int nt = P.n_tuple_layers;
    const float* Hlast = ws.H[nt - 1];
    const int la = P.alpha_layer;
    float* dZ = ws.dZ[0];
    // also: weight + bias gradient of the alpha head and the bias gradient of the last tuple layer (reductions over the same rows)
    if (d.W <= 256)
        launch(agg_ksum_bwd_kernel<8>, item_grid(agg_ksum_bwd_kernel<8>, Tm), 256, 0, st, in, d, K, T_ptr, Tm, ws.tuple_src, ws.tuple_cs, ws.wc, ws.weight_n, Hlast, ws.araw,
               weights[la], dF, d.W, d_decoded, dZ, ws.d_araw, g.conf, d_weights ? d_weights[la] : nullptr, d_biases ? d_biases[la] : nullptr,
               d_biases ? d_biases[nt - 1] : nullptr);
    else
        launch(agg_ksum_bwd_kernel<AGG_MAX_W / 32>, item_grid(Tm), 256, 0, st, in, d, K, T_ptr, Tm, ws.tuple_src, ws.tuple_cs, ws.wc, ws.weight_n, Hlast,
               ws.araw, weights[la], dF, d.W, d_decoded, dZ, ws.d_araw, g.conf, d_weights ? d_weights[la] : nullptr, d_biases ? d_biases[la] : nullptr,
               d_biases ? d_biases[nt - 1] : nullptr);
    SGN_LAUNCH_CHECK();
    if (d_conf_coef && g.conf) {
        launch(agg_conf_out_bwd_kernel, cdiv(S * K, 256), 256, 0, st, pidx, S * K, d_conf_coef, g.conf);
        SGN_LAUNCH_CHECK();
    }

    // ---- per-tuple layers, last to first ----
    const float* dE7 = nullptr;
    for (int t = nt - 1; t >= 0; t--) {
        const LayerInfo& L = P.layers[t];
        const float* a_in = t > 0 ? ws.H[t - 1] : ws.X0;
        const int a_ld = t > 0 ? d.W : d.k0pad, a_k = t > 0 ? d.W : d.k0, a_kpad = t > 0 ? d.W : d.k0pad;
        if ((rc = wgrad(dZ, d.W, d.W, a_in, a_ld, a_k, d_weights ? d_weights[t] : nullptr, L.in, T_ptr, Tm))) return rc;
        if (L.extra == EXTRA_LABEL && d_weights)
            if ((rc = wgrad(dZ, d.W, d.W, ws.L, d.LD, d.LD, d_weights[t] + a_k, L.in, T_ptr, Tm))) return rc;
        if (L.extra == EXTRA_COLORDIR && d_weights)
            if ((rc = wgrad(dZ, d.W, d.W, ws.E7, 8, 7, d_weights[t] + a_k, L.in, T_ptr, Tm))) return rc;
        // bias gradients: the last layer's comes out of agg_ksum_bwd_kernel, the lower layers' out of the dgrad epilogues below (colsum)
        if (L.extra == EXTRA_COLORDIR && (g.color || g.dir)) {
            GemmNN q = {};
            q.A1 = dZ; q.lda1 = d.W; q.B1 = ws.Wp[t] + a_kpad; q.ldb1 = L.kpad; q.K1 = d.W;
            q.Bt1 = ws.Wt[t] + (size_t)a_kpad * L.npad; q.ldbt1 = L.npad;
            q.C = ws.dE7; q.ldc = 8; q.N = 8; q.m_ptr = T_ptr; q.m_max = Tm; q.epi = EPI_NONE;
            if ((rc = run_gemm_nn(q, tc, st))) return rc;
            dE7 = ws.dE7;
        }
        if (t > 0 || g.embedding) {
            float* dnext = t > 0 ? ws.dZ[(nt - t) & 1] : ws.dX0;
            GemmNN q = {};
            q.A1 = dZ; q.lda1 = d.W; q.B1 = ws.Wp[t]; q.ldb1 = L.kpad; q.K1 = d.W;
            q.Bt1 = ws.Wt[t]; q.ldbt1 = L.npad;
            q.C = dnext; q.ldc = a_kpad; q.N = a_kpad; q.m_ptr = T_ptr; q.m_max = Tm; q.slope = d.slope;
            if (t > 0) {
                q.epi = EPI_MUL_DLEAKY; q.aux = a_in; q.ldaux = a_ld; q.colsum = d_biases ? d_biases[t - 1] : nullptr;
                if (use_sign_masks(d)) { q.mask_in = ws.HM[t - 1]; q.ldmask = d.W / 32; }
            } else { q.epi = EPI_NONE; }
            if ((rc = run_gemm_nn(q, tc, st))) return rc;
            dZ = dnext;
        }
    }
    if (g.embedding || g.color || g.dir) {
        launch(agg_scatter_kernel, item_grid(agg_scatter_kernel, Tm, (size_t)16 * d.k0pad * sizeof(float)), 256, (size_t)16 * d.k0pad * sizeof(float), st, in, d, K, SR, T_ptr, Tm, ws.tuple_src, ws.tuple_pt, ws.X0, ws.dX0, dE7, g);
        SGN_LAUNCH_CHECK();
    }
    return SGN_OK;
}
