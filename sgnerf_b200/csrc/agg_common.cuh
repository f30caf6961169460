// agg_common.cuh -- layer plan and workspace layout shared by the aggregator paths.
#pragma once
#include "common.cuh"

namespace sgn {

constexpr int AGG_MAX_LAYERS = 24;
constexpr int64_t AGG_FP32_CHUNK = 4096;  // rays per pass of the fp32 inference path (bounds the workspace)

enum ExtraInput { EXTRA_NONE = 0, EXTRA_LABEL = 1, EXTRA_COLORDIR = 2 };

constexpr int AGG_MAX_W = 512;        // widest MLP (shading_feature_num) the warp-per-row kernels keep per-lane partial sums for

struct LayerInfo {
    int in, out;      // torch Linear shape [out, in]
    int kpad, npad;   // padded to multiples of 8
    int extra;        // what is concatenated after the `width` running features (per-tuple layers only)
};

struct AggDims {
    int C, F, FD, FV;       // feat_dim, num_feat_freqs, dist_xyz_freq, num_viewdir_freqs
    int W, WC, LD;          // shading_feature_num, colour hidden width (W/2), label embedding dim
    int k0, k0pad;          // per-tuple input width (284) and its padding (288)
    int kc0, kc0pad;        // colour-branch input width (280)
    int act_super;
    float slope;
};

struct AggPlan {
    AggDims dims;
    int n_layers;           // all Linear layers, state_dict order
    int n_tuple_layers;     // block1 + block2_bpnet + block3
    int alpha_layer;        // index of alpha_branch.0
    int color_layer0;       // index of color_branch.0
    int n_color_hidden;     // colour layers followed by an activation
    LayerInfo layers[AGG_MAX_LAYERS];
};

static inline int pad8(int x) { return (x + 7) / 8 * 8; }

static inline int make_plan(const SgnAggCfg* c, AggPlan* P)
{
    SGN_CHECK_ARG(c != nullptr, "aggregator: cfg is NULL");
    SGN_CHECK_ARG(c->feat_dim > 0 && c->feat_dim % 4 == 0, "aggregator: feat_dim must be a positive multiple of 4");
    SGN_CHECK_ARG(c->num_feat_freqs >= 0 && c->dist_xyz_freq > 0 && c->num_viewdir_freqs > 0, "aggregator: bad frequency counts");
    SGN_CHECK_ARG(c->width >= 16 && c->width % 16 == 0 && c->width <= AGG_MAX_W, "aggregator: width must be a multiple of 16, at most %d", AGG_MAX_W);
    SGN_CHECK_ARG(c->n_block1 >= 1 && c->n_block3 >= 1, "aggregator: block1 and block3 need at least one layer (canonical branch)");
    SGN_CHECK_ARG(c->n_block2_bpnet >= 0 && c->n_color >= 1, "aggregator: bad layer counts");
    SGN_CHECK_ARG(c->label_dim % 8 == 0 && c->label_dim >= 0, "aggregator: label_dim must be a multiple of 8");
    SGN_CHECK_ARG(c->label_dim == 0 || c->n_block2_bpnet > 0, "aggregator: label embedding needs block2_bpnet");
    AggDims& d = P->dims;
    d.C = c->feat_dim; d.F = c->num_feat_freqs; d.FD = c->dist_xyz_freq; d.FV = c->num_viewdir_freqs;
    d.W = c->width; d.WC = c->width / 2; d.LD = c->label_dim;
    d.k0 = d.C * (1 + 2 * d.F) + 2 * d.FD * 6; d.k0pad = pad8(d.k0);
    d.kc0 = d.W + 6 * d.FV; d.kc0pad = pad8(d.kc0);
    d.act_super = c->act_super; d.slope = c->leaky_slope;
    int n = 0;
    auto add = [&](int in, int out, int extra) {
        LayerInfo& L = P->layers[n++];
        L.in = in; L.out = out; L.kpad = pad8(in); L.npad = pad8(out); L.extra = extra;
    };
    int in = d.k0;
    SGN_CHECK_ARG(c->n_block1 + c->n_block2_bpnet + c->n_block3 + 1 + c->n_color <= AGG_MAX_LAYERS, "aggregator: too many layers");
    for (int i = 0; i < c->n_block1; i++) { add(in, d.W, EXTRA_NONE); in = d.W; }
    for (int i = 0; i < c->n_block2_bpnet; i++) { add(in + (i == 0 ? d.LD : 0), d.W, (i == 0 && d.LD > 0) ? EXTRA_LABEL : EXTRA_NONE); in = d.W; }
    for (int i = 0; i < c->n_block3; i++) { add(in + (i == 0 ? 7 : 0), d.W, i == 0 ? EXTRA_COLORDIR : EXTRA_NONE); in = d.W; }
    P->n_tuple_layers = n;
    P->alpha_layer = n;
    add(d.W, 1, EXTRA_NONE);
    P->color_layer0 = n;
    P->n_color_hidden = c->n_color - 1;
    int cin = d.kc0;
    for (int i = 0; i < c->n_color - 1; i++) { add(cin, d.WC, EXTRA_NONE); cin = d.WC; }
    add(cin, 3, EXTRA_NONE);
    P->n_layers = n;
    return SGN_OK;
}

// Device-side view of the inputs of one aggregation call (or one ray chunk of it).
struct AggIn {
    SgnPointTables tab;
    const int32_t* pidx;    // [R,SR,K]
    const float* loc_w;     // [R,SR,3]
    const float* raydir;    // [R,3]
    const float* campos;    // [3]
    const float* camrot;    // [3,3] camrotc2w, row-major
    const int32_t* smask;   // optional [R,SR] (sgn_query's sample_mask): 0 = the slot holds no sample, its sample_pidx row is all -1 and is not read
};

// the sample mask handed to sgn_agg_forward_frame_masked for the call in progress on this thread (the inner forward functions pick it up)
extern thread_local const int32_t* g_agg_sample_mask;

// fp32-path workspace.  Sized for the worst case (every slot valid) so nothing depends on device counts.
struct AggWs {
    int32_t *nvalid, *svalid, *tuple_start, *sample_cidx, *partials, *tuple_src, *csample;
    int32_t *tuple_pt, *tuple_cs;   // per compact tuple: its point, its compact sample
    float *loc_pers, *weight_n, *wc;
    float *X0, *L, *E7, *araw, *C0, *sigma, *sig;
    float* H[AGG_MAX_LAYERS];
    uint32_t* HM[AGG_MAX_LAYERS];   // sign bits of H[t] / CH[c] (training, tensor-core path): what the dgrad epilogues read instead of the activations
    uint32_t* CM[AGG_MAX_LAYERS];
    float* CH[AGG_MAX_LAYERS];
    float* Wt[AGG_MAX_LAYERS];
    float* Wp[AGG_MAX_LAYERS];
    // backward scratch (save mode only)
    float *dZ[2], *dX0, *dE7, *d_araw, *dC[2], *d_raw;
};

static inline size_t carve_ws(const AggPlan& P, int64_t Rc, int SR, int K, bool save, void* base, size_t cap, AggWs* ws)
{
    const AggDims& d = P.dims;
    Arena A(base, cap);
    const size_t S = (size_t)Rc * SR, T = S * K;
    ws->nvalid = A.take<int32_t>(S + 1);
    ws->svalid = A.take<int32_t>(S + 1);
    ws->tuple_start = A.take<int32_t>(S + 1);
    ws->sample_cidx = A.take<int32_t>(S + 1);
    ws->partials = A.take<int32_t>(2 * scan_partials_count((int64_t)S));
    ws->tuple_src = A.take<int32_t>(T + 1);
    ws->tuple_pt = A.take<int32_t>(T + 1);
    ws->tuple_cs = A.take<int32_t>(T + 1);
    ws->csample = A.take<int32_t>(S + 1);
    ws->loc_pers = A.take<float>(S * 3);
    ws->weight_n = A.take<float>(T);
    ws->wc = A.take<float>(T);
    ws->X0 = A.take<float>(T * d.k0pad);
    ws->L = A.take<float>(T * (d.LD > 0 ? d.LD : 1));
    ws->E7 = A.take<float>(T * 8);
    ws->araw = A.take<float>(T);
    ws->C0 = A.take<float>(S * d.kc0pad);
    ws->sigma = A.take<float>(S);
    ws->sig = A.take<float>(S * 4);
    const int nH = save ? P.n_tuple_layers : 2, nC = save ? (P.n_color_hidden > 0 ? P.n_color_hidden : 1) : 2;
    for (int i = 0; i < nH; i++) ws->H[i] = A.take<float>(T * d.W);
    for (int i = 0; i < nC; i++) ws->CH[i] = A.take<float>(S * d.WC);
    for (int l = 0; l < P.n_layers; l++) {
        ws->Wt[l] = A.take<float>((size_t)P.layers[l].kpad * P.layers[l].npad);
        ws->Wp[l] = A.take<float>((size_t)P.layers[l].kpad * P.layers[l].npad);
    }
    if (save) {
        for (int i = 0; i < nH; i++) ws->HM[i] = A.take<uint32_t>(T * (size_t)((d.W + 31) / 32));
        for (int i = 0; i < nC; i++) ws->CM[i] = A.take<uint32_t>(S * (size_t)((d.WC + 31) / 32));
        ws->dZ[0] = A.take<float>(T * d.W);
        ws->dZ[1] = A.take<float>(T * d.W);
        ws->dX0 = A.take<float>(T * d.k0pad);
        ws->dE7 = A.take<float>(T * 8);
        ws->d_araw = A.take<float>(T);
        ws->dC[0] = A.take<float>(S * d.W);
        ws->dC[1] = A.take<float>(S * d.W);
        ws->d_raw = A.take<float>(S * 8);
    }
    return A.off;
}

}  // namespace sgn
