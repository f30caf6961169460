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
//   layers   SGEMM (gemm_simt.cuh) with fused bias + LeakyReLU, concat inputs as a second operand pair
//   alpha    per tuple  : raw sigma ; ksum per sample : sum_k w conf (sigma, h) ; colour MLP ; rgb
// The tensor-core (bf16, tcgen05) path lives in agg_tc.cu; this file is the numerically strict one and the
// one training uses.
#include "agg_common.cuh"
#include "gemm_simt.cuh"

namespace sgn {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// One thread per sample.  point_aggregators.py:885, :917-925 (dists), :494-502 + :946-947 (weights), :953 (conf).
__global__ void agg_prepare_kernel(AggIn in, int64_t S, int K, float* __restrict__ loc_pers, float* __restrict__ wc,
                                   float* __restrict__ weight_n, float* __restrict__ weight_out, float* __restrict__ conf_out, uint8_t* __restrict__ ray_valid,
                                   int32_t* __restrict__ nvalid, int32_t* __restrict__ svalid)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float lx = in.loc_w[3 * s], ly = in.loc_w[3 * s + 1], lz = in.loc_w[3 * s + 2];
    {   // w2pers: (p - campos) @ camrotc2w, then (x/z, y/z, z)
        const float sx = lx - in.campos[0], sy = ly - in.campos[1], sz = lz - in.campos[2];
        const float* Rm = in.camrot;
        const float c0 = sx * Rm[0] + sy * Rm[3] + sz * Rm[6];
        const float c1 = sx * Rm[1] + sy * Rm[4] + sz * Rm[7];
        const float c2 = sx * Rm[2] + sy * Rm[5] + sz * Rm[8];
        loc_pers[3 * s] = c0 / c2; loc_pers[3 * s + 1] = c1 / c2; loc_pers[3 * s + 2] = c2;
    }
    const int32_t* pi = in.pidx + s * K;
    float w[SGN_MAX_K];
    float sum = 0.f;
    int n = 0;
    for (int k = 0; k < K; k++) {
        const int p = pi[k];
        float wk = 0.f;
        if (p >= 0) {
            const float dx = in.tab.xyz[3 * (int64_t)p] - lx, dy = in.tab.xyz[3 * (int64_t)p + 1] - ly, dz = in.tab.xyz[3 * (int64_t)p + 2] - lz;
            const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
            wk = 1.0f / fmaxf(nrm, 1e-6f);
            n++;
        }
        w[k] = wk;
        sum += wk;
    }
    const float den = fmaxf(sum, 1e-8f);
    for (int k = 0; k < K; k++) {
        const int p = pi[k] < 0 ? 0 : pi[k];           // the reference gathers with clamp(pidx, 0) (neural_points.py:958)
        const float cf = in.tab.conf ? fminf(fmaxf(in.tab.conf[p], 0.0001f), 1.0f) : 1.0f;
        const float wn = w[k] / den;
        wc[s * K + k] = wn * cf;
        weight_n[s * K + k] = wn;
        if (weight_out) weight_out[s * K + k] = wn;
        if (conf_out) conf_out[s * K + k] = cf;
    }
    ray_valid[s] = n > 0;
    nvalid[s] = n;
    svalid[s] = n > 0;
}

// One thread per sample: tuple j -> (sample, slot), compact sample c -> sample.
__global__ void agg_index_kernel(const int32_t* __restrict__ pidx, int64_t S, int K, const int32_t* __restrict__ tuple_start,
                                 const int32_t* __restrict__ sample_cidx, const int32_t* __restrict__ nvalid,
                                 int32_t* __restrict__ tuple_src, int32_t* __restrict__ csample)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (nvalid[s] == 0) return;
    csample[sample_cidx[s]] = (int32_t)s;
    int j = tuple_start[s];
    for (int k = 0; k < K; k++)
        if (pidx[s * K + k] >= 0) tuple_src[j++] = (int32_t)(s * K + k);
}

// One warp per tuple.  X0 layout = reference `feat` (:603-611):
//   [0,C) embedding | C + 2*(d*F+f) + {0:sin,1:cos} of emb_d * 2^f | then the same for the 6 dists with F = dist_xyz_freq.
// E7 = [colour(3) | dir - viewdir (3) | dir . viewdir | 0]  (:639-652).
__global__ void __launch_bounds__(256)
agg_gather_kernel(AggIn in, AggDims d, int K, int SR, const int32_t* __restrict__ T_ptr, int T_max, const int32_t* __restrict__ tuple_src,
                  const float* __restrict__ loc_pers, float* __restrict__ X0, float* __restrict__ L, float* __restrict__ E7)
{
    const int lane = lane_id();
    const int T = min(*T_ptr, T_max);
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= T) return;
    const int flat = tuple_src[j];
    const int64_t s = flat / K;
    const int64_t r = s / SR;
    const int64_t p = in.pidx[flat];
    float* x = X0 + j * d.k0pad;
    const int C = d.C, F = d.F, FD = d.FD;
    for (int c = lane; c < C; c += 32) {
        const float e = __ldg(in.tab.embedding + p * C + c);
        x[c] = e;
        float fr = 1.0f;
        for (int f = 0; f < F; f++) {
            const float a = e * fr;
            x[C + 2 * (c * F + f)] = sinf(a);
            x[C + 2 * (c * F + f) + 1] = cosf(a);
            fr *= 2.0f;
        }
    }
    const int base = C + 2 * C * F;
    if (lane < 6) {
        const float px = in.tab.xyz[3 * p], py = in.tab.xyz[3 * p + 1], pz = in.tab.xyz[3 * p + 2];
        float dist;
        if (lane < 3) {
            dist = in.tab.xyz[3 * p + lane] - in.loc_w[3 * s + lane];
        } else {
            // point in perspective coords (neural_points.py:845-850), then :920-922
            const float sx = px - in.campos[0], sy = py - in.campos[1], sz = pz - in.campos[2];
            const float* Rm = in.camrot;
            const float c0 = sx * Rm[0] + sy * Rm[3] + sz * Rm[6];
            const float c1 = sx * Rm[1] + sy * Rm[4] + sz * Rm[7];
            const float c2 = sx * Rm[2] + sy * Rm[5] + sz * Rm[8];
            const float xp = c0 / c2, yp = c1 / c2, zp = c2;
            const float lxp = loc_pers[3 * s], lyp = loc_pers[3 * s + 1], lzp = loc_pers[3 * s + 2];
            dist = lane == 3 ? xp * zp - lxp * lzp : (lane == 4 ? yp * zp - lyp * lzp : zp - lzp);
        }
        float fr = 1.0f;
        for (int f = 0; f < FD; f++) {
            const float a = dist * fr;
            x[base + 2 * (lane * FD + f)] = sinf(a);
            x[base + 2 * (lane * FD + f) + 1] = cosf(a);
            fr *= 2.0f;
        }
    }
    for (int c = d.k0 + lane; c < d.k0pad; c += 32) x[c] = 0.f;
    if (L) {
        for (int c = lane; c < d.LD; c += 32) L[j * d.LD + c] = __ldg(in.tab.label_emb + p * d.LD + c);
    }
    if (lane < 8) {
        float v = 0.f;
        const float vx = in.raydir[3 * r], vy = in.raydir[3 * r + 1], vz = in.raydir[3 * r + 2];
        const float dx = in.tab.dir[3 * p], dy = in.tab.dir[3 * p + 1], dz = in.tab.dir[3 * p + 2];
        if (lane < 3) v = in.tab.color[3 * p + lane];
        else if (lane == 3) v = dx - vx;
        else if (lane == 4) v = dy - vy;
        else if (lane == 5) v = dz - vz;
        else if (lane == 6) v = dx * vx + dy * vy + dz * vz;
        E7[j * 8 + lane] = v;
    }
}

// One warp per tuple: raw alpha = h . wa + ba   (alpha_branch, a single Linear)
__global__ void __launch_bounds__(256)
agg_alpha_kernel(const float* __restrict__ H, int W, const float* __restrict__ wa, const float* __restrict__ ba,
                 const int32_t* __restrict__ T_ptr, int T_max, float* __restrict__ araw)
{
    const int lane = lane_id();
    const int T = min(*T_ptr, T_max);
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= T) return;
    float acc = 0.f;
    for (int c = lane; c < W; c += 32) acc = fmaf(H[j * W + c], __ldg(wa + c), acc);
    acc = warp_sum(acc);
    if (lane == 0) araw[j] = acc + ba[0];
}

__device__ __forceinline__ float softplus1(float x)  // torch.nn.Softplus(beta=1, threshold=20)
{
    return x > 20.0f ? x : log1pf(expf(x));
}

// One warp per compact sample: sigma = sum_k wc * softplus(raw - 1), F = sum_k wc * h  -> C0[:, :W];
// C0[:, W:W+6*FV] = viewdir encoding (ori=True, first three stripped): sin(v_d 2^f) d-major, then cos (:579-585).
__global__ void __launch_bounds__(256)
agg_ksum_kernel(AggIn in, AggDims d, int K, int SR, const int32_t* __restrict__ S_ptr, int S_max, const int32_t* __restrict__ csample,
                const int32_t* __restrict__ tuple_start, const int32_t* __restrict__ nvalid, const float* __restrict__ wc,
                const float* __restrict__ H, const float* __restrict__ araw, float* __restrict__ C0, float* __restrict__ sigma)
{
    const int lane = lane_id();
    const int Sv = min(*S_ptr, S_max);
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= Sv) return;
    const int64_t s = csample[c];
    const int j0 = tuple_start[s], n = nvalid[s];
    const int W = d.W;
    float wk[SGN_MAX_K];
    {
        int q = 0;
        for (int k = 0; k < K; k++)
            if (in.pidx[s * K + k] >= 0) wk[q++] = wc[s * K + k];
    }
    float sg = 0.f;
    for (int q = 0; q < n; q++) {
        const float a = araw[j0 + q];
        sg += wk[q] * (d.act_super ? softplus1(a - 1.0f) : fmaxf(a, 0.f));
    }
    float* row = C0 + c * d.kc0pad;
    for (int col = lane; col < W; col += 32) {
        float acc = 0.f;
        for (int q = 0; q < n; q++) acc += H[(int64_t)(j0 + q) * W + col] * wk[q];
        row[col] = acc;
    }
    const int64_t r = s / SR;
    const int FV = d.FV;
    for (int i = lane; i < 3 * FV; i += 32) {
        const int dd = i / FV, f = i - dd * FV;
        const float a = in.raydir[3 * r + dd] * exp2f((float)f);
        row[W + i] = sinf(a);
        row[W + 3 * FV + i] = cosf(a);
    }
    for (int col = W + 6 * FV + lane; col < d.kc0pad; col += 32) row[col] = 0.f;
    if (lane == 0) sigma[c] = sg;
}

// One warp per compact sample: rgb = sigmoid(c . Wlast^T + b) (*1.002 - 0.001), decoded[s] = (sigma, rgb)
__global__ void __launch_bounds__(256)
agg_rgb_kernel(AggDims d, const int32_t* __restrict__ S_ptr, int S_max, const int32_t* __restrict__ csample,
               const float* __restrict__ Cin, int Wc, const float* __restrict__ Wl, const float* __restrict__ bl,
               const float* __restrict__ sigma, float* __restrict__ decoded, float* __restrict__ sig_out)
{
    const int lane = lane_id();
    const int Sv = min(*S_ptr, S_max);
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= Sv) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int col = lane; col < Wc; col += 32) {
        const float x = Cin[c * Wc + col];
        a0 = fmaf(x, __ldg(Wl + col), a0);
        a1 = fmaf(x, __ldg(Wl + Wc + col), a1);
        a2 = fmaf(x, __ldg(Wl + 2 * Wc + col), a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) {
        const float s0 = 1.0f / (1.0f + expf(-(a0 + bl[0]))), s1 = 1.0f / (1.0f + expf(-(a1 + bl[1]))), s2 = 1.0f / (1.0f + expf(-(a2 + bl[2])));
        const float m = d.act_super ? 1.002f : 1.0f, o = d.act_super ? 0.001f : 0.0f;
        const int64_t s = csample[c];
        ((float4*)decoded)[s] = make_float4(sigma[c], s0 * m - o, s1 * m - o, s2 * m - o);
        if (sig_out) ((float4*)sig_out)[c] = make_float4(s0, s1, s2, 0.f);
    }
}

// W [N,K] (torch Linear) -> Wt [Kpad, Npad] (k-major, forward B operand) and Wp [Npad, Kpad] (dgrad B operand)
__global__ void pack_weight_kernel(const float* __restrict__ W, int N, int Kin, int Npad, int Kpad, float* __restrict__ Wt, float* __restrict__ Wp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Npad * Kpad) return;
    const int n = i / Kpad, k = i - n * Kpad;
    const float v = (n < N && k < Kin) ? W[(size_t)n * Kin + k] : 0.f;
    Wp[(size_t)n * Kpad + k] = v;
    Wt[(size_t)k * Npad + n] = v;
}

// ---- backward-only kernels ----

// column sums: out[c] += sum_m X[m, c]  (bias gradients)
__global__ void __launch_bounds__(128)
colsum_kernel(const float* __restrict__ X, int ld, int ncols, const int32_t* __restrict__ m_ptr, int m_max, int rows_per_block, float* __restrict__ out)
{
    const int M = min(*m_ptr, m_max);
    const int c = blockIdx.y * 128 + threadIdx.x;
    const int m0 = blockIdx.x * rows_per_block;
    if (m0 >= M || c >= ncols) return;
    const int m1 = min(M, m0 + rows_per_block);
    float acc = 0.f;
    for (int m = m0; m < m1; m++) acc += X[(size_t)m * ld + c];
    atomicAdd(out + c, acc);
}

// One warp per compact sample: d_raw[c, 0:3] = d_rgb * scale * sig (1 - sig), padded to 8 columns
__global__ void __launch_bounds__(256)
agg_rgb_bwd_kernel(AggDims d, const int32_t* __restrict__ S_ptr, int S_max, const int32_t* __restrict__ csample,
                   const float* __restrict__ d_decoded, const float* __restrict__ sig, float* __restrict__ d_raw)
{
    const int Sv = min(*S_ptr, S_max);
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Sv) return;
    const float4 g = ((const float4*)d_decoded)[csample[c]];
    const float4 sg = ((const float4*)sig)[c];
    const float m = d.act_super ? 1.002f : 1.0f;
    float* o = d_raw + c * 8;
    o[0] = g.y * m * sg.x * (1.f - sg.x);
    o[1] = g.z * m * sg.y * (1.f - sg.y);
    o[2] = g.w * m * sg.z * (1.f - sg.z);
    o[3] = o[4] = o[5] = o[6] = o[7] = 0.f;
}

// One warp per tuple: backward of ksum + alpha.  Writes dZ = dH (.) leaky'(H) for the last tuple layer,
// d_araw[j], and accumulates d_conf (straight-through clamp: d conf_coef / d conf = 1).
__global__ void __launch_bounds__(256)
agg_ksum_bwd_kernel(AggIn in, AggDims d, int K, const int32_t* __restrict__ T_ptr, int T_max, const int32_t* __restrict__ tuple_src,
                    const int32_t* __restrict__ sample_cidx, const float* __restrict__ wc, const float* __restrict__ weight_n,
                    const float* __restrict__ H, const float* __restrict__ araw, const float* __restrict__ wa,
                    const float* __restrict__ dC0, int lddc0, const float* __restrict__ d_decoded, float* __restrict__ dZ,
                    float* __restrict__ d_araw, float* __restrict__ d_conf)
{
    const int lane = lane_id();
    const int T = min(*T_ptr, T_max);
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= T) return;
    const int flat = tuple_src[j];
    const int64_t s = flat / K;
    const int64_t c = sample_cidx[s];
    const float w = wc[flat];
    const float dsig = d_decoded[4 * s];
    const float a = araw[j];
    float act, dact;
    if (d.act_super) {
        const float x = a - 1.0f;
        act = softplus1(x);
        dact = x > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-x));
    } else {
        act = fmaxf(a, 0.f);
        dact = a > 0.f ? 1.0f : 0.f;
    }
    const float da = w * dsig * dact;
    const int W = d.W;
    float dot = 0.f;
    for (int col = lane; col < W; col += 32) {
        const float h = H[j * W + col];
        const float df = dC0[c * lddc0 + col];
        dot = fmaf(h, df, dot);
        const float dh = w * df + da * __ldg(wa + col);
        dZ[j * W + col] = dh * (h > 0.f ? 1.0f : d.slope);
    }
    dot = warp_sum(dot);
    if (lane == 0) {
        d_araw[j] = da;
        if (d_conf) {
            const float d_wc = act * dsig + dot;
            atomicAdd(d_conf + in.pidx[flat], weight_n[flat] * d_wc);
        }
    }
}

// cotangent of the conf_coefficient output (all slots, invalid ones use point 0 like the reference's clamp(pidx,0) gather)
__global__ void agg_conf_out_bwd_kernel(const int32_t* __restrict__ pidx, int64_t n, const float* __restrict__ d_conf_coef, float* __restrict__ d_conf)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = d_conf_coef[i];
    if (g != 0.f) atomicAdd(d_conf + (pidx[i] < 0 ? 0 : pidx[i]), g);
}

// One warp per tuple: scatter-add into the point tables.
//   d emb_c = dX0[c] + sum_f 2^f (cos_cf dsin_cf - sin_cf dcos_cf)   (sin/cos read back from the saved X0)
//   d colour = dE7[0:3] ; d dir = dE7[3:6] + viewdir * dE7[6]
__global__ void __launch_bounds__(256)
agg_scatter_kernel(AggIn in, AggDims d, int K, int SR, const int32_t* __restrict__ T_ptr, int T_max, const int32_t* __restrict__ tuple_src,
                   const float* __restrict__ X0, const float* __restrict__ dX0, const float* __restrict__ dE7, SgnPointGrads g)
{
    const int lane = lane_id();
    const int T = min(*T_ptr, T_max);
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= T) return;
    const int flat = tuple_src[j];
    const int64_t p = in.pidx[flat];
    const int C = d.C, F = d.F;
    if (g.embedding) {
        const float* x = X0 + j * d.k0pad;
        const float* dx = dX0 + j * d.k0pad;
        for (int c = lane; c < C; c += 32) {
            float acc = dx[c], fr = 1.0f;
            for (int f = 0; f < F; f++) {
                const int o = C + 2 * (c * F + f);
                acc += fr * (x[o + 1] * dx[o] - x[o] * dx[o + 1]);
                fr *= 2.0f;
            }
            atomicAdd(g.embedding + p * C + c, acc);
        }
    }
    if (dE7) {
        const float* e = dE7 + j * 8;
        if (g.color && lane < 3) atomicAdd(g.color + 3 * p + lane, e[lane]);
        if (g.dir && lane >= 3 && lane < 6) {
            const int64_t r = (flat / K) / SR;
            atomicAdd(g.dir + 3 * p + (lane - 3), e[lane] + in.raydir[3 * r + (lane - 3)] * e[6]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------

static int pack_weights(const AggPlan& P, const float* const* weights, AggWs& ws, cudaStream_t st)
{
    for (int l = 0; l < P.n_layers; l++) {
        const LayerInfo& L = P.layers[l];
        const int n = L.npad * L.kpad;
        launch(pack_weight_kernel, cdiv(n, 256), 256, 0, st, weights[l], L.out, L.in, L.npad, L.kpad, ws.Wt[l], ws.Wp[l]);
    }
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

// forward for one chunk of rays
static int agg_forward_chunk(const AggPlan& P, const float* const* weights, const float* const* biases, const AggIn& in, int64_t Rc,
                             int SR, int K, float* decoded, uint8_t* ray_valid, float* loc_pers_out, float* weight_out, float* conf_out,
                             AggWs& ws, bool save, cudaStream_t st)
{
    const AggDims& d = P.dims;
    const int64_t S = Rc * SR;
    const int Tm = (int)(S * K), Sm = (int)S;
    float* loc_pers = loc_pers_out ? loc_pers_out : ws.loc_pers;
    SGN_CUDA(cudaMemsetAsync(decoded, 0, sizeof(float) * 4 * (size_t)S, st));
    launch(agg_prepare_kernel, cdiv(S, 128), 128, 0, st, in, S, K, loc_pers, ws.wc, ws.weight_n, weight_out, conf_out, ray_valid, ws.nvalid, ws.svalid);
    int rc;
    if ((rc = exclusive_scan_i32(ws.nvalid, ws.tuple_start, S, ws.partials, st))) return rc;
    if ((rc = exclusive_scan_i32(ws.svalid, ws.sample_cidx, S, ws.partials, st))) return rc;
    const int32_t* T_ptr = ws.tuple_start + S;
    const int32_t* S_ptr = ws.sample_cidx + S;
    launch(agg_index_kernel, cdiv(S, 128), 128, 0, st, in.pidx, S, K, ws.tuple_start, ws.sample_cidx, ws.nvalid, ws.tuple_src, ws.csample);
    launch(agg_gather_kernel, cdiv(Tm, 8), 256, 0, st, in, d, K, SR, T_ptr, Tm, ws.tuple_src, loc_pers, ws.X0, d.LD > 0 ? ws.L : nullptr, ws.E7);
    SGN_LAUNCH_CHECK();

    // per-tuple layers
    const float* cur = ws.X0;
    int cur_ld = d.k0pad, cur_k = d.k0pad;
    for (int t = 0; t < P.n_tuple_layers; t++) {
        const LayerInfo& L = P.layers[t];
        float* out = ws.H[save ? t : (t & 1)];
        GemmNN g = {};
        g.A1 = cur; g.lda1 = cur_ld; g.B1 = ws.Wt[t]; g.ldb1 = L.npad; g.K1 = cur_k;
        if (L.extra == EXTRA_LABEL) { g.A2 = ws.L; g.lda2 = d.LD; g.B2 = ws.Wt[t] + (size_t)cur_k * L.npad; g.ldb2 = L.npad; g.K2 = d.LD; }
        if (L.extra == EXTRA_COLORDIR) { g.A2 = ws.E7; g.lda2 = 8; g.B2 = ws.Wt[t] + (size_t)cur_k * L.npad; g.ldb2 = L.npad; g.K2 = 8; }
        g.C = out; g.ldc = d.W; g.N = d.W; g.m_ptr = T_ptr; g.m_max = Tm; g.bias = biases[t]; g.epi = EPI_BIAS_LEAKY; g.slope = d.slope;
        if ((rc = launch_gemm_nn(g, st))) return rc;
        cur = out; cur_ld = d.W; cur_k = d.W;
    }
    const float* Hlast = cur;
    const int la = P.alpha_layer;
    launch(agg_alpha_kernel, cdiv(Tm, 8), 256, 0, st, Hlast, d.W, weights[la], biases[la], T_ptr, Tm, ws.araw);
    launch(agg_ksum_kernel, cdiv(Sm, 8), 256, 0, st, in, d, K, SR, S_ptr, Sm, ws.csample, ws.tuple_start, ws.nvalid, ws.wc, Hlast, ws.araw, ws.C0, ws.sigma);
    SGN_LAUNCH_CHECK();

    // colour MLP
    cur = ws.C0; cur_ld = d.kc0pad; cur_k = d.kc0pad;
    for (int c = 0; c < P.n_color_hidden; c++) {
        const int l = P.color_layer0 + c;
        const LayerInfo& L = P.layers[l];
        float* out = ws.CH[save ? c : (c & 1)];
        GemmNN g = {};
        g.A1 = cur; g.lda1 = cur_ld; g.B1 = ws.Wt[l]; g.ldb1 = L.npad; g.K1 = cur_k;
        g.C = out; g.ldc = d.WC; g.N = d.WC; g.m_ptr = S_ptr; g.m_max = Sm; g.bias = biases[l]; g.epi = EPI_BIAS_LEAKY; g.slope = d.slope;
        if ((rc = launch_gemm_nn(g, st))) return rc;
        cur = out; cur_ld = d.WC; cur_k = d.WC;
    }
    const int ll = P.n_layers - 1;
    launch(agg_rgb_kernel, cdiv(Sm, 8), 256, 0, st, d, S_ptr, Sm, ws.csample, cur, cur_ld, weights[ll], biases[ll], ws.sigma, decoded, ws.sig);
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
                         int64_t R, int SR, int K, int save, float* decoded, uint8_t* ray_valid, float* loc_pers, float* weight,
                         float* conf_coef, void* workspace, size_t workspace_bytes, cudaStream_t st)
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
        in.campos = campos; in.camrot = camrotc2w;
        rc = agg_forward_chunk(P, weights, biases, in, Rc, SR, K, decoded + r0 * SR * 4, ray_valid + r0 * SR,
                               loc_pers ? loc_pers + r0 * SR * 3 : nullptr, weight ? weight + r0 * SR * K : nullptr,
                               conf_coef ? conf_coef + r0 * SR * K : nullptr, ws, save != 0, st);
        if (rc) return rc;
    }
    return SGN_OK;
}

int sgn_agg_fp32_backward(const AggPlan& P, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                          const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                          int64_t R, int SR, int K, const float* d_decoded, const float* d_conf_coef, float* const* d_weights,
                          float* const* d_biases, const SgnPointGrads* d_tables, void* workspace, size_t workspace_bytes, cudaStream_t st)
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
    in.tab = *tables; in.pidx = pidx; in.loc_w = loc_w; in.raydir = raydir; in.campos = campos; in.camrot = camrotc2w;
    SgnPointGrads g = {};
    if (d_tables) g = *d_tables;
    int rc;
    auto bias_grad = [&](const float* dZ, int ld, int ncols, const int32_t* m_ptr, int m_max, float* out) -> int {
        if (!out || m_max <= 0) return SGN_OK;
        const int rpb = 512;
        dim3 grid(cdiv(m_max, rpb), cdiv(ncols, 128));
        launch(colsum_kernel, grid, 128, 0, st, dZ, ld, ncols, m_ptr, m_max, rpb, out);
        SGN_LAUNCH_CHECK();
        return SGN_OK;
    };
    auto wgrad = [&](const float* dZ, int ldz, int Pn, const float* act, int lda, int Q, float* out, int ldc, const int32_t* m_ptr, int m_max) -> int {
        if (!out) return SGN_OK;
        GemmTN t = {};
        t.A = dZ; t.lda = ldz; t.P = Pn; t.B = act; t.ldb = lda; t.Q = Q; t.C = out; t.ldc = ldc; t.m_ptr = m_ptr; t.m_max = m_max;
        return launch_gemm_tn(t, st);
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
        const bool to_c0 = P.n_color_hidden == 0;
        q.C = dcur; q.ldc = to_c0 ? d.W : d.WC; q.N = to_c0 ? d.W : d.WC; q.m_ptr = S_ptr; q.m_max = Sm;
        q.epi = to_c0 ? EPI_NONE : EPI_MUL_DLEAKY; q.aux = Clast; q.ldaux = Clast_ld; q.slope = d.slope;
        if ((rc = launch_gemm_nn(q, st))) return rc;
        dcur_ld = q.ldc;
    }
    for (int c = P.n_color_hidden - 1; c >= 0; c--) {
        const int l = P.color_layer0 + c;
        const float* a_in = c > 0 ? ws.CH[c - 1] : ws.C0;
        const int a_ld = c > 0 ? d.WC : d.kc0pad, a_k = c > 0 ? d.WC : d.kc0;
        if ((rc = wgrad(dcur, dcur_ld, d.WC, a_in, a_ld, a_k, d_weights ? d_weights[l] : nullptr, P.layers[l].in, S_ptr, Sm))) return rc;
        if ((rc = bias_grad(dcur, dcur_ld, d.WC, S_ptr, Sm, d_biases ? d_biases[l] : nullptr))) return rc;
        float* dnext = ws.dC[(P.n_color_hidden - c) & 1];
        GemmNN q = {};
        q.A1 = dcur; q.lda1 = dcur_ld; q.B1 = ws.Wp[l]; q.ldb1 = P.layers[l].kpad; q.K1 = d.WC;
        q.C = dnext; q.m_ptr = S_ptr; q.m_max = Sm; q.slope = d.slope;
        if (c > 0) { q.ldc = d.WC; q.N = d.WC; q.epi = EPI_MUL_DLEAKY; q.aux = a_in; q.ldaux = a_ld; }
        else { q.ldc = d.W; q.N = d.W; q.epi = EPI_NONE; }      // only dF = dC0[:, :W] is needed (view encoding has no gradient)
        if ((rc = launch_gemm_nn(q, st))) return rc;
        dcur = dnext; dcur_ld = q.ldc;
    }
    const float* dF = dcur;   // [S_v, W]

    // ---- ksum + alpha ----
    const int nt = P.n_tuple_layers;
    const float* Hlast = ws.H[nt - 1];
    const int la = P.alpha_layer;
    float* dZ = ws.dZ[0];
    launch(agg_ksum_bwd_kernel, cdiv(Tm, 8), 256, 0, st, in, d, K, T_ptr, Tm, ws.tuple_src, ws.sample_cidx, ws.wc, ws.weight_n, Hlast, ws.araw,
                                                    weights[la], dF, d.W, d_decoded, dZ, ws.d_araw, g.conf);
    SGN_LAUNCH_CHECK();
    if (d_conf_coef && g.conf) {
        launch(agg_conf_out_bwd_kernel, cdiv(S * K, 256), 256, 0, st, pidx, S * K, d_conf_coef, g.conf);
        SGN_LAUNCH_CHECK();
    }
    if ((rc = wgrad(ws.d_araw, 1, 1, Hlast, d.W, d.W, d_weights ? d_weights[la] : nullptr, d.W, T_ptr, Tm))) return rc;
    if ((rc = bias_grad(ws.d_araw, 1, 1, T_ptr, Tm, d_biases ? d_biases[la] : nullptr))) return rc;

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
        if ((rc = bias_grad(dZ, d.W, d.W, T_ptr, Tm, d_biases ? d_biases[t] : nullptr))) return rc;
        if (L.extra == EXTRA_COLORDIR && (g.color || g.dir)) {
            GemmNN q = {};
            q.A1 = dZ; q.lda1 = d.W; q.B1 = ws.Wp[t] + a_kpad; q.ldb1 = L.kpad; q.K1 = d.W;
            q.C = ws.dE7; q.ldc = 8; q.N = 8; q.m_ptr = T_ptr; q.m_max = Tm; q.epi = EPI_NONE;
            if ((rc = launch_gemm_nn(q, st))) return rc;
            dE7 = ws.dE7;
        }
        if (t > 0 || g.embedding) {
            float* dnext = t > 0 ? ws.dZ[(nt - t) & 1] : ws.dX0;
            GemmNN q = {};
            q.A1 = dZ; q.lda1 = d.W; q.B1 = ws.Wp[t]; q.ldb1 = L.kpad; q.K1 = d.W;
            q.C = dnext; q.ldc = a_kpad; q.N = a_kpad; q.m_ptr = T_ptr; q.m_max = Tm; q.slope = d.slope;
            if (t > 0) { q.epi = EPI_MUL_DLEAKY; q.aux = a_in; q.ldaux = a_ld; } else { q.epi = EPI_NONE; }
            if ((rc = launch_gemm_nn(q, st))) return rc;
            dZ = dnext;
        }
    }
    if (g.embedding || g.color || g.dir) {
        launch(agg_scatter_kernel, cdiv(Tm, 8), 256, 0, st, in, d, K, SR, T_ptr, Tm, ws.tuple_src, ws.X0, ws.dX0, dE7, g);
        SGN_LAUNCH_CHECK();
    }
    return SGN_OK;
}
