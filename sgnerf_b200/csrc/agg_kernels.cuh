// agg_kernels.cuh -- per-sample / per-tuple kernels shared by the fp32 and the tensor-core aggregator paths.
#pragma once
#include "agg_common.cuh"

namespace sgn {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// One thread per sample.  point_aggregators.py:885, :917-925 (dists), :494-502 + :946-947 (weights), :953 (conf).
static __global__ void agg_prepare_kernel(AggIn in, int64_t S, int K, float* __restrict__ loc_pers, float* __restrict__ loc_depth, float* __restrict__ wc,
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
        if (loc_depth) loc_depth[s] = c2;          // the camera depth again as a dense array: what the frame tail reads (4 B instead of a 12-B-stride gather)
    }
    const int32_t* pi = in.pidx + s * K;
    if (K == 8 && (((uintptr_t)in.pidx | (uintptr_t)wc | (uintptr_t)weight_n | (uintptr_t)weight_out | (uintptr_t)conf_out) & 15) == 0) {
        // canonical K: the eight indices as two 16-byte loads, every output row as two 16-byte stores
        // a slot without a sample (mask 0) has an all -1 row by construction: the 32 bytes are not even read
        const bool slot = !in.smask || __ldg(in.smask + s) > 0;
        int4 pa = make_int4(-1, -1, -1, -1), pb = pa;
        if (slot) { pa = __ldg((const int4*)pi); pb = __ldg((const int4*)pi + 1); }
        const int p8[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
        if ((pa.x & pa.y & pa.z & pa.w & pb.x & pb.y & pb.z & pb.w) < 0) {
            // no neighbour at all (two thirds of the slots of a frame): zero weights; the confidence column still holds point 0's
            // wc and weight_n are workspaces that are only ever read at valid tuples (tuple_src / pidx >= 0): nothing to write for this sample
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4* o;
            if (weight_out) { o = (float4*)(weight_out + s * 8); o[0] = z4; o[1] = z4; }
            if (conf_out) {
                const float c0 = in.tab.conf ? fminf(fmaxf(__ldg(in.tab.conf), 0.0001f), 1.0f) : 1.0f;
                const float4 c4 = make_float4(c0, c0, c0, c0);
                o = (float4*)(conf_out + s * 8); o[0] = c4; o[1] = c4;
            }
            ray_valid[s] = 0;
            nvalid[s] = 0;
            svalid[s] = 0;
            return;
        }
        float w8[8], cf8[8];
        float sum = 0.f;
        int n = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int p = p8[k];
            float wk = 0.f;
            if (p >= 0) {
                const float dx = in.tab.xyz[3 * (int64_t)p] - lx, dy = in.tab.xyz[3 * (int64_t)p + 1] - ly, dz = in.tab.xyz[3 * (int64_t)p + 2] - lz;
                const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
                wk = 1.0f / fmaxf(nrm, 1e-6f);
                n++;
            }
            w8[k] = wk;
            sum += wk;
            const int pc = p < 0 ? 0 : p;                  // the reference gathers with clamp(pidx, 0) (neural_points.py:958)
            cf8[k] = in.tab.conf ? fminf(fmaxf(__ldg(in.tab.conf + pc), 0.0001f), 1.0f) : 1.0f;
        }
        const float den = fmaxf(sum, 1e-8f);
#pragma unroll
        for (int k = 0; k < 8; k++) w8[k] = w8[k] / den;
        float4* o = (float4*)(wc + s * 8);
        o[0] = make_float4(w8[0] * cf8[0], w8[1] * cf8[1], w8[2] * cf8[2], w8[3] * cf8[3]);
        o[1] = make_float4(w8[4] * cf8[4], w8[5] * cf8[5], w8[6] * cf8[6], w8[7] * cf8[7]);
        if (weight_n) { o = (float4*)(weight_n + s * 8); o[0] = make_float4(w8[0], w8[1], w8[2], w8[3]); o[1] = make_float4(w8[4], w8[5], w8[6], w8[7]); }
        if (weight_out) { o = (float4*)(weight_out + s * 8); o[0] = make_float4(w8[0], w8[1], w8[2], w8[3]); o[1] = make_float4(w8[4], w8[5], w8[6], w8[7]); }
        if (conf_out) { o = (float4*)(conf_out + s * 8); o[0] = make_float4(cf8[0], cf8[1], cf8[2], cf8[3]); o[1] = make_float4(cf8[4], cf8[5], cf8[6], cf8[7]); }
        ray_valid[s] = n > 0;
        nvalid[s] = n;
        svalid[s] = n > 0;
        return;
    }
    float w[SGN_MAX_K];
    float sum = 0.f;
    int n = 0;
    const bool slot_g = !in.smask || __ldg(in.smask + s) > 0;     // mask 0: the row is all -1 by construction and may not even be written
    for (int k = 0; k < K; k++) {
        const int p = slot_g ? pi[k] : -1;
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
        const int p = (!slot_g || pi[k] < 0) ? 0 : pi[k];   // the reference gathers with clamp(pidx, 0) (neural_points.py:958)
        const float cf = in.tab.conf ? fminf(fmaxf(in.tab.conf[p], 0.0001f), 1.0f) : 1.0f;
        const float wn = w[k] / den;
        wc[s * K + k] = wn * cf;
        if (weight_n) weight_n[s * K + k] = wn;
        if (weight_out) weight_out[s * K + k] = wn;
        if (conf_out) conf_out[s * K + k] = cf;
    }
    ray_valid[s] = n > 0;
    nvalid[s] = n;
    svalid[s] = n > 0;
}

// One thread per sample: tuple j -> (sample, slot), compact sample c -> sample.
static __global__ void agg_index_kernel(const int32_t* __restrict__ pidx, int64_t S, int K, const int32_t* __restrict__ tuple_start,
                                 const int32_t* __restrict__ sample_cidx, const int32_t* __restrict__ nvalid,
                                 int32_t* __restrict__ tuple_src, int32_t* __restrict__ csample, int32_t* __restrict__ tuple_pt = nullptr,
                                 int32_t* __restrict__ tuple_cs = nullptr)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (nvalid[s] == 0) return;
    const int c = sample_cidx[s];
    csample[c] = (int32_t)s;
    int j = tuple_start[s];
    for (int k = 0; k < K; k++) {
        const int p = pidx[s * K + k];
        if (p >= 0) {
            tuple_src[j] = (int32_t)(s * K + k);
            // the training kernels walk the tuples warp by warp: point and compact sample per tuple save them two dependent loads each
            if (tuple_pt) { tuple_pt[j] = p; tuple_cs[j] = c; }
            j++;
        }
    }
}

// One warp per tuple.  X0 layout = reference `feat` (:603-611):
//   [0,C) embedding | C + 2*(d*F+f) + {0:sin,1:cos} of emb_d * 2^f | then the same for the 6 dists with F = dist_xyz_freq.
// E7 = [colour(3) | dir - viewdir (3) | dir . viewdir | 0]  (:639-652).
static __global__ void __launch_bounds__(256)
agg_gather_kernel(AggIn in, AggDims d, int K, int SR, const int32_t* __restrict__ T_ptr, int T_max, const int32_t* __restrict__ tuple_src,
                  const int32_t* __restrict__ tuple_pt, const float* __restrict__ loc_pers, float* __restrict__ X0, float* __restrict__ L, float* __restrict__ E7)
{
    extern __shared__ float4 rowbuf4[];                // 8 warps x k0pad floats (dynamic)
    float* rowbuf = (float*)rowbuf4;
    const int lane = lane_id();
    const int T = min(*T_ptr, T_max);
    // a fixed grid strides over the items: their number lives on the device, a grid sized for the maximum would be mostly empty blocks.
    // The next tuple's indices are fetched while the current one is worked on (the loop is a chain of dependent loads otherwise).
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int flat_n = 0, p_n = 0;
    if (j < T) { flat_n = tuple_src[j]; p_n = tuple_pt[j]; }
    for (; j < T; j += stride) {
    const int flat = flat_n;
    const int64_t p = p_n;
    if (j + stride < T) { flat_n = tuple_src[j + stride]; p_n = tuple_pt[j + stride]; }
    const int64_t s = flat / K;
    const int64_t r = s / SR;
    // the row is put together in shared memory (its sin / cos columns interleave with a stride of 2 F floats) and leaves as whole float4s
    float* x = rowbuf + (threadIdx.x >> 5) * d.k0pad;
    const int C = d.C, F = d.F, FD = d.FD;
    for (int c = lane; c < C; c += 32) {
        const float e = __ldg(in.tab.embedding + p * C + c);
        x[c] = e;
        // sin / cos of e 2^f: one sincosf, then the double-angle recurrence (|error| grows ~2x per octave from 1 ulp: < 1e-6 at F <= 5)
        float sn, cs;
        sincosf(e, &sn, &cs);
        for (int f = 0; f < F; f++) {
            x[C + 2 * (c * F + f)] = sn;
            x[C + 2 * (c * F + f) + 1] = cs;
            const float s2 = 2.0f * sn * cs;
            cs = fmaf(-2.0f * sn, sn, 1.0f);
            sn = s2;
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
        float sn, cs;
        sincosf(dist, &sn, &cs);
        for (int f = 0; f < FD; f++) {
            x[base + 2 * (lane * FD + f)] = sn;
            x[base + 2 * (lane * FD + f) + 1] = cs;
            const float s2 = 2.0f * sn * cs;
            cs = fmaf(-2.0f * sn, sn, 1.0f);
            sn = s2;
        }
    }
    for (int c = d.k0 + lane; c < d.k0pad; c += 32) x[c] = 0.f;
    __syncwarp();
    {
        float4* dst = (float4*)(X0 + j * d.k0pad);
        const float4* src = (const float4*)x;
        for (int i = lane; i < (d.k0pad >> 2); i += 32) dst[i] = src[i];
    }
    __syncwarp();
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
}

static __device__ __forceinline__ float softplus1(float x)  // torch.nn.Softplus(beta=1, threshold=20)
{
    return x > 20.0f ? x : log1pf(expf(x));
}

// One warp per compact sample: raw alpha of its tuples (alpha_branch, a single Linear: araw = h . wa + ba, kept for the backward),
// sigma = sum_k wc * softplus(raw - 1), F = sum_k wc * h  -> C0[:, :W] -- one pass over the sample's rows of H for all three;
// C0[:, W:W+6*FV] = viewdir encoding (ori=True, first three stripped): sin(v_d 2^f) d-major, then cos (:579-585).
template <int KT>                                  // K <= KT
static __global__ void __launch_bounds__(256)
agg_ksum_kernel(AggIn in, AggDims d, int K, int SR, const int32_t* __restrict__ S_ptr, int S_max, const int32_t* __restrict__ csample,
                const int32_t* __restrict__ tuple_start, const int32_t* __restrict__ nvalid, const float* __restrict__ wc,
                const float* __restrict__ H, const float* __restrict__ wa, const float* __restrict__ ba, float* __restrict__ araw,
                float* __restrict__ C0, float* __restrict__ sigma)
{
    const int lane = lane_id();
    const int Sv = min(*S_ptr, S_max);
    // a fixed grid strides over the items: their number lives on the device, a grid sized for the maximum would be mostly empty blocks
    for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < Sv; c += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t s = csample[c];
    const int j0 = tuple_start[s], n = nvalid[s];
    const int W = d.W;
    float wk[KT], dotq[KT];
    {
        int q = 0;
#pragma unroll
        for (int k = 0; k < KT; k++) {
            wk[k] = 0.f; dotq[k] = 0.f;
        }
        for (int k = 0; k < K; k++)
            if (in.pidx[s * K + k] >= 0) {
                const float w = wc[s * K + k];
#pragma unroll
                for (int i = 0; i < KT; i++) if (i == q) wk[i] = w;
                q++;
            }
    }
    float* row = C0 + c * d.kc0pad;
    for (int col = lane; col < W; col += 32) {
        const float wac = __ldg(wa + col);
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < KT; q++)
            if (q < n) {
                const float h = H[(int64_t)(j0 + q) * W + col];
                acc += h * wk[q];
                dotq[q] = fmaf(h, wac, dotq[q]);
            }
        row[col] = acc;
    }
    float sg = 0.f;
    const float b0 = ba[0];
#pragma unroll
    for (int q = 0; q < KT; q++)
        if (q < n) {
            const float a = warp_sum(dotq[q]) + b0;
            if (lane == 0) araw[j0 + q] = a;
            sg += wk[q] * (d.act_super ? softplus1(a - 1.0f) : fmaxf(a, 0.f));
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
}

// One warp per compact sample: rgb = sigmoid(c . Wlast^T + b) (*1.002 - 0.001), decoded[s] = (sigma, rgb)
static __global__ void __launch_bounds__(256)
agg_rgb_kernel(AggDims d, const int32_t* __restrict__ S_ptr, int S_max, const int32_t* __restrict__ csample,
               const float* __restrict__ Cin, int Wc, const float* __restrict__ Wl, const float* __restrict__ bl,
               const float* __restrict__ sigma, float* __restrict__ decoded, float* __restrict__ sig_out)
{
    const int lane = lane_id();
    const int Sv = min(*S_ptr, S_max);
    // a fixed grid strides over the items: their number lives on the device, a grid sized for the maximum would be mostly empty blocks
    for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < Sv; c += (int64_t)gridDim.x * (blockDim.x >> 5)) {
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
}

// W [N,K] (torch Linear) -> Wt [Kpad, Npad] (k-major, forward B operand) and Wp [Npad, Kpad] (dgrad B operand)
static __global__ void pack_weight_kernel(const float* __restrict__ W, int N, int Kin, int Npad, int Kpad, float* __restrict__ Wt, float* __restrict__ Wp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Npad * Kpad) return;
    const int n = i / Kpad, k = i - n * Kpad;
    const float v = (n < N && k < Kin) ? W[(size_t)n * Kin + k] : 0.f;
    Wp[(size_t)n * Kpad + k] = v;
    Wt[(size_t)k * Npad + n] = v;
}

// All layers in one launch: blockIdx.y = layer.
struct PackJobs {
    const float* W[16]; float* Wt[16]; float* Wp[16];
    int N[16], Kin[16], Npad[16], Kpad[16];
};
static __global__ void pack_weights_kernel(PackJobs J)
{
    const int l = blockIdx.y;
    const int Npad = J.Npad[l], Kpad = J.Kpad[l], N = J.N[l], Kin = J.Kin[l];
    const float* __restrict__ W = J.W[l];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Npad * Kpad; i += gridDim.x * blockDim.x) {
        const int n = i / Kpad, k = i - n * Kpad;
        const float v = (n < N && k < Kin) ? W[(size_t)n * Kin + k] : 0.f;
        J.Wp[l][(size_t)n * Kpad + k] = v;
        J.Wt[l][(size_t)k * Npad + n] = v;
    }
}

// ---- backward-only kernels ----

// Skinny reductions over the rows: out[p, c] += sum_m w_p[m] X[m, c] for p < NP, with w_p[m] = A[m, p], or 1 when A is NULL
// (bias gradients = column sums; the wgrads of the 1- and 3-output layers alpha_branch.0 / color_branch.6).
// Block = 64 column quads x 4 row groups over SKINNY_ROWS rows; four 16-byte loads in flight per thread.
constexpr int SKINNY_ROWS = 64;                    // rows per block: the remaining users reduce over the samples (tens of thousands of rows), so many small blocks
template <int NP>
static __global__ void __launch_bounds__(256)
skinny_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ X, int ld, int ncols, const int32_t* __restrict__ m_ptr, int m_max,
                 float* __restrict__ out, int ldo)
{
    const int M = min(*m_ptr, m_max);
    const int m0 = blockIdx.x * SKINNY_ROWS;
    if (m0 >= M) return;
    const int m1 = min(M, m0 + SKINNY_ROWS);
    const int cq = threadIdx.x & 63, rg = threadIdx.x >> 6;
    const int c = blockIdx.y * 256 + 4 * cq;
    const bool vec = (ld & 3) == 0 && c + 4 <= ncols && (((uintptr_t)X) & 15) == 0;
    float acc[NP][4];
#pragma unroll
    for (int p = 0; p < NP; p++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[p][j] = 0.f;
    if (c < ncols) {
#pragma unroll 4
        for (int m = m0 + rg; m < m1; m += 4) {
            const float* src = X + (size_t)m * ld + c;
            float4 x;
            if (vec) x = *(const float4*)src;
            else {
                x.x = src[0];
                x.y = c + 1 < ncols ? src[1] : 0.f;
                x.z = c + 2 < ncols ? src[2] : 0.f;
                x.w = c + 3 < ncols ? src[3] : 0.f;
            }
#pragma unroll
            for (int p = 0; p < NP; p++) {
                const float w = A ? __ldg(A + (size_t)m * lda + p) : 1.f;
                acc[p][0] = fmaf(w, x.x, acc[p][0]); acc[p][1] = fmaf(w, x.y, acc[p][1]);
                acc[p][2] = fmaf(w, x.z, acc[p][2]); acc[p][3] = fmaf(w, x.w, acc[p][3]);
            }
        }
    }
    __shared__ float red[3][64][NP * 4 + 1];
    if (rg > 0) {
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 4; j++) red[rg - 1][cq][p * 4 + j] = acc[p][j];
    }
    __syncthreads();
    if (rg == 0 && c < ncols) {
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float v = acc[p][j] + red[0][cq][p * 4 + j] + red[1][cq][p * 4 + j] + red[2][cq][p * 4 + j];
                if (c + j < ncols) atomicAdd(out + (size_t)p * ldo + c + j, v);
            }
    }
}

// One warp per compact sample: d_raw[c, 0:3] = d_rgb * scale * sig (1 - sig), padded to 8 columns
static __global__ void __launch_bounds__(256)
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
// The three reductions over the tuples that read the same rows are folded in (they were separate passes over H and dZ): the weight and
// bias gradient of the alpha head (d_wa[col] += da h[col], d_ba += da) and the bias gradient of the last tuple layer (d_blast = column
// sums of dZ); partial sums per lane over the tuples its warp walks, per block through shared memory, then one red.add per column per block.
template <int NC>                                  // columns per lane: W <= 32 * NC
static __global__ void __launch_bounds__(256, 4)
agg_ksum_bwd_kernel(AggIn in, AggDims d, int K, const int32_t* __restrict__ T_ptr, int T_max, const int32_t* __restrict__ tuple_src,
                    const int32_t* __restrict__ tuple_cs, const float* __restrict__ wc, const float* __restrict__ weight_n,
                    const float* __restrict__ H, const float* __restrict__ araw, const float* __restrict__ wa,
                    const float* __restrict__ dC0, int lddc0, const float* __restrict__ d_decoded, float* __restrict__ dZ,
                    float* __restrict__ d_araw, float* __restrict__ d_conf, float* __restrict__ d_wa, float* __restrict__ d_ba,
                    float* __restrict__ d_blast)
{
    const int lane = lane_id();
    const int T = min(*T_ptr, T_max);
    // partial sums of the three reductions: a lane keeps those of its own columns over all the tuples its warp walks
    float g_wa[NC], g_bl[NC], g_ba = 0.f;
#pragma unroll
    for (int i = 0; i < NC; i++) { g_wa[i] = 0.f; g_bl[i] = 0.f; }
    // a fixed grid strides over the items (see agg_gather_kernel); the next tuple's indices are fetched ahead
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int flat_n = 0, c_n = 0;
    if (j < T) { flat_n = tuple_src[j]; c_n = tuple_cs[j]; }
    for (; j < T; j += stride) {
    const int flat = flat_n;
    const int64_t c = c_n;
    if (j + stride < T) { flat_n = tuple_src[j + stride]; c_n = tuple_cs[j + stride]; }
    const int64_t s = flat / K;
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
#pragma unroll
    for (int i = 0; i < NC; i++) {
        const int col = lane + 32 * i;
        if (col < W) {
            const float h = H[j * W + col];
            const float df = dC0[c * lddc0 + col];
            dot = fmaf(h, df, dot);
            const float dh = w * df + da * __ldg(wa + col);
            const float dz = dh * (h > 0.f ? 1.0f : d.slope);
            dZ[j * W + col] = dz;
            g_wa[i] = fmaf(da, h, g_wa[i]);
            g_bl[i] += dz;
        }
    }
    dot = warp_sum(dot);
    if (lane == 0) {
        d_araw[j] = da;
        g_ba += da;
        if (d_conf) {
            const float d_wc = act * dsig + dot;
            atomicAdd(d_conf + in.pidx[flat], weight_n[flat] * d_wc);
        }
    }
    }
    if (!d_wa && !d_ba && !d_blast) return;
    // block: sum the warps' partials column by column (two passes over one buffer), then one red.add per column per block
    __shared__ float part[8][32 * NC + 1];
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int pass = 0; pass < 2; pass++) {
        float* out = pass == 0 ? d_wa : d_blast;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NC; i++) part[wid][lane + 32 * i] = pass == 0 ? g_wa[i] : g_bl[i];
        if (pass == 0 && lane == 0) part[wid][32 * NC] = g_ba;
        __syncthreads();
        if (out)
            for (int col = threadIdx.x; col < d.W; col += blockDim.x) {
                float a = 0.f;
                for (int w2 = 0; w2 < nw; w2++) a += part[w2][col];
                if (a != 0.f) atomicAdd(out + col, a);
            }
        if (pass == 0 && threadIdx.x == 0 && d_ba) {
            float a = 0.f;
            for (int w2 = 0; w2 < nw; w2++) a += part[w2][32 * NC];
            if (a != 0.f) atomicAdd(d_ba, a);
        }
    }
}

// cotangent of the conf_coefficient output (all slots, invalid ones use point 0 like the reference's clamp(pidx,0) gather)
static __global__ void __launch_bounds__(256)
agg_conf_out_bwd_kernel(const int32_t* __restrict__ pidx, int64_t n, const float* __restrict__ d_conf_coef, float* __restrict__ d_conf)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float to0 = 0.f;                       // invalid slots all land on point 0: one atomic per block instead of one per slot
    if (i < n) {
        const float g = d_conf_coef[i];
        const int p = pidx[i];
        if (p <= 0) to0 = g;
        else if (g != 0.f) atomicAdd(d_conf + p, g);
    }
    to0 = warp_sum(to0);
    __shared__ float part[8];
    if (lane_id() == 0) part[threadIdx.x >> 5] = to0;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; w++) s += part[w];
        if (s != 0.f) atomicAdd(d_conf, s);
    }
}

// One warp per tuple: scatter-add into the point tables.
//   d emb_c = dX0[c] + sum_f 2^f (cos_cf dsin_cf - sin_cf dcos_cf)   (sin/cos read back from the saved X0)
//   d colour = dE7[0:3] ; d dir = dE7[3:6] + viewdir * dE7[6]
static __global__ void __launch_bounds__(256)
agg_scatter_kernel(AggIn in, AggDims d, int K, int SR, const int32_t* __restrict__ T_ptr, int T_max, const int32_t* __restrict__ tuple_src,
                   const int32_t* __restrict__ tuple_pt, const float* __restrict__ X0, const float* __restrict__ dX0, const float* __restrict__ dE7, SgnPointGrads g)
{
    extern __shared__ float4 rowbuf4[];                // 8 warps x 2 x k0pad floats (dynamic)
    float* rowbuf = (float*)rowbuf4;
    const int lane = lane_id();
    const int T = min(*T_ptr, T_max);
    // a fixed grid strides over the items (see agg_gather_kernel); the next tuple's indices are fetched ahead
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int flat_n = 0, p_n = 0;
    if (j < T) { flat_n = tuple_src[j]; p_n = tuple_pt[j]; }
    for (; j < T; j += stride) {
    const int flat = flat_n;
    const int64_t p = p_n;
    if (j + stride < T) { flat_n = tuple_src[j + stride]; p_n = tuple_pt[j + stride]; }
    const int C = d.C, F = d.F;
    if (g.embedding) {
        // both rows arrive as whole float4s; the sin / cos columns are then read from shared memory
        float* x = rowbuf + (threadIdx.x >> 5) * 2 * d.k0pad;
        float* dx = x + d.k0pad;
        {
            const float4* sx = (const float4*)(X0 + j * d.k0pad);
            const float4* sd = (const float4*)(dX0 + j * d.k0pad);
            for (int i = lane; i < (d.k0pad >> 2); i += 32) { ((float4*)x)[i] = sx[i]; ((float4*)dx)[i] = sd[i]; }
        }
        __syncwarp();
        for (int c = lane; c < C; c += 32) {
            float acc = dx[c], fr = 1.0f;
            for (int f = 0; f < F; f++) {
                const int o = C + 2 * (c * F + f);
                acc += fr * (x[o + 1] * dx[o] - x[o] * dx[o + 1]);
                fr *= 2.0f;
            }
            atomicAdd(g.embedding + p * C + c, acc);
        }
        __syncwarp();
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
}

}  // namespace sgn
