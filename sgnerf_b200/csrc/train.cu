// train.cu -- the pieces of a training iteration around the aggregator / compositing backward: the reference's loss with its gradient
// in one pass, and Adam over the rows of the point tables that ever received a gradient.
//
// Reference: models/base_rendering_model.py:543-641 (compute_losses: `ray_masked_coarse_raycolor` colour MSE + 1e-6 per item, zero-one
// regulariser on conf_coefficient with --zero_epsilon), models/mvs_points_volumetric_model.py:67-109 (two torch.optim.Adam instances,
// betas (0.9, 0.999), no weight decay: MLP weights at --lr, point tables at --plr).
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>

#include "common.cuh"

namespace sgn {

constexpr int LOSS_THREADS = 256;

// number of rays that hit the cloud (the reference's masked_select row count), accumulated into *count
__global__ void loss_hit_count_kernel(const int8_t* __restrict__ ray_mask, int64_t R, float* count)
{
    int c = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) c += ray_mask[r] > 0;
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int s[LOSS_THREADS / 32];
    if (lane_id() == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < LOSS_THREADS / 32; i++) t += s[i];
        if (t) atomicAdd(count, (float)t);
    }
}

// loss = color_w * sum_hit |c - gt|^2 / (3 n) + conf_w * sum_hit (log v + log(1 - v)) / (n SR K) + const_term,  v = clamp(conf, eps, 1 - eps),
// n = max(*hit_count, 1) (the GLOBAL hit count when ranks share a step).  One pass: the loss terms are reduced per block and added to
// *loss (which the caller zeroes... the first block adds const_term), the gradients w.r.t. ray_color and conf_coefficient are written
// for every row (zero for rays that missed; zero where the clamp is active, as torch.clamp's backward gives).
__global__ void loss_fwd_bwd_kernel(const float* __restrict__ ray_color, const float* __restrict__ gt, const int8_t* __restrict__ ray_mask,
                                    const float* __restrict__ conf, int64_t R, int SRK, const float* __restrict__ hit_count, float color_w,
                                    float conf_w, float eps, float const_term, float* loss, float* __restrict__ d_color, float* __restrict__ d_conf)
{
    const float n = fmaxf(*hit_count, 1.0f);
    const float kc = color_w / (3.0f * n), kz = conf_w / (n * (float)SRK);
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = i0; i < R * 3; i += stride) {
        const bool hit = ray_mask[i / 3] > 0;
        const float d = hit ? ray_color[i] - gt[i] : 0.f;
        acc += kc * d * d;
        if (d_color) d_color[i] = 2.0f * kc * d;
    }
    if (conf) {
        const int64_t total = R * SRK;
        for (int64_t i = i0; i < total; i += stride) {
            const bool hit = ray_mask[i / SRK] > 0;
            float g = 0.f;
            if (hit) {
                const float c = conf[i];
                const float v = fminf(fmaxf(c, eps), 1.0f - eps);
                acc += kz * (logf(v) + logf(1.0f - v));
                if (c >= eps && c <= 1.0f - eps) g = kz * (1.0f / v - 1.0f / (1.0f - v));
            }
            if (d_conf) d_conf[i] = g;
        }
    }
    acc = warp_sum(acc);
    __shared__ float s[LOSS_THREADS / 32];
    if (lane_id() == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = blockIdx.x == 0 ? const_term : 0.f;
        for (int i = 0; i < LOSS_THREADS / 32; i++) t += s[i];
        atomicAdd(loss, t);
    }
}

// Adam (torch.optim.Adam arithmetic: m = b1 m + (1 - b1) g, v = b2 v + (1 - b2) g^2, p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps))
// over the rows of a [N, C] table.  One thread per row.  `active` (may be NULL = every row) remembers the rows that ever received a
// non-zero gradient: a row that never did has m = v = 0 and dense Adam leaves it where it is, so skipping it gives the same table;
// a row that did keeps being updated every step (its moments decay), exactly like the dense optimiser.  zero_grad != 0 clears the
// gradient rows that were non-zero, so the [N, C] accumulator never needs a dense memset.
template <int VEC>
__global__ void adam_rows_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ m1, float* __restrict__ m2,
                                 uint8_t* __restrict__ active, int64_t N, int C, float lr, float b1, float b2, float eps, const float* __restrict__ step,
                                 float grad_scale, int zero_grad)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= N) return;
    float* g = grad + row * C;
    bool nz = false;
    if (VEC == 4) {
        for (int c = 0; c < C; c += 4) { const float4 x = *(const float4*)(g + c); nz = nz || x.x != 0.f || x.y != 0.f || x.z != 0.f || x.w != 0.f; }
    } else {
        for (int c = 0; c < C; c++) nz = nz || g[c] != 0.f;
    }
    if (active) {
        if (nz) active[row] = 1;
        else if (!active[row]) return;
    }
    const float t = *step;
    const float bc1 = 1.0f - powf(b1, t), bc2s = sqrtf(1.0f - powf(b2, t));
    const float step_size = lr / bc1;
    float *p = param + row * C, *a = m1 + row * C, *b = m2 + row * C;
    for (int c = 0; c < C; c++) {
        const float gc = g[c] * grad_scale;
        const float m = b1 * a[c] + (1.0f - b1) * gc;
        const float v = b2 * b[c] + (1.0f - b2) * gc * gc;
        a[c] = m; b[c] = v;
        p[c] -= step_size * (m / (sqrtf(v) / bc2s + eps));
        if (zero_grad && nz) g[c] = 0.f;
    }
}

// The same for several tables that share their rows (the point tables: embedding [N,32], colour [N,3], dir [N,3], conf [N]) in ONE pass:
// eight lanes per row, lane l owns the elements [4 l, 4 l + 4) of every table (16-byte accesses on the wide table: a warp covers four
// consecutive rows = 512 contiguous bytes), the row is active when any element of any table has a non-zero gradient.
constexpr int ADAM_MAX_TABLES = 8;
struct AdamTables {
    float* param[ADAM_MAX_TABLES]; float* grad[ADAM_MAX_TABLES]; float* m1[ADAM_MAX_TABLES]; float* m2[ADAM_MAX_TABLES];
    int C[ADAM_MAX_TABLES];
    int n;
};

__global__ void adam_rows_multi_kernel(AdamTables T, uint8_t* __restrict__ active, int64_t N, float lr, float b1, float b2, float eps,
                                       const float* __restrict__ step, float grad_scale, int zero_grad)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = gid >> 3;
    const int l = (int)(gid & 7);
    const unsigned group = 0xffu << (lane_id() & 24);
    const bool inb = row < N;
    bool nz = false;
    float4 gv[ADAM_MAX_TABLES];
#pragma unroll
    for (int k = 0; k < ADAM_MAX_TABLES; k++) {
        gv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < T.n && inb) {
            const int C = T.C[k], e0 = 4 * l;
            const float* g = T.grad[k] + row * C;
            if (e0 + 4 <= C && (C & 3) == 0) gv[k] = *(const float4*)(g + e0);
            else {
                if (e0 < C) gv[k].x = g[e0];
                if (e0 + 1 < C) gv[k].y = g[e0 + 1];
                if (e0 + 2 < C) gv[k].z = g[e0 + 2];
                if (e0 + 3 < C) gv[k].w = g[e0 + 3];
            }
            nz = nz || gv[k].x != 0.f || gv[k].y != 0.f || gv[k].z != 0.f || gv[k].w != 0.f;
        }
    }
    const bool row_nz = (__ballot_sync(0xffffffffu, nz) & group) != 0u;
    if (!inb) return;
    if (active) {
        if (row_nz) { if (l == 0) active[row] = 1; }
        else if (!active[row]) return;
    }
    const float t = *step;
    const float bc1 = 1.0f - powf(b1, t), bc2s = sqrtf(1.0f - powf(b2, t));
    const float step_size = lr / bc1;
#pragma unroll
    for (int k = 0; k < ADAM_MAX_TABLES; k++) {
        if (k >= T.n) break;
        const int C = T.C[k], e0 = 4 * l;
        if (e0 >= C) continue;
        float *p = T.param[k] + row * C, *a = T.m1[k] + row * C, *b = T.m2[k] + row * C, *g = T.grad[k] + row * C;
        const float gg[4] = {gv[k].x, gv[k].y, gv[k].z, gv[k].w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = e0 + i;
            if (c >= C) break;
            const float gc = gg[i] * grad_scale;
            const float m = b1 * a[c] + (1.0f - b1) * gc;
            const float v = b2 * b[c] + (1.0f - b2) * gc * gc;
            a[c] = m; b[c] = v;
            p[c] -= step_size * (m / (sqrtf(v) / bc2s + eps));
            if (zero_grad && row_nz) g[c] = 0.f;
        }
    }
}

__global__ void step_increment_kernel(float* step) { *step += 1.0f; }

// ---- dense Adam over many small tensors in one launch (the ~30 MLP weights and biases): blockIdx.y = tensor -----------------------------
constexpr int ADAM_MAX_TENSORS = 64;
struct AdamTensors {
    float* param[ADAM_MAX_TENSORS]; float* grad[ADAM_MAX_TENSORS]; float* m1[ADAM_MAX_TENSORS]; float* m2[ADAM_MAX_TENSORS];
    int64_t n[ADAM_MAX_TENSORS];
};

__global__ void adam_dense_multi_kernel(AdamTensors T, float lr, float b1, float b2, float eps, const float* __restrict__ step, float grad_scale, int zero_grad)
{
    const int k = blockIdx.y;
    const float t = *step;
    const float bc1 = 1.0f - powf(b1, t), bc2s = sqrtf(1.0f - powf(b2, t));
    const float step_size = lr / bc1;
    float *p = T.param[k], *g = T.grad[k], *a = T.m1[k], *b = T.m2[k];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T.n[k]; i += (int64_t)gridDim.x * blockDim.x) {
        const float gc = g[i] * grad_scale;
        const float m = b1 * a[i] + (1.0f - b1) * gc;
        const float v = b2 * b[i] + (1.0f - b2) * gc * gc;
        a[i] = m; b[i] = v;
        p[i] -= step_size * (m / (sqrtf(v) / bc2s + eps));
        if (zero_grad) g[i] = 0.f;
    }
}

// ---- the same update driven by a LIST of the active rows (rows that ever received a non-zero gradient) --------------------------------
// A step touches ~5 % of a 1M-point cloud; reading every row's gradient to find them costs more than the update itself.  The rows a step
// can touch are known from the query (sample_pidx >= 0): mark_rows_kernel flags them, adam_append_kernel moves newly active ones into
// the list, adam_list_kernel updates the listed rows.  `touched` is a float array so that, with several ranks, it can ride in the same
// all-reduce as the gradients (sum > 0 = touched on some rank).
__global__ void mark_rows_kernel(const int32_t* __restrict__ pidx, int64_t n, float* __restrict__ touched)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = pidx[i];
    if (p >= 0) touched[p] = 1.0f;
}

__global__ void adam_append_kernel(AdamTables T, uint8_t* __restrict__ active, int32_t* __restrict__ list, int32_t* count, float* __restrict__ touched, int64_t N)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= N || touched[row] == 0.f) return;
    touched[row] = 0.f;
    if (active[row]) return;
    bool nz = false;
    for (int k = 0; k < T.n; k++) {
        const float* g = T.grad[k] + row * T.C[k];
        for (int c = 0; c < T.C[k]; c++) nz = nz || g[c] != 0.f;
    }
    if (!nz) return;
    active[row] = 1;
    list[atomicAdd(count, 1)] = (int32_t)row;
}

__global__ void adam_list_kernel(AdamTables T, const int32_t* __restrict__ list, const int32_t* __restrict__ count, float lr, float b1, float b2, float eps,
                                 const float* __restrict__ step, float grad_scale, int zero_grad)
{
    // a fixed grid walks the list (its length lives on the device; a grid sized for "every row active" would spend its time on empty blocks)
    const int l = (int)(threadIdx.x & 7);
    const float t = *step;
    const float bc1 = 1.0f - powf(b1, t), bc2s = sqrtf(1.0f - powf(b2, t));
    const float step_size = lr / bc1;
    const int64_t n = *count;
    for (int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; idx < n; idx += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const int64_t row = list[idx];
#pragma unroll
    for (int k = 0; k < ADAM_MAX_TABLES; k++) {
        if (k >= T.n) break;
        const int C = T.C[k];
        // rows of 4 floats or fewer belong to one lane; a different lane per table so that the small tables do not all land on lane 0
        const int e0 = C <= 4 ? (l == (k & 7) ? 0 : C) : 4 * l;
        if (e0 >= C) continue;
        float *p = T.param[k] + row * C, *a = T.m1[k] + row * C, *b = T.m2[k] + row * C, *g = T.grad[k] + row * C;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int c = e0 + i;
            if (c >= C) break;
            const float g0 = g[c];
            const float gc = g0 * grad_scale;
            const float m = b1 * a[c] + (1.0f - b1) * gc;
            const float v = b2 * b[c] + (1.0f - b2) * gc * gc;
            a[c] = m; b[c] = v;
            p[c] -= step_size * (m / (sqrtf(v) / bc2s + eps));
            if (zero_grad && g0 != 0.f) g[c] = 0.f;
        }
    }
    }
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_loss_hit_count(const int8_t* ray_mask, int64_t R, float* count, void* stream)
{
    SGN_CHECK_ARG(ray_mask && count && R >= 0, "sgn_loss_hit_count: bad argument");
    if (R == 0) return SGN_OK;
    int nb = cdiv(R, LOSS_THREADS);
    if (nb > 148 * 8) nb = 148 * 8;
    launch(loss_hit_count_kernel, nb, LOSS_THREADS, 0, (cudaStream_t)stream, ray_mask, R, count);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_loss_forward_backward(const float* ray_color, const float* gt, const int8_t* ray_mask, const float* conf_coef, int64_t R, int SR,
                                         int K, const float* hit_count, float color_weight, float conf_weight, float zero_eps, float const_term,
                                         float* loss, float* d_ray_color, float* d_conf_coef, void* stream)
{
    SGN_CHECK_ARG(ray_color && gt && ray_mask && hit_count && loss && R >= 0 && SR > 0 && K > 0, "sgn_loss_forward_backward: bad argument");
    SGN_CHECK_ARG(zero_eps > 0.f && zero_eps < 0.5f, "sgn_loss_forward_backward: zero_eps must lie in (0, 0.5)");
    auto st = (cudaStream_t)stream;
    SGN_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    const int64_t work = conf_coef ? R * SR * K : R * 3;
    int nb = cdiv(work > 0 ? work : 1, LOSS_THREADS);
    if (nb > 148 * 16) nb = 148 * 16;
    launch(loss_fwd_bwd_kernel, nb, LOSS_THREADS, 0, st, ray_color, gt, ray_mask, conf_coef, R, SR * K, hit_count, color_weight, conf_weight, zero_eps,
           const_term, loss, d_ray_color, d_conf_coef);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_adam_step_count(float* step, void* stream)
{
    SGN_CHECK_ARG(step != nullptr, "sgn_adam_step_count: NULL");
    launch(step_increment_kernel, 1, 1, 0, (cudaStream_t)stream, step);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_adam_rows(float* param, float* grad, float* exp_avg, float* exp_avg_sq, uint8_t* active, int64_t N, int C, float lr, float beta1,
                             float beta2, float eps, const float* step, float grad_scale, int zero_grad, void* stream)
{
    SGN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step && N >= 0 && C > 0, "sgn_adam_rows: bad argument");
    if (N == 0) return SGN_OK;
    auto st = (cudaStream_t)stream;
    const bool vec = (C % 4 == 0) && (((uintptr_t)grad & 15) == 0);
    if (vec)
        launch(adam_rows_kernel<4>, cdiv(N, 128), 128, 0, st, param, grad, exp_avg, exp_avg_sq, active, N, C, lr, beta1, beta2, eps, step, grad_scale, zero_grad);
    else
        launch(adam_rows_kernel<1>, cdiv(N, 128), 128, 0, st, param, grad, exp_avg, exp_avg_sq, active, N, C, lr, beta1, beta2, eps, step, grad_scale, zero_grad);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_adam_rows_multi(int n_tables, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                                   const int32_t* C, uint8_t* active, int64_t N, float lr, float beta1, float beta2, float eps, const float* step,
                                   float grad_scale, int zero_grad, void* stream)
{
    SGN_CHECK_ARG(n_tables > 0 && n_tables <= ADAM_MAX_TABLES && params && grads && exp_avg && exp_avg_sq && C && step && N >= 0,
                  "sgn_adam_rows_multi: bad argument (at most %d tables)", ADAM_MAX_TABLES);
    AdamTables T = {};
    T.n = n_tables;
    for (int k = 0; k < n_tables; k++) {
        SGN_CHECK_ARG(C[k] > 0 && C[k] <= 32 && params[k] && grads[k] && exp_avg[k] && exp_avg_sq[k], "sgn_adam_rows_multi: table %d: NULL or more than 32 columns", k);
        T.param[k] = params[k]; T.grad[k] = grads[k]; T.m1[k] = exp_avg[k]; T.m2[k] = exp_avg_sq[k]; T.C[k] = C[k];
    }
    if (N == 0) return SGN_OK;
    launch(adam_rows_multi_kernel, cdiv(N * 8, 256), 256, 0, (cudaStream_t)stream, T, active, N, lr, beta1, beta2, eps, step, grad_scale, zero_grad);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

// ---- exchange of the touched rows only (several ranks) --------------------------------------------------------------------------------
// After `touched` has been summed over the ranks every rank derives the same ordered list of rows some rank touched, packs its own
// gradient rows of that list into a dense [count, stride] block (zeros for rows only other ranks touched), the block is all-reduced
// (count * stride floats instead of N * stride), and unpacked back into the table-shaped accumulators.
namespace sgn {
struct MarkedRow {
    const float* touched;
    __device__ __forceinline__ bool operator()(int32_t i) const { return touched[i] != 0.f; }
};

template <bool UNPACK>
__global__ void rows_pack_kernel(AdamTables T, const int32_t* __restrict__ list, const int32_t* __restrict__ count, float* __restrict__ packed, int stride)
{
    const int l = (int)(threadIdx.x & 7);
    const int64_t n = *count;
    for (int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; idx < n; idx += ((int64_t)gridDim.x * blockDim.x) >> 3) {
        const int64_t row = list[idx];
        float* dst = packed + idx * stride;
        int col = 0;
#pragma unroll
        for (int k = 0; k < ADAM_MAX_TABLES; k++) {
            if (k >= T.n) break;
            const int C = T.C[k];
            const int e0 = C <= 4 ? (l == (k & 7) ? 0 : C) : 4 * l;
            float* g = T.grad[k] + row * C;
            for (int c = e0; c < min(C, e0 + 4); c++) {
                if (UNPACK) g[c] = dst[col + c];
                else dst[col + c] = g[c];
            }
            col += C;
        }
    }
}
}  // namespace sgn

static size_t rows_select_bytes(int64_t N)
{
    size_t b = 0;
    cub::DeviceSelect::If(nullptr, b, cub::CountingInputIterator<int32_t>(0), (int32_t*)nullptr, (int32_t*)nullptr, (int)N, MarkedRow{nullptr});
    return b;
}

extern "C" int sgn_rows_union_bytes(int64_t N, size_t* bytes)
{
    SGN_CHECK_ARG(N >= 0 && N < (1ll << 31) && bytes, "sgn_rows_union_bytes: bad argument");
    *bytes = align_up(rows_select_bytes(N));
    return SGN_OK;
}

extern "C" int sgn_rows_union(const float* touched, int64_t N, int32_t* list, int32_t* count, void* workspace, size_t workspace_bytes, void* stream)
{
    SGN_CHECK_ARG(touched && list && count && N > 0 && N < (1ll << 31), "sgn_rows_union: bad argument");
    size_t need = rows_select_bytes(N);
    if (workspace_bytes < need || ((uintptr_t)workspace & 255)) { set_error("sgn_rows_union: workspace too small or misaligned (need %zu bytes)", need); return SGN_E_WORKSPACE; }
    // one-pass stream compaction (decoupled look-back): ascending row numbers, the same list on every rank
    ++g_launch_count;
    SGN_CUDA(cub::DeviceSelect::If(workspace, need, cub::CountingInputIterator<int32_t>(0), list, count, (int)N, MarkedRow{touched}, (cudaStream_t)stream));
    return SGN_OK;
}

extern "C" int sgn_rows_pack(int n_tables, float* const* tables, const int32_t* C, const int32_t* list, const int32_t* count, int64_t N, float* packed,
                             int stride, int unpack, void* stream)
{
    SGN_CHECK_ARG(n_tables > 0 && n_tables <= ADAM_MAX_TABLES && tables && C && list && count && packed && N >= 0, "sgn_rows_pack: bad argument");
    AdamTables T = {};
    T.n = n_tables;
    int cols = 0;
    for (int k = 0; k < n_tables; k++) {
        SGN_CHECK_ARG(C[k] > 0 && C[k] <= 32 && tables[k], "sgn_rows_pack: table %d: NULL or more than 32 columns", k);
        T.grad[k] = tables[k]; T.C[k] = C[k];
        cols += C[k];
    }
    SGN_CHECK_ARG(stride >= cols, "sgn_rows_pack: stride %d < %d columns", stride, cols);
    if (N == 0) return SGN_OK;
    const int nb = (int)std::min<int64_t>(cdiv(N * 8, 256), 148 * 8);
    if (unpack) launch(rows_pack_kernel<true>, nb, 256, 0, (cudaStream_t)stream, T, list, count, packed, stride);
    else launch(rows_pack_kernel<false>, nb, 256, 0, (cudaStream_t)stream, T, list, count, packed, stride);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

static int adam_tables(const char* what, AdamTables& T, int n_tables, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                       const int32_t* C)
{
    SGN_CHECK_ARG(n_tables > 0 && n_tables <= ADAM_MAX_TABLES && params && grads && exp_avg && exp_avg_sq && C, "%s: bad argument (at most %d tables)", what,
                  ADAM_MAX_TABLES);
    T.n = n_tables;
    for (int k = 0; k < n_tables; k++) {
        SGN_CHECK_ARG(C[k] > 0 && C[k] <= 32 && params[k] && grads[k] && exp_avg[k] && exp_avg_sq[k], "%s: table %d: NULL or more than 32 columns", what, k);
        T.param[k] = params[k]; T.grad[k] = grads[k]; T.m1[k] = exp_avg[k]; T.m2[k] = exp_avg_sq[k]; T.C[k] = C[k];
    }
    return SGN_OK;
}

extern "C" int sgn_adam_mark_rows(const int32_t* rows, int64_t n, float* touched, void* stream)
{
    SGN_CHECK_ARG(n >= 0 && (n == 0 || (rows && touched)), "sgn_adam_mark_rows: bad argument");
    if (n == 0) return SGN_OK;
    launch(mark_rows_kernel, cdiv(n, 256), 256, 0, (cudaStream_t)stream, rows, n, touched);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_adam_rows_list(int n_tables, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq, const int32_t* C,
                                  uint8_t* active, int32_t* active_list, int32_t* active_count, float* touched, int64_t N, float lr, float beta1, float beta2,
                                  float eps, const float* step, float grad_scale, int zero_grad, void* stream)
{
    AdamTables T = {};
    int rc = adam_tables("sgn_adam_rows_list", T, n_tables, params, grads, exp_avg, exp_avg_sq, C);
    if (rc) return rc;
    SGN_CHECK_ARG(active && active_list && active_count && touched && step && N >= 0 && N < (1ll << 31), "sgn_adam_rows_list: bad argument");
    if (N == 0) return SGN_OK;
    auto st = (cudaStream_t)stream;
    launch(adam_append_kernel, cdiv(N, 256), 256, 0, st, T, active, active_list, active_count, touched, N);
    launch(adam_list_kernel, (int)std::min<int64_t>(cdiv(N * 8, 256), 148 * 8), 256, 0, st, T, active_list, active_count, lr, beta1, beta2, eps, step, grad_scale,
           zero_grad);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_adam_dense_multi(int n_tensors, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                                    const int64_t* numel, float lr, float beta1, float beta2, float eps, const float* step, float grad_scale, int zero_grad,
                                    void* stream)
{
    SGN_CHECK_ARG(n_tensors >= 0 && params && grads && exp_avg && exp_avg_sq && numel && step, "sgn_adam_dense_multi: bad argument");
    for (int t0 = 0; t0 < n_tensors; t0 += ADAM_MAX_TENSORS) {
        AdamTensors T = {};
        const int n = std::min(ADAM_MAX_TENSORS, n_tensors - t0);
        int64_t most = 0;
        for (int k = 0; k < n; k++) {
            SGN_CHECK_ARG(numel[t0 + k] >= 0 && (numel[t0 + k] == 0 || (params[t0 + k] && grads[t0 + k] && exp_avg[t0 + k] && exp_avg_sq[t0 + k])),
                          "sgn_adam_dense_multi: tensor %d: NULL pointer", t0 + k);
            T.param[k] = params[t0 + k]; T.grad[k] = grads[t0 + k]; T.m1[k] = exp_avg[t0 + k]; T.m2[k] = exp_avg_sq[t0 + k]; T.n[k] = numel[t0 + k];
            most = std::max(most, numel[t0 + k]);
        }
        if (most == 0) continue;
        const int gx = (int)std::min<int64_t>(cdiv(most, 256), 32);
        launch(adam_dense_multi_kernel, dim3(gx, n), 256, 0, (cudaStream_t)stream, T, lr, beta1, beta2, eps, step, grad_scale, zero_grad);
    }
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}
