// common.cu -- error text, version, and the int32 exclusive scan used by grid build and tuple compaction.
#include <stdarg.h>

#include "common.cuh"

namespace sgn {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------- exclusive scan ----------------
// k1: each block reduces SCAN_TILE inputs -> partials[b]
// k2: one block scans the partials in place (exclusive), writes grand total to partials[nb]
// k3: each block rescans its tile with the block offset and writes out[]; block 0 thread 0 writes out[n]
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = SCAN_TILE / SCAN_THREADS;  // 8

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /*[SCAN_THREADS/32 + 1]*/)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_THREADS / 32 ? smem[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        if (lane < SCAN_THREADS / 32) smem[lane] = si - s;
        if (lane == 31) smem[SCAN_THREADS / 32] = si;
    }
    __syncthreads();
    int r = inc - v + smem[w];
    *total = smem[SCAN_THREADS / 32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* partials)
{
    __shared__ int sm[SCAN_THREADS / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER_THREAD;
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++)
        if (base + i < n) s += in[base + i];
    int total;
    block_exclusive_scan(s, &total, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_partials_kernel(int32_t* partials, int nb)
{
    __shared__ int sm[SCAN_THREADS / 32 + 1];
    // a thread owns a run of consecutive partials: one block scan in all (a scan per SCAN_THREADS partials cost 25 us at 3600 of them)
    const int per = (nb + SCAN_THREADS - 1) / SCAN_THREADS;
    const int lo = min(nb, (int)threadIdx.x * per), hi = min(nb, lo + per);
    int s = 0;
    for (int i = lo; i < hi; i++) s += partials[i];
    int total;
    int e = block_exclusive_scan(s, &total, sm);
    for (int i = lo; i < hi; i++) { const int v = partials[i]; partials[i] = e; e += v; }
    if (threadIdx.x == 0) partials[nb] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                                                                   const int32_t* __restrict__ partials, int nb)
{
    __shared__ int sm[SCAN_THREADS / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD];
    int s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    int total;
    int e = block_exclusive_scan(s, &total, sm) + partials[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++) {
        if (base + i < n) out[base + i] = e;
        e += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = partials[nb];
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* partials, cudaStream_t st)
{
    int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (nb == 0) {
        SGN_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), st));
        return SGN_OK;
    }
    launch(scan_reduce_kernel, nb, SCAN_THREADS, 0, st, in, n, partials);
    launch(scan_partials_kernel, 1, SCAN_THREADS, 0, st, partials, nb);
    launch(scan_apply_kernel, nb, SCAN_THREADS, 0, st, in, out, n, partials, nb);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

// ---- pair scan: out_sum = exclusive scan of in, out_cnt = exclusive scan of (in > 0), one pass over `in` for both.
// partials holds 2 * (nb + 1) ints: the sums, then the counts.
__global__ void __launch_bounds__(SCAN_THREADS) scan_pair_reduce_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* partials, int nb)
{
    __shared__ int sm[SCAN_THREADS / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER_THREAD;
    int s = 0, c = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++)
        if (base + i < n) { const int v = in[base + i]; s += v; c += v > 0; }
    int total;
    block_exclusive_scan(s, &total, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = total;
    block_exclusive_scan(c, &total, sm);
    if (threadIdx.x == 0) partials[nb + 1 + blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_pair_partials_kernel(int32_t* partials, int nb)
{
    __shared__ int sm[SCAN_THREADS / 32 + 1];
    for (int which = 0; which < 2; which++) {
        int32_t* p = partials + which * (nb + 1);
        const int per = (nb + SCAN_THREADS - 1) / SCAN_THREADS;
        const int lo = min(nb, (int)threadIdx.x * per), hi = min(nb, lo + per);
        int s = 0;
        for (int i = lo; i < hi; i++) s += p[i];
        int total;
        int e = block_exclusive_scan(s, &total, sm);
        for (int i = lo; i < hi; i++) { const int v = p[i]; p[i] = e; e += v; }
        if (threadIdx.x == 0) p[nb] = total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_pair_apply_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out_sum,
                                                                        int32_t* __restrict__ out_cnt, int64_t n,
                                                                        const int32_t* __restrict__ partials, int nb)
{
    __shared__ int sm[SCAN_THREADS / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD];
    int s = 0, c = 0;
    // a thread's 8 values are 32 contiguous bytes: two 16-byte accesses per array when the run is whole and the arrays are aligned
    const bool vec = SCAN_PER_THREAD == 8 && base + SCAN_PER_THREAD <= n && ((((uintptr_t)in | (uintptr_t)out_sum | (uintptr_t)out_cnt) & 15) == 0);
    if (vec) {
        const int4 a = *(const int4*)(in + base), b = *(const int4*)(in + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_PER_THREAD; i++) v[i] = base + i < n ? in[base + i] : 0;
    }
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++) {
        s += v[i];
        c += v[i] > 0;
    }
    int total;
    int es = block_exclusive_scan(s, &total, sm) + partials[blockIdx.x];
    int ec = block_exclusive_scan(c, &total, sm) + partials[nb + 1 + blockIdx.x];
    int os[SCAN_PER_THREAD], oc[SCAN_PER_THREAD];
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++) {
        os[i] = es; oc[i] = ec;
        es += v[i];
        ec += v[i] > 0;
    }
    if (vec) {
        *(int4*)(out_sum + base) = make_int4(os[0], os[1], os[2], os[3]); *(int4*)(out_sum + base + 4) = make_int4(os[4], os[5], os[6], os[7]);
        *(int4*)(out_cnt + base) = make_int4(oc[0], oc[1], oc[2], oc[3]); *(int4*)(out_cnt + base + 4) = make_int4(oc[4], oc[5], oc[6], oc[7]);
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_PER_THREAD; i++)
            if (base + i < n) { out_sum[base + i] = os[i]; out_cnt[base + i] = oc[i]; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { out_sum[n] = partials[nb]; out_cnt[n] = partials[2 * nb + 1]; }
}

int exclusive_scan_pair_i32(const int32_t* in, int32_t* out_sum, int32_t* out_cnt, int64_t n, int32_t* partials, cudaStream_t st)
{
    int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (nb == 0) {
        SGN_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(int32_t), st));
        SGN_CUDA(cudaMemsetAsync(out_cnt, 0, sizeof(int32_t), st));
        return SGN_OK;
    }
    launch(scan_pair_reduce_kernel, nb, SCAN_THREADS, 0, st, in, n, partials, nb);
    launch(scan_pair_partials_kernel, 1, SCAN_THREADS, 0, st, partials, nb);
    launch(scan_pair_apply_kernel, nb, SCAN_THREADS, 0, st, in, out_sum, out_cnt, n, partials, nb);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

}  // namespace sgn

extern "C" const char* sgn_last_error(void) { return sgn::g_err; }
extern "C" int sgn_version(void) { return 100; }
extern "C" uint64_t sgn_launch_count(void) { return sgn::g_launch_count; }
