// common.cuh -- shared helpers for libsgnerf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sgnerf_b200.h"

namespace sgn {

void set_error(const char* fmt, ...);

#define SGN_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            ::sgn::set_error(__VA_ARGS__);       \
            return SGN_E_INVALID;                \
        }                                        \
    } while (0)

#define SGN_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            ::sgn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return SGN_E_CUDA;                                                                  \
        }                                                                                       \
    } while (0)

#define SGN_LAUNCH_CHECK() SGN_CUDA(cudaGetLastError())

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Bump allocator over a caller-supplied workspace.
struct Arena {
    char* base;
    size_t cap, off;
    Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
    template <typename T>
    T* take(size_t n)
    {
        size_t bytes = align_up(n * sizeof(T));
        T* r = (T*)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap; }
};

// Exclusive prefix sum of int32 (n up to 2^31), total written to out[n].  Three small kernels.
// partials must hold cdiv(n, SCAN_TILE) + 1 ints.
constexpr int SCAN_TILE = 2048;
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* partials, cudaStream_t st);
static inline size_t scan_partials_count(int64_t n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE + 1); }
// The same for two scans of one input in one pass: out_sum = scan of in, out_cnt = scan of (in > 0); partials: 2 * scan_partials_count(n).
int exclusive_scan_pair_i32(const int32_t* in, int32_t* out_sum, int32_t* out_cnt, int64_t n, int32_t* partials, cudaStream_t st);

// Every kernel launch of the library goes through launch(): it counts launches (sgn_launch_count) so a
// caller can state how many of OUR kernels ran inside a timed region.
extern unsigned long long g_launch_count;
template <typename... KArgs, typename... Args>
static inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    ++g_launch_count;
    kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
}

// ---- device helpers ----
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Voxel coordinate exactly as the reference computes it: (int)floor((p - shift) / vsize), IEEE fp32.
__device__ __forceinline__ int vox_coord(float p, float shift, float vsize)
{
    return (int)floorf(__fdiv_rn(__fsub_rn(p, shift), vsize));
}

// The same value as vox_coord for the hot loop of the march: multiply by the reciprocal and take the exact division only when the
// product lands too close to an integer to be sure.  With Q = a / vsize exactly, q = fl(a * fl(1 / vsize)) is within |Q| 2^-23 of Q and
// the reference's fl(a / vsize) within |Q| 2^-24, so whenever q is farther than |q| 4e-7 + 1e-6 (> 3x that) from an integer both floor
// to the same voxel.  At |q| of a few hundred the test fires for ~1e-4 of the positions, so a warp almost never takes the slow path.
__device__ __forceinline__ int vox_coord_fast(float p, float shift, float vsize, float rvsize)
{
    const float a = __fsub_rn(p, shift);
    const float q = a * rvsize;
    const float f = floorf(q);
    const float fr = q - f;
    const float thr = fmaf(fabsf(q), 4e-7f, 1e-6f);
    if (fr < thr || fr > 1.0f - thr || !(fabsf(q) < 8192.0f)) return (int)floorf(__fdiv_rn(a, vsize));
    return (int)f;
}

}  // namespace sgn
