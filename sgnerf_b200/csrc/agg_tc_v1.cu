// agg_tc_v1.cu -- (first generation, kept for A/B runs with SGN_TC_V=1 until the ping-pong kernel of agg_tc.cu is validated)
// agg_tc.cu -- bf16 tensor-core aggregator forward: ONE fused persistent kernel for
//   gather + positional encoding -> per-neighbour MLP (block1 / block3) -> alpha -> K-weighted sums,
// built on tcgen05.mma (accumulators in TMEM), bulk-TMA weight streaming and mbarrier pipelines (sm_100a).
//
// Reference: PointAggregator.viewmlp (models/aggregators/point_aggregators.py:561-786) on the canonical
// non-semantic branch; the per-sample colour MLP that follows runs as separate launches (see bottom).
//
// Tile = 128 valid (sample, neighbour) tuples (rows), compacted in sample order.  Per CTA (1 per SM, persistent):
//
//   warps 0-3  epilogue : TMEM -> registers (tcgen05.ld) -> +bias, LeakyReLU -> bf16 -> next layer's A operand in
//                         shared memory (128B-swizzled K-major panels); last layer: alpha dot product and the
//                         K-weighted segmented sums over the rows of each sample -> F[S,256], sigma[S]
//   warps 4-7  gather   : for the NEXT tile, one thread per row: point tables -> [emb | PE(emb) | PE(dists)] (bf16)
//                         straight into the swizzled X0 operand panels, plus [colour | dir-view | dir.view]
//   warp  8    producer : cp.async.bulk (TMA, UBLKCP) of pre-swizzled 32 KB weight panels into a 2-stage ring
//   warp  9    MMA      : one thread issues tcgen05.mma 128x256x16 (bf16 in, fp32 accumulate in TMEM); two
//                         256-column accumulators alternate per layer so layer l+1's MMAs on K-panel p start as soon
//                         as the epilogue of layer l has written activation panel p
//
// Shared memory (bytes): X0 5 x 16 KB | activations 4 x 16 KB (aliased by the K-sum staging) | weight ring 2 x 32 KB |
// row metadata | mbarriers  = ~211 KB.  TMEM: 512 columns (2 accumulators of 128 lanes x 256 fp32 columns).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "agg_kernels.cuh"

namespace sgn {
namespace v1 {

constexpr int TC_ROWS = 128;
constexpr int TC_W = 256;                 // layer width == accumulator columns
constexpr int TC_C = 32, TC_F = 3, TC_FD = 5;
constexpr int TC_K0 = TC_C * (1 + 2 * TC_F) + 2 * TC_FD * 6;   // 284
constexpr int TC_MAX_LAYERS = 6;
constexpr int PANEL_A = TC_ROWS * 128;    // 16 KB: 128 rows x 64 bf16
constexpr int PANEL_B = TC_W * 128;       // 32 KB: 256 rows x 64 bf16
constexpr int X0_PANELS = 5, AM_PANELS = 4, B_STAGES = 2;
constexpr int E7_COL0 = 32;               // inside X0 panel 4: cols [32,48) tile parity 0, [48,64) parity 1

constexpr int OFF_X0 = 0;
constexpr int OFF_AM = OFF_X0 + X0_PANELS * PANEL_A;            // 81920
constexpr int OFF_B = OFF_AM + AM_PANELS * PANEL_A;             // 147456
constexpr int OFF_META = OFF_B + B_STAGES * PANEL_B;            // 212992
constexpr int META_BYTES = 2 * TC_ROWS * 16 + 2 * TC_ROWS * 4 + TC_ROWS * 4;   // per row {wc, keep, dest row, -} x2 | raw alpha x2 | sigma terms
constexpr int OFF_BIAS = OFF_META + META_BYTES;                 // [TC_MAX_LAYERS][256] biases + wa[256]
constexpr int BIAS_BYTES = (TC_MAX_LAYERS + 1) * TC_W * 4;
constexpr int OFF_BAR = OFF_BIAS + BIAS_BYTES;
constexpr int A_CHUNKS = TC_W / 32;           // activation hand-over granularity: 32 columns = two K-steps
constexpr int N_BARS = 2 * B_STAGES + 2 + A_CHUNKS + 4 + 2;
constexpr int OFF_TMEMPTR = OFF_BAR + N_BARS * 8;
constexpr int TC_SMEM = OFF_TMEMPTR + 16 + 1024;                // + slack for the 1024 B alignment of the base
constexpr int TC_EPI_WARPS = 8, TC_GATHER_WARP0 = 8, TC_PRODUCER_WARP = 12, TC_MMA_WARP = 13, TC_THREADS = 14 * 32;
static_assert(TC_SMEM <= 232448, "exceeds the 227 KB shared memory limit");

enum { LAYER_FROM_X0 = 0, LAYER_FROM_ACT = 1, LAYER_FROM_ACT_E7 = 2 };

struct TcParams {
    AggIn in;
    int K, SR;
    const int32_t* T_ptr; int T_max;
    const int32_t* tuple_src; const int32_t* tuple_start; const int32_t* sample_cidx;
    int n_tiles_cap;
    const float* loc_pers; const float* wc;
    const uint8_t* wpack;                  // pre-swizzled bf16 weight panels, all layers back to back
    int n_layers;
    int kind[TC_MAX_LAYERS];
    int first_panel[TC_MAX_LAYERS + 1];
    const float* bias[TC_MAX_LAYERS];
    const float* wa; const float* ba;
    float slope; int act_super;
    float* F; float* sigma;                // outputs [S_cap + tiles + 1][256] / [..]: per compact sample, then one carry row per tile, then a dummy row
    int S_cap;
    int dbg;                               // SGN_TC_DEBUG bitmask (profiling experiments only; results invalid when != 0)
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or the hint expires) instead of
// spinning in the issue slots of the epilogue / gather warps that share its scheduler
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
// the wait names the destination registers as in/out operands so no use of them can be scheduled above it
__device__ __forceinline__ void tc_wait_ld(uint32_t (&v)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 B apart (SBO), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16, A = B = bf16 (K-major), D = f32, M = 128, N = 256
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_W >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);

// last layer, operands swapped (D^T = W H^T): M = 128 features (two halves), N = 128 tuples
constexpr uint32_t TC_IDESC_T = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_ROWS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// byte offset of element (row, col) inside a 128B-swizzled K-major panel set (64 columns per panel)
__device__ __forceinline__ uint32_t sw_off(int row, int col)
{
    return (uint32_t)((col >> 6) * PANEL_A + row * 128 + ((((col >> 3) & 7) ^ (row & 7)) << 4) + (col & 7) * 2);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void stsf(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float ldsf(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int ldsi(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr)
{
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// One row of the K-sum walk, as a single asm so that none of its (chunk-invariant) bit tests can be hoisted into registers:
// row address = base + popc(heads & prefix) * ld_bytes; plain store if (st_plain & bit), reduction if (st_atom & bit).
__device__ __forceinline__ void ksum_store(const float* base, float v, uint32_t st_plain, uint32_t st_atom, uint32_t heads, uint32_t bit,
                                           uint32_t prefix, uint32_t ld_bytes)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b32 t, n;\n\t.reg .b64 off, a;\n\t"
        "and.b32 t, %2, %5;\n\tsetp.ne.u32 p, t, 0;\n\t"
        "and.b32 t, %3, %5;\n\tsetp.ne.u32 q, t, 0;\n\t"
        "and.b32 n, %4, %6;\n\tpopc.b32 n, n;\n\t"
        "mul.wide.u32 off, n, %7;\n\tadd.s64 a, %0, off;\n\t"
        "@p st.global.f32 [a], %1;\n\t@q red.global.add.f32 [a], %1;\n\t}"
        ::"l"(base), "f"(v), "r"(st_plain), "r"(st_atom), "r"(heads), "r"(bit), "r"(prefix), "r"(ld_bytes) : "memory");
}
__device__ __forceinline__ void st_global_f32(float* addr, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void red_shared_f32(uint32_t addr, float v) { asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void st_global_pred(float* addr, float v, uint32_t flag)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(addr), "f"(v), "r"(flag) : "memory");
}
__device__ __forceinline__ void red_global_pred(float* addr, float v, uint32_t flag)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.global.add.f32 [%0], %1;\n\t}" ::"l"(addr), "f"(v), "r"(flag) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(TC_THREADS, 1) agg_tuple_tc_kernel(const __grid_constant__ TcParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // per-row metadata of a tile, double buffered: float4 {wc, keep (0 at the first row of a sample, else 1), dest row (int bits), 0}
    float4* meta = (float4*)(smem + OFF_META);                     // [2][128]
    float* araw_sh = (float*)(smem + OFF_META + 2 * TC_ROWS * 16); // [2][128] raw alpha, summed over the 8 epilogue warps
    float* sig_sh = araw_sh + 2 * TC_ROWS;                         // [128] wc * act(alpha) per row
    const uint32_t bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    // barrier indices
    const int B_FULL = 0, B_EMPTY = B_STAGES, X0_FULL = 2 * B_STAGES, X0_EMPTY = X0_FULL + 1, A_FULL = X0_EMPTY + 1,
              D_FULL = A_FULL + A_CHUNKS, D_EMPTY = D_FULL + 2, META_FREE = D_EMPTY + 2;
    uint32_t* tmem_ptr_smem = (uint32_t*)(smem + OFF_TMEMPTR);

    const int T = min(*p.T_ptr, p.T_max);
    const int ntiles = (T + TC_ROWS - 1) / TC_ROWS;

    if (tid == 0) {
        for (int s = 0; s < B_STAGES; s++) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
        mbar_init(BAR(X0_FULL), 128); mbar_init(BAR(X0_EMPTY), 1);
        for (int i = 0; i < A_CHUNKS; i++) mbar_init(BAR(A_FULL + i), 128);
        for (int i = 0; i < 2; i++) { mbar_init(BAR(D_FULL + i), 1); mbar_init(BAR(D_EMPTY + i), TC_EPI_WARPS * 32); mbar_init(BAR(META_FREE + i), TC_EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < (p.n_layers + 1) * TC_W; i += blockDim.x) {
        const int l = i / TC_W, c = i - l * TC_W;
        ((float*)(smem + OFF_BIAS))[i] = l < p.n_layers ? p.bias[l][c] : p.wa[c];
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp < TC_EPI_WARPS) {
        // =========================================================== EPILOGUE: warp = (column half, TMEM lane quadrant)
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;
        // hidden layers: this warp owns the 32-column chunks half, half+2, half+4, half+6, so the two chunks of an activation
        // panel are produced side by side by the two halves and the MMA issuer can consume the panels in order

        uint32_t ph_dfull[2] = {0, 0};
        uint32_t lcount = 0;                                   // global layer counter -> accumulator buffer
        uint32_t tcount = 0;
        long long pf_wait = 0, pf_mid = 0, pf_last = 0, pf_sigma = 0, pf_t0 = 0;
        const bool prof = (p.dbg & 32) != 0;
        const uint32_t bias_a = sbase + OFF_BIAS;                              // [n_layers][256] f32, then wa[256]
        const uint32_t wa_a = bias_a + (uint32_t)(p.n_layers * TC_W) * 4u;
        const uint32_t lane_field = (uint32_t)(quad * 32) << 16;
        const uint32_t act_row = sbase + OFF_AM + row * 128;
        const float slope = p.slope;

        // bias + LeakyReLU on one 32-column chunk
        auto activate = [&](const uint32_t(&vv)[32], uint32_t bias_chunk, float(&h)[32]) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 bb = lds128f(bias_chunk + i * 4);
                const float x0 = __uint_as_float(vv[i]) + bb.x, x1 = __uint_as_float(vv[i + 1]) + bb.y;
                const float x2 = __uint_as_float(vv[i + 2]) + bb.z, x3 = __uint_as_float(vv[i + 3]) + bb.w;
                h[i] = fmaxf(x0, x0 * slope); h[i + 1] = fmaxf(x1, x1 * slope);
                h[i + 2] = fmaxf(x2, x2 * slope); h[i + 3] = fmaxf(x3, x3 * slope);
            }
        };
        // hidden layer: bf16 activations into the next layer's A operand (panel c/2, 16-byte chunks (c&1)*4 .. +3 of this row)
        auto mid_chunk = [&](int c, const uint32_t(&vv)[32], uint32_t bias_l) {
            if (!(p.dbg & 8)) {
                float h[32];
                activate(vv, bias_l + c * 128, h);
                const uint32_t rowbase = act_row + (c >> 1) * PANEL_A;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int ch = (c & 1) * 4 + q;
                    sts128(rowbase + ((ch ^ (row & 7)) << 4), pack_bf16(h[8 * q], h[8 * q + 1]), pack_bf16(h[8 * q + 2], h[8 * q + 3]),
                           pack_bf16(h[8 * q + 4], h[8 * q + 5]), pack_bf16(h[8 * q + 6], h[8 * q + 7]));
                }
            }
            fence_proxy_async();
            mbar_arrive(BAR(A_FULL + c));
        };

        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
            const int mb = tcount & 1;
            for (int l = 0; l < p.n_layers; l++, lcount++) {
                const int db = lcount & 1;
                const bool last = (l == p.n_layers - 1);
                if (prof) pf_t0 = clock64();
                mbar_wait(BAR(D_FULL + db), ph_dfull[db]);
                ph_dfull[db] ^= 1;
                tc_fence_after();
                if (prof) { const long long t1 = clock64(); pf_wait += t1 - pf_t0; pf_t0 = t1; }
                const uint32_t bias_l = bias_a + (uint32_t)(l * TC_W) * 4u;
                const uint32_t acc_addr = tmem_base + (uint32_t)(db * TC_W + half * 32) + lane_field;
                // software pipeline over this warp's 4 chunks: the TMEM load of the next chunk is in flight while one is processed
                uint32_t v0[32], v1[32];
                if (!last) {
                    tc_ld32_nowait(acc_addr, v0);
#pragma unroll 1
                    for (int cp = 0; cp < 2; cp++) {
                        tc_wait_ld(v0);
                        tc_ld32_nowait(acc_addr + (uint32_t)(cp * 128 + 64), v1);
                        mid_chunk(half + 4 * cp, v0, bias_l);
                        tc_wait_ld(v1);
                        if (cp == 0) tc_ld32_nowait(acc_addr + 128u, v0);
                        mid_chunk(half + 4 * cp + 2, v1, bias_l);
                    }
                    if (prof) { const long long t1 = clock64(); pf_mid += t1 - pf_t0; pf_t0 = t1; }
                    tc_fence_before();
                    mbar_arrive(BAR(D_EMPTY + db));
                } else {
                    // last layer, computed transposed (D^T = W H^T): TMEM lane = output feature, TMEM column = tuple row.  This thread
                    // owns feature f for all 128 rows of the tile, so the K-weighted sum over the consecutive rows of a sample is a
                    // sequential, branch-free walk in registers: acc = acc * keep + wc * h, stored to the sample's row of F after every
                    // tuple (later rows of the same sample overwrite earlier ones; 32 lanes = 32 features = one 128-byte store).
                    const int f = half * 128 + quad * 32 + lane;
                    const float bias_f = ldsf(bias_l + f * 4), wa_f = ldsf(wa_a + f * 4);
                    const uint32_t meta_a = sbase + OFF_META + (uint32_t)(mb * TC_ROWS) * 16u;
                    const uint32_t araw_a = sbase + OFF_META + 2 * TC_ROWS * 16 + (uint32_t)(mb * TC_ROWS) * 4u;
                    const uint32_t accT = tmem_base + (uint32_t)(db * TC_W + half * 128) + lane_field;
                    float* fcol = p.F + f;
                    float acc = 0.f;
                    auto last_chunk = [&](int cc, uint32_t(&vv)[32]) {
                        if (p.dbg & 24) return;
                        float pa[32];
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const float x = __uint_as_float(vv[i]) + bias_f;
                            const float h = fmaxf(x, x * slope);
                            const float4 m = lds128f(meta_a + (uint32_t)(cc * 32 + i) * 16u);          // warp-uniform address
                            acc = fmaf(acc, m.y, h * m.x);
                            st_global_f32(fcol + (size_t)(uint32_t)__float_as_int(m.z) * TC_W, acc);
                            pa[i] = h * wa_f;
                        }
                        // alpha: sum over the 32 features of this warp for each of the 32 rows (transpose-reduce, 31 shuffles),
                        // lane i ends up with the partial of row cc*32 + i; the 8 warps meet in shared memory
#pragma unroll
                        for (int sft = 16; sft >= 1; sft >>= 1) {
                            const bool up = (lane & sft) != 0;
#pragma unroll
                            for (int i = 0; i < sft; i++) {
                                const float send = up ? pa[i] : pa[i + sft];
                                const float keepv = up ? pa[i + sft] : pa[i];
                                pa[i] = keepv + __shfl_xor_sync(0xffffffffu, send, sft);
                            }
                        }
                        red_shared_f32(araw_a + (uint32_t)(cc * 32 + lane) * 4u, pa[0]);
                    };
                    tc_ld32_nowait(accT, v0);
#pragma unroll 1
                    for (int cp = 0; cp < 2; cp++) {
                        tc_wait_ld(v0);
                        tc_ld32_nowait(accT + (uint32_t)(cp * 64 + 32), v1);
                        last_chunk(2 * cp, v0);
                        tc_wait_ld(v1);
                        if (cp == 0) tc_ld32_nowait(accT + 64u, v0);
                        last_chunk(2 * cp + 1, v1);
                    }
                    if (prof) { const long long t1 = clock64(); pf_last += t1 - pf_t0; pf_t0 = t1; }
                    tc_fence_before();
                    mbar_arrive(BAR(D_EMPTY + db));
                    // sigma = sum over the rows of a sample of wc * act(alpha): the lower four warps take one row each
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (half == 0) {
                        const int r = quad * 32 + lane;
                        const float4 m = meta[mb * TC_ROWS + r];
                        const float a = araw_sh[mb * TC_ROWS + r] + p.ba[0];
                        const float act = p.act_super ? softplus1(a - 1.0f) : fmaxf(a, 0.f);
                        sig_sh[r] = (p.dbg & 24) ? 0.f : act * m.x;
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                        if (m.y == 0.f) {                        // first row of a sample (within this tile)
                            float sum = sig_sh[r];
                            for (int q = r + 1; q < TC_ROWS && meta[mb * TC_ROWS + q].y != 0.f; q++) sum += sig_sh[q];
                            p.sigma[(uint32_t)__float_as_int(m.z)] = sum;
                        }
                        __syncwarp();
                    }
                    mbar_arrive(BAR(META_FREE + mb));
                    if (prof) { const long long t1 = clock64(); pf_sigma += t1 - pf_t0; pf_t0 = t1; }
                }
            }
        }
        if (prof && blockIdx.x == 0 && lane == 0)
            printf("epi warp %d: tiles %u wait %lld mid(%d layers) %lld last %lld sigma %lld (cycles/tile)\n", warp, tcount,
                   pf_wait / max(tcount, 1u), p.n_layers - 1, pf_mid / max(tcount, 1u), pf_last / max(tcount, 1u), pf_sigma / max(tcount, 1u));
    } else if (warp < TC_PRODUCER_WARP) {
        // =========================================================== GATHER (one thread per row) for this CTA's tiles, one ahead
        const int row = tid - TC_GATHER_WARP0 * 32;
        uint32_t ph_x0empty = 1, ph_meta[2] = {1, 1};
        uint32_t tcount = 0;
        long long gf_load = 0, gf_wait = 0, gf_write = 0, gf_t0 = 0;
        const bool prof = (p.dbg & 32) != 0;
        const float* Rm = p.in.camrot;
        const float r00 = Rm[0], r01 = Rm[1], r02 = Rm[2], r10 = Rm[3], r11 = Rm[4], r12 = Rm[5], r20 = Rm[6], r21 = Rm[7], r22 = Rm[8];
        const float cpx = p.in.campos[0], cpy = p.in.campos[1], cpz = p.in.campos[2];
        const uint32_t x0 = sbase + OFF_X0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
            const int mb = tcount & 1;
            if (prof) gf_t0 = clock64();
            const int64_t j = (int64_t)tile * TC_ROWS + row;
            const bool live = j < T;
            float wcv = 0.f, keepv = 0.f; int drow = p.S_cap + p.n_tiles_cap;       // dead rows: weight 0, dummy destination row
            float emb[TC_C];
            float dist[6];
            float e7[8];
#pragma unroll
            for (int i = 0; i < TC_C; i++) emb[i] = 0.f;
#pragma unroll
            for (int i = 0; i < 6; i++) dist[i] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) e7[i] = 0.f;
            if (live && !(p.dbg & 4)) {
                const int flat = p.tuple_src[j];
                const int64_t s = flat / p.K;
                const int64_t r = s / p.SR;
                const int64_t pt = p.in.pidx[flat];
                wcv = p.wc[flat];
                const int st = p.tuple_start[s];
                // a sample continued from the previous tile accumulates into this tile's carry row (added back by the colour kernel)
                drow = st < tile * TC_ROWS ? p.S_cap + tile : p.sample_cidx[s];
                keepv = (j == st || row == 0) ? 0.f : 1.f;
                const float4* ep = (const float4*)(p.in.tab.embedding + pt * TC_C);
#pragma unroll
                for (int i = 0; i < TC_C / 4; i++) {
                    const float4 e = __ldg(ep + i);
                    emb[4 * i] = e.x; emb[4 * i + 1] = e.y; emb[4 * i + 2] = e.z; emb[4 * i + 3] = e.w;
                }
                const float px = p.in.tab.xyz[3 * pt], py = p.in.tab.xyz[3 * pt + 1], pz = p.in.tab.xyz[3 * pt + 2];
                dist[0] = px - p.in.loc_w[3 * s]; dist[1] = py - p.in.loc_w[3 * s + 1]; dist[2] = pz - p.in.loc_w[3 * s + 2];
                const float sx = px - cpx, sy = py - cpy, sz = pz - cpz;
                const float c0 = sx * r00 + sy * r10 + sz * r20, c1 = sx * r01 + sy * r11 + sz * r21, c2 = sx * r02 + sy * r12 + sz * r22;
                const float xp = c0 / c2, yp = c1 / c2;
                const float lxp = p.loc_pers[3 * s], lyp = p.loc_pers[3 * s + 1], lzp = p.loc_pers[3 * s + 2];
                dist[3] = xp * c2 - lxp * lzp; dist[4] = yp * c2 - lyp * lzp; dist[5] = c2 - lzp;
                const float vx = p.in.raydir[3 * r], vy = p.in.raydir[3 * r + 1], vz = p.in.raydir[3 * r + 2];
                const float dx = p.in.tab.dir[3 * pt], dy = p.in.tab.dir[3 * pt + 1], dz = p.in.tab.dir[3 * pt + 2];
                e7[0] = p.in.tab.color[3 * pt]; e7[1] = p.in.tab.color[3 * pt + 1]; e7[2] = p.in.tab.color[3 * pt + 2];
                e7[3] = dx - vx; e7[4] = dy - vy; e7[5] = dz - vz; e7[6] = dx * vx + dy * vy + dz * vz;
            }
            // the global loads above are in flight while the previous tile still owns the X0 panels
            if (prof) { const long long t1 = clock64(); gf_load += t1 - gf_t0; gf_t0 = t1; }
            mbar_wait(BAR(X0_EMPTY), ph_x0empty); ph_x0empty ^= 1;
            mbar_wait(BAR(META_FREE + mb), ph_meta[mb]); ph_meta[mb] ^= 1;
            if (prof) { const long long t1 = clock64(); gf_wait += t1 - gf_t0; gf_t0 = t1; }
            // cols [0,32): embedding
#pragma unroll
            for (int q = 0; q < 4; q++)
                sts128(x0 + sw_off(row, 8 * q), pack_bf16(emb[8 * q], emb[8 * q + 1]), pack_bf16(emb[8 * q + 2], emb[8 * q + 3]),
                       pack_bf16(emb[8 * q + 4], emb[8 * q + 5]), pack_bf16(emb[8 * q + 6], emb[8 * q + 7]));
            // cols 32 + 2*(c*F + f) + {0: sin, 1: cos}: base angle by sincosf, octaves by the double-angle recurrence
            if (!(p.dbg & 2))
#pragma unroll
            for (int c = 0; c < TC_C; c++) {
                float sn, cs_;
                __sincosf(emb[c], &sn, &cs_);
#pragma unroll
                for (int f = 0; f < TC_F; f++) {
                    sts32(x0 + sw_off(row, TC_C + 2 * (c * TC_F + f)), pack_bf16(sn, cs_));
                    const float s2 = 2.0f * sn * cs_, c2 = 1.0f - 2.0f * sn * sn;
                    sn = s2; cs_ = c2;
                }
            }
            constexpr int DB = TC_C * (1 + 2 * TC_F);     // 224
#pragma unroll
            for (int d = 0; d < 6; d++) {
                float sn, cs_;
                __sincosf(dist[d], &sn, &cs_);
#pragma unroll
                for (int f = 0; f < TC_FD; f++) {
                    sts32(x0 + sw_off(row, DB + 2 * (d * TC_FD + f)), pack_bf16(sn, cs_));
                    const float s2 = 2.0f * sn * cs_, c2 = 1.0f - 2.0f * sn * sn;
                    sn = s2; cs_ = c2;
                }
            }
            sts32(x0 + sw_off(row, TC_K0), 0u); sts32(x0 + sw_off(row, TC_K0 + 2), 0u);      // cols 284..287 = 0
            // E7 slot of this tile parity: cols 256 + 32 + 16*mb .. +15 (panel 4)
            {
                const int col = 4 * 64 + E7_COL0 + 16 * mb;
                sts128(x0 + sw_off(row, col), pack_bf16(e7[0], e7[1]), pack_bf16(e7[2], e7[3]), pack_bf16(e7[4], e7[5]), pack_bf16(e7[6], 0.f));
                sts128(x0 + sw_off(row, col + 8), 0u, 0u, 0u, 0u);
            }
            meta[mb * TC_ROWS + row] = make_float4(wcv, keepv, __int_as_float(drow), 0.f);
            araw_sh[mb * TC_ROWS + row] = 0.f;
            fence_proxy_async();
            mbar_arrive(BAR(X0_FULL));
            if (prof) { const long long t1 = clock64(); gf_write += t1 - gf_t0; gf_t0 = t1; }
        }
        if (prof && blockIdx.x == 0 && lane == 0)
            printf("gather warp %d: tiles %u issue-loads %lld wait-slot %lld expand+write %lld (cycles/tile)\n", warp, tcount, gf_load / max(tcount, 1u), gf_wait / max(tcount, 1u), gf_write / max(tcount, 1u));
    } else if (warp == TC_PRODUCER_WARP) {
        // =========================================================== PRODUCER: weight panels through the ring
        if (lane == 0) {
            uint32_t ph_empty[B_STAGES];
            for (int s = 0; s < B_STAGES; s++) ph_empty[s] = 1;
            uint32_t n = 0;
            const int total_panels = p.first_panel[p.n_layers];
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int pi = 0; pi < total_panels; pi++, n++) {
                    const int s = n % B_STAGES;
                    mbar_wait(BAR(B_EMPTY + s), ph_empty[s]); ph_empty[s] ^= 1;
                    if (p.dbg & 1) { mbar_arrive(BAR(B_FULL + s)); continue; }
                    mbar_expect_tx(BAR(B_FULL + s), PANEL_B);
                    bulk_g2s(sbase + OFF_B + s * PANEL_B, p.wpack + (size_t)pi * PANEL_B, PANEL_B, BAR(B_FULL + s));
                }
            }
        }
    } else {
        // =========================================================== MMA issuer
        if (lane == 0) {
            uint32_t ph_full[B_STAGES];
            for (int s = 0; s < B_STAGES; s++) ph_full[s] = 0;
            uint32_t ph_x0full = 0, ph_afull = 0, ph_dempty[2] = {1, 1};   // ph_afull: one phase bit per activation chunk
            uint32_t n = 0, lcount = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
                const int mb = tcount & 1;
                for (int l = 0; l < p.n_layers; l++, lcount++) {
                    const int db = lcount & 1;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(db * TC_W);
                    mbar_wait(BAR(D_EMPTY + db), ph_dempty[db]); ph_dempty[db] ^= 1;
                    const int np = p.first_panel[l + 1] - p.first_panel[l];
                    const int kind = p.kind[l];
                    if (kind == LAYER_FROM_X0) { mbar_wait(BAR(X0_FULL), ph_x0full); ph_x0full ^= 1; }
                    uint32_t acc = 0;
                    const bool swapped = (l == p.n_layers - 1);          // last layer: D^T = W H^T (see the epilogue)
                    auto issue = [&](uint32_t a_addr, uint32_t b_addr, int k0, int k1) {
                        for (int k = k0; k < k1; k++) {
                            if (!swapped) {
                                tc_mma(d_tmem, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), TC_IDESC, acc);
                            } else {
                                tc_mma(d_tmem, umma_desc(b_addr + k * 32), umma_desc(a_addr + k * 32), TC_IDESC_T, acc);
                                tc_mma(d_tmem + 128u, umma_desc(b_addr + 128 * 128 + k * 32), umma_desc(a_addr + k * 32), TC_IDESC_T, acc);
                            }
                            acc = 1;
                        }
                    };
                    for (int kp = 0; kp < np; kp++, n++) {
                        const int s = n % B_STAGES;
                        const uint32_t b_addr = sbase + OFF_B + s * PANEL_B;
                        if (kind != LAYER_FROM_X0 && kp < AM_PANELS) {
                            // activation panel kp arrives as two 32-column chunks; each is two K-steps
                            const uint32_t a_addr = sbase + OFF_AM + kp * PANEL_A;
                            for (int hc = 0; hc < 2; hc++) {
                                const int c = 2 * kp + hc;
                                mbar_wait(BAR(A_FULL + c), (ph_afull >> c) & 1u); ph_afull ^= 1u << c;
                                if (hc == 0) { mbar_wait(BAR(B_FULL + s), ph_full[s]); ph_full[s] ^= 1; }
                                tc_fence_after();
                                issue(a_addr, b_addr, 2 * hc, 2 * hc + 2);
                            }
                        } else {
                            uint32_t a_addr;
                            int ksteps = 4;
                            if (kind == LAYER_FROM_X0) {
                                a_addr = sbase + OFF_X0 + kp * PANEL_A;
                                if (kp == 4) ksteps = 2;                      // cols 256..287
                            } else {                                          // [colour | dir - view | dir.view] K-step of block3.0
                                a_addr = sbase + OFF_X0 + 4 * PANEL_A + (E7_COL0 + 16 * mb) * 2;
                                ksteps = 1;
                            }
                            mbar_wait(BAR(B_FULL + s), ph_full[s]); ph_full[s] ^= 1;
                            tc_fence_after();
                            issue(a_addr, b_addr, 0, ksteps);
                        }
                        tc_commit(BAR(B_EMPTY + s));
                    }
                    if (kind == LAYER_FROM_X0) tc_commit(BAR(X0_EMPTY));
                    tc_commit(BAR(D_FULL + db));
                }
            }
        }
    }
    __syncthreads();
    if (warp == TC_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}


// ================================================================================================ colour branch
// Per-sample colour MLP on tensor cores (point_aggregators.py:298-309 raw2out_color, :771-786): one persistent CTA per SM,
// tile = 128 compact samples.  All hidden-layer weights stay resident in shared memory (bf16, 128B-swizzled K-major
// panels); the A operand of the first layer is streamed: loader warps read the fp32 K-sums F[c, 0:256] written by the
// per-neighbour kernel, round to bf16 and write 16 KB swizzled panels into a 2-stage ring, the fifth panel holds the
// view-direction encoding.  The 128-wide activations live in place in two panels; the last Linear (128 -> 3), the
// sigmoid and the (sigma, r, g, b) store are fused into the last epilogue.
//   warps 0-3 epilogue | warps 4-7 loaders | warp 8 MMA issuer (and the one-off weight load)
constexpr int CW = 128;                                   // colour hidden width
constexpr int C_PANEL = CW * 128;                         // 16 KB: 128 rows x 64 bf16 (A and B panels alike)
constexpr int C_K0_PANELS = 5, C_RING = 2, C_MAX_HIDDEN = 3;
constexpr int COFF_W = 0;                                                   // resident weights: 5 + 2 + 2 panels
constexpr int C_W_PANELS = C_K0_PANELS + 2 * (C_MAX_HIDDEN - 1);
constexpr int COFF_RING = COFF_W + C_W_PANELS * C_PANEL;
constexpr int COFF_ACT = COFF_RING + C_RING * C_PANEL;
constexpr int COFF_BIAS = COFF_ACT + 2 * C_PANEL;                           // [3][128] hidden biases
constexpr int COFF_WL = COFF_BIAS + C_MAX_HIDDEN * CW * 4;                  // [3][128] last Linear + its bias [4]
constexpr int COFF_CARRY = COFF_WL + 3 * CW * 4 + 16;                       // [2][128] carry row of each sample of the tile (-1: none)
constexpr int COFF_BAR = COFF_CARRY + 2 * TC_ROWS * 4;
constexpr int C_NBARS = 1 + 2 * C_RING + 2 + 4;
constexpr int COFF_TMEMPTR = COFF_BAR + C_NBARS * 8;
constexpr int C_SMEM = COFF_TMEMPTR + 16 + 1024;
static_assert(C_SMEM <= 232448, "colour kernel exceeds the 227 KB shared memory limit");
constexpr uint32_t C_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CW >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);

struct ColParams {
    const int32_t* S_ptr; int S_max;
    const int32_t* csample;            // compact sample -> sample
    const float* F;                    // [S_cap + tiles + 1][256] K-sums per compact sample, then the per-tile carry rows
    const float* sigma;                // same row indexing
    const int32_t* tuple_start; const int32_t* nvalid; int S_cap;
    const float* raydir; int SR;
    const uint8_t* wpack;              // packed hidden-layer weights, C_PANEL each, layer after layer
    int n_hidden;                      // colour layers followed by an activation (1..3)
    int fv;                            // num_viewdir_freqs
    const float* bias[C_MAX_HIDDEN];
    const float* wl; const float* bl;  // last Linear [3,128], [3]
    float slope; int act_super;
    float* decoded;                    // [S,4]
    int dbg;
};

__global__ void __launch_bounds__(288, 1) agg_color_tc_kernel(const __grid_constant__ ColParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = sbase + COFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    const int W_FULL = 0, R_FULL = 1, R_EMPTY = R_FULL + C_RING, A_FULL = R_EMPTY + C_RING, D_FULL = A_FULL + 2, D_EMPTY = D_FULL + 2;
    uint32_t* tmem_ptr_smem = (uint32_t*)(smem + COFF_TMEMPTR);
    float* s_bias = (float*)(smem + COFF_BIAS);
    float* s_wl = (float*)(smem + COFF_WL);

    const int Sv = min(*p.S_ptr, p.S_max);
    const int ntiles = (Sv + TC_ROWS - 1) / TC_ROWS;
    const int n_wpanels = C_K0_PANELS + 2 * (p.n_hidden - 1);

    if (tid == 0) {
        mbar_init(BAR(W_FULL), 1);
        for (int s = 0; s < C_RING; s++) { mbar_init(BAR(R_FULL + s), 128); mbar_init(BAR(R_EMPTY + s), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(BAR(A_FULL + i), 128); mbar_init(BAR(D_FULL + i), 1); mbar_init(BAR(D_EMPTY + i), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < p.n_hidden * CW; i += blockDim.x) s_bias[i] = p.bias[i / CW][i % CW];
    for (int i = tid; i < 3 * CW; i += blockDim.x) s_wl[i] = p.wl[i];
    if (tid < 3) s_wl[3 * CW + tid] = p.bl[tid];
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp < 4) {
        // =========================================================== EPILOGUE
        const int row = tid;
        uint32_t ph_dfull[2] = {0, 0};
        uint32_t lcount = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t c = (int64_t)tile * TC_ROWS + row;
            for (int l = 0; l < p.n_hidden; l++, lcount++) {
                const int db = lcount & 1;
                const bool last = (l == p.n_hidden - 1);
                mbar_wait(BAR(D_FULL + db), ph_dfull[db]);
                ph_dfull[db] ^= 1;
                tc_fence_after();
                float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 1
                for (int ch = 0; ch < CW / 32; ch++) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + (uint32_t)(db * CW + ch * 32) + ((uint32_t)(warp * 32) << 16), v);
                    float h[32];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 bb = *(const float4*)(s_bias + l * CW + ch * 32 + i);
                        float x0 = __uint_as_float(v[i]) + bb.x, x1 = __uint_as_float(v[i + 1]) + bb.y;
                        float x2 = __uint_as_float(v[i + 2]) + bb.z, x3 = __uint_as_float(v[i + 3]) + bb.w;
                        h[i] = fmaxf(x0, x0 * p.slope); h[i + 1] = fmaxf(x1, x1 * p.slope);
                        h[i + 2] = fmaxf(x2, x2 * p.slope); h[i + 3] = fmaxf(x3, x3 * p.slope);
                    }
                    if (!last) {
                        const uint32_t rowbase = sbase + COFF_ACT + (ch >> 1) * C_PANEL + row * 128;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int k = (ch & 1) * 4 + q;
                            sts128(rowbase + ((k ^ (row & 7)) << 4), pack_bf16(h[8 * q], h[8 * q + 1]), pack_bf16(h[8 * q + 2], h[8 * q + 3]),
                                   pack_bf16(h[8 * q + 4], h[8 * q + 5]), pack_bf16(h[8 * q + 6], h[8 * q + 7]));
                        }
                        if (ch & 1) {
                            fence_proxy_async();
                            mbar_arrive(BAR(A_FULL + (ch >> 1)));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 w0 = *(const float4*)(s_wl + ch * 32 + i);
                            const float4 w1 = *(const float4*)(s_wl + CW + ch * 32 + i);
                            const float4 w2 = *(const float4*)(s_wl + 2 * CW + ch * 32 + i);
                            o0 = fmaf(h[i], w0.x, o0); o0 = fmaf(h[i + 1], w0.y, o0); o0 = fmaf(h[i + 2], w0.z, o0); o0 = fmaf(h[i + 3], w0.w, o0);
                            o1 = fmaf(h[i], w1.x, o1); o1 = fmaf(h[i + 1], w1.y, o1); o1 = fmaf(h[i + 2], w1.z, o1); o1 = fmaf(h[i + 3], w1.w, o1);
                            o2 = fmaf(h[i], w2.x, o2); o2 = fmaf(h[i + 1], w2.y, o2); o2 = fmaf(h[i + 2], w2.z, o2); o2 = fmaf(h[i + 3], w2.w, o2);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(BAR(D_EMPTY + db));
                if (last && c < Sv && !(p.dbg & 256)) {
                    const float s0 = 1.0f / (1.0f + __expf(-(o0 + s_wl[3 * CW]))), s1 = 1.0f / (1.0f + __expf(-(o1 + s_wl[3 * CW + 1]))),
                                s2 = 1.0f / (1.0f + __expf(-(o2 + s_wl[3 * CW + 2])));
                    const float m = p.act_super ? 1.002f : 1.0f, o = p.act_super ? 0.001f : 0.0f;
                    const int sidx = p.csample[c];
                    const int st = p.tuple_start[sidx], t2 = (st + p.nvalid[sidx] - 1) >> 7;
                    float sg = p.sigma[c];
                    if ((st >> 7) != t2) sg += p.sigma[p.S_cap + t2];
                    ((float4*)p.decoded)[sidx] = make_float4(sg, s0 * m - o, s1 * m - o, s2 * m - o);
                }
            }
        }
    } else if (warp < 8) {
        // =========================================================== LOADERS: F (fp32, global) -> bf16 ring panels
        const int lt = tid - 128;
        const int sub = lt & 15, rgrp = lt >> 4;              // 16 threads per row (float4 each), 8 rows per pass
        uint32_t ph_empty[C_RING];
        for (int s = 0; s < C_RING; s++) ph_empty[s] = 1;
        uint32_t n = 0;
        int32_t* s_carry = (int32_t*)(smem + COFF_CARRY);
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
            const int64_t c0 = (int64_t)tile * TC_ROWS;
            // samples whose tuples straddle two tiles of the per-neighbour kernel left their second part in that tile's carry row
            int32_t* carry = s_carry + (tcount & 1) * TC_ROWS;
            {
                int cr = -1;
                if (c0 + lt < Sv) {
                    const int sidx = p.csample[c0 + lt];
                    const int st = p.tuple_start[sidx], t2 = (st + p.nvalid[sidx] - 1) >> 7;
                    if ((st >> 7) != t2) cr = p.S_cap + t2;
                }
                carry[lt] = cr;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            for (int kp = 0; kp < C_K0_PANELS; kp++, n++) {
                const int s = n % C_RING;
                const uint32_t base = sbase + COFF_RING + s * C_PANEL;
                if (kp < 4) {
                    float4 f[16];
#pragma unroll
                    for (int pass = 0; pass < 16; pass++) {
                        const int r = pass * 8 + rgrp;
                        f[pass] = (c0 + r < Sv && !(p.dbg & 128)) ? __ldg((const float4*)(p.F + (size_t)(c0 + r) * TC_W + kp * 64 + sub * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const int cr = carry[r];
                        if (cr >= 0) {
                            const float4 g = __ldg((const float4*)(p.F + (size_t)cr * TC_W + kp * 64 + sub * 4));
                            f[pass].x += g.x; f[pass].y += g.y; f[pass].z += g.z; f[pass].w += g.w;
                        }
                    }
                    mbar_wait(BAR(R_EMPTY + s), ph_empty[s]); ph_empty[s] ^= 1;
#pragma unroll
                    for (int pass = 0; pass < 16; pass++) {
                        const int r = pass * 8 + rgrp;
                        const uint32_t a = base + r * 128 + (((sub >> 1) ^ (r & 7)) << 4) + (sub & 1) * 8;
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(pack_bf16(f[pass].x, f[pass].y)), "r"(pack_bf16(f[pass].z, f[pass].w)) : "memory");
                    }
                } else {
                    // view-direction encoding (ori=True, first three stripped): sin(v_d 2^f) d-major, then the cosines; cols [6 fv, 32) = 0
                    const int r = lt;
                    float vals[32];
#pragma unroll
                    for (int i = 0; i < 32; i++) vals[i] = 0.f;
                    if (c0 + r < Sv && !(p.dbg & 64)) {
                        const int64_t ray = p.csample[c0 + r] / p.SR;
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            if (i < 3 * p.fv) {
                                const int dd = i / p.fv, f = i - dd * p.fv;
                                float sn, cs;
                                sincosf(p.raydir[3 * ray + dd] * exp2f((float)f), &sn, &cs);
#pragma unroll
                                for (int j = 0; j < 32; j++) {
                                    if (j == i) vals[j] = sn;
                                    if (j == i + 3 * p.fv) vals[j] = cs;
                                }
                            }
                        }
                    }
                    mbar_wait(BAR(R_EMPTY + s), ph_empty[s]); ph_empty[s] ^= 1;
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        sts128(base + r * 128 + ((q ^ (r & 7)) << 4), pack_bf16(vals[8 * q], vals[8 * q + 1]), pack_bf16(vals[8 * q + 2], vals[8 * q + 3]),
                               pack_bf16(vals[8 * q + 4], vals[8 * q + 5]), pack_bf16(vals[8 * q + 6], vals[8 * q + 7]));
                }
                fence_proxy_async();
                mbar_arrive(BAR(R_FULL + s));
            }
        }
    } else {
        // =========================================================== MMA issuer (+ one-off resident weight load)
        if (lane == 0) {
            mbar_expect_tx(BAR(W_FULL), (uint32_t)n_wpanels * C_PANEL);
            for (int i = 0; i < n_wpanels; i++) bulk_g2s(sbase + COFF_W + i * C_PANEL, p.wpack + (size_t)i * C_PANEL, C_PANEL, BAR(W_FULL));
            mbar_wait(BAR(W_FULL), 0);
            uint32_t ph_full[C_RING];
            for (int s = 0; s < C_RING; s++) ph_full[s] = 0;
            uint32_t ph_afull[2] = {0, 0}, ph_dempty[2] = {1, 1};
            uint32_t n = 0, lcount = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < p.n_hidden; l++, lcount++) {
                    const int db = lcount & 1;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(db * CW);
                    mbar_wait(BAR(D_EMPTY + db), ph_dempty[db]); ph_dempty[db] ^= 1;
                    uint32_t acc = 0;
                    if (l == 0) {
                        for (int kp = 0; kp < C_K0_PANELS; kp++, n++) {
                            const int s = n % C_RING;
                            mbar_wait(BAR(R_FULL + s), ph_full[s]); ph_full[s] ^= 1;
                            tc_fence_after();
                            const uint32_t a_addr = sbase + COFF_RING + s * C_PANEL, b_addr = sbase + COFF_W + kp * C_PANEL;
                            const int ksteps = kp < 4 ? 4 : 2;
                            for (int k = 0; k < ksteps; k++) {
                                tc_mma(d_tmem, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), C_IDESC, acc);
                                acc = 1;
                            }
                            tc_commit(BAR(R_EMPTY + s));
                        }
                    } else {
                        for (int kp = 0; kp < 2; kp++) {
                            mbar_wait(BAR(A_FULL + kp), ph_afull[kp]); ph_afull[kp] ^= 1;
                            tc_fence_after();
                            const uint32_t a_addr = sbase + COFF_ACT + kp * C_PANEL;
                            const uint32_t b_addr = sbase + COFF_W + (C_K0_PANELS + 2 * (l - 1) + kp) * C_PANEL;
                            for (int k = 0; k < 4; k++) {
                                tc_mma(d_tmem, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), C_IDESC, acc);
                                acc = 1;
                            }
                        }
                    }
                    tc_commit(BAR(D_FULL + db));
                }
            }
        }
    }
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// ------------------------------------------------------------------------------------------------ small kernels
// torch Linear weight [N=256, K_in] fp32 -> bf16 panels of 64 K-columns in the 128B-swizzled smem image (32 KB each)
__global__ void tc_pack_weight_kernel(const float* __restrict__ W, int Nrows, int Kin, int npanels, uint8_t* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk (8 bf16)
    if (i >= npanels * Nrows * 8) return;
    const int ch = i & 7, n = (i >> 3) % Nrows, pnl = i / (Nrows * 8);
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int k = pnl * 64 + ch * 8 + 2 * e;
        const float a = k < Kin ? W[(size_t)n * Kin + k] : 0.f, b = (k + 1) < Kin ? W[(size_t)n * Kin + k + 1] : 0.f;
        w[e] = pack_bf16(a, b);
    }
    uint4* dst = (uint4*)(out + (size_t)pnl * Nrows * 128 + n * 128 + ((ch ^ (n & 7)) << 4));
    *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------------------------------------------ host side
constexpr int64_t TC_CHUNK = 65536;       // rays per pass: bounds the worst-case (every slot valid) workspace

struct TcWs {
    int32_t *nvalid, *svalid, *tuple_start, *sample_cidx, *partials, *tuple_src, *csample;
    float *loc_pers, *weight_n, *wc, *C0, *sigma;
    uint8_t *wpack, *cpack;
};

static size_t tc_carve(const AggPlan& P, int64_t Rc, int SR, int K, void* base, size_t cap, TcWs* ws)
{
    const AggDims& d = P.dims;
    Arena A(base, cap);
    const size_t S = (size_t)Rc * SR, T = S * K;
    ws->nvalid = A.take<int32_t>(S + 1); ws->svalid = A.take<int32_t>(S + 1);
    ws->tuple_start = A.take<int32_t>(S + 1); ws->sample_cidx = A.take<int32_t>(S + 1);
    ws->partials = A.take<int32_t>(scan_partials_count((int64_t)S));
    ws->tuple_src = A.take<int32_t>(T + 1); ws->csample = A.take<int32_t>(S + 1);
    ws->loc_pers = A.take<float>(S * 3); ws->weight_n = A.take<float>(T); ws->wc = A.take<float>(T);
    const size_t ext = S + T / TC_ROWS + 2;          // compact samples + one carry row per tile + dummy row
    ws->C0 = A.take<float>(ext * d.W); ws->sigma = A.take<float>(ext);
    size_t panels = 0;
    for (int t = 0; t < P.n_tuple_layers; t++) panels += (size_t)(P.layers[t].in + 63) / 64;
    ws->wpack = A.take<uint8_t>(panels * PANEL_B);
    ws->cpack = A.take<uint8_t>((size_t)C_W_PANELS * C_PANEL);
    return A.off;
}

static int tc_supported(const AggPlan& P)
{
    const AggDims& d = P.dims;
    SGN_CHECK_ARG(d.C == TC_C && d.F == TC_F && d.FD == TC_FD && d.W == TC_W,
                  "bf16 tensor-core path is built for feat_dim=32, num_feat_freqs=3, dist_xyz_freq=5, width=256 (got %d,%d,%d,%d); use SGN_PRECISION_FP32",
                  d.C, d.F, d.FD, d.W);
    SGN_CHECK_ARG(d.LD == 0, "bf16 tensor-core path does not take the label embedding yet (label_dim=%d); use SGN_PRECISION_FP32", d.LD);
    SGN_CHECK_ARG(P.n_tuple_layers <= TC_MAX_LAYERS, "bf16 tensor-core path: at most %d per-neighbour layers", TC_MAX_LAYERS);
    SGN_CHECK_ARG(P.n_color_hidden >= 1 && P.n_color_hidden <= C_MAX_HIDDEN && d.WC == CW && d.FV <= 5,
                  "bf16 tensor-core path: colour branch must have 2..%d layers of width %d and num_viewdir_freqs <= 5; use SGN_PRECISION_FP32",
                  C_MAX_HIDDEN + 1, CW);
    for (int t = 0; t < P.n_tuple_layers; t++)
        SGN_CHECK_ARG(P.layers[t].extra != EXTRA_LABEL, "bf16 tensor-core path: label input unsupported");
    return SGN_OK;
}

}  // namespace v1
}  // namespace sgn

using namespace sgn;
using namespace sgn::v1;

int sgn_agg_tc_v1_workspace_bytes(const AggPlan& P, int64_t R, int SR, int K, size_t* bytes)
{
    int rc = tc_supported(P);
    if (rc) return rc;
    TcWs ws;
    *bytes = tc_carve(P, R < TC_CHUNK ? R : TC_CHUNK, SR, K, nullptr, 0, &ws);
    return SGN_OK;
}

int sgn_agg_tc_v1_forward(const AggPlan& P, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                       const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                       int64_t R, int SR, int K, float* decoded, uint8_t* ray_valid, float* loc_pers_out, float* weight_out, float* conf_out,
                       void* workspace, size_t workspace_bytes, cudaStream_t st)
{
    int rc = tc_supported(P);
    if (rc) return rc;
    const AggDims& d = P.dims;
    const int64_t chunk = R < TC_CHUNK ? R : TC_CHUNK;
    TcWs ws;
    const size_t need = tc_carve(P, chunk, SR, K, workspace, workspace_bytes, &ws);
    if (need > workspace_bytes || ((uintptr_t)workspace & 255)) {
        set_error("sgn_agg_forward(bf16): workspace too small or misaligned (need %zu bytes, got %zu)", need, workspace_bytes);
        return SGN_E_WORKSPACE;
    }
    static bool attr_set = false;
    if (!attr_set) {
        SGN_CUDA(cudaFuncSetAttribute(agg_tuple_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        SGN_CUDA(cudaFuncSetAttribute(agg_color_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
        attr_set = true;
    }
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);

    // weights -> bf16 panels in the 128B-swizzled shared-memory image (per call: they may have been updated by the optimiser)
    TcParams tp = {};
    tp.first_panel[0] = 0;
    for (int t = 0; t < P.n_tuple_layers; t++) {
        const LayerInfo& L = P.layers[t];
        const int np = (L.in + 63) / 64;
        launch(tc_pack_weight_kernel, cdiv((int64_t)np * TC_W * 8, 256), 256, 0, st, weights[t], TC_W, L.in, np, ws.wpack + (size_t)tp.first_panel[t] * PANEL_B);
        tp.first_panel[t + 1] = tp.first_panel[t] + np;
        tp.kind[t] = t == 0 ? LAYER_FROM_X0 : (L.extra == EXTRA_COLORDIR ? LAYER_FROM_ACT_E7 : LAYER_FROM_ACT);
        tp.bias[t] = biases[t];
    }
    ColParams cp = {};
    {
        int pnl = 0;
        for (int c = 0; c < P.n_color_hidden; c++) {
            const int l = P.color_layer0 + c;
            const int np = c == 0 ? C_K0_PANELS : 2;
            launch(tc_pack_weight_kernel, cdiv((int64_t)np * CW * 8, 256), 256, 0, st, weights[l], CW, P.layers[l].in, np, ws.cpack + (size_t)pnl * C_PANEL);
            pnl += np;
            cp.bias[c] = biases[l];
        }
    }
    SGN_LAUNCH_CHECK();
    tp.n_layers = P.n_tuple_layers;
    tp.wpack = ws.wpack;
    tp.wa = weights[P.alpha_layer]; tp.ba = biases[P.alpha_layer];
    tp.slope = d.slope; tp.act_super = d.act_super;
    tp.K = K; tp.SR = SR;
    { const char* e = getenv("SGN_TC_DEBUG"); tp.dbg = e ? atoi(e) : 0; }
    cp.dbg = tp.dbg;
    cp.wpack = ws.cpack; cp.n_hidden = P.n_color_hidden; cp.fv = d.FV;
    cp.wl = weights[P.n_layers - 1]; cp.bl = biases[P.n_layers - 1];
    cp.slope = d.slope; cp.act_super = d.act_super; cp.SR = SR;

    for (int64_t r0 = 0; r0 < R; r0 += chunk) {
        const int64_t Rc = R - r0 < chunk ? R - r0 : chunk;
        const int64_t S = Rc * SR;
        const int Tm = (int)(S * K), Sm = (int)S;
        AggIn in;
        in.tab = *tables;
        in.pidx = pidx + r0 * SR * K; in.loc_w = loc_w + r0 * SR * 3; in.raydir = raydir + r0 * 3; in.campos = campos; in.camrot = camrotc2w;
        float* dec = decoded + r0 * SR * 4;
        float* loc_pers = loc_pers_out ? loc_pers_out + r0 * SR * 3 : ws.loc_pers;
        SGN_CUDA(cudaMemsetAsync(dec, 0, sizeof(float) * 4 * (size_t)S, st));
        launch(agg_prepare_kernel, cdiv(S, 128), 128, 0, st, in, S, K, loc_pers, ws.wc, ws.weight_n, weight_out ? weight_out + r0 * SR * K : nullptr,
               conf_out ? conf_out + r0 * SR * K : nullptr, ray_valid + r0 * SR, ws.nvalid, ws.svalid);
        if ((rc = exclusive_scan_i32(ws.nvalid, ws.tuple_start, S, ws.partials, st))) return rc;
        if ((rc = exclusive_scan_i32(ws.svalid, ws.sample_cidx, S, ws.partials, st))) return rc;
        const int32_t* T_ptr = ws.tuple_start + S;
        const int32_t* S_ptr = ws.sample_cidx + S;
        launch(agg_index_kernel, cdiv(S, 128), 128, 0, st, in.pidx, S, K, ws.tuple_start, ws.sample_cidx, ws.nvalid, ws.tuple_src, ws.csample);

        tp.in = in;
        tp.T_ptr = T_ptr; tp.T_max = Tm;
        tp.tuple_src = ws.tuple_src; tp.tuple_start = ws.tuple_start; tp.sample_cidx = ws.sample_cidx;
        tp.S_cap = Sm; tp.n_tiles_cap = cdiv(Tm, TC_ROWS);
        tp.loc_pers = loc_pers; tp.wc = ws.wc;
        tp.F = ws.C0; tp.sigma = ws.sigma;
        const int max_tiles = cdiv(Tm, TC_ROWS);
        launch(agg_tuple_tc_kernel, max_tiles < n_sm ? max_tiles : n_sm, TC_THREADS, TC_SMEM, st, tp);

        // per-sample colour MLP + rgb + (sigma, r, g, b) store
        cp.S_ptr = S_ptr; cp.S_max = Sm; cp.csample = ws.csample; cp.F = ws.C0; cp.sigma = ws.sigma;
        cp.tuple_start = ws.tuple_start; cp.nvalid = ws.nvalid; cp.S_cap = Sm;
        cp.raydir = in.raydir; cp.decoded = dec;
        const int max_ctiles = cdiv(Sm, TC_ROWS);
        launch(agg_color_tc_kernel, max_ctiles < n_sm ? max_ctiles : n_sm, 288, C_SMEM, st, cp);
        SGN_LAUNCH_CHECK();
    }
    return SGN_OK;
}
