// agg_tc.cu -- bf16 tensor-core aggregator forward (sm_100a): tcgen05.mma with accumulators in TMEM, bulk-TMA weight
// streaming, mbarrier pipelines.  Two fused persistent kernels:
//   agg_tuple_tc_kernel : gather -> per-neighbour MLP (block1 / block3) -> alpha -> K-weighted sums per sample
//   agg_color_tc_kernel : per-sample colour MLP -> sigmoid -> (sigma, r, g, b)
//
// Reference: PointAggregator.viewmlp (models/aggregators/point_aggregators.py:561-786) on the canonical non-semantic branch.
//
// Tile = the valid (sample, neighbour) tuples of consecutive samples, at most 128 rows; tiles are sample aligned (a sample's
// rows never straddle two tiles): tile t owns the samples whose first tuple index lies in [t*TW, (t+1)*TW), TW = 128-(K-1).
//
// Per CTA (one per SM, persistent) two tile SLOTS ping-pong: while the epilogue warps turn the accumulator of one slot into
// the next layer's operand, the tensor pipe runs the other slot's MMAs.
//   warps 0-7  epilogue : TMEM -> registers (tcgen05.ld) -> +bias, LeakyReLU -> bf16 -> written IN PLACE over the slot's A
//                         panels (128B-swizzled, K-major) as the next layer's operand.  Last layer: also the alpha dot
//                         product, then the selection matrix Sel[sample slot][row] = w*conf (bf16) next to the activations.
//   warps 8-11 gather   : one thread per row: cp.async of the point's precomputed bf16 row [emb | PE(emb)] (448 B) into the
//                         swizzled panels, PE(dists) and [colour | dir-view | dir.view] computed per tuple
//   warp  12   producer : cp.async.bulk (TMA) of pre-swizzled 32 KB weight panels into a 2-stage ring
//   warp  13   MMA      : one thread issues tcgen05.mma 128x256x16 per K-step; after the last layer one more small MMA,
//                         F^T[feature][sample slot] = H^T (MN-major view of the same activation panels) x Sel^T,
//                         does the K-weighted sum over the rows of every sample on the tensor core
// Shared memory: 2 slots x 5 panels x 16 KB | weight ring 2 x 32 KB | 1.5 KB row scratch | mbarriers  (~226 KB).
// TMEM: 512 columns = one 128 x 256 fp32 accumulator per slot.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "agg_kernels.cuh"
#include "tc_ptx.cuh"

namespace sgn {

constexpr int TC_ROWS = 128;
constexpr int TC_W = 256;                 // layer width == accumulator columns
constexpr int TC_C = 32, TC_F = 3, TC_FD = 5;
constexpr int TC_PT_COLS = TC_C * (1 + 2 * TC_F);              // 224 per-point columns: emb | PE(emb)
// block1.0 has 284 inputs: 224 per point (hoisted, see tc_point_gemm_kernel) + 60 per tuple
constexpr int TC_KD = 2 * TC_FD * 6;                           // 60 PE(dists) columns: the per-tuple K of the first layer, + 2 bias columns (1.0)
constexpr int X0_PANEL = 3;                                    // panel of the slot that holds [PE(dists) | 1 | 1 | 0 | 0] during the first layer
static_assert(TC_KD + 2 <= 64, "PE(dists) and the two bias columns must fit in one K panel");
constexpr int TC_MAX_LAYERS = 6;
constexpr int PANEL_A = TC_PANEL_BYTES;   // 16 KB: 128 rows x 64 bf16
constexpr int PANEL_B = TC_W * 128;       // 32 KB: 256 rows x 64 bf16
constexpr int PANEL_BH = PANEL_B / 2;     // 16 KB: this CTA's half (128 of the 256 output columns) of a weight panel
constexpr int SLOT_PANELS = 5, SLOT_BYTES = SLOT_PANELS * PANEL_A, B_STAGES = 4;
constexpr int E7_COL0 = 32;               // inside panel 4: cols [32,48) hold [colour | dir-view | dir.view | 1 | 1 | 0..]
constexpr int ONES_KSTEP = 1;             // K-step of panel 4 (cols 16..31) that is zero except for 1.0 at its columns 12, 13 (bias K-step operand)
constexpr int META_CHUNK = 6;             // 16-byte chunks 6, 7 of a row of panel 4: {w*conf, sample slot, first padded sample, #slots | passes << 16}, {tuple index, point index}
constexpr int KS_SLOTS = 56;              // sample slots per CTA and K-sum pass (Sel^T = 2 K-panels x 56 x 128 B, inside panel 4)
constexpr int KS_N = 2 * KS_SLOTS;        // the pair's K-sum MMA: columns [0,56) = leader CTA's samples, [56,112) = peer CTA's
constexpr int BIAS_PANEL_B = TC_W * 32;   // 8 KB: compact (unswizzled) [256 x 16] bias K-step (4 KB per CTA)
constexpr int ALPHA_N = 16;               // alpha_branch as an N = 16 MMA (row 0 = the weight vector; 8 rows per CTA)
constexpr int ALPHA_PANEL_B = 4 * ALPHA_N * 128;                // 8 KB: per CTA four 1 KB K panels of 8 rows
constexpr int ALPHA_COL = 2 * KS_N;       // accumulator columns [224, 240) of the slot receive alpha

constexpr int OFF_SLOT0 = 0;
constexpr int OFF_WRING = 2 * SLOT_BYTES;                       // 163840
constexpr int OFF_BAR = OFF_WRING + B_STAGES * PANEL_BH;        // 229376
constexpr int N_BARS = 3 * B_STAGES + 8;
constexpr int OFF_TMEMPTR = OFF_BAR + N_BARS * 8;
constexpr int TC_SMEM = OFF_TMEMPTR + 16 + 1024;                // + slack for the 1024 B alignment of the base
// 16 epilogue warps (four per TMEM lane quadrant, one 64-column activation panel each) | 8 gather warps | producer, MMA issuer, 2 idle
// (whole warpgroups for setmaxnreg).  The epilogue is bound by the latency of its TMEM loads and shared-memory stores, not by issue
// slots: twice the warps on half the columns each halve the time a layer's epilogue takes, which is what the other slot's MMAs hide.
constexpr int TC_EPI_WARPS = 16, TC_GATHER_WARP0 = 16, TC_PRODUCER_WARP = 24, TC_MMA_WARP = 25, TC_THREADS = 28 * 32;
constexpr int REGS_EPI = 72, REGS_GATHER = 80, REGS_MISC = 56;       // 16*72 + 8*80 + 4*56 = 2016 = 28 * 72 (the launch allocation per warp-lane)
static_assert(TC_SMEM <= 232448, "exceeds the 227 KB shared memory limit");

enum { LAYER_FROM_X0 = 0, LAYER_FROM_ACT = 1, LAYER_FROM_ACT_E7 = 2 };
// cta_group::2: one MMA spans the CTA pair, M = 256 = this CTA's 128 rows + the peer's, each CTA supplies half of B's N rows
constexpr uint32_t IDESC_LAYER = tc_idesc(2 * TC_ROWS, TC_W);
constexpr uint32_t IDESC_ALPHA = tc_idesc(2 * TC_ROWS, ALPHA_N);
constexpr uint32_t IDESC_KSUM = tc_idesc(256, KS_N, 1);        // A = H^T read MN-major from the activation panels

struct TcParams {
    AggIn in;
    int K, SR;
    const int32_t* ntiles_ptr; int ntiles_cap;
    const int2* tile_tab;                  // [ntiles + 1] {first tuple, first compact sample} of every tile, then the totals
    const int32_t* cpad0;                  // [ntiles + 1] first PADDED sample index of every tile (each tile's samples padded to a multiple of 8)
    const int32_t* tuple_src; const int32_t* sample_cidx;
    const float* loc_pers; const float* wc;
    const uint8_t* padd[TC_MAX_LAYERS];    // per layer: [N][256] bf16 added to the accumulator in the epilogue (point-only part of the
                                           // layer's input, hoisted into tc_point_gemm_kernel), or NULL: layer 0 (emb | PE(emb)), block2_bpnet.0 (label)
    const uint8_t* wpack;                  // pre-swizzled bf16 weight panels, all layers back to back
    const uint8_t* bpack[TC_MAX_LAYERS];   // compact bias K-step of the layer, or NULL when the bias rides in a weight panel
    const uint8_t* apack;                  // alpha_branch panel
    int n_layers;
    int kind[TC_MAX_LAYERS];
    int first_panel[TC_MAX_LAYERS + 1];
    const float* ba;
    float slope; int act_super;
    uint8_t* F; float* sigrow;             // outputs: F per padded sample (bf16, colour-kernel operand image, see f_image_off), w*conf*act(alpha) per tuple
    int dbg;                               // SGN_TC_DEBUG bitmask (timing experiments only; results invalid when != 0)
};

// K-weighted feature sums are stored as the colour kernel's first-layer A operand, MN-major (the sample index is the contiguous
// dimension, which is how they leave the K-sum MMA: one thread = one feature, consecutive registers = consecutive samples).
// Per 128 padded samples: four 16 KB stages of 64 features; a stage = 2 sample atoms (64 samples) x 8 feature groups x 8 features x
// 128 bytes, 16-byte chunks XOR-swizzled with the feature index (canonical SWIZZLE_128B MN-major, LBO 8 KB, SBO 1 KB), so the
// colour kernel loads a stage with one bulk copy.  cp8 must be a multiple of 8: returns the offset of 8 consecutive samples (16 B).
constexpr int F_TILE_BYTES = 4 * TC_PANEL_BYTES;
__device__ __forceinline__ size_t f_image_off8(int cp8, int f)
{
    return (size_t)(cp8 >> 7) * F_TILE_BYTES + (f >> 6) * TC_PANEL_BYTES + ((cp8 >> 6) & 1) * 8192 + ((f >> 3) & 7) * 1024 + (f & 7) * 128 +
           ((((cp8 & 63) >> 3) ^ (f & 7)) << 4);
}

// A operand read MN-major (M = 64-element atoms 16 KB apart (LBO), K = 8-row groups 1 KB apart (SBO)), 128-byte swizzle
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(PANEL_A >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// K-major operand without swizzle: 8 x 16-byte core matrices, the two K halves 128 B apart (LBO), 8-row groups 256 B apart (SBO)
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46);
}
// ---- cluster (CTA pair) primitives
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier of this CTA whose arrivals may come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
    } while (!ok);
}
// pair-wide MMA, issued by the leader CTA only
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// completion of all MMAs issued so far -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t r;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(r));
    return r != 0;
}
// 32-byte global load (whole sector per lane): two consecutive uint4
__device__ __forceinline__ void ldg256(const uint4* ptr, uint4& a, uint4& b)
{
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(ptr));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void sts16(uint32_t addr, uint16_t v) { asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }     // all epilogue warps
// 16-column TMEM load of this warp's 32 lanes, and its wait (the destination registers are in/out operands of the wait so that no use
// of them can be scheduled above it)
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld16(uint32_t (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :: "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t tc_ld1(uint32_t taddr)
{
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}
// LeakyReLU on a pair, after rounding to bf16: max(x, slope * x)
__device__ __forceinline__ uint32_t leaky_pack(uint32_t a, uint32_t b, __nv_bfloat162 slope2)
{
    const __nv_bfloat162 x = __floats2bfloat162_rn(__uint_as_float(a), __uint_as_float(b));
    const __nv_bfloat162 h = __hmax2(x, __hmul2(x, slope2));
    return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------ the kernel
// Launched as clusters of two CTAs (a CTA pair on the two SMs of a TPC).  Every CTA gathers, activates and reduces its OWN
// tiles; only the MMAs are shared: the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2, M = 256 = its 128 rows +
// the peer's 128 rows, and each CTA streams and holds only its half of every weight panel (N split) -- half the shared-memory
// operand traffic and half the L2 weight traffic per SM.  Per pair-cycle a cluster works on four tiles: (slot 0|1) x (rank 0|1).
template <bool kDbg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) agg_tuple_tc_kernel(const __grid_constant__ TcParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    const uint32_t bar0 = sbase + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    // W_FULL / W_EMPTY / D_FULL / BUF_FREE: one per CTA.  PEER_W (the peer's weight stage has landed), X_FULL, A_READY: the
    // leader's instances collect arrivals from both CTAs (the peer arrives through the cluster address).
    const int W_FULL = 0, W_EMPTY = B_STAGES, PEER_W = 2 * B_STAGES, X_FULL = 3 * B_STAGES, BUF_FREE = X_FULL + 2, D_FULL = BUF_FREE + 2, A_READY = D_FULL + 2;
    auto LEADER_BAR = [&](int i) { return map_to_cta(BAR(i), 0); };
    uint32_t* tmem_ptr_smem = (uint32_t*)(smem + OFF_TMEMPTR);

    const int ntiles = min(*p.ntiles_ptr, p.ntiles_cap);
    const int ncycles = (ntiles + 3) >> 2;                          // pair-cycles of the whole grid: 4 tiles each
    const int cl0 = blockIdx.x >> 1, ncl = gridDim.x >> 1;

    if (tid == 0) {
        for (int s = 0; s < B_STAGES; s++) { mbar_init(BAR(W_FULL + s), 1); mbar_init(BAR(W_EMPTY + s), 1); mbar_init(BAR(PEER_W + s), 1); }
        for (int s = 0; s < 2; s++) {
            mbar_init(BAR(X_FULL + s), 2 * 4); mbar_init(BAR(BUF_FREE + s), 1);
            mbar_init(BAR(D_FULL + s), 1); mbar_init(BAR(A_READY + s), 2 * TC_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                                 // both CTAs' barriers are initialised before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // hand the registers to the warps that need them (whole warpgroups: epilogue 0-15 keep the launch allocation of 72, the four
    // producer / MMA / idle warps 24-27 give theirs up first, then the gather warps 16-23 take 80)
    if (warp >= TC_PRODUCER_WARP) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_MISC));
    else if (warp >= TC_GATHER_WARP0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_GATHER));

    if (warp < TC_EPI_WARPS) {
        // =========================================================== EPILOGUE: warp = (64-column panel q4, TMEM lane quadrant)
        const int quad = warp & 3, q4 = warp >> 2;
        const int row = quad * 32 + lane;
        const int et = tid;                                      // 0..511
        uint32_t ph_d = 0;
        const uint32_t lane_field = (uint32_t)(quad * 32) << 16;
        const __nv_bfloat162 slope2 = __float2bfloat162_rn(p.slope);
        const float ba = p.ba[0];
        long long pf_wait = 0, pf_mid = 0, pf_last = 0, pf_drain = 0, pf_t0 = 0;
        const bool prof = kDbg && (kDbg && (p.dbg & 32)) != 0;
        uint32_t tcount = 0;
        const uint32_t a_ready0 = LEADER_BAR(A_READY);
        // this warp's part of "slot s is ready for the next MMA": every lane's writes are fenced, then one arrival for the warp
        auto warp_ready = [&](int s) {
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(a_ready0 + 8u * s);
        };

        for (int cyc = cl0; cyc < ncycles; cyc += ncl) {
            int nslots0 = 0, nslots1 = 0, npass0 = 1, npass1 = 1, c00 = 0, c01 = 0, slr0 = -1, slr1 = -1, j0 = 0, j1 = 0;
            float wcr0 = 0.f, wcr1 = 0.f;
            int pt0 = 0, pt1 = 0;                                     // point of this row, per slot
            tcount += 2;
            for (int l = 0; l < p.n_layers; l++) {
                const bool last = (l == p.n_layers - 1);
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    if (prof) pf_t0 = clock64();
                    mbar_wait(BAR(D_FULL + s), (ph_d >> s) & 1u); ph_d ^= 1u << s;
                    tc_fence_after();
                    if (prof) { const long long t1 = clock64(); pf_wait += t1 - pf_t0; pf_t0 = t1; }
                    const uint32_t slot_base = sbase + OFF_SLOT0 + s * SLOT_BYTES;
                    if (l == 0) {
                        // this row's {w*conf, sample slot, first compact sample of the tile, #sample slots | K-sum passes << 16}
                        const uint4 m = lds128u(slot_base + 4 * PANEL_A + row * 128 + ((META_CHUNK ^ (row & 7)) << 4));
                        const uint4 m2 = lds128u(slot_base + 4 * PANEL_A + row * 128 + (((META_CHUNK + 1) ^ (row & 7)) << 4));
                        const int j = (int)m2.x;
                        if (s == 0) pt0 = (int)m2.y; else pt1 = (int)m2.y;
                        if (s == 0) { wcr0 = __uint_as_float(m.x); slr0 = (int)m.y; c00 = (int)m.z; nslots0 = (int)(m.w & 0xffffu); npass0 = (int)(m.w >> 16); j0 = j; }
                        else { wcr1 = __uint_as_float(m.x); slr1 = (int)m.y; c01 = (int)m.z; nslots1 = (int)(m.w & 0xffffu); npass1 = (int)(m.w >> 16); j1 = j; }
                    }
                    const float wcr = s == 0 ? wcr0 : wcr1;
                    const int slr = s == 0 ? slr0 : slr1;
                    // this warp's 64 accumulator columns = activation panel q4, this thread's row of it (128 bytes)
                    const uint32_t acc_addr = tmem_base + (uint32_t)(s * TC_W + q4 * 64) + lane_field;
                    const uint32_t act_row = slot_base + q4 * PANEL_A + row * 128;
                    uint32_t v0[16], v1[16];
                    if (p.padd[l]) {
                        // the layer's point table (bf16 [N][256]) is added to the accumulator before the activation: this row's 64 columns
                        // = 128 bytes, fetched in two halves, each ahead of its use; x = acc + P0 in fp32, then LeakyReLU in bf16
                        const uint4* p0row = (const uint4*)(p.padd[l] + (size_t)(s == 0 ? pt0 : pt1) * (TC_W * 2)) + q4 * 8;
                        uint4 pa[4];
                        ldg256(p0row, pa[0], pa[1]); ldg256(p0row + 2, pa[2], pa[3]);
                        auto chunk0 = [&](int c, const uint32_t(&vv)[16], const uint4& pq0, const uint4& pq1) {
#pragma unroll
                            for (int q = 0; q < 2; q++) {
                                const uint4 pq = q == 0 ? pq0 : pq1;
                                const uint32_t w[4] = {pq.x, pq.y, pq.z, pq.w};
                                uint32_t o[4];
#pragma unroll
                                for (int e = 0; e < 4; e++)
                                    o[e] = leaky_pack(__float_as_uint(__uint_as_float(vv[8 * q + 2 * e]) + __uint_as_float(w[e] << 16)),
                                                      __float_as_uint(__uint_as_float(vv[8 * q + 2 * e + 1]) + __uint_as_float(w[e] & 0xffff0000u)), slope2);
                                sts128(act_row + (((2 * c + q) ^ (row & 7)) << 4), o[0], o[1], o[2], o[3]);
                            }
                        };
                        tc_ld16_nowait(acc_addr, v0);
                        tc_wait_ld16(v0);
                        tc_ld16_nowait(acc_addr + 16u, v1);
                        chunk0(0, v0, pa[0], pa[1]);
                        tc_wait_ld16(v1);
                        tc_ld16_nowait(acc_addr + 32u, v0);
                        chunk0(1, v1, pa[2], pa[3]);
                        ldg256(p0row + 4, pa[0], pa[1]); ldg256(p0row + 6, pa[2], pa[3]);
                        tc_wait_ld16(v0);
                        tc_ld16_nowait(acc_addr + 48u, v1);
                        chunk0(2, v0, pa[0], pa[1]);
                        tc_wait_ld16(v1);
                        chunk0(3, v1, pa[2], pa[3]);
                    } else {
                        // one 16-column chunk: (bias is part of the GEMM) LeakyReLU in bf16 -> 32 bytes of this row
                        auto chunk = [&](int c, const uint32_t(&vv)[16]) {
                            if (kDbg && (p.dbg & 8)) return;
#pragma unroll
                            for (int q = 0; q < 2; q++)
                                sts128(act_row + (((2 * c + q) ^ (row & 7)) << 4), leaky_pack(vv[8 * q], vv[8 * q + 1], slope2),
                                       leaky_pack(vv[8 * q + 2], vv[8 * q + 3], slope2), leaky_pack(vv[8 * q + 4], vv[8 * q + 5], slope2),
                                       leaky_pack(vv[8 * q + 6], vv[8 * q + 7], slope2));
                        };
                        tc_ld16_nowait(acc_addr, v0);
                        tc_wait_ld16(v0);
                        tc_ld16_nowait(acc_addr + 16u, v1);
                        chunk(0, v0);
                        tc_wait_ld16(v1);
                        tc_ld16_nowait(acc_addr + 32u, v0);
                        chunk(1, v1);
                        tc_wait_ld16(v0);
                        tc_ld16_nowait(acc_addr + 48u, v1);
                        chunk(2, v0);
                        tc_wait_ld16(v1);
                        chunk(3, v1);
                    }
                    if (last) {
                        // H is in the panels; selection matrix of K-sum pass 0 next to it: Sel[sample slot][row] = w*conf (bf16)
                        const uint32_t sel_base = slot_base + 4 * PANEL_A;
                        const uint32_t z = sel_base + et * 32;
                        sts128(z, 0u, 0u, 0u, 0u); sts128(z + 16, 0u, 0u, 0u, 0u);
                        epi_bar();
                        if (q4 == 0 && slr >= 0 && slr < KS_SLOTS) {
                            const __nv_bfloat16 wb = __float2bfloat16_rn(wcr);
                            sts16(sel_base + (row >> 6) * (KS_SLOTS * 128) + slr * 128 + ((((row & 63) >> 3) ^ (slr & 7)) << 4) + (row & 7) * 2,
                                  *reinterpret_cast<const uint16_t*>(&wb));
                        }
                    }
                    warp_ready(s);
                    if (prof) { const long long t1 = clock64(); if (last) pf_last += t1 - pf_t0; else pf_mid += t1 - pf_t0; pf_t0 = t1; }
                }
            }
            // ---- alpha / sigma, and the K-sums: F^T[feature = TMEM lane][sample slot = column] -> F image.  Warp (q4, quad): feature half
            // h2 = q4 >> 1, features h2*128 + quad*32 + lane, and the 32-column half jh = q4 & 1 of this CTA's (up to) 56 sample slots
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int npass = s == 0 ? npass0 : npass1, nslots = s == 0 ? nslots0 : nslots1, c0 = s == 0 ? c00 : c01, slr = s == 0 ? slr0 : slr1;
                const float wcr = s == 0 ? wcr0 : wcr1;
                const int h2 = q4 >> 1, jh = q4 & 1;
                const int f = h2 * 128 + quad * 32 + lane;
                const uint32_t sel_base = sbase + OFF_SLOT0 + s * SLOT_BYTES + 4 * PANEL_A;
                for (int pass = 0; pass < npass; pass++) {
                    if (prof) pf_t0 = clock64();
                    mbar_wait(BAR(D_FULL + s), (ph_d >> s) & 1u); ph_d ^= 1u << s;
                    tc_fence_after();
                    if (pass == 0 && q4 == 0) {                                            // warp-uniform: the TMEM load is .sync.aligned
                        // this tuple's term of sigma = sum_k w*conf*act(alpha_k); alpha came out of the alpha_branch MMA, the colour
                        // kernel adds up the (consecutive) terms of a sample
                        const float a = __uint_as_float(tc_ld1(tmem_base + (uint32_t)(s * TC_W + ALPHA_COL) + lane_field)) + ba;
                        const float act = p.act_super ? softplus1(a - 1.0f) : fmaxf(a, 0.f);
                        if (slr >= 0) p.sigrow[s == 0 ? j0 : j1] = (kDbg && (p.dbg & 24)) ? 0.f : act * wcr;
                    }
                    const int ns8 = min(KS_SLOTS, ((nslots + 7) & ~7) - pass * KS_SLOTS);   // padded slots of this pass, a multiple of 8
                    const int cbase = c0 + pass * KS_SLOTS;                                 // first padded sample index of the pass (multiple of 8)
                    // this CTA's samples are columns [rank*56, rank*56+56) of each feature half's [128 x 112] block
                    const uint32_t d_addr = tmem_base + (uint32_t)(s * TC_W + h2 * KS_N + rank * KS_SLOTS) + lane_field;
                    if (32 * jh < ns8) {                                                    // (warp-uniform)
#pragma unroll
                        for (int hh = 0; hh < 2; hh++) {
                            if (32 * jh + 16 * hh >= ns8) break;
                            uint32_t v[16];
                            tc_ld16_nowait(d_addr + 32 * jh + 16 * hh, v);               // jh = 1, hh = 1 reads 8 columns past the 56 (in bounds, unused)
                            tc_wait_ld16(v);
                            if (!(kDbg && (p.dbg & 16))) {
#pragma unroll
                                for (int g = 0; g < 2; g++)
                                    if (32 * jh + 16 * hh + 8 * g < ns8) {
                                        const uint4 o = make_uint4(pack_bf16(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                                                                   pack_bf16(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                                                                   pack_bf16(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                                                                   pack_bf16(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
                                        *(uint4*)(p.F + f_image_off8(cbase + 32 * jh + 16 * hh + 8 * g, f)) = o;
                                    }
                            }
                        }
                    }
                    if (pass + 1 < npass) {
                        // next 56 sample slots: rebuild Sel (the MMA of this pass has completed, nobody reads it now)
                        const uint32_t z = sel_base + et * 32;
                        sts128(z, 0u, 0u, 0u, 0u); sts128(z + 16, 0u, 0u, 0u, 0u);
                        epi_bar();
                        const int n = slr - (pass + 1) * KS_SLOTS;
                        if (q4 == 0 && n >= 0 && n < KS_SLOTS) {
                            const __nv_bfloat16 wb = __float2bfloat16_rn(wcr);
                            sts16(sel_base + (row >> 6) * (KS_SLOTS * 128) + n * 128 + ((((row & 63) >> 3) ^ (n & 7)) << 4) + (row & 7) * 2,
                                  *reinterpret_cast<const uint16_t*>(&wb));
                        }
                    }
                    warp_ready(s);
                    if (prof) { const long long t1 = clock64(); pf_drain += t1 - pf_t0; pf_t0 = t1; }
                }
            }
        }
        if (prof && blockIdx.x == 0 && lane == 0)
            printf("epi warp %d: tiles %u (of %d) wait %lld hidden %lld last %lld drain %lld (cycles/tile)\n", warp, tcount, ntiles, pf_wait / max(tcount, 1u),
                   pf_mid / max(tcount, 1u), pf_last / max(tcount, 1u), pf_drain / max(tcount, 1u));
    } else if (warp < TC_PRODUCER_WARP) {
        // =========================================================== GATHER: warps 8-11 feed slot 0, warps 12-15 slot 1; one thread per row
        const int s = (warp - TC_GATHER_WARP0) >> 2;
        const int row = tid - (TC_GATHER_WARP0 + 4 * s) * 32;
        uint32_t ph_free = 1;
        const float* Rm = p.in.camrot;
        const float r00 = Rm[0], r01 = Rm[1], r02 = Rm[2], r10 = Rm[3], r11 = Rm[4], r12 = Rm[5], r20 = Rm[6], r21 = Rm[7], r22 = Rm[8];
        const float cpx = p.in.campos[0], cpy = p.in.campos[1], cpz = p.in.campos[2];
        long long gf_load = 0, gf_wait = 0, gf_write = 0, gf_t0 = 0;
        const bool prof = kDbg && (kDbg && (p.dbg & 32)) != 0;
        uint32_t tcount = 0;
        const uint32_t x0 = sbase + OFF_SLOT0 + s * SLOT_BYTES;
        const uint32_t one_one = pack_bf16(1.0f, 1.0f);
        const uint32_t x_full = LEADER_BAR(X_FULL + s);
        for (int cyc = cl0; cyc < ncycles; cyc += ncl) {
            const int tile = 4 * cyc + 2 * s + (int)rank, ptile = tile ^ 1;          // ptile: the peer CTA's tile in the same slot
            tcount++;
            if (prof) gf_t0 = clock64();
            // ---- everything that does not need the slot: indices, point data, PE(dists), [colour | dir-view | dir.view], in registers
            int2 ta = make_int2(0, 0), tb = make_int2(0, 0);
            int cpad = 0;
            if (tile < ntiles) { ta = p.tile_tab[tile]; tb = p.tile_tab[tile + 1]; cpad = p.cpad0[tile]; }
            int pslots = 0;
            if (ptile < ntiles) pslots = p.tile_tab[ptile + 1].y - p.tile_tab[ptile].y;
            const int nsl = tb.y - ta.y;
            const int npass = max(1, (max(nsl, pslots) + KS_SLOTS - 1) / KS_SLOTS);   // the pair runs the same number of K-sum passes
            const bool live = row < tb.x - ta.x && !(kDbg && (p.dbg & 4));
            uint32_t pe[32], e7p[4] = {0u, 0u, 0u, 0u};
            float wcv = 0.f; int slot = -1; int ptv = 0;
            if (live) {
                const int flat = p.tuple_src[ta.x + row];
                const int64_t sm = flat / p.K;
                const int64_t r = sm / p.SR;
                const int64_t pt = p.in.pidx[flat];
                ptv = (int)pt;
                for (int l = 0; l < p.n_layers; l++)                             // the epilogues read these rows: bring them into L2 now
                    if (p.padd[l]) {
                        const uint8_t* src = p.padd[l] + (size_t)pt * (TC_W * 2);
                        prefetch_l2(src); prefetch_l2(src + 128); prefetch_l2(src + 256); prefetch_l2(src + 384);
                    }
                wcv = p.wc[flat];
                slot = p.sample_cidx[sm] - ta.y;
                float dist[6];
                const float px = p.in.tab.xyz[3 * pt], py = p.in.tab.xyz[3 * pt + 1], pz = p.in.tab.xyz[3 * pt + 2];
                dist[0] = px - p.in.loc_w[3 * sm]; dist[1] = py - p.in.loc_w[3 * sm + 1]; dist[2] = pz - p.in.loc_w[3 * sm + 2];
                const float sx = px - cpx, sy = py - cpy, sz = pz - cpz;
                const float c0 = sx * r00 + sy * r10 + sz * r20, c1 = sx * r01 + sy * r11 + sz * r21, c2 = sx * r02 + sy * r12 + sz * r22;
                const float xp = c0 / c2, yp = c1 / c2;
                const float lxp = p.loc_pers[3 * sm], lyp = p.loc_pers[3 * sm + 1], lzp = p.loc_pers[3 * sm + 2];
                dist[3] = xp * c2 - lxp * lzp; dist[4] = yp * c2 - lyp * lzp; dist[5] = c2 - lzp;
                const float vx = p.in.raydir[3 * r], vy = p.in.raydir[3 * r + 1], vz = p.in.raydir[3 * r + 2];
                const float dx = p.in.tab.dir[3 * pt], dy = p.in.tab.dir[3 * pt + 1], dz = p.in.tab.dir[3 * pt + 2];
                e7p[0] = pack_bf16(p.in.tab.color[3 * pt], p.in.tab.color[3 * pt + 1]);
                e7p[1] = pack_bf16(p.in.tab.color[3 * pt + 2], dx - vx);
                e7p[2] = pack_bf16(dy - vy, dz - vz);
                e7p[3] = pack_bf16(dx * vx + dy * vy + dz * vz, 1.0f);
                // pe[d*FD + f] = (sin, cos)(dist_d * 2^f): base angle by sincosf, octaves by the double-angle recurrence
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    float sn, cs_;
                    __sincosf(dist[d], &sn, &cs_);
#pragma unroll
                    for (int f = 0; f < TC_FD; f++) {
                        pe[d * TC_FD + f] = pack_bf16(sn, cs_);
                        const float s2 = 2.0f * sn * cs_, c2 = 1.0f - 2.0f * sn * sn;
                        sn = s2; cs_ = c2;
                    }
                }
                pe[30] = one_one; pe[31] = 0u;                                   // cols 60, 61 = 1 (bias columns), 62, 63 = 0
            }
            if (prof) { const long long t1 = clock64(); gf_load += t1 - gf_t0; gf_t0 = t1; }
            mbar_wait(BAR(BUF_FREE + s), ph_free); ph_free ^= 1;
            if (prof) { const long long t1 = clock64(); gf_wait += t1 - gf_t0; gf_t0 = t1; }
            // first-layer operand of the row, panel 3: [PE(dists) (60) | 1 | 1 | 0 | 0]; dead rows are all zero (what is left in the
            // panel from the previous tile must not reach the MMAs).  Panel 4: K-step 1 = (0 x12, 1, 1, 0, 0) for the bias K-steps,
            // K-step 2 = [colour | dir-view | dir.view | 1 | 1 | 0..] for block3.0, chunks 6, 7 = per-row metadata for the epilogue.
            const uint32_t xrow = x0 + X0_PANEL * PANEL_A + row * 128, prow = x0 + 4 * PANEL_A + row * 128;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (live) sts128(xrow + ((q ^ (row & 7)) << 4), pe[4 * q], pe[4 * q + 1], pe[4 * q + 2], pe[4 * q + 3]);
                else sts128(xrow + ((q ^ (row & 7)) << 4), 0u, 0u, 0u, 0u);
            }
            sts128(prow + ((2 ^ (row & 7)) << 4), 0u, 0u, 0u, 0u);
            sts128(prow + ((3 ^ (row & 7)) << 4), 0u, 0u, live ? one_one : 0u, 0u);
            sts128(prow + ((4 ^ (row & 7)) << 4), e7p[0], e7p[1], e7p[2], e7p[3]);
            sts128(prow + ((5 ^ (row & 7)) << 4), live ? pack_bf16(1.0f, 0.f) : 0u, 0u, 0u, 0u);
            sts128(x0 + 4 * PANEL_A + row * 128 + ((META_CHUNK ^ (row & 7)) << 4), __float_as_uint(wcv), (uint32_t)slot, (uint32_t)cpad,
                   (uint32_t)nsl | ((uint32_t)npass << 16));
            sts128(x0 + 4 * PANEL_A + row * 128 + (((META_CHUNK + 1) ^ (row & 7)) << 4), (uint32_t)(ta.x + row), (uint32_t)ptv, 0u, 0u);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(x_full);
            if (prof) { const long long t1 = clock64(); gf_write += t1 - gf_t0; gf_t0 = t1; }
        }
        if (prof && blockIdx.x == 0 && lane == 0)
            printf("gather warp %d: tiles %u prepare %lld wait-slot %lld copy %lld (cycles/tile)\n", warp, tcount, gf_load / max(tcount, 1u),
                   gf_wait / max(tcount, 1u), gf_write / max(tcount, 1u));
    } else if (warp == TC_PRODUCER_WARP) {
        // =========================================================== PRODUCER: this CTA's half of every weight panel, once per (layer, slot)
        if (lane == 0) {
            uint32_t ph_empty = (1u << B_STAGES) - 1u;
            uint32_t n = 0;
            auto push = [&](const uint8_t* srcp, uint32_t bytes) {
                const uint32_t st = n % B_STAGES;
                mbar_wait(BAR(W_EMPTY + st), (ph_empty >> st) & 1u); ph_empty ^= 1u << st;
                n++;
                if (kDbg && (p.dbg & 1)) { mbar_arrive(BAR(W_FULL + st)); return; }
                mbar_expect_tx(BAR(W_FULL + st), bytes);
                bulk_g2s(sbase + OFF_WRING + st * PANEL_BH, srcp, bytes, BAR(W_FULL + st));
            };
            for (int cyc = cl0; cyc < ncycles; cyc += ncl) {
                for (int l = 0; l < p.n_layers; l++)
                    for (int s = 0; s < 2; s++) {
                        for (int pi = p.first_panel[l]; pi < p.first_panel[l + 1]; pi++) push(p.wpack + (size_t)pi * PANEL_B + rank * PANEL_BH, PANEL_BH);
                        if (p.bpack[l]) push(p.bpack[l] + rank * (BIAS_PANEL_B / 2), BIAS_PANEL_B / 2);
                    }
                for (int s = 0; s < 2; s++) push(p.apack + rank * (ALPHA_PANEL_B / 2), ALPHA_PANEL_B / 2);
            }
        }
    } else if (warp == TC_MMA_WARP && !leader) {
        // =========================================================== peer CTA: forward "my half of stage st has landed" to the leader
        if (lane == 0) {
            uint32_t ph_full = 0, n = 0;
            const uint32_t peer_w0 = LEADER_BAR(PEER_W);
            int per_cycle = 2;                                           // alpha panels
            for (int l = 0; l < p.n_layers; l++) per_cycle += 2 * (p.first_panel[l + 1] - p.first_panel[l] + (p.bpack[l] ? 1 : 0));
            for (int cyc = cl0; cyc < ncycles; cyc += ncl)
                for (int i = 0; i < per_cycle; i++, n++) {
                    const uint32_t st = n % B_STAGES;
                    mbar_wait(BAR(W_FULL + st), (ph_full >> st) & 1u); ph_full ^= 1u << st;
                    mbar_arrive_cluster(peer_w0 + 8u * st);
                }
        }
    } else if (warp == TC_MMA_WARP) {
        // =========================================================== leader CTA: MMA issuer for the pair.  The whole warp walks the
        // (warp-uniform) schedule, one elected lane issues each tcgen05 instruction
        {
            uint32_t ph_full = 0, ph_peer = 0, ph_x = 0, ph_a = 3;          // phase bits, one per stage / slot
            uint32_t n = 0;
            const bool prof = kDbg && (p.dbg & 32) != 0;
            long long mf_w = 0, mf_pw = 0, mf_a = 0, mf_x = 0, mf_ks = 0, mf_t0 = 0, mf_start = prof ? clock64() : 0;
            auto next_stage = [&]() -> uint32_t {
                const uint32_t st = n % B_STAGES;
                if (prof) mf_t0 = clock64();
                mbar_wait(BAR(W_FULL + st), (ph_full >> st) & 1u); ph_full ^= 1u << st;
                if (prof) { const long long t1 = clock64(); mf_w += t1 - mf_t0; mf_t0 = t1; }
                mbar_wait_cluster(BAR(PEER_W + st), (ph_peer >> st) & 1u); ph_peer ^= 1u << st;
                if (prof) mf_pw += clock64() - mf_t0;
                tc_fence_after();
                return st;
            };
            auto release_stage = [&](uint32_t st) {
                if (elect_one()) tc_commit2(BAR(W_EMPTY + st));
                n++;
            };
            for (int cyc = cl0; cyc < ncycles; cyc += ncl) {
                int npass[2];
                for (int s = 0; s < 2; s++) {
                    int m = 0;
                    for (int r = 0; r < 2; r++) {
                        const int tile = 4 * cyc + 2 * s + r;
                        if (tile < ntiles) m = max(m, p.tile_tab[tile + 1].y - p.tile_tab[tile].y);
                    }
                    npass[s] = max(1, (m + KS_SLOTS - 1) / KS_SLOTS);
                }
                for (int l = 0; l < p.n_layers; l++) {
                    const int np = p.first_panel[l + 1] - p.first_panel[l];
                    const int kind = p.kind[l];
                    const bool own_bias_step = p.bpack[l] != nullptr;
                    for (int s = 0; s < 2; s++) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(s * TC_W);
                        const uint32_t a_base = sbase + OFF_SLOT0 + s * SLOT_BYTES;
                        if (prof) mf_t0 = clock64();
                        mbar_wait_cluster(BAR(A_READY + s), (ph_a >> s) & 1u); ph_a ^= 1u << s;
                        if (prof) { const long long t1 = clock64(); mf_a += t1 - mf_t0; mf_t0 = t1; }
                        if (kind == LAYER_FROM_X0) { mbar_wait_cluster(BAR(X_FULL + s), (ph_x >> s) & 1u); ph_x ^= 1u << s; }
                        if (prof) mf_x += clock64() - mf_t0;
                        tc_fence_after();
                        uint32_t acc = 0;
                        for (int kp = 0; kp < np; kp++) {
                            uint32_t a_addr = a_base + kp * PANEL_A;
                            int ksteps = 4;
                            if (kind == LAYER_FROM_X0) a_addr = a_base + X0_PANEL * PANEL_A;            // [PE(dists) | 1 | 1 | 0 | 0], one panel
                            else if (kp == 4) { a_addr += E7_COL0 * 2; ksteps = 1; }                    // [colour | dir-view | dir.view | 1 | 1] of block3.0
                            const uint32_t st = next_stage();
                            const uint64_t ad = umma_desc(a_addr), bd = umma_desc(sbase + OFF_WRING + st * PANEL_BH);
                            if (elect_one()) {                                                           // one lane issues the panel's K-steps back to back
                                tc_mma2(d_tmem, ad, bd, IDESC_LAYER, acc);
                                if (ksteps > 1) tc_mma2(d_tmem, ad + 2, bd + 2, IDESC_LAYER, 1u);        // next K-step: +32 bytes
                                if (ksteps > 2) {
                                    tc_mma2(d_tmem, ad + 4, bd + 4, IDESC_LAYER, 1u);
                                    tc_mma2(d_tmem, ad + 6, bd + 6, IDESC_LAYER, 1u);
                                }
                            }
                            acc = 1;
                            release_stage(st);
                        }
                        if (own_bias_step) {
                            // bias: one more K-step, A = the operand columns that hold (.., 1, 1, 0, 0), B = the compact bias panel
                            const uint32_t st = next_stage();
                            if (elect_one())
                                tc_mma2(d_tmem, umma_desc(a_base + 4 * PANEL_A + ONES_KSTEP * 32), umma_desc_nosw(sbase + OFF_WRING + st * PANEL_BH), IDESC_LAYER, 1u);
                            release_stage(st);
                        }
                        if (elect_one()) tc_commit2(BAR(D_FULL + s));
                    }
                }
                // alpha = H x wa^T (N = 16, column 0), then the K-weighted sums of both CTAs in one MMA per feature half:
                // F^T[128 features per CTA][112 = 56 leader + 56 peer sample slots] = H^T x Sel^T, K = the tile's 128 rows;
                // each CTA supplies its own selection matrix as its half of B and reads back only its own 56 columns.
                for (int s = 0; s < 2; s++) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(s * TC_W);
                    const uint32_t a_base = sbase + OFF_SLOT0 + s * SLOT_BYTES;
                    const uint32_t sel = a_base + 4 * PANEL_A;
                    for (int pass = 0; pass < npass[s]; pass++) {
                        if (prof) mf_t0 = clock64();
                        mbar_wait_cluster(BAR(A_READY + s), (ph_a >> s) & 1u); ph_a ^= 1u << s;
                        if (prof) mf_ks += clock64() - mf_t0;
                        tc_fence_after();
                        if (pass == 0) {
                            const uint32_t st = next_stage();
                            const uint32_t b_addr = sbase + OFF_WRING + st * PANEL_BH;
                            if (elect_one()) {
#pragma unroll
                                for (int kp = 0; kp < 4; kp++) {
                                    const uint64_t ad = umma_desc(a_base + kp * PANEL_A), bd = umma_desc(b_addr + kp * (ALPHA_N / 2 * 128));
#pragma unroll
                                    for (int k = 0; k < 4; k++) tc_mma2(d_tmem + ALPHA_COL, ad + 2 * k, bd + 2 * k, IDESC_ALPHA, (kp | k) != 0);
                                }
                            }
                            release_stage(st);
                        }
                        if (elect_one()) {
#pragma unroll
                            for (int half = 0; half < 2; half++) {
                                const uint64_t ad = umma_desc_mn(a_base + (2 * half) * PANEL_A);
#pragma unroll
                                for (int ks = 0; ks < TC_ROWS / 16; ks++)                                // +16 rows of the tile per K-step
                                    tc_mma2(d_tmem + (uint32_t)(half * KS_N), ad + ks * (2048 >> 4),
                                            umma_desc(sel + (ks >> 2) * (KS_SLOTS * 128) + (ks & 3) * 32), IDESC_KSUM, ks > 0);
                            }
                            tc_commit2(BAR(D_FULL + s));
                        }
                    }
                    if (elect_one()) tc_commit2(BAR(BUF_FREE + s));
                }
            }
            if (prof && blockIdx.x == 0 && lane == 0) {
                const long long cyc = max((ncycles - cl0 + ncl - 1) / ncl, 1);
                printf("mma issuer: %lld cycles/pair-cycle; waiting: own weights %lld, peer weights %lld, epilogue (layers) %lld, gather %lld, epilogue (K-sum) %lld\n",
                       (clock64() - mf_start) / cyc, mf_w / cyc, mf_pw / cyc, mf_a / cyc, mf_x / cyc, mf_ks / cyc);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                                 // the leader's MMAs read the peer's shared memory and write its TMEM
    if (warp == TC_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ================================================================================================ colour branch
// Per-sample colour MLP on tensor cores (point_aggregators.py:298-309 raw2out_color, :771-786): one persistent CTA per SM,
// tile = 128 compact samples.  All hidden-layer weights stay resident in shared memory (bf16, 128B-swizzled K-major
// panels); the A operand of the first layer is streamed: the per-neighbour kernel left the K-sums F as bf16 in exactly the
// swizzled panel image, so four bulk copies (TMA) per tile bring them into a 2-stage ring (the next tile is prefetched into
// L2 meanwhile); the fifth panel, the view-direction encoding, is computed by the loader warps.  The 128-wide activations live in place in two panels; the last Linear (128 -> 3), the
// sigmoid and the (sigma, r, g, b) store are fused into the last epilogue.
//   warps 0-15 epilogue (four column quarters x four TMEM lane quadrants) | warps 16-19 loaders | warp 20 MMA issuer (and the one-off weight load)
constexpr int CW = 128;                                   // colour hidden width
constexpr int C_PANEL = CW * 128;                         // 16 KB: 128 rows x 64 bf16 (A and B panels alike)
constexpr int C_K0_PANELS = 5, C_RING = 2, C_MAX_HIDDEN = 3;
constexpr int COFF_W = 0;                                                   // resident weights: 5 + 2 + 2 panels
constexpr int C_W_PANELS = C_K0_PANELS + 2 * (C_MAX_HIDDEN - 1);
constexpr int COFF_RING = COFF_W + C_W_PANELS * C_PANEL;
constexpr int COFF_ACT = COFF_RING + C_RING * C_PANEL;
constexpr int COFF_BIAS = COFF_ACT + 2 * C_PANEL;                           // [3][128] hidden biases
constexpr int COFF_WL = COFF_BIAS + C_MAX_HIDDEN * CW * 4;                  // [3][128] last Linear + its bias [4]
constexpr int COFF_PART = COFF_WL + 3 * CW * 4 + 16;                        // [3 upper column quarters][128][4] rgb partial sums
constexpr int COFF_SIG = COFF_PART + 3 * TC_ROWS * 16;                      // [2 tile parities][128] {sample index, sigma}: gathered by the loader warps
constexpr int COFF_BAR = COFF_SIG + 2 * TC_ROWS * 8;
constexpr int C_EPI_WARPS = 16, C_LOADER_WARP0 = 16, C_MMA_WARP = 20, C_THREADS = 21 * 32;   // warps 0-15 epilogue | 16-19 loaders | 20 MMA issuer
constexpr int C_STAGES = 4;                                // first-layer operand stages: the two activation panels, then the two ring panels
constexpr int C_NBARS = 1 + 2 * C_STAGES + 2 + 4;
constexpr int COFF_TMEMPTR = COFF_BAR + C_NBARS * 8;
constexpr int C_SMEM = COFF_TMEMPTR + 16 + 1024;
static_assert(C_SMEM <= 232448, "colour kernel exceeds the 227 KB shared memory limit");
constexpr uint32_t C_IDESC = tc_idesc(TC_ROWS, CW);
constexpr uint32_t C_IDESC_F = tc_idesc(TC_ROWS, CW, 1);     // the F stages are MN-major (sample index contiguous)
// MN-major SWIZZLE_128B A operand inside a 16 KB F stage: the two 64-sample atoms 8 KB apart (LBO), 8-feature groups 1 KB apart (SBO)
__device__ __forceinline__ uint64_t umma_desc_fstage(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(8192 >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

struct ColParams {
    const int32_t* S_ptr; int S_max;   // number of PADDED samples (every tuple tile's samples padded to a multiple of 8)
    const int32_t* csample;            // padded sample -> sample, -1 for padding
    const uint8_t* F;                  // K-weighted feature sums per padded sample, bf16 MN-major operand image (f_image_off8)
    const float* sigrow;               // w*conf*act(alpha) per tuple; sigma of a sample = sum over its nvalid consecutive tuples
    const int32_t* tuple_start; const int32_t* nvalid;
    const uint8_t* vtab; int SR;       // per ray 32 bf16: view-direction encoding (see tc_viewdir_rows_kernel)
    const uint8_t* wpack;              // packed hidden-layer weights, C_PANEL each, layer after layer
    int n_hidden;                      // colour layers followed by an activation (1..3)
    const float* bias[C_MAX_HIDDEN];
    const float* wl; const float* bl;  // last Linear [3,128], [3]
    float slope; int act_super;
    float* decoded;                    // [S,4]
    int dbg;
};

// SGN_TC_DEBUG & 4096: a wait that gives up after ~2 s, says where it was and traps (deadlock diagnosis)
__device__ __noinline__ void mbar_wait_or_report(uint32_t bar, uint32_t parity, int tag, int tile, int extra)
{
    for (int it = 0; it < 2000; it++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (ok) return;
    }
    printf("STUCK block %d thread %d tag %d tile %d extra %d parity %u\n", blockIdx.x, threadIdx.x, tag, tile, extra, parity);
    __trap();
}

__global__ void __launch_bounds__(C_THREADS, 1) agg_color_tc_kernel(const __grid_constant__ ColParams p)
{
#define CWAIT(tag, bar, par) do { if (p.dbg & 4096) mbar_wait_or_report(bar, par, tag, (int)blockIdx.x, 0); else mbar_wait(bar, par); } while (0)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = sbase + COFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    const int W_FULL = 0, R_FULL = 1, R_EMPTY = R_FULL + C_STAGES, A_FULL = R_EMPTY + C_STAGES, D_FULL = A_FULL + 2, D_EMPTY = D_FULL + 2;
    // The five first-layer operand panels of a tile (four F panels + the view-direction panel) go through FOUR stages: the tile's two
    // activation panels (free from the moment the previous tile's last layer has been multiplied until this tile's first epilogue
    // writes them) and the two ring panels, the first ring panel a second time for the view directions.  All four F panels of a tile
    // can therefore be in flight before its first MMA is due (with a two-stage ring the issuer waited ~2500 cycles per tile for them).
    auto stage_of = [](int kp) { return kp < 4 ? kp : 2; };
    auto stage_addr = [&](int st) { return sbase + (uint32_t)(st < 2 ? COFF_ACT + st * C_PANEL : COFF_RING + (st - 2) * C_PANEL); };
    uint32_t* tmem_ptr_smem = (uint32_t*)(smem + COFF_TMEMPTR);
    float* s_bias = (float*)(smem + COFF_BIAS);
    float* s_wl = (float*)(smem + COFF_WL);

    const int Sv = min(*p.S_ptr, p.S_max);
    const int ntiles = (Sv + TC_ROWS - 1) / TC_ROWS;
    const int n_wpanels = C_K0_PANELS + 2 * (p.n_hidden - 1);

    if (tid == 0) {
        mbar_init(BAR(W_FULL), 1);
        for (int s = 0; s < C_STAGES; s++) { mbar_init(BAR(R_FULL + s), 1); mbar_init(BAR(R_EMPTY + s), 1); }
        // one arrival per WARP (its lanes fence, synchronise, lane 0 arrives): hundreds of per-thread arrivals on one barrier serialise
        for (int i = 0; i < 2; i++) { mbar_init(BAR(A_FULL + i), C_EPI_WARPS / 2); mbar_init(BAR(D_FULL + i), 1); mbar_init(BAR(D_EMPTY + i), C_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < p.n_hidden * CW; i += blockDim.x) s_bias[i] = p.bias[i / CW][i % CW];
    for (int i = tid; i < 3 * CW; i += blockDim.x) s_wl[i] = p.wl[i];
    if (tid < 3) s_wl[3 * CW + tid] = p.bl[tid];
    if (warp == C_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp < C_EPI_WARPS) {
        // =========================================================== EPILOGUE: warp = (column quarter cq, TMEM lane quadrant): 32 accumulator
        // columns = half of activation panel cq / 2.  Sixteen warps: the per-layer chain TMEM -> registers -> bias, LeakyReLU -> shared
        // memory is what the issuer waits for between the layers of a tile; four warps per scheduler hide each other's latencies.
        const int quad = warp & 3, cq = warp >> 2;
        const int row = quad * 32 + lane;
        float* s_part = (float*)(smem + COFF_PART);            // [3][128][4] partial rgb sums of column quarters 1..3
        uint32_t ph_dfull[2] = {0, 0};
        uint32_t lcount = 0, tcnt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            // (the sample index and sigma of this row come from the loader warps through shared memory: their chain of dependent global
            // loads -- compact sample -> first tuple, count -> the tuples' sigma terms -- must not sit in front of the first epilogue)
            const int2* s_sig = (const int2*)(smem + COFF_SIG) + (tcnt & 1) * TC_ROWS;
            tcnt++;
            for (int l = 0; l < p.n_hidden; l++, lcount++) {
                const int db = lcount & 1;
                const bool last = (l == p.n_hidden - 1);
                CWAIT(1, BAR(D_FULL + db), ph_dfull[db]);
                ph_dfull[db] ^= 1;
                tc_fence_after();
                float o0 = 0.f, o1 = 0.f, o2 = 0.f;
                // the accumulator is handed back to the issuer as soon as this warp's columns are in registers, before the activation math
                uint32_t v[32];
                tc_ld32_nowait(tmem_base + (uint32_t)(db * CW + cq * 32) + ((uint32_t)(quad * 32) << 16), v);
                tc_wait_ld(v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(D_EMPTY + db));
                float h[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bb = *(const float4*)(s_bias + l * CW + cq * 32 + i);
                    float x0 = __uint_as_float(v[i]) + bb.x, x1 = __uint_as_float(v[i + 1]) + bb.y;
                    float x2 = __uint_as_float(v[i + 2]) + bb.z, x3 = __uint_as_float(v[i + 3]) + bb.w;
                    h[i] = fmaxf(x0, x0 * p.slope); h[i + 1] = fmaxf(x1, x1 * p.slope);
                    h[i + 2] = fmaxf(x2, x2 * p.slope); h[i + 3] = fmaxf(x3, x3 * p.slope);
                }
                if (!last) {
                    const uint32_t rowbase = sbase + COFF_ACT + (cq >> 1) * C_PANEL + row * 128;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int k = (cq & 1) * 4 + q;
                        sts128(rowbase + ((k ^ (row & 7)) << 4), pack_bf16(h[8 * q], h[8 * q + 1]), pack_bf16(h[8 * q + 2], h[8 * q + 3]),
                               pack_bf16(h[8 * q + 4], h[8 * q + 5]), pack_bf16(h[8 * q + 6], h[8 * q + 7]));
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(A_FULL + (cq >> 1)));     // 2 quarters x 4 lane quadrants = 8 warps arrive per panel
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 w0 = *(const float4*)(s_wl + cq * 32 + i);
                        const float4 w1 = *(const float4*)(s_wl + CW + cq * 32 + i);
                        const float4 w2 = *(const float4*)(s_wl + 2 * CW + cq * 32 + i);
                        o0 = fmaf(h[i], w0.x, o0); o0 = fmaf(h[i + 1], w0.y, o0); o0 = fmaf(h[i + 2], w0.z, o0); o0 = fmaf(h[i + 3], w0.w, o0);
                        o1 = fmaf(h[i], w1.x, o1); o1 = fmaf(h[i + 1], w1.y, o1); o1 = fmaf(h[i + 2], w1.z, o1); o1 = fmaf(h[i + 3], w1.w, o1);
                        o2 = fmaf(h[i], w2.x, o2); o2 = fmaf(h[i + 1], w2.y, o2); o2 = fmaf(h[i + 2], w2.z, o2); o2 = fmaf(h[i + 3], w2.w, o2);
                    }
                    // the four column quarters of a row meet in shared memory
                    if (cq > 0) { float* d = s_part + ((cq - 1) * TC_ROWS + row) * 4; d[0] = o0; d[1] = o1; d[2] = o2; }
                    asm volatile("bar.sync 2, 512;" ::: "memory");
                    if (cq == 0) {
#pragma unroll
                        for (int qq = 0; qq < 3; qq++) { const float* d = s_part + (qq * TC_ROWS + row) * 4; o0 += d[0]; o1 += d[1]; o2 += d[2]; }
                    }
                    asm volatile("bar.sync 2, 512;" ::: "memory");
                    const int2 ss = cq == 0 ? s_sig[row] : make_int2(-1, 0);
                    const int sidx = ss.x;
                    const float sg = __int_as_float(ss.y);
                    if (cq == 0 && sidx >= 0) {
                        const float s0 = 1.0f / (1.0f + __expf(-(o0 + s_wl[3 * CW]))), s1 = 1.0f / (1.0f + __expf(-(o1 + s_wl[3 * CW + 1]))),
                                    s2 = 1.0f / (1.0f + __expf(-(o2 + s_wl[3 * CW + 2])));
                        const float m = p.act_super ? 1.002f : 1.0f, o = p.act_super ? 0.001f : 0.0f;
                        ((float4*)p.decoded)[sidx] = make_float4(sg, s0 * m - o, s1 * m - o, s2 * m - o);
                    }
                }
            }
        }
    } else if (warp < C_MMA_WARP) {
        // =========================================================== LOADERS: F panels by bulk copy, view-direction panel computed
        const int lt = tid - C_LOADER_WARP0 * 32;
        uint32_t ph_empty[C_STAGES];
        for (int s = 0; s < C_STAGES; s++) ph_empty[s] = 1;
        uint32_t ltile = 0;
        // This row's view-direction encoding (32 bf16 of its ray, cols [6 fv, 32) = 0), sample index and sigma (= sum of its tuples' terms:
        // agg_tuple_tc_kernel stored w * conf * act(alpha) per tuple) are gathered here, by the loader warps, after the tile's F panels have
        // been requested: the chain of dependent global loads (compact sample -> ray / first tuple, count -> terms) must not sit in front
        // of the epilogue warps' first layer (it did: the issuer waited ~5000 cycles per tile for activations).
        const int r = lt;
        uint4 v0, v1, v2, v3;
        int sidx_r;
        float sg;
        auto gather = [&](int tile) {
            v0 = make_uint4(0u, 0u, 0u, 0u); v1 = v0; v2 = v0; v3 = v0;
            sg = 0.f;
            const int64_t c = (int64_t)tile * TC_ROWS + r;
            sidx_r = (tile < ntiles && c < Sv && !(p.dbg & 64)) ? p.csample[c] : -1;
            if (sidx_r >= 0) {
                const uint4* src = (const uint4*)(p.vtab + (size_t)(sidx_r / p.SR) * 64);
                v0 = __ldg(src); v1 = __ldg(src + 1); v2 = __ldg(src + 2); v3 = __ldg(src + 3);
                const int j0 = p.tuple_start[sidx_r], nv = p.nvalid[sidx_r];
                for (int q = 0; q < nv; q++) sg += p.sigrow[j0 + q];
            }
        };
        // SGN_TC_DEBUG & 1024 gathers one tile AHEAD instead (hides the latency completely, 8.8k instead of 11k cycles per tile) -- it
        // deadlocks after a few launches for a reason not yet found, so the default gathers in place
        const bool ahead = (p.dbg & 1024) != 0;
        if (ahead) gather(blockIdx.x);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ltile++) {
            if (lt == 0 && tile + (int)gridDim.x < ntiles) {
                const uint8_t* nxt = p.F + (size_t)(tile + gridDim.x) * F_TILE_BYTES;
                for (int i = 0; i < 4; i++)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt + i * C_PANEL), "r"((uint32_t)C_PANEL) : "memory");
            }
            // load order: the ring panels first (free since the previous tile's first layer: F2, F3 travel during its later layers), then
            // the activation panels (free once its last layer has been multiplied), then the view-direction panel
            for (int oi = 0; oi < C_K0_PANELS; oi++) {
                const int kp = oi < 2 ? oi + 2 : (oi < 4 ? oi - 2 : 4);
                const int s = stage_of(kp);
                const uint32_t base = stage_addr(s);
                if (kp < 4) {
                    // every loader thread follows every phase of the stage (a parity wait is only unambiguous one phase ahead)
                    CWAIT(2, BAR(R_EMPTY + s), ph_empty[s]);
                    if (lt == 0) {
                        if (p.dbg & 128) mbar_arrive(BAR(R_FULL + s));
                        else {
                            mbar_expect_tx(BAR(R_FULL + s), C_PANEL);
                            bulk_g2s(base, p.F + (size_t)tile * F_TILE_BYTES + kp * C_PANEL, C_PANEL, BAR(R_FULL + s));
                        }
                    }
                    ph_empty[s] ^= 1;
                } else {
                    if (!ahead) gather(tile);
                    CWAIT(3, BAR(R_EMPTY + s), ph_empty[s]); ph_empty[s] ^= 1;
                    ((int2*)(smem + COFF_SIG))[(ltile & 1) * TC_ROWS + r] = make_int2((p.dbg & 256) ? -1 : sidx_r, __float_as_int(sg));
                    sts128(base + r * 128 + ((0 ^ (r & 7)) << 4), v0.x, v0.y, v0.z, v0.w);
                    sts128(base + r * 128 + ((1 ^ (r & 7)) << 4), v1.x, v1.y, v1.z, v1.w);
                    sts128(base + r * 128 + ((2 ^ (r & 7)) << 4), v2.x, v2.y, v2.z, v2.w);
                    sts128(base + r * 128 + ((3 ^ (r & 7)) << 4), v3.x, v3.y, v3.z, v3.w);
                    fence_proxy_async();
                    __syncwarp();
                    asm volatile("barrier.sync 3, 128;" ::: "memory");
                    if (lt == 0) mbar_arrive(BAR(R_FULL + s));
                }
            }
            if (ahead) gather(tile + (int)gridDim.x);            // next tile's row data, while this tile is in the tensor pipe
        }
    } else {
        // =========================================================== MMA issuer (+ one-off resident weight load)
        if (lane == 0) {
            mbar_expect_tx(BAR(W_FULL), (uint32_t)n_wpanels * C_PANEL);
            for (int i = 0; i < n_wpanels; i++) bulk_g2s(sbase + COFF_W + i * C_PANEL, p.wpack + (size_t)i * C_PANEL, C_PANEL, BAR(W_FULL));
            CWAIT(4, BAR(W_FULL), 0);
            uint32_t ph_full[C_STAGES];
            for (int s = 0; s < C_STAGES; s++) ph_full[s] = 0;
            uint32_t ph_afull[2] = {0, 0}, ph_dempty[2] = {1, 1};
            uint32_t lcount = 0;
            const bool prof = (p.dbg & 32) != 0;                  // SGN_TC_DEBUG=32: where the issuer waits (cycles per tile)
            long long w_ring = 0, w_act = 0, w_acc = 0, t0 = 0, t_start = prof ? clock64() : 0;
            int ntl = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                ntl++;
                for (int l = 0; l < p.n_hidden; l++, lcount++) {
                    const int db = lcount & 1;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(db * CW);
                    if (prof) t0 = clock64();
                    CWAIT(5, BAR(D_EMPTY + db), ph_dempty[db]); ph_dempty[db] ^= 1;
                    if (prof) w_acc += clock64() - t0;
                    uint32_t acc = 0;
                    if (l == 0) {
                        for (int kp = 0; kp < C_K0_PANELS; kp++) {
                            const int s = stage_of(kp);
                            if (prof) t0 = clock64();
                            CWAIT(6, BAR(R_FULL + s), ph_full[s]); ph_full[s] ^= 1;
                            if (prof) w_ring += clock64() - t0;
                            tc_fence_after();
                            const uint32_t a_addr = stage_addr(s), b_addr = sbase + COFF_W + kp * C_PANEL;
                            const int ksteps = kp < 4 ? 4 : 2;
                            for (int k = 0; k < ksteps; k++) {
                                // F stages: MN-major A, one K-step = 16 features = two 1 KB feature groups; the view panel is K-major
                                if (kp < 4) tc_mma(d_tmem, umma_desc_fstage(a_addr + k * 2048), umma_desc(b_addr + k * 32), C_IDESC_F, acc);
                                else tc_mma(d_tmem, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), C_IDESC, acc);
                                acc = 1;
                            }
                            if (s >= 2) tc_commit(BAR(R_EMPTY + s));           // a ring panel is free once its MMAs have read it
                        }
                    } else {
                        for (int kp = 0; kp < 2; kp++) {
                            if (prof) t0 = clock64();
                            CWAIT(7, BAR(A_FULL + kp), ph_afull[kp]); ph_afull[kp] ^= 1;
                            if (prof) w_act += clock64() - t0;
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
                    if (l == p.n_hidden - 1) {
                        // the activation panels are free for the next tile's first two F panels once this tile's last layer has been multiplied
                        tc_commit(BAR(R_EMPTY + 0));
                        tc_commit(BAR(R_EMPTY + 1));
                    }
                }
            }
            if (prof && blockIdx.x == 0)
                printf("colour mma issuer: %lld cycles/tile (%d tiles); waiting: F ring %lld, activations (epilogue) %lld, accumulator free %lld\n",
                       (clock64() - t_start) / max(ntl, 1), ntl, w_ring / max(ntl, 1), w_act / max(ntl, 1), w_acc / max(ntl, 1));
        }
    }
    __syncthreads();
    if (warp == C_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

#undef CWAIT
// ------------------------------------------------------------------------------------------------ small kernels
// Point part of the first per-neighbour layer, hoisted out of the per-tuple work: block1.0 is linear in its input
// [emb | PE(emb) | PE(dists)], and the first 224 columns depend on the POINT only, so
//     P0[pt] = W0[:, 0:224] . [emb_pt | sin, cos of emb_pt,c * 2^f]        (fp32 accumulate, stored bf16, [N][256])
// is computed once per call for every point (N rows) instead of once per tuple (13x more rows at C1); the per-tuple kernel
// adds it to its K = 64 GEMM over PE(dists) in the first epilogue.  One CTA per SM, tile = 128 points, the four weight panels
// (cols 0..255 of W0; the operand's cols 224..255 are zero) resident in shared memory, tcgen05.mma 128x256x16, cta_group::1.
constexpr int P0_OFF_W = 0, P0_OFF_A = 4 * PANEL_B, P0_OFF_BAR = P0_OFF_A + 4 * PANEL_A, P0_SMEM = P0_OFF_BAR + 64 + 1024;
// raw_cols == 0: operand = [emb | PE(emb)] of the 32-channel embedding (224 columns, 4 panels); raw_cols > 0: operand = the table's
// raw_cols columns as they are (label embedding for block2_bpnet.0), ceil(raw_cols / 64) panels.
// rows != NULL: only the n_rows points listed there (an incremental update after point edits), tile = 128 list entries.
__global__ void __launch_bounds__(128, 1) tc_point_gemm_kernel(const float* __restrict__ emb, int raw_cols, int64_t N, const uint8_t* __restrict__ wpack0,
                                                                uint8_t* __restrict__ p0tab, const int32_t* __restrict__ rows, int64_t n_rows)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar_w = sbase + P0_OFF_BAR, bar_d = bar_w + 8;
    uint32_t* tmem_ptr_smem = (uint32_t*)(smem + P0_OFF_BAR + 16);
    if (tid == 0) {
        mbar_init(bar_w, 1); mbar_init(bar_d, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int npan = raw_cols > 0 ? (raw_cols + 63) / 64 : 4;
    if (tid == 0) {
        mbar_expect_tx(bar_w, (uint32_t)npan * PANEL_B);
        for (int i = 0; i < npan; i++) bulk_g2s(sbase + P0_OFF_W + i * PANEL_B, wpack0 + (size_t)i * PANEL_B, PANEL_B, bar_w);
    }
    const int64_t n_work = rows ? n_rows : N;
    const int64_t ntile = (n_work + TC_ROWS - 1) / TC_ROWS;
    uint32_t ph_d = 0;
    const uint32_t x0 = sbase + P0_OFF_A;
    const int row = tid;
    for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int64_t wi = tile * TC_ROWS + row;
        // the point of this row; out-of-range rows (tile padding, bad list entries) compute on zeros and store nothing
        int64_t pt = rows ? (wi < n_rows ? (int64_t)rows[wi] : N) : wi;
        if (pt < 0) pt = N;
        if (raw_cols > 0) {
            // ---- operand row: the table's columns as they are, zero padded to whole panels
            for (int q = 0; q < npan * 8; q++) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = (pt < N && 8 * q + i < raw_cols) ? __ldg(emb + pt * raw_cols + 8 * q + i) : 0.f;
                sts128(x0 + sw_off(row, 8 * q), pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
        } else {
        // ---- operand row: [emb | sin, cos of emb_c 2^f (c, f major) | 0 (cols 224..255)]
        float e[TC_C];
        if (pt < N) {
#pragma unroll
            for (int i = 0; i < TC_C / 4; i++) {
                const float4 v = __ldg((const float4*)(emb + pt * TC_C) + i);
                e[4 * i] = v.x; e[4 * i + 1] = v.y; e[4 * i + 2] = v.z; e[4 * i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < TC_C; i++) e[i] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
            sts128(x0 + sw_off(row, 8 * q), pack_bf16(e[8 * q], e[8 * q + 1]), pack_bf16(e[8 * q + 2], e[8 * q + 3]),
                   pack_bf16(e[8 * q + 4], e[8 * q + 5]), pack_bf16(e[8 * q + 6], e[8 * q + 7]));
#pragma unroll
        for (int c = 0; c < TC_C; c++) {
            float sn, cs_;
            __sincosf(e[c], &sn, &cs_);
#pragma unroll
            for (int f = 0; f < TC_F; f++) {
                sts32(x0 + sw_off(row, TC_C + 2 * (c * TC_F + f)), pack_bf16(sn, cs_));
                const float s2 = 2.0f * sn * cs_, c2 = 1.0f - 2.0f * sn * sn;
                sn = s2; cs_ = c2;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) sts128(x0 + sw_off(row, TC_PT_COLS + 8 * q), 0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            mbar_wait(bar_w, 0);                               // weights resident (completes once; later waits return at once)
            tc_fence_after();
            for (int kp = 0; kp < npan; kp++)
                for (int k = 0; k < 4; k++)
                    tc_mma(tmem_base, umma_desc(x0 + kp * PANEL_A + k * 32), umma_desc(sbase + P0_OFF_W + kp * PANEL_B + k * 32), tc_idesc(TC_ROWS, TC_W),
                           (kp | k) != 0);
            tc_commit(bar_d);
        }
        mbar_wait(bar_d, ph_d); ph_d ^= 1;
        tc_fence_after();
        // ---- epilogue: 256 fp32 accumulators of this point -> bf16 row
        uint8_t* dst = p0tab + (size_t)pt * (TC_W * 2);
#pragma unroll 1
        for (int ch = 0; ch < TC_W / 32; ch++) {
            uint32_t v[32];
            tc_ld32(tmem_base + (uint32_t)(ch * 32) + ((uint32_t)(warp * 32) << 16), v);
            if (pt < N) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    *(uint4*)(dst + ch * 64 + q * 16) = make_uint4(pack_bf16(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1])),
                                                                   pack_bf16(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])),
                                                                   pack_bf16(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])),
                                                                   pack_bf16(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
            }
        }
        tc_fence_before();
        __syncthreads();                                       // accumulator and operand tile are free again
        tc_fence_after();
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// Per-ray view-direction encoding (ori=True with the first three stripped, point_aggregators.py:579-585): 32 bf16 per ray,
// [sin(v_d 2^f) d-major (3 fv) | cos(..) (3 fv) | 0...]; the colour kernel's fifth operand panel is a copy of these rows.
__global__ void tc_viewdir_rows_kernel(const float* __restrict__ raydir, int64_t R, int fv, uint8_t* __restrict__ vtab)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * 32) return;
    const int64_t ray = i >> 5;
    const int j = (int)(i & 31);
    float v = 0.f;
    if (j < 6 * fv) {
        const int jj = j < 3 * fv ? j : j - 3 * fv;
        const int dd = jj / fv, f = jj - dd * fv;
        float sn, cs;
        sincosf(raydir[3 * ray + dd] * exp2f((float)f), &sn, &cs);
        v = j < 3 * fv ? sn : cs;
    }
    ((__nv_bfloat16*)vtab)[i] = __float2bfloat16_rn(v);
}

// Tile table: tile t owns the compact samples whose first tuple index lies in [t*TW, (t+1)*TW).  One thread per compact sample.
__global__ void tc_tile_kernel(const int32_t* __restrict__ S_ptr, int S_max, const int32_t* __restrict__ T_ptr, const int32_t* __restrict__ csample,
                               const int32_t* __restrict__ tuple_start, int TW, int2* __restrict__ tile_tab, int32_t* __restrict__ ntiles)
{
    const int Sv = min(*S_ptr, S_max);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && Sv == 0) *ntiles = 0;
    if (c >= Sv) return;
    const int st = tuple_start[csample[c]];
    const int t = st / TW;
    const int tprev = c == 0 ? -1 : tuple_start[csample[c - 1]] / TW;
    if (t != tprev) tile_tab[t] = make_int2(st, c);
    if (c == Sv - 1) {
        tile_tab[t + 1] = make_int2(*T_ptr, Sv);
        *ntiles = t + 1;
    }
}

// Padded sample index space: every tile's samples are padded to a multiple of 8 so that the K-sum drain stores whole 16-byte groups.
__global__ void tc_tile_pad_kernel(const int32_t* __restrict__ ntiles_ptr, int ncap, const int2* __restrict__ tile_tab, int32_t* __restrict__ padslots)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ncap) return;
    const int nt = min(*ntiles_ptr, ncap);
    padslots[t] = t < nt ? ((tile_tab[t + 1].y - tile_tab[t].y + 7) & ~7) : 0;
}
// padded sample -> sample (the buffer is preset to -1)
__global__ void tc_csample_pad_kernel(const int32_t* __restrict__ S_ptr, int S_max, const int32_t* __restrict__ csample, const int32_t* __restrict__ tuple_start,
                                      int TW, const int2* __restrict__ tile_tab, const int32_t* __restrict__ cpad0, int32_t* __restrict__ csample_pad)
{
    const int Sv = min(*S_ptr, S_max);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Sv) return;
    const int sidx = csample[c];
    const int t = tuple_start[sidx] / TW;
    csample_pad[cpad0[t] + (c - tile_tab[t].y)] = sidx;
}

// torch Linear weight [Nvalid, K_in] fp32 (row stride ldw) -> bf16 panels of 64 K-columns in the 128B-swizzled smem image (Nrows x 128 bytes each).
// With `bias`, columns K_in and K_in + 1 carry the bias split into two bf16 (hi + lo): the matching operand columns hold 1.0.
__global__ void tc_pack_weight_kernel(const float* __restrict__ W, int ldw, int Nrows, int Nvalid, int Kin, int npanels, const float* __restrict__ bias,
                                      uint8_t* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk (8 bf16)
    if (i >= npanels * Nrows * 8) return;
    const int ch = i & 7, n = (i >> 3) % Nrows, pnl = i / (Nrows * 8);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int k = pnl * 64 + ch * 8 + e;
        float x = 0.f;
        if (n < Nvalid) {
            if (k < Kin) x = W[(size_t)n * ldw + k];
            else if (bias && k == Kin) x = __bfloat162float(__float2bfloat16_rn(bias[n]));
            else if (bias && k == Kin + 1) x = bias[n] - __bfloat162float(__float2bfloat16_rn(bias[n]));
        }
        v[e] = x;
    }
    uint4* dst = (uint4*)(out + (size_t)pnl * Nrows * 128 + n * 128 + ((ch ^ (n & 7)) << 4));
    *dst = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// Compact bias K-step [256 x 16] bf16, K-major without swizzle (8 x 16-byte core matrices: offset(n, k) = (n/8)*256 + (k/8)*128 + (n%8)*16
// + (k%8)*2): bias hi / lo at k = 12, 13 -- the positions of the 1.0 columns inside operand K-step ONES_KSTEP of panel 4.
__global__ void tc_pack_bias_kernel(const float* __restrict__ bias, uint8_t* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TC_W * 16) return;
    const int n = i >> 4, k = i & 15;
    const float hi = __bfloat162float(__float2bfloat16_rn(bias[n]));
    const float x = k == 12 ? hi : (k == 13 ? bias[n] - hi : 0.f);
    *(__nv_bfloat16*)(out + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) = __float2bfloat16_rn(x);
}

// ------------------------------------------------------------------------------------------------ host side
constexpr int64_t TC_CHUNK = 1 << 19;     // rays per pass: bounds the worst-case (every slot valid) workspace (~21 KB per ray) and int32 indexing

struct TcWs {
    int32_t *nvalid, *svalid, *tuple_start, *sample_cidx, *partials, *tuple_src, *csample, *ntiles, *padslots, *cpad0, *csample_pad, *tpartials;
    int2* tile_tab;
    float *loc_pers, *wc, *sigrow;
    uint8_t *wpack, *cpack, *p0tab, *p0pack, *pltab, *plpack, *bpack, *apack, *F, *vtab;
};

static inline int tile_width(int K) { return TC_ROWS - (K - 1); }
// rays per pass: as few equal passes as keep each one under TC_CHUNK rays and 2^27 (sample, neighbour) slots (the worst-case workspace
// is ~110 bytes per slot: 14 GB)
static inline int64_t tc_chunk_rays(int64_t R, int SR, int K)
{
    const char* e = getenv("SGN_TC_CHUNK");                  // tests force small passes through this
    int64_t cap = e && atoll(e) > 0 ? atoll(e) : TC_CHUNK;
    const int64_t by_slots = ((int64_t)1 << 27) / ((int64_t)SR * K);
    if (!(e && atoll(e) > 0) && by_slots < cap) cap = by_slots > 1 ? by_slots : 1;
    const int64_t n = (R + cap - 1) / cap;
    return n <= 1 ? R : (R + n - 1) / n;
}

static size_t tc_carve(const AggPlan& P, int64_t N, int64_t Rc, int SR, int K, void* base, size_t cap, TcWs* ws)
{
    const AggDims& d = P.dims;
    Arena A(base, cap);
    const size_t S = (size_t)Rc * SR, T = S * K;
    ws->nvalid = A.take<int32_t>(S + 1); ws->svalid = A.take<int32_t>(S + 1);
    ws->tuple_start = A.take<int32_t>(S + 1); ws->sample_cidx = A.take<int32_t>(S + 1);
    ws->partials = A.take<int32_t>(2 * scan_partials_count((int64_t)S));
    ws->tuple_src = A.take<int32_t>(T + 1); ws->csample = A.take<int32_t>(S + 1);
    ws->ntiles = A.take<int32_t>(4);
    const size_t ncap = T / tile_width(K) + 1;            // upper bound of the number of tiles
    const size_t spad = S + 7 * ncap + 8;                  // upper bound of the padded sample count
    ws->tile_tab = A.take<int2>(ncap + 2);
    ws->padslots = A.take<int32_t>(ncap + 2); ws->cpad0 = A.take<int32_t>(ncap + 2);
    ws->tpartials = A.take<int32_t>(scan_partials_count((int64_t)ncap + 1));
    ws->csample_pad = A.take<int32_t>(spad);
    ws->loc_pers = A.take<float>(S * 3); ws->wc = A.take<float>(T);
    ws->F = A.take<uint8_t>((spad / TC_ROWS + 2) * F_TILE_BYTES); ws->sigrow = A.take<float>(T + 1);
    size_t panels = 0;
    for (int t = 0; t < P.n_tuple_layers; t++)
        panels += t == 0 ? 1 : (size_t)((P.layers[t].extra == EXTRA_LABEL ? TC_W : P.layers[t].in) + 63) / 64;
    ws->wpack = A.take<uint8_t>(panels * PANEL_B);
    ws->cpack = A.take<uint8_t>((size_t)C_W_PANELS * C_PANEL);
    ws->bpack = A.take<uint8_t>((size_t)TC_MAX_LAYERS * BIAS_PANEL_B);
    ws->apack = A.take<uint8_t>(ALPHA_PANEL_B);
    ws->vtab = A.take<uint8_t>((size_t)Rc * 64);
    ws->p0pack = A.take<uint8_t>((size_t)4 * PANEL_B);
    ws->plpack = A.take<uint8_t>((size_t)4 * PANEL_B);
    ws->p0tab = A.take<uint8_t>((size_t)(N + TC_ROWS) * (TC_W * 2));
    ws->pltab = A.take<uint8_t>(d.LD > 0 ? (size_t)(N + TC_ROWS) * (TC_W * 2) : 256);
    return A.off;
}

static int tc_supported(const AggPlan& P, int K)
{
    const AggDims& d = P.dims;
    SGN_CHECK_ARG(d.C == TC_C && d.F == TC_F && d.FD == TC_FD && d.W == TC_W,
                  "bf16 tensor-core path is built for feat_dim=32, num_feat_freqs=3, dist_xyz_freq=5, width=256 (got %d,%d,%d,%d); use SGN_PRECISION_FP32",
                  d.C, d.F, d.FD, d.W);
    SGN_CHECK_ARG(d.LD >= 0 && d.LD <= 256, "bf16 tensor-core path: label_dim must be at most 256 (got %d)", d.LD);
    SGN_CHECK_ARG(P.n_tuple_layers <= TC_MAX_LAYERS, "bf16 tensor-core path: at most %d per-neighbour layers", TC_MAX_LAYERS);
    SGN_CHECK_ARG(P.n_color_hidden >= 1 && P.n_color_hidden <= C_MAX_HIDDEN && d.WC == CW && d.FV <= 5,
                  "bf16 tensor-core path: colour branch must have 2..%d layers of width %d and num_viewdir_freqs <= 5; use SGN_PRECISION_FP32",
                  C_MAX_HIDDEN + 1, CW);
    SGN_CHECK_ARG(K <= 32, "bf16 tensor-core path: K <= 32");
    for (int t = 0; t < P.n_tuple_layers; t++)
        SGN_CHECK_ARG(P.layers[t].extra != EXTRA_LABEL || (t > 0 && P.layers[t].in == TC_W + d.LD),
                      "bf16 tensor-core path: the label embedding must enter a hidden layer next to the %d running features", TC_W);
    return SGN_OK;
}

}  // namespace sgn

using namespace sgn;

// Point cache layout: [packed W0[:, :224] (4 panels) | packed label weights (4 panels) | P0 table | label table]
static inline size_t pc_off_p0() { return (size_t)8 * PANEL_B; }
static inline size_t pc_off_pl(int64_t N) { return pc_off_p0() + align_up((size_t)(N + TC_ROWS) * (TC_W * 2)); }

int sgn_agg_tc_point_cache_bytes(const AggPlan& P, int64_t N, size_t* bytes)
{
    int rc = tc_supported(P, 1);
    if (rc) return rc;
    *bytes = pc_off_pl(N) + (P.dims.LD > 0 ? align_up((size_t)(N + TC_ROWS) * (TC_W * 2)) : 256);
    return SGN_OK;
}

int sgn_agg_tc_point_cache_build(const AggPlan& P, const float* const* weights, const SgnPointTables* tables, void* cache, size_t cache_bytes, cudaStream_t st)
{
    size_t need = 0;
    int rc = sgn_agg_tc_point_cache_bytes(P, tables->N, &need);
    if (rc) return rc;
    if (need > cache_bytes || ((uintptr_t)cache & 255)) {
        set_error("sgn_agg_point_cache_build: cache too small or misaligned (need %zu bytes, got %zu)", need, cache_bytes);
        return SGN_E_WORKSPACE;
    }
    SGN_CUDA(cudaFuncSetAttribute(tc_point_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P0_SMEM));
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    uint8_t* base = (uint8_t*)cache;
    uint8_t *p0pack = base, *plpack = base + 4 * PANEL_B;
    const int64_t ptiles = (tables->N + TC_ROWS - 1) / TC_ROWS;
    const int grid = (int)(ptiles < n_sm ? ptiles : n_sm);
    launch(tc_pack_weight_kernel, cdiv((int64_t)4 * TC_W * 8, 256), 256, 0, st, weights[0], P.layers[0].in, TC_W, TC_W, TC_PT_COLS, 4, (const float*)nullptr, p0pack);
    launch(tc_point_gemm_kernel, grid, 128, P0_SMEM, st, tables->embedding, 0, tables->N, p0pack, base + pc_off_p0(), (const int32_t*)nullptr, (int64_t)0);
    for (int t = 1; t < P.n_tuple_layers; t++)
        if (P.layers[t].extra == EXTRA_LABEL) {
            const int lp = (P.dims.LD + 63) / 64;
            launch(tc_pack_weight_kernel, cdiv((int64_t)lp * TC_W * 8, 256), 256, 0, st, weights[t] + TC_W, P.layers[t].in, TC_W, TC_W, P.dims.LD, lp,
                   (const float*)nullptr, plpack);
            launch(tc_point_gemm_kernel, grid, 128, P0_SMEM, st, tables->label_emb, P.dims.LD, tables->N, plpack, base + pc_off_pl(tables->N),
                   (const int32_t*)nullptr, (int64_t)0);
        }
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

// Rows of an existing cache recomputed for the listed points only (their embeddings changed: point edits); weights as at build time.
int sgn_agg_tc_point_cache_update(const AggPlan& P, const SgnPointTables* tables, void* cache, size_t cache_bytes, const int32_t* rows, int64_t n_rows,
                                  cudaStream_t st)
{
    size_t need = 0;
    int rc = sgn_agg_tc_point_cache_bytes(P, tables->N, &need);
    if (rc) return rc;
    if (need > cache_bytes || ((uintptr_t)cache & 255)) {
        set_error("sgn_agg_point_cache_update: cache too small or misaligned (need %zu bytes, got %zu)", need, cache_bytes);
        return SGN_E_WORKSPACE;
    }
    if (n_rows <= 0) return SGN_OK;
    SGN_CUDA(cudaFuncSetAttribute(tc_point_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P0_SMEM));
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    uint8_t* base = (uint8_t*)cache;
    const int64_t ptiles = (n_rows + TC_ROWS - 1) / TC_ROWS;
    const int grid = (int)(ptiles < n_sm ? ptiles : n_sm);
    launch(tc_point_gemm_kernel, grid, 128, P0_SMEM, st, tables->embedding, 0, tables->N, base, base + pc_off_p0(), rows, n_rows);
    for (int t = 1; t < P.n_tuple_layers; t++)
        if (P.layers[t].extra == EXTRA_LABEL)
            launch(tc_point_gemm_kernel, grid, 128, P0_SMEM, st, tables->label_emb, P.dims.LD, tables->N, base + 4 * PANEL_B, base + pc_off_pl(tables->N), rows, n_rows);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

// Optional timing of the dominant kernel (agg_tuple_tc_kernel) with CUDA events on the launching stream, for bench.py's roofline.
static bool g_tc_timing = false;
static cudaEvent_t g_tc_ev[2] = {nullptr, nullptr};
static float g_tc_ms_sum = 0.f;
static int g_tc_ms_n = 0, g_tc_pending = 0;

static void tc_timing_collect()
{
    if (g_tc_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(g_tc_ev[1]) == cudaSuccess && cudaEventElapsedTime(&ms, g_tc_ev[0], g_tc_ev[1]) == cudaSuccess) { g_tc_ms_sum += ms; g_tc_ms_n++; }
        g_tc_pending = 0;
    }
}

extern "C" int sgn_agg_kernel_timing(int enable)
{
    tc_timing_collect();
    g_tc_timing = enable != 0;
    g_tc_ms_sum = 0.f; g_tc_ms_n = 0;
    if (g_tc_timing && !g_tc_ev[0]) { SGN_CUDA(cudaEventCreate(&g_tc_ev[0])); SGN_CUDA(cudaEventCreate(&g_tc_ev[1])); }
    return SGN_OK;
}

extern "C" int sgn_agg_kernel_timing_read(float* total_ms, int* launches)
{
    tc_timing_collect();
    if (total_ms) *total_ms = g_tc_ms_sum;
    if (launches) *launches = g_tc_ms_n;
    return SGN_OK;
}

int sgn_agg_tc_workspace_bytes(const AggPlan& P, int64_t N, int64_t R, int SR, int K, size_t* bytes)
{
    int rc = tc_supported(P, K);
    if (rc) return rc;
    TcWs ws;
    *bytes = tc_carve(P, N, tc_chunk_rays(R, SR, K), SR, K, nullptr, 0, &ws);
    return SGN_OK;
}

int sgn_agg_tc_forward(const AggPlan& P, const float* const* weights, const float* const* biases, const SgnPointTables* tables,
                       const int32_t* pidx, const float* loc_w, const float* raydir, const float* campos, const float* camrotc2w,
                       int64_t R, int SR, int K, float* decoded, uint8_t* ray_valid, float* loc_pers_out, float* loc_depth, float* weight_out, float* conf_out,
                       void* workspace, size_t workspace_bytes, const void* point_cache, cudaStream_t st)
{
    int rc = tc_supported(P, K);
    if (rc) return rc;
    const AggDims& d = P.dims;
    const int64_t chunk = tc_chunk_rays(R, SR, K);
    TcWs ws;
    const size_t need = tc_carve(P, tables->N, chunk, SR, K, workspace, workspace_bytes, &ws);
    if (need > workspace_bytes || ((uintptr_t)workspace & 255)) {
        set_error("sgn_agg_forward(bf16): workspace too small or misaligned (need %zu bytes, got %zu)", need, workspace_bytes);
        return SGN_E_WORKSPACE;
    }
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    // the dynamic shared-memory opt-in is per device: once per device ordinal of this process
    static bool attr_set[64] = {};
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        SGN_CUDA(cudaFuncSetAttribute(agg_tuple_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        SGN_CUDA(cudaFuncSetAttribute(agg_tuple_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        SGN_CUDA(cudaFuncSetAttribute(agg_color_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
        SGN_CUDA(cudaFuncSetAttribute(tc_point_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P0_SMEM));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);

    // per call (weights and embeddings may have been updated by the optimiser): bf16 weight panels in the 128B-swizzled
    // shared-memory image, and the per-point operand rows [emb | PE(emb)]
    TcParams tp = {};
    tp.first_panel[0] = 0;
    for (int t = 0; t < P.n_tuple_layers; t++) {
        const LayerInfo& L = P.layers[t];
        tp.kind[t] = t == 0 ? LAYER_FROM_X0 : (L.extra == EXTRA_COLORDIR ? LAYER_FROM_ACT_E7 : LAYER_FROM_ACT);
        // the bias rides in the weight panel where the operand has spare columns (first layer, colour/dir layer), else in its own K-step
        const bool folded = tp.kind[t] != LAYER_FROM_ACT;
        uint8_t* dst = ws.wpack + (size_t)tp.first_panel[t] * PANEL_B;
        int np;
        if (t == 0) {
            // first layer: the per-tuple part only, W0[:, 224:284] + bias -> one panel; the point part W0[:, 0:224] feeds tc_point_gemm_kernel
            np = 1;
            launch(tc_pack_weight_kernel, cdiv((int64_t)np * TC_W * 8, 256), 256, 0, st, weights[t] + TC_PT_COLS, L.in, TC_W, TC_W, TC_KD, np, biases[t], dst);
            launch(tc_pack_weight_kernel, cdiv((int64_t)4 * TC_W * 8, 256), 256, 0, st, weights[t], L.in, TC_W, TC_W, TC_PT_COLS, 4, (const float*)nullptr, ws.p0pack);
        } else if (L.extra == EXTRA_LABEL) {
            // block2_bpnet.0: the running features' part W[:, 0:256] stays per tuple; the label part W[:, 256:256+LD] depends on the point
            // only and becomes a second per-point table added in this layer's epilogue
            np = TC_W / 64;
            launch(tc_pack_weight_kernel, cdiv((int64_t)np * TC_W * 8, 256), 256, 0, st, weights[t], L.in, TC_W, TC_W, TC_W, np, (const float*)nullptr, dst);
            const int lp = (d.LD + 63) / 64;
            launch(tc_pack_weight_kernel, cdiv((int64_t)lp * TC_W * 8, 256), 256, 0, st, weights[t] + TC_W, L.in, TC_W, TC_W, d.LD, lp, (const float*)nullptr, ws.plpack);
            tp.padd[t] = ws.pltab;
        } else {
            np = (L.in + 63) / 64;
            launch(tc_pack_weight_kernel, cdiv((int64_t)np * TC_W * 8, 256), 256, 0, st, weights[t], L.in, TC_W, TC_W, L.in, np,
                   folded ? biases[t] : (const float*)nullptr, dst);
        }
        tp.first_panel[t + 1] = tp.first_panel[t] + np;
        tp.bpack[t] = nullptr;
        if (!folded) {
            tp.bpack[t] = ws.bpack + (size_t)t * BIAS_PANEL_B;
            launch(tc_pack_bias_kernel, cdiv(TC_W * 16, 256), 256, 0, st, biases[t], ws.bpack + (size_t)t * BIAS_PANEL_B);
        }
    }
    ColParams cp = {};
    {
        int pnl = 0;
        for (int c = 0; c < P.n_color_hidden; c++) {
            const int l = P.color_layer0 + c;
            const int np = c == 0 ? C_K0_PANELS : 2;
            launch(tc_pack_weight_kernel, cdiv((int64_t)np * CW * 8, 256), 256, 0, st, weights[l], P.layers[l].in, CW, CW, P.layers[l].in, np, (const float*)nullptr,
                   ws.cpack + (size_t)pnl * C_PANEL);
            pnl += np;
            cp.bias[c] = biases[l];
        }
    }
    // alpha panel: per CTA of the pair four K panels of 8 rows; rank 0's row 0 is the weight vector, everything else zero
    SGN_CUDA(cudaMemsetAsync(ws.apack, 0, ALPHA_PANEL_B, st));
    launch(tc_pack_weight_kernel, cdiv((int64_t)4 * 8 * 8, 256), 256, 0, st, weights[P.alpha_layer], TC_W, 8, 1, TC_W, 4, (const float*)nullptr, ws.apack);
    // the per-point tables: from the caller's cache (inference on a static cloud with fixed weights), else computed now
    const uint8_t* p0tab = ws.p0tab;
    const uint8_t* pltab = ws.pltab;
    if (point_cache) {
        p0tab = (const uint8_t*)point_cache + pc_off_p0();
        pltab = (const uint8_t*)point_cache + pc_off_pl(tables->N);
    } else {
        const int64_t ptiles = (tables->N + TC_ROWS - 1) / TC_ROWS;
        launch(tc_point_gemm_kernel, (int)(ptiles < n_sm ? ptiles : n_sm), 128, P0_SMEM, st, tables->embedding, 0, tables->N, ws.p0pack, ws.p0tab, (const int32_t*)nullptr, (int64_t)0);
        if (d.LD > 0)
            launch(tc_point_gemm_kernel, (int)(ptiles < n_sm ? ptiles : n_sm), 128, P0_SMEM, st, tables->label_emb, d.LD, tables->N, ws.plpack, ws.pltab, (const int32_t*)nullptr, (int64_t)0);
    }
    for (int t = 1; t < P.n_tuple_layers; t++)
        if (tp.padd[t]) tp.padd[t] = pltab;
    SGN_LAUNCH_CHECK();
    tp.n_layers = P.n_tuple_layers;
    tp.wpack = ws.wpack; tp.padd[0] = p0tab;
    tp.apack = ws.apack; tp.ba = biases[P.alpha_layer];
    tp.slope = d.slope; tp.act_super = d.act_super;
    tp.K = K; tp.SR = SR;
    { const char* e = getenv("SGN_TC_DEBUG"); tp.dbg = e ? atoi(e) : 0; }
    cp.dbg = tp.dbg;
    cp.wpack = ws.cpack; cp.n_hidden = P.n_color_hidden;
    cp.wl = weights[P.n_layers - 1]; cp.bl = biases[P.n_layers - 1];
    cp.slope = d.slope; cp.act_super = d.act_super; cp.SR = SR;
    const int TW = tile_width(K);

    for (int64_t r0 = 0; r0 < R; r0 += chunk) {
        const int64_t Rc = R - r0 < chunk ? R - r0 : chunk;
        const int64_t S = Rc * SR;
        const int Tm = (int)(S * K), Sm = (int)S;
        AggIn in;
        in.tab = *tables;
        in.pidx = pidx + r0 * SR * K; in.loc_w = loc_w + r0 * SR * 3; in.raydir = raydir + r0 * 3; in.campos = campos; in.camrot = camrotc2w;
        in.smask = g_agg_sample_mask ? g_agg_sample_mask + r0 * SR : nullptr;
        float* dec = decoded + r0 * SR * 4;
        float* loc_pers = loc_pers_out ? loc_pers_out + r0 * SR * 3 : ws.loc_pers;
        SGN_CUDA(cudaMemsetAsync(dec, 0, sizeof(float) * 4 * (size_t)S, st));
        launch(agg_prepare_kernel, cdiv(S, 128), 128, 0, st, in, S, K, loc_pers, loc_depth ? loc_depth + r0 * SR : nullptr, ws.wc, (float*)nullptr, weight_out ? weight_out + r0 * SR * K : nullptr,
               conf_out ? conf_out + r0 * SR * K : nullptr, ray_valid + r0 * SR, ws.nvalid, ws.svalid);
        // compaction offsets of the tuples (scan of nvalid) and of the samples with a neighbour (scan of nvalid > 0), one pass
        if ((rc = exclusive_scan_pair_i32(ws.nvalid, ws.tuple_start, ws.sample_cidx, S, ws.partials, st))) return rc;
        const int32_t* T_ptr = ws.tuple_start + S;
        const int32_t* S_ptr = ws.sample_cidx + S;
        launch(agg_index_kernel, cdiv(S, 128), 128, 0, st, in.pidx, S, K, ws.tuple_start, ws.sample_cidx, ws.nvalid, ws.tuple_src, ws.csample, (int32_t*)nullptr, (int32_t*)nullptr);
        launch(tc_tile_kernel, cdiv(S, 256), 256, 0, st, S_ptr, Sm, T_ptr, ws.csample, ws.tuple_start, TW, ws.tile_tab, ws.ntiles);
        const int ncap = Tm / TW + 1;
        const int spad = Sm + 7 * ncap + 8;
        launch(tc_tile_pad_kernel, cdiv(ncap + 1, 256), 256, 0, st, ws.ntiles, ncap, ws.tile_tab, ws.padslots);
        if ((rc = exclusive_scan_i32(ws.padslots, ws.cpad0, ncap + 1, ws.tpartials, st))) return rc;
        SGN_CUDA(cudaMemsetAsync(ws.csample_pad, 0xFF, sizeof(int32_t) * (size_t)spad, st));
        launch(tc_csample_pad_kernel, cdiv(S, 256), 256, 0, st, S_ptr, Sm, ws.csample, ws.tuple_start, TW, ws.tile_tab, ws.cpad0, ws.csample_pad);

        tp.in = in;
        tp.ntiles_ptr = ws.ntiles; tp.ntiles_cap = ncap;
        tp.tile_tab = ws.tile_tab; tp.cpad0 = ws.cpad0;
        tp.tuple_src = ws.tuple_src; tp.sample_cidx = ws.sample_cidx;
        tp.loc_pers = loc_pers; tp.wc = ws.wc;
        tp.F = ws.F; tp.sigrow = ws.sigrow;
        // clusters of two CTAs (a TPC's SM pair); every cluster takes four tiles per cycle
        const int max_clusters = (tp.ntiles_cap + 3) / 4;
        const int n_clusters = max_clusters < n_sm / 2 ? max_clusters : n_sm / 2;
        if (g_tc_timing) { tc_timing_collect(); cudaEventRecord(g_tc_ev[0], st); }
        if (tp.dbg) launch(agg_tuple_tc_kernel<true>, 2 * n_clusters, TC_THREADS, TC_SMEM, st, tp);
        else launch(agg_tuple_tc_kernel<false>, 2 * n_clusters, TC_THREADS, TC_SMEM, st, tp);
        if (g_tc_timing) { cudaEventRecord(g_tc_ev[1], st); g_tc_pending = 1; }

        // per-sample colour MLP + rgb + (sigma, r, g, b) store
        cp.S_ptr = ws.cpad0 + (ncap + 1); cp.S_max = spad; cp.csample = ws.csample_pad; cp.F = ws.F; cp.sigrow = ws.sigrow;
        cp.tuple_start = ws.tuple_start; cp.nvalid = ws.nvalid;
        launch(tc_viewdir_rows_kernel, cdiv(Rc * 32, 256), 256, 0, st, in.raydir, Rc, d.FV, ws.vtab);
        cp.vtab = ws.vtab; cp.decoded = dec;
        const int max_ctiles = cdiv(spad, TC_ROWS);
        launch(agg_color_tc_kernel, max_ctiles < n_sm ? max_ctiles : n_sm, C_THREADS, C_SMEM, st, cp);
        SGN_LAUNCH_CHECK();
    }
    return SGN_OK;
}
