// gemm_tc.cuh -- TF32 tensor-core GEMMs (tcgen05.mma kind::tf32, fp32 accumulators in TMEM) for the layer-wise training path.
//
//   gemm_tc_nn : C[M,N] = epi( A1[M,K1] Bt1[N,K1]^T (+ A2[M,K2] Bt2[N,K2]^T) )        forward layers and dgrad
//   gemm_tc_tn : C[P,Q] += sum_m A[m,p] B[m,q]   (m split over CTAs, red.global.add)   wgrad
//
// The reference runs these contractions in cuBLAS through torch.nn.Linear (models/aggregators/point_aggregators.py:111-175);
// its pinned torch (1.10) has allow_tf32 on, so TF32 operands with fp32 accumulation is the reference's own arithmetic.
//
// Operands stay fp32 in HBM (tf32 = the top 19 bits, the tensor core ignores the rest) and go to shared memory with 16-byte
// cp.async copies whose destination addresses carry the UMMA swizzle:
//   * nn: A and Bt rows are K-contiguous -> K-major SWIZZLE_128B tiles (row = 128 B = 32 floats, 16-byte chunk ^= row & 7);
//   * tn: both operands are m-major ([m, p] rows) -> MN-major tiles; for 32-bit types the only MN-major layout of the tensor
//     core is SWIZZLE_128B_BASE32B (32-byte chunk ^= k-row & 3; MN atoms of 32 floats LBO apart, 4-row k-groups SBO apart).
// M (valid tuples / samples) is only known on the device: the kernels read it from *m_ptr and are persistent over the tiles
// (nn) or split the m range evenly over the grid (tn), so the host never synchronises.
// Warp roles, tn: 0-3 epilogue (TMEM lane quarter = warp), 4-7 cp.async loaders, 8 MMA issuer.
// nn: operands arrive by TMA tile loads (one producer thread, warp 9), warps 0-7 are the epilogue (two per lane quarter), warp 8 issues the MMAs.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include <cuda.h>                 // CUtensorMap and its enums; cuTensorMapEncodeTiled itself is fetched through cudaGetDriverEntryPoint

#include "tc_ptx.cuh"

namespace sgn {

constexpr int GT_THREADS = 288;
constexpr int GT_LAG = 2;                     // cp.async groups a loader thread keeps in flight

// ---- PTX helpers specific to these kernels
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// instruction descriptor, kind::tf32: A = B = tf32 (format 2), D = f32; a_mn / b_mn = operand is MN-major
__host__ __device__ constexpr uint32_t tc_idesc_tf32(int M, int N, int a_mn, int b_mn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major SWIZZLE_128B_BASE32B operand: MN atoms (32 floats) `lbo` bytes apart, 4-row k-groups 512 B apart (rows packed at 128 B)
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t saddr, uint32_t lbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}
// ordered with respect to the surrounding shared-memory stores (tc_ptx.cuh's lds128f is a pure asm the compiler may hoist)
__device__ __forceinline__ float4 lds128f_v(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32_v(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory"); }

// 2-D TMA tile load (cp.async.bulk.tensor): box {32 floats, rows} at element (k0, row0) of the tensor behind `map`, SWIZZLE_128B, bytes
// reported to `bar`; out-of-range elements arrive as zeros
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int k0, int row0, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(k0), "r"(row0), "r"(bar) : "memory");
}

// Row-major fp32 matrix [rows, cols] with leading dimension ld (floats) as a tensor map whose box is 32 columns x box_rows rows
static inline int make_tmap_f32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                                CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !ptr) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return SGN_E_CUDA;
        }
        fn = (EncodeFn)ptr;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a [%lld, %lld] matrix, ld %lld, box rows %d", (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows);
        return SGN_E_CUDA;
    }
    return SGN_OK;
}

// ------------------------------------------------------------------------------------------------ nn
struct GemmTcNN {
    alignas(64) CUtensorMap mA1, mB1, mA2, mB2, mC;                    // boxes: A {32, 128}, Bt {32, BN}, C {32, 32} (stores)
    const float* A1; int lda1; const float* Bt1; int ldb1; int K1;     // Bt: [N, ldb], K contiguous
    const float* A2; int lda2; const float* Bt2; int ldb2; int K2;
    float* C; int ldc; int N; int BN;                                  // BN = columns per CTA (multiple of 16, <= 256)
    const int* m_ptr; int m_max;
    const float* bias; const float* aux; int ldaux; int epi; float slope;
    float* colsum;                                                     // optional [N]: += column sums of the stored C (valid rows)
    uint32_t* mask_out; const uint32_t* mask_in; int ldmask;           // sign bits of C (see GemmNN)
};

constexpr int NN_STAGES = 4;
constexpr int NN_A_BYTES = 128 * 128, NN_B_BYTES = 256 * 128, NN_STAGE_BYTES = NN_A_BYTES + NN_B_BYTES;
constexpr int NN_THREADS = 10 * 32;                               // warps 0-7: epilogue (two per TMEM lane quarter), 8 MMA issuer, 9 TMA producer
constexpr int NN_OFF_EPI = NN_STAGES * NN_STAGE_BYTES;            // 8 warps x (32 rows x 128 B) staging
constexpr int NN_OFF_BAR = NN_OFF_EPI + 8 * 4096;
constexpr int NN_SMEM = NN_OFF_BAR + 256 + 1024;

static __global__ void __launch_bounds__(NN_THREADS, 1) gemm_tc_nn_kernel(const __grid_constant__ GemmTcNN p)
{
    const int M = min(*p.m_ptr, p.m_max);
    const int ntile = (M + 127) >> 7;
    if ((int)blockIdx.x >= ntile) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = sbase + NN_OFF_BAR, bar_empty = bar_full + 8 * NN_STAGES, bar_accf = bar_empty + 8 * NN_STAGES, bar_acce = bar_accf + 16;
    uint32_t* tmem_ptr_smem = (uint32_t*)(smem + NN_OFF_BAR + 8 * (2 * NN_STAGES + 4));
    if (tid == 0) {
        for (int s = 0; s < NN_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(bar_accf + 8 * a, 1); mbar_init(bar_acce + 8 * a, 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int n0 = blockIdx.y * p.BN;
    const int nst1 = (p.K1 + 31) >> 5, nst = nst1 + (p.A2 ? (p.K2 + 31) >> 5 : 0);

    if (warp == 9) {
        // ------------------------------------------------------------------ producer: two TMA tile loads per stage (A rows, weight rows)
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t stage_bytes = NN_A_BYTES + (uint32_t)p.BN * 128u;
            for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
                const int m0 = tile << 7;
                for (int i = 0; i < nst; i++, it++) {
                    const bool second = i >= nst1;
                    const int k0 = (second ? i - nst1 : i) << 5;
                    const uint32_t s = it % NN_STAGES, ph = (it / NN_STAGES) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t sa = sbase + s * NN_STAGE_BYTES, sb = sa + NN_A_BYTES;
                    mbar_expect_tx(bar_full + 8 * s, stage_bytes);
                    tma_load_2d(sa, second ? &p.mA2 : &p.mA1, k0, m0, bar_full + 8 * s);
                    tma_load_2d(sb, second ? &p.mB2 : &p.mB1, k0, n0, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = tc_idesc_tf32(128, p.BN, 0, 0);
            uint32_t it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x, tl++) {
                const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
                mbar_wait(bar_acce + 8 * acc, aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int i = 0; i < nst; i++, it++) {
                    const bool second = i >= nst1;
                    const int kleft = (second ? p.K2 - ((i - nst1) << 5) : p.K1 - (i << 5));
                    const int ksteps = kleft >= 32 ? 4 : kleft >> 3;
                    const uint32_t s = it % NN_STAGES, ph = (it / NN_STAGES) & 1;
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint64_t ad = umma_desc(sbase + s * NN_STAGE_BYTES), bd = umma_desc(sbase + s * NN_STAGE_BYTES + NN_A_BYTES);
                    for (int k = 0; k < ksteps; k++) tc_mma_tf32(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (i | k) != 0);
                    tc_commit(bar_empty + 8 * s);
                }
                tc_commit(bar_accf + 8 * acc);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: TMEM -> registers (one row per lane) -> swizzled staging
        // -> TMA tile store.  A lane owns row `lane` of its warp's 32 x 32 chunk: bias / LeakyReLU / the dLeakyReLU factor are applied in
        // registers, the row's 32 sign bits are one word (written for the forward layers, read by the dgrad ones), column sums (bias
        // gradients) are taken from the staged tile.  Rows past *m_ptr (but inside m_max) are stored too: nothing reads them.
        // two warps per TMEM lane quarter (a warp reaches the 32 lanes of quarter warp % 4): the first takes column chunks 0-3, the second 4-7
        const int quarter = warp & 3, half = warp >> 2;
        const uint32_t stg = sbase + NN_OFF_EPI + (quarter + 4 * half) * 4096;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};                    // column sums: chunk 4 * half + i, column `lane`
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x, tl++) {
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            const int m0 = (tile << 7) + quarter * 32;
            const int m = m0 + lane;
            mbar_wait(bar_accf + 8 * acc, aph);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int ch = 4 * half + i;
                if (ch * 32 >= p.BN) break;
                const int nc0 = n0 + ch * 32;                  // first column of the chunk
                if (nc0 >= p.N) break;
                uint32_t mword = 0u;
                if (p.epi == EPI_MUL_DLEAKY && p.mask_in && m < M) mword = __ldg(p.mask_in + (size_t)m * p.ldmask + (nc0 >> 5));
                uint32_t v[32];
                tc_ld32(tmem_base + acc * 256 + (uint32_t)(ch * 32) + ((uint32_t)(quarter * 32) << 16), v);
                float x[32];
#pragma unroll
                for (int c = 0; c < 32; c++) x[c] = __uint_as_float(v[c]);
                if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_LEAKY) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        if (nc0 + 4 * q < p.N) {
                            const float4 bb = __ldg((const float4*)(p.bias + nc0 + 4 * q));
                            x[4 * q] += bb.x; x[4 * q + 1] += bb.y; x[4 * q + 2] += bb.z; x[4 * q + 3] += bb.w;
                        }
                    }
                    if (p.epi == EPI_BIAS_LEAKY) {
#pragma unroll
                        for (int c = 0; c < 32; c++) x[c] = x[c] > 0.f ? x[c] : x[c] * p.slope;
                    }
                } else if (p.epi == EPI_MUL_DLEAKY) {
                    if (p.mask_in) {
#pragma unroll
                        for (int c = 0; c < 32; c++) x[c] *= ((mword >> c) & 1u) ? 1.f : p.slope;
                    } else if (m < M) {
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            if (nc0 + 4 * q < p.N) {
                                const float4 a = __ldg((const float4*)(p.aux + (size_t)m * p.ldaux + nc0 + 4 * q));
                                x[4 * q] *= a.x > 0.f ? 1.f : p.slope; x[4 * q + 1] *= a.y > 0.f ? 1.f : p.slope;
                                x[4 * q + 2] *= a.z > 0.f ? 1.f : p.slope; x[4 * q + 3] *= a.w > 0.f ? 1.f : p.slope;
                            }
                        }
                    }
                }
                if (p.mask_out && m < M) {
                    uint32_t word = 0u;
#pragma unroll
                    for (int c = 0; c < 32; c++) word |= (x[c] > 0.f ? 1u : 0u) << c;
                    p.mask_out[(size_t)m * p.ldmask + (nc0 >> 5)] = word;
                }
                // the staging tile is free once the previous chunk's TMA store has read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; q++)
                    sts128(stg + lane * 128 + ((q ^ (lane & 7)) << 4), __float_as_uint(x[4 * q]), __float_as_uint(x[4 * q + 1]), __float_as_uint(x[4 * q + 2]),
                           __float_as_uint(x[4 * q + 3]));
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&p.mC), "r"(nc0), "r"(m0), "r"(stg) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (p.colsum) {
                    // column `lane` of the staged tile over its valid rows (a row of the tile is 128 contiguous bytes: no bank conflicts)
                    const int rows = min(32, M - m0);
                    float a = 0.f;
                    for (int r = 0; r < rows; r++) a += lds_f32_v(stg + r * 128 + ((((uint32_t)lane >> 2) ^ ((uint32_t)r & 7)) << 4) + ((uint32_t)lane & 3) * 4);
                    cs[i] += a;
                }
            }
            tc_fence_before();
            mbar_arrive(bar_acce + 8 * acc);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (p.colsum) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int n = n0 + (4 * half + i) * 32 + lane;
                if ((4 * half + i) * 32 < p.BN && n < p.N && cs[i] != 0.f) red_add_f32(p.colsum + n, cs[i]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

static inline bool gemm_tc_nn_ok(const GemmNN& g)
{
    auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    if (g.K1 % 8 || (g.lda1 & 3) || (g.ldbt1 & 3) || !al(g.A1) || !al(g.Bt1) || !g.Bt1) return false;
    if (g.A2 && (g.K2 % 8 || (g.lda2 & 3) || (g.ldbt2 & 3) || !al(g.A2) || !al(g.Bt2) || !g.Bt2)) return false;
    if ((g.N & 3) || (g.ldc & 3) || !al(g.C)) return false;
    if (g.epi == EPI_MUL_DLEAKY && ((g.ldaux & 3) || !al(g.aux))) return false;
    if ((g.epi == EPI_BIAS || g.epi == EPI_BIAS_LEAKY) && !al(g.bias)) return false;
    if (g.colsum && !al(g.colsum)) return false;
    if ((g.mask_out || g.mask_in) && (g.N & 31)) return false;
    return true;
}

static inline int launch_gemm_tc_nn(const GemmNN& g, cudaStream_t st)
{
    if (g.m_max <= 0) return SGN_OK;
    static bool attr_set[64] = {};                        // the attribute is per device (one process may drive several)
    int dev_id = 0;
    SGN_CUDA(cudaGetDevice(&dev_id));
    if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
        SGN_CUDA(cudaFuncSetAttribute(gemm_tc_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_SMEM));
        if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
    }
    GemmTcNN p = {};
    p.A1 = g.A1; p.lda1 = g.lda1; p.Bt1 = g.Bt1; p.ldb1 = g.ldbt1; p.K1 = g.K1;
    p.A2 = g.A2; p.lda2 = g.lda2; p.Bt2 = g.Bt2; p.ldb2 = g.ldbt2; p.K2 = g.K2;
    p.C = g.C; p.ldc = g.ldc; p.N = g.N; p.m_ptr = g.m_ptr; p.m_max = g.m_max;
    p.bias = g.bias; p.aux = g.aux; p.ldaux = g.ldaux; p.epi = g.epi; p.slope = g.slope; p.colsum = g.colsum;
    p.mask_out = g.mask_out; p.mask_in = g.mask_in; p.ldmask = g.ldmask;
    const int nsplit = cdiv(g.N, 256);
    // columns per CTA: the epilogue stores whole 32-column chunks (TMA clips at N), so a split must fall on a multiple of 32
    p.BN = nsplit > 1 ? (cdiv(g.N, nsplit) + 31) / 32 * 32 : (g.N + 15) / 16 * 16;
    int rc = make_tmap_f32(&p.mA1, g.A1, g.m_max, g.K1, g.lda1, 128);
    if (!rc) rc = make_tmap_f32(&p.mB1, g.Bt1, g.N, g.K1, g.ldbt1, p.BN);
    if (!rc && g.A2) rc = make_tmap_f32(&p.mA2, g.A2, g.m_max, g.K2, g.lda2, 128);
    if (!rc && g.A2) rc = make_tmap_f32(&p.mB2, g.Bt2, g.N, g.K2, g.ldbt2, p.BN);
    if (!rc) rc = make_tmap_f32(&p.mC, g.C, g.m_max, g.N, g.ldc, 32);
    if (rc) return rc;
    const int tiles = cdiv(g.m_max, 128);
    const int gx = tiles < 148 / nsplit ? tiles : 148 / nsplit;
    launch(gemm_tc_nn_kernel, dim3(gx, nsplit), NN_THREADS, NN_SMEM, st, p);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

// ------------------------------------------------------------------------------------------------ tn (wgrad)
struct GemmTcTN {
    alignas(64) CUtensorMap mA, mB;        // boxes {32 columns, 32 m-rows}, SWIZZLE_128B_ATOM_32B: one box = one MN atom of a stage
    const float* A; int lda; int P;        // dZ  [M, lda], P columns used (multiple of 32, <= 256)
    const float* B; int ldb; int Q;        // act [M, ldb], Q columns used
    float* C; int ldc;                     // grad [P, ldc] (+=)
    const int* m_ptr; int m_max; int BQ;   // BQ = columns of B per CTA (multiple of 16, <= 256)
    int vec;                               // gradient rows are 16-byte aligned: the epilogue reduces with red.global.add.v4.f32
};

constexpr int TN_STAGES = 3;
constexpr int TN_ATOM = 32 * 128;                                  // one MN atom of a stage: 32 m-rows x 128 B
constexpr int TN_A_BYTES = 8 * TN_ATOM, TN_B_BYTES = 8 * TN_ATOM, TN_STAGE_BYTES = TN_A_BYTES + TN_B_BYTES;
constexpr int TN_OFF_EPI = TN_STAGES * TN_STAGE_BYTES;             // 4 warps x 32 x 33 floats
constexpr int TN_EPI_WARP = 32 * 33 * 4;
constexpr int TN_OFF_BAR = TN_OFF_EPI + 4 * TN_EPI_WARP + 128;
constexpr int TN_SMEM = TN_OFF_BAR + 256 + 1024;     // barriers: full[3], empty[3], done, tail, the TMEM address

static __global__ void __launch_bounds__(GT_THREADS, 1) gemm_tc_tn_kernel(const __grid_constant__ GemmTcTN p)
{
    const int M = min(*p.m_ptr, p.m_max);
    const int mpb = ((M + (int)gridDim.x - 1) / (int)gridDim.x + 31) & ~31;
    const int mb = blockIdx.x * mpb;
    if (mb >= M) return;
    const int me = min(M, mb + mpb);
    const int nst = (me - mb + 31) >> 5;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = sbase + TN_OFF_BAR, bar_empty = bar_full + 8 * TN_STAGES, bar_done = bar_empty + 8 * TN_STAGES, bar_tail = bar_done + 8;
    uint32_t* tmem_ptr_smem = (uint32_t*)(smem + TN_OFF_BAR + 8 * (2 * TN_STAGES + 2));
    if (tid == 0) {
        for (int s = 0; s < TN_STAGES; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_done, 1);
        mbar_init(bar_tail, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int q0 = blockIdx.y * p.BQ;
    const int atoms_a = p.P >> 5, atoms_b = (p.BQ + 31) >> 5;
    const int nhalf = (p.P + 127) >> 7;

    if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------------ producer (warp 4): one TMA box per MN atom (32 columns x 32 m-rows,
        // 128-byte rows, 32-byte chunk ^= row & 3: the tensor core's MN-major layout for 32-bit operands).  The CTA's last stage may reach past
        // its m range (the row count lives on the device, the tensor map only knows m_max): its extra rows are zeroed in shared memory.
        if (warp == 4) {
            const uint32_t stage_bytes = (uint32_t)(atoms_a + atoms_b) * TN_ATOM;
            for (int i = 0; i < nst; i++) {
                const uint32_t s = i % TN_STAGES, ph = (i / TN_STAGES) & 1;
                const uint32_t sa = sbase + s * TN_STAGE_BYTES, sb = sa + TN_A_BYTES;
                const int m0 = mb + (i << 5);
                const int live = min(32, me - m0);
                const uint32_t bar = live == 32 ? bar_full + 8 * s : bar_tail;
                if (lane == 0) {
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    mbar_expect_tx(bar, stage_bytes);
                    for (int a = 0; a < atoms_a; a++) tma_load_2d(sa + a * TN_ATOM, &p.mA, 32 * a, m0, bar);
                    for (int b = 0; b < atoms_b; b++) tma_load_2d(sb + b * TN_ATOM, &p.mB, q0 + 32 * b, m0, bar);
                }
                if (live < 32) {
                    mbar_wait(bar_tail, 0);
                    for (int at = 0; at < atoms_a + atoms_b; at++) {
                        const uint32_t base = at < atoms_a ? sa + at * TN_ATOM : sb + (at - atoms_a) * TN_ATOM;
                        for (int idx = lane; idx < (32 - live) * 8; idx += 32) sts128(base + (live + (idx >> 3)) * 128 + (idx & 7) * 16, 0u, 0u, 0u, 0u);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 8) {
        if (lane == 0) {
            const uint32_t idesc = tc_idesc_tf32(128, p.BQ, 1, 1);
            for (int i = 0; i < nst; i++) {
                const uint32_t s = i % TN_STAGES, ph = (i / TN_STAGES) & 1;
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint32_t sa = sbase + s * TN_STAGE_BYTES, sb = sa + TN_A_BYTES;
                for (int k = 0; k < 4; k++) {                  // K-step = 8 m-rows = 1 KB further down every atom
                    const uint64_t bd = umma_desc_mn32(sb + k * 1024, TN_ATOM);
                    for (int h = 0; h < nhalf; h++)
                        tc_mma_tf32(tmem_base + h * 256, umma_desc_mn32(sa + h * 4 * TN_ATOM + k * 1024, TN_ATOM), bd, idesc, (i | k) != 0);
                }
                tc_commit(bar_empty + 8 * s);
            }
            tc_commit(bar_done);
        }
    } else {
        if (p.vec) {
            // ------------------------------------------------------------------ epilogue: TMEM -> registers -> vector reductions.  A lane holds 32
            // consecutive columns of one gradient row: eight red.global.add.v4.f32 (a quarter of the L2 atomic operations of scalar reds; every CTA
            // of the grid adds into the same [P, Q] block, so the atomic units are what this phase waits for)
            mbar_wait(bar_done, 0);
            tc_fence_after();
            const int qend = min(p.Q, q0 + p.BQ);
            for (int h = 0; h < nhalf; h++) {
                const int pp = h * 128 + warp * 32 + lane;
                for (int ch = 0; ch * 32 < p.BQ; ch++) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + h * 256 + (uint32_t)(ch * 32) + ((uint32_t)(warp * 32) << 16), v);
                    if (pp < p.P) {
                        float* row = p.C + (size_t)pp * p.ldc;
    #pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const int q = q0 + ch * 32 + 4 * j;
                            if (q + 3 < qend && (((uintptr_t)(row + q)) & 15) == 0)
                                red_add_v4(row + q, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                            else
                                for (int e = 0; e < 4; e++)
                                    if (q + e < qend) red_add_f32(row + q + e, __uint_as_float(v[4 * j + e]));
                        }
                    }
                }
            }
    
        } else {
            // ------------------------------------------------------------------ epilogue: TMEM -> padded staging -> coalesced reductions
            const uint32_t stg = sbase + TN_OFF_EPI + warp * TN_EPI_WARP;
            mbar_wait(bar_done, 0);
            tc_fence_after();
            const int qend = min(p.Q, q0 + p.BQ);
            for (int h = 0; h < nhalf; h++) {
                const int prow0 = h * 128 + warp * 32;
                for (int ch = 0; ch * 32 < p.BQ; ch++) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + h * 256 + (uint32_t)(ch * 32) + ((uint32_t)(warp * 32) << 16), v);
    #pragma unroll
                    for (int j = 0; j < 32; j++) sts32(stg + (lane * 33 + j) * 4, v[j]);
                    __syncwarp();
                    const int q = q0 + ch * 32 + lane;
                    if (q < qend) {
    #pragma unroll 8
                        for (int r = 0; r < 32; r++) {
                            const int pp = prow0 + r;
                            if (pp < p.P) red_add_f32(p.C + (size_t)pp * p.ldc + q, ldsf(stg + (r * 33 + lane) * 4));
                        }
                    }
                    __syncwarp();
                }
            }
    
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

static inline bool gemm_tc_tn_ok(const GemmTN& t)
{
    auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    return t.P >= 32 && t.P <= 256 && t.P % 32 == 0 && (t.lda & 3) == 0 && (t.ldb & 3) == 0 && t.lda >= t.P && al(t.A) && al(t.B);
}

static inline int launch_gemm_tc_tn(const GemmTN& t, cudaStream_t st)
{
    if (t.m_max <= 0) return SGN_OK;
    static bool attr_set[64] = {};                        // the attribute is per device (one process may drive several)
    int dev_id = 0;
    SGN_CUDA(cudaGetDevice(&dev_id));
    if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
        SGN_CUDA(cudaFuncSetAttribute(gemm_tc_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM));
        if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
    }
    GemmTcTN p = {};
    p.A = t.A; p.lda = t.lda; p.P = t.P; p.B = t.B; p.ldb = t.ldb; p.Q = t.Q; p.C = t.C; p.ldc = t.ldc; p.m_ptr = t.m_ptr; p.m_max = t.m_max;
    const int nq = cdiv(t.Q, 256);
    p.BQ = (cdiv(t.Q, nq) + 15) / 16 * 16;
    p.vec = (t.ldc & 3) == 0 && ((uintptr_t)t.C & 15) == 0;
    int rc = make_tmap_f32(&p.mA, t.A, t.m_max, t.P, t.lda, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (!rc) rc = make_tmap_f32(&p.mB, t.B, t.m_max, t.ldb, t.ldb, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    int gx = 148 / nq;
    const int most = cdiv(t.m_max, 32);
    if (gx > most) gx = most;
    launch(gemm_tc_tn_kernel, dim3(gx, nq), GT_THREADS, TN_SMEM, st, p);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

}  // namespace sgn
