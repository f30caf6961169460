// gemm_simt.cuh -- hand-written fp32 SIMT GEMMs for the strict-parity (fp32) aggregator path.
//
//   gemm_nn : C[M,N] = epi( A1[M,K1] B1[K1,N] (+ A2[M,K2] B2[K2,N]) )      forward layers and dgrad
//   gemm_tn : C[P,Q] += sum_m A[m,p] B[m,q]  (split over m, atomicAdd)      wgrad
//
// M (the number of valid neighbour tuples / samples) is only known on the device: both kernels read it
// from *m_ptr and blocks past the end exit at once, so the host never synchronises.
// 128x128x8 tiles, 256 threads, 8x8 outputs per thread, double-buffered shared memory.
#pragma once
#include "common.cuh"

namespace sgn {

enum Epilogue { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_LEAKY = 2, EPI_MUL_DLEAKY = 3 };

struct GemmNN {
    const float* A1; int lda1; const float* B1; int ldb1; int K1;
    const float* A2; int lda2; const float* B2; int ldb2; int K2;   // optional second operand pair (concat input)
    const float* Bt1; int ldbt1; const float* Bt2; int ldbt2;       // the same B operands transposed ([N, ldbt], K contiguous): what the tensor-core path reads
    float* C; int ldc; int N;
    const int* m_ptr; int m_max;
    const float* bias;        // [N] for EPI_BIAS*
    const float* aux; int ldaux;  // saved activation for EPI_MUL_DLEAKY: C *= (aux > 0 ? 1 : slope)
    int epi; float slope;
    float* colsum;            // optional [N]: += column sums of the stored C over the valid rows (bias gradient of the next wgrad); tensor-core path only
    // sign bits of the stored C, one 32-bit word per 32 columns ([M, ldmask] words): written by EPI_BIAS_LEAKY (mask_out), read instead of `aux`
    // by EPI_MUL_DLEAKY (mask_in) -- 1/32 of the bytes of the activation itself.  Tensor-core path only; N must be a multiple of 32.
    uint32_t* mask_out; const uint32_t* mask_in; int ldmask;
};

constexpr int GBM = 128, GBN = 128, GBK = 8, GTHREADS = 256, GPAD = 4;

static __global__ void __launch_bounds__(GTHREADS) gemm_nn_kernel(GemmNN p)
{
    const int M = min(*p.m_ptr, p.m_max);
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    if (m0 >= M) return;
    __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
    __shared__ __align__(16) float Bs[2][GBK][GBN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int a_row = tid >> 1, a_k = (tid & 1) * 4;
    const int b_k = tid >> 5, b_n = (tid & 31) * 4;
    const int nk1 = p.K1 / GBK, nk = nk1 + (p.A2 ? p.K2 / GBK : 0);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    auto load = [&](int kt, float4& a, float4& b) {
        const bool second = kt >= nk1;
        const float* A = second ? p.A2 : p.A1;
        const float* B = second ? p.B2 : p.B1;
        const int lda = second ? p.lda2 : p.lda1, ldb = second ? p.ldb2 : p.ldb1;
        const int k0 = (second ? kt - nk1 : kt) * GBK;
        a = make_float4(0.f, 0.f, 0.f, 0.f);
        b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + a_row < M) a = *(const float4*)(A + (size_t)(m0 + a_row) * lda + k0 + a_k);
        if (n0 + b_n < p.N) b = __ldg((const float4*)(B + (size_t)(k0 + b_k) * ldb + n0 + b_n));
    };
    auto stash = [&](int buf, const float4& a, const float4& b) {
        As[buf][a_k + 0][a_row] = a.x; As[buf][a_k + 1][a_row] = a.y;
        As[buf][a_k + 2][a_row] = a.z; As[buf][a_k + 3][a_row] = a.w;
        *(float4*)&Bs[buf][b_k][b_n] = b;
    };

    float4 ra, rb;
    load(0, ra, rb);
    stash(0, ra, rb);
    __syncthreads();
    for (int kt = 0; kt < nk; kt++) {
        const int cur = kt & 1;
        if (kt + 1 < nk) load(kt + 1, ra, rb);
#pragma unroll
        for (int k = 0; k < GBK; k++) {
            const float4 a0 = *(const float4*)&As[cur][k][ty * 4];
            const float4 a1 = *(const float4*)&As[cur][k][64 + ty * 4];
            const float4 b0 = *(const float4*)&Bs[cur][k][tx * 4];
            const float4 b1 = *(const float4*)&Bs[cur][k][64 + tx * 4];
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            stash(cur ^ 1, ra, rb);
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int n = n0 + h * 64 + tx * 4;
            if (n >= p.N) continue;
            float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
            if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_LEAKY) {
                const float4 bb = __ldg((const float4*)(p.bias + n));
                v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
                if (p.epi == EPI_BIAS_LEAKY) {
#pragma unroll
                    for (int q = 0; q < 4; q++) v[q] = v[q] > 0.f ? v[q] : v[q] * p.slope;
                }
            } else if (p.epi == EPI_MUL_DLEAKY) {
                const float4 x = *(const float4*)(p.aux + (size_t)m * p.ldaux + n);
                v[0] *= x.x > 0.f ? 1.f : p.slope; v[1] *= x.y > 0.f ? 1.f : p.slope;
                v[2] *= x.z > 0.f ? 1.f : p.slope; v[3] *= x.w > 0.f ? 1.f : p.slope;
            }
            *(float4*)(p.C + (size_t)m * p.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

static inline int launch_gemm_nn(const GemmNN& p, cudaStream_t st)
{
    if (p.m_max <= 0) return SGN_OK;
    dim3 grid(cdiv(p.m_max, GBM), cdiv(p.N, GBN));
    launch(gemm_nn_kernel, grid, GTHREADS, 0, st, p);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

// ---- wgrad: C[P,Q] += sum_m A[m,p] * B[m,q] ----
struct GemmTN {
    const float* A; int lda; int P;     // dZ   [M, lda], first P columns used
    const float* B; int ldb; int Q;     // act  [M, ldb], first Q columns used
    float* C; int ldc;                  // grad [P, ldc] (+=)
    const int* m_ptr; int m_max; int m_per_block;
};

static __global__ void __launch_bounds__(GTHREADS) gemm_tn_kernel(GemmTN p)
{
    const int M = min(*p.m_ptr, p.m_max);
    const int mb = blockIdx.z * p.m_per_block;
    if (mb >= M) return;
    const int me = min(M, mb + p.m_per_block);
    const int p0 = blockIdx.x * GBM, q0 = blockIdx.y * GBN;
    __shared__ __align__(16) float As[2][GBK][GBM];
    __shared__ __align__(16) float Bs[2][GBK][GBN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int l_m = tid >> 5, l_c = (tid & 31) * 4;
    const bool vecA = (p.lda & 3) == 0 && (p.P & 3) == 0, vecB = (p.ldb & 3) == 0 && (p.Q & 3) == 0;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    auto load1 = [&](const float* X, int ld, int lim, int c0, bool vec, int m) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < me) {
            const float* src = X + (size_t)m * ld + c0 + l_c;
            if (vec) {
                if (c0 + l_c < lim) v = *(const float4*)src;
            } else {
                if (c0 + l_c + 0 < lim) v.x = src[0];
                if (c0 + l_c + 1 < lim) v.y = src[1];
                if (c0 + l_c + 2 < lim) v.z = src[2];
                if (c0 + l_c + 3 < lim) v.w = src[3];
            }
        }
        return v;
    };

    float4 ra = load1(p.A, p.lda, p.P, p0, vecA, mb + l_m), rb = load1(p.B, p.ldb, p.Q, q0, vecB, mb + l_m);
    *(float4*)&As[0][l_m][l_c] = ra;
    *(float4*)&Bs[0][l_m][l_c] = rb;
    __syncthreads();
    const int nt = (me - mb + GBK - 1) / GBK;
    for (int t = 0; t < nt; t++) {
        const int cur = t & 1;
        if (t + 1 < nt) {
            ra = load1(p.A, p.lda, p.P, p0, vecA, mb + (t + 1) * GBK + l_m);
            rb = load1(p.B, p.ldb, p.Q, q0, vecB, mb + (t + 1) * GBK + l_m);
        }
#pragma unroll
        for (int k = 0; k < GBK; k++) {
            const float4 a0 = *(const float4*)&As[cur][k][ty * 4];
            const float4 a1 = *(const float4*)&As[cur][k][64 + ty * 4];
            const float4 b0 = *(const float4*)&Bs[cur][k][tx * 4];
            const float4 b1 = *(const float4*)&Bs[cur][k][64 + tx * 4];
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (t + 1 < nt) {
            *(float4*)&As[cur ^ 1][l_m][l_c] = ra;
            *(float4*)&Bs[cur ^ 1][l_m][l_c] = rb;
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int pp = p0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (pp >= p.P) continue;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int qq = q0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (qq < p.Q) atomicAdd(p.C + (size_t)pp * p.ldc + qq, acc[i][j]);
        }
    }
}

static inline int launch_gemm_tn(GemmTN p, cudaStream_t st)
{
    if (p.m_max <= 0) return SGN_OK;
    // enough m-splits to fill the machine a few times over, at least 256 rows per block
    const int tiles = cdiv(p.P, GBM) * cdiv(p.Q, GBN);
    int splits = (148 * 4 + tiles - 1) / tiles;
    int mpb = cdiv(p.m_max, splits);
    mpb = ((mpb < 256 ? 256 : mpb) + GBK - 1) / GBK * GBK;
    p.m_per_block = mpb;
    dim3 grid(cdiv(p.P, GBM), cdiv(p.Q, GBN), cdiv(p.m_max, mpb));
    launch(gemm_tn_kernel, grid, GTHREADS, 0, st, p);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

}  // namespace sgn
