// query_pers.cu -- the PERSPECTIVE-frustum neural-point querier (--wcoord_query 0), SURVEY.md section 8f-4.
//
// Reference: get_occ_vox, near_vox_full, insert_vox_points, query_neigh_along_ray_layered / query_rand_along_ray and their torch glue,
// models/neural_points/query_point_indices.py:263-782 (file P below).  The grid lives in the camera's perspective coordinates
// (x/z, y/z, z), so it is rebuilt per camera; all rays of the call share it (the reference rebuilds it per 784-ray batch).
//
// What the reference computes, independent of its atomics' order: (1) occupancy = the query_size box around every voxel that holds a point
// (P:263-311); (2) per ray, the first SR occupied depths of its pixel column, ray_mask = "the column's last occupied depth is > 0"
// (P:313-365); (3) per voxel the list of its points -- with the coordinates TRUNCATED, not floored, at insertion (P:381-386), first P in
// point order, then the reservoir of P:399-405; (4) per sample the K nearest points of the kernel_size block by the layered walk of
// P:493-590 (or the random pick of P:411-490 for NN <= 0).  The reference builds lists only for the voxels "selected" by the pixel
// columns of the call: the point voxels inside the QUERY_size box of a listed sample (P:345-355 -- the launch at :656-676 hands
// query_size to that kernel -- and :695-696).  When the walk of (4) stays inside the query_size box, every voxel it can visit is selected
// by the sample's own column and the lists of ALL point voxels give the same neighbours; otherwise (kernel_size larger than query_size:
// the result then depends on which other rays are in the call) a selection pass marks the voxels exactly as the reference does.
// Not reproduced (and excluded by the oracle, oracle/query_pers_ref.c): the int8 overflow of P:696 and a max_o smaller than a column's
// selected-voxel count.
//
// B200 design: points are sorted by voxel with a stable radix sort (lists in point order without atomics), voxel -> list through a dense
// int2 volume (15 M voxels at 640x480 / vscale 2 / D 400: 123 MB, rebuilt per camera), occupancy as a bit volume whose pixel columns are
// contiguous runs of bits (a ray finds its first SR samples with a handful of ffs), one thread per ray for the samples and one per
// (ray, sample) for the neighbours.
#include <cub/device/device_radix_sort.cuh>
#include <curand_kernel.h>

#include "common.cuh"

namespace sgn {

struct PersParams {
    float sx, sy, sz, vx, vy, vz, rvx, rvy, rvz;
    int X, Y, Z, qx, qy, qz, kx, ky, kz, scx, scy;
    int P, SR, K, NN, inverse;
    float radius2, depth2;
};

__device__ __forceinline__ int64_t pers_cell(const PersParams& g, int x, int y, int z) { return ((int64_t)x * g.Y + y) * g.Z + z; }

// P:283-292: floor coordinates, early range checks; marks the point voxel
__global__ void pers_point_kernel(const float* __restrict__ xyz, int64_t n, PersParams g, uint32_t* pt_bits)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c0 = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * i], g.sx), g.vx));
    if (c0 < 0 || c0 >= g.X) return;
    const int c1 = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * i + 1], g.sy), g.vy));
    if (c1 < 0 || c1 >= g.Y) return;
    float z = xyz[3 * i + 2];
    if (g.inverse > 0) z = __fdiv_rn(1.0f, z);
    const int c2 = (int)floorf(__fdiv_rn(__fsub_rn(z, g.sz), g.vz));
    if (c2 < 0 || c2 >= g.Z) return;
    const int64_t c = pers_cell(g, c0, c1, c2);
    atomicOr(pt_bits + (c >> 5), 1u << (c & 31));
}

// P:298-309: the query_size box around every point voxel becomes occupied.  One thread per word of the point-voxel mask.
__global__ void pers_dilate_kernel(PersParams g, int64_t nwords, const uint32_t* __restrict__ pt_bits, uint32_t* occ_bits)
{
    const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwords) return;
    for (uint32_t w = pt_bits[wi]; w; w &= w - 1) {
        const int64_t c = wi * 32 + (__ffs(w) - 1);
        const int c2 = (int)(c % g.Z), c1 = (int)((c / g.Z) % g.Y), c0 = (int)(c / ((int64_t)g.Z * g.Y));
        for (int x = max(0, c0 - g.qx / 2); x < min(g.X, c0 + (g.qx + 1) / 2); x++)
            for (int y = max(0, c1 - g.qy / 2); y < min(g.Y, c1 + (g.qy + 1) / 2); y++)
                for (int z = max(0, c2 - g.qz / 2); z < min(g.Z, c2 + (g.qz + 1) / 2); z++) {
                    const int64_t cj = pers_cell(g, x, y, z);
                    const uint32_t bit = 1u << (cj & 31);
                    if (!(occ_bits[cj >> 5] & bit)) atomicOr(occ_bits + (cj >> 5), bit);
                }
    }
}

// P:381-388: the voxel a point is INSERTED into -- coordinates truncated towards zero -- provided that voxel holds a point by the floor rule
__global__ void pers_key_kernel(const float* __restrict__ xyz, int64_t n, PersParams g, const uint32_t* __restrict__ pt_bits, uint32_t* keys, int32_t* vals)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = (int)__fdiv_rn(__fsub_rn(xyz[3 * i], g.sx), g.vx);
    const int cy = (int)__fdiv_rn(__fsub_rn(xyz[3 * i + 1], g.sy), g.vy);
    float z = xyz[3 * i + 2];
    if (g.inverse > 0) z = __fdiv_rn(1.0f, z);
    const int cz = (int)__fdiv_rn(__fsub_rn(z, g.sz), g.vz);
    uint32_t key = 0xffffffffu;
    if (cx >= 0 && cx < g.X && cy >= 0 && cy < g.Y && cz >= 0 && cz < g.Z) {
        const int64_t c = pers_cell(g, cx, cy, cz);
        if ((pt_bits[c >> 5] >> (c & 31)) & 1u) key = (uint32_t)c;
    }
    keys[i] = key;
    vals[i] = (int32_t)i;
}

// per voxel (= run of equal keys in the sorted order): the P-cap reservoir in point order (P:395-405, seed = index + seconds), voxel -> list
__global__ void pers_list_kernel(const uint32_t* __restrict__ keys, int32_t* vals, int64_t n, PersParams g, uint64_t seconds, int2* cell_list)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    if (k == 0xffffffffu || (i > 0 && keys[i - 1] == k)) return;
    int m = 0;
    for (int64_t j = i; j < n && keys[j] == k; j++, m++)
        if (m >= g.P) {
            curandState state;
            curand_init((unsigned long long)vals[j] + seconds, 0, 0, &state);
            const int ins = (int)(ceilf(curand_uniform(&state) * (float)(m + 1)) - 1.0f);
            if (ins < g.P) vals[i + ins] = vals[j];
        }
    cell_list[k] = make_int2((int)i, m < g.P ? m : g.P);
}

// P:313-365 + the sample positions of P:452-463: one thread per ray
__global__ void pers_sample_kernel(PersParams g, const int32_t* __restrict__ pixel_idx, int64_t R, const uint32_t* __restrict__ occ_bits,
                                   int32_t* __restrict__ coorz, float* __restrict__ sample_loc, int8_t* __restrict__ ray_mask)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int px = pixel_idx[2 * r], py = pixel_idx[2 * r + 1];
    const int fx = px / g.scx, fy = py / g.scy;
    const int64_t col0 = ((int64_t)fx * g.Y + fy) * g.Z;
    int cnt = 0, last = -1;
    int32_t* cz = coorz + r * g.SR;
    // the column's Z bits are consecutive bits of the volume
    for (int64_t w0 = col0 >> 5; w0 <= (col0 + g.Z - 1) >> 5; w0++) {
        uint32_t w = occ_bits[w0];
        if (w0 == (col0 >> 5)) w &= 0xffffffffu << (col0 & 31);
        const int64_t endbit = col0 + g.Z - (w0 << 5);
        if (endbit < 32) w &= (1u << endbit) - 1u;
        for (; w; w &= w - 1) {
            const int z = (int)((w0 << 5) + (__ffs(w) - 1) - col0);
            if (cnt < g.SR) cz[cnt++] = z;
            last = z;
        }
    }
    for (int s = cnt; s < g.SR; s++) cz[s] = -1;
    ray_mask[r] = last > 0 ? 1 : 0;                        // far_id > 0 (P:341): a column whose only occupied depth is 0 is masked out
    // sample positions: centre of the sub-pixel in x / y, centre of the voxel in depth (fp32 FMA + double, as nvcc compiles P:457-459)
    const float cxf = (float)((double)__fmaf_rn((float)fx, g.vx, g.sx) + (px % g.scx + 0.5) * (double)g.rvx);
    const float cyf = (float)((double)__fmaf_rn((float)fy, g.vy, g.sy) + (py % g.scy + 0.5) * (double)g.rvy);
    for (int s = 0; s < g.SR; s++) {
        float czf = (float)((double)g.sz + (cz[s] + 0.5) * (double)g.vz);
        if (g.inverse > 0) czf = __fdiv_rn(1.0f, czf);
        float* sl = sample_loc + (r * g.SR + s) * 3;
        sl[0] = cxf; sl[1] = cyf; sl[2] = czf;
    }
}

// P:345-355: the point voxels inside the query_size box of every listed sample become "selected".  One thread per (ray, sample).
__global__ void pers_select_kernel(PersParams g, const int32_t* __restrict__ pixel_idx, int64_t R, const int32_t* __restrict__ coorz,
                                   const uint32_t* __restrict__ pt_bits, uint32_t* sel_bits)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * g.SR) return;
    const int d = coorz[idx];
    if (d < 0) return;
    const int64_t r = idx / g.SR;
    const int vx = pixel_idx[2 * r] / g.scx, vy = pixel_idx[2 * r + 1] / g.scy;
    for (int x = max(0, vx - g.qx / 2); x < min(g.X, vx + (g.qx + 1) / 2); x++)
        for (int y = max(0, vy - g.qy / 2); y < min(g.Y, vy + (g.qy + 1) / 2); y++)
            for (int z = max(0, d - g.qz / 2); z < min(g.Z, d + (g.qz + 1) / 2); z++) {
                const int64_t c = pers_cell(g, x, y, z);
                const uint32_t bit = 1u << (c & 31);
                if ((pt_bits[c >> 5] & bit) && !(sel_bits[c >> 5] & bit)) atomicOr(sel_bits + (c >> 5), bit);
            }
}

// P:493-590 (NN > 0) / P:411-490: one thread per (ray, sample)
template <int KT>
__global__ void pers_knn_kernel(PersParams g, const float* __restrict__ xyz, const int32_t* __restrict__ pixel_idx, int64_t R, const int32_t* __restrict__ coorz,
                                const float* __restrict__ sample_loc, const int8_t* __restrict__ ray_mask, const int32_t* __restrict__ ray_rank,
                                const int2* __restrict__ cell_list, const int32_t* __restrict__ vals, uint64_t seconds, int32_t* __restrict__ sample_pidx)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * g.SR) return;
    const int64_t r = idx / g.SR;
    int32_t* out = sample_pidx + idx * g.K;
    int32_t o[KT];
    float buf[KT];
#pragma unroll
    for (int i = 0; i < KT; i++) { o[i] = -1; buf[i] = 0.f; }
    const int fz = coorz[idx];
    if (ray_mask[r] > 0 && fz >= 0) {
        const int fx = pixel_idx[2 * r] / g.scx, fy = pixel_idx[2 * r + 1] / g.scy;
        const float cx = sample_loc[3 * idx], cy = sample_loc[3 * idx + 1], cz = sample_loc[3 * idx + 2];
        int kid = 0, far_ind = 0;
        float far2 = 0.f;
        auto visit = [&](int x, int y, int z, bool layered) {
            const int2 li = cell_list[pers_cell(g, x, y, z)];
            for (int q = li.x; q < li.x + li.y; q++) {
                const int pi = vals[q];
                const float p0 = xyz[3 * (int64_t)pi], p1 = xyz[3 * (int64_t)pi + 1], p2 = xyz[3 * (int64_t)pi + 2];
                if (layered) {
                    const float xv = g.NN < 2 ? __fsub_rn(p0, cx) : __fmaf_rn(p0, p2, -__fmul_rn(cx, cz));
                    const float yv = g.NN < 2 ? __fsub_rn(p1, cy) : __fmaf_rn(p1, p2, -__fmul_rn(cy, cz));
                    const float xy2 = __fmaf_rn(xv, xv, __fmul_rn(yv, yv));
                    const float zd = __fsub_rn(p2, cz);
                    const float z2 = __fmul_rn(zd, zd);
                    const float d2 = __fadd_rn(xy2, z2);
                    if ((g.radius2 == 0.f || xy2 <= g.radius2) && (g.depth2 == 0.f || z2 <= g.depth2)) {
                        if (kid++ < g.K) {
#pragma unroll
                            for (int i = 0; i < KT; i++) if (i == kid - 1) { o[i] = pi; buf[i] = d2; }
                            if (d2 > far2) { far2 = d2; far_ind = kid - 1; }
                        } else if (d2 < far2) {
#pragma unroll
                            for (int i = 0; i < KT; i++) if (i == far_ind) { o[i] = pi; buf[i] = d2; }
                            far2 = d2;
#pragma unroll
                            for (int i = 0; i < KT; i++) if (i < g.K && buf[i] > far2) { far2 = buf[i]; far_ind = i; }
                        }
                    }
                } else {
                    const float dx = __fsub_rn(p0, cx), dy = __fsub_rn(p1, cy), dz = __fsub_rn(p2, cz);
                    if ((g.radius2 == 0.f || __fmaf_rn(dx, dx, __fmul_rn(dy, dy)) <= g.radius2) && (g.depth2 == 0.f || __fmul_rn(dz, dz) <= g.depth2)) {
                        int slot = -1;
                        if (kid++ < g.K) slot = kid - 1;
                        else {
                            // the kernel's `index` counts the rays that passed the mask (P:688) -- ray_rank -- times SR plus the slot
                            curandState state;
                            curand_init((unsigned long long)((int64_t)ray_rank[r] * g.SR + (idx - r * g.SR)) + seconds, 0, 0, &state);
                            const int ins = (int)(ceilf(curand_uniform(&state) * (float)kid) - 1.0f);
                            if (ins < g.K) slot = ins;
                        }
#pragma unroll
                        for (int i = 0; i < KT; i++) if (i == slot) o[i] = pi;
                    }
                }
            }
        };
        if (g.NN > 0) {
            for (int layer = 0; layer < (g.kx + 1) / 2; layer++) {
                const int zlayer = min((g.kz + 1) / 2 - 1, layer);
                for (int x = max(-fx, -layer); x < min(g.X - fx, layer + 1); x++)
                    for (int y = max(-fy, -layer); y < min(g.Y - fy, layer + 1); y++)
                        for (int z = max(-fz, -zlayer); z < min(g.Z - fz, zlayer + 1); z++) {
                            if (max(abs(x), abs(y)) != layer && ((zlayer == layer) ? (abs(z) != zlayer) : true)) continue;
                            visit(fx + x, fy + y, fz + z, true);
                        }
            }
        } else {
            for (int x = max(0, fx - g.kx / 2); x < min(g.X, fx + (g.kx + 1) / 2); x++)
                for (int y = max(0, fy - g.ky / 2); y < min(g.Y, fy + (g.ky + 1) / 2); y++)
                    for (int z = max(0, fz - g.kz / 2); z < min(g.Z, fz + (g.kz + 1) / 2); z++) visit(x, y, z, false);
        }
    }
#pragma unroll
    for (int i = 0; i < KT; i++) if (i < g.K) out[i] = o[i];
}

__global__ void pers_ray_rank_kernel(const int8_t* __restrict__ ray_mask, int64_t R, int32_t* flag)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) flag[r] = ray_mask[r] > 0;
}

static size_t pers_sort_bytes(int64_t N)
{
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, (int)N, 0, 32);
    return b;
}

static int pers_check(int64_t N, int64_t R, const SgnPersCfg* c)
{
    SGN_CHECK_ARG(c != nullptr && N >= 0 && N < (1ll << 31) && R >= 0, "sgn_pers_query: bad N / R / cfg");
    SGN_CHECK_ARG(c->dim[0] > 0 && c->dim[1] > 0 && c->dim[2] > 0 && (int64_t)c->dim[0] * c->dim[1] * c->dim[2] < 0xffffffffll, "sgn_pers_query: bad grid size");
    SGN_CHECK_ARG(c->SR > 0 && c->K > 0 && c->K <= SGN_MAX_K && c->P > 0, "sgn_pers_query: bad SR / K / P");
    SGN_CHECK_ARG(c->vscale[0] > 0 && c->vscale[1] > 0, "sgn_pers_query: bad vscale");
    return SGN_OK;
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_pers_query_bytes(int64_t N, int64_t R, const SgnPersCfg* cfg, size_t* bytes)
{
    int rc = pers_check(N, R, cfg);
    if (rc) return rc;
    SGN_CHECK_ARG(bytes != nullptr, "sgn_pers_query_bytes: bytes is NULL");
    const int64_t vol = (int64_t)cfg->dim[0] * cfg->dim[1] * cfg->dim[2], nwords = (vol + 31) / 32;
    const size_t n = (size_t)(N > 0 ? N : 1), r = (size_t)(R > 0 ? R : 1);
    *bytes = align_up(pers_sort_bytes(N)) + 2 * align_up(4 * n) + 2 * align_up(4 * n) + 3 * align_up(4 * (size_t)nwords) + align_up(8 * (size_t)vol) +
             align_up(4 * r * cfg->SR) + 2 * align_up(4 * (r + 1)) + align_up(4 * scan_partials_count(R));
    return SGN_OK;
}

extern "C" int sgn_pers_query(const float* xyz_pers, int64_t N, const int32_t* pixel_idx, int64_t R, const SgnPersCfg* cfg, void* workspace,
                              size_t workspace_bytes, int32_t* sample_pidx, float* sample_loc, int8_t* ray_mask, void* stream)
{
    int rc = pers_check(N, R, cfg);
    if (rc) return rc;
    if (R == 0) return SGN_OK;
    SGN_CHECK_ARG((xyz_pers || N == 0) && pixel_idx && sample_pidx && sample_loc && ray_mask, "sgn_pers_query: NULL argument");
    size_t need;
    sgn_pers_query_bytes(N, R, cfg, &need);
    if (workspace_bytes < need || ((uintptr_t)workspace & 255)) { set_error("sgn_pers_query: workspace too small or misaligned (need %zu bytes)", need); return SGN_E_WORKSPACE; }
    auto st = (cudaStream_t)stream;
    PersParams g;
    g.sx = cfg->shift[0]; g.sy = cfg->shift[1]; g.sz = cfg->shift[2]; g.vx = cfg->vsize[0]; g.vy = cfg->vsize[1]; g.vz = cfg->vsize[2];
    g.rvx = cfg->ray_vsize[0]; g.rvy = cfg->ray_vsize[1]; g.rvz = cfg->ray_vsize[2];
    g.X = cfg->dim[0]; g.Y = cfg->dim[1]; g.Z = cfg->dim[2]; g.qx = cfg->query_size[0]; g.qy = cfg->query_size[1]; g.qz = cfg->query_size[2];
    g.kx = cfg->kernel_size[0]; g.ky = cfg->kernel_size[1]; g.kz = cfg->kernel_size[2]; g.scx = cfg->vscale[0]; g.scy = cfg->vscale[1];
    g.P = cfg->P; g.SR = cfg->SR; g.K = cfg->K; g.NN = cfg->NN; g.inverse = cfg->inverse; g.radius2 = cfg->radius2; g.depth2 = cfg->depth2;
    const int64_t vol = (int64_t)g.X * g.Y * g.Z, nwords = (vol + 31) / 32;
    Arena A(workspace, workspace_bytes);
    size_t sb = pers_sort_bytes(N);
    void* sort_tmp = A.take<char>(sb);
    uint32_t* keys = A.take<uint32_t>(N > 0 ? N : 1); uint32_t* keys2 = A.take<uint32_t>(N > 0 ? N : 1);
    int32_t* vals = A.take<int32_t>(N > 0 ? N : 1); int32_t* vals2 = A.take<int32_t>(N > 0 ? N : 1);
    uint32_t* pt_bits = A.take<uint32_t>(nwords); uint32_t* occ_bits = A.take<uint32_t>(nwords); uint32_t* sel_bits = A.take<uint32_t>(nwords);
    int2* cell_list = A.take<int2>(vol);
    int32_t* coorz = A.take<int32_t>(R * cfg->SR);
    int32_t* flag = A.take<int32_t>(R + 1); int32_t* ray_rank = A.take<int32_t>(R + 1);
    int32_t* partials = A.take<int32_t>(scan_partials_count(R));
    SGN_CUDA(cudaMemsetAsync(pt_bits, 0, 4 * (size_t)nwords, st));
    SGN_CUDA(cudaMemsetAsync(occ_bits, 0, 4 * (size_t)nwords, st));
    SGN_CUDA(cudaMemsetAsync(cell_list, 0, 8 * (size_t)vol, st));
    const int T = 256;
    // does the neighbour walk stay inside the query_size box of its sample?  (box of size s: offsets -s/2 .. (s+1)/2-1)
    bool inside = true;
    for (int a = 0; a < 3; a++) {
        const int ks = cfg->NN > 0 ? cfg->kernel_size[a == 2 ? 2 : 0] : cfg->kernel_size[a], qs = cfg->query_size[a];
        const int lo = cfg->NN > 0 ? (ks + 1) / 2 - 1 : ks / 2, hi = (ks + 1) / 2 - 1;
        inside = inside && lo <= qs / 2 && hi <= (qs + 1) / 2 - 1;
    }
    if (N > 0) {
        launch(pers_point_kernel, cdiv(N, T), T, 0, st, xyz_pers, N, g, pt_bits);
        launch(pers_dilate_kernel, cdiv(nwords, T), T, 0, st, g, nwords, pt_bits, occ_bits);
    }
    launch(pers_sample_kernel, cdiv(R, 128), 128, 0, st, g, pixel_idx, R, occ_bits, coorz, sample_loc, ray_mask);
    const uint32_t* list_bits = pt_bits;
    if (!inside && N > 0) {
        SGN_CUDA(cudaMemsetAsync(sel_bits, 0, 4 * (size_t)nwords, st));
        launch(pers_select_kernel, cdiv(R * cfg->SR, T), T, 0, st, g, pixel_idx, R, coorz, pt_bits, sel_bits);
        list_bits = sel_bits;
    }
    if (N > 0) {
        launch(pers_key_kernel, cdiv(N, T), T, 0, st, xyz_pers, N, g, list_bits, keys, vals);
        ++g_launch_count;
        SGN_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, sb, keys, keys2, vals, vals2, (int)N, 0, 32, st));
        launch(pers_list_kernel, cdiv(N, T), T, 0, st, keys2, vals2, N, g, cfg->seconds_insert, cell_list);
    }
    launch(pers_ray_rank_kernel, cdiv(R, T), T, 0, st, ray_mask, R, flag);
    rc = exclusive_scan_i32(flag, ray_rank, R, partials, st);
    if (rc) return rc;
    if (cfg->K <= 8)
        launch(pers_knn_kernel<8>, cdiv(R * cfg->SR, 128), 128, 0, st, g, xyz_pers, pixel_idx, R, coorz, sample_loc, ray_mask, ray_rank, cell_list, vals2,
               cfg->seconds_query, sample_pidx);
    else
        launch(pers_knn_kernel<SGN_MAX_K>, cdiv(R * cfg->SR, 128), 128, 0, st, g, xyz_pers, pixel_idx, R, coorz, sample_loc, ray_mask, ray_rank, cell_list,
               vals2, cfg->seconds_query, sample_pidx);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}
