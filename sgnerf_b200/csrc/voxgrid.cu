// voxgrid.cu -- the two voxel-grid helpers around the render path that the world-coordinate querier does not cover:
//
//   sgn_voxel_downsample   point-cloud initialisation: one point per occupied voxel of a vox_res^3 grid, the one closest to the voxel's
//                          centroid.  Reference: construct_vox_points_closest, models/mvs/mvs_utils.py:536-561 (torch.unique over the
//                          voxel coordinates + scatter_mean + scatter_min), called from run/train_ft.py:141, :715.
//   sgn_query_vox_grid     NN < 0 "grid" query: the 8 corners of the grid cell a shading sample falls in, all -1 unless the 8 exist.
//                          Reference: NeuralPoints.query_vox_grid, models/neural_points/neural_points.py:814-826.
//
// Down-sampling is a sort: keys = linearised voxel coordinates (their order is the lexicographic order torch.unique(dim=0) returns),
// values = point indices; cub's radix sort is stable, so the points of a voxel stay in index order and the centroid sums and the
// first-minimum rule are deterministic.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace sgn {

struct VoxParams {
    float mnx, mny, mnz, sx, sy, sz;
    int res;          // coordinates lie in [0, res]
};

__global__ void vox_key_kernel(const float* __restrict__ xyz, int64_t n, VoxParams g, uint64_t* keys, int32_t* vals)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // floor((xyz - space_min) / construct_vox_sz) with separate IEEE operations, as torch evaluates it (:548-549)
    const int cx = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * i], g.mnx), g.sx));
    const int cy = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * i + 1], g.mny), g.sy));
    const int cz = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * i + 2], g.mnz), g.sz));
    // lexicographic order of (x, y, z) incl. negative coordinates: bias every coordinate into 21 unsigned bits
    const uint64_t B = 1u << 20;
    keys[i] = ((uint64_t)(cx + (int64_t)B) << 42) | ((uint64_t)(cy + (int64_t)B) << 21) | (uint64_t)(cz + (int64_t)B);
    vals[i] = (int32_t)i;
}

__global__ void vox_head_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* head)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per sorted element that starts a voxel: walk the voxel's points (in index order)
__global__ void vox_reduce_kernel(const float* __restrict__ xyz, const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                  const int32_t* __restrict__ head, const int32_t* __restrict__ vid, int64_t n, float* __restrict__ centroid,
                                  int32_t* __restrict__ grid_idx, int64_t* __restrict__ min_idx, int32_t* count)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = vid[n];
    if (i >= n || !head[i]) return;
    const int v = vid[i];
    const uint64_t k = keys[i];
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int m = 0;
    for (int64_t j = i; j < n && keys[j] == k; j++, m++) {
        const int64_t p = vals[j];
        sx += xyz[3 * p]; sy += xyz[3 * p + 1]; sz += xyz[3 * p + 2];
    }
    const float cx = sx / (float)m, cy = sy / (float)m, cz = sz / (float)m;
    float best = INFINITY;
    int64_t bi = vals[i];
    for (int64_t j = i; j < i + m; j++) {
        const int64_t p = vals[j];
        const float dx = xyz[3 * p] - cx, dy = xyz[3 * p + 1] - cy, dz = xyz[3 * p + 2] - cz;
        const float r = sqrtf(dx * dx + dy * dy + dz * dz);
        if (r < best) { best = r; bi = p; }
    }
    centroid[3 * v] = cx; centroid[3 * v + 1] = cy; centroid[3 * v + 2] = cz;
    const int64_t B = 1 << 20;
    grid_idx[3 * v] = (int)((int64_t)(k >> 42) - B); grid_idx[3 * v + 1] = (int)((int64_t)((k >> 21) & 0x1fffff) - B);
    grid_idx[3 * v + 2] = (int)((int64_t)(k & 0x1fffff) - B);
    min_idx[v] = bi;
}

__global__ void query_vox_grid_kernel(const float* __restrict__ loc_w, int64_t S, const int32_t* __restrict__ full_grid, int G,
                                      float mnx, float mny, float mnz, float vsz, int64_t* __restrict__ out)
{
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    // floor((loc - space_min) / grid_vox_sz) (:817); the 8 corners in the reference's order (:818)
    const long long vx = (long long)floorf(__fdiv_rn(__fsub_rn(loc_w[3 * s], mnx), vsz)), vy = (long long)floorf(__fdiv_rn(__fsub_rn(loc_w[3 * s + 1], mny), vsz)),
                    vz = (long long)floorf(__fdiv_rn(__fsub_rn(loc_w[3 * s + 2], mnz), vsz));
    const int dx[8] = {0, 1, 0, 0, 1, 0, 1, 1}, dy[8] = {0, 0, 1, 0, 0, 1, 1, 1}, dz[8] = {0, 0, 0, 1, 1, 1, 0, 1};
    int64_t idx[8];
    bool bad = false;
    const int64_t E = (int64_t)G + 1;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        long long x = vx + dx[c], y = vy + dy[c], z = vz + dz[c];
        if (x < 0 || x > G || y < 0 || y > G || z < 0 || z > G) bad = true;         // :820 (any corner outside [0, grid_res])
        x = x < 0 ? 0 : (x > G ? G : x); y = y < 0 ? 0 : (y > G ? G : y); z = z < 0 ? 0 : (z > G ? G : z);
        idx[c] = full_grid[(x * E + y) * E + z];
        if (idx[c] < 0) bad = true;                                                  // :825 (-1 for all 8 corners)
    }
#pragma unroll
    for (int c = 0; c < 8; c++) out[8 * s + c] = bad ? -1 : idx[c];
}

static size_t vox_sort_bytes(int64_t N)
{
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, (int)N, 0, 63);
    return b;
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_voxel_downsample_bytes(int64_t N, size_t* bytes)
{
    SGN_CHECK_ARG(bytes && N >= 0 && N < (1ll << 31), "sgn_voxel_downsample_bytes: bad argument");
    const size_t n = (size_t)(N > 0 ? N : 1);
    *bytes = align_up(vox_sort_bytes(N)) + 2 * align_up(8 * n) + 2 * align_up(4 * n) + 2 * align_up(4 * (n + 1)) + align_up(4 * scan_partials_count(N));
    return SGN_OK;
}

extern "C" int sgn_voxel_downsample(const float* xyz, int64_t N, const float* space_min, const float* vox_size, int vox_res, void* workspace,
                                    size_t workspace_bytes, float* centroid, int32_t* grid_idx, int64_t* min_idx, int32_t* count, void* stream)
{
    SGN_CHECK_ARG(xyz && space_min && vox_size && centroid && grid_idx && min_idx && count && N >= 0 && N < (1ll << 31), "sgn_voxel_downsample: bad argument");
    SGN_CHECK_ARG(vox_res > 0 && vox_res < (1 << 19), "sgn_voxel_downsample: vox_res out of range");
    size_t need;
    sgn_voxel_downsample_bytes(N, &need);
    if (workspace_bytes < need || ((uintptr_t)workspace & 255)) { set_error("sgn_voxel_downsample: workspace too small or misaligned (need %zu bytes)", need); return SGN_E_WORKSPACE; }
    auto st = (cudaStream_t)stream;
    if (N == 0) { SGN_CUDA(cudaMemsetAsync(count, 0, 4, st)); return SGN_OK; }
    Arena A(workspace, workspace_bytes);
    size_t sb = vox_sort_bytes(N);
    void* sort_tmp = A.take<char>(sb);
    uint64_t* keys = A.take<uint64_t>(N); uint64_t* keys2 = A.take<uint64_t>(N);
    int32_t* vals = A.take<int32_t>(N); int32_t* vals2 = A.take<int32_t>(N);
    int32_t* head = A.take<int32_t>(N + 1); int32_t* vid = A.take<int32_t>(N + 1);
    int32_t* partials = A.take<int32_t>(scan_partials_count(N));
    VoxParams g = {space_min[0], space_min[1], space_min[2], vox_size[0], vox_size[1], vox_size[2], vox_res};
    launch(vox_key_kernel, cdiv(N, 256), 256, 0, st, xyz, N, g, keys, vals);
    ++g_launch_count;
    SGN_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, sb, keys, keys2, vals, vals2, (int)N, 0, 63, st));
    launch(vox_head_kernel, cdiv(N, 256), 256, 0, st, keys2, N, head);
    int rc = exclusive_scan_i32(head, vid, N, partials, st);
    if (rc) return rc;
    launch(vox_reduce_kernel, cdiv(N, 128), 128, 0, st, xyz, keys2, vals2, head, vid, N, centroid, grid_idx, min_idx, count);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_query_vox_grid(const float* sample_loc_w, int64_t n_samples, const int32_t* full_grid_idx, int grid_res, const float* space_min,
                                  float grid_vox_sz, int64_t* out, void* stream)
{
    SGN_CHECK_ARG(sample_loc_w && full_grid_idx && space_min && out && n_samples >= 0 && grid_res > 0 && grid_vox_sz > 0.f, "sgn_query_vox_grid: bad argument");
    if (n_samples == 0) return SGN_OK;
    launch(query_vox_grid_kernel, cdiv(n_samples, 256), 256, 0, (cudaStream_t)stream, sample_loc_w, n_samples, full_grid_idx, grid_res, space_min[0],
           space_min[1], space_min[2], grid_vox_sz, out);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}
