// grid.cu -- occupancy grid build: a counting sort of the points by voxel that reproduces the
// reference's *sequential-order* result without depending on atomic arrival order.
//
// Reference: claim_occ / map_coor2occ / fill_occ2pnts and their driver build_occ_vox
// (models/neural_points/query_point_indices_worldcoords.py:265-410, :706-778).
//
// Canonical semantics (threads of each reference kernel run one after another in index order):
//   slot(voxel)   = rank of the voxel among all occupied voxels ordered by their smallest point index
//                   (first visitor claims the slot, :295-311); beyond max_o a reservoir pick
//                   j = ceilf(curand_uniform(seed = i + 2*seconds) * (slot+1)) - 1 overwrites record j
//                   (:313-321), last writer in index order wins;
//   list(voxel)   = its points in increasing index order, first P kept; point number m >= P replaces
//                   entry j = ceilf(u * (m+1)) - 1 if j < P (:397-406);
//   slot 0        never receives points (`voxel_idx > 0`, :395);
//   occupancy     = 1 on the query_size box around every surviving record (:353-360).
//
// Data layout in HBM (persistent): cell_slot int32[X*Y*Z] (dense, -1 = empty), occ_bits (1 bit / voxel),
// slot_start int32[max_o+1], cand float4[<=N] = (x, y, z, bits(point index)) grouped by slot so the
// K-NN kernel streams a voxel's candidates as consecutive 16-byte loads.
#include <curand_kernel.h>

#include "common.cuh"
#include "grid.cuh"


namespace sgn {

struct GridParams {
    float ox, oy, oz, vx, vy, vz;
    int dx, dy, dz;
    int qx, qy, qz;
    int P, max_o;
};

__device__ __forceinline__ bool point_cell(const float* __restrict__ xyz, int64_t i, const GridParams& g, int& cx, int& cy, int& cz,
                                           int64_t& cell)
{
    cx = vox_coord(xyz[3 * i + 0], g.ox, g.vx);
    cy = vox_coord(xyz[3 * i + 1], g.oy, g.vy);
    cz = vox_coord(xyz[3 * i + 2], g.oz, g.vz);
    if (cx < 0 || cx >= g.dx || cy < 0 || cy >= g.dy || cz < 0 || cz >= g.dz) return false;
    cell = ((int64_t)cx * g.dy + cy) * g.dz + cz;
    return true;
}

// reservoir index of :315 / :403
__device__ __forceinline__ int reservoir_index(int64_t index, uint64_t seconds, int tmp)
{
    curandState state;
    curand_init((unsigned long long)index + 2ull * seconds, 0, 0, &state);
    return (int)(ceilf(curand_uniform(&state) * (float)(tmp + 1)) - 1.0f);
}

// K1: smallest point index per voxel (cell_slot pre-filled with 0xFFFFFFFF)
__global__ void cell_min_kernel(const float* __restrict__ xyz, int64_t n, GridParams g, uint32_t* cell_min)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx, cy, cz;
    int64_t cell;
    if (point_cell(xyz, i, g, cx, cy, cz, cell)) atomicMin(cell_min + cell, (uint32_t)i);
}

// K2: flag[i] = 1 iff point i is the first visitor of its voxel
__global__ void first_flag_kernel(const float* __restrict__ xyz, int64_t n, int64_t n_total, GridParams g,
                                  const uint32_t* __restrict__ cell_min, int32_t* flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    int f = 0;
    if (i < n) {
        int cx, cy, cz;
        int64_t cell;
        if (point_cell(xyz, i, g, cx, cy, cz, cell)) f = cell_min[cell] == (uint32_t)i;
    }
    flag[i] = f;
}

// K3a: every first visitor bids for a record: its own slot, or a reservoir pick once slot >= max_o.
//      record_writer[j] = largest bidding point index (= last writer in sequential order).
__global__ void claim_bid_kernel(int64_t n, const int32_t* __restrict__ flag, const int32_t* __restrict__ rank, int max_o,
                                 uint64_t seconds, int32_t* record_writer, int32_t* counters)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) counters[0] = rank[n];  // occ_idx: number of claimed voxels
    if (i >= n || !flag[i]) return;
    int slot = rank[i];
    int j = slot < max_o ? slot : reservoir_index(i, seconds, slot);
    if (j < max_o) atomicMax(record_writer + j, (int32_t)i);
}

// K3b: winners write their record and publish cell -> slot; losers' voxels end up with no slot (-1),
//      exactly like the reference after coor_2_occ is reset (:735) and refilled from the records (:343-351).
__global__ void claim_resolve_kernel(const float* __restrict__ xyz, int64_t n, GridParams g, const int32_t* __restrict__ flag,
                                     const int32_t* __restrict__ rank, uint64_t seconds, const int32_t* __restrict__ record_writer,
                                     int32_t* cell_slot, int32_t* slot_coor)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx, cy, cz;
    int64_t cell;
    if (!point_cell(xyz, i, g, cx, cy, cz, cell)) return;
    if (!flag[i]) return;
    int slot = rank[i];
    int j = slot < g.max_o ? slot : reservoir_index(i, seconds, slot);
    if (j < g.max_o && record_writer[j] == (int32_t)i) {
        slot_coor[3 * j + 0] = cx; slot_coor[3 * j + 1] = cy; slot_coor[3 * j + 2] = cz;
        cell_slot[cell] = j;
    } else {
        cell_slot[cell] = -1;
    }
}

// K4: points per slot (uncapped, like occ_numpnts); slot 0 is skipped (`voxel_idx > 0`)
__global__ void slot_count_kernel(const float* __restrict__ xyz, int64_t n, GridParams g, const int32_t* __restrict__ cell_slot,
                                  int32_t* slot_count)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx, cy, cz;
    int64_t cell;
    if (!point_cell(xyz, i, g, cx, cy, cz, cell)) return;
    int s = cell_slot[cell];
    if (s > 0) atomicAdd(slot_count + s, 1);
}

// K5: scatter point indices into their slot's segment (arrival order, fixed by K6)
__global__ void slot_fill_kernel(const float* __restrict__ xyz, int64_t n, GridParams g, const int32_t* __restrict__ cell_slot,
                                 const int32_t* __restrict__ seg_start, int32_t* cursor, int32_t* seg)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx, cy, cz;
    int64_t cell;
    if (!point_cell(xyz, i, g, cx, cy, cz, cell)) return;
    int s = cell_slot[cell];
    if (s > 0) seg[seg_start[s] + atomicAdd(cursor + s, 1)] = (int32_t)i;
}

__device__ void sift_down(int32_t* a, int start, int end)
{
    int root = start;
    while (2 * root + 1 <= end) {
        int child = 2 * root + 1, sw = root;
        if (a[sw] < a[child]) sw = child;
        if (child + 1 <= end && a[sw] < a[child + 1]) sw = child + 1;
        if (sw == root) return;
        int t = a[root]; a[root] = a[sw]; a[sw] = t;
        root = sw;
    }
}

// K6: per slot: sort the segment by point index, apply the P-cap reservoir in index order,
//     ncap[s] = min(count, P).  One thread per slot; segments are a handful of points.
__global__ void slot_canon_kernel(int max_o, int P, uint64_t seconds, const int32_t* __restrict__ slot_count,
                                  const int32_t* __restrict__ seg_start, int32_t* seg, int32_t* ncap)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= max_o) return;
    int n = slot_count[s];
    int32_t* a = seg + seg_start[s];
    if (n <= 48) {
        for (int i = 1; i < n; i++) {
            int v = a[i], j = i - 1;
            while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
            a[j + 1] = v;
        }
    } else {  // heapsort: O(n log n), in place
        for (int st = (n - 2) / 2; st >= 0; st--) sift_down(a, st, n - 1);
        for (int end = n - 1; end > 0; end--) {
            int t = a[end]; a[end] = a[0]; a[0] = t;
            sift_down(a, 0, end - 1);
        }
    }
    for (int m = P; m < n; m++) {
        int j = reservoir_index(a[m], seconds, m);
        if (j < P) a[j] = a[m];
    }
    ncap[s] = n < P ? n : P;
}

// K7: candidate records (x, y, z, index) grouped by slot
__global__ void emit_cand_kernel(int max_o, const float* __restrict__ xyz, const int32_t* __restrict__ seg_start,
                                 const int32_t* __restrict__ seg, const int32_t* __restrict__ slot_start, float4* cand,
                                 int32_t* counters)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) counters[1] = slot_start[max_o];
    if (s >= max_o) return;
    int b = slot_start[s], n = slot_start[s + 1] - b;
    const int32_t* a = seg + seg_start[s];
    for (int g = 0; g < n; g++) {
        int p = a[g];
        cand[b + g] = make_float4(xyz[3 * (int64_t)p], xyz[3 * (int64_t)p + 1], xyz[3 * (int64_t)p + 2], __int_as_float(p));
    }
}

// K8: occupancy bits on the query_size box around every surviving record (:353-360)
__global__ void dilate_kernel(GridParams g, const int32_t* __restrict__ counters, const int32_t* __restrict__ slot_coor, uint32_t* occ_bits,
                              uint32_t* coarse_bits)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    int n_rec = counters[0] < g.max_o ? counters[0] : g.max_o;
    if (s >= n_rec) return;
    int c0 = slot_coor[3 * s];
    if (c0 < 0) return;
    int c1 = slot_coor[3 * s + 1], c2 = slot_coor[3 * s + 2];
    int x0 = max(0, c0 - g.qx / 2), x1 = min(g.dx, c0 + (g.qx + 1) / 2);
    int y0 = max(0, c1 - g.qy / 2), y1 = min(g.dy, c1 + (g.qy + 1) / 2);
    int z0 = max(0, c2 - g.qz / 2), z1 = min(g.dz, c2 + (g.qz + 1) / 2);
    for (int x = x0; x < x1; x++)
        for (int y = y0; y < y1; y++)
            for (int z = z0; z < z1; z++) {
                int64_t c = ((int64_t)x * g.dy + y) * g.dz + z;
                uint32_t bit = 1u << (c & 31);
                if (!(occ_bits[c >> 5] & bit)) atomicOr(occ_bits + (c >> 5), bit);
            }
    // coarse mask: every 8^3 brick that holds one of these voxels or a voxel adjacent to one (the one-voxel margin absorbs the rounding
    // of the approximate coordinate march_kernel tests the brick with)
    const int cdx = (g.dx + 7) >> 3, cdy = (g.dy + 7) >> 3, cdz = (g.dz + 7) >> 3;
    const int bx0 = max(0, (x0 - 1) >> 3), bx1 = min(cdx - 1, x1 >> 3);
    const int by0 = max(0, (y0 - 1) >> 3), by1 = min(cdy - 1, y1 >> 3);
    const int bz0 = max(0, (z0 - 1) >> 3), bz1 = min(cdz - 1, z1 >> 3);
    for (int x = bx0; x <= bx1; x++)
        for (int y = by0; y <= by1; y++)
            for (int z = bz0; z <= bz1; z++) {
                const int c = (x * cdy + y) * cdz + z;
                const uint32_t bit = 1u << (c & 31);
                if (!(coarse_bits[c >> 5] & bit)) atomicOr(coarse_bits + (c >> 5), bit);
            }
}

// K9-K11: the K-NN index.  A voxel is "listed" when it owns a surviving record with at least one candidate (slot 0 never does, :395).
__device__ __forceinline__ void brick_of(const GridParams& g, int c0, int c1, int c2, int64_t& b, int& bit)
{
    b = ((int64_t)(c0 >> 2) * brick4(g.dy) + (c1 >> 2)) * brick4(g.dz) + (c2 >> 2);
    bit = ((c0 & 3) * 4 + (c1 & 3)) * 4 + (c2 & 3);
}

__global__ void brick_mark_kernel(GridParams g, const int32_t* __restrict__ counters, const int32_t* __restrict__ slot_coor,
                                  const int32_t* __restrict__ slot_start, uint4* knn_brick)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    int n_rec = counters[0] < g.max_o ? counters[0] : g.max_o;
    if (s >= n_rec || slot_coor[3 * s] < 0 || slot_start[s + 1] <= slot_start[s]) return;
    int64_t b; int bit;
    brick_of(g, slot_coor[3 * s], slot_coor[3 * s + 1], slot_coor[3 * s + 2], b, bit);
    atomicOr((unsigned long long*)(knn_brick + b), 1ull << bit);
}

__global__ void brick_count_kernel(int64_t nb, const uint4* __restrict__ knn_brick, int32_t* cnt)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint4 e = knn_brick[b];
    cnt[b] = __popc(e.x) + __popc(e.y);
}

__global__ void brick_base_kernel(int64_t nb, const int32_t* __restrict__ base, uint4* knn_brick)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    knn_brick[b].z = (uint32_t)base[b];
}

__global__ void brick_list_kernel(GridParams g, const int32_t* __restrict__ counters, const int32_t* __restrict__ slot_coor,
                                  const int32_t* __restrict__ slot_start, const uint4* __restrict__ knn_brick, int2* knn_list)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    int n_rec = counters[0] < g.max_o ? counters[0] : g.max_o;
    if (s >= n_rec || slot_coor[3 * s] < 0 || slot_start[s + 1] <= slot_start[s]) return;
    int64_t b; int bit;
    brick_of(g, slot_coor[3 * s], slot_coor[3 * s + 1], slot_coor[3 * s + 2], b, bit);
    const uint4 e = knn_brick[b];
    const unsigned long long m = ((unsigned long long)e.y << 32) | e.x;
    const int rank = (int)e.z + __popcll(m & ((1ull << bit) - 1ull));
    knn_list[rank] = make_int2(slot_start[s], slot_start[s + 1] - slot_start[s]);
}

// K12-K14: neighbour lists per sample voxel (grid.cuh).
__global__ void word_popc_kernel(int64_t nwords, const uint32_t* __restrict__ bits, int32_t* cnt)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nwords) cnt[w] = __popc(bits[w]);
}

// One thread per 32-voxel WORD of the occupancy mask (the volume is almost empty: a thread per voxel spent 4 ms on 390 M voxels at C4).
template <bool FILL>
__global__ void nbr_list_kernel(GridParams g, int nby, int nbz, int64_t nwords, const uint32_t* __restrict__ occ_bits, const int32_t* __restrict__ occ_rank,
                                const uint4* __restrict__ knn_brick, const int2* __restrict__ knn_list, int32_t* nbr_cnt_or_off, uint32_t* nbr_ent)
{
    const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwords) return;
    uint32_t w = occ_bits[wi];
    if (!w) return;
    int rank = occ_rank[wi];
    for (; w; w &= w - 1, rank++) {
        const int64_t c = wi * 32 + (__ffs(w) - 1);
        const int fz = (int)(c % g.dz), fy = (int)((c / g.dz) % g.dy), fx = (int)(c / ((int64_t)g.dz * g.dy));
        int n = 0;
        uint32_t shell1 = 0x80000000u;
        uint32_t* out = FILL ? nbr_ent + nbr_cnt_or_off[rank] : nullptr;
        for (int i = 0; i < 27; i++) {
            const int l = i == 0 ? 13 : (i - 1 < 13 ? i - 1 : i);       // centre first, then the 26 others with x outer / z inner (:617-627)
            const int vx = fx + l / 9 - 1, vy = fy + (l / 3) % 3 - 1, vz = fz + l % 3 - 1;
            if ((unsigned)vx >= (unsigned)g.dx || (unsigned)vy >= (unsigned)g.dy || (unsigned)vz >= (unsigned)g.dz) continue;
            const uint4 e = knn_brick[((int64_t)(vx >> 2) * nby + (vy >> 2)) * nbz + (vz >> 2)];
            const int bit = ((vx & 3) * 4 + (vy & 3)) * 4 + (vz & 3);
            const unsigned long long m = ((unsigned long long)e.y << 32) | e.x;
            if (!((m >> bit) & 1ull)) continue;
            if (FILL) {
                const int2 li = knn_list[(int)e.z + __popcll(m & ((1ull << bit) - 1ull))];
                uint32_t flag = 0u;
                if (i > 0) { flag = shell1; shell1 = 0u; }
                out[n] = (uint32_t)li.x | ((uint32_t)li.y << 24) | flag;
            }
            n++;
        }
        if (!FILL) nbr_cnt_or_off[rank] = n;
    }
}

static GridParams make_params(const SgnGridCfg* c)
{
    GridParams g;
    g.ox = c->origin[0]; g.oy = c->origin[1]; g.oz = c->origin[2];
    g.vx = c->vsize[0]; g.vy = c->vsize[1]; g.vz = c->vsize[2];
    g.dx = c->dim[0]; g.dy = c->dim[1]; g.dz = c->dim[2];
    g.qx = c->query_size[0]; g.qy = c->query_size[1]; g.qz = c->query_size[2];
    g.P = c->P; g.max_o = c->max_o;
    return g;
}

struct GridLayout {
    size_t cell_slot, occ_bits, slot_coor, slot_count, slot_start, cand, counters, coarse_bits, knn_brick, knn_list, occ_rank, nbr_off, nbr_ent, total;
};

// upper bounds known before the build: every surviving record dilates to at most q^3 sample voxels and is listed by at most 27 of them
static int64_t listed_max(int64_t N, const SgnGridCfg* c) { return N < c->max_o ? N : c->max_o; }
static int64_t sample_voxels_max(int64_t N, const SgnGridCfg* c)
{
    const int64_t q = (int64_t)(c->query_size[0] > 0 ? c->query_size[0] : 1) * (c->query_size[1] > 0 ? c->query_size[1] : 1) * (c->query_size[2] > 0 ? c->query_size[2] : 1);
    const int64_t v = q * listed_max(N, c), vol = (int64_t)c->dim[0] * c->dim[1] * c->dim[2];
    return v < vol ? v : vol;
}
static bool nbr_lists_ok(int64_t N, const SgnGridCfg* c) { return N < (1ll << 24) && c->P < 128; }

static int64_t brick_count(const SgnGridCfg* c) { return (int64_t)brick4(c->dim[0]) * brick4(c->dim[1]) * brick4(c->dim[2]); }

static int64_t grid_vol(const SgnGridCfg* c) { return (int64_t)c->dim[0] * c->dim[1] * c->dim[2]; }

static size_t coarse_words(const SgnGridCfg* c)
{
    const size_t cvol = (size_t)((c->dim[0] + 7) >> 3) * ((c->dim[1] + 7) >> 3) * ((c->dim[2] + 7) >> 3);
    return (cvol + 31) / 32;
}

static GridLayout persistent_layout(int64_t N, const SgnGridCfg* c)
{
    GridLayout L;
    size_t off = 0;
    int64_t vol = grid_vol(c);
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    L.cell_slot = take(sizeof(int32_t) * (size_t)vol);
    L.occ_bits = take(sizeof(uint32_t) * (size_t)((vol + 31) / 32));
    L.slot_coor = take(sizeof(int32_t) * 3 * (size_t)c->max_o);
    L.slot_count = take(sizeof(int32_t) * (size_t)c->max_o);
    L.slot_start = take(sizeof(int32_t) * ((size_t)c->max_o + 1));
    L.cand = take(sizeof(float4) * (size_t)(N > 0 ? N : 1));
    L.counters = take(sizeof(int32_t) * 4);
    L.coarse_bits = take(sizeof(uint32_t) * coarse_words(c));
    L.knn_brick = take(sizeof(uint4) * (size_t)brick_count(c));
    L.knn_list = take(sizeof(int2) * ((size_t)c->max_o + 1));
    L.occ_rank = take(sizeof(int32_t) * (size_t)((vol + 31) / 32 + 1));
    L.nbr_off = take(nbr_lists_ok(N, c) ? sizeof(int32_t) * (size_t)(sample_voxels_max(N, c) + 1) : 0);
    L.nbr_ent = take(nbr_lists_ok(N, c) ? sizeof(uint32_t) * (size_t)(27 * listed_max(N, c) + 1) : 0);
    L.total = off;
    return L;
}

static size_t scratch_bytes_for(int64_t N, const SgnGridCfg* c)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { off += align_up(bytes); };
    int64_t nscan = N > c->max_o ? N : c->max_o;
    if (brick_count(c) > nscan) nscan = brick_count(c);
    const int64_t nwords = ((int64_t)c->dim[0] * c->dim[1] * c->dim[2] + 31) / 32;
    if (nwords > nscan) nscan = nwords;
    if (sample_voxels_max(N, c) > nscan) nscan = sample_voxels_max(N, c);
    take(sizeof(int32_t) * (size_t)(nwords + 1));          // set bits per occupancy word
    take(sizeof(int32_t) * (size_t)(sample_voxels_max(N, c) + 1));   // neighbours per sample voxel
    take(sizeof(int32_t) * (size_t)(brick_count(c) + 1)); // listed voxels per brick
    take(sizeof(int32_t) * (size_t)(brick_count(c) + 1)); // its exclusive scan
    take(sizeof(int32_t) * (size_t)(N + 1));              // flag
    take(sizeof(int32_t) * (size_t)(N + 1));              // rank
    take(sizeof(int32_t) * scan_partials_count(nscan));    // scan partials
    take(sizeof(int32_t) * (size_t)c->max_o);             // record_writer
    take(sizeof(int32_t) * ((size_t)c->max_o + 1));       // seg_start
    take(sizeof(int32_t) * (size_t)c->max_o);             // cursor
    take(sizeof(int32_t) * (size_t)(N + 1));              // seg
    take(sizeof(int32_t) * ((size_t)c->max_o + 1));       // ncap
    return off;
}

static int check_cfg(int64_t N, const SgnGridCfg* c)
{
    SGN_CHECK_ARG(c != nullptr, "grid: cfg is NULL");
    SGN_CHECK_ARG(N >= 0 && N < (1ll << 31), "grid: N=%lld out of range", (long long)N);
    SGN_CHECK_ARG(c->dim[0] > 0 && c->dim[1] > 0 && c->dim[2] > 0, "grid: non-positive dim");
    // the reference indexes voxels with a 32-bit int (:293); keep the same domain
    SGN_CHECK_ARG(grid_vol(c) < (1ll << 31), "grid: %d x %d x %d voxels exceed the int32 index domain", c->dim[0], c->dim[1], c->dim[2]);
    SGN_CHECK_ARG(c->P > 0 && c->max_o > 0, "grid: P and max_o must be positive");
    SGN_CHECK_ARG(c->vsize[0] > 0 && c->vsize[1] > 0 && c->vsize[2] > 0, "grid: voxel size must be positive");
    return SGN_OK;
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_grid_workspace_bytes(int64_t N, const SgnGridCfg* cfg, size_t* persistent_bytes, size_t* scratch_bytes)
{
    int rc = check_cfg(N, cfg);
    if (rc) return rc;
    if (persistent_bytes) *persistent_bytes = persistent_layout(N, cfg).total;
    if (scratch_bytes) *scratch_bytes = scratch_bytes_for(N, cfg);
    return SGN_OK;
}

extern "C" int sgn_grid_build(const float* xyz, int64_t N, int64_t actual_n, const SgnGridCfg* cfg, void* persistent,
                              size_t persistent_bytes, void* scratch, size_t scratch_bytes, SgnGrid** out, void* stream)
{
    return sgn_grid_build_flags(xyz, N, actual_n, cfg, persistent, persistent_bytes, scratch, scratch_bytes, 0, out, stream);
}

extern "C" int sgn_grid_build_flags(const float* xyz, int64_t N, int64_t actual_n, const SgnGridCfg* cfg, void* persistent,
                                    size_t persistent_bytes, void* scratch, size_t scratch_bytes, int flags, SgnGrid** out, void* stream)
{
    int rc = check_cfg(N, cfg);
    if (rc) return rc;
    SGN_CHECK_ARG(out != nullptr, "sgn_grid_build: out is NULL");
    SGN_CHECK_ARG(actual_n >= 0 && actual_n <= N, "sgn_grid_build: actual_n out of range");
    GridLayout L = persistent_layout(N, cfg);
    if (persistent_bytes < L.total || scratch_bytes < scratch_bytes_for(N, cfg) || ((uintptr_t)persistent & 255) || ((uintptr_t)scratch & 255)) {
        set_error("sgn_grid_build: workspace too small or misaligned (need %zu + %zu bytes, 256-aligned)", L.total, scratch_bytes_for(N, cfg));
        return SGN_E_WORKSPACE;
    }
    auto st = (cudaStream_t)stream;
    const int64_t vol = grid_vol(cfg);
    const int max_o = cfg->max_o;
    char* pb = (char*)persistent;
    SgnGrid* G = new SgnGrid();
    G->cfg = *cfg; G->N = N; G->vol = vol;
    G->cell_slot = (int32_t*)(pb + L.cell_slot);
    G->occ_bits = (uint32_t*)(pb + L.occ_bits);
    G->slot_coor = (int32_t*)(pb + L.slot_coor);
    G->slot_count = (int32_t*)(pb + L.slot_count);
    G->slot_start = (int32_t*)(pb + L.slot_start);
    G->cand = (float4*)(pb + L.cand);
    G->counters = (int32_t*)(pb + L.counters);
    G->coarse_bits = (uint32_t*)(pb + L.coarse_bits);
    G->knn_brick = (uint4*)(pb + L.knn_brick);
    G->knn_list = (int2*)(pb + L.knn_list);
    G->nbx = brick4(cfg->dim[0]); G->nby = brick4(cfg->dim[1]); G->nbz = brick4(cfg->dim[2]);
    const int64_t nbrick = brick_count(cfg);

    Arena A(scratch, scratch_bytes);
    G->occ_rank = (int32_t*)(pb + L.occ_rank);
    G->nbr_ok = (nbr_lists_ok(N, cfg) && !(flags & SGN_GRID_NO_NEIGHBOUR_LISTS)) ? 1 : 0;
    G->nbr_off = G->nbr_ok ? (int32_t*)(pb + L.nbr_off) : nullptr;
    G->nbr_ent = G->nbr_ok ? (uint32_t*)(pb + L.nbr_ent) : nullptr;
    int64_t nscan = N > max_o ? N : max_o;
    if (nbrick > nscan) nscan = nbrick;
    const int64_t nwords = (vol + 31) / 32, nsv = sample_voxels_max(N, cfg);
    if (nwords > nscan) nscan = nwords;
    if (nsv > nscan) nscan = nsv;
    int32_t* word_cnt = A.take<int32_t>(nwords + 1);
    int32_t* nbr_cnt = A.take<int32_t>(nsv + 1);
    int32_t* brick_cnt = A.take<int32_t>(nbrick + 1);
    int32_t* brick_base = A.take<int32_t>(nbrick + 1);
    int32_t* flag = A.take<int32_t>(N + 1);
    int32_t* rank = A.take<int32_t>(N + 1);
    int32_t* partials = A.take<int32_t>(scan_partials_count(nscan));
    int32_t* record_writer = A.take<int32_t>(max_o);
    int32_t* seg_start = A.take<int32_t>((size_t)max_o + 1);
    int32_t* cursor = A.take<int32_t>(max_o);
    int32_t* seg = A.take<int32_t>(N + 1);
    int32_t* ncap = A.take<int32_t>((size_t)max_o + 1);

    GridParams g = make_params(cfg);
    const int T = 256;
    auto fail = [&](int code) { delete G; return code; };
#define GRID_TRY(expr) do { int rc__ = (expr); if (rc__) return fail(rc__); } while (0)
#define GRID_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_error("%s -> %s", #expr, cudaGetErrorString(e__)); return fail(SGN_E_CUDA); } } while (0)

    GRID_CUDA(cudaMemsetAsync(G->cell_slot, 0xFF, sizeof(int32_t) * (size_t)vol, st));
    GRID_CUDA(cudaMemsetAsync(G->occ_bits, 0, sizeof(uint32_t) * (size_t)((vol + 31) / 32), st));
    GRID_CUDA(cudaMemsetAsync(G->slot_coor, 0xFF, sizeof(int32_t) * 3 * (size_t)max_o, st));
    GRID_CUDA(cudaMemsetAsync(G->slot_count, 0, sizeof(int32_t) * (size_t)max_o, st));
    GRID_CUDA(cudaMemsetAsync(G->counters, 0, sizeof(int32_t) * 4, st));
    GRID_CUDA(cudaMemsetAsync(G->coarse_bits, 0, sizeof(uint32_t) * coarse_words(cfg), st));
    GRID_CUDA(cudaMemsetAsync(record_writer, 0xFF, sizeof(int32_t) * (size_t)max_o, st));
    GRID_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)max_o, st));
    GRID_CUDA(cudaMemsetAsync(G->knn_brick, 0, sizeof(uint4) * (size_t)nbrick, st));

    if (actual_n > 0) {
        const int nb = cdiv(actual_n, T);
        launch(cell_min_kernel, nb, T, 0, st, xyz, actual_n, g, (uint32_t*)G->cell_slot);
        launch(first_flag_kernel, cdiv(N, T), T, 0, st, xyz, actual_n, N, g, (const uint32_t*)G->cell_slot, flag);
        GRID_TRY(exclusive_scan_i32(flag, rank, N, partials, st));
        launch(claim_bid_kernel, cdiv(N, T), T, 0, st, N, flag, rank, max_o, cfg->seconds_claim, record_writer, G->counters);
        launch(claim_resolve_kernel, nb, T, 0, st, xyz, actual_n, g, flag, rank, cfg->seconds_claim, record_writer, G->cell_slot, G->slot_coor);
        launch(slot_count_kernel, nb, T, 0, st, xyz, actual_n, g, G->cell_slot, G->slot_count);
        GRID_TRY(exclusive_scan_i32(G->slot_count, seg_start, max_o, partials, st));
        launch(slot_fill_kernel, nb, T, 0, st, xyz, actual_n, g, G->cell_slot, seg_start, cursor, seg);
        launch(slot_canon_kernel, cdiv(max_o, T), T, 0, st, max_o, cfg->P, cfg->seconds_fill, G->slot_count, seg_start, seg, ncap);
        GRID_TRY(exclusive_scan_i32(ncap, G->slot_start, max_o, partials, st));
        launch(emit_cand_kernel, cdiv(max_o, T), T, 0, st, max_o, xyz, seg_start, seg, G->slot_start, G->cand, G->counters);
        launch(dilate_kernel, cdiv(max_o, T), T, 0, st, g, G->counters, G->slot_coor, G->occ_bits, G->coarse_bits);
        launch(brick_mark_kernel, cdiv(max_o, T), T, 0, st, g, G->counters, G->slot_coor, G->slot_start, G->knn_brick);
        launch(brick_count_kernel, cdiv(nbrick, T), T, 0, st, nbrick, G->knn_brick, brick_cnt);
        GRID_TRY(exclusive_scan_i32(brick_cnt, brick_base, nbrick, partials, st));
        launch(brick_base_kernel, cdiv(nbrick, T), T, 0, st, nbrick, brick_base, G->knn_brick);
        launch(brick_list_kernel, cdiv(max_o, T), T, 0, st, g, G->counters, G->slot_coor, G->slot_start, G->knn_brick, G->knn_list);
        launch(word_popc_kernel, cdiv(nwords, T), T, 0, st, nwords, G->occ_bits, word_cnt);
        GRID_TRY(exclusive_scan_i32(word_cnt, G->occ_rank, nwords, partials, st));
        if (G->nbr_ok) {
            GRID_CUDA(cudaMemsetAsync(nbr_cnt, 0, sizeof(int32_t) * (size_t)(nsv + 1), st));
            launch(nbr_list_kernel<false>, cdiv(nwords, T), T, 0, st, g, G->nby, G->nbz, nwords, G->occ_bits, G->occ_rank, G->knn_brick, G->knn_list, nbr_cnt, (uint32_t*)nullptr);
            GRID_TRY(exclusive_scan_i32(nbr_cnt, G->nbr_off, nsv, partials, st));
            launch(nbr_list_kernel<true>, cdiv(nwords, T), T, 0, st, g, G->nby, G->nbz, nwords, G->occ_bits, G->occ_rank, G->knn_brick, G->knn_list, G->nbr_off, G->nbr_ent);
        }
        GRID_CUDA(cudaGetLastError());
    } else {
        GRID_CUDA(cudaMemsetAsync(G->slot_start, 0, sizeof(int32_t) * ((size_t)max_o + 1), st));
        GRID_CUDA(cudaMemsetAsync(G->occ_rank, 0, sizeof(int32_t) * (size_t)(nwords + 1), st));
        if (G->nbr_ok) GRID_CUDA(cudaMemsetAsync(G->nbr_off, 0, sizeof(int32_t) * (size_t)(nsv + 1), st));
    }
#undef GRID_TRY
#undef GRID_CUDA
    *out = G;
    return SGN_OK;
}

extern "C" int sgn_grid_destroy(SgnGrid* g)
{
    delete g;
    return SGN_OK;
}

extern "C" int sgn_grid_buffer(const SgnGrid* g, int which, void** ptr, int64_t* n)
{
    SGN_CHECK_ARG(g && ptr && n, "sgn_grid_buffer: NULL argument");
    switch (which) {
        case 0: *ptr = g->cell_slot; *n = g->vol; break;
        case 1: *ptr = g->occ_bits; *n = (g->vol + 31) / 32; break;
        case 2: *ptr = g->slot_coor; *n = 3ll * g->cfg.max_o; break;
        case 3: *ptr = g->slot_count; *n = g->cfg.max_o; break;
        case 4: *ptr = g->slot_start; *n = (int64_t)g->cfg.max_o + 1; break;
        case 5: *ptr = g->cand; *n = g->N; break;
        case 6: *ptr = g->counters; *n = 4; break;
        case 7: *ptr = g->coarse_bits; *n = (int64_t)coarse_words(&g->cfg); break;
        case 8: *ptr = g->knn_brick; *n = 4 * brick_count(&g->cfg); break;
        case 9: *ptr = g->knn_list; *n = 2 * ((int64_t)g->cfg.max_o + 1); break;
        default: set_error("sgn_grid_buffer: unknown buffer %d", which); return SGN_E_INVALID;
    }
    return SGN_OK;
}
