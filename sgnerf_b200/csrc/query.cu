// query.cu -- ray march over the occupancy bitmask + K-nearest-within-radius search, all rays in one pass.
//
// Reference: ray positions (models/rendering/diff_ray_marching.py:387), mask_raypos, the cumsum glue,
// get_shadingloc[_with_semantic], query_neigh_along_ray_layered[_semantic_guidance]
// (models/neural_points/query_point_indices_worldcoords.py:413-681, :811-938).
//
// march_kernel: one warp per ray.  Lanes test 32 consecutive depth candidates against the 1-bit/voxel
// occupancy mask; ballot + prefix popcount gives the running index the reference obtains with
// torch.cumsum (:843-844); the first SR occupied candidates become the shading samples.
// knn_kernel: one thread per shading sample, visiting voxels and candidates in exactly the reference's
// order (shell by shell, x outer / z inner, list order) with the same replace-farthest rule, so the K
// slots come out in the same order as the sequential reference, ties included.  Candidates of a voxel
// are consecutive float4 (x,y,z,index) records, so the inner loop is one 16-byte load per candidate.
#include "common.cuh"
#include "grid.cuh"


namespace sgn {

struct QueryGrid {
    float ox, oy, oz, vx, vy, vz;
    int dx, dy, dz;
    const int32_t* cell_slot;
    const uint32_t* occ_bits;
    const int32_t* slot_start;
    const float4* cand;
    const uint32_t* coarse_bits;   // 8^3-voxel bricks that may hold an occupied voxel (grid.cu:dilate_kernel)
    int cdx, cdy, cdz;
};

constexpr int MARCH_WARPS = 8;

__global__ void __launch_bounds__(MARCH_WARPS * 32)
march_kernel(QueryGrid g, const float* __restrict__ campos, const float* __restrict__ raydir, const float* __restrict__ t,
             int t_per_ray, int64_t R, int D, int SR, const int32_t* __restrict__ ray_label, float* __restrict__ sample_loc_w,
             int32_t* __restrict__ sample_mask, int32_t* __restrict__ sample_label, int8_t* __restrict__ ray_mask)
{
    const int lane = lane_id();
    const int64_t r = (int64_t)blockIdx.x * MARCH_WARPS + (threadIdx.x >> 5);
    if (r >= R) return;
    const float cx = campos[0], cy = campos[1], cz = campos[2];
    const float dx = raydir[3 * r], dy = raydir[3 * r + 1], dz = raydir[3 * r + 2];
    const float* tr = t_per_ray ? t + r * D : t;
    const int label = ray_label ? ray_label[r] : 0;
    const float rvx = 1.0f / g.vx, rvy = 1.0f / g.vy, rvz = 1.0f / g.vz;
    const float fdx = (float)g.dx + 0.01f, fdy = (float)g.dy + 0.01f, fdz = (float)g.dz + 0.01f;
    __shared__ uint16_t s_queue[MARCH_WARPS][64];            // per warp: depth indices that passed the brick test, in ray order
    uint16_t* queue = s_queue[threadIdx.x >> 5];
    int cnt = 0, qn = 0;
    // exact test of up to 32 queued candidates (lane i takes entry i): voxel coordinate as the reference computes it, occupancy bit,
    // rank among the ray's occupied candidates by ballot + popcount (the reference's torch.cumsum, :843-844), sample store
    auto drain = [&](int n) {
        bool occ = false;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (lane < n) {
            const float tv = __ldg(tr + queue[lane]);
            // campos + raydir * t with separate fp32 multiply and add, as torch evaluates it
            px = __fadd_rn(cx, __fmul_rn(dx, tv));
            py = __fadd_rn(cy, __fmul_rn(dy, tv));
            pz = __fadd_rn(cz, __fmul_rn(dz, tv));
            const int vx = vox_coord_fast(px, g.ox, g.vx, rvx), vy = vox_coord_fast(py, g.oy, g.vy, rvy), vz = vox_coord_fast(pz, g.oz, g.vz, rvz);
            if ((unsigned)vx < (unsigned)g.dx && (unsigned)vy < (unsigned)g.dy && (unsigned)vz < (unsigned)g.dz) {
                const uint32_t c = ((uint32_t)vx * (uint32_t)g.dy + (uint32_t)vy) * (uint32_t)g.dz + (uint32_t)vz;   // < 2^31 (grid.cu:check_cfg)
                occ = (__ldg(g.occ_bits + (c >> 5)) >> (c & 31)) & 1u;
            }
        }
        const unsigned b = __ballot_sync(0xffffffffu, occ);
        const int rank = cnt + __popc(b & ((1u << lane) - 1u));
        if (occ && rank < SR) {
            const int64_t o = r * SR + rank;
            sample_loc_w[3 * o] = px; sample_loc_w[3 * o + 1] = py; sample_loc_w[3 * o + 2] = pz;
            sample_mask[o] = 1;
            if (sample_label) sample_label[o] = label;
        }
        cnt += __popc(b);
    };
    for (int base = 0; base < D && cnt < SR; base += 32) {
        const int d = base + lane;
        bool maybe = false;
        if (d < D) {
            // brick test: approximate voxel coordinate (a few 1e-5 off at most), 8^3-voxel brick, one bit.  The mask covers every brick
            // with an occupied voxel inside or one voxel away, so a candidate that fails it cannot be occupied whatever the rounding;
            // only the others are queued for the exact test, and the queue keeps them in ray order.
            const float tv = __ldg(tr + d);
            const float ax = (__fadd_rn(cx, __fmul_rn(dx, tv)) - g.ox) * rvx, ay = (__fadd_rn(cy, __fmul_rn(dy, tv)) - g.oy) * rvy,
                        az = (__fadd_rn(cz, __fmul_rn(dz, tv)) - g.oz) * rvz;
            if (ax > -0.01f && ay > -0.01f && az > -0.01f && ax < fdx && ay < fdy && az < fdz) {
                const int bx = min((int)(ax * 0.125f), g.cdx - 1), by = min((int)(ay * 0.125f), g.cdy - 1), bz = min((int)(az * 0.125f), g.cdz - 1);
                const int bc = (bx * g.cdy + by) * g.cdz + bz;
                maybe = (__ldg(g.coarse_bits + (bc >> 5)) >> (bc & 31)) & 1u;
            }
        }
        const unsigned mb = __ballot_sync(0xffffffffu, maybe);
        if (mb == 0u) continue;
        if (maybe) queue[qn + __popc(mb & ((1u << lane) - 1u))] = (uint16_t)d;
        qn += __popc(mb);
        __syncwarp();
        if (qn >= 32) {
            drain(32);
            const uint16_t carry = queue[32 + lane];             // at most 31 entries stay behind
            __syncwarp();
            queue[lane] = carry;
            qn -= 32;
            __syncwarp();
        }
    }
    if (qn > 0 && cnt < SR) drain(qn);
    cnt = cnt < SR ? cnt : SR;
    for (int s = cnt + lane; s < SR; s += 32) {  // unused slots stay at world (0,0,0), mask 0 (:835, :845)
        const int64_t o = r * SR + s;
        sample_loc_w[3 * o] = 0.f; sample_loc_w[3 * o + 1] = 0.f; sample_loc_w[3 * o + 2] = 0.f;
        sample_mask[o] = 0;
        if (sample_label) sample_label[o] = 0;
    }
    if (lane == 0) ray_mask[r] = 0;
}

constexpr int KNN_SLOTS = 1536;   // sample slots per block (64 rays at SR = 24): their occupied samples are compacted in shared memory so that all lanes work
constexpr int KNN_THREADS = 128;
constexpr int KNN_MAX_CELLS = 27;  // cell-list entries per thread kept in shared memory (a 3^3 block; larger kernels take the nested-loop path)

// One candidate against the K slots, exactly the reference's rule (:650-676): fill the first K slots in order, then replace the
// farthest one when the new candidate is strictly nearer and rescan for the new farthest (first index wins ties).
template <int KT>
__device__ __forceinline__ void knn_insert(int pidx, float d2, int K, int& kid, int& far_ind, float& far2, int32_t (&out)[KT], float (&buf)[KT])
{
    if (kid++ < K) {
#pragma unroll
        for (int i = 0; i < KT; i++)
            if (i == kid - 1) { out[i] = pidx; buf[i] = d2; }
        if (d2 > far2) { far2 = d2; far_ind = kid - 1; }
    } else if (d2 < far2) {
#pragma unroll
        for (int i = 0; i < KT; i++)
            if (i == far_ind) { out[i] = pidx; buf[i] = d2; }
        far2 = d2;
#pragma unroll
        for (int i = 0; i < KT; i++)
            if (i < K && buf[i] > far2) { far2 = buf[i]; far_ind = i; }
    }
}

// knn_kernel, two phases per sample so that the lanes of a warp (32 different samples) stay in step:
//   1. walk the voxels of the block in the reference's order (shell by shell, x outer / z inner) and append the occupied ones --
//      (first candidate, end, "opens a new shell") -- to a per-thread list in shared memory: a uniform 27-iteration loop;
//   2. one flat loop over the candidates of that list, one candidate per lane per iteration: a warp runs max-over-lanes of the
//      candidate counts instead of the sum over voxels of the per-voxel maxima the nested loops cost.
// The visiting order and the insertion rule are unchanged, so the K slots come out as in the sequential reference.
template <int KT, bool SEMANTIC>
__global__ void __launch_bounds__(KNN_THREADS, KT == 8 ? 8 : 3)
knn_kernel(QueryGrid g, int64_t R, int SR, int K, int nlayer, float radius2, const float* __restrict__ sample_loc_w,
           const int32_t* __restrict__ sample_mask, const int32_t* __restrict__ sample_label, const int32_t* __restrict__ pt_label,
           const int32_t* __restrict__ pt_label_prob_bits, uint64_t seconds, int32_t* __restrict__ sample_pidx,
           int8_t* __restrict__ ray_mask, int rays_per_block, int flat_ok)
{
    extern __shared__ int32_t s_list[];                       // [rays_per_block * SR] sample indices (relative to the block's first slot)
    __shared__ uint32_t s_cells[KNN_MAX_CELLS][KNN_THREADS];  // per-thread list of occupied voxels: begin (24 bits) | count (7 bits) << 24 | new-shell flag (bit 31)
    __shared__ int s_count;
    const int64_t slot0 = (int64_t)blockIdx.x * rays_per_block * SR;
    const int nslot = (int)min((int64_t)rays_per_block * SR, R * SR - slot0);
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    // pass 1: unoccupied slots get their empty neighbour lists right away, occupied ones are queued (any order: results go by index)
    for (int i = threadIdx.x; i < ((nslot + 31) & ~31); i += blockDim.x) {
        const bool occ = i < nslot && __ldg(sample_mask + slot0 + i) > 0;
        const unsigned b = __ballot_sync(0xffffffffu, occ);
        int base = 0;
        if (lane_id() == 0 && b) base = atomicAdd(&s_count, __popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (occ) s_list[base + __popc(b & ((1u << lane_id()) - 1u))] = i;
        else if (i < nslot) {
            int32_t* o = sample_pidx + (slot0 + i) * K;
            for (int k = 0; k < K; k++) o[k] = -1;
        }
    }
    __syncthreads();
    const int nq = s_count;
    const bool flat = nlayer <= 2 && flat_ok;                 // 3^3 block, candidate indices < 2^24, at most 127 candidates per voxel: the list fits
    for (int qi = threadIdx.x; qi < nq; qi += blockDim.x) {
        const int64_t idx = slot0 + s_list[qi];
        const int64_t r = idx / SR;
        int32_t out[KT];
        float buf[KT];
#pragma unroll
        for (int i = 0; i < KT; i++) { out[i] = -1; buf[i] = 0.f; }
        int kid = 0;
        const float cx = sample_loc_w[3 * idx], cy = sample_loc_w[3 * idx + 1], cz = sample_loc_w[3 * idx + 2];
        const int fx = vox_coord(cx, g.ox, g.vx), fy = vox_coord(cy, g.oy, g.vy), fz = vox_coord(cz, g.oz, g.vz);
        const int center_label = SEMANTIC ? sample_label[idx] : 0;
        int far_ind = 0;
        float far2 = 0.0f;
        auto candidate = [&](const float4 c) {
            const int pidx = __float_as_int(c.w);
            if (SEMANTIC) {
                const int label_v = pt_label[pidx];
                const int label_prob = (int)(__int_as_float(pt_label_prob_bits[(int64_t)pidx * 20 + label_v]) * 10.0f);
                const bool ok = (center_label == label_v) || (label_v == 0) || (center_label == 0) ||
                                ((center_label != label_v) && ((seconds % 10) <= (uint64_t)(int64_t)(1 - label_prob)));
                if (!ok) return;
            }
            const float xv = __fsub_rn(c.x, cx), yv = __fsub_rn(c.y, cy), zv = __fsub_rn(c.z, cz);
            // nvcc contracts the reference's x*x + y*y + z*z to fma(z,z, fma(x,x, y*y))
            const float d2 = __fmaf_rn(zv, zv, __fmaf_rn(xv, xv, __fmul_rn(yv, yv)));
            if (radius2 == 0.0f || d2 <= radius2) knn_insert<KT>(pidx, d2, K, kid, far_ind, far2, out, buf);
        };
        if (flat) {
            // ---- phase 1: the voxels of the 3^3 block in visiting order -- centre (shell 0), then x outer / y / z inner without the
            // centre (shell 1; nlayer == 1 stops after the centre).  All voxel reads are issued before any is used, then the
            // candidate ranges of the occupied ones nine at a time: two or three memory latencies instead of one per voxel.
            int occ[27];
#pragma unroll
            for (int i = 0; i < 27; i++) {
                const int l = i == 0 ? 13 : (i - 1 < 13 ? i - 1 : i);
                const int x = l / 9 - 1, y = (l / 3) % 3 - 1, z = l % 3 - 1;
                const bool in = (i == 0 || nlayer > 1) && (unsigned)(fx + x) < (unsigned)g.dx && (unsigned)(fy + y) < (unsigned)g.dy &&
                                (unsigned)(fz + z) < (unsigned)g.dz;
                occ[i] = in ? __ldg(g.cell_slot + ((int64_t)(fx + x) * g.dy + (fy + y)) * g.dz + (fz + z)) : -1;
            }
            int nc = 0;
            unsigned shell1 = 0x80000000u;
#pragma unroll
            for (int grp = 0; grp < 3; grp++) {
                int b[9], e[9];
#pragma unroll
                for (int j = 0; j < 9; j++) {
                    const int o = occ[9 * grp + j];
                    b[j] = o >= 0 ? __ldg(g.slot_start + o) : 0;
                    e[j] = o >= 0 ? __ldg(g.slot_start + o + 1) : 0;
                }
#pragma unroll
                for (int j = 0; j < 9; j++) {
                    if (e[j] > b[j]) {
                        unsigned flag = 0u;
                        if (9 * grp + j > 0) { flag = shell1; shell1 = 0u; }
                        s_cells[nc++][threadIdx.x] = (uint32_t)b[j] | ((uint32_t)(e[j] - b[j]) << 24) | flag;
                    }
                }
            }
            // ---- phase 2: the candidates of the list, one per iteration; the next candidate's record is requested before the
            // current one is processed (its address never depends on the outcome, only whether it is used does)
            int ci = 0, q = 0, e = 0;
            bool live = nc > 0;
            float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) { const uint32_t c0 = s_cells[0][threadIdx.x]; q = (int)(c0 & 0xffffffu); e = q + (int)((c0 >> 24) & 127u); cur = __ldg(g.cand + q); }
            while (live) {
                bool nxt_live = true, nxt_shell = false;
                if (++q == e) {
                    if (++ci == nc) nxt_live = false;
                    else {
                        const uint32_t cn = s_cells[ci][threadIdx.x];
                        nxt_shell = (cn >> 31) != 0u;
                        q = (int)(cn & 0xffffffu); e = q + (int)((cn >> 24) & 127u);
                    }
                }
                float4 nxt = cur;
                if (nxt_live) nxt = __ldg(g.cand + q);
                candidate(cur);
                if (nxt_shell && kid >= K) nxt_live = false;          // a new shell starts: the reference stops once K were found (:678)
                cur = nxt; live = nxt_live;
            }
        } else {
            for (int layer = 0; layer < nlayer; layer++) {
                const int xlo = max(-fx, -layer), xhi = min(g.dx - fx, layer + 1);
                const int ylo = max(-fy, -layer), yhi = min(g.dy - fy, layer + 1);
                const int zlo = max(-fz, -layer), zhi = min(g.dz - fz, layer + 1);
                for (int x = xlo; x < xhi; x++) {
                    for (int y = ylo; y < yhi; y++) {
                        const int64_t rowbase = ((int64_t)(fx + x) * g.dy + (fy + y)) * g.dz + fz;
                        const bool inner = max(abs(x), abs(y)) != layer;
                        for (int z = zlo; z < zhi; z++) {
                            if (inner && abs(z) != layer) continue;
                            const int occ = __ldg(g.cell_slot + rowbase + z);
                            if (occ < 0) continue;
                            const int b = __ldg(g.slot_start + occ), e = __ldg(g.slot_start + occ + 1);
                            for (int q = b; q < e; q++) candidate(__ldg(g.cand + q));
                        }
                    }
                }
                if (kid >= K) break;
            }
        }
        int32_t* o = sample_pidx + idx * K;
        if (KT == 8 && K == 8) {
            ((int4*)o)[0] = make_int4(out[0], out[1], out[2], out[3]);
            ((int4*)o)[1] = make_int4(out[4], out[5], out[6], out[7]);
        } else {
#pragma unroll
            for (int i = 0; i < KT; i++)
                if (i < K) o[i] = out[i];
        }
        if (kid > 0) ray_mask[r] = 1;  // every writer stores the same value
    }
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_query(const SgnGrid* G, const float* campos, const float* raydir, const float* t, int t_per_ray, int64_t R, int D,
                         int SR, int K, int kernel_size0, float radius2, const int32_t* ray_label, const int32_t* pt_label,
                         const int32_t* pt_label_prob_bits, uint64_t seconds_query, int32_t* sample_pidx, float* sample_loc_w,
                         int32_t* sample_mask, int32_t* sample_label, int8_t* ray_mask, void* stream)
{
    SGN_CHECK_ARG(G != nullptr, "sgn_query: grid is NULL");
    SGN_CHECK_ARG(R >= 0 && D > 0 && D <= 65535 && SR > 0 && SR <= 4096, "sgn_query: bad R/D/SR (D at most 65535, SR at most 4096)");
    SGN_CHECK_ARG(K > 0 && K <= SGN_MAX_K, "sgn_query: K=%d out of range (1..%d)", K, SGN_MAX_K);
    SGN_CHECK_ARG(sample_pidx && sample_loc_w && sample_mask && ray_mask, "sgn_query: NULL output");
    const bool semantic = ray_label != nullptr;
    SGN_CHECK_ARG(!semantic || (pt_label && pt_label_prob_bits && sample_label), "sgn_query: semantic guidance needs pt_label, pt_label_prob_bits and sample_label");
    if (R == 0) return SGN_OK;
    auto st = (cudaStream_t)stream;
    QueryGrid g;
    g.ox = G->cfg.origin[0]; g.oy = G->cfg.origin[1]; g.oz = G->cfg.origin[2];
    g.vx = G->cfg.vsize[0]; g.vy = G->cfg.vsize[1]; g.vz = G->cfg.vsize[2];
    g.dx = G->cfg.dim[0]; g.dy = G->cfg.dim[1]; g.dz = G->cfg.dim[2];
    g.cell_slot = G->cell_slot; g.occ_bits = G->occ_bits; g.slot_start = G->slot_start; g.cand = G->cand;
    g.coarse_bits = G->coarse_bits; g.cdx = (g.dx + 7) >> 3; g.cdy = (g.dy + 7) >> 3; g.cdz = (g.dz + 7) >> 3;

    launch(march_kernel, cdiv(R, MARCH_WARPS), MARCH_WARPS * 32, 0, st, g, campos, raydir, t, t_per_ray, R, D, SR, ray_label, sample_loc_w,
                                                                   sample_mask, semantic ? sample_label : nullptr, ray_mask);
    const int nlayer = (kernel_size0 + 1) / 2;
    // rays per block: up to KNN_SLOTS sample slots (6 KB list), fewer for small ray counts (a training patch) so that the grid still
    // covers the 148 SMs several times over, but never fewer slots than the block has threads
    int rpb = KNN_SLOTS / SR > 0 ? KNN_SLOTS / SR : 1;
    const int rpb_fill = (int)(R / (148 * 8)), rpb_min = cdiv(128, SR);
    if (rpb > rpb_fill) rpb = rpb_fill > rpb_min ? rpb_fill : (rpb_min < rpb ? rpb_min : rpb);
    const int nb = cdiv(R, rpb);
    const int flat_ok = (G->N < (1ll << 24) && G->cfg.P < 128) ? 1 : 0;
    const size_t ksm = (size_t)rpb * SR * sizeof(int32_t);
    if (K == 8) {
        if (semantic)
            launch(knn_kernel<8, true>, nb, KNN_THREADS, ksm, st, g, R, SR, K, nlayer, radius2, sample_loc_w, sample_mask, sample_label, pt_label,
                                                     pt_label_prob_bits, seconds_query, sample_pidx, ray_mask, rpb, flat_ok);
        else
            launch(knn_kernel<8, false>, nb, KNN_THREADS, ksm, st, g, R, SR, K, nlayer, radius2, sample_loc_w, sample_mask, nullptr, nullptr, nullptr,
                                                      seconds_query, sample_pidx, ray_mask, rpb, flat_ok);
    } else {
        if (semantic)
            launch(knn_kernel<SGN_MAX_K, true>, nb, KNN_THREADS, ksm, st, g, R, SR, K, nlayer, radius2, sample_loc_w, sample_mask, sample_label, pt_label,
                                                             pt_label_prob_bits, seconds_query, sample_pidx, ray_mask, rpb, flat_ok);
        else
            launch(knn_kernel<SGN_MAX_K, false>, nb, KNN_THREADS, ksm, st, g, R, SR, K, nlayer, radius2, sample_loc_w, sample_mask, nullptr, nullptr,
                                                              nullptr, seconds_query, sample_pidx, ray_mask, rpb, flat_ok);
    }
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

__global__ void gather_rows_kernel(const float* __restrict__ table, int C, const int32_t* __restrict__ pidx, int64_t n, float* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    int64_t row = i / C;
    int c = (int)(i - row * C);
    int p = pidx[row];
    p = p < 0 ? 0 : p;
    out[i] = __ldg(table + (int64_t)p * C + c);
}

extern "C" int sgn_gather_rows(const float* table, int C, const int32_t* pidx, int64_t n_rows, float* out, void* stream)
{
    SGN_CHECK_ARG(C > 0 && n_rows >= 0, "sgn_gather_rows: bad sizes");
    if (n_rows == 0) return SGN_OK;
    launch(gather_rows_kernel, cdiv(n_rows * C, 256), 256, 0, (cudaStream_t)stream, table, C, pidx, n_rows, out);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}
