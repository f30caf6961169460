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
    const uint4* knn_brick;        // 4^3-voxel bricks: mask of voxels with a candidate list + rank of the first one (grid.cu:brick_*_kernel)
    const int2* knn_list;          // rank -> (first candidate, count)
    int nby, nbz;
    const int32_t* occ_rank;       // neighbour lists per sample voxel (grid.cuh); nbr_off == nullptr when they were not built
    const int32_t* nbr_off;
    const uint32_t* nbr_ent;
};

constexpr int MARCH_THREADS = 128;
constexpr int64_t MARCH_WARP_BELOW = 131072;   // fewer rays than this: one warp per ray (march_warp_kernel)
constexpr int MARCH_RANGES = 14;   // candidate ranges kept per ray (more are merged into the last one: tests a few extra candidates, never fewer)

// march_kernel: empty space is skipped with a 3-D DDA over the 8^3-voxel brick mask, one THREAD per ray: a step per brick (0.128 m at
// the canonical voxel size) instead of a test per depth candidate (0.02 m).  The walk only produces index ranges [k_lo, k_hi] of the
// candidates whose depth falls in a run of set bricks; the exact test of those -- voxel coordinate exactly as the reference computes
// it, occupancy bit, running count (the reference's [R,D] mask + torch.cumsum, :843-844) -- is then done by the whole WARP, ray by ray
// for its 32 rays, 32 candidates per step in depth order with ballot + popcount ranks, so the first SR occupied candidates land in the
// same slots.  Why no candidate is lost: the bricks' intervals partition the ray's parameter range and candidate k is assigned to the
// brick whose interval holds t[k]; that brick contains a point within ~1e-5 m (fp32 rounding of the DDA planes and of campos +
// raydir * t) of the candidate's position, and grid.cu:dilate_kernel sets every brick holding a voxel equal or ADJACENT to an
// occupied one, so the brick of an occupied candidate is always set.  Rays of a warp are neighbouring pixels and step through
// nearly the same bricks.
__global__ void __launch_bounds__(MARCH_THREADS)
march_kernel(QueryGrid g, const float* __restrict__ campos, const float* __restrict__ raydir, const float* __restrict__ t,
             int t_per_ray, int64_t R, int D, int SR, const int32_t* __restrict__ ray_label, float* __restrict__ sample_loc_w,
             int32_t* __restrict__ sample_mask, int32_t* __restrict__ sample_label, int8_t* __restrict__ ray_mask)
{
    __shared__ uint32_t s_rng[MARCH_RANGES][MARCH_THREADS];      // k_lo | k_hi << 16
    __shared__ uint8_t s_nr[MARCH_THREADS];
    const int lane = lane_id();
    const int64_t r = (int64_t)blockIdx.x * MARCH_THREADS + threadIdx.x;
    const float cx = campos[0], cy = campos[1], cz = campos[2];
    const float rvx = 1.0f / g.vx, rvy = 1.0f / g.vy, rvz = 1.0f / g.vz;
    // ---- phase A: brick walk of this thread's ray -> candidate ranges
    int nr = 0;
    if (r < R) {
        const float dx = raydir[3 * r], dy = raydir[3 * r + 1], dz = raydir[3 * r + 2];
        const float* tr = t_per_ray ? t + r * D : t;
        // brick coordinates along the ray: q(t) = A + B t
        const float Ax = (cx - g.ox) * rvx * 0.125f, Ay = (cy - g.oy) * rvy * 0.125f, Az = (cz - g.oz) * rvz * 0.125f;
        const float Bx = dx * rvx * 0.125f, By = dy * rvy * 0.125f, Bz = dz * rvz * 0.125f;
        const float t_first = __ldg(tr), t_last = __ldg(tr + D - 1);
        const float inv_dt = D > 1 ? (float)(D - 1) / fmaxf(t_last - t_first, 1e-20f) : 0.f;
        // clip to the grid box grown by 0.02 brick (2.5 mm): a candidate in a boundary voxel is never clipped away by rounding
        const float PAD = 0.02f;
        float t0 = t_first - 1e-3f, t1 = t_last + 1e-3f;
        auto slab = [&](float A, float B, float hi) {
            if (fabsf(B) < 1e-12f) { if (A < -PAD || A > hi + PAD) t1 = -1e30f; return; }
            const float ib = 1.0f / B;
            float ta = (-PAD - A) * ib, tb = (hi + PAD - A) * ib;
            if (ta > tb) { const float x = ta; ta = tb; tb = x; }
            t0 = fmaxf(t0, ta); t1 = fminf(t1, tb);
        };
        slab(Ax, Bx, (float)g.dx * 0.125f); slab(Ay, By, (float)g.dy * 0.125f); slab(Az, Bz, (float)g.dz * 0.125f);
        if (t1 >= t0) {
            auto clampi = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
            int bx = clampi((int)floorf(Ax + Bx * t0), g.cdx - 1), by = clampi((int)floorf(Ay + By * t0), g.cdy - 1),
                bz = clampi((int)floorf(Az + Bz * t0), g.cdz - 1);
            const int sx = Bx > 0.f ? 1 : -1, sy = By > 0.f ? 1 : -1, sz = Bz > 0.f ? 1 : -1;
            const float ibx = fabsf(Bx) < 1e-12f ? 0.f : 1.0f / Bx, iby = fabsf(By) < 1e-12f ? 0.f : 1.0f / By, ibz = fabsf(Bz) < 1e-12f ? 0.f : 1.0f / Bz;
            // parameter at which the ray leaves brick b along one axis (planes are recomputed, not accumulated: no drift)
            auto cross = [](int b, int sgn, float A, float ib) { return ib == 0.f ? 3e38f : ((float)(b + (sgn > 0 ? 1 : 0)) - A) * ib; };
            float tx = cross(bx, sx, Ax, ibx), ty = cross(by, sy, Ay, iby), tz = cross(bz, sz, Az, ibz);
            int k = 0;
            // first index with t[k] > tt, never moving backwards: guess from the mean spacing, then exact
            auto seek = [&](float tt) {
                int kg = (int)((tt - t_first) * inv_dt) - (t_per_ray ? 12 : 2);
                if (kg > D) kg = D;
                const int k0 = k;
                if (kg > k) k = kg;
                while (k > k0 && __ldg(tr + k - 1) > tt) k--;
                while (k < D && __ldg(tr + k) <= tt) k++;
            };
            float tcur = t0;
            // candidates at or before the box entry can only belong to it through rounding: keep them with the first brick
            { int kg = (int)((t0 - t_first) * inv_dt) - (t_per_ray ? 12 : 2); if (kg > D) kg = D; if (kg > 0) k = kg;
              while (k > 0 && __ldg(tr + k - 1) >= t0) k--;
              while (k < D && __ldg(tr + k) < t0) k++; }
            bool open = false;
            int lo = 0;
            auto push = [&](int a, int b2) {
                if (b2 < a) return;
                if (nr == MARCH_RANGES) { s_rng[nr - 1][threadIdx.x] = (s_rng[nr - 1][threadIdx.x] & 0xffffu) | ((uint32_t)b2 << 16); return; }
                s_rng[nr++][threadIdx.x] = (uint32_t)a | ((uint32_t)b2 << 16);
            };
            while (k < D && tcur <= t1) {
                const float tnext = fminf(fminf(tx, ty), tz);
                const int bc = (bx * g.cdy + by) * g.cdz + bz;
                const bool set = (__ldg(g.coarse_bits + (bc >> 5)) >> (bc & 31)) & 1u;
                if (set && !open) { if (tcur > t0) seek(tcur); lo = k; open = true; }
                else if (!set && open) { seek(tcur); push(lo, k - 1); open = false; }
                // step into the next brick
                bool out_of_box;
                if (tx <= ty && tx <= tz) { bx += sx; out_of_box = (unsigned)bx >= (unsigned)g.cdx; tx = cross(bx, sx, Ax, ibx); }
                else if (ty <= tz)        { by += sy; out_of_box = (unsigned)by >= (unsigned)g.cdy; ty = cross(by, sy, Ay, iby); }
                else                      { bz += sz; out_of_box = (unsigned)bz >= (unsigned)g.cdz; tz = cross(bz, sz, Az, ibz); }
                tcur = tnext;
                if (out_of_box) break;
            }
            if (open) { seek(fminf(tcur, t1)); push(lo, k - 1); }
        }
        ray_mask[r] = 0;
    }
    s_nr[threadIdx.x] = (uint8_t)nr;
    __syncwarp();
    // ---- phase B: the warp tests the candidates of its 32 rays, ray by ray, 32 candidates at a time in depth order
    const int64_t r0 = r - lane;
    const int w0 = threadIdx.x - lane;
    for (int j = 0; j < 32; j++) {
        const int64_t rj = r0 + j;
        if (rj >= R) break;
        const int n_rng = s_nr[w0 + j];
        int cnt = 0;
        if (n_rng > 0) {
            const float dx = raydir[3 * rj], dy = raydir[3 * rj + 1], dz = raydir[3 * rj + 2];
            const float* tr = t_per_ray ? t + rj * D : t;
            const int label = ray_label ? ray_label[rj] : 0;
            // lane i holds range i; the candidates of all ranges form one sequence, 32 consecutive elements per step
            const uint32_t mine = lane < n_rng ? s_rng[lane][w0 + j] : 0u;
            const int my_lo = (int)(mine & 0xffffu), my_len = lane < n_rng ? (int)(mine >> 16) - my_lo + 1 : 0;
            int inc = my_len;                                  // inclusive prefix of the range lengths
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
            const int total = __shfl_sync(0xffffffffu, inc, MARCH_RANGES - 1 < 31 ? 15 : 31);
            for (int base = 0; base < total && cnt < SR; base += 32) {
                const int p = base + lane;
                int k = -1;
                for (int i = 0; i < n_rng; i++) {              // range that holds element p (n_rng is the same for the whole warp)
                    const int end_i = __shfl_sync(0xffffffffu, inc, i);
                    const uint32_t rg = __shfl_sync(0xffffffffu, mine, i);
                    if (k < 0 && p < end_i) k = (int)(rg >> 16) - (end_i - 1 - p);
                }
                bool occ = false;
                float px = 0.f, py = 0.f, pz = 0.f;
                if (k >= 0 && p < total) {
                    const float tv = __ldg(tr + k);
                    // campos + raydir * t with separate fp32 multiply and add, as torch evaluates it
                    px = __fadd_rn(cx, __fmul_rn(dx, tv)); py = __fadd_rn(cy, __fmul_rn(dy, tv)); pz = __fadd_rn(cz, __fmul_rn(dz, tv));
                    const int vx = vox_coord_fast(px, g.ox, g.vx, rvx), vy = vox_coord_fast(py, g.oy, g.vy, rvy), vz = vox_coord_fast(pz, g.oz, g.vz, rvz);
                    if ((unsigned)vx < (unsigned)g.dx && (unsigned)vy < (unsigned)g.dy && (unsigned)vz < (unsigned)g.dz) {
                        const uint32_t c = ((uint32_t)vx * (uint32_t)g.dy + (uint32_t)vy) * (uint32_t)g.dz + (uint32_t)vz;   // < 2^31 (grid.cu:check_cfg)
                        occ = (__ldg(g.occ_bits + (c >> 5)) >> (c & 31)) & 1u;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, occ);
                const int rank = cnt + __popc(bal & ((1u << lane) - 1u));
                if (occ && rank < SR) {
                    const int64_t o = rj * SR + rank;
                    sample_loc_w[3 * o] = px; sample_loc_w[3 * o + 1] = py; sample_loc_w[3 * o + 2] = pz;
                    sample_mask[o] = 1;
                    if (sample_label) sample_label[o] = label;
                }
                cnt += __popc(bal);
            }
        }
        // unused slots stay at world (0,0,0), mask 0 (:835, :845)
        for (int sl = (cnt < SR ? cnt : SR) + lane; sl < SR; sl += 32) {
            const int64_t o = rj * SR + sl;
            sample_loc_w[3 * o] = 0.f; sample_loc_w[3 * o + 1] = 0.f; sample_loc_w[3 * o + 2] = 0.f;
            sample_mask[o] = 0;
            if (sample_label) sample_label[o] = 0;
        }
    }
}

constexpr int MARCH_WARPS = 8;

// march_warp_kernel: one WARP per ray, used for small ray counts (a training patch), where one thread per ray would leave the GPU
// empty: lanes test 32 consecutive depth candidates against the brick mask, the survivors are queued in ray order and exact-tested
// 32 at a time (same arithmetic, same ranks as march_kernel).

__global__ void __launch_bounds__(MARCH_WARPS * 32)
march_warp_kernel(QueryGrid g, const float* __restrict__ campos, const float* __restrict__ raydir, const float* __restrict__ t,
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
constexpr int KNN_BINS = 64;      // candidate-count bins of the in-block counting sort

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
__global__ void __launch_bounds__(KNN_THREADS, KT == 8 ? 10 : 3)
knn_kernel(QueryGrid g, int64_t R, int SR, int K, int nlayer, float radius2, const float* __restrict__ sample_loc_w,
           const int32_t* __restrict__ sample_mask, const int32_t* __restrict__ sample_label, const int32_t* __restrict__ pt_label,
           const int32_t* __restrict__ pt_label_prob_bits, uint64_t seconds, int32_t* __restrict__ sample_pidx,
           int8_t* __restrict__ ray_mask, int rays_per_block, int flat_ok, int write_empty)
{
    extern __shared__ int32_t s_list[];                       // [slots_cap] sample indices (relative to the block's first slot), then s_off / s_sorted / s_nent / s_key
    const int slots_cap = (rays_per_block * SR + 3) & ~3;
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
        else if (i < nslot && write_empty) {            // sgn_query_frame leaves these rows unwritten: its consumer takes the mask
            int32_t* o = sample_pidx + (slot0 + i) * K;
            if (KT == 8 && K == 8) { ((int4*)o)[0] = make_int4(-1, -1, -1, -1); ((int4*)o)[1] = make_int4(-1, -1, -1, -1); }
            else for (int k = 0; k < K; k++) o[k] = -1;
        }
    }
    __syncthreads();
    const int nq = s_count;
    const bool flat = nlayer <= 2 && flat_ok;                 // 3^3 (or 1^3) block and prebuilt neighbour lists (candidate indices < 2^24, P < 128)
    // pass 2 (lists only): the sample's voxel -> its neighbour list (grid.cu:nbr_list_kernel) and the number of candidates it holds; the
    // queue is then counting-sorted by that number, so that the 32 samples a warp takes together run loops of nearly the same length
    // (samples are independent and results go by index, so the processing order is free).
    int32_t* s_off = s_list + slots_cap;                      // first list entry of the queued sample
    uint16_t* s_sorted = (uint16_t*)(s_off + slots_cap);      // queue positions ordered by candidate count
    uint8_t* s_nent = (uint8_t*)(s_sorted + slots_cap);       // entries in the list (<= 27)
    uint8_t* s_key = s_nent + slots_cap;
    __shared__ int s_hist[KNN_BINS + 1];
    if (flat) {
        for (int i = threadIdx.x; i <= KNN_BINS; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int qi = threadIdx.x; qi < nq; qi += blockDim.x) {
            const int64_t idx = slot0 + s_list[qi];
            const int fx = vox_coord(sample_loc_w[3 * idx], g.ox, g.vx), fy = vox_coord(sample_loc_w[3 * idx + 1], g.oy, g.vy),
                      fz = vox_coord(sample_loc_w[3 * idx + 2], g.oz, g.vz);
            const uint32_t c = ((uint32_t)fx * (uint32_t)g.dy + (uint32_t)fy) * (uint32_t)g.dz + (uint32_t)fz;
            const uint32_t w = __ldg(g.occ_bits + (c >> 5));
            const int vr = __ldg(g.occ_rank + (c >> 5)) + __popc(w & ((1u << (c & 31)) - 1u));
            const int ci = __ldg(g.nbr_off + vr), ce = __ldg(g.nbr_off + vr + 1);
            int total = 0;
            for (int j = ci; j < ce; j++) total += (int)((__ldg(g.nbr_ent + j) >> 24) & 127u);
            const int key = total < KNN_BINS - 1 ? total : KNN_BINS - 1;
            s_off[qi] = ci; s_nent[qi] = (uint8_t)(ce - ci); s_key[qi] = (uint8_t)key;
            atomicAdd(&s_hist[key], 1);
        }
        __syncthreads();
        if (threadIdx.x < 32) {                               // exclusive scan of the histogram, largest counts first
            int carry = 0;
            for (int b0 = 0; b0 < KNN_BINS; b0 += 32) {
                const int bin = KNN_BINS - 1 - (b0 + lane_id());
                const int v = s_hist[bin];
                int inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, inc, o); if (lane_id() >= o) inc += n; }
                __syncwarp();
                s_hist[bin] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncthreads();
        for (int qi = threadIdx.x; qi < nq; qi += blockDim.x) s_sorted[atomicAdd(&s_hist[s_key[qi]], 1)] = (uint16_t)qi;
        __syncthreads();
    }
    for (int pos = threadIdx.x; pos < nq; pos += blockDim.x) {
        const int qi = flat ? (int)s_sorted[pos] : pos;
        const int64_t idx = slot0 + s_list[qi];
        const int64_t r = idx / SR;
        int32_t out[KT];
        float buf[KT];
#pragma unroll
        for (int i = 0; i < KT; i++) { out[i] = -1; buf[i] = 0.f; }
        int kid = 0;
        const float cx = sample_loc_w[3 * idx], cy = sample_loc_w[3 * idx + 1], cz = sample_loc_w[3 * idx + 2];
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
            // the prebuilt neighbour list of the sample's voxel: the occupied voxels of the 3^3 block in the reference's visiting order,
            // one 4-byte entry each (first candidate | count << 24 | "shell 1 starts here" << 31)
            int ci = s_off[qi];
            const int ce = ci + (int)s_nent[qi];
            // one flat loop over the candidates of the list; the next candidate's record (and the next list entry) is requested before
            // the current one is processed -- its address never depends on the outcome, only whether it is used does
            int q = 0, e = 0;
            bool live = ci < ce;
            float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t nent = 0u;
            if (live) {
                const uint32_t c0 = __ldg(g.nbr_ent + ci);
                q = (int)(c0 & 0xffffffu); e = q + (int)((c0 >> 24) & 127u); cur = __ldg(g.cand + q);
                if (ci + 1 < ce) nent = __ldg(g.nbr_ent + ci + 1);
            }
            while (live) {
                bool nxt_live = true, nxt_shell = false;
                if (++q == e) {
                    if (++ci == ce) nxt_live = false;
                    else {
                        nxt_shell = (nent >> 31) != 0u;
                        q = (int)(nent & 0xffffffu); e = q + (int)((nent >> 24) & 127u);
                        if (ci + 1 < ce) nent = __ldg(g.nbr_ent + ci + 1);
                    }
                }
                float4 nxt = cur;
                if (nxt_live) nxt = __ldg(g.cand + q);
                candidate(cur);
                // a new shell starts: the reference stops once K were found (:678); with a 1^3 kernel it never looks past the centre
                if (nxt_shell && (kid >= K || nlayer < 2)) nxt_live = false;
                cur = nxt; live = nxt_live;
            }
        } else {
            const int fx = vox_coord(cx, g.ox, g.vx), fy = vox_coord(cy, g.oy, g.vy), fz = vox_coord(cz, g.oz, g.vz);
            for (int layer = 0; layer < nlayer; layer++) {
                const int xlo = max(-fx, -layer), xhi = min(g.dx - fx, layer + 1);
                const int ylo = max(-fy, -layer), yhi = min(g.dy - fy, layer + 1);
                const int zlo = max(-fz, -layer), zhi = min(g.dz - fz, layer + 1);
                for (int x = xlo; x < xhi; x++) {
                    for (int y = ylo; y < yhi; y++) {
                        const bool inner = max(abs(x), abs(y)) != layer;
                        for (int z = zlo; z < zhi; z++) {
                            if (inner && abs(z) != layer) continue;
                            const int vx = fx + x, vy = fy + y, vz = fz + z;
                            const uint4 be = __ldg(g.knn_brick + ((int64_t)(vx >> 2) * g.nby + (vy >> 2)) * g.nbz + (vz >> 2));
                            const int bit = ((vx & 3) * 4 + (vy & 3)) * 4 + (vz & 3);
                            const unsigned long long m = ((unsigned long long)be.y << 32) | be.x;
                            if (!((m >> bit) & 1ull)) continue;
                            const int2 li = __ldg(g.knn_list + (int)be.z + __popcll(m & ((1ull << bit) - 1ull)));
                            for (int q = li.x; q < li.x + li.y; q++) candidate(__ldg(g.cand + q));
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

static int g_march_mode = 0;   // 0: by ray count, 1: always the thread-per-ray brick walk, 2: always the warp-per-ray kernel

extern "C" int sgn_query_march_mode(int mode)
{
    SGN_CHECK_ARG(mode >= 0 && mode <= 2, "sgn_query_march_mode: mode must be 0 (auto), 1 (brick walk) or 2 (warp per ray)");
    g_march_mode = mode;
    return SGN_OK;
}

static int query_impl(const SgnGrid* G, const float* campos, const float* raydir, const float* t, int t_per_ray, int64_t R, int D,
                      int SR, int K, int kernel_size0, float radius2, const int32_t* ray_label, const int32_t* pt_label,
                      const int32_t* pt_label_prob_bits, uint64_t seconds_query, int32_t* sample_pidx, float* sample_loc_w,
                      int32_t* sample_mask, int32_t* sample_label, int8_t* ray_mask, int write_empty, void* stream)
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
    g.knn_brick = G->knn_brick; g.knn_list = G->knn_list; g.nby = G->nby; g.nbz = G->nbz;
    g.occ_rank = G->occ_rank; g.nbr_off = G->nbr_off; g.nbr_ent = G->nbr_ent;

    // thread-per-ray brick walk for frames; warp-per-ray for small ray counts (its latency is ~10x lower when the GPU is not full)
    if (g_march_mode == 1 || (g_march_mode == 0 && R >= MARCH_WARP_BELOW))
        launch(march_kernel, cdiv(R, MARCH_THREADS), MARCH_THREADS, 0, st, g, campos, raydir, t, t_per_ray, R, D, SR, ray_label, sample_loc_w,
                                                                    sample_mask, semantic ? sample_label : nullptr, ray_mask);
    else
        launch(march_warp_kernel, cdiv(R, MARCH_WARPS), MARCH_WARPS * 32, 0, st, g, campos, raydir, t, t_per_ray, R, D, SR, ray_label, sample_loc_w,
                                                                            sample_mask, semantic ? sample_label : nullptr, ray_mask);
    const int nlayer = (kernel_size0 + 1) / 2;
    // rays per block: up to KNN_SLOTS sample slots (6 KB list), fewer for small ray counts (a training patch) so that the grid still
    // covers the 148 SMs several times over, but never fewer slots than the block has threads
    int rpb = KNN_SLOTS / SR > 0 ? KNN_SLOTS / SR : 1;
    const int rpb_fill = (int)(R / (148 * 8)), rpb_min = cdiv(128, SR);
    if (rpb > rpb_fill) rpb = rpb_fill > rpb_min ? rpb_fill : (rpb_min < rpb ? rpb_min : rpb);
    const int nb = cdiv(R, rpb);
    const int flat_ok = G->nbr_ok;
    const size_t ksm = (size_t)((rpb * SR + 3) & ~3) * (4 + 4 + 2 + 1 + 1);
    if (K == 8) {
        if (semantic)
            launch(knn_kernel<8, true>, nb, KNN_THREADS, ksm, st, g, R, SR, K, nlayer, radius2, sample_loc_w, sample_mask, sample_label, pt_label,
                                                     pt_label_prob_bits, seconds_query, sample_pidx, ray_mask, rpb, flat_ok, write_empty);
        else
            launch(knn_kernel<8, false>, nb, KNN_THREADS, ksm, st, g, R, SR, K, nlayer, radius2, sample_loc_w, sample_mask, nullptr, nullptr, nullptr,
                                                      seconds_query, sample_pidx, ray_mask, rpb, flat_ok, write_empty);
    } else {
        if (semantic)
            launch(knn_kernel<SGN_MAX_K, true>, nb, KNN_THREADS, ksm, st, g, R, SR, K, nlayer, radius2, sample_loc_w, sample_mask, sample_label, pt_label,
                                                             pt_label_prob_bits, seconds_query, sample_pidx, ray_mask, rpb, flat_ok, write_empty);
        else
            launch(knn_kernel<SGN_MAX_K, false>, nb, KNN_THREADS, ksm, st, g, R, SR, K, nlayer, radius2, sample_loc_w, sample_mask, nullptr, nullptr,
                                                              nullptr, seconds_query, sample_pidx, ray_mask, rpb, flat_ok, write_empty);
    }
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_query(const SgnGrid* G, const float* campos, const float* raydir, const float* t, int t_per_ray, int64_t R, int D,
                         int SR, int K, int kernel_size0, float radius2, const int32_t* ray_label, const int32_t* pt_label,
                         const int32_t* pt_label_prob_bits, uint64_t seconds_query, int32_t* sample_pidx, float* sample_loc_w,
                         int32_t* sample_mask, int32_t* sample_label, int8_t* ray_mask, void* stream)
{
    return query_impl(G, campos, raydir, t, t_per_ray, R, D, SR, K, kernel_size0, radius2, ray_label, pt_label, pt_label_prob_bits, seconds_query,
                      sample_pidx, sample_loc_w, sample_mask, sample_label, ray_mask, 1, stream);
}

extern "C" int sgn_query_frame(const SgnGrid* G, const float* campos, const float* raydir, const float* t, int t_per_ray, int64_t R, int D,
                               int SR, int K, int kernel_size0, float radius2, const int32_t* ray_label, const int32_t* pt_label,
                               const int32_t* pt_label_prob_bits, uint64_t seconds_query, int32_t* sample_pidx, float* sample_loc_w,
                               int32_t* sample_mask, int32_t* sample_label, int8_t* ray_mask, void* stream)
{
    return query_impl(G, campos, raydir, t, t_per_ray, R, D, SR, K, kernel_size0, radius2, ray_label, pt_label, pt_label_prob_bits, seconds_query,
                      sample_pidx, sample_loc_w, sample_mask, sample_label, ray_mask, 0, stream);
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
