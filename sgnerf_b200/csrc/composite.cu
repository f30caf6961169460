// composite.cu -- alpha compositing forward/backward as warp scans, step-size glue, fill_invalid.
//
// Reference: ray_march / alpha_ray_march (models/rendering/diff_ray_marching.py:509-573),
// radiance_render / alpha_blend / alpha2_blend (models/rendering/diff_render_func.py:36-49),
// ray_dist glue and fill_invalid (models/neural_points_volumetric_model.py:569-577, :158-195).
//
// Layout: one warp per ray, lane = sample (chunks of 32 for SR > 32), so every global access is a
// coalesced run of SR consecutive elements.  HBM-bound: forward moves SR*(16+4+1) B in and
// SR*4*(#outputs) + 16 B out per ray; nothing is re-read.
#include "common.cuh"

namespace sgn {

constexpr int COMP_WARPS = 8;        // warps per block
constexpr int COMP_MAX_CHUNKS = 32;  // SR <= 1024

__device__ __forceinline__ float warp_incl_prod(float v)
{
    const int lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v *= t;
    }
    return v;
}

__device__ __forceinline__ float warp_incl_max(float v)
{
    const int lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = fmaxf(v, t);
    }
    return v;
}

// suffix (reverse inclusive) sum across lanes
__device__ __forceinline__ float warp_rincl_sum(float v)
{
    const int lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += t;
    }
    return v;
}

template <int BLEND>
__global__ void __launch_bounds__(COMP_WARPS * 32)
composite_fwd_kernel(const float4* __restrict__ decoded, const float* __restrict__ ray_dist, const uint8_t* __restrict__ valid,
                     const float* __restrict__ bg, int64_t R, int SR, float* __restrict__ ray_color, float* __restrict__ opacity,
                     float* __restrict__ acc_t, float* __restrict__ blend_w, float* __restrict__ bg_t)
{
    const int lane = lane_id();
    const int64_t warp0 = (int64_t)blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * COMP_WARPS;
    for (int64_t r = warp0; r < R; r += nwarps) {
        float carry = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f;
        for (int base = 0; base < SR; base += 32) {
            const int s = base + lane;
            const bool act = s < SR;
            const int64_t i = r * SR + s;
            float4 f = act ? __ldg(decoded + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            float dist = act ? __ldg(ray_dist + i) : 0.f;
            float v = (act && __ldg(valid + i)) ? 1.0f : 0.0f;
            float sigma = f.x * v;
            float o = 1.0f - expf(-sigma * dist);
            float x = act ? (1.0f - o) + 1e-10f : 1.0f;
            float incl = warp_incl_prod(x);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            float T = carry * excl;
            float w = BLEND == 0 ? o * T : o * T * T;
            if (act) {
                if (opacity) opacity[i] = o;
                if (acc_t) acc_t[i] = T;
                if (blend_w) blend_w[i] = w;
                cr += f.y * w; cg += f.z * w; cb += f.w * w;
            }
            carry *= __shfl_sync(0xffffffffu, incl, 31);
        }
        cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
        if (lane == 0) {
            if (ray_color) {
                float b0 = 0.f, b1 = 0.f, b2 = 0.f;
                if (bg) { b0 = bg[0]; b1 = bg[1]; b2 = bg[2]; }
                ray_color[r * 3 + 0] = cr + b0 * carry;
                ray_color[r * 3 + 1] = cg + b1 * carry;
                ray_color[r * 3 + 2] = cb + b2 * carry;
            }
            if (bg_t) bg_t[r] = carry;
        }
    }
}

// Backward.  With x_i = 1 - o_i + eps, T_i = prod_{j<i} x_j, T_end = prod_j x_j, w_i = o_i T_i (alpha) or
// o_i T_i^2 (alpha2), colour = sum_i w_i c_i + bg T_end:
//   dL/dc_i  = w_i g_colour
//   gw_i     = g_colour . c_i + g_blend_i
//   dL/do_i  = gw_i dw/do + g_opacity_i  -  (1/x_i) [ sum_{j>i} gT_j T_j + gT_end T_end ]
//   gT_j     = gw_j dw_j/dT_j ,  gT_end = g_colour . bg + g_bgT
//   dL/dsigma_i = dL/do_i * dist_i * (1 - o_i) ;  d decoded.x = valid * dL/dsigma
template <int BLEND>
__global__ void __launch_bounds__(COMP_WARPS * 32)
composite_bwd_kernel(const float4* __restrict__ decoded, const float* __restrict__ ray_dist, const uint8_t* __restrict__ valid,
                     const float* __restrict__ bg, int64_t R, int SR, const float* __restrict__ g_color,
                     const float* __restrict__ g_opacity, const float* __restrict__ g_blend, const float* __restrict__ g_bgt,
                     float4* __restrict__ d_decoded)
{
    const int lane = lane_id();
    const int64_t warp0 = (int64_t)blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * COMP_WARPS;
    const int nchunk = (SR + 31) / 32;
    for (int64_t r = warp0; r < R; r += nwarps) {
        // pass 1: carry-in of every chunk (product of x over all earlier chunks)
        float carry_in[COMP_MAX_CHUNKS];
        float carry = 1.0f;
#pragma unroll 1
        for (int c = 0; c < nchunk; c++) {
            const int s = c * 32 + lane;
            const bool act = s < SR;
            const int64_t i = r * SR + s;
            float x = 1.0f;
            if (act) {
                float v = __ldg(valid + i) ? 1.0f : 0.0f;
                float o = 1.0f - expf(-(__ldg(decoded + i).x * v) * __ldg(ray_dist + i));
                x = (1.0f - o) + 1e-10f;
            }
            carry_in[c] = carry;
            float incl = warp_incl_prod(x);
            carry *= __shfl_sync(0xffffffffu, incl, 31);
        }
        const float T_end = carry;
        float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f;
        if (g_color) { gc0 = g_color[r * 3]; gc1 = g_color[r * 3 + 1]; gc2 = g_color[r * 3 + 2]; }
        float gT_end = g_bgt ? g_bgt[r] : 0.f;
        if (bg) gT_end += gc0 * bg[0] + gc1 * bg[1] + gc2 * bg[2];
        float suffix = gT_end * T_end;  // sum over later samples of gT_j T_j, plus the background term
        // pass 2: chunks in reverse
#pragma unroll 1
        for (int c = nchunk - 1; c >= 0; c--) {
            const int s = c * 32 + lane;
            const bool act = s < SR;
            const int64_t i = r * SR + s;
            float4 f = act ? __ldg(decoded + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            float dist = act ? __ldg(ray_dist + i) : 0.f;
            float v = (act && __ldg(valid + i)) ? 1.0f : 0.0f;
            float o = 1.0f - expf(-(f.x * v) * dist);
            float x = act ? (1.0f - o) + 1e-10f : 1.0f;
            float incl = warp_incl_prod(x);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            float T = carry_in[c] * excl;
            float w = BLEND == 0 ? o * T : o * T * T;
            float gw = gc0 * f.y + gc1 * f.z + gc2 * f.w + ((act && g_blend) ? g_blend[i] : 0.f);
            float dw_do = BLEND == 0 ? T : T * T;
            float dw_dT = BLEND == 0 ? o : 2.0f * o * T;
            float gT_T = act ? gw * dw_dT * T : 0.f;
            float rs = warp_rincl_sum(gT_T);            // inclusive suffix within the chunk
            float later = (rs - gT_T) + suffix;         // strictly later samples + later chunks + background
            float go = gw * dw_do + ((act && g_opacity) ? g_opacity[i] : 0.f) - later / x;
            if (act) {
                float dsig = go * dist * (1.0f - o) * v;
                d_decoded[i] = make_float4(dsig, w * gc0, w * gc1, w * gc2);
            }
            suffix += __shfl_sync(0xffffffffu, rs, 0);
        }
    }
}

// ray_dist glue, models/neural_points_volumetric_model.py:569-577.
__global__ void __launch_bounds__(COMP_WARPS * 32)
ray_dist_kernel(const float* __restrict__ loc_pers, const uint8_t* __restrict__ ray_valid, float vsize_z, int mode_unit,
                int64_t R, int SR, float* __restrict__ out)
{
    const int lane = lane_id();
    const int64_t warp0 = (int64_t)blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * COMP_WARPS;
    for (int64_t r = warp0; r < R; r += nwarps) {
        float carry = -INFINITY;
        for (int base = 0; base < SR; base += 32) {
            const int s = base + lane;
            const bool act = s < SR;
            const int64_t i = r * SR + s;
            float z = act ? __ldg(loc_pers + i * 3 + 2) : -INFINITY;
            float cm = fmaxf(carry, warp_incl_max(z));
            float d;
            if (s + 1 < SR) {
                float zn = __ldg(loc_pers + (i + 1) * 3 + 2);
                d = fmaxf(cm, zn) - cm;
            } else {
                d = vsize_z;
            }
            bool bad = d < 1e-8f;
            if (mode_unit > 0) bad = bad || (d > 2.0f * vsize_z);
            float m = bad ? 1.0f : 0.0f;
            d = d * (1.0f - m) + m * vsize_z;
            if (act) out[i] = d * (__ldg(ray_valid + i) ? 1.0f : 0.0f);
            carry = __shfl_sync(0xffffffffu, cm, 31);
        }
    }
}

// Inference frame tail in one pass: step sizes (ray_dist_kernel) -> compositing (composite_fwd_kernel) -> fill_invalid, per ray,
// nothing but the outputs written: reads SR*(16 + 12 + 1) + 1 B, writes SR*4 (opacity) + 16 B per ray.  Rays that missed the cloud
// (ray_mask <= 0) get the background without reading their samples -- the same values the three separate kernels produce.
template <int BLEND>
__global__ void __launch_bounds__(COMP_WARPS * 32)
render_composite_kernel(const float4* __restrict__ decoded, const float* __restrict__ zsrc, int zstride, const uint8_t* __restrict__ valid,
                        const int8_t* __restrict__ ray_mask, float vsize_z, int mode_unit, const float* __restrict__ bg, int64_t R, int SR,
                        float* __restrict__ ray_color, float* __restrict__ opacity, float* __restrict__ bg_t, float* __restrict__ depth)
{
    const int lane = lane_id();
    const int64_t warp0 = (int64_t)blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * COMP_WARPS;
    const float b0 = bg ? bg[0] : 0.f, b1 = bg ? bg[1] : 0.f, b2 = bg ? bg[2] : 0.f;
    // The per-ray chain mask -> valid -> decoded is three dependent memory latencies; the first two (and the depth) of the NEXT ray are
    // requested before the current ray is processed, so a warp waits for one latency per ray, not three.
    const bool pipelined = SR <= 32;
    int8_t m_next = 1;
    uint8_t v_next = 0;
    float z_next = -INFINITY;
    if (pipelined && warp0 < R) {
        m_next = ray_mask ? ray_mask[warp0] : (int8_t)1;
        if (lane < SR) { v_next = __ldg(valid + warp0 * SR + lane); z_next = __ldg(zsrc + (warp0 * SR + lane) * zstride); }
    }
    for (int64_t r = warp0; r < R; r += nwarps) {
        const int8_t m_cur = pipelined ? m_next : (ray_mask ? ray_mask[r] : (int8_t)1);
        const uint8_t v_cur = v_next;
        const float z_cur = z_next;
        if (pipelined && r + nwarps < R) {
            const int64_t rn = r + nwarps;
            m_next = ray_mask ? ray_mask[rn] : (int8_t)1;
            if (lane < SR) { v_next = __ldg(valid + rn * SR + lane); z_next = __ldg(zsrc + (rn * SR + lane) * zstride); }
        }
        if (m_cur <= 0) {
            if (opacity)
                for (int s = lane; s < SR; s += 32) opacity[r * SR + s] = 0.f;
            if (lane == 0) {
                if (ray_color) { ray_color[r * 3] = b0; ray_color[r * 3 + 1] = b1; ray_color[r * 3 + 2] = b2; }
                if (bg_t) bg_t[r] = 1.0f;
                if (depth) depth[r] = 0.f;
            }
            continue;
        }
        float carry = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, zcarry = -INFINITY, dsum = 0.f, wsum = 0.f;
        for (int base = 0; base < SR; base += 32) {
            const int s = base + lane;
            const bool act = s < SR;
            const int64_t i = r * SR + s;
            const float v = (act && (pipelined ? v_cur : __ldg(valid + i))) ? 1.0f : 0.0f;
            // a sample without neighbours has sigma * valid = 0, hence weight 0: its 16 bytes are not read (two thirds of a frame's slots)
            const float4 f = v > 0.f ? __ldg(decoded + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            // step size: running maximum of the camera depth, difference to the next sample, voxel size where degenerate
            // camera depth of the sample: zsrc = loc_pers + 2 with stride 3, or the dense [R,SR] depth array with stride 1
            const float z = act ? (pipelined ? z_cur : __ldg(zsrc + i * zstride)) : -INFINITY;
            const float cm = fmaxf(zcarry, warp_incl_max(z));
            float zn = __shfl_down_sync(0xffffffffu, z, 1);
            if (lane == 31 && s + 1 < SR) zn = __ldg(zsrc + (i + 1) * zstride);
            float d = (s + 1 < SR) ? fmaxf(cm, zn) - cm : vsize_z;
            bool bad = d < 1e-8f;
            if (mode_unit > 0) bad = bad || (d > 2.0f * vsize_z);
            const float m = bad ? 1.0f : 0.0f;
            d = d * (1.0f - m) + m * vsize_z;
            const float dist = act ? d * v : 0.f;
            zcarry = __shfl_sync(0xffffffffu, cm, 31);
            // compositing
            const float sigma = f.x * v;
            const float o = 1.0f - expf(-sigma * dist);
            const float x = act ? (1.0f - o) + 1e-10f : 1.0f;
            const float incl = warp_incl_prod(x);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float T = carry * excl;
            const float w = BLEND == 0 ? o * T : o * T * T;
            if (act) {
                if (opacity) opacity[i] = o;
                cr += f.y * w; cg += f.z * w; cb += f.w * w;
                dsum += (o * T) * z; wsum += o * T;           // depth: opacity * acc_transmission weights whatever the blend (:621)
            }
            carry *= __shfl_sync(0xffffffffu, incl, 31);
        }
        cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
        if (depth) { dsum = warp_sum(dsum); wsum = warp_sum(wsum); }
        if (lane == 0) {
            if (ray_color) {
                ray_color[r * 3 + 0] = cr + b0 * carry;
                ray_color[r * 3 + 1] = cg + b1 * carry;
                ray_color[r * 3 + 2] = cb + b2 * carry;
            }
            if (bg_t) bg_t[r] = carry;
            if (depth) depth[r] = dsum / (wsum + 1e-6f);
        }
    }
}

// The same frame tail with one THREAD per ray (SR <= 40): the warp-per-ray kernel above spends ~320 warp instructions per ray on its
// scans and reductions (it is bound by instruction issue at 0.12 ms per 640x480 frame); a thread that walks its ray's samples in order
// needs ~25 instructions per sample.  A block owns 128 consecutive rays and stages their rows through shared memory with coalesced
// accesses: camera depth + validity of all samples, the decoded (sigma, r, g, b) rows eight samples at a time (only the samples that
// have neighbours are fetched), the opacities on the way out.  Transmittance is the running product in sample order.
constexpr int ROWS_RAYS = 128, ROWS_CHUNK = 8;

template <int BLEND>
__global__ void __launch_bounds__(ROWS_RAYS)
render_composite_rows_kernel(const float4* __restrict__ decoded, const float* __restrict__ zsrc, int zstride, const uint8_t* __restrict__ valid,
                             const int8_t* __restrict__ ray_mask, float vsize_z, int mode_unit, const float* __restrict__ bg, int64_t R, int SR,
                             float* __restrict__ ray_color, float* __restrict__ opacity, float* __restrict__ bg_t, float* __restrict__ depth)
{
    extern __shared__ float4 s_rows[];                                     // [128][CHUNK + 1] decoded rows (padded: conflict-free 16-byte reads)
    float* s_z = (float*)(s_rows + ROWS_RAYS * (ROWS_CHUNK + 1));          // [128][SR + 1] camera depth, overwritten by the opacity
    uint8_t* s_v = (uint8_t*)(s_z + ROWS_RAYS * (SR + 1));                 // [128][SR]
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS_RAYS;
    const int nr = (int)min((int64_t)ROWS_RAYS, R - r0);
    const int zs = SR + 1;
    // stage depth and validity of the block's rays (consecutive rays are consecutive in memory)
    for (int i = tid; i < nr * SR; i += ROWS_RAYS) {
        const int ray = i / SR, sm = i - ray * SR;
        s_z[ray * zs + sm] = __ldg(zsrc + (r0 * SR + i) * zstride);
        s_v[ray * SR + sm] = __ldg(valid + r0 * SR + i);
    }
    const int64_t r = r0 + tid;
    const bool live = tid < nr;
    const bool hit = live && (!ray_mask || ray_mask[r] > 0);
    const float b0 = bg ? bg[0] : 0.f, b1 = bg ? bg[1] : 0.f, b2 = bg ? bg[2] : 0.f;
    float T = 1.0f, cr = 0.f, cg = 0.f, cb = 0.f, zmax = -INFINITY, dsum = 0.f, wsum = 0.f;
    __syncthreads();
    for (int c0 = 0; c0 < SR; c0 += ROWS_CHUNK) {
        const int cn = min(ROWS_CHUNK, SR - c0);
        // decoded rows of this chunk: thread i takes (ray i / cn, sample i % cn): a warp covers whole 16 * cn-byte runs
        for (int i = tid; i < nr * cn; i += ROWS_RAYS) {
            const int ray = i / cn, j = i - ray * cn;
            if (s_v[ray * SR + c0 + j]) s_rows[ray * (ROWS_CHUNK + 1) + j] = __ldg(decoded + (r0 + ray) * SR + c0 + j);
        }
        __syncthreads();
        if (live) {
            for (int j = 0; j < cn; j++) {
                const int sm = c0 + j;
                const float v = s_v[tid * SR + sm] ? 1.0f : 0.0f;
                const float z = s_z[tid * zs + sm];
                // step size: running maximum of the camera depth, difference to the next sample, voxel size where degenerate
                const float cm = fmaxf(zmax, z);
                float d = (sm + 1 < SR) ? fmaxf(cm, s_z[tid * zs + sm + 1]) - cm : vsize_z;
                bool bad = d < 1e-8f;
                if (mode_unit > 0) bad = bad || (d > 2.0f * vsize_z);
                const float m = bad ? 1.0f : 0.0f;
                d = d * (1.0f - m) + m * vsize_z;
                zmax = cm;
                float o = 0.f;
                if (hit) {
                    const float4 f = v > 0.f ? s_rows[tid * (ROWS_CHUNK + 1) + j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    o = 1.0f - expf(-(f.x * v) * (d * v));
                    const float w = BLEND == 0 ? o * T : o * T * T;
                    cr += f.y * w; cg += f.z * w; cb += f.w * w;
                    dsum += (o * T) * z; wsum += o * T;
                    T *= (1.0f - o) + 1e-10f;
                }
                s_z[tid * zs + sm] = o;                                      // the depth of this sample is not needed again
            }
        }
        __syncthreads();
    }
    if (live) {
        if (ray_color) {
            ray_color[r * 3 + 0] = hit ? cr + b0 * T : b0;
            ray_color[r * 3 + 1] = hit ? cg + b1 * T : b1;
            ray_color[r * 3 + 2] = hit ? cb + b2 * T : b2;
        }
        if (bg_t) bg_t[r] = hit ? T : 1.0f;
        if (depth) depth[r] = hit ? dsum / (wsum + 1e-6f) : 0.f;
    }
    if (opacity)
        for (int i = tid; i < nr * SR; i += ROWS_RAYS) {
            const int ray = i / SR, sm = i - ray * SR;
            opacity[r0 * SR + i] = s_z[ray * zs + sm];
        }
}

// `prob == 1` outputs (models/neural_points_volumetric_model.py:633-656): one warp per ray.  Lanes find the first sample of largest
// opacity (torch.max returns the first maximal index), then lane k < K reads neighbour k of that sample -- invalid slots read point 0,
// as the reference's clamp(pidx, 0) gather does -- and the K-wide reductions are warp shuffles.  Rays with ray_mask <= 0 (no row in
// the reference's compacted tensors) get zeros.
__global__ void __launch_bounds__(COMP_WARPS * 32)
probe_kernel(const float* __restrict__ opacity, const float* __restrict__ loc_w, const int32_t* __restrict__ pidx, const float* __restrict__ weight,
             const float* __restrict__ conf_coef, const int8_t* __restrict__ ray_mask, SgnPointTables tab, int C, int64_t R, int SR, int K,
             float* __restrict__ max_opacity, float* __restrict__ max_loc_w, float* __restrict__ far_dist, float* __restrict__ avg_color,
             float* __restrict__ avg_dir, float* __restrict__ avg_conf, float* __restrict__ avg_emb)
{
    const int lane = lane_id();
    const int64_t warp0 = (int64_t)blockIdx.x * COMP_WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * COMP_WARPS;
    for (int64_t r = warp0; r < R; r += nwarps) {
        const bool hit = !ray_mask || ray_mask[r] > 0;
        float best = -INFINITY;
        int bi = 0x7fffffff;
        if (hit)
            for (int s = lane; s < SR; s += 32) {
                const float o = opacity[r * SR + s];
                if (o > best) { best = o; bi = s; }          // strict: keeps the first maximum of this lane's samples
            }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (!hit || bi >= SR) { best = 0.f; bi = 0; }
        const int64_t smp = r * SR + bi;
        const float lx = hit ? loc_w[3 * smp] : 0.f, ly = hit ? loc_w[3 * smp + 1] : 0.f, lz = hit ? loc_w[3 * smp + 2] : 0.f;
        float w = 0.f, dmin = INFINITY, cr = 0.f, cg = 0.f, cb = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, cf = 0.f;
        int p = 0;
        // K <= 32: lane k owns neighbour k
        if (hit && lane < K) {
            const int pi = pidx[smp * K + lane];
            p = pi < 0 ? 0 : pi;
            w = weight[smp * K + lane] * conf_coef[smp * K + lane];
            const float ex = tab.xyz[3 * (int64_t)p] - lx, ey = tab.xyz[3 * (int64_t)p + 1] - ly, ez = tab.xyz[3 * (int64_t)p + 2] - lz;
            dmin = sqrtf(ex * ex + ey * ey + ez * ez);
            cr = tab.color[3 * (int64_t)p] * w; cg = tab.color[3 * (int64_t)p + 1] * w; cb = tab.color[3 * (int64_t)p + 2] * w;
            dx = tab.dir[3 * (int64_t)p] * w; dy = tab.dir[3 * (int64_t)p + 1] * w; dz = tab.dir[3 * (int64_t)p + 2] * w;
            cf = (tab.conf ? tab.conf[p] : 1.0f) * w;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, off));
        cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
        dx = warp_sum(dx); dy = warp_sum(dy); dz = warp_sum(dz); cf = warp_sum(cf);
        if (lane == 0) {
            if (max_opacity) max_opacity[r] = best;
            if (max_loc_w) { max_loc_w[3 * r] = lx; max_loc_w[3 * r + 1] = ly; max_loc_w[3 * r + 2] = lz; }
            if (far_dist) far_dist[r] = hit ? dmin : 0.f;
            if (avg_color) { avg_color[3 * r] = cr; avg_color[3 * r + 1] = cg; avg_color[3 * r + 2] = cb; }
            if (avg_dir) { avg_dir[3 * r] = dx; avg_dir[3 * r + 1] = dy; avg_dir[3 * r + 2] = dz; }
            if (avg_conf) avg_conf[r] = cf;
        }
        if (avg_emb)
            for (int c = lane; c < C; c += 32) {           // embedding average: lane = channel, the K weights broadcast one by one
                float acc = 0.f;
                for (int k = 0; k < K; k++) {
                    const float wk = __shfl_sync(0xffffffffu, w, k);
                    const int pk = __shfl_sync(0xffffffffu, p, k);
                    acc += hit ? tab.embedding[(int64_t)pk * C + c] * wk : 0.f;
                }
                avg_emb[r * C + c] = acc;
            }
    }
}

__global__ void fill_invalid_kernel(const int8_t* __restrict__ ray_mask, const float* __restrict__ bg, int64_t R, int SR,
                                    float* ray_color, float* opacity, float* bg_t)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R || ray_mask[r] > 0) return;
    if (ray_color) { ray_color[r * 3] = bg[0]; ray_color[r * 3 + 1] = bg[1]; ray_color[r * 3 + 2] = bg[2]; }
    if (bg_t) bg_t[r] = 1.0f;
    if (opacity)
        for (int s = 0; s < SR; s++) opacity[r * SR + s] = 0.f;
}

static int comp_grid(int64_t R)
{
    int64_t b = (R + COMP_WARPS - 1) / COMP_WARPS;
    int64_t cap = 148 * 8 * 4;  // a few waves of 148 SMs x 8 resident blocks
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_ray_dist(const float* loc_pers, const uint8_t* ray_valid, float vsize_z, int raydist_mode_unit, int64_t R,
                            int SR, float* ray_dist, void* stream)
{
    SGN_CHECK_ARG(R >= 0 && SR > 0, "sgn_ray_dist: bad R/SR");
    if (R == 0) return SGN_OK;
    launch(ray_dist_kernel, comp_grid(R), COMP_WARPS * 32, 0, (cudaStream_t)stream, loc_pers, ray_valid, vsize_z, raydist_mode_unit, R, SR, ray_dist);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_composite_forward(const float* decoded, const float* ray_dist, const uint8_t* valid, const float* bg, int blend,
                                     int64_t R, int SR, float* ray_color, float* opacity, float* acc_transmission,
                                     float* blend_weight, float* bg_transmission, void* stream)
{
    SGN_CHECK_ARG(R >= 0 && SR > 0 && SR <= 32 * COMP_MAX_CHUNKS, "sgn_composite_forward: SR=%d out of range", SR);
    SGN_CHECK_ARG(blend == 0 || blend == 1, "sgn_composite_forward: blend must be 0 (alpha) or 1 (alpha2)");
    if (R == 0) return SGN_OK;
    auto st = (cudaStream_t)stream;
    if (blend == 0)
        launch(composite_fwd_kernel<0>, comp_grid(R), COMP_WARPS * 32, 0, st, (const float4*)decoded, ray_dist, valid, bg, R, SR, ray_color,
                                                                         opacity, acc_transmission, blend_weight, bg_transmission);
    else
        launch(composite_fwd_kernel<1>, comp_grid(R), COMP_WARPS * 32, 0, st, (const float4*)decoded, ray_dist, valid, bg, R, SR, ray_color,
                                                                         opacity, acc_transmission, blend_weight, bg_transmission);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_composite_backward(const float* decoded, const float* ray_dist, const uint8_t* valid, const float* bg, int blend,
                                      int64_t R, int SR, const float* d_ray_color, const float* d_opacity,
                                      const float* d_blend_weight, const float* d_bg_transmission, float* d_decoded, void* stream)
{
    SGN_CHECK_ARG(R >= 0 && SR > 0 && SR <= 32 * COMP_MAX_CHUNKS, "sgn_composite_backward: SR=%d out of range", SR);
    SGN_CHECK_ARG(blend == 0 || blend == 1, "sgn_composite_backward: blend must be 0 or 1");
    if (R == 0) return SGN_OK;
    auto st = (cudaStream_t)stream;
    if (blend == 0)
        launch(composite_bwd_kernel<0>, comp_grid(R), COMP_WARPS * 32, 0, st, (const float4*)decoded, ray_dist, valid, bg, R, SR, d_ray_color,
                                                                         d_opacity, d_blend_weight, d_bg_transmission, (float4*)d_decoded);
    else
        launch(composite_bwd_kernel<1>, comp_grid(R), COMP_WARPS * 32, 0, st, (const float4*)decoded, ray_dist, valid, bg, R, SR, d_ray_color,
                                                                         d_opacity, d_blend_weight, d_bg_transmission, (float4*)d_decoded);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

static int render_composite_impl(const float* decoded, const float* zsrc, int zstride, const uint8_t* ray_valid, const int8_t* ray_mask, float vsize_z,
                                 int raydist_mode_unit, const float* bg, int blend, int64_t R, int SR, float* ray_color, float* opacity,
                                 float* bg_transmission, float* depth, void* stream)
{
    SGN_CHECK_ARG(R >= 0 && SR > 0 && SR <= 32 * COMP_MAX_CHUNKS, "sgn_render_composite: SR=%d out of range", SR);
    SGN_CHECK_ARG(blend == 0 || blend == 1, "sgn_render_composite: blend must be 0 (alpha) or 1 (alpha2)");
    SGN_CHECK_ARG(decoded && zsrc && ray_valid, "sgn_render_composite: NULL input");
    if (R == 0) return SGN_OK;
    auto st = (cudaStream_t)stream;
    if (SR <= 40) {                                        // its shared memory stays under the 48 KB a kernel gets without opting in
        const size_t sm = sizeof(float4) * ROWS_RAYS * (ROWS_CHUNK + 1) + sizeof(float) * ROWS_RAYS * (SR + 1) + (size_t)ROWS_RAYS * SR;
        if (blend == 0)
            launch(render_composite_rows_kernel<0>, cdiv(R, ROWS_RAYS), ROWS_RAYS, sm, st, (const float4*)decoded, zsrc, zstride, ray_valid, ray_mask,
                   vsize_z, raydist_mode_unit, bg, R, SR, ray_color, opacity, bg_transmission, depth);
        else
            launch(render_composite_rows_kernel<1>, cdiv(R, ROWS_RAYS), ROWS_RAYS, sm, st, (const float4*)decoded, zsrc, zstride, ray_valid, ray_mask,
                   vsize_z, raydist_mode_unit, bg, R, SR, ray_color, opacity, bg_transmission, depth);
        SGN_LAUNCH_CHECK();
        return SGN_OK;
    }
    if (blend == 0)
        launch(render_composite_kernel<0>, comp_grid(R), COMP_WARPS * 32, 0, st, (const float4*)decoded, zsrc, zstride, ray_valid, ray_mask, vsize_z,
                                                                            raydist_mode_unit, bg, R, SR, ray_color, opacity, bg_transmission, depth);
    else
        launch(render_composite_kernel<1>, comp_grid(R), COMP_WARPS * 32, 0, st, (const float4*)decoded, zsrc, zstride, ray_valid, ray_mask, vsize_z,
                                                                            raydist_mode_unit, bg, R, SR, ray_color, opacity, bg_transmission, depth);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_render_composite(const float* decoded, const float* loc_pers, const uint8_t* ray_valid, const int8_t* ray_mask, float vsize_z,
                                    int raydist_mode_unit, const float* bg, int blend, int64_t R, int SR, float* ray_color, float* opacity,
                                    float* bg_transmission, float* depth, void* stream)
{
    return render_composite_impl(decoded, loc_pers ? loc_pers + 2 : nullptr, 3, ray_valid, ray_mask, vsize_z, raydist_mode_unit, bg, blend, R, SR,
                                 ray_color, opacity, bg_transmission, depth, stream);
}

extern "C" int sgn_render_composite_depth(const float* decoded, const float* loc_depth, const uint8_t* ray_valid, const int8_t* ray_mask, float vsize_z,
                                          int raydist_mode_unit, const float* bg, int blend, int64_t R, int SR, float* ray_color, float* opacity,
                                          float* bg_transmission, float* depth, void* stream)
{
    return render_composite_impl(decoded, loc_depth, 1, ray_valid, ray_mask, vsize_z, raydist_mode_unit, bg, blend, R, SR, ray_color, opacity,
                                 bg_transmission, depth, stream);
}

extern "C" int sgn_probe_outputs(const float* opacity, const float* sample_loc_w, const int32_t* sample_pidx, const float* weight,
                                 const float* conf_coef, const int8_t* ray_mask, const SgnPointTables* tables, int feat_dim, int64_t R, int SR,
                                 int K, float* ray_max_shading_opacity, float* ray_max_sample_loc_w, float* ray_max_far_dist,
                                 float* shading_avg_color, float* shading_avg_dir, float* shading_avg_conf, float* shading_avg_embedding,
                                 void* stream)
{
    SGN_CHECK_ARG(R >= 0 && SR > 0 && K > 0 && K <= 32, "sgn_probe_outputs: bad R/SR/K (K at most 32)");
    SGN_CHECK_ARG(opacity && sample_loc_w && sample_pidx && weight && conf_coef && tables, "sgn_probe_outputs: NULL input");
    SGN_CHECK_ARG(tables->xyz && tables->color && tables->dir && (tables->embedding || !shading_avg_embedding), "sgn_probe_outputs: missing table");
    SGN_CHECK_ARG(feat_dim % 32 == 0 || !shading_avg_embedding, "sgn_probe_outputs: feat_dim must be a multiple of 32");
    if (R == 0) return SGN_OK;
    launch(probe_kernel, comp_grid(R), COMP_WARPS * 32, 0, (cudaStream_t)stream, opacity, sample_loc_w, sample_pidx, weight, conf_coef, ray_mask,
           *tables, feat_dim, R, SR, K, ray_max_shading_opacity, ray_max_sample_loc_w, ray_max_far_dist, shading_avg_color, shading_avg_dir,
           shading_avg_conf, shading_avg_embedding);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}

extern "C" int sgn_fill_invalid(const int8_t* ray_mask, const float* bg, int64_t R, int SR, float* ray_color, float* opacity,
                                float* bg_transmission, void* stream)
{
    if (R == 0) return SGN_OK;
    launch(fill_invalid_kernel, cdiv(R, 256), 256, 0, (cudaStream_t)stream, ray_mask, bg, R, SR, ray_color, opacity, bg_transmission);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
}
