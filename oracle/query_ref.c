/*
 * oracle/query_ref.c -- TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
 *
 * Sequential ("thread-index order") CPU restatement of the reference's world-coordinate
 * neural-point query kernels.  Every function cites the reference lines it follows; paths are
 * relative to the reference root, file Q = models/neural_points/query_point_indices_worldcoords.py.
 *
 * The reference kernels use atomics, so slot order / list order depend on thread arrival.
 * The canonical form restated here is "threads run one after another in index order"
 * (hard part 1 and appendix A.3-A.5); wall-clock `seconds` is an input.
 *
 * Floating point: compile with -ffp-contract=off.  Wherever nvcc (default -fmad=true)
 * contracts the reference source into an FMA this file calls fmaf() explicitly; everything
 * else is plain IEEE fp32 (the reference is built without fast-math, so '/' is div.rn).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- cuRAND XORWOW, seq 0 / offset 0 (CUDA toolkit curand_kernel.h: _curand_init_scratch,
 * curand(), curand_uniform.h:_curand_uniform).  Call sites: Q:313-315 and Q:401-403. ---- */
static float orc_curand_uniform_first(uint64_t seed)
{
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    uint32_t d = 6615241u + t1 + t0;
    uint32_t v0 = 123456789u + t0;
    uint32_t v4 = 5783321u + t0;
    /* one curand() step; only v[0], v[4] and d take part in the first output */
    uint32_t t = v0 ^ (v0 >> 2);
    v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
    d += 362437u;
    uint32_t x = v4 + d;
    /* x * 2^-32 + 2^-33, contracted to one FMA by nvcc */
    return fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}

float orc_curand_uniform(uint64_t seed) { return orc_curand_uniform_first(seed); }

/* reservoir index, Q:315 / Q:403:  int insrtidx = ceilf(curand_uniform(&state) * (tmp+1)) - 1; */
static int orc_reservoir_index(int64_t index, uint64_t seconds, int tmp)
{
    uint64_t seed = (uint64_t)index + 2ull * seconds;
    float u = orc_curand_uniform_first(seed);
    return (int)(ceilf(u * (float)(tmp + 1)) - 1.0f);
}

/* voxel coordinate of a position, Q:288-290 (also Q:388-390, Q:429-431, Q:623-625) */
static inline int orc_vox(float p, float shift, float vsize)
{
    return (int)floorf((p - shift) / vsize);
}

/*
 * Occupancy build: claim_occ (Q:265-326) -> coor_2_occ reset (Q:735) -> map_coor2occ
 * (Q:328-363, called with query_size in the kernel_size slot, Q:797) -> fill_occ2pnts
 * (Q:365-410), driven as build_occ_vox does (Q:706-778).  B = 1.
 *
 * Outputs must be pre-sized by the caller:
 *   coor_occ    int32[X*Y*Z]   (zeroed here)
 *   coor_2_occ  int32[X*Y*Z]   (set to -1 here)
 *   occ_2_coor  int32[max_o*3] (set to -1 here)
 *   occ_numpnts int32[max_o]   (zeroed here)
 *   occ_2_pnts  int32[max_o*P] (set to -1 here)
 *   occ_idx     int32[1]
 */
void orc_build_occ_vox(const float *xyz, int N, int actual_n,
                       const float *shift, const float *vsize, const int *dim, const int *query_size,
                       int max_o, int P, uint64_t seconds_claim, uint64_t seconds_fill,
                       int *coor_occ, int *coor_2_occ, int *occ_2_coor, int *occ_idx,
                       int *occ_numpnts, int *occ_2_pnts)
{
    const int64_t vol = (int64_t)dim[0] * dim[1] * dim[2];
    const int YZ = dim[1] * dim[2];
    memset(coor_occ, 0, sizeof(int) * (size_t)vol);
    for (int64_t i = 0; i < vol; i++) coor_2_occ[i] = -1;
    for (int64_t i = 0; i < (int64_t)max_o * 3; i++) occ_2_coor[i] = -1;
    memset(occ_numpnts, 0, sizeof(int) * (size_t)max_o);
    for (int64_t i = 0; i < (int64_t)max_o * P; i++) occ_2_pnts[i] = -1;
    int counter = 0;

    /* claim_occ, Q:280-325 */
    for (int i = 0; i < N; i++) {
        if (i >= actual_n) continue;
        int c0 = orc_vox(xyz[3 * i + 0], shift[0], vsize[0]);
        int c1 = orc_vox(xyz[3 * i + 1], shift[1], vsize[1]);
        int c2 = orc_vox(xyz[3 * i + 2], shift[2], vsize[2]);
        if (c0 < 0 || c0 >= dim[0] || c1 < 0 || c1 >= dim[1] || c2 < 0 || c2 >= dim[2]) continue;
        int64_t ci = (int64_t)c0 * YZ + (int64_t)c1 * dim[2] + c2;
        if (coor_2_occ[ci] == -1) {
            coor_2_occ[ci] = 0;                         /* atomicCAS(-1 -> 0), Q:297-300 */
            int tmp = counter++;                        /* atomicAdd(occ_idx, 1), Q:305 */
            if (tmp < max_o) {
                occ_2_coor[3 * tmp + 0] = c0; occ_2_coor[3 * tmp + 1] = c1; occ_2_coor[3 * tmp + 2] = c2;
            } else {                                    /* reservoir, Q:313-321 */
                int j = orc_reservoir_index(i, seconds_claim, tmp);
                if (j < max_o) {
                    occ_2_coor[3 * j + 0] = c0; occ_2_coor[3 * j + 1] = c1; occ_2_coor[3 * j + 2] = c2;
                }
            }
        }
    }
    occ_idx[0] = counter;

    /* Q:735: coor_2_occ is replaced by a fresh all -1 tensor before map_coor2occ */
    for (int64_t i = 0; i < vol; i++) coor_2_occ[i] = -1;

    /* map_coor2occ, Q:339-362 (kernel_size argument := query_size) */
    for (int s = 0; s < max_o; s++) {
        if (!(s < counter && s < max_o)) continue;
        int c0 = occ_2_coor[3 * s + 0];
        if (c0 < 0) continue;
        int c1 = occ_2_coor[3 * s + 1], c2 = occ_2_coor[3 * s + 2];
        coor_2_occ[(int64_t)c0 * YZ + (int64_t)c1 * dim[2] + c2] = s;
        int x0 = c0 - query_size[0] / 2, x1 = c0 + (query_size[0] + 1) / 2;
        int y0 = c1 - query_size[1] / 2, y1 = c1 + (query_size[1] + 1) / 2;
        int z0 = c2 - query_size[2] / 2, z1 = c2 + (query_size[2] + 1) / 2;
        if (x0 < 0) x0 = 0; if (x1 > dim[0]) x1 = dim[0];
        if (y0 < 0) y0 = 0; if (y1 > dim[1]) y1 = dim[1];
        if (z0 < 0) z0 = 0; if (z1 > dim[2]) z1 = dim[2];
        for (int x = x0; x < x1; x++)
            for (int y = y0; y < y1; y++)
                for (int z = z0; z < z1; z++)
                    coor_occ[(int64_t)x * YZ + (int64_t)y * dim[2] + z] = 1;
    }

    /* fill_occ2pnts, Q:381-409.  Note the `voxel_idx > 0` guard (Q:395): slot 0 never gets points. */
    for (int i = 0; i < N; i++) {
        if (i >= actual_n) continue;
        int c0 = orc_vox(xyz[3 * i + 0], shift[0], vsize[0]);
        int c1 = orc_vox(xyz[3 * i + 1], shift[1], vsize[1]);
        int c2 = orc_vox(xyz[3 * i + 2], shift[2], vsize[2]);
        if (c0 < 0 || c0 >= dim[0] || c1 < 0 || c1 >= dim[1] || c2 < 0 || c2 >= dim[2]) continue;
        int v = coor_2_occ[(int64_t)c0 * YZ + (int64_t)c1 * dim[2] + c2];
        if (v > 0) {
            int tmp = occ_numpnts[v]++;
            if (tmp < P) {
                occ_2_pnts[(int64_t)v * P + tmp] = i;
            } else {
                int j = orc_reservoir_index(i, seconds_fill, tmp);
                if (j < P) occ_2_pnts[(int64_t)v * P + j] = i;
            }
        }
    }
}

/* mask_raypos, Q:413-437.  raypos [R*D*3], mask [R*D] must be zeroed by the caller (Q:811). */
void orc_mask_raypos(const float *raypos, const int *coor_occ, int64_t RD,
                     const float *shift, const int *dim, const float *vsize, int *raypos_mask)
{
    const int YZ = dim[1] * dim[2];
    for (int64_t i = 0; i < RD; i++) {
        int c0 = orc_vox(raypos[3 * i + 0], shift[0], vsize[0]);
        int c1 = orc_vox(raypos[3 * i + 1], shift[1], vsize[1]);
        int c2 = orc_vox(raypos[3 * i + 2], shift[2], vsize[2]);
        if (c0 >= 0 && c0 < dim[0] && c1 >= 0 && c1 < dim[1] && c2 >= 0 && c2 < dim[2])
            raypos_mask[i] = coor_occ[(int64_t)c0 * YZ + (int64_t)c1 * dim[2] + c2];
    }
}

/* get_shadingloc / get_shadingloc_with_semantic, Q:439-487.  raylabel/sample_label may be NULL. */
void orc_get_shadingloc(const float *raypos, const int *raylabel, const int *raypos_slot,
                        int R, int D, int SR, float *sample_loc, int *sample_label, int *sample_loc_mask)
{
    for (int64_t i = 0; i < (int64_t)R * D; i++) {
        int temp = raypos_slot[i];
        if (temp >= 0) {
            int r = (int)(i / D);
            int64_t li = (int64_t)r * SR + temp;
            sample_loc[3 * li + 0] = raypos[3 * i + 0];
            sample_loc[3 * li + 1] = raypos[3 * i + 1];
            sample_loc[3 * li + 2] = raypos[3 * i + 2];
            if (raylabel && sample_label) sample_label[li] = raylabel[i];
            sample_loc_mask[li] = 1;
        }
    }
}

/*
 * query_neigh_along_ray_layered (Q:594-681) and ..._semantic_guidance (Q:489-591).
 * in_label == NULL selects the plain kernel.  in_label_prob_bits is the int32 tensor the
 * reference passes where the kernel declares `const float*` (Q:492 vs Q:916): it is read as
 * float bits, multiplied by 10 and truncated to int (Q:549).
 * Squared distance: nvcc contracts `x*x + y*y + z*z` (Q:651) to fma(z,z, fma(x,x, y*y))
 * (checked in the SASS of the reference source built for sm_100a, see oracle/build_ref.py).
 * sample_pidx [R*SR*K] must be -1-filled by the caller (Q:836).
 */
void orc_query_neigh(const float *xyz, const int *in_label, const int *in_label_prob_bits,
                     int R, int SR, int P, int K, float radius2,
                     const float *shift, const int *dim, const float *vsize, const int *kernel_size,
                     const int *occ_numpnts, const int *occ_2_pnts, const int *coor_2_occ,
                     const float *sample_loc, const int *sample_loc_mask, const int *sample_label,
                     int *sample_pidx, uint64_t seconds)
{
    const int YZ = dim[1] * dim[2];
    float *buf = (float *)malloc(sizeof(float) * (size_t)K);
    for (int64_t idx = 0; idx < (int64_t)R * SR; idx++) {
        if (sample_loc_mask[idx] <= 0) continue;
        float cx = sample_loc[3 * idx + 0], cy = sample_loc[3 * idx + 1], cz = sample_loc[3 * idx + 2];
        int center_label = (in_label && sample_label) ? sample_label[idx] : 0;
        int fx = orc_vox(cx, shift[0], vsize[0]);
        int fy = orc_vox(cy, shift[1], vsize[1]);
        int fz = orc_vox(cz, shift[2], vsize[2]);
        int kid = 0, far_ind = 0;
        float far2 = 0.0f;
        int nlayer = (kernel_size[0] + 1) / 2;
        for (int layer = 0; layer < nlayer; layer++) {
            int xlo = -fx > -layer ? -fx : -layer, xhi = dim[0] - fx < layer + 1 ? dim[0] - fx : layer + 1;
            int ylo = -fy > -layer ? -fy : -layer, yhi = dim[1] - fy < layer + 1 ? dim[1] - fy : layer + 1;
            int zlo = -fz > -layer ? -fz : -layer, zhi = dim[2] - fz < layer + 1 ? dim[2] - fz : layer + 1;
            for (int x = xlo; x < xhi; x++) {
                for (int y = ylo; y < yhi; y++) {
                    for (int z = zlo; z < zhi; z++) {
                        int ax = abs(x), ay = abs(y), az = abs(z);
                        int m = ax > ay ? ax : ay; if (az > m) m = az;
                        if (m != layer) continue;
                        int occ = coor_2_occ[(int64_t)(fx + x) * YZ + (int64_t)(fy + y) * dim[2] + (fz + z)];
                        if (occ < 0) continue;
                        int cnt = occ_numpnts[occ] < P ? occ_numpnts[occ] : P;
                        for (int g = 0; g < cnt; g++) {
                            int pidx = occ_2_pnts[(int64_t)occ * P + g];
                            if (in_label) {                              /* Q:548-553 */
                                int label_v = in_label[pidx];
                                float pf; int32_t bits = in_label_prob_bits[(int64_t)pidx * 20 + label_v];
                                memcpy(&pf, &bits, 4);
                                int label_prob = (int)(pf * 10.0f);
                                /* `seconds` is unsigned long: the int (1-label_prob) is converted to
                                 * unsigned for the comparison, so a negative value compares as huge. */
                                int ok = (center_label == label_v) || (label_v == 0) || (center_label == 0) ||
                                         ((center_label != label_v) && ((seconds % 10) <= (uint64_t)(int64_t)(1 - label_prob)));
                                if (!ok) continue;
                            }
                            float xv = xyz[3 * pidx + 0] - cx;
                            float yv = xyz[3 * pidx + 1] - cy;
                            float zv = xyz[3 * pidx + 2] - cz;
                            float d2 = fmaf(zv, zv, fmaf(xv, xv, yv * yv));
                            if (radius2 == 0.0f || d2 <= radius2) {
                                if (kid++ < K) {
                                    sample_pidx[idx * K + kid - 1] = pidx;
                                    buf[kid - 1] = d2;
                                    if (d2 > far2) { far2 = d2; far_ind = kid - 1; }
                                } else if (d2 < far2) {
                                    sample_pidx[idx * K + far_ind] = pidx;
                                    buf[far_ind] = d2;
                                    far2 = d2;
                                    for (int i = 0; i < K; i++)
                                        if (buf[i] > far2) { far2 = buf[i]; far_ind = i; }
                                }
                            }
                        }
                    }
                }
            }
            if (kid >= K) break;
        }
    }
    free(buf);
}
