/*
 * oracle/query_pers_ref.c -- TEST INFRASTRUCTURE ONLY (never linked or called by the product path).
 *
 * Sequential ("thread-index order") CPU restatement of the reference's PERSPECTIVE-frustum neural-point query kernels
 * (--wcoord_query 0): get_occ_vox, near_vox_full, insert_vox_points, query_neigh_along_ray_layered / query_rand_along_ray and the torch
 * glue between them.  File P = models/neural_points/query_point_indices.py of the reference.  B = 1.
 *
 * The kernels use atomics; the canonical form restated here is "threads run one after another in index order" (lists in point-index
 * order).  Two overflow behaviours of the reference are NOT reproduced and must not occur in the inputs: the int8 cumsum of P:696
 * (more than 127 selected voxels in one pixel column) and a given max_o smaller than a column's selected-voxel count (P:349, :474 then
 * index into the next column's lists).
 *
 * Floating point: compile with -ffp-contract=off; nvcc's contractions are written as fmaf() where they matter (distances).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

float orc_curand_uniform(uint64_t seed);   /* query_ref.c: first curand_uniform of XORWOW(seed, 0, 0) */

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/*
 * xyz_pers [N,3]; pixel_idx [R,2] (x, y of every ray); shift = ranges[:3]; vsize = scaled voxel size; dim = scaled_vdim;
 * ray_vsize = vsize / vscale (P:711).  max_o: lists per column (>= selected voxels of any column).
 * Outputs, UNCOMPACTED (row r = input ray r; the reference compacts the rays with ray_mask > 0, P:688):
 *   ray_mask int8 [R], sample_pidx int32 [R,SR,K] (-1), sample_loc f32 [R,SR,3] (rows of masked-out rays stay 0).
 */
int orc_pers_query(const float* xyz, int N, const int* pixel_idx, int R, const float* shift, const float* vsize, const int* dim,
                   const int* vscale, const int* kernel_size, const int* query_size, const float* ray_vsize, int SR, int K, int P,
                   int max_o, float radius2, float depth2, int NN, int inverse, uint64_t seconds_insert, uint64_t seconds_query,
                   int8_t* ray_mask, int32_t* sample_pidx, float* sample_loc)
{
    const int X = dim[0], Y = dim[1], Z = dim[2];
    const int64_t pixel_size = (int64_t)X * Y, vol = pixel_size * Z;
    uint8_t* coor_occ = (uint8_t*)calloc(vol, 1);
    int8_t* loc = (int8_t*)calloc(vol, 1);                 /* loc_coor_counter */
    int* near_id = (int*)malloc(sizeof(int) * pixel_size);
    int* far_id = (int*)malloc(sizeof(int) * pixel_size);
    short* coorz = (short*)malloc(sizeof(short) * pixel_size * SR);
    uint8_t* pixel_map = (uint8_t*)calloc(pixel_size, 1);
    int* rank = (int*)malloc(sizeof(int) * vol);           /* loc_coor_counter after P:695-696 (as int: no int8 wrap) */
    short* cnt = (short*)calloc(pixel_size * max_o, sizeof(short));
    int* lists = (int*)calloc(pixel_size * max_o * (int64_t)P, sizeof(int));
    if (!coor_occ || !loc || !near_id || !far_id || !coorz || !pixel_map || !rank || !cnt || !lists) return -1;
    for (int64_t i = 0; i < pixel_size; i++) { near_id[i] = Z; far_id[i] = -1; }
    for (int64_t i = 0; i < pixel_size * SR; i++) coorz[i] = -1;

    /* ---- get_occ_vox, P:263-311: the first point of a voxel marks the query_size box around it and the columns' depth ranges ---- */
    for (int i = 0; i < N; i++) {
        const float* p = xyz + 3 * i;
        int c0 = (int)floorf((p[0] - shift[0]) / vsize[0]);
        if (c0 < 0 || c0 >= X) continue;
        int c1 = (int)floorf((p[1] - shift[1]) / vsize[1]);
        if (c1 < 0 || c1 >= Y) continue;
        float z = p[2];
        if (inverse > 0) z = 1.0f / z;
        int c2 = (int)floorf((z - shift[2]) / vsize[2]);
        if (c2 < 0 || c2 >= Z) continue;
        int64_t ci = ((int64_t)c0 * Y + c1) * Z + c2;
        if (loc[ci] < 0) continue;
        loc[ci] = -1;                                       /* atomicAdd(-1) from 0: only the first visitor goes on */
        for (int x = imax(0, c0 - query_size[0] / 2); x < imin(X, c0 + (query_size[0] + 1) / 2); x++)
            for (int y = imax(0, c1 - query_size[1] / 2); y < imin(Y, c1 + (query_size[1] + 1) / 2); y++)
                for (int zz = imax(0, c2 - query_size[2] / 2); zz < imin(Z, c2 + (query_size[2] + 1) / 2); zz++) {
                    int64_t col = (int64_t)x * Y + y, cj = col * Z + zz;
                    if (coor_occ[cj]) continue;
                    coor_occ[cj] = 1;
                    if (zz < near_id[col]) near_id[col] = zz;
                    if (zz > far_id[col]) far_id[col] = zz;
                }
    }
    /* ---- near_vox_full, P:313-365: per ray its column's mask; the first ray of a column lists its first SR occupied depths and selects
     * (loc = 1) the point voxels in the query_size box around each of them (the launch at P:656-676 hands query_size_gpu to the
     * kernel's `kernel_size` parameter) ---- */
    for (int r = 0; r < R; r++) {
        int vx = pixel_idx[2 * r] / vscale[0], vy = pixel_idx[2 * r + 1] / vscale[1];
        int64_t col = (int64_t)vx * Y + vy;
        int nid = near_id[col], fid = far_id[col];
        ray_mask[r] = fid > 0 ? 1 : 0;
        if (pixel_map[col]) continue;
        pixel_map[col] = 1;
        int counter = 0;
        for (int d = nid; d <= fid; d++) {
            if (!coor_occ[col * Z + d]) continue;
            coorz[col * SR + counter] = (short)d;
            for (int x = imax(0, vx - query_size[0] / 2); x < imin(X, vx + (query_size[0] + 1) / 2); x++)
                for (int y = imax(0, vy - query_size[1] / 2); y < imin(Y, vy + (query_size[1] + 1) / 2); y++)
                    for (int zz = imax(0, d - query_size[2] / 2); zz < imin(Z, d + (query_size[2] + 1) / 2); zz++) {
                        int64_t cj = ((int64_t)x * Y + y) * Z + zz;
                        if (loc[cj] < 0) loc[cj] = 1;
                    }
            if (counter >= SR - 1) break;
            counter++;
        }
    }
    /* ---- P:695-696: rank of the selected voxels inside their column, -1 elsewhere ---- */
    for (int64_t col = 0; col < pixel_size; col++) {
        int c = 0;
        for (int zz = 0; zz < Z; zz++) {
            int sel = loc[col * Z + zz] > 0;
            c += sel;
            rank[col * Z + zz] = sel * c - 1;
        }
        if (c > max_o) return -2;                            /* the reference would index into the next column here */
    }
    /* ---- insert_vox_points, P:368-408: (int) TRUNCATION of the coordinates (not floor); lists in point order, P-cap reservoir ---- */
    for (int i = 0; i < N; i++) {
        const float* p = xyz + 3 * i;
        int cx = (int)((p[0] - shift[0]) / vsize[0]);
        int cy = (int)((p[1] - shift[1]) / vsize[1]);
        float z = p[2];
        if (inverse > 0) z = 1.0f / z;
        int cz = (int)((z - shift[2]) / vsize[2]);
        if (cx < 0 || cx >= X || cy < 0 || cy >= Y || cz < 0 || cz >= Z) continue;
        int64_t col = (int64_t)cx * Y + cy;
        int rk = rank[col * Z + cz];
        if (rk < 0) continue;
        int64_t v = col * max_o + rk;
        int n = (int)cnt[v]++;
        if (n < P) lists[v * P + n] = i;
        else {
            float u = orc_curand_uniform((uint64_t)i + seconds_insert);
            int j = (int)(ceilf(u * (float)(n + 1)) - 1.0f);
            if (j < P) lists[v * P + j] = i;
        }
    }
    /* ---- query_neigh_along_ray_layered (NN > 0, P:493-590) / query_rand_along_ray (P:411-490), for the rays with ray_mask > 0 ---- */
    float* buf = (float*)malloc(sizeof(float) * (K > 0 ? K : 1));
    int cr = -1;                                            /* index among the compacted rays: the kernels' `index` feeds the seed */
    for (int r = 0; r < R; r++) {
        if (!ray_mask[r]) continue;
        cr++;
        int px = pixel_idx[2 * r], py = pixel_idx[2 * r + 1];
        int fx = px / vscale[0], fy = py / vscale[1];
        int64_t col = (int64_t)fx * Y + fy;
        for (int s = 0; s < SR; s++) {
            int fz = (int)coorz[col * SR + s];
            /* P:457-459 / :537-539: float + int*float + (int % int + 0.5 [double]) * float, evaluated as the C expression is */
            /* (nvcc contracts shift + frust * vsize into one fp32 FMA; the rest is promoted to double by the 0.5) */
            float cxf = (float)((double)fmaf((float)fx, vsize[0], shift[0]) + (px % vscale[0] + 0.5) * (double)ray_vsize[0]);
            float cyf = (float)((double)fmaf((float)fy, vsize[1], shift[1]) + (py % vscale[1] + 0.5) * (double)ray_vsize[1]);
            float czf = (float)((double)shift[2] + (fz + 0.5) * (double)vsize[2]);
            if (inverse > 0) czf = 1.0f / czf;
            float* sl = sample_loc + ((int64_t)r * SR + s) * 3;
            sl[0] = cxf; sl[1] = cyf; sl[2] = czf;
            if (fz < 0) continue;
            int32_t* out = sample_pidx + ((int64_t)r * SR + s) * K;
            int kid = 0, far_ind = 0;
            float far2 = 0.0f;
            const int64_t index = (int64_t)cr * SR + s;
            if (NN > 0) {
                for (int layer = 0; layer < (kernel_size[0] + 1) / 2; layer++) {
                    int zlayer = imin((kernel_size[2] + 1) / 2 - 1, layer);
                    for (int x = imax(-fx, -layer); x < imin(X - fx, layer + 1); x++)
                        for (int y = imax(-fy, -layer); y < imin(Y - fy, layer + 1); y++) {
                            int64_t pcol = (int64_t)(fx + x) * Y + (fy + y);
                            for (int zz = imax(-fz, -zlayer); zz < imin(Z - fz, zlayer + 1); zz++) {
                                if (imax(abs(x), abs(y)) != layer && ((zlayer == layer) ? (abs(zz) != zlayer) : 1)) continue;
                                int rk = rank[pcol * Z + fz + zz];
                                if (rk < 0) continue;
                                int64_t v = pcol * max_o + rk;
                                for (int g = 0; g < imin(P, (int)cnt[v]); g++) {
                                    int pi = lists[v * P + g];
                                    const float* q = xyz + 3 * pi;
                                    float xv = (NN < 2) ? (q[0] - cxf) : fmaf(q[0], q[2], -(cxf * czf));   /* nvcc: FMUL + FFMA (oracle/_ref SASS) */
                                    float yv = (NN < 2) ? (q[1] - cyf) : fmaf(q[1], q[2], -(cyf * czf));
                                    float xy2 = fmaf(xv, xv, yv * yv);
                                    float zd = q[2] - czf;
                                    float z2 = zd * zd;
                                    float xyz2 = xy2 + z2;
                                    if ((radius2 == 0 || xy2 <= radius2) && (depth2 == 0 || z2 <= depth2)) {
                                        if (kid++ < K) {
                                            out[kid - 1] = pi; buf[kid - 1] = xyz2;
                                            if (xyz2 > far2) { far2 = xyz2; far_ind = kid - 1; }
                                        } else if (xyz2 < far2) {
                                            out[far_ind] = pi; buf[far_ind] = xyz2; far2 = xyz2;
                                            for (int i = 0; i < K; i++) if (buf[i] > far2) { far2 = buf[i]; far_ind = i; }
                                        }
                                    }
                                }
                            }
                        }
                }
            } else {
                for (int x = imax(0, fx - kernel_size[0] / 2); x < imin(X, fx + (kernel_size[0] + 1) / 2); x++)
                    for (int y = imax(0, fy - kernel_size[1] / 2); y < imin(Y, fy + (kernel_size[1] + 1) / 2); y++) {
                        int64_t pcol = (int64_t)x * Y + y;
                        for (int zz = imax(0, fz - kernel_size[2] / 2); zz < imin(Z, fz + (kernel_size[2] + 1) / 2); zz++) {
                            int rk = rank[pcol * Z + zz];
                            if (rk < 0) continue;
                            int64_t v = pcol * max_o + rk;
                            for (int g = 0; g < imin(P, (int)cnt[v]); g++) {
                                int pi = lists[v * P + g];
                                const float* q = xyz + 3 * pi;
                                float dx = q[0] - cxf, dy = q[1] - cyf, dz = q[2] - czf;
                                if ((radius2 == 0 || fmaf(dx, dx, dy * dy) <= radius2) && (depth2 == 0 || dz * dz <= depth2)) {
                                    if (kid++ < K) out[kid - 1] = pi;
                                    else {
                                        float u = orc_curand_uniform((uint64_t)index + seconds_query);
                                        int j = (int)(ceilf(u * (float)kid) - 1.0f);
                                        if (j < K) out[j] = pi;
                                    }
                                }
                            }
                        }
                    }
            }
        }
    }
    free(buf); free(coor_occ); free(loc); free(near_id); free(far_id); free(coorz); free(pixel_map); free(rank); free(cnt); free(lists);
    return 0;
}
