// grid.cuh -- the opaque handle behind SgnGrid* (include/sgnerf_b200.h): device buffers of one occupancy grid.
// Built by grid.cu (sgn_grid_build), read by query.cu (sgn_query).
#pragma once
#include "common.cuh"

struct SgnGrid {
    SgnGridCfg cfg;
    int64_t N, vol;
    int32_t* cell_slot;
    uint32_t* occ_bits;
    int32_t* slot_coor;
    int32_t* slot_count;
    int32_t* slot_start;
    float4* cand;
    int32_t* counters;
    uint32_t* coarse_bits;   // 1 bit per 8^3 voxels: some voxel of the brick, or one next to it, is set in occ_bits (march_kernel's skip test)
    // compact voxel -> candidate-list index of the K-NN kernel (replaces the dense cell_slot volume on the query path):
    // knn_brick[b] = (64-bit mask of the 4^3 voxels of brick b that own a non-empty candidate list, rank of the brick's first such voxel);
    // knn_list[rank] = (first candidate, number of candidates) in `cand`.  16 bytes per 64 voxels: L2-resident (8.8 MB at C1, 73 MB at C4).
    uint4* knn_brick;
    int2* knn_list;
    int nbx, nby, nbz;
    // neighbour lists of the K-NN kernel, one per voxel that can hold a shading sample (= set in occ_bits): the voxels of its 3^3 block
    // that own candidates, in the reference's visiting order (centre, then x outer / y / z inner), as
    // (first candidate : 24 | count : 7 | "first entry of shell 1" : 1).  occ_rank[w] = number of set occ_bits below word w, so
    // list(voxel c) = nbr_ent[nbr_off[r] .. nbr_off[r + 1]) with r = occ_rank[c >> 5] + popc(occ_bits[c >> 5] & below(c & 31)).
    // Built only when candidate indices fit 24 bits and P < 128 (nbr_ok); the kernel falls back to the brick index otherwise.
    int32_t* occ_rank;
    int32_t* nbr_off;
    uint32_t* nbr_ent;
    int nbr_ok;
};

namespace sgn {
__host__ __device__ inline int brick4(int d) { return (d + 3) >> 2; }
}
