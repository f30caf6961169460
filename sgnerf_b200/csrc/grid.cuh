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
};
