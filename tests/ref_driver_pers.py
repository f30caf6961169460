"""TEST INFRASTRUCTURE: drive the REFERENCE's own PERSPECTIVE query kernels (compiled unchanged into oracle/_ref/libref_query_pers_K<K>.so by
oracle/build_ref.py) on torch CUDA tensors, in the order and with the torch glue of lighting_fast_querier.query_grid_point_index
(models/neural_points/query_point_indices.py:617-782; line numbers below refer to that file)."""
import ctypes as C
import os

import numpy as np
import torch

from oracle import build_ref

_libs = {}


def available(K=8):
    return os.path.exists(build_ref.so_path(K, pers=True))


def lib(K=8):
    if K not in _libs:
        _libs[K] = C.CDLL(build_ref.so_path(K, pers=True))
    return _libs[K]


def _p(t):
    return C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"reference kernel {what}: cuda error {rc}")


def query_grid_point_index(L, pixel_idx, xyz_pers, opt, hp, seconds=(0, 0), max_o=None):
    """pixel_idx int32 [1,R,2] cuda, xyz_pers f32 [1,N,3] cuda.  Returns the reference's four outputs (compacted to the masked-in rays)."""
    dev = xyz_pers.device
    xyz_pers = xyz_pers.reshape(1, -1, 3)
    B, N = 1, xyz_pers.shape[1]
    dim = [int(v) for v in hp.scaled_vdim]
    pixel_size, vol = dim[0] * dim[1], dim[0] * dim[1] * dim[2]
    SR, K, P = opt.SR, opt.K, opt.P
    f32 = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device=dev)
    i32 = lambda a: torch.tensor(np.asarray(a, dtype=np.int32), device=dev)
    shift, vsize, dim_t, vscale = f32(hp.ranges[:3]), f32(hp.scaled_vsize), i32(hp.scaled_vdim), i32(hp.vscale)
    ks, qs, ray_vsize = i32(opt.kernel_size), i32(opt.query_size), f32(hp.ray_vsize)
    xyz = xyz_pers.contiguous()
    actual = torch.full([B], N, dtype=torch.int32, device=dev)
    pix = pixel_idx.reshape(B, -1, 2).to(torch.int32).clone()
    R = pix.shape[1]
    coor_occ = torch.zeros([B] + dim, dtype=torch.uint8, device=dev)                                   # :627-630
    counter = torch.zeros([B] + dim, dtype=torch.int8, device=dev)
    near_id = torch.full([B, dim[0], dim[1]], dim[2], dtype=torch.int32, device=dev)
    far_id = torch.full([B, dim[0], dim[1]], -1, dtype=torch.int32, device=dev)
    _chk(L.refp_get_occ_vox(_p(xyz), _p(actual), B, N, _p(shift), _p(vsize), _p(dim_t), _p(qs), pixel_size, vol, _p(coor_occ), _p(counter),
                            _p(near_id), _p(far_id), int(opt.inverse), _st()), "get_occ_vox")
    coorz = torch.full([B, dim[0], dim[1], SR], -1, dtype=torch.int16, device=dev)                      # :653-655
    pixel_map = torch.zeros([B, dim[0], dim[1]], dtype=torch.uint8, device=dev)
    ray_mask = torch.zeros([B, R], dtype=torch.int8, device=dev)
    _chk(L.refp_near_vox_full(B, SR, _p(pix), R, _p(vscale), _p(dim_t), pixel_size, vol, _p(qs), _p(pixel_map), _p(ray_mask), _p(coor_occ),
                              _p(counter), _p(near_id), _p(far_id), _p(coorz), _st()), "near_vox_full")
    occ_per_column = coor_occ.sum(-1, dtype=torch.int32)
    pix = torch.masked_select(pix, (ray_mask > 0)[..., None].expand(-1, -1, 2)).reshape(1, -1, 2)      # :688
    R1 = int(torch.max(torch.sum(ray_mask, dim=-1)).cpu().numpy())
    sel_count = (counter > 0).sum(-1, dtype=torch.int32)                                                # int8 cumsum of :696 overflows past 127
    counter = (counter > 0).to(torch.int8)
    counter = counter * torch.cumsum(counter, dtype=torch.int8, dim=-1) - 1                             # :695-696
    if max_o is None:
        max_o = int(torch.max(counter).cpu().numpy().astype(np.int32)) + 1                              # :698-699
    pnt_counter = torch.zeros([B, dim[0], dim[1], max_o], dtype=torch.int16, device=dev)
    pntidx = torch.zeros([B, dim[0], dim[1], max_o, P], dtype=torch.int32, device=dev)
    _chk(L.refp_insert_vox_points(_p(xyz), _p(actual), B, N, P, max_o, pixel_size, vol, _p(shift), _p(dim_t), _p(vsize), _p(counter),
                                  _p(pnt_counter), _p(pntidx), C.c_ulong(int(seconds[0])), int(opt.inverse), _st()), "insert_vox_points")
    sample_pidx = torch.full([B, R1, SR, K], -1, dtype=torch.int32, device=dev)
    sample_loc = torch.full([B, R1, SR, 3], 0.0, dtype=torch.float32, device=dev)
    if R1 > 0:
        fn = L.refp_query_neigh_along_ray_layered if opt.NN > 0 else L.refp_query_rand_along_ray
        _chk(fn(_p(xyz), B, SR, R1, max_o, P, K, pixel_size, vol, C.c_float(float(hp.radius2)), C.c_float(float(hp.depth2)), _p(shift), _p(dim_t),
                _p(vsize), _p(ray_vsize), _p(vscale), _p(ks), _p(pix), _p(counter), _p(coorz), _p(pnt_counter), _p(pntidx), _p(sample_pidx),
                _p(sample_loc), C.c_ulong(int(seconds[1])), int(opt.NN), int(opt.inverse), _st()), "query_along_ray")
    torch.cuda.synchronize()
    info = dict(max_o=max_o, max_selected_per_column=int(sel_count.max()), max_points_per_voxel=int(pnt_counter.max()),
                max_occupied_per_column=int(occ_per_column.max()))
    return sample_pidx, sample_loc, pix, ray_mask, info
