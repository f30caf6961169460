"""TEST INFRASTRUCTURE: drive the REFERENCE's own CUDA kernels (compiled from /root/reference into
oracle/_ref/libref_query_K<K>.so by oracle/build_ref.py) on torch CUDA tensors, in the order and with the
torch glue of the reference's build_occ_vox / query_grid_point_index
(models/neural_points/query_point_indices_worldcoords.py:706-778, :782-954).
Used to pin the oracle on the GPU box and as the "reference query on one B200" timing in bench.py."""
import ctypes as C
import os

import numpy as np
import torch

from oracle import build_ref

_libs = {}


def available(K=8):
    return os.path.exists(build_ref.so_path(K))


def lib(K=8):
    if K not in _libs:
        _libs[K] = C.CDLL(build_ref.so_path(K))
    return _libs[K]


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"reference kernel {what}: cuda error {rc}")


def build_occ_vox(L, xyz, opt, hp, seconds=(0, 0)):
    """:706-778.  xyz [1,N,3] cuda."""
    dev = xyz.device
    B, N = 1, xyz.shape[1]
    dim = [int(v) for v in hp.scaled_vdim]
    vol = dim[0] * dim[1] * dim[2]
    i32 = dict(dtype=torch.int32, device=dev)
    coor_occ = torch.zeros([B] + dim, **i32)
    occ_2_pnts = torch.full([B, opt.max_o, opt.P], -1, **i32)
    occ_2_coor = torch.full([B, opt.max_o, 3], -1, **i32)
    occ_numpnts = torch.zeros([B, opt.max_o], **i32)
    coor_2_occ = torch.full([B] + dim, -1, **i32)
    occ_idx = torch.zeros([B], **i32)
    actual = torch.full([B], N, **i32)
    shift = torch.from_numpy(np.ascontiguousarray(hp.ranges[:3])).to(dev)
    vsize = torch.from_numpy(np.ascontiguousarray(hp.scaled_vsize)).to(dev)
    dim_t = torch.tensor(dim, **i32)
    qs = torch.tensor(list(opt.query_size), **i32)
    _chk(L.ref_claim_occ(_p(xyz), _p(actual), B, N, _p(shift), _p(vsize), _p(dim_t), vol, opt.max_o, _p(occ_idx), _p(coor_2_occ),
                         _p(occ_2_coor), C.c_ulong(int(seconds[0])), _st()), "claim_occ")
    coor_2_occ = torch.full([B] + dim, -1, **i32)                                       # :735
    _chk(L.ref_map_coor2occ(B, _p(dim_t), _p(qs), vol, opt.max_o, _p(occ_idx), _p(coor_occ), _p(coor_2_occ), _p(occ_2_coor), _st()),
         "map_coor2occ")
    _chk(L.ref_fill_occ2pnts(_p(xyz), _p(actual), B, N, opt.P, _p(shift), _p(vsize), _p(dim_t), vol, opt.max_o, _p(coor_2_occ),
                             _p(occ_2_pnts), _p(occ_numpnts), C.c_ulong(int(seconds[1])), _st()), "fill_occ2pnts")
    return dict(coor_occ=coor_occ, occ_2_coor=occ_2_coor, coor_2_occ=coor_2_occ, occ_idx=occ_idx, occ_numpnts=occ_numpnts,
                occ_2_pnts=occ_2_pnts, shift=shift, vsize=vsize, dim_t=dim_t, vol=vol)


def query_grid_point_index(L, raypos, xyz, opt, hp, seconds=(0, 0, 0), grid=None, raylabel=None, points_label=None,
                           points_label_prob=None):
    """:782-954.  raypos [1,R,D,3] cuda, xyz [1,N,3] cuda.  With raylabel (int32 [1,R], one label per ray, repeated over D as :110 does),
    points_label (int32 [N]) and points_label_prob (int32 [N,20]: what `.to(torch.int32)` of :916 produced) the semantic-guidance
    kernels run (get_shadingloc_with_semantic + query_neigh_along_ray_layered_semantic_guidance, :853-870 / :910-938).
    Returns sample_pidx [1,R'',SR,K], sample_loc [1,R'',SR,3], ray_mask int8 [1,R], grid dict."""
    dev = xyz.device
    B, R, D = 1, raypos.shape[1], raypos.shape[2]
    SR, K = opt.SR, opt.K
    g = grid if grid is not None else build_occ_vox(L, xyz, opt, hp, seconds[:2])
    i32 = dict(dtype=torch.int32, device=dev)
    raypos = raypos.contiguous()
    raypos_mask = torch.zeros([B, R, D], **i32)
    _chk(L.ref_mask_raypos(_p(raypos), _p(g["coor_occ"]), B, R, D, g["vol"], _p(g["shift"]), _p(g["dim_t"]), _p(g["vsize"]),
                           _p(raypos_mask), _st()), "mask_raypos")
    ray_mask = torch.max(raypos_mask, dim=-1)[0] > 0
    R1 = int(torch.max(torch.sum(ray_mask.to(torch.int32))).cpu().numpy())
    sample_loc = torch.zeros([B, R1, SR, 3], dtype=torch.float32, device=dev)
    sample_pidx = torch.full([B, R1, SR, K], -1, **i32)
    if R1 > 0:
        raypos = torch.masked_select(raypos, ray_mask[..., None, None].expand(-1, -1, D, 3)).reshape(B, R1, D, 3)
        raypos_mask = torch.masked_select(raypos_mask, ray_mask[..., None].expand(-1, -1, D)).reshape(B, R1, D)
        sem = raylabel is not None
        if sem:
            raylabel_d = raylabel.reshape(B, R, 1, 1).to(torch.int32).expand(-1, -1, D, 1).contiguous()                    # :110
            raylabel_d = torch.masked_select(raylabel_d, ray_mask[..., None, None].expand(-1, -1, D, 1)).reshape(B, R1, D, 1)  # :840
        cum = torch.cumsum(raypos_mask, dim=-1).to(torch.int32)
        raypos_mask = (raypos_mask * cum * (cum <= SR)) - 1
        sample_loc_mask = torch.zeros([B, R1, SR], **i32)
        ks = torch.tensor(list(opt.kernel_size), **i32)
        if sem:
            sample_label = torch.zeros([B, R1, SR], **i32)
            _chk(L.ref_get_shadingloc_with_semantic(_p(raypos), _p(raylabel_d.contiguous()), _p(raypos_mask.contiguous()), B, R1, D, SR,
                                                    _p(sample_loc), _p(sample_label), _p(sample_loc_mask), _st()), "get_shadingloc_with_semantic")
            lab = points_label.reshape(-1).to(torch.int32).contiguous()
            prob = points_label_prob.reshape(-1, 20).to(torch.int32).contiguous()
            _chk(L.ref_query_neigh_along_ray_layered_semantic_guidance(
                _p(xyz), _p(lab), _p(prob), B, SR, R1, opt.max_o, opt.P, K, g["vol"], C.c_float(float(hp.radius2)), _p(g["shift"]),
                _p(g["dim_t"]), _p(g["vsize"]), _p(ks), _p(g["occ_numpnts"]), _p(g["occ_2_pnts"]), _p(g["coor_2_occ"]), _p(sample_loc),
                _p(sample_loc_mask), _p(sample_label), _p(sample_pidx), C.c_ulong(int(seconds[2])), opt.NN, _st()),
                "query_neigh_along_ray_layered_semantic_guidance")
        else:
            _chk(L.ref_get_shadingloc(_p(raypos), _p(raypos_mask.contiguous()), B, R1, D, SR, _p(sample_loc), _p(sample_loc_mask), _st()),
                 "get_shadingloc")
            _chk(L.ref_query_neigh_along_ray_layered(
                _p(xyz), B, SR, R1, opt.max_o, opt.P, K, g["vol"], C.c_float(float(hp.radius2)), _p(g["shift"]), _p(g["dim_t"]),
                _p(g["vsize"]), _p(ks), _p(g["occ_numpnts"]), _p(g["occ_2_pnts"]), _p(g["coor_2_occ"]), _p(sample_loc),
                _p(sample_loc_mask), _p(sample_pidx), C.c_ulong(int(seconds[2])), opt.NN, _st()), "query_neigh_along_ray_layered")
        valid_ray = torch.sum(sample_pidx.view(B, R1, -1) >= 0, dim=-1) > 0
        R2 = int(torch.max(torch.sum(valid_ray.to(torch.int32), dim=-1)).cpu().numpy())
        ray_mask.masked_scatter_(ray_mask, valid_ray)
        sample_pidx = torch.masked_select(sample_pidx, valid_ray[..., None, None].expand(-1, -1, SR, K)).reshape(B, R2, SR, K)
        sample_loc = torch.masked_select(sample_loc, valid_ray[..., None, None].expand(-1, -1, SR, 3)).reshape(B, R2, SR, 3)
    return sample_pidx, sample_loc, ray_mask.to(torch.int8), g
