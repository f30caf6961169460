"""oracle/query_ref.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's world-coordinate neural-point query
(`lighting_fast_querier`, file Q = models/neural_points/query_point_indices_worldcoords.py in the
reference tree).  The kernels live in oracle/query_ref.c (sequential thread-index order); this file
restates the host side: grid hyper-parameters (Q:66-92), ray positions
(models/rendering/diff_ray_marching.py:349-393), the torch glue of query_grid_point_index
(Q:782-954) and the tail of query_points (Q:114-132).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product package (sgnerf_b200/) never does.

Parity status: the reference ships no golden vectors for this path and its kernels cannot run in
the CPU container (PyCUDA).  The restatement is pinned on the GPU box against the reference's own
kernels compiled from /root/reference into oracle/_ref/ (see oracle/build_ref.py and
tests/test_query_vs_reference_kernels.py), and the vectors produced there are committed under
tests/golden/.
"""
import ctypes
import os
import subprocess
from types import SimpleNamespace

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "liborc_query.so")
_lib = None


def build(force=False):
    """gcc the C restatement.  -ffp-contract=off: FMAs are explicit in the source."""
    srcs = [os.path.join(_HERE, "query_ref.c"), os.path.join(_HERE, "query_pers_ref.c")]
    os.makedirs(_BUILD, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO] + srcs + ["-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_curand_uniform.restype = ctypes.c_float
        _lib.orc_curand_uniform.argtypes = [ctypes.c_uint64]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def default_opt(**kw):
    """Canonical hyper-parameters (pointnerf/run/checkpoints/scannet/scene0710_00640480/opt.txt)."""
    o = SimpleNamespace(
        vsize=[0.008, 0.008, 0.008], vscale=[2, 2, 2], kernel_size=[3, 3, 3], query_size=[3, 3, 3],
        ranges=[-10.0, -10.0, -10.0, 10.0, 10.0, 10.0], radius_limit_scale=4.0, depth_limit_scale=0.0,
        max_o=610000, P=26, SR=24, K=8, NN=2, z_depth_dim=400, inverse=0, is_train=0,
        semantic_guidance=0, split="train")
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def get_hyperparameters(opt, xyz_w):
    """Q:66-92.  xyz_w: torch f32 [1,N,3].  numpy/torch promotions are restated literally."""
    vsize_np = opt.vsize
    min_xyz, max_xyz = torch.min(xyz_w, dim=-2)[0][0], torch.max(xyz_w, dim=-2)[0][0]
    vscale_np = np.array(opt.vscale, dtype=np.int32)
    scaled_vsize_np = (vsize_np * vscale_np).astype(np.float32)            # f64 product -> f32
    ranges = opt.ranges
    if ranges is not None:
        min_xyz = torch.max(torch.stack([min_xyz, torch.as_tensor(ranges[:3], dtype=torch.float32)], dim=0), dim=0)[0]
        max_xyz = torch.min(torch.stack([max_xyz, torch.as_tensor(ranges[3:], dtype=torch.float32)], dim=0), dim=0)[0]
    pad = torch.as_tensor(scaled_vsize_np * opt.kernel_size / 2, dtype=torch.float32)   # f32*int list -> f64 -> f32
    min_xyz = min_xyz - pad
    max_xyz = max_xyz + pad
    ranges_np = torch.cat([min_xyz, max_xyz], dim=-1).numpy().astype(np.float32)
    vdim_np = (max_xyz - min_xyz).numpy() / vsize_np                        # f32 / list -> f64
    scaled_vdim_np = np.ceil(vdim_np / vscale_np).astype(np.int32)
    radius_limit_np = np.asarray(opt.radius_limit_scale * max(vsize_np[0], vsize_np[1])).astype(np.float32)
    depth_limit_np = np.asarray(opt.depth_limit_scale * vsize_np[2]).astype(np.float32)
    return SimpleNamespace(radius_limit=radius_limit_np, depth_limit=depth_limit_np, ranges=ranges_np,
                           vsize=vsize_np, scaled_vsize=scaled_vsize_np, scaled_vdim=scaled_vdim_np,
                           vscale=vscale_np, radius2=np.float32(radius_limit_np ** 2))


def near_far_linear_ray_generation(campos, raydir, point_count, near, far, jitter=0.0, rand=None):
    """models/rendering/diff_ray_marching.py:349-393.  rand: optional U[0,1) tensor [B,R,D] standing in
    for torch.rand (only its distribution matters to the reference).  Returns raypos, middle_point_ts."""
    tvals = torch.linspace(0, 1, point_count + 1).view(1, -1)
    tvals = near * (1 - tvals) + far * tvals
    if rand is None:
        rand = torch.rand((raydir.shape[0], raydir.shape[1], point_count))
    segment_length = (tvals[..., 1:] - tvals[..., :-1]) * (1 + jitter * (rand - 0.5))
    end_point_ts = torch.cumsum(segment_length, dim=2)
    end_point_ts = torch.cat([torch.zeros((end_point_ts.shape[0], end_point_ts.shape[1], 1)), end_point_ts], dim=2)
    end_point_ts = near + end_point_ts
    middle_point_ts = (end_point_ts[:, :, :-1] + end_point_ts[:, :, 1:]) / 2
    raypos = campos[:, None, None, :] + raydir[:, :, None, :] * middle_point_ts[:, :, :, None]
    return raypos, middle_point_ts


def raypos_from_t(campos, raydir, t):
    """raypos = campos + raydir * t with separate fp32 multiply and add (diff_ray_marching.py:387).
    t: [D] or [R,D] torch f32."""
    if t.dim() == 1:
        t = t[None, :].expand(raydir.shape[1], -1)
    return campos[:, None, None, :] + raydir[:, :, None, :] * t[None, :, :, None]


def w2pers(point_xyz_w, camrotc2w, campos):
    """Q:125-132 (querier copy; the NeuralPoints copy at neural_points.py:838-850 is the same math)."""
    xyz_w_shift = point_xyz_w - campos[:, None, :]
    xyz_c = torch.sum(xyz_w_shift[..., None, :] * torch.transpose(camrotc2w, 1, 2)[:, None, None, ...], dim=-1)
    z_pers = xyz_c[..., 2]
    x_pers = xyz_c[..., 0] / xyz_c[..., 2]
    y_pers = xyz_c[..., 1] / xyz_c[..., 2]
    return torch.stack([x_pers, y_pers, z_pers], dim=-1)


def build_occ_vox(opt, hp, xyz_np, actual_n=None, seconds_claim=0, seconds_fill=0):
    """Q:706-778 (build_occ_vox as called from Q:797: query_size sits in the kernel_size slot)."""
    N = xyz_np.shape[0]
    dim = np.ascontiguousarray(hp.scaled_vdim, dtype=np.int32)
    vol = int(dim[0]) * int(dim[1]) * int(dim[2])
    g = SimpleNamespace(
        coor_occ=np.empty(vol, np.int32), coor_2_occ=np.empty(vol, np.int32),
        occ_2_coor=np.empty((opt.max_o, 3), np.int32), occ_idx=np.zeros(1, np.int32),
        occ_numpnts=np.empty(opt.max_o, np.int32), occ_2_pnts=np.empty((opt.max_o, opt.P), np.int32),
        dim=dim, shift=np.ascontiguousarray(hp.ranges[:3], dtype=np.float32),
        vsize=np.ascontiguousarray(hp.scaled_vsize, dtype=np.float32))
    qs = np.asarray(opt.query_size, dtype=np.int32)
    lib().orc_build_occ_vox(_p(xyz_np), ctypes.c_int(N), ctypes.c_int(N if actual_n is None else actual_n),
                            _p(g.shift), _p(g.vsize), _p(dim), _p(qs), ctypes.c_int(opt.max_o), ctypes.c_int(opt.P),
                            ctypes.c_uint64(int(seconds_claim)), ctypes.c_uint64(int(seconds_fill)),
                            _p(g.coor_occ), _p(g.coor_2_occ), _p(g.occ_2_coor), _p(g.occ_idx),
                            _p(g.occ_numpnts), _p(g.occ_2_pnts))
    return g


def query_grid_point_index(opt, hp, raypos, xyz_w, raylabel=None, points_label=None, points_label_prob=None,
                           seconds=(0, 0, 0), grid=None):
    """Q:782-954.  raypos torch f32 [1,R,D,3]; xyz_w torch f32 [1,N,3].
    seconds = (claim, fill, query) stand in for the three time.time() reads (Q:715, Q:751, Q:881)."""
    B, R, D = raypos.shape[0], raypos.shape[1], raypos.shape[2]
    assert B == 1
    SR, K = opt.SR, opt.K
    xyz_np = np.ascontiguousarray(xyz_w[0].numpy(), dtype=np.float32)
    g = grid if grid is not None else build_occ_vox(opt, hp, xyz_np, None, seconds[0], seconds[1])
    L = lib()
    raypos_np = np.ascontiguousarray(raypos[0].numpy(), dtype=np.float32)
    raypos_mask = np.zeros((R, D), np.int32)
    L.orc_mask_raypos(_p(raypos_np), _p(g.coor_occ), ctypes.c_int64(R * D), _p(g.shift), _p(g.dim), _p(g.vsize),
                      _p(raypos_mask))
    ray_mask = raypos_mask.max(axis=-1) > 0 if D > 0 else np.zeros(R, bool)        # Q:833
    R1 = int(ray_mask.sum())
    sample_loc = np.zeros((R1, SR, 3), np.float32)
    sample_pidx = np.full((R1, SR, K), -1, np.int32)
    info = SimpleNamespace(grid=g, R1=R1, ray_mask1=ray_mask.copy())
    if R1 > 0:
        raypos_sel = np.ascontiguousarray(raypos_np[ray_mask])                      # Q:838
        mask_sel = raypos_mask[ray_mask]
        cum = np.cumsum(mask_sel, axis=-1).astype(np.int32)                         # Q:843
        slot = np.ascontiguousarray(mask_sel * cum * (cum <= SR) - 1, dtype=np.int32)   # Q:844
        sample_loc_mask = np.zeros((R1, SR), np.int32)
        sample_label = np.zeros((R1, SR), np.int32)
        raylabel_sel = None
        sem = opt.semantic_guidance == 1
        if sem:
            # Q:110 repeats the per-ray label over D; Q:840 selects the hit rays
            raylabel_sel = np.ascontiguousarray(
                np.repeat(raylabel.reshape(R, 1).astype(np.int32), D, axis=1)[ray_mask])
        L.orc_get_shadingloc(_p(raypos_sel), _p(raylabel_sel), _p(slot), ctypes.c_int(R1), ctypes.c_int(D),
                             ctypes.c_int(SR), _p(sample_loc), _p(sample_label if sem else None), _p(sample_loc_mask))
        ks = np.asarray(opt.kernel_size, dtype=np.int32)
        lab = prob = None
        if sem:
            lab = np.ascontiguousarray(points_label.reshape(-1).astype(np.int32))           # Q:915
            prob = np.ascontiguousarray(points_label_prob.reshape(-1, 20).astype(np.int32))  # Q:916 (.to(int32))
        L.orc_query_neigh(_p(xyz_np), _p(lab), _p(prob), ctypes.c_int(R1), ctypes.c_int(SR), ctypes.c_int(opt.P),
                          ctypes.c_int(K), ctypes.c_float(float(hp.radius2)), _p(g.shift), _p(g.dim), _p(g.vsize),
                          _p(ks), _p(g.occ_numpnts), _p(g.occ_2_pnts), _p(g.coor_2_occ), _p(sample_loc),
                          _p(sample_loc_mask), _p(sample_label), _p(sample_pidx), ctypes.c_uint64(int(seconds[2])))
        info.sample_loc_mask1 = sample_loc_mask
        info.sample_pidx1 = sample_pidx.copy()
        info.sample_loc1 = sample_loc.copy()
        masked_valid_ray = (sample_pidx.reshape(R1, -1) >= 0).sum(axis=-1) > 0      # Q:944
        ray_mask[ray_mask] = masked_valid_ray                                        # Q:948 masked_scatter_
        sample_pidx = sample_pidx[masked_valid_ray]
        sample_loc = sample_loc[masked_valid_ray]
    return (torch.from_numpy(sample_pidx)[None], torch.from_numpy(sample_loc)[None],
            torch.from_numpy(ray_mask.astype(np.int8))[None], info)


def query_points(opt, xyz_w, near, far, ray_dirs, cam_pos, cam_rot, t=None, jitter_rand=None,
                 ray_label=None, points_label=None, points_label_prob=None, seconds=(0, 0, 0)):
    """Q:95-122.  Returns what the reference returns plus an `info` namespace of intermediates.
    `t` (middle_point_ts, [D] or [R,D]) may be supplied so that a device path and this oracle
    consume bit-identical depths; otherwise it is computed as the reference does."""
    hp = get_hyperparameters(opt, xyz_w)
    if t is None:
        raypos, mid = near_far_linear_ray_generation(cam_pos, ray_dirs, opt.z_depth_dim, near, far,
                                                     jitter=0.3 if opt.is_train > 0 else 0.0, rand=jitter_rand)
    else:
        raypos, mid = raypos_from_t(cam_pos, ray_dirs, t), t
    pidx, loc_w, ray_mask, info = query_grid_point_index(
        opt, hp, raypos, xyz_w, raylabel=ray_label, points_label=points_label,
        points_label_prob=points_label_prob, seconds=seconds)
    sel = ray_mask[0] > 0
    sample_ray_dirs = ray_dirs[:, sel, :][..., None, :].expand(-1, -1, opt.SR, -1).contiguous()   # Q:114
    info.hp, info.t = hp, mid
    return pidx, w2pers(loc_w, cam_rot, cam_pos), loc_w, sample_ray_dirs, ray_mask, hp.vsize, hp.ranges, info
