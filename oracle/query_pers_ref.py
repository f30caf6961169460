"""TEST INFRASTRUCTURE (oracle): the reference's PERSPECTIVE-frustum querier (--wcoord_query 0), host side.

File P = models/neural_points/query_point_indices.py of the reference: get_hyperparameters (P:48-73), query_grid_point_index (P:600-782; the
kernels are restated sequentially in oracle/query_pers_ref.c), pers2w (P:95-107), query_points (P:76-93).  Only tests/ may import this.
"""
import ctypes
from types import SimpleNamespace

import numpy as np
import torch

from . import query_ref as qr


def default_opt(**kw):
    o = SimpleNamespace(vscale=[2, 2, 2], kernel_size=[3, 3, 3], query_size=[3, 3, 3], radius_limit_scale=4.0, depth_limit_scale=1.3,
                        max_o=64, P=16, SR=24, K=8, NN=2, z_depth_dim=400, inverse=0, is_train=0, shpnt_jitter="passfunc")
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def get_hyperparameters(opt, h, w, intrinsic, near_depth, far_depth):
    """P:48-73."""
    x_rl, x_rh = -intrinsic[0, 2] / intrinsic[0, 0], (w - intrinsic[0, 2]) / intrinsic[0, 0]
    y_rl, y_rh = -intrinsic[1, 2] / intrinsic[1, 1], (h - intrinsic[1, 2]) / intrinsic[1, 1]
    z_r = (far_depth - near_depth) if opt.inverse == 0 else (1.0 / near_depth - 1.0 / far_depth)
    ranges = np.array([x_rl, y_rl, near_depth, x_rh, y_rh, far_depth], dtype=np.float32) if opt.inverse == 0 else \
        np.array([x_rl, y_rl, 1.0 / far_depth, x_rh, y_rh, 1.0 / near_depth], dtype=np.float32)
    vdim = np.array([w, h, opt.z_depth_dim], dtype=np.int32)
    vsize = np.array([(x_rh - x_rl) / vdim[0], (y_rh - y_rl) / vdim[1], z_r / vdim[2]], dtype=np.float32)
    vscale = np.array(opt.vscale, dtype=np.int32)
    scaled_vdim = np.ceil(vdim / vscale).astype(np.int32)
    scaled_vsize = (vsize * vscale).astype(np.float32)
    radius_limit, depth_limit = opt.radius_limit_scale * max(vsize[0], vsize[1]), opt.depth_limit_scale * vsize[2]
    ray_vsize = (scaled_vsize / vscale).astype(np.float32)                                # P:711
    return SimpleNamespace(radius_limit=np.float32(radius_limit), depth_limit=np.float32(depth_limit), ranges=ranges, vsize=vsize, vdim=vdim,
                           scaled_vsize=scaled_vsize, scaled_vdim=scaled_vdim, vscale=vscale, ray_vsize=ray_vsize,
                           radius2=np.float32(np.float32(radius_limit) ** 2), depth2=np.float32(np.float32(depth_limit) ** 2))


def pers2w(point_xyz_pers, camrotc2w, campos):
    """P:95-107."""
    x_pers = point_xyz_pers[..., 0] * point_xyz_pers[..., 2]
    y_pers = point_xyz_pers[..., 1] * point_xyz_pers[..., 2]
    z_pers = point_xyz_pers[..., 2]
    xyz_c = torch.stack([x_pers, y_pers, z_pers], dim=-1)
    xyz_w_shift = torch.sum(xyz_c[..., None, :] * camrotc2w, dim=-1)
    ray_dirs = xyz_w_shift / (torch.linalg.norm(xyz_w_shift, dim=-1, keepdims=True) + 1e-7)
    return xyz_w_shift + campos[:, None, :], ray_dirs


def query_uncompacted(opt, hp, pixel_idx, xyz_pers, seconds=(0, 0)):
    """The kernels on all R rays, rows per input ray.  pixel_idx int [R,2], xyz_pers f32 [N,3] (numpy or torch)."""
    L = qr.lib()
    xyz = np.ascontiguousarray(np.asarray(xyz_pers, dtype=np.float32).reshape(-1, 3))
    pix = np.ascontiguousarray(np.asarray(pixel_idx, dtype=np.int32).reshape(-1, 2))
    R, N = pix.shape[0], xyz.shape[0]
    ray_mask = np.zeros(R, np.int8)
    pidx = np.full((R, opt.SR, opt.K), -1, np.int32)
    loc = np.zeros((R, opt.SR, 3), np.float32)
    a = lambda v, t: np.ascontiguousarray(np.asarray(v, dtype=t))
    p = lambda v: v.ctypes.data_as(ctypes.c_void_p)
    shift, vs, dim, vsc = a(hp.ranges[:3], np.float32), a(hp.scaled_vsize, np.float32), a(hp.scaled_vdim, np.int32), a(hp.vscale, np.int32)
    ks, qs, rv = a(opt.kernel_size, np.int32), a(opt.query_size, np.int32), a(hp.ray_vsize, np.float32)
    L.orc_pers_query.restype = ctypes.c_int
    rc = L.orc_pers_query(p(xyz), ctypes.c_int(N), p(pix), ctypes.c_int(R), p(shift), p(vs), p(dim), p(vsc), p(ks), p(qs), p(rv), ctypes.c_int(opt.SR),
                          ctypes.c_int(opt.K), ctypes.c_int(opt.P), ctypes.c_int(opt.max_o), ctypes.c_float(float(hp.radius2)), ctypes.c_float(float(hp.depth2)),
                          ctypes.c_int(opt.NN), ctypes.c_int(opt.inverse), ctypes.c_uint64(int(seconds[0])), ctypes.c_uint64(int(seconds[1])),
                          p(ray_mask), p(pidx), p(loc))
    if rc != 0:
        raise RuntimeError(f"orc_pers_query: code {rc} (-2: a pixel column selects more voxels than max_o)")
    return pidx, loc, ray_mask


def query_points(opt, pixel_idx, xyz_pers, h, w, intrinsic, near, far, cam_pos, cam_rot, seconds=(0, 0)):
    """P:76-93 for is_train = 0 (no sample jitter): compacted outputs as the reference returns them."""
    hp = get_hyperparameters(opt, h, w, intrinsic, near, far)
    pidx, loc, mask = query_uncompacted(opt, hp, pixel_idx, xyz_pers, seconds)
    sel = mask > 0
    sample_pidx = torch.from_numpy(pidx[sel])[None]
    sample_loc = torch.from_numpy(loc[sel])[None]
    B, Rr = 1, sample_loc.shape[1]
    loc_w, dirs = pers2w(sample_loc.reshape(1, -1, 3), cam_rot, cam_pos)
    return (sample_pidx, sample_loc, loc_w.reshape(B, Rr, opt.SR, 3), dirs.reshape(B, Rr, opt.SR, 3), torch.from_numpy(mask)[None], hp.vsize, hp.ranges, hp)
