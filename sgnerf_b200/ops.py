"""Tensor-level host API over the C ABI (include/sgnerf_b200.h).

PyTorch is plumbing here: device memory, the current stream and autograd bookkeeping.  All arithmetic of
the hot path happens in libsgnerf_b200.so; there is no eager/PyTorch fallback -- CPU tensors are rejected.
"""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from ._lib import SgnAggCfg, SgnGridCfg, SgnPointGrads, SgnPointTables

PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISION_TF32 = 2     # layer-wise path with TF32 tensor-core GEMMs (forward + backward): the training precision


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, dtype, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"sgnerf_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _workspace(nbytes, device):
    return torch.empty((max(int(nbytes), 256) + 255) // 256 * 64, dtype=torch.float32, device=device)


# ---------------------------------------------------------------------------------------------------------
# occupancy grid
# ---------------------------------------------------------------------------------------------------------
def grid_hyperparameters(xyz, vsize, vscale, kernel_size, ranges, radius_limit_scale, alive=None):
    """Host-side grid parameters exactly as lighting_fast_querier.get_hyperparameters derives them
    (reference models/neural_points/query_point_indices_worldcoords.py:66-92): fp32 min/max of the cloud,
    clamp to `ranges`, pad by scaled_vsize*kernel/2 (a float64 product rounded to fp32), ceil of a float64
    quotient for the dims, radius from the UNscaled voxel size.  Two small D2H reads; call once per cloud."""
    vsize64 = np.asarray(vsize, dtype=np.float64)
    vscale_i = np.asarray(vscale, dtype=np.int32)
    scaled_vsize = (vsize64 * vscale_i).astype(np.float32)
    pts = xyz.reshape(-1, 3)
    if alive is None:
        mn, mx = pts.min(dim=0)[0].float(), pts.max(dim=0)[0].float()
    else:        # rows of pruned points (holes kept for index stability, RenderScene.edit) do not count
        a = alive.reshape(-1, 1)
        mn = torch.where(a, pts, torch.full_like(pts, float("inf"))).min(dim=0)[0].float()
        mx = torch.where(a, pts, torch.full_like(pts, float("-inf"))).max(dim=0)[0].float()
    if ranges is not None:
        r = torch.as_tensor(np.asarray(ranges, dtype=np.float64), dtype=torch.float32, device=xyz.device)
        mn = torch.maximum(mn, r[:3])
        mx = torch.minimum(mx, r[3:])
    pad64 = scaled_vsize.astype(np.float64) * np.asarray(kernel_size, dtype=np.int64) / 2
    pad = torch.as_tensor(pad64, dtype=torch.float32, device=xyz.device)
    mn = mn - pad
    mx = mx + pad
    ranges_np = torch.cat([mn, mx]).cpu().numpy().astype(np.float32)
    vdim = (mx - mn).cpu().numpy().astype(np.float64) / vsize64
    scaled_vdim = np.ceil(vdim / vscale_i).astype(np.int32)
    radius = np.float32(radius_limit_scale * max(vsize[0], vsize[1]))
    return SimpleNamespace(ranges=ranges_np, scaled_vsize=scaled_vsize, scaled_vdim=scaled_vdim,
                           radius=radius, radius2=np.float32(radius * radius), vsize=list(vsize))


class OccGrid:
    """Device-resident occupancy grid (sgn_grid_build).  Build once per point-cloud version."""

    def __init__(self, xyz, origin, scaled_vsize, dim, query_size, P, max_o, seconds_claim=0, seconds_fill=0,
                 actual_n=None, neighbour_lists=True):
        self.xyz = _dev(xyz.reshape(-1, 3), torch.float32, "xyz")
        N = self.xyz.shape[0]
        cfg = SgnGridCfg()
        for i in range(3):
            cfg.origin[i] = float(origin[i]); cfg.vsize[i] = float(scaled_vsize[i])
            cfg.dim[i] = int(dim[i]); cfg.query_size[i] = int(query_size[i])
        cfg.P, cfg.max_o = int(P), int(max_o)
        cfg.seconds_claim, cfg.seconds_fill = int(seconds_claim), int(seconds_fill)
        self.cfg = cfg
        pb, sb = C.c_size_t(), C.c_size_t()
        _lib.call("sgn_grid_workspace_bytes", N, C.byref(cfg), C.byref(pb), C.byref(sb))
        self._persistent = _workspace(pb.value, self.xyz.device)
        scratch = _workspace(sb.value, self.xyz.device)
        handle = C.c_void_p()
        _lib.call("sgn_grid_build_flags", _ptr(self.xyz), N, N if actual_n is None else int(actual_n), C.byref(cfg),
                  _ptr(self._persistent), self._persistent.numel() * 4, _ptr(scratch), scratch.numel() * 4,
                  0 if neighbour_lists else 1, C.byref(handle), _stream())
        self._handle = handle
        self._scratch = scratch  # stream-ordered: keep alive until the build kernels have run
        self.N, self.P, self.max_o = N, int(P), int(max_o)
        self.dim = [int(d) for d in dim]

    def buffer(self, which, dtype=torch.int32):
        """Zero-copy view of an internal buffer (tests/tooling); see sgn_grid_buffer."""
        p, n = C.c_void_p(), C.c_int64()
        _lib.call("sgn_grid_buffer", self._handle, which, C.byref(p), C.byref(n))
        base = self._persistent.data_ptr()
        off = (p.value - base) // 4
        width = 4 if which == 5 else 1
        flat = self._persistent[off:off + n.value * width]
        if which == 5:
            return flat.view(n.value, 4)
        return flat.view(dtype)

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            _lib.load().sgn_grid_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def query(grid, campos, raydir, t, SR, K, kernel_size0, radius2, ray_label=None, pt_label=None,
          pt_label_prob_bits=None, seconds_query=0, sparse_rows=False):
    """sgn_query.  campos [3], raydir [R,3], t [D] or [R,D] (middle_point_ts).  Uncompacted outputs:
    sample_pidx int32 [R,SR,K], sample_loc_w f32 [R,SR,3], sample_mask int32 [R,SR], ray_mask int8 [R].
    sparse_rows=True (sgn_query_frame): the sample_pidx rows of slots with sample_mask == 0 are left unwritten -- only for a consumer that
    is given the mask (aggregate(sample_mask=...))."""
    raydir = _dev(raydir.reshape(-1, 3), torch.float32, "raydir")
    campos = _dev(campos.reshape(3), torch.float32, "campos")
    t = _dev(t, torch.float32, "t")
    R = raydir.shape[0]
    per_ray = 1 if t.dim() == 2 else 0
    D = t.shape[-1]
    if per_ray and t.shape[0] != R:
        raise ValueError("t must be [D] or [R,D]")
    dev = raydir.device
    pidx = torch.empty(R, SR, K, dtype=torch.int32, device=dev)
    loc_w = torch.empty(R, SR, 3, dtype=torch.float32, device=dev)
    smask = torch.empty(R, SR, dtype=torch.int32, device=dev)
    rmask = torch.empty(R, dtype=torch.int8, device=dev)
    slabel = None
    if ray_label is not None:
        ray_label = _dev(ray_label.reshape(-1), torch.int32, "ray_label")
        pt_label = _dev(pt_label.reshape(-1), torch.int32, "pt_label")
        pt_label_prob_bits = _dev(pt_label_prob_bits.reshape(-1, 20), torch.int32, "pt_label_prob_bits")
        slabel = torch.empty(R, SR, dtype=torch.int32, device=dev)
    _lib.call("sgn_query_frame" if sparse_rows else "sgn_query", grid._handle, _ptr(campos), _ptr(raydir), _ptr(t), per_ray, R, D, SR, K, int(kernel_size0),
              float(radius2), _ptr(ray_label), _ptr(pt_label), _ptr(pt_label_prob_bits), int(seconds_query),
              _ptr(pidx), _ptr(loc_w), _ptr(smask), _ptr(slabel), _ptr(rmask), _stream())
    return pidx, loc_w, smask, rmask


def gather_rows(table, pidx):
    """table [N,C] f32, pidx int32 [...] -> [..., C] with clamp(pidx, 0) (neural_points.py:956-972)."""
    table = _dev(table, torch.float32, "table")
    pidx = _dev(pidx, torch.int32, "pidx")
    Cc = table.shape[-1]
    out = torch.empty(pidx.shape + (Cc,), dtype=torch.float32, device=table.device)
    _lib.call("sgn_gather_rows", _ptr(table), Cc, _ptr(pidx), pidx.numel(), _ptr(out), _stream())
    return out


# ---------------------------------------------------------------------------------------------------------
# voxel-grid helpers
# ---------------------------------------------------------------------------------------------------------
def voxel_downsample(xyz_val, vox_res):
    """construct_vox_points_closest (models/mvs/mvs_utils.py:536-561, the call of run/train_ft.py:141 / :715): the host glue (bounds,
    voxel size) with the reference's torch ops, the grouping / centroids / closest points by sgn_voxel_downsample.
    Returns (xyz_centroid [V,3], sparse_grid_idx int32 [V,3], min_idx int64 [V])."""
    xyz = _dev(xyz_val.reshape(-1, 3), torch.float32, "xyz")
    N = xyz.shape[0]
    xyz_min, xyz_max = torch.min(xyz, dim=-2)[0], torch.max(xyz, dim=-2)[0]
    space_edge = torch.max(xyz_max - xyz_min) * 1.05
    space_min = (xyz_max + xyz_min) / 2 - space_edge / 2
    vox_sz = space_edge / vox_res
    mn = (C.c_float * 3)(*[float(v) for v in space_min.cpu()])
    sz = (C.c_float * 3)(*([float(vox_sz.cpu())] * 3))
    nbytes = C.c_size_t()
    _lib.call("sgn_voxel_downsample_bytes", N, C.byref(nbytes))
    ws = _workspace(nbytes.value, xyz.device)
    centroid = torch.empty(N, 3, dtype=torch.float32, device=xyz.device)
    grid_idx = torch.empty(N, 3, dtype=torch.int32, device=xyz.device)
    min_idx = torch.empty(N, dtype=torch.int64, device=xyz.device)
    count = torch.zeros(1, dtype=torch.int32, device=xyz.device)
    _lib.call("sgn_voxel_downsample", _ptr(xyz), N, mn, sz, int(vox_res), _ptr(ws), ws.numel() * 4, _ptr(centroid), _ptr(grid_idx), _ptr(min_idx),
              _ptr(count), _stream())
    V = int(count.item())
    return centroid[:V], grid_idx[:V], min_idx[:V]


def query_vox_grid(sample_loc_w, full_grid_idx, space_min, grid_vox_sz, grid_res):
    """NeuralPoints.query_vox_grid (neural_points.py:814-826): sample_loc_w [...,3] -> int64 [...,8] corner indices (sgn_query_vox_grid)."""
    loc = _dev(sample_loc_w, torch.float32, "sample_loc_w")
    grid = _dev(full_grid_idx, torch.int32, "full_grid_idx")
    out = torch.empty(loc.shape[:-1] + (8,), dtype=torch.int64, device=loc.device)
    mn = (C.c_float * 3)(*[float(v) for v in torch.as_tensor(space_min).cpu().reshape(3)])
    _lib.call("sgn_query_vox_grid", _ptr(loc), loc.numel() // 3, _ptr(grid), int(grid_res), mn, float(grid_vox_sz), _ptr(out), _stream())
    return out


def pers_hyperparameters(h, w, intrinsic, near_depth, far_depth, z_depth_dim, vscale, radius_limit_scale, depth_limit_scale, inverse=0):
    """lighting_fast_querier.get_hyperparameters of the perspective querier (models/neural_points/query_point_indices.py:48-73): host numpy
    arithmetic in the reference's order and dtypes (float32 arrays built from python / float64 scalars)."""
    from types import SimpleNamespace
    intrinsic = np.asarray(intrinsic)
    x_rl, x_rh = -intrinsic[0, 2] / intrinsic[0, 0], (w - intrinsic[0, 2]) / intrinsic[0, 0]
    y_rl, y_rh = -intrinsic[1, 2] / intrinsic[1, 1], (h - intrinsic[1, 2]) / intrinsic[1, 1]
    z_r = (far_depth - near_depth) if inverse == 0 else (1.0 / near_depth - 1.0 / far_depth)
    if inverse == 0:
        ranges = np.array([x_rl, y_rl, near_depth, x_rh, y_rh, far_depth], dtype=np.float32)
    else:
        ranges = np.array([x_rl, y_rl, 1.0 / far_depth, x_rh, y_rh, 1.0 / near_depth], dtype=np.float32)
    vdim = np.array([w, h, z_depth_dim], dtype=np.int32)
    vsize = np.array([(x_rh - x_rl) / vdim[0], (y_rh - y_rl) / vdim[1], z_r / vdim[2]], dtype=np.float32)
    vscale = np.array(vscale, dtype=np.int32)
    scaled_vdim = np.ceil(vdim / vscale).astype(np.int32)
    scaled_vsize = (vsize * vscale).astype(np.float32)
    radius_limit = np.float32(radius_limit_scale * max(vsize[0], vsize[1]))
    depth_limit = np.float32(depth_limit_scale * vsize[2])
    return SimpleNamespace(radius_limit=radius_limit, depth_limit=depth_limit, ranges=ranges, vsize=vsize, vdim=vdim, scaled_vsize=scaled_vsize,
                           scaled_vdim=scaled_vdim, vscale=vscale, ray_vsize=(scaled_vsize / vscale).astype(np.float32),          # :711
                           radius2=np.float32(radius_limit ** 2), depth2=np.float32(depth_limit ** 2))                            # :749-750


def pers_query(xyz_pers, pixel_idx, hp, kernel_size, query_size, SR, K, P, NN=2, inverse=0, seconds=(0, 0)):
    """query_grid_point_index of the perspective querier (query_point_indices.py:600-782) for all rays of one camera, rows per input ray.
    xyz_pers f32 [N,3] perspective coordinates, pixel_idx int [R,2].  Returns sample_pidx int32 [R,SR,K], sample_loc f32 [R,SR,3] (perspective),
    ray_mask int8 [R]."""
    xyz = _dev(xyz_pers.reshape(-1, 3), torch.float32, "xyz_pers")
    pix = _dev(pixel_idx.reshape(-1, 2).to(torch.int32), torch.int32, "pixel_idx")
    N, R = xyz.shape[0], pix.shape[0]
    cfg = _lib.SgnPersCfg()
    for name, src in (("shift", hp.ranges[:3]), ("vsize", hp.scaled_vsize), ("ray_vsize", hp.ray_vsize)):
        for i in range(3):
            getattr(cfg, name)[i] = float(src[i])
    for name, src in (("dim", hp.scaled_vdim), ("vscale", hp.vscale), ("kernel_size", kernel_size), ("query_size", query_size)):
        for i in range(3):
            getattr(cfg, name)[i] = int(src[i])
    cfg.P, cfg.SR, cfg.K, cfg.NN, cfg.inverse = int(P), int(SR), int(K), int(NN), int(inverse)
    cfg.radius2, cfg.depth2 = float(hp.radius2), float(hp.depth2)
    cfg.seconds_insert, cfg.seconds_query = int(seconds[0]), int(seconds[1])
    nbytes = C.c_size_t()
    _lib.call("sgn_pers_query_bytes", N, R, C.byref(cfg), C.byref(nbytes))
    ws = _workspace(nbytes.value, xyz.device)
    pidx = torch.empty((R, SR, K), dtype=torch.int32, device=xyz.device)
    loc = torch.empty((R, SR, 3), dtype=torch.float32, device=xyz.device)
    mask = torch.empty((R,), dtype=torch.int8, device=xyz.device)
    _lib.call("sgn_pers_query", _ptr(xyz), N, _ptr(pix), R, C.byref(cfg), _ptr(ws), nbytes.value, _ptr(pidx), _ptr(loc), _ptr(mask), _stream())
    return pidx, loc, mask


# ---------------------------------------------------------------------------------------------------------
# compositing
# ---------------------------------------------------------------------------------------------------------
def ray_dist(loc_pers, ray_valid, vsize_z, raydist_mode_unit=1):
    loc_pers = _dev(loc_pers, torch.float32, "loc_pers")
    valid = _dev(ray_valid, torch.uint8, "ray_valid")
    R, SR = loc_pers.shape[-3], loc_pers.shape[-2]
    out = torch.empty(loc_pers.shape[:-1], dtype=torch.float32, device=loc_pers.device)
    _lib.call("sgn_ray_dist", _ptr(loc_pers), _ptr(valid), float(vsize_z), int(raydist_mode_unit), R, SR, _ptr(out), _stream())
    return out


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, decoded, ray_dist_, valid_u8, bg, blend):
        R, SR = decoded.shape[-3], decoded.shape[-2]
        dev = decoded.device
        lead = decoded.shape[:-2]
        ray_color = torch.empty(lead + (3,), dtype=torch.float32, device=dev)
        opacity = torch.empty(lead + (SR,), dtype=torch.float32, device=dev)
        acc = torch.empty_like(opacity)
        bw = torch.empty_like(opacity)
        bgt = torch.empty(lead, dtype=torch.float32, device=dev)
        n_rays = int(np.prod(lead)) if len(lead) else 1
        _lib.call("sgn_composite_forward", _ptr(decoded), _ptr(ray_dist_), _ptr(valid_u8), _ptr(bg), blend, n_rays, SR,
                  _ptr(ray_color), _ptr(opacity), _ptr(acc), _ptr(bw), _ptr(bgt), _stream())
        ctx.save_for_backward(decoded, ray_dist_, valid_u8, bg)
        ctx.blend, ctx.n_rays, ctx.SR = blend, n_rays, SR
        ctx.mark_non_differentiable(acc)
        return ray_color, opacity, acc, bw, bgt

    @staticmethod
    def backward(ctx, g_color, g_opacity, g_acc, g_bw, g_bgt):
        decoded, ray_dist_, valid_u8, bg = ctx.saved_tensors
        c = lambda g: None if g is None else g.contiguous().float()
        g_color, g_opacity, g_bw, g_bgt = c(g_color), c(g_opacity), c(g_bw), c(g_bgt)
        d_dec = torch.empty_like(decoded)
        _lib.call("sgn_composite_backward", _ptr(decoded), _ptr(ray_dist_), _ptr(valid_u8), _ptr(bg), ctx.blend, ctx.n_rays,
                  ctx.SR, _ptr(g_color), _ptr(g_opacity), _ptr(g_bw), _ptr(g_bgt), _ptr(d_dec), _stream())
        return d_dec, None, None, None, None


def composite(decoded, ray_dist_, ray_valid, bg_color=None, blend=0):
    """Alpha compositing (differentiable w.r.t. `decoded`).  decoded [...,R,SR,4], ray_dist/ray_valid [...,R,SR].
    Returns ray_color [...,R,3], opacity, acc_transmission, blend_weight [...,R,SR], bg_transmission [...,R]."""
    decoded = _dev(decoded, torch.float32, "decoded")
    ray_dist_ = _dev(ray_dist_, torch.float32, "ray_dist")
    valid = _dev(ray_valid, torch.uint8, "ray_valid")
    bg = _dev(bg_color.reshape(3), torch.float32, "bg_color") if bg_color is not None else None
    if decoded.shape[-1] != 4:
        raise ValueError("decoded must have 4 channels (sigma, r, g, b)")
    return _Composite.apply(decoded, ray_dist_, valid, bg, int(blend))


def render_composite(decoded, loc_pers, ray_valid, ray_mask, vsize_z, bg_color, blend=0, raydist_mode_unit=1, depth_array=False):
    """Inference tail of a frame in one kernel (no autograd): ray_dist -> composite -> fill_invalid.  loc_pers is [R,SR,3], or with
    depth_array=True the dense camera-depth array [R,SR] of aggregate(depth_only=True) (sgn_render_composite_depth).
    Returns ray_color [R,3], opacity [R,SR], bg_transmission [R] -- the values the three separate calls give -- and depth [R]
    (`coarse_depth`: opacity * transmittance weighted camera depth of the samples, 0 for rays that missed)."""
    decoded = _dev(decoded.detach(), torch.float32, "decoded")
    loc_pers = _dev(loc_pers, torch.float32, "loc_pers")
    valid = _dev(ray_valid, torch.uint8, "ray_valid")
    R, SR = decoded.shape[-3], decoded.shape[-2]
    dev = decoded.device
    ray_color = torch.empty(R, 3, dtype=torch.float32, device=dev)
    opacity = torch.empty(R, SR, dtype=torch.float32, device=dev)
    bgt = torch.empty(R, dtype=torch.float32, device=dev)
    depth = torch.empty(R, dtype=torch.float32, device=dev)
    _lib.call("sgn_render_composite_depth" if depth_array else "sgn_render_composite", _ptr(decoded), _ptr(loc_pers), _ptr(valid),
              _ptr(_dev(ray_mask, torch.int8, "ray_mask")), float(vsize_z),
              int(raydist_mode_unit), _ptr(_dev(bg_color.reshape(3), torch.float32, "bg")), int(blend), R, SR, _ptr(ray_color), _ptr(opacity),
              _ptr(bgt), _ptr(depth), _stream())
    return ray_color, opacity, bgt, depth


def probe_outputs(opacity, sample_loc_w, sample_pidx, weight, conf_coef, ray_mask, xyz, embedding, color, dirs, conf):
    """The reference's `prob == 1` outputs (sgn_probe_outputs), rows per input ray.  Returns a dict with the reference's key names:
    ray_max_shading_opacity [R,1], ray_max_sample_loc_w [R,3], ray_max_far_dist [R,1], shading_avg_color [R,3], shading_avg_dir [R,3],
    shading_avg_conf [R,1], shading_avg_embedding [R,C]."""
    f32 = torch.float32
    opacity = _dev(opacity.detach(), f32, "opacity")
    R, SR = opacity.shape[-2], opacity.shape[-1]
    pidx = _dev(sample_pidx, torch.int32, "sample_pidx")
    K = pidx.shape[-1]
    xyz = _dev(xyz.reshape(-1, 3), f32, "xyz")
    N = xyz.shape[0]
    embedding = _dev(embedding.detach().reshape(N, -1), f32, "embedding")
    C_ = embedding.shape[1]
    tb = _tables(xyz, embedding, _dev(color.detach().reshape(N, 3), f32, "color"), _dev(dirs.detach().reshape(N, 3), f32, "dirs"),
                 _dev(conf.detach().reshape(N), f32, "conf") if conf is not None else None, None)
    dev = opacity.device
    o = {"ray_max_shading_opacity": torch.empty(R, 1, device=dev), "ray_max_sample_loc_w": torch.empty(R, 3, device=dev),
         "ray_max_far_dist": torch.empty(R, 1, device=dev), "shading_avg_color": torch.empty(R, 3, device=dev),
         "shading_avg_dir": torch.empty(R, 3, device=dev), "shading_avg_conf": torch.empty(R, 1, device=dev),
         "shading_avg_embedding": torch.empty(R, C_, device=dev)}
    _lib.call("sgn_probe_outputs", _ptr(opacity), _ptr(_dev(sample_loc_w, f32, "sample_loc_w")), _ptr(pidx), _ptr(_dev(weight.detach(), f32, "weight")),
              _ptr(_dev(conf_coef.detach(), f32, "conf_coef")), _ptr(_dev(ray_mask, torch.int8, "ray_mask") if ray_mask is not None else None),
              C.byref(tb), C_, R, SR, K, _ptr(o["ray_max_shading_opacity"]), _ptr(o["ray_max_sample_loc_w"]), _ptr(o["ray_max_far_dist"]),
              _ptr(o["shading_avg_color"]), _ptr(o["shading_avg_dir"]), _ptr(o["shading_avg_conf"]), _ptr(o["shading_avg_embedding"]), _stream())
    return o


def fill_invalid(ray_mask, bg_color, ray_color, opacity=None, bg_transmission=None):
    """In-place fill_invalid for uncompacted rows (models/neural_points_volumetric_model.py:158-195)."""
    R = ray_mask.numel()
    SR = opacity.shape[-1] if opacity is not None else 1
    _lib.call("sgn_fill_invalid", _ptr(_dev(ray_mask, torch.int8, "ray_mask")), _ptr(_dev(bg_color.reshape(3), torch.float32, "bg")),
              R, SR, _ptr(ray_color), _ptr(opacity), _ptr(bg_transmission), _stream())
    return ray_color


# ---------------------------------------------------------------------------------------------------------
# aggregation
# ---------------------------------------------------------------------------------------------------------
def agg_cfg(feat_dim=32, num_feat_freqs=3, dist_xyz_freq=5, num_viewdir_freqs=4, width=256, n_block1=2, n_block2_bpnet=0,
            label_dim=0, n_block3=2, n_color=4, act_super=1, leaky_slope=0.01):
    return SgnAggCfg(feat_dim, num_feat_freqs, dist_xyz_freq, num_viewdir_freqs, width, n_block1, n_block2_bpnet, label_dim,
                     n_block3, n_color, act_super, leaky_slope)


def agg_layer_shapes(cfg):
    n = _lib.load().sgn_agg_num_layers(C.byref(cfg))
    if n < 0:
        _lib.check(n, "sgn_agg_num_layers")
    out = []
    for i in range(n):
        a, b = C.c_int(), C.c_int()
        _lib.call("sgn_agg_layer_shape", C.byref(cfg), i, C.byref(a), C.byref(b))
        out.append((a.value, b.value))
    return out


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def _tables(xyz, embedding, color, dirs, conf, label_emb):
    tb = SgnPointTables()
    tb.xyz, tb.embedding, tb.color, tb.dir = xyz.data_ptr(), embedding.data_ptr(), color.data_ptr(), dirs.data_ptr()
    tb.conf = conf.data_ptr() if conf is not None else None
    tb.label_emb = label_emb.data_ptr() if label_emb is not None else None
    tb.N = xyz.shape[0]
    return tb


# Test hook: run the backward GEMMs in this precision instead of the forward call's (both read the same saved workspace, so a
# TF32 forward followed by an fp32 and a TF32 backward isolates the rounding of the backward GEMMs from LeakyReLU-kink flips).
BACKWARD_PRECISION_OVERRIDE = None


class _Aggregate(torch.autograd.Function):
    """decoded, ray_valid, loc_pers, weight, conf_coef = f(embedding, color, dir, conf, *weights, *biases)."""

    @staticmethod
    def forward(ctx, meta, embedding, color, dirs, conf, *wb):
        nl = len(wb) // 2
        weights, biases = list(wb[:nl]), list(wb[nl:])
        cfg, xyz, label_emb, pidx, loc_w, raydir, campos, camrot, precision, want_aux = (
            meta.cfg, meta.xyz, meta.label_emb, meta.pidx, meta.loc_w, meta.raydir, meta.campos, meta.camrot,
            meta.precision, meta.want_aux)
        R, SR, K = pidx.shape
        dev = pidx.device
        need_grad = meta.grad_enabled and any(ctx.needs_input_grad[1:])    # forward() itself always runs with grad mode off
        if need_grad and precision == PRECISION_BF16:
            raise RuntimeError("sgnerf_b200: training runs the fp32 or tf32 path; the bf16 tensor-core path is forward-only")
        nbytes = C.c_size_t()
        _lib.call("sgn_agg_workspace_bytes", C.byref(cfg), xyz.shape[0], R, SR, K, precision, int(need_grad), C.byref(nbytes))
        ws = _workspace(nbytes.value, dev)
        decoded = torch.empty(R, SR, 4, dtype=torch.float32, device=dev)
        ray_valid = torch.empty(R, SR, dtype=torch.uint8, device=dev)
        # frame mode (inference tail): the samples' camera depth as a dense [R,SR] array instead of the [R,SR,3] perspective positions
        depth_only = bool(getattr(meta, "depth_only", False)) and not need_grad
        loc_pers = None if depth_only else torch.empty(R, SR, 3, dtype=torch.float32, device=dev)
        loc_depth = torch.empty(R, SR, dtype=torch.float32, device=dev) if depth_only else None
        weight = torch.empty(R, SR, K, dtype=torch.float32, device=dev) if want_aux else None
        conf_coef = torch.empty(R, SR, K, dtype=torch.float32, device=dev) if want_aux else None
        tb = _tables(xyz, embedding, color, dirs, conf, label_emb)
        cache = meta.point_cache if (precision == PRECISION_BF16 and not need_grad) else None
        smask = getattr(meta, "sample_mask", None)            # sgn_query's sample_mask: slots without a sample are not read
        if smask is not None:
            smask = _dev(smask, torch.int32, "sample_mask")
        _lib.call("sgn_agg_forward_frame_masked", C.byref(cfg), _ptr_array(weights), _ptr_array(biases), C.byref(tb), _ptr(pidx), _ptr(smask),
                  _ptr(loc_w), _ptr(raydir), _ptr(campos), _ptr(camrot), R, SR, K, precision, int(need_grad), _ptr(decoded), _ptr(ray_valid),
                  _ptr(loc_pers), _ptr(loc_depth), _ptr(weight), _ptr(conf_coef), _ptr(ws), ws.numel() * 4, _ptr(cache), _stream())
        if depth_only:
            loc_pers = loc_depth
        if need_grad:
            ctx.meta, ctx.ws, ctx.nl = meta, ws, nl
            ctx.save_for_backward(embedding, color, dirs, conf, *wb)
        if weight is None:
            weight = torch.empty(0, device=dev)
            conf_coef = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(ray_valid, loc_pers, weight)
        return decoded, ray_valid, loc_pers, weight, conf_coef

    @staticmethod
    def backward(ctx, g_decoded, g_valid, g_loc, g_weight, g_conf):
        meta, nl = ctx.meta, ctx.nl
        embedding, color, dirs, conf, *wb = ctx.saved_tensors
        weights, biases = list(wb[:nl]), list(wb[nl:])
        R, SR, K = meta.pidx.shape
        need = ctx.needs_input_grad
        z = lambda t, flag: torch.zeros_like(t) if (flag and t is not None) else None
        d_emb, d_col, d_dir, d_conf = z(embedding, need[1]), z(color, need[2]), z(dirs, need[3]), z(conf, need[4])
        d_w = [z(w, need[5 + i]) for i, w in enumerate(weights)]
        d_b = [z(b, need[5 + nl + i]) for i, b in enumerate(biases)]
        g_decoded = g_decoded.contiguous().float() if g_decoded is not None else torch.zeros(R, SR, 4, device=meta.pidx.device)
        g_conf_c = None
        if g_conf is not None and meta.want_aux and g_conf.numel() == R * SR * K:
            g_conf_c = g_conf.contiguous().float()
        tb = _tables(meta.xyz, embedding, color, dirs, conf, meta.label_emb)
        gr = SgnPointGrads()
        gr.embedding = d_emb.data_ptr() if d_emb is not None else None
        gr.color = d_col.data_ptr() if d_col is not None else None
        gr.dir = d_dir.data_ptr() if d_dir is not None else None
        gr.conf = d_conf.data_ptr() if d_conf is not None else None
        _lib.call("sgn_agg_backward_prec", C.byref(meta.cfg), _ptr_array(weights), _ptr_array(biases), C.byref(tb), _ptr(meta.pidx),
                  _ptr(meta.loc_w), _ptr(meta.raydir), _ptr(meta.campos), _ptr(meta.camrot), R, SR, K,
                  meta.precision if BACKWARD_PRECISION_OVERRIDE is None else int(BACKWARD_PRECISION_OVERRIDE), _ptr(g_decoded),
                  _ptr(g_conf_c), _ptr_array(d_w), _ptr_array(d_b), C.byref(gr), _ptr(ctx.ws), ctx.ws.numel() * 4, _stream())
        # the saved workspace stays with ctx (a second backward under retain_graph reads it again); autograd frees it with the graph
        return (None, d_emb, d_col, d_dir, d_conf, *d_w, *d_b)


# ---------------------------------------------------------------------------------------------------------
# training step without autograd (sgnerf_b200/train.py): the same entry points, called in order, gradients into caller-owned buffers
# ---------------------------------------------------------------------------------------------------------
def aggregate_train_forward(cfg, weights, biases, xyz, embedding, color, dirs, conf, label_emb, pidx, loc_w, raydir, campos, camrotc2w, precision):
    """sgn_agg_forward with save_for_backward = 1.  All tensors are contiguous fp32 CUDA tensors already (no checks: hot loop).
    Returns (decoded, ray_valid, loc_pers, weight, conf_coef, workspace, tables)."""
    R, SR, K = pidx.shape
    dev = pidx.device
    nbytes = C.c_size_t()
    _lib.call("sgn_agg_workspace_bytes", C.byref(cfg), xyz.shape[0], R, SR, K, precision, 1, C.byref(nbytes))
    ws = _workspace(nbytes.value, dev)
    decoded = torch.empty(R, SR, 4, dtype=torch.float32, device=dev)
    ray_valid = torch.empty(R, SR, dtype=torch.uint8, device=dev)
    loc_pers = torch.empty(R, SR, 3, dtype=torch.float32, device=dev)
    weight = torch.empty(R, SR, K, dtype=torch.float32, device=dev)
    conf_coef = torch.empty(R, SR, K, dtype=torch.float32, device=dev)
    tb = _tables(xyz, embedding, color, dirs, conf, label_emb)
    _lib.call("sgn_agg_forward", C.byref(cfg), _ptr_array(weights), _ptr_array(biases), C.byref(tb), _ptr(pidx), _ptr(loc_w), _ptr(raydir),
              _ptr(campos), _ptr(camrotc2w), R, SR, K, precision, 1, _ptr(decoded), _ptr(ray_valid), _ptr(loc_pers), _ptr(weight), _ptr(conf_coef),
              _ptr(ws), ws.numel() * 4, _stream())
    return decoded, ray_valid, loc_pers, weight, conf_coef, ws, tb


def aggregate_train_backward(cfg, weights, biases, tb, pidx, loc_w, raydir, campos, camrotc2w, precision, d_decoded, d_conf_coef, d_weights, d_biases,
                             d_emb, d_color, d_dir, d_conf, ws):
    """sgn_agg_backward_prec: gradients are ACCUMULATED (+=) into d_weights / d_biases (lists, entries may be None) and the point-table
    accumulators d_emb / d_color / d_dir / d_conf (any may be None)."""
    R, SR, K = pidx.shape
    gr = SgnPointGrads()
    gr.embedding = d_emb.data_ptr() if d_emb is not None else None
    gr.color = d_color.data_ptr() if d_color is not None else None
    gr.dir = d_dir.data_ptr() if d_dir is not None else None
    gr.conf = d_conf.data_ptr() if d_conf is not None else None
    _lib.call("sgn_agg_backward_prec", C.byref(cfg), _ptr_array(weights), _ptr_array(biases), C.byref(tb), _ptr(pidx), _ptr(loc_w), _ptr(raydir),
              _ptr(campos), _ptr(camrotc2w), R, SR, K, precision, _ptr(d_decoded), _ptr(d_conf_coef), _ptr_array(d_weights), _ptr_array(d_biases),
              C.byref(gr), _ptr(ws), ws.numel() * 4, _stream())


def composite_forward_raw(decoded, ray_dist_, valid_u8, bg, blend=0):
    """sgn_composite_forward, ray colour only (no autograd)."""
    R, SR = decoded.shape[0], decoded.shape[1]
    ray_color = torch.empty(R, 3, dtype=torch.float32, device=decoded.device)
    _lib.call("sgn_composite_forward", _ptr(decoded), _ptr(ray_dist_), _ptr(valid_u8), _ptr(bg), blend, R, SR, _ptr(ray_color), None, None, None, None,
              _stream())
    return ray_color


def composite_backward_raw(decoded, ray_dist_, valid_u8, bg, d_ray_color, blend=0):
    R, SR = decoded.shape[0], decoded.shape[1]
    d_dec = torch.empty_like(decoded)
    _lib.call("sgn_composite_backward", _ptr(decoded), _ptr(ray_dist_), _ptr(valid_u8), _ptr(bg), blend, R, SR, _ptr(d_ray_color), None, None, None,
              _ptr(d_dec), _stream())
    return d_dec


def loss_hit_count(ray_mask, count):
    _lib.call("sgn_loss_hit_count", _ptr(ray_mask), ray_mask.numel(), _ptr(count), _stream())


def loss_forward_backward(ray_color, gt, ray_mask, conf_coef, hit_count, loss, color_weight=1.0, conf_weight=1e-4, zero_eps=1e-3, const_term=1e-6):
    """The reference's loss (base_rendering_model.py:543-641) and its gradients in one pass (sgn_loss_forward_backward).
    Returns (d_ray_color [R,3], d_conf_coef [R,SR,K] or None); `loss` (device scalar) is overwritten."""
    R = ray_color.shape[0]
    d_color = torch.empty_like(ray_color)
    SR, K = (conf_coef.shape[1], conf_coef.shape[2]) if conf_coef is not None else (1, 1)
    d_conf = torch.empty_like(conf_coef) if conf_coef is not None else None
    _lib.call("sgn_loss_forward_backward", _ptr(ray_color), _ptr(gt), _ptr(ray_mask), _ptr(conf_coef), R, SR, K, _ptr(hit_count), float(color_weight),
              float(conf_weight), float(zero_eps), float(const_term), _ptr(loss), _ptr(d_color), _ptr(d_conf), _stream())
    return d_color, d_conf


def adam_rows(param, grad, exp_avg, exp_avg_sq, active, step, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, zero_grad=True):
    """sgn_adam_rows on a [N,C] (or [N]) table; `step` is the device step counter (already incremented for this step)."""
    N = param.shape[0]
    Cc = param.numel() // max(N, 1)
    _lib.call("sgn_adam_rows", _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(active), N, Cc, float(lr), float(betas[0]), float(betas[1]),
              float(eps), _ptr(step), float(grad_scale), int(zero_grad), _stream())


def adam_rows_multi(params, grads, exp_avgs, exp_avg_sqs, active, step, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, zero_grad=True):
    """sgn_adam_rows_multi: several [N, C_k] tables that share their rows (the point tables) in one pass, one `active` flag per row."""
    N = params[0].shape[0]
    Cs = (C.c_int32 * len(params))(*[p.numel() // max(N, 1) for p in params])
    _lib.call("sgn_adam_rows_multi", len(params), _ptr_array(params), _ptr_array(grads), _ptr_array(exp_avgs), _ptr_array(exp_avg_sqs), Cs, _ptr(active), N,
              float(lr), float(betas[0]), float(betas[1]), float(eps), _ptr(step), float(grad_scale), int(zero_grad), _stream())


def adam_dense_multi(params, grads, exp_avgs, exp_avg_sqs, step, lr, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, zero_grad=False):
    """sgn_adam_dense_multi: torch.optim.Adam's update on a list of small dense tensors (the MLP) in one launch."""
    n = (C.c_int64 * len(params))(*[p.numel() for p in params])
    _lib.call("sgn_adam_dense_multi", len(params), _ptr_array(params), _ptr_array(grads), _ptr_array(exp_avgs), _ptr_array(exp_avg_sqs), n, float(lr),
              float(betas[0]), float(betas[1]), float(eps), _ptr(step), float(grad_scale), int(zero_grad), _stream())


def adam_mark_rows(rows, touched):
    """sgn_adam_mark_rows: touched[r] = 1 for every r >= 0 of the int32 tensor `rows`."""
    rows = _dev(rows, torch.int32, "rows")
    _lib.call("sgn_adam_mark_rows", _ptr(rows), rows.numel(), _ptr(touched), _stream())


def adam_rows_list(params, grads, exp_avgs, exp_avg_sqs, active, active_list, active_count, touched, step, lr, betas=(0.9, 0.999), eps=1e-8,
                   grad_scale=1.0, zero_grad=True):
    """sgn_adam_rows_list: adam_rows_multi driven by the list of active rows (nothing is read for rows that never received a gradient)."""
    N = params[0].shape[0]
    Cs = (C.c_int32 * len(params))(*[p.numel() // max(N, 1) for p in params])
    _lib.call("sgn_adam_rows_list", len(params), _ptr_array(params), _ptr_array(grads), _ptr_array(exp_avgs), _ptr_array(exp_avg_sqs), Cs, _ptr(active),
              _ptr(active_list), _ptr(active_count), _ptr(touched), N, float(lr), float(betas[0]), float(betas[1]), float(eps), _ptr(step),
              float(grad_scale), int(zero_grad), _stream())


def rows_union(touched, row_list, count):
    """sgn_rows_union: ascending list of the rows with touched != 0 into row_list (int32 [N]), their number into count (int32 [1])."""
    N = touched.numel()
    nbytes = C.c_size_t()
    _lib.call("sgn_rows_union_bytes", N, C.byref(nbytes))
    ws = _workspace(nbytes.value, touched.device)
    _lib.call("sgn_rows_union", _ptr(touched), N, _ptr(row_list), _ptr(count), _ptr(ws), nbytes.value, _stream())


def rows_pack(tables, row_list, count, packed, stride, unpack=False):
    """sgn_rows_pack: rows `row_list[:count]` of the tables [N, C_k] <-> packed [count, stride]."""
    N = tables[0].shape[0]
    Cs = (C.c_int32 * len(tables))(*[t.numel() // max(N, 1) for t in tables])
    _lib.call("sgn_rows_pack", len(tables), _ptr_array(tables), Cs, _ptr(row_list), _ptr(count), N, _ptr(packed), int(stride), int(unpack), _stream())


def adam_step_count(step):
    _lib.call("sgn_adam_step_count", _ptr(step), _stream())


def build_point_cache(cfg, weights, embedding, label_emb=None):
    """Inference cache of the bf16 path (sgn_agg_point_cache_build): the point-only part of block1.0 (and block2_bpnet.0) per point.
    Valid until the embeddings or the aggregator weights change; pass it to aggregate(point_cache=...)."""
    f32 = torch.float32
    embedding = _dev(embedding.reshape(-1, embedding.shape[-1]), f32, "embedding")
    N = embedding.shape[0]
    label_emb = _dev(label_emb.reshape(N, -1), f32, "label_emb") if label_emb is not None else None
    ws_ = [_dev(w, f32, "weight") for w in weights]
    nbytes = C.c_size_t()
    _lib.call("sgn_agg_point_cache_bytes", C.byref(cfg), N, C.byref(nbytes))
    cache = _workspace(nbytes.value, embedding.device)
    tb = SgnPointTables()
    tb.embedding = embedding.data_ptr()
    tb.label_emb = label_emb.data_ptr() if label_emb is not None else None
    tb.N = N
    _lib.call("sgn_agg_point_cache_build", C.byref(cfg), _ptr_array(ws_), C.byref(tb), _ptr(cache), cache.numel() * 4, _stream())
    return cache


def update_point_cache(cfg, cache, embedding, rows, label_emb=None):
    """sgn_agg_point_cache_update: recompute the cache rows of the points listed in `rows` (int32 device tensor) after their embeddings
    changed (point edits); the layer weights are the ones the cache was built with."""
    f32 = torch.float32
    embedding = _dev(embedding.reshape(-1, embedding.shape[-1]), f32, "embedding")
    N = embedding.shape[0]
    tb = SgnPointTables()
    tb.embedding = embedding.data_ptr()
    label_emb = _dev(label_emb.reshape(N, -1), f32, "label_emb") if label_emb is not None else None
    tb.label_emb = label_emb.data_ptr() if label_emb is not None else None
    tb.N = N
    rows = _dev(rows.reshape(-1), torch.int32, "rows")
    _lib.call("sgn_agg_point_cache_update", C.byref(cfg), C.byref(tb), _ptr(cache), cache.numel() * 4, _ptr(rows), rows.numel(), _stream())


def aggregate(cfg, weights, biases, xyz, embedding, color, dirs, conf, label_emb, pidx, loc_w, raydir, campos, camrotc2w,
              precision=PRECISION_FP32, want_aux=True, point_cache=None, depth_only=False, sample_mask=None):
    """Fused gather + aggregation MLPs (sgn_agg_forward / sgn_agg_backward).

    Tables: xyz [N,3], embedding [N,C], color [N,3], dirs [N,3], conf [N] (or None), label_emb [N,E] (or None).
    Query outputs: pidx int32 [R,SR,K], loc_w [R,SR,3]; raydir [R,3]; campos [3]; camrotc2w [3,3].
    Returns decoded [R,SR,4], ray_valid uint8 [R,SR], loc_pers [R,SR,3], weight [R,SR,K], conf_coef [R,SR,K].
    depth_only=True (inference, no autograd): the third result is the samples' camera depth [R,SR] (= loc_pers[..., 2]) for
    render_composite(depth_array=True) instead of the full perspective positions (sgn_agg_forward_frame).
    sample_mask: ops.query's third result; slots it marks empty have all -1 index rows, which are then not read (same results)."""
    f32 = torch.float32
    meta = SimpleNamespace(cfg=cfg, xyz=_dev(xyz.reshape(-1, 3), f32, "xyz"), label_emb=_dev(label_emb, f32, "label_emb"),
                           pidx=_dev(pidx, torch.int32, "pidx"), loc_w=_dev(loc_w, f32, "loc_w"),
                           raydir=_dev(raydir.reshape(-1, 3), f32, "raydir"), campos=_dev(campos.reshape(3), f32, "campos"),
                           camrot=_dev(camrotc2w.reshape(3, 3), f32, "camrotc2w"), precision=int(precision), want_aux=bool(want_aux),
                           grad_enabled=torch.is_grad_enabled(), point_cache=point_cache, depth_only=depth_only, sample_mask=sample_mask)
    N = meta.xyz.shape[0]
    embedding = _dev(embedding.reshape(N, -1), f32, "embedding")
    color = _dev(color.reshape(N, 3), f32, "color")
    dirs = _dev(dirs.reshape(N, 3), f32, "dirs")
    conf = _dev(conf.reshape(N), f32, "conf") if conf is not None else None
    if meta.label_emb is not None:
        meta.label_emb = meta.label_emb.reshape(N, -1)
    shapes = agg_layer_shapes(cfg)
    if len(weights) != len(shapes) or len(biases) != len(shapes):
        raise ValueError(f"expected {len(shapes)} weight/bias tensors, got {len(weights)}/{len(biases)}")
    ws_, bs_ = [], []
    for (cin, cout), w, b in zip(shapes, weights, biases):
        if tuple(w.shape) != (cout, cin) or b.numel() != cout:
            raise ValueError(f"layer shape mismatch: expected weight {(cout, cin)}, got {tuple(w.shape)}")
        ws_.append(_dev(w, f32, "weight")); bs_.append(_dev(b, f32, "bias"))
    return _Aggregate.apply(meta, embedding, color, dirs, conf, *ws_, *bs_)
