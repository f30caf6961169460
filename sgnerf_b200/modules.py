"""Reference-shaped modules over the C ABI: `lighting_fast_querier`, `NeuralPoints`, `PointAggregator`, `ray_march`,
`alpha_ray_march` -- same names, argument lists, return tuples and state_dict keys as the reference classes they replace:

    lighting_fast_querier   models/neural_points/query_point_indices_worldcoords.py:47-132
    NeuralPoints            models/neural_points/neural_points.py:77, :312-420 (ctor), :520-665 (prune/grow/set_*), :942-988 (forward)
    PointAggregator         models/aggregators/point_aggregators.py:12, :256-296 (ctor), :312-418 (viewmlp_init), :868-959 (forward)
    ray_march               models/rendering/diff_ray_marching.py:509-555, alpha_ray_march :558-573

They are drop-ins for models/neural_points_volumetric_model.py.  Only the canonical branch of the
reference is built (which_agg_model=viewmlp, agg_distance_kernel=linear, agg_dist_pers=20, agg_intrp_order=2,
wcoord_query=1, near_far_linear ray generation, radiance render, alpha/alpha2 blend); any other option raises.  All arithmetic
runs in libsgnerf_b200.so -- there is no PyTorch fallback, CPU tensors are rejected.
"""
import math
import time
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from . import ops, pipeline


def _opt(opt, name, default):
    return getattr(opt, name, default)


def _register_flags(parser, table):
    """Add the reference's command-line flags (name, type, default, nargs) to `parser`; a flag another class already added
    (argparse raises on duplicates) is left as it is, as happens when both the reference's and these classes are imported."""
    have = set(parser._option_string_actions)
    for name, typ, default, nargs in table:
        if "--" + name in have:
            continue
        kw = dict(type=typ, default=default, help=f"{name} (same flag as the reference)")
        if nargs is not None:
            kw["nargs"] = nargs
        parser.add_argument("--" + name, **kw)
    return parser


# Flags of NeuralPoints.modify_commandline_options (models/neural_points/neural_points.py:80-309): same names, types, defaults, nargs.
# `wcoord_query` is declared int with the STRING default '0' in the reference (argparse converts string defaults through `type`).
NEURAL_POINTS_FLAGS = (
    ("semantic_guidance", int, 0, None), ("load_points", int, 1, None), ("point_noise", str, "", None), ("num_point", int, 8192, None),
    ("construct_res", int, 0, None), ("grid_res", int, 0, None), ("cloud_path", str, "", None), ("shpnt_jitter", str, "passfunc", None),
    ("point_features_dim", int, 64, None), ("gpu_maxthr", int, 1024, None), ("z_depth_dim", int, 400, None), ("SR", int, 24, None),
    ("K", int, 32, None), ("max_o", int, None, None), ("P", int, 16, None), ("NN", int, 0, None), ("radius_limit_scale", float, 5.0, None),
    ("depth_limit_scale", float, 1.3, None), ("default_conf", float, -1.0, None), ("vscale", int, (2, 2, 1), "+"),
    ("kernel_size", int, (7, 7, 1), "+"), ("query_size", int, (0, 0, 0), "+"), ("xyz_grad", int, 0, None), ("feat_grad", int, 1, None),
    ("conf_grad", int, 1, None), ("color_grad", int, 1, None), ("bp_embedding_grad", int, 0, None), ("dir_grad", int, 0, None),
    ("feedforward", int, 0, None), ("inverse", int, 0, None), ("point_conf_mode", str, "0", None), ("point_color_mode", str, "0", None),
    ("point_dir_mode", str, "0", None), ("vsize", float, (0.005, 0.005, 0.005), "+"), ("wcoord_query", int, "0", None),
    ("ranges", float, (-100.0, -100.0, -100.0, 100.0, 100.0, 100.0), "+"),
)

# Flags of PointAggregator.modify_commandline_options (models/aggregators/point_aggregators.py:15-253).
POINT_AGGREGATOR_FLAGS = (
    ("feature_init_method", str, "rand", None), ("which_agg_model", str, "viewmlp", None), ("agg_distance_kernel", str, "quadric", None),
    ("sh_degree", int, 4, None), ("sh_dist_func", str, "sh_quadric", None), ("sh_act", str, "sigmoid", None),
    ("agg_axis_weight", float, None, "+"), ("agg_dist_pers", int, 1, None), ("apply_pnt_mask", int, 1, None), ("modulator_concat", int, 0, None),
    ("agg_intrp_order", int, 0, None), ("shading_feature_mlp_layer0", int, 0, None), ("shading_feature_mlp_layer1", int, 2, None),
    ("shading_feature_mlp_layer2", int, 0, None), ("shading_feature_mlp_layer2_bpnet", int, 0, None), ("shading_feature_mlp_layer3", int, 0, None),
    ("shading_feature_mlp_layer4", int, 1, None), ("shading_feature_mlp_linear", int, 0, None), ("shading_feature_num", int, 256, None),
    ("point_hyper_dim", int, 256, None), ("shading_alpha_mlp_layer", int, 1, None), ("shading_color_mlp_layer", int, 1, None),
    ("shading_color_channel_num", int, 3, None), ("num_feat_freqs", int, 0, None), ("num_hyperfeat_freqs", int, 0, None),
    ("dist_xyz_freq", int, 2, None), ("dist_xyz_deno", float, 0, None), ("weight_xyz_freq", int, 2, None), ("weight_feat_dim", int, 8, None),
    ("agg_weight_norm", int, 1, None), ("view_ori", int, 0, None), ("agg_feat_xyz_mode", str, "None", None),
    ("agg_alpha_xyz_mode", str, "None", None), ("agg_color_xyz_mode", str, "None", None), ("act_type", str, "ReLU", None),
    ("act_super", int, 1, None), ("predict_semantic", int, 0, None), ("layers_2d", int, 34, None), ("classes", int, 20, None),
    ("arch_3d", str, "MinkUNet18A", None), ("bpnetweight", str, "../bpnetInitmodel/bpnet_5cm.pth.tar", None),
)


# ---------------------------------------------------------------------------------------------------------
# lazy gathers
# ---------------------------------------------------------------------------------------------------------
class GatheredRows:
    """Stand-in for one of NeuralPoints.forward's gathered tensors ([1,R,SR,K,C] = table[clamp(pidx,0)],
    neural_points.py:956-972).  The fused aggregator never needs the dense tensor; `.materialize()` (or
    torch.as_tensor(handle.materialize())) builds it with sgn_gather_rows for code that does."""

    def __init__(self, ctx, table, cols=None):
        self.ctx, self.table, self.cols = ctx, table, cols

    @property
    def shape(self):
        C = self.table.shape[-1] if self.cols is None else self.cols.stop - self.cols.start
        return torch.Size(tuple(self.ctx.pidx.shape) + (C,))

    def materialize(self):
        """The dense tensor (sgn_gather_rows); built once per handle, detached as the reference's callers use it (prob outputs)."""
        if getattr(self, "_dense", None) is None:
            t = self.table.reshape(-1, self.table.shape[-1])
            out = ops.gather_rows(t.detach(), self.ctx.pidx.clamp(min=0))
            if self.cols is not None:
                out = out[..., self.cols]
            self._dense = out
        return self._dense

    # The reference's caller treats these entries as tensors in its `opt.prob == 1` block (torch.gather(sampled_xyz, 2, ...),
    # sampled_color * weight, .shape[-2] -- neural_points_volumetric_model.py:633-668).  Any torch function that meets a handle
    # gets the dense tensor; so do indexing and arithmetic.
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        dense = lambda a: a.materialize() if isinstance(a, GatheredRows) else a
        args = tuple(dense(a) for a in args)
        kwargs = {k: dense(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)

    def __getitem__(self, item):
        return self.materialize()[item]

    def __getattr__(self, name):                # .dtype, .device, .dim(), .detach(), .squeeze(), ... of the dense tensor
        if name.startswith("_") or name in ("ctx", "table", "cols"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)



def _dense_binary(op):
    def f(self, other):
        return getattr(self.materialize(), op)(other.materialize() if isinstance(other, GatheredRows) else other)
    return f


for _n in ("__mul__", "__rmul__", "__add__", "__radd__", "__sub__", "__rsub__", "__truediv__"):
    setattr(GatheredRows, _n, _dense_binary(_n))


class GatheredPers(GatheredRows):
    """sampled_xyz_pers: the gathered points in camera-perspective coordinates (neural_points.py:762, :838-850)."""

    def materialize(self):
        xyz = super().materialize()                                            # [1,R,SR,K,3]
        c = self.ctx
        flat = lighting_fast_querier.w2pers(xyz.reshape(1, -1, 3), c.camrotc2w[None], c.campos[None])
        return flat.reshape(xyz.shape)


# ---------------------------------------------------------------------------------------------------------
# querier
# ---------------------------------------------------------------------------------------------------------
class lighting_fast_querier:
    """World-coordinate voxel-grid querier.  The occupancy grid is built once per point-cloud version (keyed on the
    tensor's storage and version counter) instead of once per call."""

    def __init__(self, device, opt):
        self.device, self.opt = device, opt
        self._grid, self._grid_key, self._hp = None, None, None
        if _opt(opt, "inverse", 0) > 0:
            raise NotImplementedError("sgnerf_b200: --inverse ray generation is not built")

    def clean_up(self):                      # a no-op in the reference too (query_point_indices_worldcoords.py:62-64)
        pass

    def invalidate(self):
        if self._grid is not None:
            self._grid.close()
        self._grid, self._grid_key, self._hp = None, None, None

    def _get_grid(self, xyz):
        key = (xyz.data_ptr(), xyz._version, tuple(xyz.shape))
        if self._grid is None or key != self._grid_key:
            self.invalidate()
            o = self.opt
            self._hp = ops.grid_hyperparameters(xyz, o.vsize, o.vscale, o.kernel_size, _opt(o, "ranges", None), o.radius_limit_scale)
            # reservoir seeds of the max_o / P overflows: two time.time() reads as in the reference (:715, :751)
            self._grid = ops.OccGrid(xyz, self._hp.ranges[:3], self._hp.scaled_vsize, self._hp.scaled_vdim, o.query_size, o.P, o.max_o,
                                     seconds_claim=self._seconds(), seconds_fill=self._seconds())
            self._grid_key = key
        return self._grid, self._hp

    def _seconds(self):
        """np.uint64(time.time()) of the reference's kernel calls (:715, :751, :916); `opt.sgn_seconds` pins it (tests)."""
        fixed = _opt(self.opt, "sgn_seconds", None)
        return int(time.time()) if fixed is None else int(fixed)

    @staticmethod
    def w2pers(point_xyz_w, camrotc2w, campos):
        shift = point_xyz_w - campos[:, None, :]
        c = torch.sum(shift[..., None, :] * torch.transpose(camrotc2w, 1, 2)[:, None, None, ...], dim=-1)
        return torch.stack([c[..., 0] / c[..., 2], c[..., 1] / c[..., 2], c[..., 2]], dim=-1)

    def query_uncompacted(self, point_xyz_w_tensor, near_depth, far_depth, ray_dirs_tensor, cam_pos_tensor,
                          points_label_tensor=None, points_label_prob_tensor=None, ray_label_tensor=None, t=None):
        """Row r of the outputs belongs to input ray r (no host synchronisation).  Used by the fused path."""
        o = self.opt
        xyz = point_xyz_w_tensor.reshape(-1, 3)
        grid, hp = self._get_grid(xyz)
        raydir = ray_dirs_tensor.reshape(-1, 3)
        if t is None:
            jitter = 0.3 if _opt(o, "is_train", 0) > 0 else 0.0
            t = pipeline.middle_point_ts(float(near_depth), float(far_depth), o.z_depth_dim, raydir.device, jitter=jitter,
                                         n_rays=raydir.shape[0])
        kw = {}
        if _opt(o, "semantic_guidance", 0) == 1:
            prob = points_label_prob_tensor.reshape(-1, points_label_prob_tensor.shape[-1])
            # the reference converts the probabilities with .to(torch.int32) (values 0 / 1) and its kernel then reads that tensor
            # through a `const float*` (:916, :547): the same int32 tensor goes to sgn_query, which reinterprets the bits likewise
            bits = prob.to(torch.int32)
            kw = dict(ray_label=ray_label_tensor.reshape(-1).to(torch.int32), pt_label=points_label_tensor.reshape(-1).to(torch.int32),
                      pt_label_prob_bits=bits, seconds_query=self._seconds())
        pidx, loc_w, smask, rmask = ops.query(grid, cam_pos_tensor.reshape(3), raydir, t, o.SR, o.K, o.kernel_size[0], hp.radius2, **kw)
        return pidx, loc_w, smask, rmask, hp

    def query_points(self, pixel_idx_tensor, point_xyz_pers_tensor, point_xyz_w_tensor, actual_numpoints_tensor, h, w, intrinsic,
                     near_depth, far_depth, ray_dirs_tensor, cam_pos_tensor, cam_rot_tensor, pixel_label_tensor=None,
                     points_label_tensor=None, points_label_prob_tensor=None, ray_label_tensor=None):
        near_depth, far_depth = np.asarray(near_depth).item(), np.asarray(far_depth).item()
        pidx, loc_w, _, rmask, hp = self.query_uncompacted(point_xyz_w_tensor, near_depth, far_depth, ray_dirs_tensor, cam_pos_tensor,
                                                           points_label_tensor, points_label_prob_tensor, ray_label_tensor)
        sel = rmask > 0                                           # the reference compacts the rays with >= 1 neighbour (:946-952)
        sample_pidx = pidx[sel][None]
        sample_loc_w = loc_w[sel][None]
        sample_ray_dirs = ray_dirs_tensor.reshape(-1, 3)[sel][None, :, None, :].expand(-1, -1, self.opt.SR, -1).contiguous()
        sample_loc = self.w2pers(sample_loc_w, cam_rot_tensor, cam_pos_tensor)
        return (sample_pidx, sample_loc, sample_loc_w, sample_ray_dirs, rmask[None], np.asarray(hp.vsize, dtype=np.float32), hp.ranges)


class lighting_fast_querier_p:
    """Perspective-frustum querier (--wcoord_query 0): lighting_fast_querier of models/neural_points/query_point_indices.py:28-128.
    The grid lives in the camera's perspective coordinates, so every call rebuilds it -- for all rays of the call at once
    (ops.pers_query -> sgn_pers_query).  Like the reference's, query_points takes no semantic arguments (:76)."""

    def __init__(self, device, opt):
        self.device, self.opt = device, opt
        self.inverse = _opt(opt, "inverse", 0)

    def clean_up(self):
        pass

    def invalidate(self):
        pass

    _seconds = lighting_fast_querier._seconds

    def get_hyperparameters(self, h, w, intrinsic, near_depth, far_depth):
        o = self.opt
        return ops.pers_hyperparameters(h, w, intrinsic, near_depth, far_depth, o.z_depth_dim, o.vscale, o.radius_limit_scale,
                                        o.depth_limit_scale, self.inverse)

    @staticmethod
    def pers2w(point_xyz_pers, camrotc2w, campos):
        """:95-107 -- also the (normalised) ray direction of every sample."""
        xyz_c = torch.stack([point_xyz_pers[..., 0] * point_xyz_pers[..., 2], point_xyz_pers[..., 1] * point_xyz_pers[..., 2],
                             point_xyz_pers[..., 2]], dim=-1)
        xyz_w_shift = torch.sum(xyz_c[..., None, :] * camrotc2w, dim=-1)
        ray_dirs = xyz_w_shift / (torch.linalg.norm(xyz_w_shift, dim=-1, keepdims=True) + 1e-7)
        return xyz_w_shift + campos[:, None, :], ray_dirs

    @staticmethod
    def gaussian(input, vsize):                                            # :109-113
        B, R, SR, _ = input.shape
        jitters = torch.normal(mean=torch.zeros([B, R, SR], dtype=torch.float32, device=input.device),
                               std=torch.full([B, R, SR], vsize[2] / 4, dtype=torch.float32, device=input.device))
        input[..., 2] = input[..., 2] + torch.clamp(jitters, min=-vsize[2] / 2, max=vsize[2] / 2)
        return input

    @staticmethod
    def uniform(input, vsize):                                             # :115-119
        B, R, SR, _ = input.shape
        jitters = torch.rand([B, R, SR], dtype=torch.float32, device=input.device) - 0.5
        input[..., 2] = input[..., 2] + jitters * vsize[2]
        return input

    @staticmethod
    def passfunc(input, vsize):                                            # :121-122
        return input

    def query_uncompacted(self, pixel_idx_tensor, point_xyz_pers_tensor, h, w, intrinsic, near_depth, far_depth):
        """Rows per input ray, no host synchronisation: sample_pidx [R,SR,K], sample_loc (perspective) [R,SR,3], ray_mask int8 [R], hp."""
        o = self.opt
        hp = self.get_hyperparameters(int(h), int(w), np.asarray(intrinsic), np.asarray(near_depth).item(), np.asarray(far_depth).item())
        pidx, loc, mask = ops.pers_query(point_xyz_pers_tensor.reshape(-1, 3), pixel_idx_tensor.reshape(-1, 2), hp, o.kernel_size, o.query_size,
                                         o.SR, o.K, o.P, NN=o.NN, inverse=self.inverse, seconds=(self._seconds(), self._seconds()))
        return pidx, loc, mask, hp

    def query_points(self, pixel_idx_tensor, point_xyz_pers_tensor, point_xyz_w_tensor, actual_numpoints_tensor, h, w, intrinsic,
                     near_depth, far_depth, ray_dirs_tensor, cam_pos_tensor, cam_rot_tensor):
        pidx, loc, mask, hp = self.query_uncompacted(pixel_idx_tensor, point_xyz_pers_tensor, h, w, intrinsic, near_depth, far_depth)
        sel = mask > 0                                            # the reference keeps the rays whose pixel column is occupied (:688)
        sample_pidx, sample_loc = pidx[sel][None], loc[sel][None]
        if _opt(self.opt, "is_train", 0):
            sample_loc = getattr(self, _opt(self.opt, "shpnt_jitter", "passfunc"))(sample_loc, hp.vsize)
        B, R = 1, sample_loc.shape[1]
        loc_w, dirs = self.pers2w(sample_loc.reshape(1, -1, 3), cam_rot_tensor, cam_pos_tensor)
        return (sample_pidx, sample_loc, loc_w.reshape(B, R, self.opt.SR, 3), dirs.reshape(B, R, self.opt.SR, 3), mask[None], hp.vsize, hp.ranges)


# ---------------------------------------------------------------------------------------------------------
# NeuralPoints
# ---------------------------------------------------------------------------------------------------------
class NeuralPoints(nn.Module):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        """Registers the flags of the reference's NeuralPoints.modify_commandline_options (neural_points.py:80-309): the reference
        calls it on the imported class (neural_points_volumetric_model.py:66), so after the import swap this is where they come from."""
        return _register_flags(parser, NEURAL_POINTS_FLAGS)

    def __init__(self, num_channels, size, opt, device, checkpoint=None, feature_init_method='rand', reg_weight=0., feedforward=0):
        super().__init__()
        self.opt, self.device = opt, device
        self.grid_vox_sz = 0
        self.points_conf = self.points_dir = self.points_color = self.eulers = None
        self.points_label = self.points_label_prob = self.bpnet_points_embedding = self.points_feats = None
        self.Rw2c = torch.eye(3, device=device, dtype=torch.float32)
        self.reg_weight = reg_weight
        # neural_points.py:425: query_size falls back to kernel_size when left at its (0, 0, 0) default
        if hasattr(opt, "query_size") and hasattr(opt, "kernel_size") and opt.query_size[0] == 0:
            opt.query_size = opt.kernel_size
        # neural_points.py:426-427
        self.lighting_fast_querier = lighting_fast_querier if _opt(opt, "wcoord_query", 1) > 0 else lighting_fast_querier_p
        self.querier = self.lighting_fast_querier(device, opt)
        self.xyz, self.points_embeding = None, None
        if _opt(opt, "load_points", 1) == 1 and not feedforward:
            saved = None
            if checkpoint:
                saved = torch.load(checkpoint, map_location=device) if isinstance(checkpoint, str) else checkpoint
            if saved is not None and "neural_points.xyz" in saved:
                g = lambda k: saved.get("neural_points." + k)
                self.xyz = nn.Parameter(g("xyz").to(device))
                for name in ("points_embeding", "points_conf", "points_dir", "points_color", "eulers"):
                    v = g(name)
                    setattr(self, name, nn.Parameter(v.to(device)) if v is not None else None)
                for name in ("points_feats", "points_label"):        # constants kept in the state_dict (requires_grad False)
                    v = g(name)
                    if v is not None:
                        setattr(self, name, nn.Parameter(v.to(device), requires_grad=False))
                if g("Rw2c") is not None:
                    self.Rw2c = nn.Parameter(g("Rw2c").to(device), requires_grad=False)
            else:
                # no checkpoint: `size` points whose positions arrive through set_points; features by feature_init_method
                # (neural_points.py:386-409: rand = U(-.5, .5), zeros, ones, gau_<std>)
                n = int(size)
                self.xyz = nn.Parameter(torch.zeros(n, 3, device=device))
                shape = (1, n, num_channels)
                if feature_init_method == "rand":
                    emb = torch.rand(shape, device=device) - 0.5
                elif feature_init_method == "zeros":
                    emb = torch.zeros(shape, device=device)
                elif feature_init_method == "ones":
                    emb = torch.ones(shape, device=device)
                elif feature_init_method.startswith("gau"):
                    emb = torch.normal(mean=torch.zeros(shape, device=device), std=float(feature_init_method.split("_")[1]))
                else:
                    raise ValueError(feature_init_method)
                self.points_embeding = nn.Parameter(emb)
                self.points_conf = torch.ones_like(self.points_embeding[..., 0:1])
            self.xyz.requires_grad = _opt(opt, "xyz_grad", 0) > 0
            for name, flag in (("points_embeding", "feat_grad"), ("points_conf", "conf_grad"), ("points_dir", "dir_grad"),
                               ("points_color", "color_grad")):
                t = getattr(self, name)
                if isinstance(t, nn.Parameter):
                    t.requires_grad = _opt(opt, flag, 1 if flag != "dir_grad" else 0) > 0
            if self.eulers is not None:
                self.eulers.requires_grad = False

    # ---- point-cloud edits: every one of them moves the tensors, so the querier's grid key changes on its own ----
    def reset_querier(self):
        self.querier.clean_up()
        self.querier.invalidate()

    def set_points(self, points_xyz, points_feats, points_embeding, points_label=None, points_color=None, points_dir=None,
                   points_conf=None, points_semantic=None, parameter=False, Rw2c=None, eulers=None):
        o = self.opt
        grad = lambda flag, default: _opt(o, flag, default) > 0
        wrap = (lambda t, g=True: nn.Parameter(t, requires_grad=g)) if parameter else (lambda t, g=True: t)
        three = lambda t: None if t is None else (t if t.dim() == 3 else t[None, ...])
        if points_embeding.shape[-1] > _opt(o, "point_features_dim", points_embeding.shape[-1]):
            points_embeding = points_embeding[..., :o.point_features_dim]
        dc = _opt(o, "default_conf", -1.0)
        if 0.0 < dc <= 1.0 and points_conf is not None:
            points_conf = torch.ones_like(points_conf) * dc
        for name, mode in (("conf", "point_conf_mode"), ("dir", "point_dir_mode"), ("color", "point_color_mode")):
            if "0" in list(_opt(o, mode, "1")):
                raise NotImplementedError(f"sgnerf_b200: --{mode} 0 (attribute concatenated into the embedding) is not built; use '1'")
        for name in ("xyz", "points_feats", "points_label", "points_embeding", "points_conf", "points_dir", "points_color", "eulers", "Rw2c"):
            self._parameters.pop(name, None)          # a plain tensor may replace what was a Parameter, and back
        self.xyz = wrap(points_xyz.reshape(-1, 3), grad("xyz_grad", 0))
        self.points_feats = None if points_feats is None else wrap(points_feats, False)
        self.points_label = None if points_label is None else wrap(points_label, False)
        self.points_embeding = wrap(three(points_embeding), grad("feat_grad", 1))
        self.points_conf = None if points_conf is None else wrap(three(points_conf), grad("conf_grad", 1))
        self.points_dir = None if points_dir is None else wrap(three(points_dir), grad("dir_grad", 0))
        self.points_color = None if points_color is None else wrap(three(points_color), grad("color_grad", 1))
        self.eulers = None if eulers is None else wrap(eulers, False)
        # the reference resets Rw2c to the identity when none is given (neural_points.py:647-651)
        self.Rw2c = torch.eye(3, device=self.xyz.device, dtype=torch.float32) if Rw2c is None else nn.Parameter(Rw2c, requires_grad=False)
        self.reset_querier()

    def editing_set_points(self, points_xyz, points_embeding, points_color=None, points_dir=None, points_conf=None, parameter=False,
                           Rw2c=None, eulers=None):
        self.set_points(points_xyz, None, points_embeding, points_color=points_color, points_dir=points_dir, points_conf=points_conf,
                        parameter=parameter, Rw2c=Rw2c, eulers=eulers)

    def set_bpnet_feats(self, points_label_prob, points_label, bpnet_points_embedding):
        self.points_label_prob = points_label_prob
        self.points_label = points_label
        self.bpnet_points_embedding = bpnet_points_embedding if bpnet_points_embedding.dim() == 3 else bpnet_points_embedding[None, ...]

    def prune(self, thresh):
        mask = self.points_conf[0, :, 0] >= thresh
        self.xyz = nn.Parameter(self.xyz[mask, :], requires_grad=self.xyz.requires_grad)
        for name in ("points_embeding", "points_conf", "points_dir", "points_color"):
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, nn.Parameter(t[:, mask, :]))
        if self.points_label is not None:
            self.points_label = self.points_label[mask]
        self.reset_querier()

    def grow_points(self, add_xyz, add_embedding, add_color, add_dir, add_conf, add_label=None, add_eulers=None, add_Rw2c=None):
        self.xyz = nn.Parameter(torch.cat([self.xyz, add_xyz], dim=0), requires_grad=self.xyz.requires_grad)
        cat = lambda cur, add: nn.Parameter(torch.cat([cur, add[None, ...] if add.dim() == 2 else add], dim=1))
        if self.points_embeding is not None:
            self.points_embeding = cat(self.points_embeding, add_embedding)
        if self.points_conf is not None:
            self.points_conf = cat(self.points_conf, add_conf)
        if self.points_dir is not None:
            self.points_dir = cat(self.points_dir, add_dir)
        if self.points_color is not None:
            self.points_color = cat(self.points_color, add_color)
        if self.points_label is not None and add_label is not None:
            self.points_label = torch.cat([self.points_label, add_label], dim=0)
        self.reset_querier()

    def construct_grid_points(self, xyz):
        """neural_points.py:685-712: a regular grid of grid_res^3 cells over the cloud's bounding cube (x 1.1); the grid points of every
        construction voxel (construct_res^3) that holds a point become the neural points, `full_grid_idx` maps a grid coordinate to its
        point (or -1).  Init-time host logic (torch); the per-sample lookup is sgn_query_vox_grid.  Returns (xyz, sparse_grid_idx,
        full_grid_idx) and keeps space_min / grid_vox_sz / full_grid_idx for forward()."""
        o = self.opt
        xyz_min, xyz_max = torch.min(xyz, dim=-2)[0], torch.max(xyz, dim=-2)[0]
        self.space_edge = torch.max(xyz_max - xyz_min) * 1.1
        mid = (xyz_max + xyz_min) / 2
        self.space_min, self.space_max = mid - self.space_edge / 2, mid + self.space_edge / 2
        self.construct_vox_sz = self.space_edge / o.construct_res
        self.grid_vox_sz = self.space_edge / o.grid_res
        vox = torch.unique(torch.floor((xyz - self.space_min[None]) / self.construct_vox_sz[None]).to(torch.int16), dim=0)
        ratio = int(o.grid_res / o.construct_res)
        g = torch.arange(0, ratio + 1, device=vox.device, dtype=vox.dtype)
        cell = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), dim=-1).view(1, -1, 3)
        sparse = torch.unique((vox[:, None, :] * ratio + cell).view(-1, 3), dim=0).to(torch.int64)
        full = torch.full([o.grid_res + 1] * 3, -1, device=xyz.device, dtype=torch.int32)
        full[sparse[:, 0], sparse[:, 1], sparse[:, 2]] = torch.arange(0, sparse.shape[0], device=xyz.device, dtype=torch.int32)
        self.full_grid_idx = full
        return self.space_min[None] + sparse * self.grid_vox_sz, sparse, full

    def getPointsData(self):
        return self.xyz.data.cpu().numpy().copy(), None if self.points_feats is None else self.points_feats.data.cpu().numpy().copy()

    def null_grad(self):
        for name in ("points_embeding", "xyz"):
            t = getattr(self, name)
            if t is not None:
                t.grad = None

    def reg_loss(self):
        return self.reg_weight * torch.mean(torch.pow(self.points_embeding, 2))

    def w2pers(self, point_xyz, camrotc2w, campos):
        return lighting_fast_querier.w2pers(point_xyz[None, ...] if point_xyz.dim() == 2 else point_xyz, camrotc2w, campos)

    def forward(self, inputs):
        camrotc2w, campos = inputs["camrotc2w"], inputs["campos"]
        near, far = float(torch.min(inputs["near"])), float(torch.max(inputs["far"]))
        raydir = inputs["raydir"]
        vox_query = _opt(self.opt, "NN", 2) < 0
        if vox_query and getattr(self, "full_grid_idx", None) is None:
            raise RuntimeError("sgnerf_b200: --NN < 0 queries the construction grid: call construct_grid_points(xyz) first "
                               "(the reference builds it in its constructor when --construct_res > 0, neural_points.py:344)")
        semantic = _opt(self.opt, "semantic_guidance", 0) == 1 or _opt(self.opt, "predict_semantic", 0) == 1
        kw = {}
        if _opt(self.opt, "semantic_guidance", 0) == 1:
            kw = dict(points_label_tensor=self.points_label, points_label_prob_tensor=self.points_label_prob,
                      ray_label_tensor=inputs["pixel_label"])
        pers = isinstance(self.querier, lighting_fast_querier_p)
        if pers:
            # neural_points.py:760-792 with the perspective querier: per-sample ray directions from pers2w, sample_loc straight from the kernel
            xyz_pers = self.w2pers(self.xyz, camrotc2w, campos)
            sample_pidx, sample_loc_p, sample_loc_w, sample_ray_dirs_p, rmask1, vsize_p, _ = self.querier.query_points(
                inputs["pixel_idx"].to(torch.int32), xyz_pers, self.xyz[None, ...], None, int(torch.max(inputs["h"])), int(torch.max(inputs["w"])),
                inputs["intrinsic"].cpu().numpy()[0], near, far, raydir, campos, camrotc2w, **kw)
            rmask, sel = rmask1[0], rmask1[0] > 0
            sample_pidx, sample_loc_w = sample_pidx.contiguous(), sample_loc_w.contiguous()
            hp = SimpleNamespace(vsize=vsize_p)
        else:
            pidx_u, loc_w_u, _, rmask, hp = self.querier.query_uncompacted(self.xyz, near, far, raydir, campos, **kw)
            sel = rmask > 0                                      # host sync, as in the reference (:946); the fused frame path avoids it
            sample_pidx = pidx_u[sel][None].contiguous()
            sample_loc_w = loc_w_u[sel][None].contiguous()
        if vox_query:
            # neural_points.py:799-803: the 8 corners of the construction-grid cell of every sample replace the K-NN result
            sample_pidx = ops.query_vox_grid(sample_loc_w, self.full_grid_idx, self.space_min, float(self.grid_vox_sz),
                                             self.opt.grid_res).to(torch.int32) if sample_pidx.shape[1] > 0 else \
                torch.zeros(1, 0, self.opt.SR, 8, device=sample_pidx.device, dtype=torch.int32)
        if pers:
            # all samples of a ray sit on the line through its sub-pixel centre: one direction per ray for the fused aggregator
            sample_ray_dirs, sample_loc = sample_ray_dirs_p, sample_loc_p
            rd = sample_ray_dirs_p[0, :, 0, :]
        else:
            rd = raydir.reshape(-1, 3)[sel]
            sample_ray_dirs = rd[None, :, None, :].expand(-1, -1, self.opt.SR, -1).contiguous()
            sample_loc = lighting_fast_querier.w2pers(sample_loc_w, camrotc2w, campos)
        ctx = SimpleNamespace(neural_points=self, pidx=sample_pidx, loc_w=sample_loc_w, raydir=rd.contiguous(), campos=campos.reshape(3),
                              camrotc2w=camrotc2w.reshape(3, 3), vsize=hp.vsize)
        xyz_tab = self.xyz[None, ...]
        g = lambda table, cols=None: None if table is None else GatheredRows(ctx, table, cols)
        sampled_label_embedding = g(self.bpnet_points_embedding) if semantic and self.bpnet_points_embedding is not None else None
        ctx.use_label = sampled_label_embedding is not None
        vsize = np.asarray(hp.vsize, dtype=np.float32)
        return (g(self.points_color), sampled_label_embedding, self.Rw2c, g(self.points_dir), g(self.points_conf), g(self.points_embeding),
                GatheredPers(ctx, xyz_tab), g(xyz_tab), sample_pidx >= 0, sample_loc, sample_loc_w, sample_ray_dirs, rmask[None], vsize,
                self.grid_vox_sz)


# ---------------------------------------------------------------------------------------------------------
# PointAggregator
# ---------------------------------------------------------------------------------------------------------
def _init_seq(seq, slope):
    """Xavier-uniform initialisation of helpers/networks.py:120-172 (init_seq): gain('leaky_relu', slope) for a Linear
    followed by the activation, gain 1 for the last module of the Sequential; biases 0."""
    mods = list(seq)

    def xavier(m, gain):
        if isinstance(m, nn.Linear):
            fan_out, fan_in = m.weight.shape
            std = gain * math.sqrt(2.0 / (fan_in + fan_out))
            m.weight.data.uniform_(-std * math.sqrt(3.0), std * math.sqrt(3.0))
            if m.bias is not None:
                m.bias.data.zero_()
    for a, b in zip(mods[:-1], mods[1:]):
        xavier(a, nn.init.calculate_gain('leaky_relu', slope) if isinstance(b, nn.LeakyReLU) else 1.0)
    xavier(mods[-1], 1.0)


class PointAggregator(nn.Module):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        """Registers the flags of the reference's PointAggregator.modify_commandline_options (point_aggregators.py:15-253), called on
        the imported class at neural_points_volumetric_model.py:67."""
        return _register_flags(parser, POINT_AGGREGATOR_FLAGS)

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        if _opt(opt, "which_agg_model", "viewmlp") != "viewmlp" or _opt(opt, "agg_distance_kernel", "linear") != "linear":
            raise NotImplementedError("sgnerf_b200: only which_agg_model=viewmlp with agg_distance_kernel=linear is built")
        if _opt(opt, "agg_dist_pers", 20) != 20 or _opt(opt, "agg_intrp_order", 2) != 2 or _opt(opt, "act_type", "LeakyReLU") != "LeakyReLU":
            raise NotImplementedError("sgnerf_b200: canonical agg_dist_pers=20 / agg_intrp_order=2 / LeakyReLU only")
        if _opt(opt, "shading_feature_mlp_layer2", 0) != 0 or _opt(opt, "shading_feature_mlp_layer0", 0) != 0:
            raise NotImplementedError("sgnerf_b200: block0 / block2 (shading_feature_mlp_layer0/2) are not built")
        slope = 0.01
        C, W = opt.point_features_dim, opt.shading_feature_num
        self.label_dim = 96 if _opt(opt, "shading_feature_mlp_layer2_bpnet", 0) > 0 else 0
        in_ch = C + 2 * opt.num_feat_freqs * C + 2 * opt.dist_xyz_freq * 6
        seqs = []

        def mlp(cin, widths, final_act):
            mods = []
            for i, w in enumerate(widths):
                mods.append(nn.Linear(cin, w))
                if final_act or i < len(widths) - 1:
                    mods.append(nn.LeakyReLU(slope, inplace=True))
                cin = w
            return nn.Sequential(*mods)
        self.block1 = mlp(in_ch, [W] * opt.shading_feature_mlp_layer1, True); seqs.append(self.block1)
        n2 = _opt(opt, "shading_feature_mlp_layer2_bpnet", 0)
        if n2 > 0:
            self.block2_bpnet = mlp(W + self.label_dim, [W] * n2, True); seqs.append(self.block2_bpnet)
        self.block3 = mlp(W + 3 + 4, [W] * opt.shading_feature_mlp_layer3, True); seqs.append(self.block3)
        na = opt.shading_alpha_mlp_layer
        if na != 1:
            raise NotImplementedError("sgnerf_b200: shading_alpha_mlp_layer must be 1")
        self.alpha_branch = mlp(W, [1], False); seqs.append(self.alpha_branch)
        nc = opt.shading_color_mlp_layer
        self.color_branch = mlp(W + 2 * opt.num_viewdir_freqs * 3, [W // 2] * (nc - 1) + [3], False); seqs.append(self.color_branch)
        for s in seqs:
            _init_seq(s, slope)
        self.cfg = ops.agg_cfg(feat_dim=C, num_feat_freqs=opt.num_feat_freqs, dist_xyz_freq=opt.dist_xyz_freq,
                               num_viewdir_freqs=opt.num_viewdir_freqs, width=W, n_block1=opt.shading_feature_mlp_layer1,
                               n_block2_bpnet=n2, label_dim=self.label_dim, n_block3=opt.shading_feature_mlp_layer3, n_color=nc,
                               act_super=_opt(opt, "act_super", 1), leaky_slope=slope)

    def _linears(self):
        seqs = [self.block1] + ([self.block2_bpnet] if hasattr(self, "block2_bpnet") else []) + [self.block3, self.alpha_branch, self.color_branch]
        return [m for s in seqs for m in s if isinstance(m, nn.Linear)]

    def _require_identity_rw2c(self, Rw2c):
        """The kernels take Rw2c = identity (include/sgnerf_b200.h; what every ScanNet script of the reference passes:
        data/scannet_ft_dataset.py:124).  The reference rotates view directions, point offsets and point directions by Rw2c^T
        (point_aggregators.py:563-653): any other matrix, or a per-point one, is refused instead of rendering wrong colours.
        One device read per Rw2c tensor version."""
        if Rw2c is None:
            return
        if Rw2c.dim() != 2:
            raise NotImplementedError("sgnerf_b200: per-point Rw2c is not built (uniform identity only)")
        key = (Rw2c.data_ptr(), Rw2c._version)
        if getattr(self, "_rw2c_ok", None) != key:
            if tuple(Rw2c.shape) != (3, 3) or not torch.equal(Rw2c.detach().float().cpu(), torch.eye(3)):
                raise NotImplementedError("sgnerf_b200: Rw2c must be the 3x3 identity (a rotated point frame is not built)")
            self._rw2c_ok = key

    def forward(self, sampled_color, sampled_label_embedding, sampled_Rw2c, sampled_dir, sampled_conf, sampled_embedding, sampled_xyz_pers,
                sampled_xyz, sample_pnt_mask, sample_loc, sample_loc_w, sample_ray_dirs, vsize, grid_vox_sz):
        if not isinstance(sampled_embedding, GatheredRows):
            raise TypeError("sgnerf_b200.PointAggregator takes the handles returned by sgnerf_b200.NeuralPoints.forward (fused gather); "
                            "dense gathered tensors are not accepted and there is no PyTorch fallback")
        ctx = sampled_embedding.ctx
        npnts = ctx.neural_points
        self._require_identity_rw2c(sampled_Rw2c)
        lin = self._linears()
        training = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        # opt.sgn_precision: "auto" = TF32 tensor-core GEMMs when training (the reference's cuBLAS default at its pinned torch), bf16
        # tensor-core kernels for inference; "fp32" = strict fp32 SIMT everywhere; "tf32" = the layer-wise TF32 path everywhere
        mode = _opt(self.opt, "sgn_precision", "auto")
        if mode == "fp32":
            precision = ops.PRECISION_FP32
        elif training or mode == "tf32":
            precision = ops.PRECISION_TF32
        else:
            precision = ops.PRECISION_BF16
        opt = self.opt
        want = not ((_opt(opt, "sparse_loss_weight", 0) <= 0) and ("conf_coefficient" not in _opt(opt, "zero_one_loss_items", "")) and _opt(opt, "prob", 0) == 0)
        label = npnts.bpnet_points_embedding[0] if (self.label_dim > 0 and npnts.bpnet_points_embedding is not None) else None
        decoded, ray_valid, _, weight, conf = ops.aggregate(
            self.cfg, [m.weight for m in lin], [m.bias for m in lin], npnts.xyz, npnts.points_embeding[0], npnts.points_color[0],
            npnts.points_dir[0], None if npnts.points_conf is None else npnts.points_conf[0, :, 0], label, ctx.pidx[0], ctx.loc_w[0],
            ctx.raydir, ctx.campos, ctx.camrotc2w, precision=precision, want_aux=want)
        out = (decoded[None], ray_valid[None].bool(), weight[None] if want else None, conf[None] if want else None)
        return out


# ---------------------------------------------------------------------------------------------------------
# ray marching
# ---------------------------------------------------------------------------------------------------------
def _blend_id(blend_func):
    name = getattr(blend_func, "__name__", str(blend_func))
    if name == "alpha_blend":
        return 0
    if name == "alpha2_blend":
        return 1
    raise NotImplementedError(f"sgnerf_b200: blend function {name} is not built (alpha / alpha2 only)")


def ray_march(ray_dist, ray_valid, ray_features, render_func, blend_func, bg_color=None):
    """7-tuple of diff_ray_marching.py:509-555: (ray_color, point_color, opacity, acc_transmission, blend_weight[...,None],
    background_transmission[...,None], background_blend_weight)."""
    if getattr(render_func, "__name__", "") not in ("radiance_render",):
        raise NotImplementedError("sgnerf_b200: only the radiance render function is built")
    blend = _blend_id(blend_func)
    bg = None if bg_color is None else bg_color.to(ray_features.device).float().reshape(-1)[:3]
    ray_color, opacity, acc, bw, bgt = ops.composite(ray_features, ray_dist, ray_valid, bg, blend=blend)
    bgt = bgt[..., None]
    return ray_color, ray_features[..., 1:4], opacity, acc, bw[..., None], bgt, (bgt if blend == 0 else bgt * bgt)


def alpha_ray_march(ray_dist, ray_valid, ray_features, blend_func):
    """5-tuple of diff_ray_marching.py:558-573."""
    blend = _blend_id(blend_func)
    _, opacity, acc, bw, bgt = ops.composite(ray_features, ray_dist, ray_valid, None, blend=blend)
    bgt = bgt[..., None]
    return opacity, acc, bw[..., None], bgt, (bgt if blend == 0 else bgt * bgt)
