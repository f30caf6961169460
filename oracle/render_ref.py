"""oracle/render_ref.py -- TEST INFRASTRUCTURE ONLY.

torch-CPU (fp32, autograd-capable) restatement of the floating-point half of the reference's
per-ray render path, for the canonical configuration:

  gather            models/neural_points/neural_points.py:956-988  (+ w2pers :838-850)
  aggregation       models/aggregators/point_aggregators.py:868-959 (forward), :494-502 (linear),
                    :561-786 (viewmlp), :298-309 (activations), :863-865 (gradiant_clamp)
  pos. encoding     models/helpers/networks.py:175-192, weight init :120-172
  step sizes        models/neural_points_volumetric_model.py:569-577
  compositing       models/rendering/diff_ray_marching.py:509-573, diff_render_func.py:36-49
  fill_invalid      models/neural_points_volumetric_model.py:158-195

It is checked against the reference's own modules (imported from /root/reference in the build
container, see tests/golden/make_golden.py) and against the golden vectors that script commits.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F


def agg_config(**kw):
    """Aggregator options of the canonical config.  semantic=True adds block2_bpnet (352->256) and the
    96-d label embedding input (shading_feature_mlp_layer2_bpnet=1, predict_semantic=1)."""
    c = SimpleNamespace(
        point_features_dim=32, num_feat_freqs=3, dist_xyz_freq=5, num_viewdir_freqs=4,
        shading_feature_num=256, shading_feature_mlp_layer1=2, shading_feature_mlp_layer2_bpnet=0,
        shading_feature_mlp_layer3=2, shading_alpha_mlp_layer=1, shading_color_mlp_layer=4,
        shading_color_channel_num=3, label_embedding_dim=0, act_super=1, leaky_slope=0.01,
        return_weight=True)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def semantic_config(**kw):
    return agg_config(shading_feature_mlp_layer2_bpnet=1, label_embedding_dim=96, **kw)


def layer_shapes(cfg):
    """(name, in, out) of every Linear, in reference state_dict order (point_aggregators.py:312-418)."""
    C, W = cfg.point_features_dim, cfg.shading_feature_num
    dist_dim = 6                                                    # agg_dist_pers == 20
    in_ch = C + 2 * cfg.num_feat_freqs * C + 2 * cfg.dist_xyz_freq * dist_dim
    out = []
    for i in range(cfg.shading_feature_mlp_layer1):
        out.append((f"block1.{2 * i}", in_ch, W)); in_ch = W
    if cfg.shading_feature_mlp_layer2_bpnet > 0:
        in_ch += cfg.label_embedding_dim
        for i in range(cfg.shading_feature_mlp_layer2_bpnet):
            out.append((f"block2_bpnet.{2 * i}", in_ch, W)); in_ch = W
    if cfg.shading_feature_mlp_layer3 > 0:
        in_ch += 3 + 4                                               # point_color_mode "1", point_dir_mode "1"
        for i in range(cfg.shading_feature_mlp_layer3):
            out.append((f"block3.{2 * i}", in_ch, W)); in_ch = W
    a_in = W
    for i in range(cfg.shading_alpha_mlp_layer - 1):
        out.append((f"alpha_branch.{2 * i}", a_in, W // 2)); a_in = W // 2
    out.append((f"alpha_branch.{2 * (cfg.shading_alpha_mlp_layer - 1)}", a_in, 1))
    c_in = W + 2 * cfg.num_viewdir_freqs * 3
    for i in range(cfg.shading_color_mlp_layer - 1):
        out.append((f"color_branch.{2 * i}", c_in, W // 2)); c_in = W // 2
    out.append((f"color_branch.{2 * (cfg.shading_color_mlp_layer - 1)}", c_in, cfg.shading_color_channel_num))
    return out


def init_params(cfg, seed=0, bias_scale=0.0):
    """Xavier-uniform init of helpers/networks.py:120-172: gain('leaky_relu', slope) when an activation
    follows, 1 for the last Linear of each Sequential; bias 0 (bias_scale>0 randomises it for tests)."""
    g = torch.Generator().manual_seed(seed)
    gain_act = math.sqrt(2.0 / (1 + cfg.leaky_slope ** 2))
    shapes = layer_shapes(cfg)
    last = {}
    for name, _, _ in shapes:
        last[name.split(".")[0]] = name
    P = {}
    for name, cin, cout in shapes:
        is_last = last[name.split(".")[0]] == name
        # block1/2/3 end with an activation, so their last Linear is still followed by LeakyReLU in the
        # Sequential; init_seq applies gain to every (Linear, act) pair and gain 1 only to s[-1], which
        # for those blocks is the activation module itself (a no-op).
        followed_by_act = (not is_last) or name.startswith("block")
        gain = gain_act if followed_by_act else 1.0
        bound = gain * math.sqrt(2.0 / (cin + cout)) * math.sqrt(3.0)
        P[name + ".weight"] = (torch.rand(cout, cin, generator=g) * 2 - 1) * bound
        P[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bias_scale
    return P


def positional_encoding(positions, freqs, ori=False):
    """helpers/networks.py:175-192."""
    freq_bands = 2.0 ** torch.arange(freqs, dtype=torch.float32)
    pts = (positions[..., None] * freq_bands).reshape(positions.shape[:-1] + (freqs * positions.shape[-1],))
    if ori:
        return torch.cat([positions, torch.sin(pts), torch.cos(pts)], dim=-1)
    return torch.stack([torch.sin(pts), torch.cos(pts)], dim=-1).reshape(pts.shape[:-1] + (pts.shape[-1] * 2,))


def w2pers_points(xyz, camrotc2w, campos):
    """neural_points.py:838-850.  xyz [N,3] -> [1,N,3]."""
    shift = xyz[None, ...] - campos[:, None, :]
    c = torch.sum(camrotc2w[:, None, :, :] * shift[:, :, :, None], dim=-2)
    return torch.stack([c[:, :, 0] / c[:, :, 2], c[:, :, 1] / c[:, :, 2], c[:, :, 2]], dim=-1)


def gather_neighbors(tables, sample_pidx, camrotc2w, campos):
    """neural_points.py:956-988.  tables: namespace(xyz [N,3], embedding [1,N,C], color [1,N,3],
    dir [1,N,3], conf [1,N,1], label_embedding [1,N,E] or None).  sample_pidx int [1,R,SR,K]."""
    B, R, SR, K = sample_pidx.shape
    mask = sample_pidx >= 0
    idx = torch.clamp(sample_pidx, min=0).view(-1).long()
    xyz_pers = w2pers_points(tables.xyz, camrotc2w, campos)
    C = tables.embedding.shape[2]
    cat = torch.cat([tables.xyz[None, ...], xyz_pers, tables.embedding], dim=-1)
    g = torch.index_select(cat, 1, idx).view(B, R, SR, K, C + 6)
    sel = lambda t: None if t is None else torch.index_select(t, 1, idx).view(B, R, SR, K, t.shape[2])
    return SimpleNamespace(color=sel(tables.color), label_embedding=sel(getattr(tables, "label_embedding", None)),
                           dir=sel(tables.dir), conf=sel(tables.conf), embedding=g[..., 6:], xyz_pers=g[..., 3:6],
                           xyz=g[..., :3], pnt_mask=mask)


def _mlp(P, prefix, n_layers, x, slope, last_act=True):
    for i in range(n_layers):
        x = F.linear(x, P[f"{prefix}.{2 * i}.weight"], P[f"{prefix}.{2 * i}.bias"])
        if last_act or i < n_layers - 1:
            x = F.leaky_relu(x, slope)
    return x


def aggregator_forward(P, cfg, sampled_color, sampled_label_embedding, sampled_dir, sampled_conf, sampled_embedding,
                       sampled_xyz_pers, sampled_xyz, sample_pnt_mask, sample_loc, sample_loc_w, sample_ray_dirs):
    """point_aggregators.py:868-959 + viewmlp :561-786 on the canonical branch: agg_dist_pers=20,
    linear kernel with unit axis weights, agg_weight_norm=1, agg_intrp_order=2, apply_pnt_mask=1,
    dist_xyz_deno=0, Rw2c = identity, agg_*_xyz_mode = None.
    Returns decoded [B,R,SR,4], ray_valid [B,R,SR], weight [B,R,SR,K], conf_coefficient [B,R,SR,K]."""
    B, R, SR, K = sample_pnt_mask.shape
    ray_valid = torch.any(sample_pnt_mask, dim=-1).view(-1)
    total = ray_valid.numel()
    if total == 0 or int(ray_valid.sum()) == 0:                                       # :888-890
        return torch.zeros(B, R, SR, cfg.shading_color_channel_num + 1), ray_valid.view(B, R, SR), None, None
    # :917-925  dists = [xyz - loc_w | (x_p z_p - s_x s_z, y_p z_p - s_y s_z, z_p - s_z)]
    xd = sampled_xyz_pers[..., 0] * sampled_xyz_pers[..., 2] - sample_loc[:, :, :, None, 0] * sample_loc[:, :, :, None, 2]
    yd = sampled_xyz_pers[..., 1] * sampled_xyz_pers[..., 2] - sample_loc[:, :, :, None, 1] * sample_loc[:, :, :, None, 2]
    zd = sampled_xyz_pers[..., 2] - sample_loc[:, :, :, None, 2]
    dists = torch.cat([sampled_xyz - sample_loc_w[..., None, :], torch.stack([xd, yd, zd], dim=-1)], dim=-1)
    # :494-502 + :946-947
    weight = sample_pnt_mask * (1.0 / torch.clamp(torch.norm(dists[..., :3], dim=-1), min=1e-6))
    weight = weight / torch.clamp(torch.sum(weight, dim=-1, keepdim=True), min=1e-8)
    # :863-865, :953  straight-through clamp
    conf = sampled_conf[..., 0]
    conf_coefficient = conf - (conf - torch.clamp(conf, min=0.0001, max=1)).detach()
    w = (weight * conf_coefficient).view(B * R * SR, K, 1)

    m = sample_pnt_mask.view(-1)
    # viewdir features: ori=True, first 3 stripped (:579-585)
    vd = sample_ray_dirs.view(-1, 3)
    vd_pe = positional_encoding(vd, cfg.num_viewdir_freqs, ori=True)
    ori_vd, vd_feat = vd_pe[..., :3], vd_pe[..., 3:]
    # per-neighbour input (:594-611)
    dists_flat = dists.view(-1, 6)[m, :]
    dists_flat = positional_encoding(dists_flat, cfg.dist_xyz_freq)
    feat = sampled_embedding.reshape(-1, sampled_embedding.shape[-1])[m, :]
    feat = torch.cat([feat, positional_encoding(feat, cfg.num_feat_freqs)], dim=-1)
    feat = torch.cat([feat, dists_flat], dim=-1)
    feat = _mlp(P, "block1", cfg.shading_feature_mlp_layer1, feat, cfg.leaky_slope)                    # :620
    if cfg.shading_feature_mlp_layer2_bpnet > 0:                                                       # :629-636
        if sampled_label_embedding is not None:
            feat = torch.cat([feat, sampled_label_embedding.reshape(-1, sampled_label_embedding.shape[-1])[m, :]], dim=-1)
        feat = _mlp(P, "block2_bpnet", cfg.shading_feature_mlp_layer2_bpnet, feat, cfg.leaky_slope)
    if cfg.shading_feature_mlp_layer3 > 0:                                                              # :638-653
        col = sampled_color.reshape(-1, 3)[m, :]
        d = sampled_dir.reshape(-1, 3)[m, :]
        ov = ori_vd[..., None, :].repeat(1, K, 1).view(-1, 3)[m, :]
        feat = torch.cat([feat, col, d - ov, torch.sum(d * ov, dim=-1, keepdim=True)], dim=-1)
        feat = _mlp(P, "block3", cfg.shading_feature_mlp_layer3, feat, cfg.leaky_slope)
    # :743-780
    raw_alpha = _mlp(P, "alpha_branch", cfg.shading_alpha_mlp_layer, feat, cfg.leaky_slope, last_act=False)
    alpha = F.softplus(raw_alpha - 1) if cfg.act_super > 0 else F.relu(raw_alpha)
    alpha_holder = torch.zeros(B * R * SR * K, 1).index_put((m.nonzero()[:, 0],), alpha)
    alpha_s = torch.sum(alpha_holder.view(B * R * SR, K, 1) * w, dim=-2)[ray_valid, :]
    feat_holder = torch.zeros(B * R * SR * K, feat.shape[-1]).index_put((m.nonzero()[:, 0],), feat)
    feat_s = torch.sum(feat_holder.view(B * R * SR, K, -1) * w, dim=-2)[ray_valid, :]
    color_in = torch.cat([feat_s, vd_feat[ray_valid, :]], dim=-1)
    raw_c = _mlp(P, "color_branch", cfg.shading_color_mlp_layer, color_in, cfg.leaky_slope, last_act=False)
    color = torch.sigmoid(raw_c)
    if cfg.act_super > 0:
        color = color * (1 + 2 * 0.001) - 0.001
    out = torch.zeros(total, cfg.shading_color_channel_num + 1).index_put(
        (ray_valid.nonzero()[:, 0],), torch.cat([alpha_s, color], dim=-1))
    if not cfg.return_weight:
        weight, conf_coefficient = None, None
    return out.view(B, R, SR, -1), ray_valid.view(B, R, SR), weight, conf_coefficient


def ray_dist_from_samples(sample_loc, ray_valid, vsize_z, raydist_mode_unit=1):
    """models/neural_points_volumetric_model.py:569-577.  sample_loc [B,R,SR,3] (perspective coords)."""
    z = torch.cummax(sample_loc[..., 2], dim=-1)[0]
    d = torch.cat([z[..., 1:] - z[..., :-1], torch.full((z.shape[0], z.shape[1], 1), vsize_z)], dim=-1)
    mask = d < 1e-8
    if raydist_mode_unit > 0:
        mask = torch.logical_or(mask, d > 2 * vsize_z)
    mask = mask.to(torch.float32)
    d = d * (1.0 - mask) + mask * vsize_z
    return d * ray_valid.float()


def ray_march(ray_dist, ray_valid, ray_features, bg_color=None, blend="alpha"):
    """diff_ray_marching.py:509-555 with render_func=radiance (diff_render_func.py:48-49) and
    blend_func alpha / alpha2 (:36-45).  Returns the reference's 7-tuple."""
    point_color = ray_features[..., 1:4]
    sigma = ray_features[..., 0] * ray_valid.float()
    opacity = 1 - torch.exp(-sigma * ray_dist)
    acc = torch.cumprod(1. - opacity + 1e-10, dim=-1)
    bg_t = acc[:, :, [-1]]
    acc = torch.cat([torch.ones(opacity.shape[0:2] + (1,)), acc[:, :, :-1]], dim=-1)
    bf = (lambda o, t: o * t) if blend == "alpha" else (lambda o, t: o * t * t)
    blend_weight = bf(opacity, acc)[..., None]
    ray_color = torch.sum(point_color * blend_weight, dim=-2)
    if bg_color is not None:
        ray_color = ray_color + bg_color.float().view(bg_t.shape[0], 1, 3) * bg_t
    return ray_color, point_color, opacity, acc, blend_weight, bg_t, bf(1, bg_t)


def alpha_ray_march(ray_dist, ray_valid, ray_features, blend="alpha"):
    """diff_ray_marching.py:558-573."""
    r = ray_march(ray_dist, ray_valid, ray_features, None, blend)
    return r[2], r[3], r[4], r[5], r[6]


def fill_invalid(ray_mask, ray_color, opacity, bg_transmission, bg_color):
    """models/neural_points_volumetric_model.py:158-195 (tonemap off): scatter the R'' rendered rays back
    to all R rays; misses get the background colour, opacity 0, is_background 1."""
    B, OR = ray_mask.shape
    inds = torch.nonzero(ray_mask)
    is_bg = torch.ones(B, OR, 1)
    is_bg[inds[:, 0], inds[:, 1], :] = bg_transmission
    color = torch.ones(B, OR, 3) * bg_color[None, ...]
    color[inds[:, 0], inds[:, 1], :] = ray_color
    op = torch.zeros(B, OR, opacity.shape[2])
    op[inds[:, 0], inds[:, 1], :] = opacity
    return color, op, is_bg


def probe_outputs(opacity, sample_loc_w, weight, conf_coefficient, gn):
    """`prob == 1` outputs of NeuralPointsRayMarching.forward (models/neural_points_volumetric_model.py:633-656), the inputs of
    probe_hole / point growing (run/train_ft.py:425-540): per ray, the sample of largest opacity, its world position, the distance
    to its nearest gathered neighbour (all K slots: invalid ones hold point 0, as clamp(pidx, 0) gathers it) and the
    weight * conf averages of its neighbours' colour / dir / conf / embedding.  The same torch calls as the reference.
    opacity [B,R,SR]; sample_loc_w [B,R,SR,3]; weight, conf_coefficient [B,R,SR,K]; gn = gather_neighbors(...)."""
    out = {}
    out["ray_max_shading_opacity"], opacity_ind = torch.max(opacity, dim=-1, keepdim=True)
    opacity_ind = opacity_ind[..., None]
    out["ray_max_sample_loc_w"] = torch.gather(sample_loc_w, 2, opacity_ind.expand(-1, -1, -1, sample_loc_w.shape[-1])).squeeze(2)
    w = torch.gather(weight * conf_coefficient, 2, opacity_ind.expand(-1, -1, -1, weight.shape[-1])).squeeze(2)[..., None]
    opacity_ind = opacity_ind[..., None]
    g5 = lambda t: torch.gather(t, 2, opacity_ind.expand(-1, -1, -1, t.shape[-2], t.shape[-1])).squeeze(2)
    xyz_max = g5(gn.xyz)
    out["ray_max_far_dist"] = torch.min(torch.norm(xyz_max - out["ray_max_sample_loc_w"][..., None, :], dim=-1), axis=-1, keepdim=True)[0]
    out["shading_avg_color"] = torch.sum(g5(gn.color) * w, dim=-2)
    out["shading_avg_dir"] = torch.sum(g5(gn.dir) * w, dim=-2)
    out["shading_avg_conf"] = torch.sum(g5(gn.conf) * w, dim=-2)
    out["shading_avg_embedding"] = torch.sum(g5(gn.embedding) * w, dim=-2)
    return out


def render_from_query(P, cfg, tables, sample_pidx, sample_loc, sample_loc_w, sample_ray_dirs, ray_mask,
                      camrotc2w, campos, vsize, bg_color):
    """NeuralPointsRayMarching.forward, :541-626, from the querier outputs on."""
    g = gather_neighbors(tables, sample_pidx, camrotc2w, campos)
    decoded, ray_valid, weight, conf = aggregator_forward(
        P, cfg, g.color, g.label_embedding, g.dir, g.conf, g.embedding, g.xyz_pers, g.xyz, g.pnt_mask,
        sample_loc, sample_loc_w, sample_ray_dirs)
    ray_dist = ray_dist_from_samples(sample_loc, ray_valid, float(vsize[2]))
    rm = ray_march(ray_dist, ray_valid, decoded, bg_color)
    color, op, is_bg = fill_invalid(ray_mask, rm[0], rm[2], rm[5], bg_color)
    # :620-624 (return_depth): avg_depth = sum(w ray_ts) / (sum(w) + 1e-6), w = opacity * acc_transmission.  The reference's
    # forward never defines `ray_ts` (the branch is dead there); upstream Point-NeRF feeds the samples' camera depth, used here.
    w_alpha = rm[2] * rm[3]
    depth = (w_alpha * sample_loc[..., 2]).sum(-1) / (w_alpha.sum(-1) + 1e-6)
    return SimpleNamespace(coarse_raycolor=color, coarse_point_opacity=op, coarse_is_background=is_bg, coarse_depth=depth,
                           decoded=decoded, ray_valid=ray_valid, weight=weight, conf_coefficient=conf,
                           ray_dist=ray_dist, ray_color=rm[0], opacity=rm[2], acc_transmission=rm[3],
                           blend_weight=rm[4], bg_transmission=rm[5])
