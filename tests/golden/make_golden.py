"""Generate the golden vectors under tests/golden/ from the REFERENCE's own modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's PointAggregator, positional_encoding, ray_march, alpha_ray_march and
near_far_linear_ray_generation from /root/reference (with a two-name scipy shim: utils/spherical.py
imports scipy.special.sph_harm / lpmn, removed in recent scipy and unused on this path), runs them on
small seeded inputs, checks oracle/render_ref.py against them, and stores inputs + reference outputs
as compressed .npz files.  tests/test_oracle_golden.py re-checks the oracle against the stored
vectors without the reference being present.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import scipy.special
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("SGN_REFERENCE_ROOT", "/root/reference")


def import_reference():
    if not hasattr(scipy.special, "sph_harm"):
        scipy.special.sph_harm = lambda *a, **k: None
    if not hasattr(scipy.special, "lpmn"):
        scipy.special.lpmn = lambda *a, **k: None
    sys.path.insert(0, REF)
    from models.aggregators.point_aggregators import PointAggregator
    from models.helpers.networks import positional_encoding
    from models.rendering.diff_ray_marching import ray_march, alpha_ray_march, near_far_linear_ray_generation
    from models.rendering.diff_render_func import find_render_function, find_blend_function
    return SimpleNamespace(PointAggregator=PointAggregator, positional_encoding=positional_encoding,
                           ray_march=ray_march, alpha_ray_march=alpha_ray_march,
                           near_far_linear_ray_generation=near_far_linear_ray_generation,
                           find_render_function=find_render_function, find_blend_function=find_blend_function)


def reference_opt(cfg):
    """The option fields PointAggregator reads (point_aggregators.py:255-421, 561-959), canonical values."""
    return SimpleNamespace(
        act_type="LeakyReLU", point_hyper_dim=256, point_features_dim=cfg.point_features_dim,
        agg_distance_kernel="linear", agg_dist_pers=20, agg_axis_weight=None, num_pos_freqs=10,
        num_viewdir_freqs=cfg.num_viewdir_freqs, view_ori=0, which_agg_model="viewmlp",
        dist_xyz_freq=cfg.dist_xyz_freq, agg_feat_xyz_mode="None", weight_feat_dim=8, weight_xyz_freq=2,
        sh_degree=4, num_feat_freqs=cfg.num_feat_freqs, agg_intrp_order=2,
        shading_feature_mlp_layer1=cfg.shading_feature_mlp_layer1, shading_feature_num=cfg.shading_feature_num,
        shading_feature_mlp_layer2=0, shading_feature_mlp_layer2_bpnet=cfg.shading_feature_mlp_layer2_bpnet,
        predict_semantic=1 if cfg.label_embedding_dim > 0 else 0,
        shading_feature_mlp_layer3=cfg.shading_feature_mlp_layer3, point_color_mode="1", point_dir_mode="1",
        point_conf_mode="1", agg_alpha_xyz_mode="None", shading_alpha_mlp_layer=cfg.shading_alpha_mlp_layer,
        agg_color_xyz_mode="None", shading_color_mlp_layer=cfg.shading_color_mlp_layer, act_super=cfg.act_super,
        apply_pnt_mask=1, dist_xyz_deno=0.0, agg_weight_norm=1, sparse_loss_weight=0.0,
        zero_one_loss_items=["conf_coefficient"], prob=0, shading_color_channel_num=3)


def make_agg_inputs(cfg, R, SR, K, seed):
    """Gathered-tensor inputs of PointAggregator.forward with a realistic mask pattern: whole samples
    empty, partially filled samples, and conf values on both sides of the [1e-4, 1] clamp."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    ru = lambda *s: torch.rand(*s, generator=g)
    n_valid = torch.randint(0, K + 1, (1, R, SR), generator=g)
    n_valid[ru(1, R, SR) < 0.35] = 0
    n_valid[ru(1, R, SR) < 0.3] = K
    mask = torch.arange(K)[None, None, None, :] < n_valid[..., None]
    loc_w = rn(1, R, SR, 3)
    xyz = loc_w[..., None, :] + 0.02 * rn(1, R, SR, K, 3)
    campos = torch.tensor([[0.3, -0.2, -4.0]])
    rot = torch.eye(3)[None]
    pers = lambda p: torch.stack([(p - campos)[..., 0] / (p - campos)[..., 2], (p - campos)[..., 1] / (p - campos)[..., 2],
                                  (p - campos)[..., 2]], dim=-1)
    d = rn(1, R, 1, 3).expand(-1, -1, SR, -1).contiguous()
    di = rn(1, R, SR, K, 3)
    inp = dict(
        sampled_color=ru(1, R, SR, K, 3), sampled_dir=di / di.norm(dim=-1, keepdim=True),
        sampled_conf=ru(1, R, SR, K, 1) * 1.3 - 0.1, sampled_embedding=ru(1, R, SR, K, cfg.point_features_dim) - 0.5,
        sampled_xyz_pers=pers(xyz), sampled_xyz=xyz, sample_pnt_mask=mask, sample_loc=pers(loc_w),
        sample_loc_w=loc_w, sample_ray_dirs=d)
    inp["sampled_label_embedding"] = rn(1, R, SR, K, cfg.label_embedding_dim) if cfg.label_embedding_dim > 0 else None
    return inp


def run_reference_aggregator(ref, cfg, P_state, inp, cot):
    opt = reference_opt(cfg)
    agg = ref.PointAggregator(opt)
    if P_state is None:
        torch.manual_seed(1)
        for n, p in agg.named_parameters():
            if n.endswith("bias"):
                p.data.uniform_(-0.1, 0.1)
        P_state = {k: v.detach().clone() for k, v in agg.state_dict().items()}
    else:
        agg.load_state_dict(P_state)
    leaves = {}
    args = {}
    for k, v in inp.items():
        if v is not None and v.dtype == torch.float32 and k in ("sampled_color", "sampled_dir", "sampled_conf",
                                                                 "sampled_embedding", "sampled_label_embedding"):
            v = v.clone().requires_grad_(True)
            leaves[k] = v
        args[k] = v
    out, ray_valid, weight, conf = agg(args["sampled_color"], args["sampled_label_embedding"], torch.eye(3),
                                       args["sampled_dir"], args["sampled_conf"], args["sampled_embedding"],
                                       args["sampled_xyz_pers"], args["sampled_xyz"], args["sample_pnt_mask"],
                                       args["sample_loc"], args["sample_loc_w"], args["sample_ray_dirs"],
                                       np.array([0.008, 0.008, 0.008]), 0)
    loss = (out * cot["decoded"]).sum() + (conf * cot["conf"]).sum()
    loss.backward()
    grads = {"g_" + k: v.grad.detach() for k, v in leaves.items()}
    grads.update({"gw_" + n: p.grad.detach() for n, p in agg.named_parameters()})
    return P_state, dict(decoded=out.detach(), ray_valid=ray_valid, weight=weight.detach(), conf=conf.detach()), grads


def run_oracle_aggregator(cfg, P_state, inp, cot):
    from oracle import render_ref as rr
    P = {k: v.clone().requires_grad_(True) for k, v in P_state.items()}
    leaves = {}
    a = {}
    for k, v in inp.items():
        if v is not None and v.dtype == torch.float32 and k in ("sampled_color", "sampled_dir", "sampled_conf",
                                                                 "sampled_embedding", "sampled_label_embedding"):
            v = v.clone().requires_grad_(True)
            leaves[k] = v
        a[k] = v
    out, ray_valid, weight, conf = rr.aggregator_forward(
        P, cfg, a["sampled_color"], a["sampled_label_embedding"], a["sampled_dir"], a["sampled_conf"],
        a["sampled_embedding"], a["sampled_xyz_pers"], a["sampled_xyz"], a["sample_pnt_mask"], a["sample_loc"],
        a["sample_loc_w"], a["sample_ray_dirs"])
    loss = (out * cot["decoded"]).sum() + (conf * cot["conf"]).sum()
    loss.backward()
    grads = {"g_" + k: v.grad.detach() for k, v in leaves.items()}
    grads.update({"gw_" + n: p.grad.detach() for n, p in P.items()})
    return dict(decoded=out.detach(), ray_valid=ray_valid, weight=weight.detach(), conf=conf.detach()), grads


def maxdiff(a, b):
    return float((a.double() - b.double()).abs().max()) if a.numel() else 0.0


def golden_aggregator(ref, name, cfg, R, SR, K, seed, store_weights):
    from oracle import render_ref as rr
    inp = make_agg_inputs(cfg, R, SR, K, seed)
    g = torch.Generator().manual_seed(seed + 100)
    cot = dict(decoded=torch.randn(1, R, SR, 4, generator=g), conf=torch.randn(1, R, SR, K, generator=g) * 0.1)
    P_state = None if store_weights else rr.init_params(cfg, seed=seed, bias_scale=0.1)
    P_state, ref_out, ref_grads = run_reference_aggregator(ref, cfg, P_state, inp, cot)
    orc_out, orc_grads = run_oracle_aggregator(cfg, P_state, inp, cot)
    worst = 0.0
    for k in ref_out:
        d = maxdiff(ref_out[k].float(), orc_out[k].float()); worst = max(worst, d)
    for k in ref_grads:
        scale = float(ref_grads[k].abs().max()) + 1e-12
        d = maxdiff(ref_grads[k], orc_grads[k]) / scale; worst = max(worst, d)
    print(f"[{name}] oracle vs reference: worst abs/rel diff {worst:.3e}")
    assert worst < 2e-5, worst
    save = {"in_" + k: v.numpy() for k, v in inp.items() if v is not None}
    save.update({"cot_" + k: v.numpy() for k, v in cot.items()})
    save.update({"out_" + k: v.numpy() for k, v in ref_out.items()})
    # weight grads are large for the canonical width: store them only for the small config, and a
    # per-tensor (sum, abs-sum) signature otherwise
    for k, v in ref_grads.items():
        if store_weights or not k.startswith("gw_"):
            save[k] = v.numpy()
        else:
            save["sig_" + k] = np.array([float(v.double().sum()), float(v.double().abs().sum())])
    if store_weights:
        save.update({"P_" + k: v.numpy() for k, v in P_state.items()})
    else:
        save["P_seed"] = np.array(seed)
        save["P_signature"] = np.array([float(sum(v.double().sum() for v in P_state.values())),
                                        float(sum(v.double().abs().sum() for v in P_state.values()))])
    save["cfg"] = np.array(repr(vars(cfg)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)


def golden_ray_march(ref):
    g = torch.Generator().manual_seed(7)
    B, R, SR = 1, 37, 24
    feats = torch.rand(B, R, SR, 4, generator=g)
    feats[..., 0] = feats[..., 0] * 40.0
    valid = torch.rand(B, R, SR, generator=g) > 0.4
    valid[:, 0] = False
    valid[:, 1] = True
    dist = torch.rand(B, R, SR, generator=g) * 0.016 * valid.float()
    bg = torch.tensor([1.0, 1.0, 1.0])
    save = dict(feats=feats.numpy(), valid=valid.numpy(), dist=dist.numpy(), bg=bg.numpy())
    from oracle import render_ref as rr
    for blend in ("alpha", "alpha2"):
        f = feats.clone().requires_grad_(True)
        out = ref.ray_march(dist, valid, f, ref.find_render_function("radiance"), ref.find_blend_function(blend), bg)
        cot_c = torch.randn(B, R, 3, generator=g)
        cot_o = torch.randn(B, R, SR, generator=g)
        ((out[0] * cot_c).sum() + (out[2] * cot_o).sum()).backward()
        names = ["ray_color", "point_color", "opacity", "acc_transmission", "blend_weight", "bg_transmission", "bg_blend_weight"]
        mine = rr.ray_march(dist, valid, feats, bg, blend)
        for n, a, b in zip(names, out, mine):
            assert maxdiff(a.detach(), b) < 1e-6, (n, maxdiff(a.detach(), b))
            save[f"{blend}_{n}"] = a.detach().numpy()
        save[f"{blend}_cot_color"] = cot_c.numpy()
        save[f"{blend}_cot_opacity"] = cot_o.numpy()
        save[f"{blend}_grad_feats"] = f.grad.numpy()
        a5 = ref.alpha_ray_march(dist, valid, feats, ref.find_blend_function(blend))
        m5 = rr.alpha_ray_march(dist, valid, feats, blend)
        for a, b in zip(a5, m5):
            assert maxdiff(a, b) < 1e-6
    np.savez_compressed(os.path.join(HERE, "ray_march.npz"), **save)
    print("[ray_march] oracle vs reference OK")


def golden_pe_and_rays(ref):
    from oracle import render_ref as rr
    from oracle import query_ref as qr
    g = torch.Generator().manual_seed(11)
    x = torch.randn(5, 7, 3, generator=g)
    save = dict(pe_x=x.numpy())
    for F, ori in ((3, False), (5, False), (4, True)):
        a = ref.positional_encoding(x, F, ori=ori)
        assert torch.equal(a, rr.positional_encoding(x, F, ori=ori))
        save[f"pe_{F}_{int(ori)}"] = a.numpy()
    campos = torch.tensor([[0.5, -1.0, 1.5]])
    raydir = torch.randn(1, 9, 3, generator=g)
    for jit in (0.0, 0.3):
        torch.manual_seed(5)
        raypos, seg, valid, mid = ref.near_far_linear_ray_generation(campos, raydir, 400, near=0.1, far=8.0, jitter=jit)
        torch.manual_seed(5)
        rp2, mid2 = qr.near_far_linear_ray_generation(campos, raydir, 400, 0.1, 8.0, jitter=jit)
        assert torch.equal(raypos, rp2) and torch.equal(mid, mid2)
        assert torch.equal(qr.raypos_from_t(campos, raydir, mid[0]), raypos)
        save[f"rays_mid_{jit}"] = mid.numpy()
        save[f"rays_pos_{jit}"] = raypos.numpy()
    save["rays_campos"], save["rays_dir"] = campos.numpy(), raydir.numpy()
    np.savez_compressed(os.path.join(HERE, "pe_rays.npz"), **save)
    print("[pe/rays] oracle vs reference bit-exact")


def main():
    ref = import_reference()
    from oracle import render_ref as rr
    golden_pe_and_rays(ref)
    golden_ray_march(ref)
    small = rr.agg_config(shading_feature_num=64)
    golden_aggregator(ref, "agg_small_plain", small, R=4, SR=10, K=8, seed=3, store_weights=True)
    small_sem = rr.semantic_config(shading_feature_num=64)
    golden_aggregator(ref, "agg_small_semantic", small_sem, R=4, SR=10, K=8, seed=4, store_weights=True)
    golden_aggregator(ref, "agg_canonical_plain", rr.agg_config(), R=3, SR=24, K=8, seed=5, store_weights=False)
    golden_aggregator(ref, "agg_canonical_semantic", rr.semantic_config(), R=3, SR=24, K=8, seed=6, store_weights=False)


if __name__ == "__main__":
    main()
