"""GPU: the drop-in boundary under the reference's REAL caller.

models/neural_points_volumetric_model.py is imported UNMODIFIED from baseline/_ref (staged by baseline/stage_reference.py) with
INTEGRATION.md's three-import swap applied from outside (tests/ref_import.py): its own option parser
(NeuralPointsVolumetricModel.modify_commandline_options -> the swapped classes' flag tables), its NeuralPointsRayMarching.forward
(:435-671, including the torch step-size glue and the `opt.prob == 1` block that indexes the gathered tensors) and its fill_invalid
(:158-195) run on sgnerf_b200's NeuralPoints / PointAggregator / ray_march.  coarse_raycolor, opacity and the prob outputs are compared
with the oracle (C query restatement + torch render restatement) on the same scene.
"""
from types import SimpleNamespace

import argparse
import numpy as np
import pytest
import torch

from oracle import query_ref as qr
from oracle import render_ref as rr
from tests import ref_import, util

pytestmark = pytest.mark.gpu

CANONICAL = ("--K 8 --SR 24 --P 26 --NN 2 --z_depth_dim 400 --vsize 0.008 0.008 0.008 --vscale 2 2 2 --kernel_size 3 3 3 --query_size 3 3 3 "
             "--radius_limit_scale 4 --depth_limit_scale 0 --max_o 610000 --ranges -10 -10 -10 10 10 10 --wcoord_query 1 --point_features_dim 32 "
             "--agg_distance_kernel linear --agg_dist_pers 20 --agg_intrp_order 2 --act_type LeakyReLU --num_feat_freqs 3 --dist_xyz_freq 5 "
             "--shading_feature_mlp_layer1 2 --shading_feature_mlp_layer3 2 --shading_alpha_mlp_layer 1 --shading_color_mlp_layer 4 "
             "--shading_feature_num 256 --point_conf_mode 1 --point_dir_mode 1 --point_color_mode 1 --num_viewdir_freqs 4 --num_pos_freqs 10 "
             "--raydist_mode_unit 1 --which_render_func radiance --which_blend_func alpha --which_tonemap_func off "
             "--zero_one_loss_items conf_coefficient --zero_one_loss_weights 0.0001 --bg_color white --color_grad 1 --dir_grad 1")


def _reference_options(mod, extra=""):
    """The reference's own parser for this model (its BaseRenderingModel flags + the swapped classes' flags), then its canonical
    ScanNet command line."""
    p = argparse.ArgumentParser()
    mod.NeuralPointsVolumetricModel.modify_commandline_options(p, is_train=True)
    opt, _ = p.parse_known_args((CANONICAL + " " + extra).split())
    opt.is_train, opt.sgn_precision = False, "fp32"
    return opt


@pytest.mark.skipif(not ref_import.available(), reason="baseline/_ref not staged")
@pytest.mark.parametrize("prob", [0, 1])
def test_reference_caller_on_swapped_modules_vs_oracle(prob):
    mod = ref_import.volumetric_model(swap=True)
    ref = ref_import.reference_modules()
    from sgnerf_b200 import modules, synth
    assert mod.NeuralPoints is modules.NeuralPoints and mod.PointAggregator is modules.PointAggregator
    dev = "cuda"
    opt = _reference_options(mod, f"--prob {prob}")
    assert opt.K == 8 and opt.prob == prob and opt.zero_epsilon == 1e-3
    s = synth.scene_c0(n_points=20_000, n_rays=300)
    cfg = rr.agg_config()
    P = rr.init_params(cfg, seed=0, bias_scale=0.05)
    tabs = synth.make_point_tables(s.xyz.shape[0], 32, 0, seed=0, conf_spread=0.2)
    npnts = mod.NeuralPoints(opt.point_features_dim, s.xyz.shape[0], opt, dev, feedforward=1)
    npnts.set_points(torch.from_numpy(s.xyz).to(dev), None, tabs.embedding.to(dev), points_color=tabs.color.to(dev), points_dir=tabs.dir.to(dev),
                     points_conf=tabs.conf.to(dev), parameter=True)
    agg = mod.PointAggregator(opt).to(dev)
    agg.load_state_dict(P)
    net = mod.NeuralPointsRayMarching(tonemap_func=ref.find_tone_map(opt.which_tonemap_func), render_func=ref.find_render_function(opt.which_render_func),
                                      blend_func=ref.find_blend_function(opt.which_blend_func), aggregator=agg, bpnet=None, is_compute_depth=False,
                                      neural_points=npnts, opt=opt, num_pos_freqs=opt.num_pos_freqs, num_viewdir_freqs=opt.num_viewdir_freqs)
    R = s.raydir.shape[0]
    inputs = {"campos": torch.from_numpy(s.campos)[None].to(dev), "raydir": torch.from_numpy(s.raydir)[None].to(dev),
              "camrotc2w": torch.from_numpy(s.camrotc2w)[None].to(dev), "bg_color": torch.ones(1, 3, device=dev),
              "pixel_idx": torch.zeros(1, R, 2, device=dev), "near": torch.tensor([[s.near]], device=dev), "far": torch.tensor([[s.far]], device=dev),
              "h": torch.tensor([480]), "w": torch.tensor([640]), "intrinsic": torch.eye(3, device=dev)[None],
              "gt_semantic_img": torch.zeros(1, 4, 4, 1, device=dev)}
    with torch.no_grad():
        output = net(inputs)                                                      # the reference's forward, line for line
        me = SimpleNamespace(input=inputs, opt=opt, tonemap_func=ref.find_tone_map("off"))
        me.unmask = lambda *a: mod.NeuralPointsVolumetricModel.unmask(me, *a)
        output = mod.NeuralPointsVolumetricModel.fill_invalid(me, output, inputs)   # and its fill_invalid
    # oracle on the same inputs
    t = util.shared_t(s.near, s.far, opt.z_depth_dim)
    o_pidx, o_loc, o_loc_w, o_dirs, o_mask, o_vsize, _, _ = util.oracle_query(s, qr.default_opt(SR=24), t)
    tables = SimpleNamespace(xyz=torch.from_numpy(s.xyz), embedding=tabs.embedding, color=tabs.color, dir=tabs.dir, conf=tabs.conf, label_embedding=None)
    want = rr.render_from_query(P, cfg, tables, o_pidx, o_loc, o_loc_w, o_dirs, o_mask, torch.from_numpy(s.camrotc2w)[None],
                                torch.from_numpy(s.campos)[None], o_vsize, torch.ones(3))
    assert np.array_equal(output["ray_mask"][0].cpu().numpy(), o_mask[0].numpy())
    assert int(o_mask.sum()) > 30
    torch.testing.assert_close(output["coarse_raycolor"].cpu(), want.coarse_raycolor, rtol=0, atol=1e-3)
    torch.testing.assert_close(output["coarse_point_opacity"].cpu(), want.coarse_point_opacity, rtol=0, atol=1e-3)
    torch.testing.assert_close(output["coarse_is_background"].cpu(), want.coarse_is_background, rtol=0, atol=1e-3)
    assert tuple(output["coarse_raycolor"].shape) == (1, R, 3) and tuple(output["queried_shading"].shape) == (1, R, 3)
    assert output["conf_coefficient"] is not None and tuple(output["weight"].shape) == (1, int(o_mask.sum()), opt.SR, opt.K)
    if prob == 1:
        gn = rr.gather_neighbors(tables, o_pidx, torch.from_numpy(s.camrotc2w)[None], torch.from_numpy(s.campos)[None])
        po = rr.probe_outputs(want.opacity, o_loc_w, want.weight, want.conf_coefficient, gn)
        sel = o_mask[0] > 0
        # the index of the largest opacity must be unambiguous for a tolerance comparison: keep rays whose top two differ clearly
        top2 = torch.topk(want.opacity[0], 2, dim=-1)[0]
        clear = (top2[:, 0] - top2[:, 1]) > 1e-3
        assert int(clear.sum()) > 10
        for k, v in po.items():
            got = output[k][0].cpu()[sel]                    # unmask() scattered the rows back to all R rays
            torch.testing.assert_close(got[clear], v[0][clear].reshape(got[clear].shape), rtol=0, atol=1e-3, msg=k)
            assert float(output[k][0].cpu()[~sel].abs().sum()) == 0.0


@pytest.mark.skipif(not ref_import.available(), reason="baseline/_ref not staged")
def test_swapped_aggregator_equals_the_reference_class_on_the_same_gathers():
    """The reference's own PointAggregator (torch, on cuda) fed with the dense tensors our NeuralPoints handles materialise, against our fused
    PointAggregator.forward on the handles: decoded / ray_valid / weight / conf_coefficient, fp32, 1e-3."""
    mod = ref_import.volumetric_model(swap=True)
    ref = ref_import.reference_modules()
    from sgnerf_b200 import synth
    dev = "cuda"
    opt = _reference_options(mod)
    opt.agg_axis_weight = None
    s = synth.scene_c0(n_points=20_000, n_rays=200)
    P = rr.init_params(rr.agg_config(), seed=0, bias_scale=0.05)
    tabs = synth.make_point_tables(s.xyz.shape[0], 32, 0, seed=0, conf_spread=0.2)
    npnts = mod.NeuralPoints(32, s.xyz.shape[0], opt, dev, feedforward=1)
    npnts.set_points(torch.from_numpy(s.xyz).to(dev), None, tabs.embedding.to(dev), points_color=tabs.color.to(dev), points_dir=tabs.dir.to(dev),
                     points_conf=tabs.conf.to(dev), parameter=True)
    ours = mod.PointAggregator(opt).to(dev)
    ours.load_state_dict(P)
    theirs = ref.PointAggregator(opt).to(dev)
    theirs.load_state_dict(P)
    inputs = {"campos": torch.from_numpy(s.campos)[None].to(dev), "raydir": torch.from_numpy(s.raydir)[None].to(dev),
              "camrotc2w": torch.from_numpy(s.camrotc2w)[None].to(dev), "near": torch.tensor([[s.near]]), "far": torch.tensor([[s.far]]),
              "pixel_idx": None, "h": None, "w": None, "intrinsic": None, "pixel_label": None}
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            tup = npnts(inputs)
            a = ours(*tup[:12], tup[13], tup[14])
            dense = [x.materialize() if hasattr(x, "materialize") else x for x in tup[:12]]
            b = theirs(*dense, tup[13], tup[14])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    torch.testing.assert_close(a[0], b[0], rtol=0, atol=1e-3)
    assert torch.equal(a[1], b[1])
    torch.testing.assert_close(a[2], b[2], rtol=0, atol=1e-5)
    torch.testing.assert_close(a[3], b[3], rtol=0, atol=1e-6)
