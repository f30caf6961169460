"""GPU: the perspective-frustum querier (--wcoord_query 0, sgn_pers_query) against the sequential oracle (exact, slot order included) and
the oracle + CUDA path against the REFERENCE's own perspective kernels (oracle/_ref/libref_query_pers_K8.so, per-sample sorted sets: the
reference's list order follows its atomics)."""
import numpy as np
import pytest
import torch

from oracle import query_pers_ref as qp
from sgnerf_b200 import ops, synth
from tests import ref_driver_pers, util

pytestmark = pytest.mark.gpu


_scene = util.pers_scene


def _cuda(s, opt, hp, seconds=(0, 0)):
    pidx, loc, mask = ops.pers_query(s.xyz_pers.cuda(), s.pixel_idx.cuda(), hp, opt.kernel_size, opt.query_size, opt.SR, opt.K, opt.P, NN=opt.NN,
                                     inverse=opt.inverse, seconds=seconds)
    torch.cuda.synchronize()
    return pidx.cpu().numpy(), loc.cpu().numpy(), mask.cpu().numpy()


CASES = {
    "canonical": dict(),                                                                                       # vscale 2, kernel 3^3, NN 2
    "wide": dict(vscale=[4, 4, 4], kernel_size=[5, 5, 3], query_size=[5, 5, 3], radius_limit_scale=16.0, depth_limit_scale=4.0),
    "nn1": dict(NN=1, vscale=[4, 4, 4], kernel_size=[5, 5, 3], query_size=[3, 3, 3], radius_limit_scale=16.0, depth_limit_scale=4.0),
    "p-cap": dict(P=2, vscale=[8, 8, 8], kernel_size=[3, 3, 3], radius_limit_scale=24.0, depth_limit_scale=8.0),   # reservoir of :399-405
    "rand": dict(NN=0, vscale=[4, 4, 4], kernel_size=[5, 5, 3], query_size=[3, 3, 3], radius_limit_scale=16.0, depth_limit_scale=4.0, K=4),
    "inverse": dict(inverse=1, vscale=[4, 4, 4], kernel_size=[5, 5, 3], query_size=[3, 3, 3], radius_limit_scale=16.0, depth_limit_scale=40.0),
    "k16-sr8": dict(K=16, SR=8, vscale=[4, 4, 4], kernel_size=[7, 7, 1], query_size=[3, 3, 1], radius_limit_scale=16.0, depth_limit_scale=4.0),
}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("patch", [False, True], ids=["random-pixels", "pixel-patch"])
def test_pers_query_equals_oracle(case, patch):
    opt = qp.default_opt(max_o=127, **CASES[case])
    s = _scene(400_000, 1024, full_patch=patch)
    hp = qp.get_hyperparameters(opt, s.height, s.width, synth.SCANNET_INTRINSIC, s.near, s.far)
    hp_c = ops.pers_hyperparameters(s.height, s.width, synth.SCANNET_INTRINSIC, s.near, s.far, opt.z_depth_dim, opt.vscale, opt.radius_limit_scale,
                                    opt.depth_limit_scale, opt.inverse)
    for k in ("ranges", "vsize", "scaled_vsize", "scaled_vdim", "ray_vsize"):
        assert np.array_equal(getattr(hp, k), getattr(hp_c, k)), k
    assert hp.radius2 == hp_c.radius2 and hp.depth2 == hp_c.depth2
    seconds = (1_700_000_001, 1_700_000_002)
    o_pidx, o_loc, o_mask = qp.query_uncompacted(opt, hp, s.pixel_idx, s.xyz_pers, seconds)
    c_pidx, c_loc, c_mask = _cuda(s, opt, hp_c, seconds)
    assert np.array_equal(c_mask, o_mask)
    assert o_mask.sum() > 100
    sel = o_mask > 0
    assert np.array_equal(c_loc[sel].view(np.int32), o_loc[sel].view(np.int32)), "sample positions differ (bitwise)"
    assert np.array_equal(c_pidx, o_pidx), f"{(c_pidx != o_pidx).any(-1).sum()} samples differ"
    assert (o_pidx >= 0).sum() > 1000, "fixture too sparse to mean anything"
    if case in ("wide", "p-cap", "k16-sr8"):
        assert ((o_pidx >= 0).sum(-1) == opt.K).any(), "no sample filled all K slots"


def test_pers_query_empty_and_tiny():
    opt = qp.default_opt(max_o=127)
    s = _scene(1000, 64)
    hp = ops.pers_hyperparameters(s.height, s.width, synth.SCANNET_INTRINSIC, s.near, s.far, opt.z_depth_dim, opt.vscale, opt.radius_limit_scale,
                                  opt.depth_limit_scale)
    pidx, loc, mask = ops.pers_query(s.xyz_pers.cuda(), s.pixel_idx[:0].cuda(), hp, opt.kernel_size, opt.query_size, opt.SR, opt.K, opt.P)
    assert pidx.shape == (0, opt.SR, opt.K) and mask.shape == (0,)
    # every point behind the camera / outside the frustum: nothing occupied
    far_pts = torch.full((100, 3), 50.0)
    pidx, loc, mask = ops.pers_query(far_pts.cuda(), s.pixel_idx.cuda(), hp, opt.kernel_size, opt.query_size, opt.SR, opt.K, opt.P)
    assert int(mask.sum()) == 0 and int((pidx >= 0).sum()) == 0
    o = qp.query_uncompacted(opt, hp, s.pixel_idx, s.xyz_pers)
    c = _cuda(s, opt, hp)
    assert np.array_equal(c[0], o[0]) and np.array_equal(c[2], o[2])


@pytest.mark.skipif(not ref_driver_pers.available(8), reason="oracle/_ref/libref_query_pers_K8.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", ["canonical", "wide", "nn1", "inverse"])
def test_reference_perspective_kernels_vs_oracle_and_cuda(case):
    """The reference's kernels, launched with its own geometry and torch glue, give the oracle's neighbours: same ray mask, same sample
    positions bit for bit, same neighbour SETS per sample (P large enough that no voxel list overflows; ties in distance are the only
    order-dependent outcome left and are counted)."""
    opt = qp.default_opt(max_o=127, P=64, **CASES[case])
    s = _scene(400_000, 1024)
    hp = qp.get_hyperparameters(opt, s.height, s.width, synth.SCANNET_INTRINSIC, s.near, s.far)
    L = ref_driver_pers.lib(8)
    r_pidx, r_loc, r_pix, r_mask, info = ref_driver_pers.query_grid_point_index(L, s.pixel_idx.cuda()[None], s.xyz_pers.cuda()[None], opt, hp)
    assert info["max_selected_per_column"] <= 127 and info["max_points_per_voxel"] <= opt.P            # inside the documented envelope
    o_pidx, o_loc, o_mask = qp.query_uncompacted(opt, hp, s.pixel_idx, s.xyz_pers)
    c_pidx, c_loc, c_mask = _cuda(s, opt, hp)
    r_mask = r_mask[0].cpu().numpy()
    assert np.array_equal(r_mask, o_mask) and np.array_equal(c_mask, o_mask)
    sel = o_mask > 0
    r_pidx, r_loc = r_pidx[0].cpu().numpy(), r_loc[0].cpu().numpy()
    assert np.array_equal(r_loc.view(np.int32), o_loc[sel].view(np.int32)), "sample positions differ from the reference kernels (bitwise)"
    rs, os_ = np.sort(r_pidx, -1), np.sort(o_pidx[sel], -1)
    differ = (rs != os_).any(-1)
    assert (os_ >= 0).sum() > 1000
    assert differ.mean() < 1e-3, f"{differ.sum()} of {differ.size} samples have different neighbour sets"
    assert np.array_equal(np.sort(c_pidx[sel], -1), os_)
    if case == "wide":                                    # vectors of the reference's kernels for the CPU re-check (tests/test_oracle_golden.py)
        import os
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        keep = np.arange(0, rs.shape[0], 4)
        np.savez_compressed(os.path.join(out, "pers_reference_kernels_c0.npz"), ray_rows=keep.astype(np.int32), ref_pidx_sorted=rs[keep],
                            ref_loc=r_loc[keep], ref_ray_mask=r_mask, n_points=np.array(400_000), n_rays=np.array(1024), P=np.array(opt.P),
                            samples_with_other_sets=np.nonzero(differ.reshape(-1))[0].astype(np.int32))


@pytest.mark.parametrize("train", [False, True])
def test_neural_points_with_perspective_querier_renders_like_the_oracle(train):
    """--wcoord_query 0 through the reference-shaped modules: NeuralPoints (perspective querier) -> PointAggregator -> ray_march against the
    oracle's querier (query_points, compacted) + the oracle's fp32 render of its outputs."""
    from types import SimpleNamespace

    from oracle import render_ref as rr
    from sgnerf_b200 import modules
    from tests.test_modules import make_opt
    dev = "cuda"
    kw = dict(vscale=[4, 4, 4], kernel_size=[5, 5, 3], query_size=[5, 5, 3], radius_limit_scale=16.0, depth_limit_scale=4.0)
    s = _scene(400_000, 600)
    opt = make_opt(wcoord_query=0, P=16, sgn_seconds=7, shpnt_jitter="uniform", is_train=0, **kw)
    cfg = rr.agg_config()
    P = rr.init_params(cfg, seed=0, bias_scale=0.05)
    tabs = synth.make_point_tables(s.xyz.shape[0], 32, 0, seed=0, conf_spread=0.2)
    npnts = modules.NeuralPoints(32, s.xyz.shape[0], opt, dev, feedforward=1)
    npnts.set_points(torch.from_numpy(s.xyz).to(dev), None, tabs.embedding.to(dev), points_color=tabs.color.to(dev), points_dir=tabs.dir.to(dev),
                     points_conf=tabs.conf.to(dev), parameter=True)
    assert isinstance(npnts.querier, modules.lighting_fast_querier_p)
    agg = modules.PointAggregator(opt).to(dev)
    agg.load_state_dict(P)
    if not train:
        agg.requires_grad_(False)
    campos, rot = torch.from_numpy(s.campos)[None].to(dev), torch.from_numpy(s.camrotc2w)[None].to(dev)
    intr = torch.from_numpy(synth.SCANNET_INTRINSIC)[None]
    inputs = {"pixel_idx": s.pixel_idx[None].to(dev), "camrotc2w": rot, "campos": campos, "near": torch.tensor([s.near]), "far": torch.tensor([s.far]),
              "h": torch.tensor([s.height]), "w": torch.tensor([s.width]), "intrinsic": intr, "raydir": torch.from_numpy(s.raydir)[None].to(dev),
              "pixel_label": None}
    with (torch.enable_grad() if train else torch.no_grad()):
        out = npnts(inputs)
        (sampled_color, sampled_label_embedding, sampled_Rw2c, sampled_dir, sampled_conf, sampled_embedding, sampled_xyz_pers, sampled_xyz,
         sample_pnt_mask, sample_loc, sample_loc_w, sample_ray_dirs, ray_mask_tensor, vsize, grid_vox_sz) = out
        decoded, ray_valid, weight, conf_coefficient = agg(*out[:12], vsize, grid_vox_sz)
        ray_dist = torch.cummax(sample_loc[..., 2], dim=-1)[0]
        ray_dist = torch.cat([ray_dist[..., 1:] - ray_dist[..., :-1], torch.full((1, ray_dist.shape[1], 1), float(vsize[2]), device=dev)], dim=-1)
        mask = torch.logical_or(ray_dist < 1e-8, ray_dist > 2 * float(vsize[2])).float()
        ray_dist = (ray_dist * (1.0 - mask) + mask * float(vsize[2])) * ray_valid.float()
        blend = lambda opacity, acc: opacity * acc
        blend.__name__ = "alpha_blend"
        render = lambda f: f[..., 1:4]
        render.__name__ = "radiance_render"
        ray_color = modules.ray_march(ray_dist, ray_valid, decoded, render, blend, torch.ones(1, 3, device=dev))[0]
    # the oracle's querier on the same inputs (compacted outputs, query_point_indices.py:76-93) ...
    o_opt = qp.default_opt(max_o=127, P=16, **kw)
    xyz_pers = npnts.w2pers(npnts.xyz.detach(), rot, campos).reshape(-1, 3).cpu()      # the points as the module's own (GPU) w2pers places them
    o = qp.query_points(o_opt, s.pixel_idx, xyz_pers, s.height, s.width, synth.SCANNET_INTRINSIC, s.near, s.far, torch.from_numpy(s.campos)[None],
                        torch.from_numpy(s.camrotc2w)[None], seconds=(7, 7))
    o_pidx, o_loc, o_loc_w, o_dirs, o_mask, o_vsize, _, _ = o
    assert np.array_equal(ray_mask_tensor[0].cpu().numpy(), o_mask[0].numpy())
    assert np.array_equal(sampled_embedding.ctx.pidx.cpu().numpy(), o_pidx.numpy())
    assert np.array_equal(sample_loc.cpu().numpy(), o_loc.numpy())
    torch.testing.assert_close(sample_loc_w.cpu(), o_loc_w, rtol=0, atol=2e-6)
    torch.testing.assert_close(sample_ray_dirs.cpu(), o_dirs, rtol=0, atol=1e-6)
    assert np.array_equal(np.asarray(vsize), o_vsize)
    # ... and its fp32 render
    tables = SimpleNamespace(xyz=torch.from_numpy(s.xyz), embedding=tabs.embedding, color=tabs.color, dir=tabs.dir, conf=tabs.conf, label_embedding=None)
    ref = rr.render_from_query(P, cfg, tables, o_pidx, o_loc, o_loc_w, o_dirs, o_mask, torch.from_numpy(s.camrotc2w)[None],
                               torch.from_numpy(s.campos)[None], o_vsize, torch.ones(3))
    sel = o_mask[0] > 0
    assert int((o_pidx >= 0).sum()) > 2000
    tol = 1e-3 if train else 1e-2          # TF32 layer-wise training path / bf16 inference kernels, as in tests/test_modules.py
    torch.testing.assert_close(ray_color[0].detach().cpu(), ref.coarse_raycolor[0][sel], rtol=0, atol=tol)
    if train:
        (ray_color ** 2).mean().backward()
        assert float(npnts.points_embeding.grad.abs().sum()) > 0
