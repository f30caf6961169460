"""GPU: the perspective-frustum querier (--wcoord_query 0, sgn_pers_query) against the sequential oracle (exact, slot order included) and
the oracle + CUDA path against the REFERENCE's own perspective kernels (oracle/_ref/libref_query_pers_K8.so, per-sample sorted sets: the
reference's list order follows its atomics)."""
import numpy as np
import pytest
import torch

from oracle import query_pers_ref as qp
from oracle import query_ref as qr
from sgnerf_b200 import ops, synth
from tests import ref_driver_pers

pytestmark = pytest.mark.gpu


def _scene(n_points, n_rays, seed=1234, full_patch=False):
    s = synth.scene_c0(n_points=n_points, n_rays=n_rays, seed=seed)
    if full_patch:                                                   # a dense 64 x 48 pixel patch: several rays share one frustum column
        px, py = np.meshgrid(np.arange(200, 264), np.arange(150, 198))
        s.px, s.py = px.reshape(-1).astype(np.float32), py.reshape(-1).astype(np.float32)
    xyz_pers = qr.w2pers(torch.from_numpy(s.xyz)[None], torch.from_numpy(s.camrotc2w)[None], torch.from_numpy(s.campos)[None])[0]
    s.xyz_pers = xyz_pers.contiguous()
    s.pixel_idx = torch.from_numpy(np.stack([s.px, s.py], -1).astype(np.int32))
    return s


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
