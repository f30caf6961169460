"""GPU: pin the oracle (and the CUDA path) against the REFERENCE's OWN query kernels, compiled unchanged from
/root/reference into oracle/_ref/ and launched with the reference's geometry (tests/ref_driver.py).

The reference's slot numbering and list order depend on atomic arrival order, so equality is per-sample *sorted sets*
on an overflow-free fixture, and samples whose 3^3 block touches a slot-0 voxel (the voxel the `voxel_idx > 0` bug
empties -- a different voxel on every run of the reference) are compared separately.  The vectors of one run are
written to gpurun_out/ so they can be committed under tests/golden/ and re-checked on CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import query_ref as qr
from sgnerf_b200 import synth
from tests import ref_driver, util

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[1, 2], ids=["brick-walk", "warp-per-ray"])
def march_kernel_choice(request):
    """Every query test runs with each of sgn_query's two march kernels forced (they must give identical results)."""
    from sgnerf_b200 import _lib
    _lib.call("sgn_query_march_mode", request.param)
    yield
    _lib.call("sgn_query_march_mode", 0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _near_voxel(loc_w, hp, coor):
    """samples whose voxel is within Chebyshev distance 1 of voxel `coor`"""
    c = np.floor((loc_w - hp.ranges[:3]) / hp.scaled_vsize)
    return (np.abs(c - np.asarray(coor, np.float32)).max(-1) <= 1)


@pytest.mark.skipif(not ref_driver.available(8), reason="oracle/_ref/libref_query_K8.so not built (needs /root/reference at build time)")
def test_reference_kernels_vs_oracle_and_cuda():
    s = synth.scene_c0(n_points=100_000, n_rays=1024)
    opt = qr.default_opt(SR=24)
    t = util.shared_t(s.near, s.far, opt.z_depth_dim)
    orc = util.oracle_query(s, opt, t)
    o_pidx, _, o_loc_w, _, o_ray_mask, _, _, info = orc
    hp = info.hp
    assert int(info.grid.occ_idx[0]) < opt.max_o and int(info.grid.occ_numpnts.max()) <= opt.P   # overflow-free

    L = ref_driver.lib(8)
    xyz = torch.from_numpy(s.xyz).cuda()[None]
    raypos = qr.raypos_from_t(torch.from_numpy(s.campos)[None], torch.from_numpy(s.raydir)[None], t).cuda()
    r_pidx, r_loc, r_mask, g = ref_driver.query_grid_point_index(L, raypos, xyz, opt, hp)
    torch.cuda.synchronize()

    # structures that do not depend on arrival order
    assert int(g["occ_idx"][0]) == int(info.grid.occ_idx[0])
    assert np.array_equal(g["coor_occ"].cpu().numpy().reshape(-1), info.grid.coor_occ)
    assert np.array_equal((g["coor_2_occ"].cpu().numpy().reshape(-1) >= 0), (info.grid.coor_2_occ >= 0))

    ref_slot0 = g["occ_2_coor"][0, 0].cpu().numpy()
    orc_slot0 = info.grid.occ_2_coor[0]
    # rays: identical masks except rays that only see slot-0-affected samples
    r_mask_np, o_mask_np = r_mask[0].cpu().numpy(), o_ray_mask[0].numpy()
    both = (r_mask_np > 0) & (o_mask_np > 0)
    assert (r_mask_np != o_mask_np).sum() <= 4
    ridx = np.cumsum(r_mask_np > 0) - 1
    oidx = np.cumsum(o_mask_np > 0) - 1
    rp, op_ = r_pidx[0].cpu().numpy()[ridx[both]], o_pidx[0].numpy()[oidx[both]]
    rl, ol = r_loc[0].cpu().numpy()[ridx[both]], o_loc_w[0].numpy()[oidx[both]]
    assert np.array_equal(rl.view(np.int32), ol.view(np.int32)), "shading sample positions differ from the reference kernels"
    touched = _near_voxel(ol, hp, ref_slot0) | _near_voxel(ol, hp, orc_slot0)
    same = (np.sort(rp, -1) == np.sort(op_, -1)).all(-1)
    assert same[~touched].all(), "neighbour sets differ from the reference kernels away from the slot-0 voxels"
    assert (~touched).mean() > 0.95
    # the CUDA path equals the oracle exactly (order included), hence the reference as sets
    cu = util.cuda_query(s, opt, t)
    util.assert_query_equal(cu, orc)

    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    keep = np.nonzero(both)[0][:128]                      # a small slice of rays is enough for the CPU re-check
    sel = np.isin(np.nonzero(both)[0], keep)
    np.savez_compressed(os.path.join(out, "query_reference_kernels_c0.npz"),
                        ray_ids=keep.astype(np.int32), ref_pidx_sorted=np.sort(rp[sel], -1), ref_loc_w=rl[sel],
                        ref_slot0=ref_slot0, ref_ray_mask=r_mask_np, ref_occ_idx=g["occ_idx"].cpu().numpy(),
                        n_points=np.array(100_000), n_rays=np.array(1024), SR=np.array(24))


@pytest.mark.skipif(not ref_driver.available(8), reason="oracle/_ref/libref_query_K8.so not built")
def test_reference_kernel_reservoir_matches_oracle_xorwow():
    """With one thread block the reference's claim order is still racy, but the *set* of surviving records under
    max_o overflow must have the reference's size, and the per-voxel P-cap keeps min(count, P) entries."""
    s = synth.scene_c0(n_points=20_000, n_rays=64, seed=9)
    opt = qr.default_opt(SR=24, P=2, max_o=3000, vsize=[0.02, 0.02, 0.02])
    xyz = torch.from_numpy(s.xyz)[None]
    hp = qr.get_hyperparameters(opt, xyz)
    g = ref_driver.build_occ_vox(ref_driver.lib(8), xyz.cuda(), opt, hp, seconds=(123, 456))
    og = qr.build_occ_vox(opt, hp, s.xyz, None, 123, 456)
    assert int(g["occ_idx"][0]) == int(og.occ_idx[0]) > opt.max_o
    # total number of points claimed by voxels is order independent only up to which voxels survive; check bounds
    cnt = g["occ_numpnts"][0].cpu().numpy()
    lists = g["occ_2_pnts"][0].cpu().numpy()
    assert ((lists >= 0).sum(-1) == np.minimum(cnt, opt.P)).all()


@pytest.mark.skipif(not ref_driver.available(8), reason="oracle/_ref/libref_query_K8.so not built")
@pytest.mark.parametrize("sec", [1_700_000_001, 1_700_000_005])          # seconds % 10 = 1 opens the cross-label gate, 5 closes it
def test_reference_semantic_kernels_vs_oracle_and_cuda(sec):
    """The semantic-guidance branch (get_shadingloc_with_semantic + query_neigh_along_ray_layered_semantic_guidance, :462-591, host side
    :853-938) of the reference's own compiled kernels against the oracle's restatement and the CUDA path.  The branch holds a reference
    quirk the restatement reproduces: the probabilities go through `.to(torch.int32)` (:916) and the kernel reads that tensor through a
    `const float*` (:547), so `label_prob = int(bits_as_float * 10)`; the int `1 - label_prob` is then compared with the unsigned long
    `seconds % 10`.  Point probabilities here are (a) what `.to(int32)` of softmax values gives (0, and 1 for an exact 1.0) and (b) raw
    int32 patterns that read back as 0.35f / 0.05f, which drive `1 - label_prob` negative (= a huge unsigned: always accepted) or to 1."""
    s = synth.scene_c0(n_points=100_000, n_rays=1024)
    rng = np.random.default_rng(3)
    N, R = s.xyz.shape[0], s.raydir.shape[0]
    opt = qr.default_opt(SR=24, semantic_guidance=1)
    pt_label = rng.integers(0, 20, N).astype(np.int32)                     # label 0 = "accept always" on either side
    prob = (rng.random((N, 20)) < 0.1).astype(np.int32)                    # .to(int32) of probabilities: 0, or 1 where p == 1.0
    wild = rng.random(N) < 0.2
    prob[wild] = np.where(rng.random((int(wild.sum()), 20)) < 0.5, np.float32(0.35).view(np.int32), np.float32(0.05).view(np.int32))
    ray_label = rng.integers(0, 20, R).astype(np.int32)
    t = util.shared_t(s.near, s.far, opt.z_depth_dim)
    kw = dict(ray_label=ray_label, points_label=pt_label, points_label_prob=prob)
    orc = util.oracle_query(s, opt, t, seconds=(0, 0, sec), **kw)
    o_pidx, _, o_loc_w, _, o_ray_mask, _, _, info = orc
    hp = info.hp
    L = ref_driver.lib(8)
    xyz = torch.from_numpy(s.xyz).cuda()[None]
    raypos = qr.raypos_from_t(torch.from_numpy(s.campos)[None], torch.from_numpy(s.raydir)[None], t).cuda()
    r_pidx, r_loc, r_mask, g = ref_driver.query_grid_point_index(
        L, raypos, xyz, opt, hp, seconds=(0, 0, sec), raylabel=torch.from_numpy(ray_label).cuda()[None],
        points_label=torch.from_numpy(pt_label).cuda(), points_label_prob=torch.from_numpy(prob).cuda())
    torch.cuda.synchronize()
    ref_slot0, orc_slot0 = g["occ_2_coor"][0, 0].cpu().numpy(), info.grid.occ_2_coor[0]
    r_mask_np, o_mask_np = r_mask[0].cpu().numpy(), o_ray_mask[0].numpy()
    both = (r_mask_np > 0) & (o_mask_np > 0)
    assert (r_mask_np != o_mask_np).sum() <= 4 and both.sum() > 500
    ridx, oidx = np.cumsum(r_mask_np > 0) - 1, np.cumsum(o_mask_np > 0) - 1
    rp, op_ = r_pidx[0].cpu().numpy()[ridx[both]], o_pidx[0].numpy()[oidx[both]]
    rl, ol = r_loc[0].cpu().numpy()[ridx[both]], o_loc_w[0].numpy()[oidx[both]]
    assert np.array_equal(rl.view(np.int32), ol.view(np.int32))
    touched = _near_voxel(ol, hp, ref_slot0) | _near_voxel(ol, hp, orc_slot0)
    same = (np.sort(rp, -1) == np.sort(op_, -1)).all(-1)
    assert same[~touched].all(), "semantic-guidance neighbour sets differ from the reference's own kernel"
    assert (~touched).mean() > 0.95
    # the gate must matter in this fixture: closed (seconds % 10 = 5) it rejects most cross-label neighbours, open it rejects none
    plain = util.oracle_query(s, qr.default_opt(SR=24), t)
    n_plain, n_sem = int((plain[0] >= 0).sum()), int((o_pidx >= 0).sum())
    assert (n_sem == n_plain) if sec % 10 <= 1 else (0 < n_sem < 0.5 * n_plain)
    cu = util.cuda_query(s, opt, t, seconds=(0, 0, sec), **kw)
    util.assert_query_equal(cu, orc)
