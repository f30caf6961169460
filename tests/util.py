"""Shared helpers for the parity tests: run the CUDA path and the oracle on the same seeded inputs."""
from types import SimpleNamespace

import numpy as np
import torch

from oracle import query_ref as qr


def shared_t(near, far, D):
    """middle_point_ts for jitter 0, computed once on the CPU with the reference's torch ops and fed to both sides."""
    _, mid = qr.near_far_linear_ray_generation(torch.zeros(1, 3), torch.zeros(1, 1, 3), D, near, far, jitter=0.0)
    return mid[0, 0].contiguous()


def jittered_t(near, far, D, R, seed):
    g = torch.Generator().manual_seed(seed)
    rand = torch.rand(1, R, D, generator=g)
    _, mid = qr.near_far_linear_ray_generation(torch.zeros(1, 3), torch.zeros(1, R, 3), D, near, far, jitter=0.3, rand=rand)
    return mid[0].contiguous()


def oracle_query(scene, opt, t, seconds=(0, 0, 0), ray_label=None, points_label=None, points_label_prob=None):
    xyz = torch.from_numpy(scene.xyz)[None]
    return qr.query_points(opt, xyz, scene.near, scene.far, torch.from_numpy(scene.raydir)[None],
                           torch.from_numpy(scene.campos)[None], torch.from_numpy(scene.camrotc2w)[None], t=t,
                           ray_label=ray_label, points_label=points_label, points_label_prob=points_label_prob,
                           seconds=seconds)


def cuda_grid(scene, opt, seconds=(0, 0)):
    from sgnerf_b200 import ops
    xyz = torch.from_numpy(scene.xyz).cuda()
    hp = ops.grid_hyperparameters(xyz, opt.vsize, opt.vscale, opt.kernel_size, opt.ranges, opt.radius_limit_scale)
    grid = ops.OccGrid(xyz, hp.ranges[:3], hp.scaled_vsize, hp.scaled_vdim, opt.query_size, opt.P, opt.max_o,
                       seconds_claim=seconds[0], seconds_fill=seconds[1])
    return grid, hp


def cuda_query(scene, opt, t, seconds=(0, 0, 0), ray_label=None, points_label=None, points_label_prob=None, grid=None):
    from sgnerf_b200 import ops
    if grid is None:
        grid, hp = cuda_grid(scene, opt, seconds[:2])
    else:
        grid, hp = grid
    kw = {}
    if ray_label is not None:
        kw = dict(ray_label=torch.as_tensor(ray_label).cuda(), pt_label=torch.as_tensor(points_label).to(torch.int32).cuda(),
                  pt_label_prob_bits=torch.as_tensor(points_label_prob).to(torch.int32).cuda())
    pidx, loc_w, smask, rmask = ops.query(grid, torch.from_numpy(scene.campos).cuda(), torch.from_numpy(scene.raydir).cuda(),
                                          t.cuda(), opt.SR, opt.K, opt.kernel_size[0], hp.radius2, seconds_query=seconds[2], **kw)
    torch.cuda.synchronize()
    return SimpleNamespace(pidx=pidx, loc_w=loc_w, smask=smask, rmask=rmask, grid=grid, hp=hp)


def assert_query_equal(cu, orc):
    """Exact equality with the sequential oracle: ray mask, neighbour indices INCLUDING slot order, sample positions bit for bit."""
    o_pidx, _, o_loc_w, _, o_ray_mask, _, o_ranges, info = orc
    rmask = cu.rmask.cpu().numpy()
    assert np.array_equal(rmask, o_ray_mask[0].numpy()), "ray_mask differs"
    sel = rmask > 0
    pidx = cu.pidx.cpu().numpy()
    assert np.array_equal(pidx[sel], o_pidx[0].numpy()), "sample_pidx differs"
    assert (pidx[~sel] == -1).all(), "rows of missed rays must be empty"
    loc = cu.loc_w.cpu().numpy()
    assert np.array_equal(loc[sel].view(np.int32), o_loc_w[0].numpy().view(np.int32)), "sample_loc_w differs (bitwise)"
    assert np.array_equal(cu.hp.ranges, o_ranges)
    assert np.array_equal(cu.hp.scaled_vdim, info.hp.scaled_vdim)
    assert cu.hp.radius2 == info.hp.radius2


def assert_grid_equal(grid, g):
    """Device grid vs the oracle's build_occ_vox arrays."""
    n_claimed = int(grid.buffer(6).cpu()[0])
    assert n_claimed == int(g.occ_idx[0])
    cell_slot = grid.buffer(0).cpu().numpy()
    assert np.array_equal(cell_slot, g.coor_2_occ), "coor_2_occ differs"
    bits = grid.buffer(1).cpu().numpy().view(np.uint32)
    occ = np.unpackbits(bits.view(np.uint8), bitorder="little")[:g.coor_occ.size]
    assert np.array_equal(occ.astype(np.int32), g.coor_occ), "coor_occ differs"
    n_rec = min(n_claimed, grid.max_o)
    assert np.array_equal(grid.buffer(2).cpu().numpy().reshape(-1, 3)[:n_rec], g.occ_2_coor[:n_rec]), "occ_2_coor differs"
    assert np.array_equal(grid.buffer(3).cpu().numpy(), g.occ_numpnts), "occ_numpnts differs"
    start = grid.buffer(4).cpu().numpy()
    cand = grid.buffer(5, torch.float32).cpu().numpy()
    cand_idx = cand[:, 3].copy().view(np.int32)
    ncap = np.minimum(g.occ_numpnts, grid.P)
    assert np.array_equal(np.diff(start), ncap)
    # per-slot lists, order included
    flat_ref = np.concatenate([g.occ_2_pnts[s, :ncap[s]] for s in np.nonzero(ncap)[0]]) if ncap.sum() else np.zeros(0, np.int32)
    assert np.array_equal(cand_idx[:start[-1]], flat_ref), "occ_2_pnts differs"


def pers_scene(n_points, n_rays, seed=1234, full_patch=False):
    """Scene C0 for the perspective querier: pixel indices and the points in the camera's perspective coordinates.  w2pers
    (neural_points.py:838-850) is restated with float32 numpy ELEMENTWISE operations in a fixed order, so the fixture is bit-identical on
    every machine (the golden vectors of the reference's kernels are re-checked on CPU against it)."""
    from sgnerf_b200 import synth
    s = synth.scene_c0(n_points=n_points, n_rays=n_rays, seed=seed)
    if full_patch:                                                   # a dense 64 x 48 pixel patch: several rays share one frustum column
        px, py = np.meshgrid(np.arange(200, 264), np.arange(150, 198))
        s.px, s.py = px.reshape(-1).astype(np.float32), py.reshape(-1).astype(np.float32)
    d = (s.xyz - s.campos[None]).astype(np.float32)
    R = s.camrotc2w.astype(np.float32)
    cam = [d[:, 0] * R[0, j] + d[:, 1] * R[1, j] + d[:, 2] * R[2, j] for j in range(3)]          # shift @ R  (xyz_c = R^T shift)
    s.xyz_pers = torch.from_numpy(np.ascontiguousarray(np.stack([cam[0] / cam[2], cam[1] / cam[2], cam[2]], -1).astype(np.float32)))
    s.pixel_idx = torch.from_numpy(np.stack([s.px, s.py], -1).astype(np.int32))
    return s
