"""GPU: occupancy build + ray march + K-NN (through the C ABI) against the sequential oracle.
Bar: bit-exact -- ray mask, neighbour indices including slot order, sample positions, grid structures."""
import numpy as np
import pytest
import torch

from oracle import query_ref as qr
from sgnerf_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[1, 2], ids=["brick-walk", "warp-per-ray"])
def march_kernel_choice(request):
    """Every query test runs with each of sgn_query's two march kernels forced (they must give identical results)."""
    from sgnerf_b200 import _lib
    _lib.call("sgn_query_march_mode", request.param)
    yield
    _lib.call("sgn_query_march_mode", 0)


@pytest.fixture(scope="module")
def scene_c0():
    return synth.scene_c0(n_points=100_000, n_rays=1024)


@pytest.mark.parametrize("SR", [24, 80])
def test_c0_test_mode(scene_c0, SR):
    opt = qr.default_opt(SR=SR)
    t = util.shared_t(scene_c0.near, scene_c0.far, opt.z_depth_dim)
    cu = util.cuda_query(scene_c0, opt, t)
    orc = util.oracle_query(scene_c0, opt, t)
    util.assert_query_equal(cu, orc)
    util.assert_grid_equal(cu.grid, orc[-1].grid)
    assert int(cu.rmask.sum()) > 500          # the fixture actually exercises the path


def test_c0_train_mode_per_ray_jitter(scene_c0):
    opt = qr.default_opt(SR=24, is_train=1)
    t = util.jittered_t(scene_c0.near, scene_c0.far, opt.z_depth_dim, scene_c0.raydir.shape[0], seed=3)
    util.assert_query_equal(util.cuda_query(scene_c0, opt, t), util.oracle_query(scene_c0, opt, t))


@pytest.mark.parametrize("K", [4, 16])
def test_other_k(scene_c0, K):
    opt = qr.default_opt(SR=24, K=K)
    t = util.shared_t(scene_c0.near, scene_c0.far, opt.z_depth_dim)
    util.assert_query_equal(util.cuda_query(scene_c0, opt, t), util.oracle_query(scene_c0, opt, t))


def test_overflow_reservoirs_match_curand():
    """More voxels than max_o and more points per voxel than P: both cuRAND XORWOW reservoirs are hit, with
    non-zero `seconds` seeds.  Also pins the oracle's XORWOW restatement against the device generator."""
    s = synth.scene_c0(n_points=60_000, n_rays=512, seed=7)
    opt = qr.default_opt(SR=24, P=2, max_o=9000, vsize=[0.02, 0.02, 0.02])
    t = util.shared_t(s.near, s.far, opt.z_depth_dim)
    seconds = (1_700_000_123, 1_700_000_124, 1_700_000_125)
    cu = util.cuda_query(s, opt, t, seconds=seconds)
    orc = util.oracle_query(s, opt, t, seconds=seconds)
    g = orc[-1].grid
    assert int(g.occ_idx[0]) > opt.max_o and int(g.occ_numpnts.max()) > opt.P
    util.assert_grid_equal(cu.grid, g)
    util.assert_query_equal(cu, orc)


@pytest.mark.parametrize("sec", [1_700_000_001, 1_700_000_005])   # seconds % 10 <= 1 passes the gate, 5 does not
def test_semantic_guidance(scene_c0, sec):
    rng = np.random.default_rng(0)
    N, R = scene_c0.xyz.shape[0], scene_c0.raydir.shape[0]
    opt = qr.default_opt(SR=24, semantic_guidance=1)
    pt_label = rng.integers(0, 20, N).astype(np.int32)
    pt_prob = rng.random((N, 20)).astype(np.float32)     # the reference casts this to int32 before the kernel reads it as float
    ray_label = rng.integers(0, 20, R).astype(np.int32)
    t = util.shared_t(scene_c0.near, scene_c0.far, opt.z_depth_dim)
    kw = dict(ray_label=ray_label, points_label=pt_label, points_label_prob=pt_prob)
    cu = util.cuda_query(scene_c0, opt, t, seconds=(0, 0, sec), **kw)
    orc = util.oracle_query(scene_c0, opt, t, seconds=(0, 0, sec), **kw)
    util.assert_query_equal(cu, orc)


def test_edge_cases():
    opt = qr.default_opt(SR=24)
    s = synth.scene_c0(n_points=2000, n_rays=64, seed=5)
    t = util.shared_t(s.near, s.far, opt.z_depth_dim)
    # camera looking away from everything: no ray hits
    away = synth.scene_c0(n_points=2000, n_rays=64, seed=5)
    away.campos = np.array([50.0, 50.0, 50.0], np.float32)
    cu, orc = util.cuda_query(away, opt, t), util.oracle_query(away, opt, t)
    util.assert_query_equal(cu, orc)
    assert int(cu.rmask.sum()) == 0 and orc[0].shape[1] == 0
    # points outside `ranges` are dropped by both; a single in-range point
    far_pts = synth.scene_c0(n_points=2000, n_rays=64, seed=5)
    far_pts.xyz = far_pts.xyz.copy()
    far_pts.xyz[::2] += 40.0
    util.assert_query_equal(util.cuda_query(far_pts, opt, t), util.oracle_query(far_pts, opt, t))
    one = synth.scene_c0(n_points=2000, n_rays=64, seed=5)
    one.xyz = one.xyz[:1].copy()
    cu, orc = util.cuda_query(one, opt, t), util.oracle_query(one, opt, t)
    util.assert_query_equal(cu, orc)
    util.assert_grid_equal(cu.grid, orc[-1].grid)
    # duplicated points: ties in distance are broken by list order, identically on both sides
    dup = synth.scene_c0(n_points=4000, n_rays=256, seed=6)
    dup.xyz = np.ascontiguousarray(np.concatenate([dup.xyz[:2000], dup.xyz[:2000]]))
    util.assert_query_equal(util.cuda_query(dup, opt, t), util.oracle_query(dup, opt, t))


def test_full_size_properties():
    """BASELINE config C1 (1M points, 640x480) -- too large for the oracle, so check what the domain guarantees:
    determinism, every neighbour within radius and inside the 3^3 voxel block, and a brute-force K-NN on a sample."""
    s = synth.scene_room(1_000_000)
    opt = qr.default_opt(SR=24)
    t = util.shared_t(s.near, s.far, opt.z_depth_dim)
    g = util.cuda_grid(s, opt)
    a = util.cuda_query(s, opt, t, grid=g)
    b = util.cuda_query(s, opt, t)       # fresh build
    assert torch.equal(a.pidx, b.pidx) and torch.equal(a.loc_w, b.loc_w) and torch.equal(a.rmask, b.rmask)
    pidx, loc = a.pidx, a.loc_w
    valid = pidx >= 0
    assert torch.equal((valid.flatten(1).any(1)).to(torch.int8), a.rmask)
    assert bool(((a.smask > 0) | ~valid.any(-1)).all())
    xyz = torch.from_numpy(s.xyz).cuda()
    nb = xyz[pidx.clamp(min=0).long()]                                    # [R,SR,K,3]
    d2 = ((nb - loc[..., None, :]) ** 2).sum(-1)
    assert float(d2[valid].max()) <= float(a.hp.radius2) * (1 + 1e-5)
    origin = torch.from_numpy(a.hp.ranges[:3]).cuda()
    vs = torch.from_numpy(a.hp.scaled_vsize).cuda()
    cell_s = torch.floor((loc - origin) / vs)
    cell_p = torch.floor((nb - origin) / vs)
    assert bool(((cell_p - cell_s[..., None, :]).abs().amax(-1)[valid] <= 1).all())
    # brute force on 64 random valid samples: the chosen set is the K nearest, within radius, among the points the
    # reference keeps in the 3^3 block (voxels that own a slot > 0: beyond max_o voxels are dropped by the reservoir,
    # and slot 0 is emptied by the `> 0` guard)
    idx = torch.nonzero(valid.any(-1))
    pick = idx[torch.randperm(idx.shape[0], generator=torch.Generator().manual_seed(0))[:64]]
    cell_all = torch.floor((xyz - origin) / vs).long()
    dims = torch.tensor(a.grid.dim, device="cuda")
    lin = (cell_all[:, 0] * dims[1] + cell_all[:, 1]) * dims[2] + cell_all[:, 2]
    kept = a.grid.buffer(0)[lin] > 0
    counts = a.grid.buffer(3)
    checked = 0
    for r, sidx in pick.tolist():
        c = cell_s[r, sidx].long()
        inblk = ((cell_all - c).abs().amax(-1) <= 1) & kept
        if int(counts[a.grid.buffer(0)[lin[inblk]].long()].max()) > opt.P:
            continue                                                       # a capped voxel: reservoir picks, not a pure K-NN
        dd = ((xyz[inblk] - loc[r, sidx]) ** 2).sum(-1)
        own = ((cell_all[inblk] - c).abs().amax(-1) == 0)
        if int(own.sum()) >= opt.K:
            dd = dd[own]                                                   # layer-0 early exit keeps only own-voxel points
        dd = dd[dd <= float(a.hp.radius2)]
        got = d2[r, sidx][valid[r, sidx]]
        k = min(opt.K, dd.numel())
        assert got.numel() == k
        torch.testing.assert_close(torch.sort(got)[0], torch.sort(dd)[0][:k], rtol=1e-4, atol=1e-9)
        checked += 1
    assert checked > 32


def test_full_size_march_against_torch_bruteforce():
    """C1, all 307200 rays x 400 candidates: sample selection recomputed with plain torch ops (position = campos + raydir * t,
    voxel = floor((p - origin) / vsize) in fp32, occupancy bit, first SR by cumsum -- the reference's own formulation, :413-437 and
    :811-845) must equal the kernel's output bit for bit: the brick-mask skip and the reciprocal fast path may not change anything.
    Also: every occupied voxel's brick (and its neighbours within one voxel) is set in the brick mask."""
    s = synth.scene_room(1_000_000)
    opt = qr.default_opt(SR=24)
    t = util.shared_t(s.near, s.far, opt.z_depth_dim)
    a = util.cuda_query(s, opt, t)
    dims = a.grid.dim
    origin = torch.from_numpy(a.hp.ranges[:3]).cuda()
    vs = torch.from_numpy(a.hp.scaled_vsize).cuda()
    bits = a.grid.buffer(1)
    campos = torch.from_numpy(s.campos).cuda()
    raydir = torch.from_numpy(s.raydir).cuda()
    tt = t.cuda()
    SR = opt.SR
    for r0 in range(0, raydir.shape[0], 32768):
        rd = raydir[r0:r0 + 32768]
        pos = campos[None, None, :] + rd[:, None, :] * tt[None, :, None]                     # [r, D, 3]
        vox = torch.floor((pos - origin) / vs).long()
        inside = ((vox >= 0) & (vox < torch.tensor(dims, device="cuda"))).all(-1)
        lin = ((vox[..., 0] * dims[1] + vox[..., 1]) * dims[2] + vox[..., 2]).clamp(min=0, max=dims[0] * dims[1] * dims[2] - 1)
        occ = inside & (((bits[lin >> 5].long() >> (lin & 31)) & 1) > 0)
        order = torch.cumsum(occ.long(), dim=1)
        take = occ & (order <= SR)
        want_mask = torch.zeros(rd.shape[0], SR, dtype=torch.int32, device="cuda")
        want_loc = torch.zeros(rd.shape[0], SR, 3, device="cuda")
        rr_, dd_ = torch.nonzero(take, as_tuple=True)
        slot = order[rr_, dd_] - 1
        want_mask[rr_, slot] = 1
        want_loc[rr_, slot] = pos[rr_, dd_]
        assert torch.equal(a.smask[r0:r0 + 32768] > 0, want_mask > 0)
        assert torch.equal(a.loc_w[r0:r0 + 32768], want_loc)
        del pos, vox, inside, lin, occ, order, take
    # brick mask covers the dilated occupancy with a one-voxel margin
    coarse = a.grid.buffer(7)
    words = torch.nonzero(bits != 0)[:, 0]
    sub = words[torch.randperm(words.numel(), generator=torch.Generator().manual_seed(0))[:200000].cuda()]
    cdim = [(d + 7) // 8 for d in dims]
    for b in range(32):
        c = sub * 32 + b
        on = ((bits[sub].long() >> b) & 1) > 0
        c = c[on]
        z, y, x = c % dims[2], (c // dims[2]) % dims[1], c // (dims[1] * dims[2])
        for dx_ in (-1, 0, 1):
            for dy_ in (-1, 0, 1):
                for dz_ in (-1, 0, 1):
                    bx, by, bz = ((x + dx_).clamp(0, dims[0] - 1)) >> 3, ((y + dy_).clamp(0, dims[1] - 1)) >> 3, ((z + dz_).clamp(0, dims[2] - 1)) >> 3
                    bc = (bx * cdim[1] + by) * cdim[2] + bz
                    assert bool((((coarse[bc >> 5].long() >> (bc & 31)) & 1) > 0).all())
        if b >= 3:
            break


def test_grid_without_neighbour_lists_gives_the_same_query(scene_c0):
    """SGN_GRID_NO_NEIGHBOUR_LISTS (grids of clouds that are edited between frames): the K-NN kernel walks the brick index itself instead
    of the prebuilt per-voxel lists -- identical outputs, slot order included, also under both overflows."""
    from sgnerf_b200 import ops
    for opt, seconds in ((qr.default_opt(SR=24), (0, 0, 0)), (qr.default_opt(SR=24, P=2, max_o=3000, vsize=[0.02, 0.02, 0.02]), (11, 22, 0))):
        t = util.shared_t(scene_c0.near, scene_c0.far, opt.z_depth_dim)
        a = util.cuda_query(scene_c0, opt, t, seconds=seconds)
        xyz = torch.from_numpy(scene_c0.xyz).cuda()
        hp = a.hp
        g2 = ops.OccGrid(xyz, hp.ranges[:3], hp.scaled_vsize, hp.scaled_vdim, opt.query_size, opt.P, opt.max_o, seconds_claim=seconds[0],
                         seconds_fill=seconds[1], neighbour_lists=False)
        b = util.cuda_query(scene_c0, opt, t, seconds=seconds, grid=(g2, hp))
        assert torch.equal(a.pidx, b.pidx) and torch.equal(a.loc_w, b.loc_w) and torch.equal(a.rmask, b.rmask) and torch.equal(a.smask, b.smask)
        assert int(a.rmask.sum()) > 100


@pytest.mark.parametrize("K", [8, 4])
def test_query_frame_leaves_only_empty_rows_unwritten_and_the_masked_aggregator_does_not_care(K):
    """sgn_query_frame (sparse_rows): identical to sgn_query on every slot with sample_mask > 0, rows of the other slots untouched (the
    output buffer is poisoned first); the aggregator given the mask renders the same from both."""
    from sgnerf_b200 import ops, pipeline
    s = synth.scene_c0(n_points=60_000, n_rays=512)
    opt = qr.default_opt(SR=24, K=K)
    grid, hp = util.cuda_grid(s, opt)
    t = util.shared_t(s.near, s.far, opt.z_depth_dim).cuda()
    campos, raydir = torch.from_numpy(s.campos).cuda(), torch.from_numpy(s.raydir).cuda()
    a = ops.query(grid, campos, raydir, t, opt.SR, K, opt.kernel_size[0], hp.radius2)
    torch.empty(512, opt.SR, K, dtype=torch.int32, device="cuda").fill_(123456789)      # poison what the allocator hands out next
    b = ops.query(grid, campos, raydir, t, opt.SR, K, opt.kernel_size[0], hp.radius2, sparse_rows=True)
    torch.cuda.synchronize()
    m = a[2] > 0
    assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and torch.equal(a[1], b[1])
    assert torch.equal(a[0][m], b[0][m]) and bool((a[0][~m] == -1).all())
    tabs = synth.make_point_tables(60_000, 32, 0, seed=0, conf_spread=0.2)
    from oracle import render_ref as rr
    from tests.test_gpu_aggregate import cfg_to_c
    cfg = rr.agg_config()
    P = rr.init_params(cfg, seed=1, bias_scale=0.05)
    names = [n for n, _, _ in rr.layer_shapes(cfg)]
    W, Bs = [P[n + ".weight"].cuda() for n in names], [P[n + ".bias"].cuda() for n in names]
    rot = torch.from_numpy(s.camrotc2w).cuda()
    args = (cfg_to_c(cfg), W, Bs, torch.from_numpy(s.xyz).cuda(), tabs.embedding.reshape(60_000, -1).cuda(), tabs.color.reshape(60_000, 3).cuda(),
            tabs.dir.reshape(60_000, 3).cuda(), tabs.conf.reshape(60_000).cuda(), None)
    with torch.no_grad():
        ra = ops.aggregate(*args, a[0], a[1], raydir, campos, rot, precision=ops.PRECISION_FP32, want_aux=True)
        poisoned = b[0].clone()
        poisoned[~m] = 123456789
        rb = ops.aggregate(*args, poisoned, b[1], raydir, campos, rot, precision=ops.PRECISION_FP32, want_aux=True, sample_mask=b[2])
    for x, y in zip(ra, rb):
        assert torch.equal(x, y)
