"""GPU: the voxel-grid helpers next to the render path against the oracle's torch restatement
(oracle/voxel_ref.py): point-cloud down-sampling (construct_vox_points_closest, mvs_utils.py:536-561) and the NN < 0 grid query
(NeuralPoints.query_vox_grid, neural_points.py:814-826).  Voxel sets, their order, the chosen points and the corner indices are exact;
centroids to fp32 rounding (the reference's scatter_mean adds with atomics in any order)."""
import numpy as np
import pytest
import torch

from oracle import voxel_ref as vr
from sgnerf_b200 import ops, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,vox_res", [(2, 10), (5000, 20), (100_000, 100), (200_000, 900)])
def test_voxel_downsample_vs_oracle(n, vox_res):
    xyz = torch.from_numpy(synth.make_room_cloud(n, (6.0, 5.0, 3.0), 6, 0.002, seed=n))
    if n > 10:
        xyz[7] = xyz[3]                                          # a duplicated point: a tie in the residual, first index wins
    c_ref, g_ref, m_ref, _, _ = vr.construct_vox_points_closest(xyz, vox_res)
    c, g, m = ops.voxel_downsample(xyz.cuda(), vox_res)
    assert np.array_equal(g.cpu().numpy(), g_ref.numpy()), "voxel set / order differs"
    torch.testing.assert_close(c.cpu(), c_ref, rtol=0, atol=2e-6)
    same = m.cpu() == m_ref
    # a different choice is only acceptable where two residuals are equal to rounding (centroid sums differ in the last bit)
    if not bool(same.all()):
        res = lambda idx: torch.norm(xyz[idx] - c_ref, dim=-1)
        assert float((res(m.cpu()) - res(m_ref)).abs().max()) <= 1e-6
    assert float(same.float().mean()) > 0.999
    assert g.dtype == torch.int32 and m.dtype == torch.int64 and c.shape == (g_ref.shape[0], 3)


def test_query_vox_grid_vs_oracle():
    g = torch.Generator().manual_seed(0)
    G = 24
    full = torch.full((G + 1, G + 1, G + 1), -1, dtype=torch.int32)
    occ = torch.rand(G + 1, G + 1, G + 1, generator=g) < 0.95
    full[occ] = torch.arange(int(occ.sum()), dtype=torch.int32)
    space_min = torch.tensor([-1.0, -0.5, 0.25])
    vsz = 0.1
    loc = space_min + (torch.rand(1, 300, 24, 3, generator=g) * 1.2 - 0.1) * (G * vsz)      # some samples outside the grid
    loc[0, 0, 0] = space_min + vsz * torch.tensor([3.0, 4.0, 5.0])                             # exactly on a grid point
    want = vr.query_vox_grid(loc.clone(), full.clone(), space_min, vsz, G)
    got = ops.query_vox_grid(loc.cuda(), full.cuda(), space_min, vsz, G)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), want)
    assert 0.05 < float((want[..., 0] >= 0).float().mean()) < 0.95
