"""GPU: compositing / step-size glue / fill_invalid through the C ABI vs the oracle and the reference's golden vectors.
Tolerance: fp32, |delta| <= 1e-5 (the warp scan multiplies in a different association than torch.cumprod)."""
import os

import numpy as np
import pytest
import torch

from oracle import render_ref as rr
from sgnerf_b200 import ops

pytestmark = pytest.mark.gpu
ATOL = 1e-5


def _inputs(R, SR, seed, frac_valid=0.6):
    g = torch.Generator().manual_seed(seed)
    feats = torch.rand(1, R, SR, 4, generator=g)
    feats[..., 0] *= 60.0
    valid = torch.rand(1, R, SR, generator=g) < frac_valid
    if R > 2:
        valid[:, 0] = False
        valid[:, 1] = True
    dist = torch.rand(1, R, SR, generator=g) * 0.016 * valid.float()
    return feats, valid, dist


@pytest.mark.parametrize("blend", ["alpha", "alpha2"])
def test_golden_ray_march(golden_dir, blend):
    g = np.load(os.path.join(golden_dir, "ray_march.npz"))
    feats = torch.from_numpy(g["feats"]).cuda().requires_grad_(True)
    out = ops.composite(feats, torch.from_numpy(g["dist"]).cuda(), torch.from_numpy(g["valid"]).cuda(),
                        torch.from_numpy(g["bg"]).cuda(), blend=0 if blend == "alpha" else 1)
    ray_color, opacity, acc, bw, bgt = out
    for name, a in (("ray_color", ray_color), ("opacity", opacity), ("acc_transmission", acc)):
        np.testing.assert_allclose(a.detach().cpu().numpy(), g[f"{blend}_{name}"], rtol=0, atol=ATOL, err_msg=name)
    np.testing.assert_allclose(bw.detach().cpu().numpy(), g[f"{blend}_blend_weight"][..., 0], rtol=0, atol=ATOL)
    np.testing.assert_allclose(bgt.detach().cpu().numpy(), g[f"{blend}_bg_transmission"][..., 0], rtol=0, atol=ATOL)
    ((ray_color * torch.from_numpy(g[f"{blend}_cot_color"]).cuda()).sum()
     + (opacity * torch.from_numpy(g[f"{blend}_cot_opacity"]).cuda()).sum()).backward()
    ref = g[f"{blend}_grad_feats"]
    np.testing.assert_allclose(feats.grad.cpu().numpy(), ref, rtol=1e-4, atol=1e-5 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("R,SR", [(1, 1), (5, 24), (33, 32), (7, 33), (3, 200), (1000, 24), (0, 24)])
@pytest.mark.parametrize("blend", [0, 1])
@pytest.mark.parametrize("use_bg", [True, False])
def test_forward_backward_vs_oracle(R, SR, blend, use_bg):
    feats, valid, dist = _inputs(R, SR, seed=R * 100 + SR)
    bg = torch.tensor([0.9, 0.5, 0.1]) if use_bg else None
    g = torch.Generator().manual_seed(1)
    cots = [torch.randn(1, R, 3, generator=g), torch.randn(1, R, SR, generator=g), torch.randn(1, R, SR, generator=g),
            torch.randn(1, R, generator=g)]
    f_ref = feats.clone().requires_grad_(True)
    r = rr.ray_march(dist, valid, f_ref, bg, "alpha" if blend == 0 else "alpha2")
    (r[0] * cots[0]).sum().add((r[2] * cots[1]).sum()).add((r[4][..., 0] * cots[2]).sum()).add((r[5][..., 0] * cots[3]).sum()).backward()
    f = feats.cuda().requires_grad_(True)
    out = ops.composite(f, dist.cuda(), valid.cuda(), None if bg is None else bg.cuda(), blend=blend)
    if R == 0:
        assert out[0].shape == (1, 0, 3)
        return
    for a, b in ((out[0], r[0]), (out[1], r[2]), (out[2], r[3]), (out[3], r[4][..., 0]), (out[4], r[5][..., 0])):
        torch.testing.assert_close(a.detach().cpu(), b.detach(), rtol=0, atol=ATOL)
    (out[0] * cots[0].cuda()).sum().add((out[1] * cots[1].cuda()).sum()).add((out[3] * cots[2].cuda()).sum()).add((out[4] * cots[3].cuda()).sum()).backward()
    scale = max(1.0, float(f_ref.grad.abs().max()))
    torch.testing.assert_close(f.grad.cpu(), f_ref.grad, rtol=1e-4, atol=2e-5 * scale)


@pytest.mark.parametrize("mode_unit", [0, 1])
def test_ray_dist(mode_unit):
    g = torch.Generator().manual_seed(2)
    R, SR = 57, 24
    loc = torch.rand(1, R, SR, 3, generator=g)
    loc[..., 2] = torch.cumsum(torch.rand(1, R, SR, generator=g) * 0.02, dim=-1)
    loc[0, 3, 5:, 2] = 0.0                      # unused slots sit at world origin -> non-monotone depth
    loc[0, 4, 7, 2] = loc[0, 4, 6, 2]           # zero step
    valid = torch.rand(1, R, SR, generator=g) < 0.7
    ref = rr.ray_dist_from_samples(loc, valid, 0.008, mode_unit)
    out = ops.ray_dist(loc.cuda(), valid.cuda(), 0.008, mode_unit)
    torch.testing.assert_close(out.cpu(), ref, rtol=0, atol=1e-7)


def test_fill_invalid():
    g = torch.Generator().manual_seed(3)
    R, SR = 100, 24
    mask = (torch.rand(1, R, generator=g) < 0.5)
    mask[0, 0], mask[0, 1] = True, False
    n = int(mask.sum())
    color, op, bgt = torch.rand(1, n, 3, generator=g), torch.rand(1, n, SR, generator=g), torch.rand(1, n, 1, generator=g)
    bg = torch.tensor([1.0, 1.0, 1.0])
    ref_c, ref_o, ref_b = rr.fill_invalid(mask, color, op, bgt, bg)
    full_c, full_o, full_b = torch.zeros(R, 3), torch.zeros(R, SR), torch.zeros(R)
    full_c[mask[0]], full_o[mask[0]], full_b[mask[0]] = color[0], op[0], bgt[0, :, 0]
    full_c, full_o, full_b = full_c.cuda(), full_o.cuda(), full_b.cuda()
    ops.fill_invalid(mask[0].to(torch.int8).cuda(), bg.cuda(), full_c, full_o, full_b)
    torch.testing.assert_close(full_c.cpu(), ref_c[0]); torch.testing.assert_close(full_o.cpu(), ref_o[0])
    torch.testing.assert_close(full_b.cpu(), ref_b[0, :, 0])


def test_full_size_properties():
    """C1-sized input: transmittance bookkeeping must close (sum of blend weights + background transmission = 1)."""
    R, SR = 640 * 480, 24
    g = torch.Generator(device="cuda").manual_seed(0)
    feats = torch.rand(R, SR, 4, device="cuda", generator=g)
    feats[..., 0] *= 100
    valid = torch.rand(R, SR, device="cuda", generator=g) < 0.3
    dist = torch.full((R, SR), 0.008, device="cuda") * valid
    color, opacity, acc, bw, bgt = ops.composite(feats, dist, valid, torch.ones(3, device="cuda"))
    torch.testing.assert_close(bw.sum(-1) + bgt, torch.ones(R, device="cuda"), rtol=0, atol=1e-5)
    assert float(color.min()) >= 0 and float(color.max()) <= 1 + 1e-5
    assert bool((opacity[~valid] == 0).all())


@pytest.mark.parametrize("R,SR", [(1, 1), (57, 24), (40, 33), (9, 200), (640 * 480, 24)])
@pytest.mark.parametrize("blend", [0, 1])
def test_fused_frame_tail_equals_three_kernels(R, SR, blend):
    """sgn_render_composite (ray_dist + composite + fill_invalid in one pass) against the three kernels: the same step sizes and opacities
    bit for bit; colour and background transmittance to rounding (the fused kernel multiplies the transmittance in sample order for
    SR <= 40, the separate kernel by a warp scan)."""
    g = torch.Generator(device="cuda").manual_seed(4)
    dec = torch.rand(R, SR, 4, device="cuda", generator=g)
    dec[..., 0] *= 80.0
    loc = torch.rand(R, SR, 3, device="cuda", generator=g)
    loc[..., 2] = torch.cumsum(torch.rand(R, SR, device="cuda", generator=g) * 0.02, dim=-1)
    if R > 8:
        loc[3, SR // 3:, 2] = 0.0                # unused slots at the origin -> non-monotone depth
        loc[4, SR // 2, 2] = loc[4, SR // 2 - 1, 2]
    valid = torch.rand(R, SR, device="cuda", generator=g) < 0.5
    mask = valid.any(-1).to(torch.int8)
    if R > 8:
        mask[5] = 0                              # a missed ray whose rows hold junk must still come out as background
    bg = torch.tensor([1.0, 0.5, 0.25], device="cuda")
    rd = ops.ray_dist(loc, valid, 0.008, 1)
    color, opacity, _, _, bgt = ops.composite(dec, rd, valid, bg, blend=blend)
    ops.fill_invalid(mask, bg, color, opacity, bgt)
    color, opacity, acc, _, bgt = ops.composite(dec, rd, valid, bg, blend=blend)
    w_alpha = opacity * acc
    want_depth = (w_alpha * loc[..., 2]).sum(-1) / (w_alpha.sum(-1) + 1e-6) * (mask > 0)
    ops.fill_invalid(mask, bg, color, opacity, bgt)
    f_color, f_opacity, f_bgt, f_depth = ops.render_composite(dec, loc, valid, mask, 0.008, bg, blend=blend)
    assert torch.equal(f_opacity, opacity)
    torch.testing.assert_close(f_color, color, rtol=0, atol=2e-6)
    torch.testing.assert_close(f_bgt, bgt, rtol=0, atol=2e-6)
    torch.testing.assert_close(f_depth, want_depth, rtol=0, atol=1e-5)
    # the dense-depth entry (sgn_render_composite_depth) is the same computation on loc[..., 2]
    d_color, d_opacity, d_bgt, d_depth = ops.render_composite(dec, loc[..., 2].contiguous(), valid, mask, 0.008, bg, blend=blend, depth_array=True)
    assert torch.equal(d_color, f_color) and torch.equal(d_opacity, f_opacity) and torch.equal(d_bgt, f_bgt) and torch.equal(d_depth, f_depth)
