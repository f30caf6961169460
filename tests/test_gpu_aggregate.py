"""GPU: fused gather + aggregation (sgn_agg_forward / sgn_agg_backward through the C ABI) against
(a) the golden vectors produced by the reference's own PointAggregator and (b) the torch oracle on larger seeded inputs.

Tolerances (fp32 strict-parity mode): decoded (sigma, rgb), weights, conf <= 2e-4 abs (north_star: rgb within 1e-3);
gradients: relative L2 <= 1e-3 per tensor.
"""
import ast
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import render_ref as rr
from sgnerf_b200 import ops

pytestmark = pytest.mark.gpu
ATOL = 2e-4


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-12))


def cfg_to_c(cfg):
    return ops.agg_cfg(feat_dim=cfg.point_features_dim, num_feat_freqs=cfg.num_feat_freqs, dist_xyz_freq=cfg.dist_xyz_freq,
                       num_viewdir_freqs=cfg.num_viewdir_freqs, width=cfg.shading_feature_num,
                       n_block1=cfg.shading_feature_mlp_layer1, n_block2_bpnet=cfg.shading_feature_mlp_layer2_bpnet,
                       label_dim=cfg.label_embedding_dim, n_block3=cfg.shading_feature_mlp_layer3,
                       n_color=cfg.shading_color_mlp_layer, act_super=cfg.act_super, leaky_slope=cfg.leaky_slope)


def param_lists(P, cfg, device="cuda", requires_grad=False):
    names = [n for n, _, _ in rr.layer_shapes(cfg)]
    w = [P[n + ".weight"].detach().clone().to(device).requires_grad_(requires_grad) for n in names]
    b = [P[n + ".bias"].detach().clone().to(device).requires_grad_(requires_grad) for n in names]
    return names, w, b


@pytest.mark.parametrize("name", ["agg_small_plain", "agg_small_semantic", "agg_canonical_plain", "agg_canonical_semantic"])
def test_against_reference_golden(golden_dir, name):
    """Every (sample, slot) of the golden gathered tensors becomes its own point, so tables + indices reproduce the
    reference's gathered inputs exactly; outputs and gradients are then compared with the reference's."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = SimpleNamespace(**ast.literal_eval(str(g["cfg"])))
    if "P_seed" in g:
        P = rr.init_params(cfg, seed=int(g["P_seed"]), bias_scale=0.1)
    else:
        P = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("P_")}
    names, W, B = param_lists(P, cfg, requires_grad=True)
    inp = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in_")}
    mask = inp["sample_pnt_mask"]
    _, R, SR, K = mask.shape
    n = R * SR * K
    pidx = torch.where(mask.reshape(-1), torch.arange(n), torch.full((n,), -1)).reshape(R, SR, K).to(torch.int32).cuda()
    leaf = lambda t, c: t.reshape(n, c).clone().cuda().requires_grad_(True)
    emb, col, dr = leaf(inp["sampled_embedding"], cfg.point_features_dim), leaf(inp["sampled_color"], 3), leaf(inp["sampled_dir"], 3)
    conf = inp["sampled_conf"].reshape(n).clone().cuda().requires_grad_(True)
    lab = inp["sampled_label_embedding"].reshape(n, -1).cuda() if "sampled_label_embedding" in inp else None
    campos = torch.tensor([0.3, -0.2, -4.0]).cuda()          # the camera make_golden.py used for the perspective coords
    out = ops.aggregate(cfg_to_c(cfg), W, B, inp["sampled_xyz"].reshape(n, 3).cuda(), emb, col, dr, conf, lab, pidx,
                        inp["sample_loc_w"][0].cuda(), inp["sample_ray_dirs"][0, :, 0].cuda(), campos, torch.eye(3).cuda())
    decoded, ray_valid, loc_pers, weight, conf_coef = out
    np.testing.assert_allclose(decoded.detach().cpu().numpy(), g["out_decoded"][0], rtol=0, atol=ATOL)
    assert np.array_equal(ray_valid.cpu().numpy().astype(bool), g["out_ray_valid"][0])
    np.testing.assert_allclose(weight.detach().cpu().numpy(), g["out_weight"][0], rtol=0, atol=1e-5)
    np.testing.assert_allclose(loc_pers.cpu().numpy(), inp["sample_loc"][0].numpy(), rtol=1e-5, atol=1e-6)
    m = mask[0].numpy()
    np.testing.assert_allclose(conf_coef.detach().cpu().numpy()[m], g["out_conf"][0][m], rtol=0, atol=1e-6)
    # invalid slots read point 0, as the reference's clamp(pidx, 0) gather does
    c0 = float(np.clip(inp["sampled_conf"].reshape(-1)[0], 1e-4, 1.0))
    np.testing.assert_allclose(conf_coef.detach().cpu().numpy()[~m], c0, rtol=0, atol=1e-6)

    cot_d, cot_c = torch.from_numpy(g["cot_decoded"][0]).cuda(), torch.from_numpy(g["cot_conf"][0]).cuda()
    ((decoded * cot_d).sum() + (conf_coef * cot_c).sum()).backward()
    mflat = mask.reshape(-1)
    for t, key, c in ((emb, "g_sampled_embedding", cfg.point_features_dim), (col, "g_sampled_color", 3), (dr, "g_sampled_dir", 3)):
        ref = torch.from_numpy(g[key]).reshape(n, c) * mflat[:, None]
        assert rel_l2(t.grad.cpu(), ref) < 1e-3, key
    ref_conf = torch.from_numpy(g["g_sampled_conf"]).reshape(n) * mflat
    ref_conf[0] += torch.from_numpy(g["cot_conf"]).reshape(n)[~mflat].sum()
    assert rel_l2(conf.grad.cpu(), ref_conf) < 1e-3
    for nme, w, b in zip(names, W, B):
        for suffix, t in ((".weight", w), (".bias", b)):
            key = "gw_" + nme + suffix
            if key in g:
                assert rel_l2(t.grad.cpu(), torch.from_numpy(g[key])) < 1e-3, key
            else:
                s = g["sig_" + key]
                got = np.array([float(t.grad.double().sum()), float(t.grad.double().abs().sum())])
                np.testing.assert_allclose(got[1], s[1], rtol=2e-3, err_msg=key)
                np.testing.assert_allclose(got[0], s[0], rtol=0, atol=2e-3 * max(s[1], 1e-6), err_msg=key)


def _random_case(cfg, N, R, SR, K, seed, prefix_mask=True):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(N, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0])
    tables = SimpleNamespace(xyz=xyz, embedding=torch.rand(1, N, cfg.point_features_dim, generator=g) - 0.5,
                             color=torch.rand(1, N, 3, generator=g), dir=torch.nn.functional.normalize(torch.randn(1, N, 3, generator=g), dim=-1),
                             conf=torch.rand(1, N, 1, generator=g) * 1.3 - 0.1,
                             label_embedding=torch.randn(1, N, cfg.label_embedding_dim, generator=g) if cfg.label_embedding_dim else None)
    pidx = torch.randint(0, N, (R, SR, K), generator=g).to(torch.int32)
    nv = torch.randint(0, K + 1, (R, SR), generator=g)
    nv[torch.rand(R, SR, generator=g) < 0.4] = 0
    if prefix_mask:
        m = torch.arange(K)[None, None, :] < nv[..., None]
    else:
        m = torch.rand(R, SR, K, generator=g) < 0.5
    pidx[~m] = -1
    loc_w = torch.rand(R, SR, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0])
    raydir = torch.randn(R, 3, generator=g)
    campos = torch.tensor([0.1, -0.1, -0.5])
    a = 0.3                                                  # a mild rotation keeps camera-space depth well away from zero
    rot = torch.tensor([[1.0, 0.0, 0.0], [0.0, float(np.cos(a)), float(-np.sin(a))], [0.0, float(np.sin(a)), float(np.cos(a))]])
    return tables, pidx, loc_w, raydir, campos, rot


@pytest.mark.parametrize("semantic", [False, True])
@pytest.mark.parametrize("prefix_mask", [True, False])
def test_forward_backward_vs_oracle(semantic, prefix_mask):
    cfg = rr.semantic_config() if semantic else rr.agg_config()
    N, R, SR, K = 3000, 37, 24, 8
    tables, pidx, loc_w, raydir, campos, rot = _random_case(cfg, N, R, SR, K, seed=11 + semantic, prefix_mask=prefix_mask)
    P = rr.init_params(cfg, seed=2, bias_scale=0.1)
    # oracle (autograd)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    tr = SimpleNamespace(xyz=tables.xyz, label_embedding=tables.label_embedding,
                         **{k: getattr(tables, k).clone().requires_grad_(True) for k in ("embedding", "color", "dir", "conf")})
    gn = rr.gather_neighbors(tr, pidx[None], rot[None], campos[None])
    loc_pers_ref = rr.w2pers_points(loc_w.reshape(-1, 3), rot[None], campos[None]).reshape(1, R, SR, 3)
    dirs = raydir[None, :, None, :].expand(1, R, SR, 3).contiguous()
    dec_r, valid_r, w_r, conf_r = rr.aggregator_forward(Pr, cfg, gn.color, gn.label_embedding, gn.dir, gn.conf, gn.embedding, gn.xyz_pers,
                                                        gn.xyz, gn.pnt_mask, loc_pers_ref, loc_w[None], dirs)
    g = torch.Generator().manual_seed(5)
    cot_d, cot_c = torch.randn(R, SR, 4, generator=g), torch.randn(R, SR, K, generator=g) * 0.1
    ((dec_r[0] * cot_d).sum() + (conf_r[0] * cot_c).sum()).backward()
    # CUDA
    names, W, B = param_lists(P, cfg, requires_grad=True)
    tc = {k: getattr(tables, k).clone().cuda().requires_grad_(True) for k in ("embedding", "color", "dir", "conf")}
    lab = tables.label_embedding.cuda() if semantic else None
    dec, valid, loc_pers, w, conf = ops.aggregate(cfg_to_c(cfg), W, B, tables.xyz.cuda(), tc["embedding"], tc["color"], tc["dir"], tc["conf"],
                                                  lab, pidx.cuda(), loc_w.cuda(), raydir.cuda(), campos.cuda(), rot.cuda())
    torch.testing.assert_close(dec.detach().cpu(), dec_r[0].detach(), rtol=0, atol=ATOL)
    assert torch.equal(valid.cpu().bool(), valid_r[0])
    torch.testing.assert_close(w.cpu(), w_r[0].detach(), rtol=0, atol=1e-5)
    torch.testing.assert_close(conf.detach().cpu(), conf_r[0].detach(), rtol=0, atol=1e-6)
    torch.testing.assert_close(loc_pers.cpu(), loc_pers_ref[0], rtol=1e-5, atol=1e-6)
    ((dec * cot_d.cuda()).sum() + (conf * cot_c.cuda()).sum()).backward()
    for k in ("embedding", "color", "dir", "conf"):
        assert rel_l2(tc[k].grad.cpu(), getattr(tr, k).grad) < 1e-3, k
    for nme, wt, bt in zip(names, W, B):
        assert rel_l2(wt.grad.cpu(), Pr[nme + ".weight"].grad) < 1e-3, nme
        assert rel_l2(bt.grad.cpu(), Pr[nme + ".bias"].grad) < 1e-3, nme


def test_inference_chunking_and_empty():
    """R larger than the fp32 inference chunk (4096 rays) must give the same rows as R processed alone; all-empty input gives zeros."""
    cfg = rr.agg_config(shading_feature_num=64)
    N, R, SR, K = 2000, 4096 + 300, 4, 8
    tables, pidx, loc_w, raydir, campos, rot = _random_case(cfg, N, R, SR, K, seed=3)
    P = rr.init_params(cfg, seed=1, bias_scale=0.1)
    _, W, B = param_lists(P, cfg)
    args = lambda sl: (tables.xyz.cuda(), tables.embedding.cuda(), tables.color.cuda(), tables.dir.cuda(), tables.conf.cuda(), None,
                       pidx[sl].cuda(), loc_w[sl].cuda(), raydir[sl].cuda(), campos.cuda(), rot.cuda())
    with torch.no_grad():
        full = ops.aggregate(cfg_to_c(cfg), W, B, *args(slice(None)))
        tail = ops.aggregate(cfg_to_c(cfg), W, B, *args(slice(4096, None)))
        assert torch.equal(full[0][4096:], tail[0]) and torch.equal(full[1][4096:], tail[1])
        empty = torch.full((5, SR, K), -1, dtype=torch.int32).cuda()
        out = ops.aggregate(cfg_to_c(cfg), W, B, tables.xyz.cuda(), tables.embedding.cuda(), tables.color.cuda(), tables.dir.cuda(),
                            tables.conf.cuda(), None, empty, loc_w[:5].cuda(), raydir[:5].cuda(), campos.cuda(), rot.cuda())
        assert float(out[0].abs().max()) == 0 and int(out[1].sum()) == 0


@pytest.mark.parametrize("R,SR,semantic", [(3, 24, False), (300, 24, False), (41, 80, False), (300, 24, True)])
def test_bf16_tensor_core_path_vs_fp32(R, SR, semantic):
    """bf16 tcgen05 path (forward only) against the fp32 strict path on the same inputs.
    Stated bf16 tolerance (operands rounded to bf16, fp32 accumulate): |d rgb| <= 1e-2, |d sigma| <= 2e-2 * max(1, |sigma|);
    validity masks, weights and conf coefficients are computed in fp32 on both paths and must be identical.
    semantic = the block2_bpnet configuration with the 96-d label embedding (its point-only part is a second hoisted table)."""
    cfg = rr.semantic_config() if semantic else rr.agg_config()
    N, K = 5000, 8
    tables, pidx, loc_w, raydir, campos, rot = _random_case(cfg, N, R, SR, K, seed=21 + R)
    P = rr.init_params(cfg, seed=3, bias_scale=0.1)
    _, W, B = param_lists(P, cfg)
    args = (tables.xyz.cuda(), tables.embedding.cuda(), tables.color.cuda(), tables.dir.cuda(), tables.conf.cuda(),
            tables.label_embedding.cuda() if semantic else None, pidx.cuda(), loc_w.cuda(), raydir.cuda(), campos.cuda(), rot.cuda())
    with torch.no_grad():
        ref = ops.aggregate(cfg_to_c(cfg), W, B, *args, precision=ops.PRECISION_FP32)
        out = ops.aggregate(cfg_to_c(cfg), W, B, *args, precision=ops.PRECISION_BF16)
    torch.cuda.synchronize()
    assert torch.equal(out[1], ref[1])
    torch.testing.assert_close(out[3], ref[3], rtol=0, atol=0)
    torch.testing.assert_close(out[4], ref[4], rtol=0, atol=0)
    d_rgb = float((out[0][..., 1:] - ref[0][..., 1:]).abs().max())
    sig_err = ((out[0][..., 0] - ref[0][..., 0]).abs() / ref[0][..., 0].abs().clamp(min=1.0))
    print(f"bf16 vs fp32: max |d rgb| = {d_rgb:.3e}, max rel |d sigma| = {float(sig_err.max()):.3e}")
    assert d_rgb <= 1e-2 and float(sig_err.max()) <= 2e-2


def test_bf16_edge_cases(monkeypatch):
    """bf16 path: ragged tiles -- many one-neighbour samples (more than 56 sample slots per 128-row tile, i.e. several K-sum
    passes), an all-empty input, a single valid tuple, and ray chunking (SGN_TC_CHUNK) must not change the result."""
    cfg = rr.agg_config()
    N, R, SR, K = 4000, 257, 24, 8
    tables, pidx, loc_w, raydir, campos, rot = _random_case(cfg, N, R, SR, K, seed=77)
    g = torch.Generator().manual_seed(3)
    keep = torch.rand(R, SR, generator=g) < 0.7                 # most samples keep exactly one neighbour -> ~120 slots per tile
    pidx = pidx.clone()
    first = torch.randint(0, N, (R, SR), generator=g).to(torch.int32)
    one = torch.full((R, SR, K), -1, dtype=torch.int32)
    one[..., 3] = first
    pidx = torch.where(keep[..., None], one, pidx)
    P = rr.init_params(cfg, seed=4, bias_scale=0.1)
    _, W, B = param_lists(P, cfg)
    tb = (tables.xyz.cuda(), tables.embedding.cuda(), tables.color.cuda(), tables.dir.cuda(), tables.conf.cuda(), None)
    cam = (raydir.cuda(), campos.cuda(), rot.cuda())

    def run(pi, lw, precision):
        with torch.no_grad():
            o = ops.aggregate(cfg_to_c(cfg), W, B, *tb, pi.cuda(), lw.cuda(), cam[0][:pi.shape[0]], cam[1], cam[2], precision=precision)
        torch.cuda.synchronize()
        return o
    ref = run(pidx, loc_w, ops.PRECISION_FP32)
    out = run(pidx, loc_w, ops.PRECISION_BF16)
    assert torch.equal(out[1], ref[1])
    assert float((out[0][..., 1:] - ref[0][..., 1:]).abs().max()) <= 1e-2
    assert float(((out[0][..., 0] - ref[0][..., 0]).abs() / ref[0][..., 0].abs().clamp(min=1.0)).max()) <= 2e-2
    monkeypatch.setenv("SGN_TC_CHUNK", "64")                    # 5 passes of 52 rays instead of one
    chunked = run(pidx, loc_w, ops.PRECISION_BF16)
    monkeypatch.delenv("SGN_TC_CHUNK")
    torch.testing.assert_close(chunked[0], out[0], rtol=0, atol=2e-3)   # tiles are cut differently: K-sum order inside the MMA changes
    assert torch.equal(chunked[1], out[1])
    empty = torch.full((7, SR, K), -1, dtype=torch.int32)
    o = run(empty, loc_w[:7], ops.PRECISION_BF16)
    assert float(o[0].abs().max()) == 0 and int(o[1].sum()) == 0
    single = empty.clone()
    single[4, 11, 6] = 123
    o = run(single, loc_w[:7], ops.PRECISION_BF16)
    r = run(single, loc_w[:7], ops.PRECISION_FP32)
    assert int(o[1].sum()) == 1 and float((o[0] - r[0]).abs().max()) <= 2e-2


@pytest.mark.parametrize("semantic", [False, True])
def test_bf16_point_cache(semantic):
    """The cached per-point tables (sgn_agg_point_cache_build + sgn_agg_forward_cached) give bit-identical results to rebuilding
    them inside the call, and a stale cache is the caller's responsibility (changing the embedding without rebuilding differs)."""
    cfg = rr.semantic_config() if semantic else rr.agg_config()
    N, R, SR, K = 3000, 64, 24, 8
    tables, pidx, loc_w, raydir, campos, rot = _random_case(cfg, N, R, SR, K, seed=5)
    P = rr.init_params(cfg, seed=6, bias_scale=0.1)
    _, W, B = param_lists(P, cfg)
    emb = tables.embedding.cuda()
    lab = tables.label_embedding.cuda() if semantic else None
    args = (tables.xyz.cuda(), emb, tables.color.cuda(), tables.dir.cuda(), tables.conf.cuda(), lab, pidx.cuda(), loc_w.cuda(), raydir.cuda(),
            campos.cuda(), rot.cuda())
    with torch.no_grad():
        cache = ops.build_point_cache(cfg_to_c(cfg), W, emb, lab)
        a = ops.aggregate(cfg_to_c(cfg), W, B, *args, precision=ops.PRECISION_BF16)
        b = ops.aggregate(cfg_to_c(cfg), W, B, *args, precision=ops.PRECISION_BF16, point_cache=cache)
        emb2 = emb + 0.25
        c = ops.aggregate(cfg_to_c(cfg), W, B, args[0], emb2, *args[2:], precision=ops.PRECISION_BF16, point_cache=cache)
        d = ops.aggregate(cfg_to_c(cfg), W, B, args[0], emb2, *args[2:], precision=ops.PRECISION_BF16)
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0])
    assert torch.equal(c[0], b[0]) and not torch.equal(d[0], b[0])


def _tf32_case(R, SR, semantic, prec, bwd_override=None, width=256):
    cfg = rr.semantic_config() if semantic else rr.agg_config()
    if width != 256:
        cfg = rr.agg_config(shading_feature_num=width)
    N, K = 5000, 8
    tables, pidx, loc_w, raydir, campos, rot = _random_case(cfg, N, R, SR, K, seed=21 + semantic, prefix_mask=True)
    P = rr.init_params(cfg, seed=4, bias_scale=0.1)
    g = torch.Generator().manual_seed(6)
    cot_d, cot_c = torch.randn(R, SR, 4, generator=g).cuda(), (torch.randn(R, SR, K, generator=g) * 0.1).cuda()
    names, W, B = param_lists(P, cfg, requires_grad=True)
    tc = {k: getattr(tables, k).clone().cuda().requires_grad_(True) for k in ("embedding", "color", "dir", "conf")}
    lab = tables.label_embedding.cuda() if semantic else None
    dec, valid, _, w, conf = ops.aggregate(cfg_to_c(cfg), W, B, tables.xyz.cuda(), tc["embedding"], tc["color"], tc["dir"], tc["conf"],
                                           lab, pidx.cuda(), loc_w.cuda(), raydir.cuda(), campos.cuda(), rot.cuda(), precision=prec)
    ops.BACKWARD_PRECISION_OVERRIDE = bwd_override
    try:
        ((dec * cot_d).sum() + (conf * cot_c).sum()).backward()
    finally:
        ops.BACKWARD_PRECISION_OVERRIDE = None
    grads = {k: v.grad for k, v in tc.items()}
    grads.update({n + ".weight": x.grad for n, x in zip(names, W)})
    grads.update({n + ".bias": x.grad for n, x in zip(names, B)})
    return dec.detach(), valid, grads


TF32_SHAPES = [(37, 24, False), (37, 24, True), (700, 24, False), (129, 80, True)]


@pytest.mark.parametrize("R,SR,semantic", TF32_SHAPES)
def test_tf32_backward_gemms_on_shared_activations(R, SR, semantic):
    """The tcgen05 kind::tf32 dgrad / wgrad kernels against the fp32 SIMT ones on the SAME saved activations (TF32 forward, then
    the backward once per arithmetic): only the rounding of the backward GEMM operands differs (the tensor core truncates fp32 operands to TF32, a bias of up to 2^-10
    per operand that compounds over the six dgrad layers: observed <= 5e-3).  Relative L2 <= 1e-2 per tensor."""
    d0, v0, g_tf = _tf32_case(R, SR, semantic, ops.PRECISION_TF32)
    d1, v1, g_fp = _tf32_case(R, SR, semantic, ops.PRECISION_TF32, bwd_override=ops.PRECISION_FP32)
    assert torch.equal(d0, d1) and torch.equal(v0, v1)          # the forward is deterministic: both backwards saw the same workspace
    for k in g_fp:
        assert rel_l2(g_tf[k], g_fp[k]) < 1e-2, (k, rel_l2(g_tf[k], g_fp[k]))


def test_tf32_backward_after_an_fp32_forward():
    """Forward and backward may run in different arithmetics: the tensor-core dgrads read the activations' sign bits, which an fp32 (SIMT)
    forward provides too (mask_from_act_kernel).  TF32 backward vs fp32 backward on the same fp32 activations: relative L2 <= 1e-2."""
    d0, v0, g_tf = _tf32_case(129, 24, True, ops.PRECISION_FP32, bwd_override=ops.PRECISION_TF32)
    d1, v1, g_fp = _tf32_case(129, 24, True, ops.PRECISION_FP32)
    assert torch.equal(d0, d1) and torch.equal(v0, v1)
    for k in g_fp:
        assert rel_l2(g_tf[k], g_fp[k]) < 1e-2, (k, rel_l2(g_tf[k], g_fp[k]))


@pytest.mark.parametrize("width", [64, 128])
def test_tf32_backward_gemms_narrow_layers(width):
    """Same check at shading_feature_num 64 / 128 (colour width 32 / 64): partial MMA tiles, one accumulator half in the wgrad."""
    d0, v0, g_tf = _tf32_case(300, 24, False, ops.PRECISION_TF32, width=width)
    d1, v1, g_fp = _tf32_case(300, 24, False, ops.PRECISION_TF32, bwd_override=ops.PRECISION_FP32, width=width)
    assert torch.equal(d0, d1)
    for k in g_fp:
        assert rel_l2(g_tf[k], g_fp[k]) < 1e-2, (k, rel_l2(g_tf[k], g_fp[k]))
    d2, _, _ = _tf32_case(300, 24, False, ops.PRECISION_FP32, width=width)
    assert bool(((d0 - d2).abs() <= 1e-2 * d2.abs().clamp(min=1.0)).all())


@pytest.mark.parametrize("R,SR,semantic", TF32_SHAPES)
def test_tf32_training_path_vs_fp32(R, SR, semantic):
    """SGN_PRECISION_TF32 (tcgen05 kind::tf32 GEMMs, fp32 storage and accumulation -- the arithmetic of the reference's cuBLAS default
    at its pinned torch 1.10) against the fp32 SIMT path end to end.  Stated TF32 tolerance: |d decoded| <= 1e-2 * max(1, |decoded|).
    End-to-end gradients differ by more than operand rounding: a pre-activation within ~1e-3 of zero changes sign between the two
    arithmetics and LeakyReLU(0.01) then scales that element's gradient by 1 instead of 0.01 (about 0.3 % of the elements, i.e.
    ~5 % relative L2 -- a CPU emulation that truncates the operands of every torch Linear to TF32 shows the same figures), so the
    end-to-end bound is direction + scale: cosine >= 0.99 and relative L2 <= 0.12; the GEMMs themselves are held to 1e-2 above."""
    dec_a, valid_a, g_a = _tf32_case(R, SR, semantic, ops.PRECISION_FP32)
    dec_b, valid_b, g_b = _tf32_case(R, SR, semantic, ops.PRECISION_TF32)
    assert torch.equal(valid_a, valid_b)
    assert bool(((dec_a - dec_b).abs() <= 1e-2 * dec_a.abs().clamp(min=1.0)).all()), float((dec_a - dec_b).abs().max())
    for k in g_a:
        a, b = g_a[k].double().flatten(), g_b[k].double().flatten()
        cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.99 and rel_l2(g_b[k], g_a[k]) < 0.12, (k, cos, rel_l2(g_b[k], g_a[k]))


@pytest.mark.parametrize("precision", [ops.PRECISION_FP32, ops.PRECISION_TF32])
def test_training_path_with_no_valid_tuple(precision):
    """Edge case: no sample has a neighbour (M = 0 for every GEMM, read from the device): forward gives zeros, backward gives zero
    gradients and nothing hangs or faults -- and a single valid tuple (M = 1, one partial tile) matches between the two arithmetics."""
    cfg = rr.agg_config()
    N, R, SR, K = 500, 9, 24, 8
    tables, pidx, loc_w, raydir, campos, rot = _random_case(cfg, N, R, SR, K, seed=5)
    P = rr.init_params(cfg, seed=1, bias_scale=0.1)

    def run(pi):
        names, W, B = param_lists(P, cfg, requires_grad=True)
        emb = tables.embedding.clone().cuda().requires_grad_(True)
        dec, valid, _, w, conf = ops.aggregate(cfg_to_c(cfg), W, B, tables.xyz.cuda(), emb, tables.color.cuda(), tables.dir.cuda(), tables.conf.cuda(),
                                               None, pi.cuda(), loc_w.cuda(), raydir.cuda(), campos.cuda(), rot.cuda(), precision=precision)
        (dec.sum() + 0.0 * conf.sum()).backward()
        torch.cuda.synchronize()
        return dec.detach(), valid, emb.grad, [x.grad for x in W]

    empty = torch.full_like(pidx, -1)
    dec, valid, g_emb, g_w = run(empty)
    assert float(dec.abs().max()) == 0.0 and int(valid.sum()) == 0
    assert float(g_emb.abs().max()) == 0.0 and all(float(g.abs().max()) == 0.0 for g in g_w)
    one = empty.clone()
    one[3, 7, 0] = 11
    dec1, valid1, g_emb1, g_w1 = run(one)
    assert int(valid1.sum()) == 1 and float(dec1[3, 7].abs().max()) > 0 and float(g_emb1[0, 11].abs().max()) > 0
    rest = g_emb1.clone()
    rest[0, 11] = 0.0
    assert float(rest.abs().max()) == 0.0                                        # only the gathered point receives a gradient


@pytest.mark.parametrize("precision", [ops.PRECISION_FP32, ops.PRECISION_BF16])
def test_sample_mask_changes_nothing(precision):
    """ops.aggregate(sample_mask=ops.query's mask): the all -1 index rows of slots without a sample are not read (sgn_agg_forward_frame_masked);
    every output is bit for bit what the unmasked call gives."""
    from sgnerf_b200 import pipeline, synth
    s = synth.scene_c0(n_points=60_000, n_rays=700)
    tabs = synth.make_point_tables(60_000, 32, 0, seed=0, conf_spread=0.2)
    cfg = rr.agg_config()
    P = rr.init_params(cfg, seed=1, bias_scale=0.05)
    names = [n for n, _, _ in rr.layer_shapes(cfg)]
    scene = pipeline.RenderScene(torch.from_numpy(s.xyz), tabs.embedding.reshape(60_000, -1), tabs.color.reshape(60_000, 3), tabs.dir.reshape(60_000, 3),
                                 tabs.conf.reshape(60_000), [P[n + ".weight"] for n in names], [P[n + ".bias"] for n in names], cfg_to_c(cfg),
                                 pipeline.query_options(), device="cuda")
    grid, hp = scene.grid()
    campos, rot = torch.from_numpy(s.campos).cuda(), torch.from_numpy(s.camrotc2w).cuda()
    raydir = torch.from_numpy(s.raydir).cuda()
    q = scene.qopt
    pidx, loc_w, smask, rmask = ops.query(grid, campos, raydir, scene.depth_candidates(s.near, s.far), q.SR, q.K, q.kernel_size[0], hp.radius2)
    assert 0 < int((smask > 0).sum()) < smask.numel() and bool((pidx[smask == 0] == -1).all())
    with torch.no_grad():
        a = ops.aggregate(scene.agg_cfg, scene.weights, scene.biases, scene.xyz, scene.embedding, scene.color, scene.dirs, scene.conf, None, pidx, loc_w,
                          raydir, campos, rot, precision=precision, want_aux=True)
        b = ops.aggregate(scene.agg_cfg, scene.weights, scene.biases, scene.xyz, scene.embedding, scene.color, scene.dirs, scene.conf, None, pidx, loc_w,
                          raydir, campos, rot, precision=precision, want_aux=True, sample_mask=smask)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
