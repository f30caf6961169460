"""CPU: the oracle restatement against the golden vectors produced by the REFERENCE's own modules
(tests/golden/make_golden.py).  Bit-exact where the restatement uses the same torch ops."""
import ast
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import query_ref as qr
from oracle import render_ref as rr


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def test_positional_encoding_and_ray_positions(golden_dir):
    g = load(golden_dir, "pe_rays")
    x = torch.from_numpy(g["pe_x"])
    for F, ori in ((3, False), (5, False), (4, True)):
        assert np.array_equal(rr.positional_encoding(x, F, ori=ori).numpy(), g[f"pe_{F}_{int(ori)}"])
    campos, raydir = torch.from_numpy(g["rays_campos"]), torch.from_numpy(g["rays_dir"])
    for jit in (0.0, 0.3):
        mid = torch.from_numpy(g[f"rays_mid_{jit}"])
        assert np.array_equal(qr.raypos_from_t(campos, raydir, mid[0]).numpy(), g[f"rays_pos_{jit}"])
    _, mid = qr.near_far_linear_ray_generation(campos, raydir, 400, 0.1, 8.0, jitter=0.0)
    assert np.array_equal(mid.numpy(), g["rays_mid_0.0"])


@pytest.mark.parametrize("blend", ["alpha", "alpha2"])
def test_ray_march(golden_dir, blend):
    g = load(golden_dir, "ray_march")
    feats = torch.from_numpy(g["feats"]).requires_grad_(True)
    out = rr.ray_march(torch.from_numpy(g["dist"]), torch.from_numpy(g["valid"]), feats, torch.from_numpy(g["bg"]), blend)
    names = ["ray_color", "point_color", "opacity", "acc_transmission", "blend_weight", "bg_transmission", "bg_blend_weight"]
    for n, a in zip(names, out):
        np.testing.assert_allclose(a.detach().numpy(), g[f"{blend}_{n}"], rtol=0, atol=1e-6, err_msg=n)
    ((out[0] * torch.from_numpy(g[f"{blend}_cot_color"])).sum() + (out[2] * torch.from_numpy(g[f"{blend}_cot_opacity"])).sum()).backward()
    np.testing.assert_allclose(feats.grad.numpy(), g[f"{blend}_grad_feats"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["agg_small_plain", "agg_small_semantic", "agg_canonical_plain", "agg_canonical_semantic"])
def test_aggregator_forward_backward(golden_dir, name):
    g = load(golden_dir, name)
    cfg = SimpleNamespace(**ast.literal_eval(str(g["cfg"])))
    if "P_seed" in g:
        P = rr.init_params(cfg, seed=int(g["P_seed"]), bias_scale=0.1)
        sig = np.array([float(sum(v.double().sum() for v in P.values())), float(sum(v.double().abs().sum() for v in P.values()))])
        np.testing.assert_allclose(sig, g["P_signature"], rtol=1e-9)       # the seeded weights are the ones the golden run used
    else:
        P = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("P_")}
    P = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    inp = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in_")}
    leaves = {}
    for k in ("sampled_color", "sampled_dir", "sampled_conf", "sampled_embedding", "sampled_label_embedding"):
        if k in inp:
            inp[k] = inp[k].clone().requires_grad_(True)
            leaves[k] = inp[k]
    out, ray_valid, weight, conf = rr.aggregator_forward(
        P, cfg, inp["sampled_color"], inp.get("sampled_label_embedding"), inp["sampled_dir"], inp["sampled_conf"],
        inp["sampled_embedding"], inp["sampled_xyz_pers"], inp["sampled_xyz"], inp["sample_pnt_mask"], inp["sample_loc"],
        inp["sample_loc_w"], inp["sample_ray_dirs"])
    np.testing.assert_allclose(out.detach().numpy(), g["out_decoded"], rtol=0, atol=1e-6)
    assert np.array_equal(ray_valid.numpy(), g["out_ray_valid"])
    np.testing.assert_allclose(weight.detach().numpy(), g["out_weight"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(conf.detach().numpy(), g["out_conf"], rtol=0, atol=1e-7)
    ((out * torch.from_numpy(g["cot_decoded"])).sum() + (conf * torch.from_numpy(g["cot_conf"])).sum()).backward()
    for k, v in leaves.items():
        ref = g["g_" + k]
        np.testing.assert_allclose(v.grad.numpy(), ref, rtol=1e-4, atol=1e-6 * max(1.0, np.abs(ref).max()), err_msg=k)
    for k, v in P.items():
        if "gw_" + k in g:
            ref = g["gw_" + k]
            np.testing.assert_allclose(v.grad.numpy(), ref, rtol=1e-4, atol=1e-5 * max(1e-3, np.abs(ref).max()), err_msg=k)
        else:
            s = g["sig_gw_" + k]
            np.testing.assert_allclose([float(v.grad.double().sum()), float(v.grad.double().abs().sum())], s, rtol=1e-4, atol=1e-6)


def test_probe_outputs_known_answer():
    """Hand-made case for the `prob == 1` restatement: the first of two equal opacities wins, an invalid slot reads point 0."""
    from types import SimpleNamespace
    xyz = torch.tensor([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 2.0, 0.0]])
    tables = SimpleNamespace(xyz=xyz, embedding=torch.tensor([[[1.0, 1.0], [2.0, 0.0], [0.0, 4.0]]]), color=torch.tensor([[[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]]),
                             dir=torch.tensor([[[0.0, 0, 1.0], [0, 1.0, 0], [1.0, 0, 0]]]), conf=torch.tensor([[[0.5], [1.0], [0.25]]]), label_embedding=None)
    pidx = torch.tensor([[[[1, 2], [2, -1]]]])                       # [1,1,2,2]: sample 1 has one invalid slot
    gn = rr.gather_neighbors(tables, pidx, torch.eye(3)[None], torch.tensor([[0.0, 0.0, -5.0]]))
    opacity = torch.tensor([[[0.7, 0.7]]])
    loc_w = torch.tensor([[[[1.0, 0.0, 1.0], [9.0, 9.0, 9.0]]]])
    weight = torch.tensor([[[[0.25, 0.75], [1.0, 0.0]]]])
    conf_c = torch.tensor([[[[1.0, 0.5], [1.0, 1.0]]]])
    o = rr.probe_outputs(opacity, loc_w, weight, conf_c, gn)
    assert torch.equal(o["ray_max_sample_loc_w"], torch.tensor([[[1.0, 0.0, 1.0]]]))                    # sample 0, not 1
    torch.testing.assert_close(o["ray_max_far_dist"], torch.tensor([[[1.0]]]))                          # point 1 at distance 1
    torch.testing.assert_close(o["shading_avg_color"], torch.tensor([[[0.0, 0.25, 0.375]]]))
    torch.testing.assert_close(o["shading_avg_conf"], torch.tensor([[[0.25 * 1.0 + 0.375 * 0.25]]]))
    torch.testing.assert_close(o["shading_avg_embedding"], torch.tensor([[[0.5, 1.5]]]))


def test_query_oracle_against_reference_kernel_vectors(golden_dir):
    """The C restatement of the query (oracle/query_ref.c) against outputs of the REFERENCE's OWN CUDA kernels: the vectors were
    written on a B200 by tests/test_gpu_reference_kernels.py (reference source compiled unchanged into oracle/_ref, scene C0:
    100k points, 1024 rays) and are re-checked here on every CPU run.  The reference's slot numbering depends on atomic arrival order,
    so neighbours compare as per-sample sorted sets, away from the voxel its `voxel_idx > 0` guard empties (a different voxel per run)."""
    import os
    from oracle import query_ref as qr
    from sgnerf_b200 import synth
    from tests import util
    g = np.load(os.path.join(golden_dir, "query_reference_kernels_c0.npz"))
    s = synth.scene_c0(n_points=int(g["n_points"]), n_rays=int(g["n_rays"]))
    opt = qr.default_opt(SR=int(g["SR"]))
    t = util.shared_t(s.near, s.far, opt.z_depth_dim)
    o_pidx, _, o_loc_w, _, o_ray_mask, _, _, info = util.oracle_query(s, opt, t)
    assert int(info.grid.occ_idx[0]) == int(g["ref_occ_idx"][0])                       # occupied voxels claimed
    o_mask = o_ray_mask[0].numpy()
    assert int((o_mask != g["ref_ray_mask"]).sum()) <= 4                               # rays that only see the slot-0 voxel
    oidx = np.cumsum(o_mask > 0) - 1
    rays = g["ray_ids"]
    assert (o_mask[rays] > 0).all()
    ol = o_loc_w[0].numpy()[oidx[rays]]
    op_ = o_pidx[0].numpy()[oidx[rays]]
    assert np.array_equal(ol.view(np.int32), g["ref_loc_w"].view(np.int32)), "shading sample positions differ from the reference kernels"
    hp = info.hp
    near = lambda coor: np.abs(np.floor((ol - hp.ranges[:3]) / hp.scaled_vsize) - np.asarray(coor, np.float32)).max(-1) <= 1
    touched = near(g["ref_slot0"]) | near(info.grid.occ_2_coor[0])
    same = (np.sort(op_, -1) == g["ref_pidx_sorted"]).all(-1)
    assert same[~touched].all() and (~touched).mean() > 0.95


def test_perspective_query_oracle_against_reference_kernel_vectors(golden_dir):
    """oracle/query_pers_ref.c against outputs of the REFERENCE's OWN perspective kernels (query_point_indices.py:263-590, compiled unchanged
    into oracle/_ref/libref_query_pers_K8.so and run on a B200 by tests/test_gpu_pers_query.py, case "wide"): ray mask, sample positions
    bit for bit, neighbours as per-sample sorted sets (the reference's list order follows its atomics)."""
    import os
    from oracle import query_pers_ref as qp
    from sgnerf_b200 import synth
    from tests import util
    from tests.test_gpu_pers_query import CASES
    g = np.load(os.path.join(golden_dir, "pers_reference_kernels_c0.npz"))
    s = util.pers_scene(int(g["n_points"]), int(g["n_rays"]))
    opt = qp.default_opt(max_o=127, P=int(g["P"]), **CASES["wide"])
    hp = qp.get_hyperparameters(opt, s.height, s.width, synth.SCANNET_INTRINSIC, s.near, s.far)
    o_pidx, o_loc, o_mask = qp.query_uncompacted(opt, hp, s.pixel_idx, s.xyz_pers)
    assert np.array_equal(o_mask, g["ref_ray_mask"])
    sel = o_mask > 0
    rows = g["ray_rows"]
    assert np.array_equal(o_loc[sel][rows].view(np.int32), g["ref_loc"].view(np.int32))
    same = (np.sort(o_pidx[sel][rows], -1) == g["ref_pidx_sorted"]).all(-1)
    assert (g["ref_pidx_sorted"] >= 0).sum() > 1000
    assert (~same).mean() < 1e-3
