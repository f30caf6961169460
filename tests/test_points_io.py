"""CPU: on-disk formats around the render path (sgnerf_b200/points_io.py): the PLY vertex reader against files written in the
reference's layout (binary little-endian, double positions + normals + uchar colours; ASCII; an element in front of the vertices),
load_init_points' crop / feature scaling (data/scannet_ft_dataset.py:463-495) and the checkpoint key layout."""
import numpy as np
import torch

from sgnerf_b200 import points_io


def test_ply_roundtrip_binary_and_ascii(tmp_path):
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(257, 3))
    nrm = rng.normal(size=(257, 3))
    rgb = rng.integers(0, 256, (257, 3)).astype(np.uint8)
    p = tmp_path / "pcd.ply"
    points_io.write_ply_vertices(p, xyz, nrm, rgb, dtype="double")
    v = points_io.read_ply_vertices(p)
    assert v.dtype.names == ("x", "y", "z", "nx", "ny", "nz", "red", "green", "blue") and v["x"].dtype == np.float64
    assert np.array_equal(np.stack([v["x"], v["y"], v["z"]], -1), xyz) and np.array_equal(np.stack([v["red"], v["green"], v["blue"]], -1), rgb)
    # ASCII body, float positions, a comment and an element in front of the vertices
    a = tmp_path / "a.ply"
    with open(a, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by hand\nelement camera 2\nproperty float fx\nelement vertex 3\nproperty float x\nproperty float y\n"
                "property float z\nproperty uchar red\nend_header\n1.0\n2.0\n0.5 1.5 -2 7\n1e-3 0 4 255\n3 3 3 0\n")
    v = points_io.read_ply_vertices(a)
    assert v["x"].dtype == np.float32 and np.allclose(v["x"], [0.5, 1e-3, 3]) and list(v["red"]) == [7, 255, 0]


def test_load_init_points_crops_and_scales(tmp_path):
    rng = np.random.default_rng(1)
    xyz = rng.uniform(-3, 3, size=(1000, 3))
    feats = rng.uniform(-1, 1, size=(1000, 3)).astype(np.float32)
    points_io.write_ply_vertices(tmp_path / "pcd.ply", xyz, np.zeros_like(xyz), np.zeros((1000, 3), np.uint8))
    torch.save((xyz.astype(np.float32), feats), tmp_path / "points.pth")
    ranges = [-1.0, -2.0, -0.5, 2.0, 1.0, 2.5]
    p, f = points_io.load_init_points(tmp_path / "pcd.ply", tmp_path / "points.pth", ranges, device="cpu")
    x32 = xyz.astype(np.float32)
    keep = np.all((x32 >= np.float32(ranges[:3])) & (x32 <= np.float32(ranges[3:])), axis=1)
    assert p.dtype == torch.float32 and np.array_equal(p.numpy(), x32[keep])
    assert np.array_equal(f.numpy(), ((feats + 1.0) * 127.5).astype(np.float32)[keep])
    p2, f2 = points_io.load_init_points(tmp_path / "pcd.ply", None, [-100.0] * 3 + [100.0] * 3, device="cpu")
    assert p2.shape[0] == 1000 and f2 is None


def test_checkpoint_layout_roundtrip(tmp_path):
    """save_ray_marching_checkpoint writes the keys of SURVEY.md appendix B; pipeline.scene_from_checkpoint's shape logic reads them back."""
    from types import SimpleNamespace
    from oracle import render_ref as rr
    from sgnerf_b200 import modules, pipeline
    opt = SimpleNamespace(point_features_dim=32, num_feat_freqs=3, dist_xyz_freq=5, num_viewdir_freqs=4, shading_feature_num=256,
                          shading_feature_mlp_layer1=2, shading_feature_mlp_layer2_bpnet=0, shading_feature_mlp_layer3=2, shading_alpha_mlp_layer=1,
                          shading_color_mlp_layer=4, act_super=1)
    agg = modules.PointAggregator(opt)
    n = 50
    npnts = SimpleNamespace(state_dict=lambda: {"xyz": torch.randn(n, 3), "points_embeding": torch.randn(1, n, 32), "points_conf": torch.ones(1, n, 1),
                                                "points_dir": torch.randn(1, n, 3), "points_color": torch.rand(1, n, 3), "Rw2c": torch.eye(3)})
    sd = points_io.save_ray_marching_checkpoint(tmp_path / "100_net_ray_marching.pth", npnts, agg)
    back = torch.load(tmp_path / "100_net_ray_marching.pth")
    assert set(back) == set(sd) and "neural_points.points_embeding" in back and "aggregator.block1.0.weight" in back
    cfg, names = pipeline.agg_cfg_from_state_dict(back)
    assert names == [k for k, _, _ in rr.layer_shapes(rr.agg_config())] and cfg.width == 256
