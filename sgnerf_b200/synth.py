"""Synthetic scenes of the shapes BASELINE.json names.

Everything is seeded: geometry from numpy.random.default_rng(seed), per-point tables and MLP
weights from a torch.Generator.  No dataset or checkpoint is read (there is no network); the
reference's own data layer (data/scannet_ft_dataset.py) is out of scope and only its ray
convention (data/data_utils.py:55-69, un-normalised directions with z_cam = 1) is followed.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch

# ScanNet colour intrinsics quoted at models/neural_points/neural_points.py:43 of the reference
SCANNET_INTRINSIC = np.array([[577.870605, 0.0, 319.5], [0.0, 577.870605, 239.5], [0.0, 0.0, 1.0]], dtype=np.float64)


def _box_faces(lo, hi):
    """Six axis-aligned faces of a box as (origin, edge_u, edge_v)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    d = hi - lo
    faces = []
    for ax in range(3):
        u, v = (ax + 1) % 3, (ax + 2) % 3
        eu, ev = np.zeros(3), np.zeros(3)
        eu[u], ev[v] = d[u], d[v]
        for side in (0, 1):
            o = lo.copy()
            o[ax] = hi[ax] if side else lo[ax]
            faces.append((o, eu, ev))
    return faces


def make_room_cloud(n_points, size=(6.0, 5.0, 3.0), n_boxes=6, noise=0.002, seed=1234):
    """Surface samples (+ N(0, noise)) of a room shell and `n_boxes` interior boxes.  Returns f32 [N,3]."""
    rng = np.random.default_rng(seed)
    size = np.asarray(size, float)
    faces = _box_faces(np.zeros(3), size)
    for _ in range(n_boxes):
        ext = rng.uniform(0.3, 1.2, 3) * np.array([1.0, 1.0, min(1.0, size[2] / 3.0)])
        lo = np.array([rng.uniform(0.3, size[0] - 0.3 - ext[0]), rng.uniform(0.3, size[1] - 0.3 - ext[1]), 0.0])
        faces += _box_faces(lo, lo + ext)
    area = np.array([np.linalg.norm(np.cross(eu, ev)) for _, eu, ev in faces])
    which = rng.choice(len(faces), size=n_points, p=area / area.sum())
    uv = rng.random((n_points, 2))
    O = np.stack([f[0] for f in faces])[which]
    U = np.stack([f[1] for f in faces])[which]
    V = np.stack([f[2] for f in faces])[which]
    pts = O + U * uv[:, :1] + V * uv[:, 1:] + rng.normal(0.0, noise, (n_points, 3))
    return np.ascontiguousarray(pts, dtype=np.float32)


def make_object_cloud(n_points, radius=1.0, noise=0.001, seed=1234):
    """NeRF-Synthetic-shaped object: a sphere plus three boxes inside [-1.5,1.5]^3 (config C3)."""
    rng = np.random.default_rng(seed)
    n_s = n_points // 2
    v = rng.normal(size=(n_s, 3))
    sph = radius * 0.8 * v / np.linalg.norm(v, axis=1, keepdims=True)
    faces = []
    for lo, hi in (((-1.2, -1.2, -1.0), (1.2, 1.2, -0.9)), ((0.5, 0.5, -0.9), (1.0, 1.0, 0.6)), ((-1.1, 0.2, -0.9), (-0.5, 0.9, 0.2))):
        faces += _box_faces(lo, hi)
    area = np.array([np.linalg.norm(np.cross(eu, ev)) for _, eu, ev in faces])
    n_b = n_points - n_s
    which = rng.choice(len(faces), size=n_b, p=area / area.sum())
    uv = rng.random((n_b, 2))
    O = np.stack([f[0] for f in faces])[which]
    U = np.stack([f[1] for f in faces])[which]
    V = np.stack([f[2] for f in faces])[which]
    pts = np.concatenate([sph, O + U * uv[:, :1] + V * uv[:, 1:]]) + rng.normal(0.0, noise, (n_points, 3))
    return np.ascontiguousarray(rng.permutation(pts), dtype=np.float32)


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """Camera-to-world rotation with x right, y down, z forward (the convention get_dtu_raydir assumes)."""
    eye, target, up = np.asarray(eye, float), np.asarray(target, float), np.asarray(up, float)
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z], axis=1)          # columns = camera axes in world


def pixel_rays(px, py, intrinsic, camrotc2w):
    """data/data_utils.py:55-69 convention: d_cam = ((u+.5-cx)/fx, (v+.5-cy)/fy, 1); d_w = d_cam @ R^T."""
    x = (px + 0.5 - intrinsic[0, 2]) / intrinsic[0, 0]
    y = (py + 0.5 - intrinsic[1, 2]) / intrinsic[1, 1]
    d = np.stack([x, y, np.ones_like(x)], axis=-1) @ camrotc2w.T
    return np.ascontiguousarray(d.reshape(-1, 3), dtype=np.float32)


def full_frame_pixels(width, height, margin=0):
    px, py = np.meshgrid(np.arange(margin, width - margin, dtype=np.float32),
                         np.arange(margin, height - margin, dtype=np.float32))
    return px.reshape(-1), py.reshape(-1)


def random_pixels(n, width, height, margin=10, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.integers(margin, width - margin, n).astype(np.float32),
            rng.integers(margin, height - margin, n).astype(np.float32))


def make_point_tables(n_points, feat_dim=32, label_dim=0, seed=0, conf_spread=0.0, device="cpu"):
    """Per-point tables in the reference's parameter shapes: embedding U(-.5,.5)
    (feature_init_method=rand, neural_points.py:386), colour U(0,1), unit dirs, conf 1 (optionally spread
    over [1-conf_spread, 1+conf_spread] so the straight-through clamp is exercised)."""
    g = torch.Generator().manual_seed(seed)
    emb = torch.rand(1, n_points, feat_dim, generator=g) - 0.5
    color = torch.rand(1, n_points, 3, generator=g)
    d = torch.randn(1, n_points, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    conf = torch.ones(1, n_points, 1)
    if conf_spread > 0:
        conf = conf + (torch.rand(1, n_points, 1, generator=g) * 2 - 1) * conf_spread
    lab = torch.randn(1, n_points, label_dim, generator=g) if label_dim > 0 else None
    to = lambda t: None if t is None else t.to(device)
    return SimpleNamespace(embedding=to(emb), color=to(color), dir=to(d), conf=to(conf), label_embedding=to(lab))


def mlp_layer_shapes(feat_dim=32, num_feat_freqs=3, dist_xyz_freq=5, num_viewdir_freqs=4, width=256,
                     layers1=2, layers2_bpnet=0, label_dim=0, layers3=2, color_layers=4):
    """(state_dict prefix, in, out) of the aggregator's Linear layers (alpha branch: one layer)."""
    in_ch = feat_dim * (1 + 2 * num_feat_freqs) + 2 * dist_xyz_freq * 6
    out = []
    for i in range(layers1):
        out.append((f"block1.{2 * i}", in_ch, width)); in_ch = width
    if layers2_bpnet > 0:
        in_ch += label_dim
        for i in range(layers2_bpnet):
            out.append((f"block2_bpnet.{2 * i}", in_ch, width)); in_ch = width
    if layers3 > 0:
        in_ch += 7
        for i in range(layers3):
            out.append((f"block3.{2 * i}", in_ch, width)); in_ch = width
    out.append(("alpha_branch.0", width, 1))
    c_in = width + 6 * num_viewdir_freqs
    for i in range(color_layers - 1):
        out.append((f"color_branch.{2 * i}", c_in, width // 2)); c_in = width // 2
    out.append((f"color_branch.{2 * (color_layers - 1)}", c_in, 3))
    return out


def make_mlp_params(shapes, seed=0, slope=0.01, bias_scale=0.0, device="cpu"):
    """Xavier-uniform with LeakyReLU gain where an activation follows (reference init, appendix A.8)."""
    g = torch.Generator().manual_seed(seed)
    gain_act = math.sqrt(2.0 / (1 + slope ** 2))
    last = {n.split(".")[0]: n for n, _, _ in shapes}
    P = {}
    for name, cin, cout in shapes:
        act = name.startswith("block") or last[name.split(".")[0]] != name
        bound = (gain_act if act else 1.0) * math.sqrt(2.0 / (cin + cout)) * math.sqrt(3.0)
        P[name + ".weight"] = ((torch.rand(cout, cin, generator=g) * 2 - 1) * bound).to(device)
        P[name + ".bias"] = ((torch.rand(cout, generator=g) * 2 - 1) * bias_scale).to(device)
    return P


def scene_c0(n_points=100_000, n_rays=1024, seed=1234):
    """Config C0 (CPU-runnable): 100k-point room, 1024 random pixels of a 640x480 ScanNet-intrinsics view."""
    xyz = make_room_cloud(n_points, (6.0, 5.0, 3.0), 6, 0.002, seed)
    eye, target = np.array([1.2, 1.0, 1.5]), np.array([4.5, 3.8, 1.1])
    R = look_at(eye, target)
    px, py = random_pixels(n_rays, 640, 480, 10, seed)
    return SimpleNamespace(xyz=xyz, campos=eye.astype(np.float32), camrotc2w=R.astype(np.float32),
                           raydir=pixel_rays(px, py, SCANNET_INTRINSIC, R), px=px, py=py,
                           near=0.1, far=8.0, width=640, height=480)


def scene_room(n_points, room=(8.0, 8.0, 3.0), width=640, height=480, seed=1234, pixels=None):
    """Configs C1/C2/C4: room cloud + a full frame (or `pixels` = (px, py)) from a camera inside the room."""
    xyz = make_room_cloud(n_points, room, 6, 0.002, seed)
    eye = np.array([room[0] * 0.2, room[1] * 0.2, 1.5])
    target = np.array([room[0] * 0.75, room[1] * 0.7, 1.1])
    R = look_at(eye, target)
    K = SCANNET_INTRINSIC.copy()
    K[0, 2], K[1, 2] = (width - 1) / 2.0, (height - 1) / 2.0
    K[0, 0] = K[1, 1] = SCANNET_INTRINSIC[0, 0] * width / 640.0
    px, py = pixels if pixels is not None else full_frame_pixels(width, height)
    return SimpleNamespace(xyz=xyz, campos=eye.astype(np.float32), camrotc2w=R.astype(np.float32),
                           raydir=pixel_rays(px, py, K, R), px=px, py=py, near=0.1, far=8.0,
                           width=width, height=height)


def scene_c3(n_points=3_000_000, width=800, height=800, seed=1234, pixels=None):
    """Config C3: NeRF-Synthetic-shaped object cloud in [-1.5,1.5]^3, 800x800 frame, focal 1111, camera on a
    sphere of radius 4 looking at the origin, near 2 / far 6.  Query options that go with it: vsize .004, P = 9, SR = 200."""
    xyz = make_object_cloud(n_points, 1.0, 0.001, seed)
    eye = 4.0 * np.array([0.6, -0.64, 0.48]) / np.linalg.norm([0.6, -0.64, 0.48])
    R = look_at(eye, np.zeros(3))
    K = np.array([[1111.0, 0.0, (width - 1) / 2.0], [0.0, 1111.0, (height - 1) / 2.0], [0.0, 0.0, 1.0]])
    K[0, 0] = K[1, 1] = 1111.0 * width / 800.0
    px, py = pixels if pixels is not None else full_frame_pixels(width, height)
    return SimpleNamespace(xyz=xyz, campos=eye.astype(np.float32), camrotc2w=R.astype(np.float32), raydir=pixel_rays(px, py, K, R),
                           px=px, py=py, near=2.0, far=6.0, width=width, height=height)


C3_QUERY = dict(vsize=(0.004, 0.004, 0.004), P=9, SR=200)


def scene_c4(n_points=10_000_000, width=1296, height=968, seed=1234, pixels=None):
    """Config C4: 10M-point room of 20 x 20 x 4 m at raw ScanNet resolution."""
    return scene_room(n_points, room=(20.0, 20.0, 4.0), width=width, height=height, seed=seed, pixels=pixels)
