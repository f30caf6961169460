"""On-disk formats of the point cloud and of the network the reference reads and writes around the render path:

    exported/pcd.ply     vertices of the scene mesh: x, y, z (float or double), nx, ny, nz, red, green, blue (uchar); written with
                         `plyfile` as binary little-endian (data/scannet_ft_dataset.py:436-461), read back by load_init_points (:463-475)
    exported/points.pth  torch.save of a pair (locs ndarray [N,3], feats ndarray [N,3] in [-1, 1]) (:479-484); BPNet's input features are
                         (feats + 1) * 127.5
    <iter>_net_ray_marching.pth    state_dict of NeuralPointsRayMarching: `neural_points.*` + `aggregator.*` (models/base_model.py:84-118)

The PLY reader / writer below is self-contained (`plyfile` is not a dependency of this package): header parsing, ASCII and
binary_little_endian / binary_big_endian bodies, scalar properties of the `vertex` element; list properties (faces) are skipped.
"""
import os

import numpy as np
import torch

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply_vertices(path):
    """Structured numpy array of the `vertex` element of a PLY file (all its scalar properties)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": [], "has_list": False}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    cur["has_list"] = True
                    cur["props"].append((tok[4], ("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]])))
                else:
                    cur["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unknown PLY format {fmt!r}")
        for el in elements:
            if el["name"] == "vertex":
                if el["has_list"]:
                    raise ValueError(f"{path}: list properties on the vertex element are not supported")
                if fmt == "ascii":
                    rows = [f.readline().split() for _ in range(el["count"])]
                    dt = np.dtype([(n, t) for n, t in el["props"]])
                    out = np.zeros(el["count"], dtype=dt)
                    for j, (n, t) in enumerate(el["props"]):
                        out[n] = np.array([r[j] for r in rows], dtype=np.float64).astype(t)
                    return out
                order = "<" if fmt == "binary_little_endian" else ">"
                dt = np.dtype([(n, order + t) for n, t in el["props"]])
                return np.frombuffer(f.read(dt.itemsize * el["count"]), dtype=dt, count=el["count"]).copy()
            # an element in front of `vertex`: skip its body
            if fmt == "ascii":
                for _ in range(el["count"]):
                    f.readline()
            elif not el["has_list"]:
                f.seek(sum(np.dtype(t).itemsize for _, t in el["props"]) * el["count"], os.SEEK_CUR)
            else:
                raise ValueError(f"{path}: cannot skip a binary list element in front of the vertices")
    raise ValueError(f"{path}: no vertex element")


def write_ply_vertices(path, xyz, normals=None, rgb=None, dtype="double"):
    """Binary little-endian PLY with the property set of the reference's pcd.ply (scannet_ft_dataset.py:447-461)."""
    xyz = np.asarray(xyz)
    n = xyz.shape[0]
    t = _PLY_TYPES[dtype]
    fields = [("x", "<" + t), ("y", "<" + t), ("z", "<" + t)]
    if normals is not None:
        fields += [("nx", "<" + t), ("ny", "<" + t), ("nz", "<" + t)]
    if rgb is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    v = np.zeros(n, dtype=np.dtype(fields))
    v["x"], v["y"], v["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if normals is not None:
        v["nx"], v["ny"], v["nz"] = normals[:, 0], normals[:, 1], normals[:, 2]
    if rgb is not None:
        v["red"], v["green"], v["blue"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    names = {"f4": "float", "f8": "double", "u1": "uchar"}
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element vertex {n}\n".encode())
        for name, ft in fields:
            f.write(f"property {names[ft.lstrip('<')]} {name}\n".encode())
        f.write(b"end_header\n")
        f.write(v.tobytes())


def load_init_points(ply_path, pth_path=None, ranges=None, device="cuda"):
    """data/scannet_ft_dataset.py:463-495: positions from pcd.ply (cast to float32), BPNet colour features from points.pth
    ((feats + 1) * 127.5), both cropped to `ranges` (when ranges[0] > -99).  Returns (points_xyz [N,3], points_feats [N,3] or None)."""
    v = read_ply_vertices(ply_path)
    xyz = torch.as_tensor(np.stack([v["x"].astype(np.float32), v["y"].astype(np.float32), v["z"].astype(np.float32)], axis=-1), device=device)
    feats = None
    if pth_path is not None:
        pth = torch.load(pth_path, weights_only=False)
        feats = torch.as_tensor(((np.asarray(pth[1]) + 1.0) * 127.5).astype(np.float32), device=device)
    if ranges is not None and ranges[0] > -99.0:
        r = torch.as_tensor(ranges, device=xyz.device, dtype=torch.float32)
        mask = torch.prod(torch.logical_and(xyz >= r[None, :3], xyz <= r[None, 3:]), dim=-1) > 0
        xyz = xyz[mask]
        if feats is not None:
            feats = feats[mask]
    return xyz, feats


def save_ray_marching_checkpoint(path, neural_points, aggregator):
    """`<iter>_net_ray_marching.pth` as base_model.py:84-95 writes it: the state_dict of the module that owns `neural_points` and
    `aggregator`, on the CPU."""
    sd = {"neural_points." + k: v.detach().cpu() for k, v in neural_points.state_dict().items()}
    sd.update({"aggregator." + k: v.detach().cpu() for k, v in aggregator.state_dict().items()})
    torch.save(sd, path)
    return sd
