#!/usr/bin/env python
"""Query stage alone on the C1 frame (profiling aid): grid build, then sgn_query a few times.  Under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'march|knn' python tools/bench_query.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgnerf_b200 import ops, pipeline, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = "cuda"
    s = synth.scene_room(1_000_000, room=(8.0, 8.0, 3.0), width=640, height=480, seed=1234)
    q = pipeline.query_options(SR=24)
    xyz = torch.from_numpy(s.xyz).to(dev)
    hp = ops.grid_hyperparameters(xyz, q.vsize, q.vscale, q.kernel_size, q.ranges, q.radius_limit_scale)
    grid = ops.OccGrid(xyz, hp.ranges[:3], hp.scaled_vsize, hp.scaled_vdim, q.query_size, q.P, q.max_o)
    campos, raydir = torch.from_numpy(s.campos).to(dev), torch.from_numpy(s.raydir).to(dev)
    t = pipeline.middle_point_ts(s.near, s.far, q.z_depth_dim, dev)
    for _ in range(n):
        out = ops.query(grid, campos, raydir, t, q.SR, q.K, q.kernel_size[0], hp.radius2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out = ops.query(grid, campos, raydir, t, q.SR, q.K, q.kernel_size[0], hp.radius2)
    b.record(); torch.cuda.synchronize()
    print("query ms", a.elapsed_time(b) / 10, "valid samples", int(out[2].sum()), "valid tuples", int((out[0] >= 0).sum()))


if __name__ == "__main__":
    main()
