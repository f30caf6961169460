#!/usr/bin/env python
"""One C3 frame (3M-point object cloud, 800x800, SR 200, P 9, vsize .004) on one GPU: stage times, for kernel work / ncu launch lists."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgnerf_b200 import ops, pipeline, synth  # noqa: E402


def main():
    dev = "cuda"
    frac = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    s = synth.scene_c3()
    tabs = synth.make_point_tables(s.xyz.shape[0], 32, 0, seed=0)
    shapes = synth.mlp_layer_shapes()
    P = synth.make_mlp_params(shapes, seed=0)
    names = [k for k, _, _ in shapes]
    scene = pipeline.RenderScene(s.xyz, tabs.embedding, tabs.color, tabs.dir, tabs.conf, [P[k + ".weight"] for k in names],
                                 [P[k + ".bias"] for k in names], ops.agg_cfg(), pipeline.query_options(**synth.C3_QUERY), device=dev)
    campos, rot = torch.from_numpy(s.campos).to(dev), torch.from_numpy(s.camrotc2w).to(dev)
    n = int(s.raydir.shape[0] * frac)
    if len(sys.argv) > 2:        # emulate rank 0 of `world` ranks: tiles of 256 rays dealt round-robin
        from sgnerf_b200 import dist as sdist
        idx = sdist.shard_rays(s.raydir.shape[0], 0, int(sys.argv[2]), tile=256)
        raydir = torch.from_numpy(s.raydir)[idx].to(dev).contiguous()
        n = raydir.shape[0]
    else:
        raydir = torch.from_numpy(s.raydir)[:n].to(dev)
    bg = torch.ones(3, device=dev)
    with torch.no_grad():
        for _ in range(2):
            o = pipeline.render_rays(scene, campos, rot, raydir, s.near, s.far, bg, precision=ops.PRECISION_BF16)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            o = pipeline.render_rays(scene, campos, rot, raydir, s.near, s.far, bg, precision=ops.PRECISION_BF16)
        b.record(); torch.cuda.synchronize()
    print("C3 rays", n, "ms/frame", a.elapsed_time(b) / 3, "hit", int(o.ray_mask.sum()))


if __name__ == "__main__":
    main()
