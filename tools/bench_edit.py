#!/usr/bin/env python
"""C4 point-edit step (prune 2 % + grow 1 % of a 10M-point cloud, then the 1296x968 frame): device time of every piece."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgnerf_b200 import ops, pipeline, synth  # noqa: E402


def main():
    dev = "cuda"
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    s = synth.scene_c4(n)
    tabs = synth.make_point_tables(n, 32, 0, seed=0, conf_spread=0.5)
    shapes = synth.mlp_layer_shapes()
    P = synth.make_mlp_params(shapes, seed=0)
    names = [k for k, _, _ in shapes]
    scene = pipeline.RenderScene(s.xyz, tabs.embedding, tabs.color, tabs.dir, tabs.conf, [P[k + ".weight"] for k in names],
                                 [P[k + ".bias"] for k in names], ops.agg_cfg(), pipeline.query_options(SR=24), device=dev)
    campos, rot = torch.from_numpy(s.campos).to(dev), torch.from_numpy(s.camrotc2w).to(dev)
    raydir = torch.from_numpy(s.raydir).to(dev)
    bg = torch.ones(3, device=dev)
    g = torch.Generator(device=dev).manual_seed(5)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn):
        a, b = ev(), ev()
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b), r

    with torch.no_grad():
        for _ in range(2):
            pipeline.render_rays(scene, campos, rot, raydir, s.near, s.far, bg, precision=ops.PRECISION_BF16)
        scene.dynamic = len(sys.argv) > 2
        for it in range(3):
            t_q, thr = timed(lambda: torch.quantile(scene.conf[:1_000_000], 0.02))
            m = scene.xyz.shape[0] // 100

            def edit():
                base = scene.xyz[torch.randint(0, 1_000_000, (m,), device=dev, generator=g)]
                base = torch.where(base < 1e29, base, torch.zeros_like(base))
                return scene.edit(prune_thresh=thr, add=(base + 0.004 * torch.randn(m, 3, device=dev, generator=g), torch.rand(m, 32, device=dev, generator=g) - 0.5,
                                                         torch.rand(m, 3, device=dev, generator=g),
                                                         torch.nn.functional.normalize(torch.randn(m, 3, device=dev, generator=g), dim=-1),
                                                         0.5 + torch.rand(m, device=dev, generator=g)))
            t_p, _ = timed(edit)
            t_g = 0.0
            t_grid, _ = timed(lambda: scene.grid())
            t_pc, _ = timed(lambda: scene.point_cache())
            t_f, _ = timed(lambda: pipeline.render_rays(scene, campos, rot, raydir, s.near, s.far, bg, precision=ops.PRECISION_BF16))
            print(f"step {it}: quantile {t_q:.2f} | edit (prune + grow, stable indices) {t_p:.2f} | - {t_g:.2f} | grid {t_grid:.2f} | point tables {t_pc:.2f} | frame {t_f:.2f} ms; "
                  f"points {scene.xyz.shape[0]}")


if __name__ == "__main__":
    main()
