#!/usr/bin/env python
"""Stage timings of the render path at config C1 (or a smaller frame), for kernel work.

    python tools/bench_agg.py [--precision bf16|fp32] [--iters 5] [--semantic] [--width 640 --height 480]

Prints one line per stage (query / aggregate / ray_dist+composite) with CUDA-event times, so a kernel change
can be judged in one short gpurun call.  SGN_TC_DEBUG=<bitmask> switches off parts of the tensor-core
kernel (timing experiments only).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--semantic", action="store_true")
    ap.add_argument("--once", action="store_true", help="one aggregate call only (for ncu)")
    args = ap.parse_args()
    from sgnerf_b200 import ops, pipeline, synth
    dev = "cuda:0"
    s = synth.scene_room(args.points, room=(8.0, 8.0, 3.0), width=args.width, height=args.height, seed=1234)
    ld = 96 if args.semantic else 0
    tabs = synth.make_point_tables(args.points, 32, ld, seed=0)
    shapes = synth.mlp_layer_shapes(layers2_bpnet=1 if args.semantic else 0, label_dim=ld)
    P = synth.make_mlp_params(shapes, seed=0)
    names = [n for n, _, _ in shapes]
    cfg = ops.agg_cfg(n_block2_bpnet=1 if args.semantic else 0, label_dim=ld)
    scene = pipeline.RenderScene(s.xyz, tabs.embedding, tabs.color, tabs.dir, tabs.conf, [P[n + ".weight"] for n in names],
                                 [P[n + ".bias"] for n in names], cfg, pipeline.query_options(SR=24), label_emb=tabs.label_embedding,
                                 device=dev)
    campos, rot = torch.from_numpy(s.campos).to(dev), torch.from_numpy(s.camrotc2w).to(dev)
    raydir = torch.from_numpy(s.raydir).to(dev)
    precision = {"fp32": ops.PRECISION_FP32, "bf16": ops.PRECISION_BF16}[args.precision]
    grid, hp = scene.grid()
    q = scene.qopt
    t = pipeline.middle_point_ts(s.near, s.far, q.z_depth_dim, dev)

    def timed(fn, n):
        out = None
        ms = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = fn(); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return out, ms

    with torch.no_grad():
        (pidx, loc_w, smask, rmask), q_ms = timed(lambda: ops.query(grid, campos, raydir, t, q.SR, q.K, q.kernel_size[0], hp.radius2),
                                                  1 if args.once else args.iters)
        T_v = int((pidx >= 0).sum()); S_v = int((pidx >= 0).any(-1).sum())
        agg = lambda: ops.aggregate(scene.agg_cfg, scene.weights, scene.biases, scene.xyz, scene.embedding, scene.color, scene.dirs,
                                    scene.conf, scene.label_emb, pidx, loc_w, raydir, campos, rot, precision=precision, want_aux=False)
        (decoded, ray_valid, loc_pers, _, _), a_ms = timed(agg, 1 if args.once else args.iters)
        if args.once:
            return

        def comp():
            rd = ops.ray_dist(loc_pers, ray_valid, hp.vsize[2], 1)
            return ops.composite(decoded, rd, ray_valid, torch.ones(3, device=dev), blend=0)
        _, c_ms = timed(comp, args.iters)
    per_t = 722944 if args.semantic else 542720
    flops = T_v * per_t + S_v * 137984
    best = min(a_ms[1:]) if len(a_ms) > 1 else a_ms[0]
    print(f"dbg={os.environ.get('SGN_TC_DEBUG', '0')} R={raydir.shape[0]} T_v={T_v} S_v={S_v}  query {min(q_ms):.3f} ms | aggregate {best:.3f} ms "
          f"({flops / best / 1e9:.1f} TFLOP/s) | ray_dist+composite {min(c_ms):.3f} ms   all agg: {[round(x, 2) for x in a_ms]}")


if __name__ == "__main__":
    main()
