#!/usr/bin/env python
"""Training step of config C2: 56x56 random-pixel patch per GPU, jittered ray samples, forward + backward
through the aggregator (TF32 tensor-core GEMMs by default, --precision fp32 for the SIMT path) with point-feature scatter gradients, loss = MSE(rgb, gt) + 1e-4 * zero-one(conf), then the
gradient all-reduce (MLP weights + point tables, one flat bucket) and an Adam step.

    python tools/bench_train.py [--steps 10 --warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py
Prints one JSON line on rank 0: ms per step (max over ranks) and its split, rays/s over all ranks.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--patch", type=int, default=56)
    ap.add_argument("--mode", default="graph", choices=["graph", "eager", "modules"],
                    help="graph: sgnerf_b200.train.TrainStep captured in a CUDA graph (no host synchronisation); eager: the same step launched "
                         "kernel by kernel; modules: the reference-shaped NeuralPoints / PointAggregator / ray_march modules with the "
                         "reference's host-side ray compaction")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"], help="GEMM arithmetic of the aggregator's forward + backward")
    args = ap.parse_args()
    from sgnerf_b200 import dist as sdist
    from sgnerf_b200 import modules, synth
    from tests.test_modules import make_opt
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(dev))
    s = synth.scene_room(args.points, room=(8.0, 8.0, 3.0), width=640, height=480, seed=1234)
    tabs = synth.make_point_tables(args.points, 32, 0, seed=0)
    opt = make_opt(is_train=1)
    opt.sgn_precision = "fp32" if args.precision == "fp32" else "auto"
    torch.manual_seed(0)                                     # identical initial weights on every rank
    campos, rot = torch.from_numpy(s.campos)[None].to(dev), torch.from_numpy(s.camrotc2w)[None].to(dev)
    all_rays = torch.from_numpy(s.raydir).to(dev)
    n_patch = args.patch * args.patch
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)      # a different patch per rank
    bg = torch.ones(1, 3, device=dev)
    if args.mode != "modules":
        from sgnerf_b200 import ops, pipeline, train
        agg0 = modules.PointAggregator(opt).to(dev)
        lin = agg0._linears()
        scene = pipeline.RenderScene(torch.from_numpy(s.xyz), tabs.embedding.reshape(args.points, -1), tabs.color.reshape(args.points, 3),
                                     tabs.dir.reshape(args.points, 3), tabs.conf.reshape(args.points),
                                     [m.weight.detach().clone() for m in lin], [m.bias.detach().clone() for m in lin], agg0.cfg,
                                     pipeline.query_options(), device=dev)
        ts = train.TrainStep(scene, n_patch, s.near, s.far, bg, lr=5e-4, plr=2e-3,
                             precision=ops.PRECISION_FP32 if args.precision == "fp32" else ops.PRECISION_TF32,
                             use_graph=args.mode == "graph")
        dgen = torch.Generator(device=dev).manual_seed(200 + rank)

        def step():
            pix = torch.randint(0, all_rays.shape[0], (n_patch,), generator=gen).to(dev, non_blocking=True)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            ts.set_inputs(campos, rot, all_rays[pix], torch.rand(n_patch, 3, device=dev, generator=dgen), ts.jittered_t(0.3, dgen))
            ts.step()
            for e in ev[1:]:
                e.record()
            return ev, int(ts.exchange_floats) * 4, ts.n_hit      # bytes all-reduced by this step (0 on one rank; the touched rows on several)
    else:
        step = None
    if step is None:
        npnts = modules.NeuralPoints(32, args.points, opt, dev, feedforward=1)
        npnts.set_points(torch.from_numpy(s.xyz).to(dev), None, tabs.embedding.to(dev), points_color=tabs.color.to(dev), points_dir=tabs.dir.to(dev),
                         points_conf=tabs.conf.to(dev), parameter=True)
        agg = modules.PointAggregator(opt).to(dev)
        params_net = list(agg.parameters())
        params_pts = [npnts.points_embeding, npnts.points_conf, npnts.points_color, npnts.points_dir]
        optim = torch.optim.Adam([{"params": params_net, "lr": 5e-4}, {"params": params_pts, "lr": 2e-3}])
        blend = lambda o, a: o * a
        blend.__name__ = "alpha_blend"
        render = lambda f: f[..., 1:4]
        render.__name__ = "radiance_render"

        def step():
            pix = torch.randint(0, all_rays.shape[0], (n_patch,), generator=gen).to(dev)
            raydir = all_rays[pix][None]
            gt = torch.rand(1, n_patch, 3, device=dev)
            inputs = {"pixel_idx": None, "camrotc2w": rot, "campos": campos, "near": torch.tensor([s.near]), "far": torch.tensor([s.far]),
                      "h": torch.tensor([480]), "w": torch.tensor([640]), "intrinsic": None, "raydir": raydir, "pixel_label": None}
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            out = npnts(inputs)
            decoded, ray_valid, weight, conf = agg(*out[:12], out[13], out[14])
            sample_loc, ray_mask, vsize = out[9], out[12], out[13]
            rd = torch.cummax(sample_loc[..., 2], dim=-1)[0]
            rd = torch.cat([rd[..., 1:] - rd[..., :-1], torch.full((1, rd.shape[1], 1), float(vsize[2]), device=dev)], dim=-1)
            m = torch.logical_or(rd < 1e-8, rd > 2 * float(vsize[2])).float()
            rd = (rd * (1.0 - m) + m * float(vsize[2])) * ray_valid.float()
            ray_color = modules.ray_march(rd, ray_valid, decoded, render, blend, bg)[0]
            sel = ray_mask[0] > 0
            loss = ((ray_color - gt[:, sel]) ** 2).mean() + 1e-6 + 1e-4 * torch.mean(torch.log(conf.reshape(-1).clamp(1e-3, 1 - 1e-3)) + torch.log(1.0 - conf.reshape(-1).clamp(1e-3, 1 - 1e-3)))
            optim.zero_grad(set_to_none=False)
            loss.backward()
            ev[1].record()
            nbytes = sdist.allreduce_grads(params_net + params_pts, average=True)
            ev[2].record()
            optim.step()
            ev[3].record()
            return ev, nbytes, sel.sum()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    tot, parts, hits = 0.0, np.zeros(3), 0.0
    for _ in range(args.steps):
        ev, nbytes, nh = step()
        torch.cuda.synchronize()
        parts += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
        tot += ev[0].elapsed_time(ev[3]); hits += float(nh)
    t = torch.tensor([tot / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({"metric": "train step ms (C2: 56x56 rays per GPU, fwd+bwd+allreduce+Adam)", "value": ms, "unit": "ms", "n_gpus": world,
                          "rays_per_s": world * n_patch / (ms * 1e-3), "fwd_bwd_ms": parts[0] / args.steps, "allreduce_ms": parts[1] / args.steps,
                          "adam_ms": parts[2] / args.steps, "allreduce_bytes": nbytes, "rays_hit_per_step": hits / args.steps, "dtype": args.precision, "mode": args.mode,
                          "points": args.points, "higher_is_better": False}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
