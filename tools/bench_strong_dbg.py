#!/usr/bin/env python
"""Debug aid: C3 frame ray-sharded over the ranks of a torchrun launch, render and gather timed separately."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgnerf_b200 import dist as sdist
from sgnerf_b200 import ops, pipeline, synth  # noqa: E402


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    s = synth.scene_c3()
    tabs = synth.make_point_tables(s.xyz.shape[0], 32, 0, seed=0)
    shapes = synth.mlp_layer_shapes()
    P = synth.make_mlp_params(shapes, seed=0)
    names = [k for k, _, _ in shapes]
    scene = pipeline.RenderScene(s.xyz, tabs.embedding, tabs.color, tabs.dir, tabs.conf, [P[k + ".weight"] for k in names],
                                 [P[k + ".bias"] for k in names], ops.agg_cfg(), pipeline.query_options(**synth.C3_QUERY), device=dev)
    campos, rot = torch.from_numpy(s.campos).to(dev), torch.from_numpy(s.camrotc2w).to(dev)
    R = s.raydir.shape[0]
    idx = sdist.shard_rays(R, rank, world, tile=256)
    mine = torch.from_numpy(s.raydir)[idx].to(dev).contiguous()
    idx_d = idx.to(dev)
    bg = torch.ones(3, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for it in range(6):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, b, c = ev(), ev(), ev()
            a.record()
            part = pipeline.render_rays(scene, campos, rot, mine, s.near, s.far, bg, precision=ops.PRECISION_BF16)
            b.record()
            frame = sdist.gather_frame(part.ray_color, idx_d, R, tile=256)
            c.record(); torch.cuda.synchronize()
            print(f"rank {rank} it {it}: render {a.elapsed_time(b):.2f} ms, gather {b.elapsed_time(c):.2f} ms", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
