#!/usr/bin/env python
"""Perspective-frustum querier (--wcoord_query 0, sgn_pers_query) on the C1 frame: 1M points, 640x480 rays, vscale 2, kernel 3^3, SR 24, K 8.
Prints ms per call (the grid is rebuilt per call: it lives in the camera's coordinates) and the neighbour statistics."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from sgnerf_b200 import modules, ops, synth
    dev = "cuda"
    s = synth.scene_room(1_000_000, room=(8.0, 8.0, 3.0), width=640, height=480, seed=1234)
    K = synth.SCANNET_INTRINSIC.copy()
    K[0, 2], K[1, 2] = (640 - 1) / 2.0, (480 - 1) / 2.0
    xyz = torch.from_numpy(s.xyz).to(dev)
    campos, rot = torch.from_numpy(s.campos)[None].to(dev), torch.from_numpy(s.camrotc2w)[None].to(dev)
    pix = torch.from_numpy(np.stack([s.px, s.py], -1).astype(np.int32)).to(dev)
    out = {}
    for name, kw in (("canonical", dict(vscale=[2, 2, 2], ks=[3, 3, 3], rl=4.0, dl=1.3)), ("wide", dict(vscale=[4, 4, 4], ks=[5, 5, 3], rl=16.0, dl=4.0))):
        hp = ops.pers_hyperparameters(480, 640, K, s.near, s.far, 400, kw["vscale"], kw["rl"], kw["dl"])
        def call():
            xyz_pers = modules.lighting_fast_querier.w2pers(xyz[None], rot, campos)[0]
            return ops.pers_query(xyz_pers, pix, hp, kw["ks"], kw["ks"], 24, 8, 16, NN=2)
        for _ in range(3):
            pidx, loc, mask = call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            pidx, loc, mask = call()
        e1.record()
        torch.cuda.synchronize()
        out[name] = dict(ms_per_call=e0.elapsed_time(e1) / 10, rays=int(pix.shape[0]), rays_masked_in=int((mask > 0).sum()),
                         valid_neighbours=int((pidx >= 0).sum()), grid=[int(v) for v in hp.scaled_vdim])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
