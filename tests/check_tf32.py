#!/usr/bin/env python
"""Per-tensor error report of the TF32 tensor-core training path against the fp32 SIMT path (debug aid for gemm_tc.cuh; test
infrastructure: it uses the oracle's parameter initialiser).  python tests/check_tf32.py [semantic]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import render_ref as rr  # noqa: E402
from sgnerf_b200 import ops  # noqa: E402
from tests.test_gpu_aggregate import _random_case, cfg_to_c, param_lists, rel_l2  # noqa: E402


def main():
    semantic = len(sys.argv) > 1 and sys.argv[1] == "semantic"
    R, SR, K, N = (int(os.environ.get("R", "700")), 24, 8, 5000)
    cfg = rr.semantic_config() if semantic else rr.agg_config()
    tables, pidx, loc_w, raydir, campos, rot = _random_case(cfg, N, R, SR, K, seed=21, prefix_mask=True)
    P = rr.init_params(cfg, seed=4, bias_scale=0.1)
    g = torch.Generator().manual_seed(6)
    cot_d, cot_c = torch.randn(R, SR, 4, generator=g).cuda(), (torch.randn(R, SR, K, generator=g) * 0.1).cuda()
    res = {}
    for prec in (ops.PRECISION_FP32, ops.PRECISION_TF32):
        names, W, B = param_lists(P, cfg, requires_grad=True)
        tc = {k: getattr(tables, k).clone().cuda().requires_grad_(True) for k in ("embedding", "color", "dir", "conf")}
        lab = tables.label_embedding.cuda() if semantic else None
        with torch.no_grad():
            d0 = ops.aggregate(cfg_to_c(cfg), W, B, tables.xyz.cuda(), tc["embedding"], tc["color"], tc["dir"], tc["conf"],
                               lab, pidx.cuda(), loc_w.cuda(), raydir.cuda(), campos.cuda(), rot.cuda(), precision=prec)[0]
        torch.cuda.synchronize()
        print("precision", prec, "forward (no grad) ok", float(d0.abs().max()), flush=True)
        dec, valid, _, w, conf = ops.aggregate(cfg_to_c(cfg), W, B, tables.xyz.cuda(), tc["embedding"], tc["color"], tc["dir"], tc["conf"],
                                               lab, pidx.cuda(), loc_w.cuda(), raydir.cuda(), campos.cuda(), rot.cuda(), precision=prec)
        torch.cuda.synchronize()
        print("precision", prec, "forward (save) ok; valid tuples", int((pidx >= 0).sum()), flush=True)
        ((dec * cot_d).sum() + (conf * cot_c).sum()).backward()
        torch.cuda.synchronize()
        print("precision", prec, "backward ok", flush=True)
        res[prec] = (d0, dec.detach(), {k: v.grad for k, v in tc.items()}, [x.grad for x in W], [x.grad for x in B])
    a, b = res[ops.PRECISION_FP32], res[ops.PRECISION_TF32]
    print("decoded(nograd) max abs diff", float((a[0] - b[0]).abs().max()), " decoded(save) max abs diff", float((a[1] - b[1]).abs().max()))
    for k in a[2]:
        print(f"  d_{k:10s} rel_l2 {rel_l2(b[2][k], a[2][k]):.3e}")
    for i, n in enumerate(names):
        print(f"  {n:22s} W {tuple(a[3][i].shape)} rel_l2 {rel_l2(b[3][i], a[3][i]):.3e}   b rel_l2 {rel_l2(b[4][i], a[4][i]):.3e}")


if __name__ == "__main__":
    main()
