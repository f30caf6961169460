#!/usr/bin/env python
"""The pieces of the training step's touched-row exchange, timed one by one with CUDA events (max over ranks):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_exchange.py

marks all-reduce (4 B / point), sgn_rows_union, sgn_rows_pack, the all-reduce of [MLP gradients | packed rows], sgn_rows_pack (unpack),
and the dense all-reduce of the whole bucket they replace.  Rows: a fraction --touched of --points per rank, random."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--touched", type=float, default=0.05)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    from sgnerf_b200 import ops
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    N, Cs, n_net, stride = args.points, [32, 3, 3, 1], 425_000, 40
    g = torch.Generator(device=dev).manual_seed(rank)
    tabs = [torch.randn(N, c, device=dev, generator=g) if c > 1 else torch.randn(N, device=dev, generator=g) for c in Cs]
    rows = torch.randperm(N, device=dev, generator=g)[: int(N * args.touched)]
    marks = torch.zeros(N + 1, device=dev)
    xbuf = torch.zeros(n_net + N * stride, device=dev)
    lst, cnt = torch.zeros(N, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    dense = torch.zeros(n_net + N * 39 + N, device=dev)
    ar = (lambda t: dist.all_reduce(t)) if world > 1 else (lambda t: None)

    def timed(fn, setup=None):
        for _ in range(3):
            if setup:
                setup()
            fn()
        best = []
        for _ in range(args.reps):
            if setup:
                setup()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(best)[len(best) // 2]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) * 1e3

    def set_marks():
        marks.zero_()
        marks[rows] = 1.0
    out = {"n_gpus": world, "points": N, "touched_per_rank": int(rows.numel())}
    out["marks_allreduce_us"] = timed(lambda: ar(marks), set_marks)
    set_marks(); ar(marks)
    out["rows_union_us"] = timed(lambda: ops.rows_union(marks[:N], lst, cnt))
    n = int(cnt)
    out["union_rows"] = n
    out["rows_pack_us"] = timed(lambda: ops.rows_pack(tabs, lst, cnt, xbuf[n_net:], stride))
    out["packed_allreduce_us"] = timed(lambda: ar(xbuf[: n_net + n * stride]))
    out["packed_allreduce_bytes"] = (n_net + n * stride) * 4
    if world > 1:
        # the same bytes as two halves on two communicators / two streams at once (does one all-reduce of this size fill NVLink?)
        g2 = dist.new_group()
        side = torch.cuda.Stream()
        half = (n_net + n * stride) // 2 // 64 * 64
        def split():
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dist.all_reduce(xbuf[half: n_net + n * stride], group=g2)
            dist.all_reduce(xbuf[:half])
            main.wait_stream(side)
        out["packed_allreduce_two_groups_us"] = timed(split)
        quarter = half // 2 // 64 * 64
        out["packed_allreduce_half_us"] = timed(lambda: ar(xbuf[:half]))
        out["packed_allreduce_quarter_us"] = timed(lambda: ar(xbuf[:quarter]))
    out["rows_unpack_us"] = timed(lambda: ops.rows_pack(tabs, lst, cnt, xbuf[n_net:], stride, unpack=True))
    out["dense_allreduce_us"] = timed(lambda: ar(dense))
    out["dense_allreduce_bytes"] = dense.numel() * 4
    out["scalar_allreduce_us"] = timed(lambda: ar(marks[N:]))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
