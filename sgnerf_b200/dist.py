"""Multi-GPU plumbing of the render path: one process per GPU (torch.distributed; NCCL over NVLink on GPUs, gloo in the CPU tests).

Rays are independent, so the path shards by rays with the point cloud and the occupancy grid replicated on every rank and NO
collective on the render path.  Training has one real exchange step: the gradients of the aggregator MLP
weights and of the point tables are summed over ranks, as one flat fp32 bucket, before the (identical) optimiser step on every rank.
The reference has nothing here (its DataParallel stub is unused).
"""
import torch
import torch.distributed as dist


def shard_rays(n_rays, rank, world, tile=256):
    """Indices of the rays rank `rank` renders: tiles of `tile` consecutive rays dealt round-robin, which balances hit/miss
    regions of a frame across ranks.  The union over ranks is exactly range(n_rays), each ray once."""
    if world <= 1:
        return torch.arange(n_rays)
    n_tiles = (n_rays + tile - 1) // tile
    mine = torch.arange(rank, n_tiles, world)
    idx = (mine[:, None] * tile + torch.arange(tile)[None, :]).reshape(-1)
    return idx[idx < n_rays]


def gather_frame(local_rgb, idx, n_rays, group=None):
    """Assemble the full frame on every rank from the per-rank ray shards (inference convenience; 12 bytes per ray)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = torch.zeros(n_rays, local_rgb.shape[-1], dtype=local_rgb.dtype, device=local_rgb.device)
    out[idx.to(local_rgb.device)] = local_rgb
    if world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)       # shards are disjoint, so the sum is the concatenation
    return out


def allreduce_grads(params, average=True, group=None):
    """Sum (or average) the .grad of `params` over all ranks through one flat fp32 bucket (one collective launch: the MLP
    gradients are 1.7 MB, the dense point-table gradients 156 MB at N = 1M, so launch latency, not link count, is what to save).
    Parameters without a gradient on this rank contribute zeros; every rank must pass the same parameter list."""
    params = [p for p in params if p is not None and p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    sizes = [p.numel() for p in params]
    dev = params[0].device
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    off = 0
    for p, n in zip(params, sizes):
        if p.grad is not None:
            flat[off:off + n].copy_(p.grad.reshape(-1))
        off += n
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
    off = 0
    for p, n in zip(params, sizes):
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel() * 4
