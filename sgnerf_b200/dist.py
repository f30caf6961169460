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


_FRAME_PERM = {}


def _frame_layout(n_rays, world, tile):
    """(rays of the largest shard, for every ray its position in the rank-major concatenation of the padded shards)."""
    key = (int(n_rays), int(world), int(tile))
    if key not in _FRAME_PERM:
        shards = [shard_rays(n_rays, r, world, tile) for r in range(world)]
        per = max(int(s.numel()) for s in shards)
        perm = torch.empty(n_rays, dtype=torch.long)
        for r, s in enumerate(shards):
            perm[s] = r * per + torch.arange(s.numel())
        _FRAME_PERM[key] = (per, perm)
    return _FRAME_PERM[key]


def gather_frame(local_rgb, idx, n_rays, group=None, tile=256):
    """Assemble the full frame on every rank from the per-rank ray shards of shard_rays(n_rays, rank, world, tile): ONE all-gather of
    the (padded) shards -- every byte crosses the links once, 12 bytes per ray -- then one indexed copy into frame order.
    `idx` is this rank's shard_rays(...) (kept for the single-process case and for checking)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world <= 1:
        out = torch.empty(n_rays, local_rgb.shape[-1], dtype=local_rgb.dtype, device=local_rgb.device)
        out[idx.to(local_rgb.device)] = local_rgb
        return out
    per, perm = _frame_layout(n_rays, world, tile)
    C_ = local_rgb.shape[-1]
    mine = local_rgb
    if mine.shape[0] != per:                                  # the last tiles may leave a shard a few rays short: pad to the common size
        mine = torch.zeros(per, C_, dtype=local_rgb.dtype, device=local_rgb.device)
        mine[:local_rgb.shape[0]] = local_rgb
    allr = torch.empty(world * per, C_, dtype=local_rgb.dtype, device=local_rgb.device)
    dist.all_gather_into_tensor(allr, mine.contiguous(), group=group)
    key = ("perm_dev", n_rays, world, tile, str(local_rgb.device))
    if key not in _FRAME_PERM:
        _FRAME_PERM[key] = perm.to(local_rgb.device)
    return allr[_FRAME_PERM[key]]


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
