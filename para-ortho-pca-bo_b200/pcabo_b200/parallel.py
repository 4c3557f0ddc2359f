"""Multi-GPU plumbing (one process per GPU, torch.distributed): run sharding and the candidate-sharded argmax.

SURVEY 8(e): independent (function, instance) runs shard over ranks with NO data-path collective; one large log-EI
candidate sweep is split into contiguous index ranges per rank and reduced by a single all_gather of
(value, global index) pairs (16 B per rank -- latency bound, NVLink bandwidth irrelevant), ties resolved to the
lowest global index like torch.argmax.
"""
import torch
import torch.distributed as dist


def shard_runs(runs, rank: int, world: int):
    """Round-robin assignment of independent runs to ranks."""
    return [r for j, r in enumerate(runs) if j % world == rank]


def candidate_range(n_candidates: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of the global candidate index space owned by `rank`."""
    per = (n_candidates + world - 1) // world
    lo = min(rank * per, n_candidates)
    return lo, min(lo + per, n_candidates)


def global_argmax(best_val: torch.Tensor, best_idx: torch.Tensor, group=None):
    """best_val [B] float64 and best_idx [B] int64 (GLOBAL candidate indices) per rank -> global (value, index) [B].

    One all_gather of the packed pairs; works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return best_val, best_idx
    world = dist.get_world_size(group)
    packed = torch.stack([best_val.to(torch.float64), best_idx.to(torch.float64)], dim=-1).contiguous()   # [B, 2]
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    allp = torch.stack(out, dim=0)                      # [world, B, 2]
    vals, idxs = allp[..., 0], allp[..., 1]
    vmax = vals.max(dim=0).values
    cand = torch.where(vals == vmax, idxs, torch.full_like(idxs, float("inf")))
    imin = cand.min(dim=0).values
    return vmax, imin.to(torch.int64)
