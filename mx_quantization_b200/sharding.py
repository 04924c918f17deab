"""Batch x heads sharding of the hot path across the GPUs of one box.

Every (batch, head) pair is independent through the whole path, so ranks take
contiguous batch slices and nothing is exchanged on the data path.  NCCL (or gloo in the CPU tests)
is used only to gather outputs / mask digests for verification, outside any timed region.
"""
from typing import List, Tuple

import torch


def shard_bounds(n_units: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n_units for `rank`; the first (n_units % world) ranks get one
    extra unit.  Empty slices are allowed (more ranks than units)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world {world_size}")
    base, extra = divmod(n_units, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch_heads(B: int, H: int, world_size: int, rank: int):
    """Shard the batch first (keeps the fused qkv slice contiguous); when there are more ranks
    than batch entries (long-sequence sweep, B < G) shard the flattened (batch, head) units.
    Returns ("batch", lo, hi) or ("heads", lo, hi) with lo/hi indexing B resp. B*H."""
    if B >= world_size:
        lo, hi = shard_bounds(B, world_size, rank)
        return "batch", lo, hi
    lo, hi = shard_bounds(B * H, world_size, rank)
    return "heads", lo, hi


def take_shard(t: torch.Tensor, B: int, H: int, world_size: int, rank: int) -> torch.Tensor:
    """The rank's slice of a (B,H,N,hd) tensor/view, as a (b',h',N,hd) view (no copy for batch
    sharding; head sharding flattens (B,H) -> units, which copies only if the view cannot)."""
    kind, lo, hi = shard_batch_heads(B, H, world_size, rank)
    if kind == "batch":
        return t[lo:hi]
    flat = t.reshape(B * H, 1, *t.shape[2:])
    return flat[lo:hi]


def gather_for_verification(local: torch.Tensor, world_size: int, group=None) -> List[torch.Tensor]:
    """all_gather of (possibly ragged) per-rank outputs; verification only."""
    import torch.distributed as dist
    if world_size == 1 or not dist.is_initialized():
        return [local]
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world_size)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(outs, pad, group=group)
    return [o[: int(s.item())] for o, s in zip(outs, sizes)]
