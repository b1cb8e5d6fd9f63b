"""Multi-GPU plumbing: pairs are independent, so ranks own contiguous blocks of pairs and the only
collective is one gather of packed matches at the end (SURVEY.md section 8 e).  Works on NCCL (GPU)
and gloo (CPU tensors, used by the world_size-2 tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` pairs owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_matches(matches0: torch.Tensor, num_matches: torch.Tensor) -> torch.Tensor:
    """(B, K) int64 matches + (B,) counts -> one int32 buffer (B, K + 1): column 0 is the count."""
    return torch.cat((num_matches.to(torch.int32)[:, None], matches0.to(torch.int32)), dim=1).contiguous()


def gather_matches(packed: torch.Tensor, per_rank: int, total: int = None) -> torch.Tensor:
    """all_gather of equally-shaped packed buffers (ranks with fewer pairs pad with -2 rows).

    One collective and no host synchronisation: the result is the (world * per_rank, K + 1) buffer itself; when the
    global number of pairs ``total`` is given (contiguous shards, ``shard_range``), the padding rows are dropped
    with index arithmetic done on the host instead of a device-side mask."""
    if not (dist.is_available() and dist.is_initialized()):
        return packed
    world = dist.get_world_size()
    if packed.shape[0] < per_rank:
        pad = packed.new_full((per_rank - packed.shape[0], packed.shape[1]), -2)
        packed = torch.cat((packed, pad), 0)
    out = packed.new_empty((world * per_rank, packed.shape[1]))
    dist.all_gather_into_tensor(out, packed.contiguous())
    if total is None or total == world * per_rank:
        return out
    parts = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        parts.append(out[r * per_rank: r * per_rank + (hi - lo)])
    return torch.cat(parts, 0)
