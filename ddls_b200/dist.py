"""Multi-GPU plumbing: episodes are independent (the reference's memo is per env instance, RCE:269-275), so the
batch shards by contiguous episode ranges, one process per GPU, with NO data-path collective.  The only exchange
is one all-gather of the per-episode metric rows per batch step, over NCCL on GPUs (gloo in the CPU
tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_episodes: int, rank: int, world: int):
    """Contiguous range [lo, hi) of global episode ids owned by `rank`: [g*B/G, (g+1)*B/G)."""
    if not (0 <= rank < world):
        raise Exception(f'rank {rank} not in [0, {world})')
    lo = (n_episodes * rank) // world
    hi = (n_episodes * (rank + 1)) // world
    return lo, hi


def shard_sizes(n_episodes: int, world: int):
    return [shard_range(n_episodes, r, world)[1] - shard_range(n_episodes, r, world)[0] for r in range(world)]


def gather_episode_metrics(local_rows: torch.Tensor, n_episodes: int = None) -> torch.Tensor:
    """All-gathers [B_local, K] metric rows into [B_global, K] in global episode order.

    Uniform shards use all_gather_into_tensor (one NCCL all-gather); ragged shards pad to the largest shard."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_rows
    world = dist.get_world_size()
    if n_episodes is None:
        n_episodes = local_rows.shape[0] * world
    sizes = shard_sizes(n_episodes, world)
    if len(set(sizes)) == 1:
        out = torch.empty((n_episodes,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out, local_rows.contiguous())
        return out
    m = max(sizes)
    padded = torch.zeros((m,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    padded[:local_rows.shape[0]] = local_rows
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
