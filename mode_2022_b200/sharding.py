"""Multi-GPU partitioning of the stereo stage.

Work items are (frame, camera pair); the 6 pairs of a frame are independent until fusion and frames are independent, so
the path shards with NO data-path collective: rank r takes items {i : i mod world == r} (weights replicated, 22 MB).
The only exchange is the all-gather of the per-pair disparity / confidence maps into the fusion stage
(2 x 2.1 MB per pair), issued on the stream that produced them.  The reference's only parallelism is nn.DataParallel
(train_disparity.py:264-265); this replaces it with one process per GPU + torch.distributed (NCCL on GPUs, gloo in tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_items(n_items: int, rank: int, world: int) -> List[int]:
  """Round-robin ownership: item i belongs to rank i % world."""
  if not (0 <= rank < world):
    raise ValueError('rank must be in [0, world)')
  return list(range(rank, n_items, world))


def max_shard(n_items: int, world: int) -> int:
  return (n_items + world - 1) // world


def gather_maps(local: torch.Tensor, n_items: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
  """All-gather per-item maps.  `local`: (n_local, C, H, W) for this rank's items (in shard order);
  returns (n_items, C, H, W) in global item order on every rank.  Ranks with fewer items are padded to the
  maximum shard size so a single fixed-shape collective is issued."""
  world = dist.get_world_size(group) if dist.is_initialized() else 1
  rank = dist.get_rank(group) if dist.is_initialized() else 0
  if world == 1:
    return local
  m = max_shard(n_items, world)
  if local.shape[0] > m:
    raise ValueError('local shard larger than the maximum shard size')
  buf = local.new_zeros((m, *local.shape[1:]))
  buf[:local.shape[0]] = local
  out = [torch.empty_like(buf) for _ in range(world)]
  dist.all_gather(out, buf, group=group)
  full = local.new_empty((n_items, *local.shape[1:]))
  for r in range(world):
    idx = shard_items(n_items, r, world)
    if idx:
      full[idx] = out[r][:len(idx)]
  return full


def run_sharded(model, left: torch.Tensor, right: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
  """Stereo stage over all (frame x pair) items with the work split across the ranks of `group`.
  left/right: (n_items, 3, H, W), identical on every rank (or only this rank's rows need to be valid).
  Returns (pred, conf), each (n_items, 1, H, W), complete on every rank -- the input of the fusion stage."""
  world = dist.get_world_size(group) if dist.is_initialized() else 1
  rank = dist.get_rank(group) if dist.is_initialized() else 0
  n = left.shape[0]
  mine = shard_items(n, rank, world)
  if mine:
    out = model(left[mine], right[mine])
    pred, conf = out if isinstance(out, tuple) else (out, torch.zeros_like(out))
  else:
    pred = left.new_zeros((0, 1, *left.shape[2:]))
    conf = left.new_zeros((0, 1, *left.shape[2:]))
  both = gather_maps(torch.cat([pred, conf], 1), n, group)
  return both[:, :1], both[:, 1:]
