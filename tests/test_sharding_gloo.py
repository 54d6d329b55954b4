"""CPU, world_size 2 (gloo): the N>1 path of the stereo stage -- item ownership and the all-gather into fusion."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeStereo(torch.nn.Module):
  """Stands in for ModeDisparity (which needs a GPU): deterministic per-item outputs so ordering can be checked."""

  def forward(self, left, right):
    pred = left[:, :1] * 2 + right[:, :1]
    return pred, pred * 0.5


def _worker(rank, world, port, n_items, results):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from mode_2022_b200 import sharding as S
    g = torch.Generator().manual_seed(0)
    left = torch.randn(n_items, 3, 8, 4, generator=g)
    right = torch.randn(n_items, 3, 8, 4, generator=g)
    pred, conf = S.run_sharded(_FakeStereo(), left, right)
    want = left[:, :1] * 2 + right[:, :1]
    ok = torch.equal(pred, want) and torch.equal(conf, want * 0.5)
    owned = S.shard_items(n_items, rank, world)
    results[rank] = (ok, owned)
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('n_items', [6, 5, 1])
def test_sharded_stage_world2(n_items):
  world = 2
  mgr = mp.Manager()
  results = mgr.dict()
  port = 29500 + (os.getpid() + n_items) % 1000
  mp.spawn(_worker, args=(world, port, n_items, results), nprocs=world, join=True)
  assert all(results[r][0] for r in range(world))
  owned = sorted(i for r in range(world) for i in results[r][1])
  assert owned == list(range(n_items))  # every (frame, pair) item is processed exactly once


def test_shard_items_partition():
  from mode_2022_b200 import sharding as S
  for world in (1, 2, 4, 8):
    for n in (0, 1, 6, 12, 13):
      got = sorted(i for r in range(world) for i in S.shard_items(n, r, world))
      assert got == list(range(n))
      assert max((len(S.shard_items(n, r, world)) for r in range(world)), default=0) == (S.max_shard(n, world) if n else 0)
  with pytest.raises(ValueError):
    S.shard_items(4, 2, 2)
