"""CPU, world_size 2 (gloo): the data-parallel training step -- bucketed, hook-driven gradient all-reduce and the global masked loss
reproduce the single-process (DataParallel-equivalent) gradients exactly (reference train_disparity.py:147-163, 264-265)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn
import torch.nn.functional as F


class _TinyStereo(nn.Module):
  """Three-head stand-in for ModeDisparity.train() (which needs a GPU): conv trunk without BatchNorm so that sharding the batch
  does not change the function."""

  def __init__(self):
    super().__init__()
    torch.manual_seed(3)
    self.f = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1), nn.ReLU())
    self.h1, self.h2, self.h3 = nn.Conv2d(16, 1, 3, padding=1), nn.Conv2d(16, 1, 3, padding=1), nn.Conv2d(16, 1, 3, padding=1)
    self.unused = nn.Parameter(torch.zeros(5))  # never touched by forward: its bucket must still be reduced on every rank

  def forward(self, left, right):
    x = torch.cat([self.f(left), self.f(right)], 1)
    return self.h1(x), self.h2(x), self.h3(x)


def _data(n):
  g = torch.Generator().manual_seed(11)
  left, right = torch.randn(n, 3, 12, 10, generator=g), torch.randn(n, 3, 12, 10, generator=g)
  disp = torch.rand(n, 1, 12, 10, generator=g) * 20
  mask = torch.rand(n, 1, 12, 10, generator=g) < 0.7
  mask[0] = False  # a sample without valid pixels: the per-rank pixel counts differ, only the GLOBAL mean is right
  mask[0, 0, 0, :3] = True
  return left, right, disp, mask


def _reference_grads(n):
  m = _TinyStereo()
  left, right, disp, mask = _data(n)
  o1, o2, o3 = m(left, right)
  tgt = disp[mask]
  loss = 0.5 * F.smooth_l1_loss(o1[mask], tgt) + 0.7 * F.smooth_l1_loss(o2[mask], tgt) + F.smooth_l1_loss(o3[mask], tgt)  # train_disparity.py:152-158
  loss.backward()
  return loss.item(), {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()}


def _worker(rank, world, port, n, bucket_bytes, results):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from mode_2022_b200 import training as T
    m = _TinyStereo()
    red = T.GradAllReduce(m.parameters(), bucket_bytes=bucket_bytes)
    left, right, disp, mask = _data(n)
    mine = list(range(rank, n, world))
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    for _ in range(2):  # twice: zero_grad() must keep the bucket views alive
      loss = T.train_step(m, red, opt, left[mine], right[mine], disp[mine], mask[mine])
    total = loss.clone()
    dist.all_reduce(total)
    grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    results[rank] = (total.item(), grads, len(red.buckets), list(red.launch_order), red.grad_bytes())
    with pytest.raises(RuntimeError):  # optimizer.zero_grad(set_to_none=True) would silently detach the buckets: detected
      for p in m.parameters():
        p.grad = None
      o = m(left[mine], right[mine])
      T.global_masked_loss(o, disp[mine], mask[mine]).backward()
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('bucket_bytes', [1 << 30, 600])
def test_bucketed_allreduce_matches_single_process_world2(bucket_bytes):
  world, n = 2, 4
  want_loss, want = _reference_grads(n)
  mgr = mp.Manager()
  results = mgr.dict()
  port = 29600 + (os.getpid() + bucket_bytes) % 1000
  mp.spawn(_worker, args=(world, port, n, bucket_bytes, results), nprocs=world, join=True)
  for r in range(world):
    loss, grads, nb, order, gbytes = results[r]
    assert abs(loss - want_loss) <= 1e-6 * max(1.0, abs(want_loss))
    for k, g in grads.items():
      assert torch.allclose(g, want[k], rtol=1e-5, atol=1e-7), (r, k)
    assert sorted(order) == list(range(nb)) and gbytes == sum(v.numel() * 4 for v in want.values())
    if bucket_bytes == 600:
      assert nb > 3 and order[0] == 0  # the heads (registered last) leave first, while the trunk's backward is still running
