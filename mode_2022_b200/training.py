"""Data-parallel training step of the stereo stage: one process per GPU + bucketed, overlapped NCCL gradient all-reduce.

The reference trains under `nn.DataParallel` (train_disparity.py:264-265): one Python process scatters the batch over the GPUs,
re-broadcasts the 22 MB of parameters EVERY forward, gathers the outputs to GPU 0, computes the loss there
(train_disparity.py:147-158) and `reduce_add_coalesced`s the replica gradients back to GPU 0.  Here every rank owns a shard of the
batch and its own replica; the only exchange is the gradient all-reduce, issued per bucket AS SOON AS the
bucket's gradients exist, so that NCCL over NVLink runs underneath the rest of the backward pass (the conv3d stack produces its
gradients last-layer-first; the feature extractor's arrive at the very end and form the last bucket).

  reducer = GradAllReduce(model.parameters())          # once
  loss = global_masked_loss(preds, disp_true, mask)    # reference loss, normalised by the GLOBAL mask count
  loss.backward()                                      # buckets all-reduce while autograd is still running
  reducer.finish()                                     # wait; .grad of every parameter now holds the global gradient
  optimizer.step(); reducer.zero_grad()

Semantics match the reference's DataParallel step: the loss is the mean over ALL masked pixels of the global batch (each rank
divides its partial sum by the all-reduced pixel count, gradients are SUMMED), BatchNorm statistics stay per replica
(DataParallel does not synchronise them either).  Works with gloo on CPU (tests) and NCCL on GPUs.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn.functional as F


def _world(group=None):
  return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class GradAllReduce:
  """Flat gradient buckets (parameters' `.grad` are views into them) + one asynchronous all-reduce per bucket, launched from a
  post-accumulate-grad hook when the last gradient of the bucket has been written.

  bucket_bytes: target bucket size.  The whole model is 22 MB of fp32 gradients; NVSwitch makes the collective latency-bound, not
  bandwidth-bound, so a few buckets of ~6 MB overlap best: small enough that the first one leaves while the
  conv3d backward is still running, large enough that launch latency does not add up."""

  def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None, bucket_bytes: int = 6 << 20):
    self.group = group
    self.world = _world(group)
    self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
    if not self.params:
      raise ValueError('GradAllReduce: no trainable parameters')
    dev, dt = self.params[0].device, self.params[0].dtype
    if any(p.device != dev or p.dtype != dt for p in self.params):
      raise ValueError('GradAllReduce: all parameters must share one device and dtype')
    # autograd produces gradients roughly in REVERSE registration order: bucket 0 holds the last-registered parameters
    self.buckets: List[dict] = []
    cur, cur_bytes = [], 0
    for p in reversed(self.params):
      cur.append(p)
      cur_bytes += p.numel() * p.element_size()
      if cur_bytes >= bucket_bytes:
        self._close(cur, dev, dt)
        cur, cur_bytes = [], 0
    if cur:
      self._close(cur, dev, dt)
    self._bucket_of = {}
    for bi, b in enumerate(self.buckets):
      for p in b['params']:
        self._bucket_of[p] = bi
    self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
    self._works: List = []
    self.launch_order: List[int] = []  # bucket indices in the order their all-reduce was issued during the last backward (diagnostics)

  def _close(self, plist: Sequence[torch.nn.Parameter], dev, dt):
    n = sum(p.numel() for p in plist)
    flat = torch.zeros(n, device=dev, dtype=dt)
    off = 0
    for p in plist:
      p.grad = flat[off:off + p.numel()].view_as(p)  # autograd accumulates in place into the view
      off += p.numel()
    self.buckets.append({'params': list(plist), 'flat': flat, 'pending': len(plist), 'n': len(plist)})

  def _hook(self, p: torch.nn.Parameter):
    b = self.buckets[self._bucket_of[p]]
    if p.grad.data_ptr() < b['flat'].data_ptr() or p.grad.data_ptr() >= b['flat'].data_ptr() + b['flat'].numel() * b['flat'].element_size():
      raise RuntimeError('GradAllReduce: a parameter gradient was re-allocated outside its bucket (use reducer.zero_grad(), not set_to_none)')
    b['pending'] -= 1
    if b['pending'] == 0:
      self.launch_order.append(self._bucket_of[p])
      if self.world > 1:
        # async_op: NCCL's stream waits for the gradients written so far and runs underneath the remaining backward kernels
        self._works.append(dist.all_reduce(b['flat'], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

  def finish(self):
    """Wait for every bucket (the current stream then sees the reduced gradients).  Buckets whose gradients never all arrived
    (unused parameters in this step) are reduced here, so every rank issues the same collectives."""
    for bi, b in enumerate(self.buckets):
      if b['pending'] != 0:
        self.launch_order.append(bi)
        if self.world > 1:
          self._works.append(dist.all_reduce(b['flat'], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
    for w in self._works:
      w.wait()
    self._works = []
    for b in self.buckets:
      b['pending'] = b['n']

  def zero_grad(self):
    for b in self.buckets:
      b['flat'].zero_()
      b['pending'] = b['n']
    self.launch_order = []

  def grad_bytes(self) -> int:
    return sum(b['flat'].numel() * b['flat'].element_size() for b in self.buckets)

  def close(self):
    for h in self._handles:
      h.remove()
    self._handles = []


def global_masked_loss(outputs: Sequence[torch.Tensor], disp_true: torch.Tensor, mask: torch.Tensor, weights: Sequence[float] = (0.5, 0.7, 1.0),
                       group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
  """The reference's training loss (train_disparity.py:152-158): 0.5/0.7/1.0-weighted smooth-L1 over the masked pixels of the three
  heads -- as DataParallel computes it, i.e. the MEAN OVER THE GLOBAL BATCH: every rank sums over its own masked pixels and divides
  by the all-reduced pixel count (one 8-byte collective), so that SUM-reduced gradients equal the single-process gradient exactly."""
  n = mask.sum().to(torch.float64)
  if _world(group) > 1:
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
  n = n.clamp_min(1.0).to(outputs[0].dtype)
  tgt = disp_true[mask]
  loss = outputs[0].new_zeros(())
  for w, o in zip(weights, outputs):
    loss = loss + w * F.smooth_l1_loss(o[mask], tgt, reduction='sum') / n
  return loss


def train_step(model, reducer: GradAllReduce, optimizer, left, right, disp_true, mask, group=None) -> torch.Tensor:
  """trainDisp of the reference (train_disparity.py:147-163) on this rank's shard of the batch; returns the (local share of the) loss."""
  model.train()
  reducer.zero_grad()
  outs = model(left, right)
  loss = global_masked_loss(outs, disp_true, mask, group=group)
  loss.backward()
  reducer.finish()
  optimizer.step()
  return loss.detach()
