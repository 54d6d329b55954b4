"""Host-to-host streaming runner for the stereo stage.

The reference's stage driver (save_output_disparity_stage.py:179-199) handles one frame at a time: upload the 6 camera pairs,
run ModeDisparity, download disparity + confidence, write files.  On a B200 the forward of a frame takes ~16 ms while its
75 MB of fp32 images take ~1.5 ms over PCIe and the 25 MB of results ~0.5 ms, so a serial loop leaves the GPU idle ~10 % of
the time.  `HostPipeline` keeps `depth` frames in flight on three streams: the upload of frame i+1 and the download of frame
i-1 overlap the compute of frame i.  Every frame still pays its own H2D and D2H; only their latency is hidden.

  pipe = HostPipeline(model, batch=6, height=1024, width=512)
  t = pipe.submit(left_host, right_host)        # pinned (B,3,H,W) fp32 host tensors; returns a ticket
  pred_host, conf_host = pipe.collect(t)        # pinned (B,1,H,W) host tensors, valid until `depth` more submits

With `use_graph=True` the forward is captured once in a CUDA graph (static shapes) and replayed.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class HostPipeline:
  def __init__(self, model, batch: int, height: int, width: int, depth: int = 2, use_graph: bool = True,
               post: Optional[Callable[[torch.Tensor, torch.Tensor], None]] = None, device=None):
    if model.training:
      raise RuntimeError('HostPipeline runs the inference plan: call model.eval() first')
    if depth < 2:
      raise ValueError('depth must be >= 2 (one frame computing, one moving)')
    self.model, self.depth, self.post = model, depth, post
    self.dev = torch.device(device) if device is not None else next(model.parameters()).device
    if self.dev.type != 'cuda':
      raise NotImplementedError('HostPipeline needs a CUDA model (no CPU fallback)')
    shape_in, shape_out = (batch, 3, height, width), (batch, 1, height, width)
    self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
    self.d_left = [torch.empty(shape_in, device=self.dev) for _ in range(depth)]
    self.d_right = [torch.empty(shape_in, device=self.dev) for _ in range(depth)]
    self.d_pred = [torch.empty(shape_out, device=self.dev) for _ in range(depth)]
    self.d_conf = [torch.empty(shape_out, device=self.dev) for _ in range(depth)]
    self.h_pred = [torch.empty(shape_out).pin_memory() for _ in range(depth)]
    self.h_conf = [torch.empty(shape_out).pin_memory() for _ in range(depth)]
    self.ev_in = [torch.cuda.Event() for _ in range(depth)]    # upload of the slot's frame finished
    self.ev_cmp = [torch.cuda.Event() for _ in range(depth)]   # compute of the slot's frame finished (inputs consumed, outputs ready)
    self.ev_out = [torch.cuda.Event() for _ in range(depth)]   # download of the slot's frame finished
    self.n = 0
    self.graph = None
    with torch.no_grad(), torch.cuda.device(self.dev):
      self._g_left, self._g_right = torch.zeros(shape_in, device=self.dev), torch.zeros(shape_in, device=self.dev)
      out = model(self._g_left, self._g_right)  # builds the plan, warms up allocator and cuDNN heuristics
      self._with_conf = isinstance(out, tuple)
      if use_graph:
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
          model(self._g_left, self._g_right)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
          self._g_out = model(self._g_left, self._g_right)
      torch.cuda.synchronize(self.dev)
    self._plan, self._plan_key = model._plan, model._plan_key  # the captured graph holds THESE folded weights

  def submit(self, left_host: torch.Tensor, right_host: torch.Tensor) -> int:
    """Queue one frame (asynchronous); host tensors must stay untouched until the ticket is collected."""
    if self.graph is not None and (self.model._plan is not self._plan or self.model._weights_key() != self._plan_key):
      raise RuntimeError('HostPipeline: the model weights changed after the CUDA graph was captured; build a new HostPipeline')
    t, k = self.n, self.n % self.depth
    self.n += 1
    with torch.no_grad(), torch.cuda.device(self.dev):
      cur = torch.cuda.current_stream(self.dev)
      self.s_in.wait_event(self.ev_cmp[k])  # the frame that used this slot `depth` submits ago has consumed its inputs
      with torch.cuda.stream(self.s_in):
        self.d_left[k].copy_(left_host, non_blocking=True)
        self.d_right[k].copy_(right_host, non_blocking=True)
        self.ev_in[k].record(self.s_in)
      cur.wait_event(self.ev_in[k])
      cur.wait_event(self.ev_out[k])  # ... and its results have left the slot's output buffers
      if self.graph is not None:
        self._g_left.copy_(self.d_left[k])
        self._g_right.copy_(self.d_right[k])
        self.graph.replay()
        out = self._g_out
      else:
        out = self.model(self.d_left[k], self.d_right[k])
      pred, conf = out if self._with_conf else (out, None)
      if self.post is not None:
        self.post(pred, conf)
      self.d_pred[k].copy_(pred)
      if conf is not None:
        self.d_conf[k].copy_(conf)
      self.ev_cmp[k].record(cur)
      self.s_out.wait_event(self.ev_cmp[k])
      with torch.cuda.stream(self.s_out):
        self.h_pred[k].copy_(self.d_pred[k], non_blocking=True)
        if conf is not None:
          self.h_conf[k].copy_(self.d_conf[k], non_blocking=True)
        self.ev_out[k].record(self.s_out)
    return t

  def collect(self, ticket: int):
    """Block until the frame's results are in host memory; returns (pred, conf) pinned host tensors (conf None without out_conf)."""
    if not (self.n - self.depth <= ticket < self.n):
      raise ValueError(f'ticket {ticket} is no longer (or not yet) in flight')
    k = ticket % self.depth
    self.ev_out[k].synchronize()
    return self.h_pred[k], (self.h_conf[k] if self._with_conf else None)


class FusionStage:
  """Stage 2 of a frame as one replayable unit: the in-memory stage boundary (disparity -> depth, rotation / z-buffer forward warp of
  the 6 pairs into camera 1's frame; reference save_output_disparity_stage.py:105-160 + the npz/PNG files it replaces) followed by
  ModeFusion (reference test_fusion.py:67-102), captured in ONE CUDA graph: ~90 small launches (geometry: ~40, fusion: ~50 cuDNN
  calls) whose launch latency would otherwise be comparable to their run time.

    stage = FusionStage(fusion_model, height=1024, width=512)
    depth = stage(disp6, conf6, rgbs)      # (6,1,H,W), (6,1,H,W), 4 x (1,3,H,W) device tensors -> (1,1,H,W), valid until the next call
  """

  def __init__(self, fusion, height: int, width: int, boundary=None, use_graph: bool = True, device=None):
    from .utils.geometry import StageBoundary
    if fusion.training:
      raise RuntimeError('FusionStage runs the inference plan: call fusion.eval() first')
    self.fusion, self.boundary = fusion, boundary or StageBoundary()
    self.dev = torch.device(device) if device is not None else next(fusion.parameters()).device
    self._disp = torch.zeros(6, 1, height, width, device=self.dev)
    self._conf = torch.zeros(6, 1, height, width, device=self.dev)
    self._rgbs = [torch.zeros(1, 3, height, width, device=self.dev) for _ in range(4)]
    self.graph = None
    with torch.no_grad(), torch.cuda.device(self.dev):
      self._disp.uniform_(1.0, 100.0)  # a plausible disparity field for the warm-up (all-zero disparities mean depth 1000 everywhere)
      out = self._run()  # builds the folded plan, the constant grids and warms up cuDNN
      if use_graph:
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
          self._run()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
          out = self._run()
      self._out = out
      torch.cuda.synchronize(self.dev)

  def _run(self):
    depths, confs = self.boundary(self._disp, self._conf)
    return self.fusion([d.float() for d in depths], [c.float() for c in confs], self._rgbs)

  def __call__(self, disp6: torch.Tensor, conf6: torch.Tensor, rgbs):
    with torch.no_grad(), torch.cuda.device(self.dev):
      self._disp.copy_(disp6)
      self._conf.copy_(conf6)
      for dst, src in zip(self._rgbs, rgbs):
        dst.copy_(src)
      if self.graph is not None:
        self.graph.replay()
        return self._out
      return self._run()
