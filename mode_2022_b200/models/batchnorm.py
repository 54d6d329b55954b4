"""nn.BatchNorm2d / nn.BatchNorm3d whose TRAINING-mode forward and backward run libmode_b200's HBM-bound kernels
(`mode_b200::batch_norm_train`) instead of cuDNN's spatial BN (49 ms of a 182 ms training step at 1024x512, D=192).

Same parameters, buffers, state-dict keys and update rule as torch's modules (reference: models/submodule.py:14-30,
models/mode_disparity.py:11-46 build plain nn.BatchNorm2d / nn.BatchNorm3d; `isinstance(m, nn.BatchNorm2d)` still holds).  Everything
the kernels do not cover -- eval mode, CPU tensors, non-fp32 input, `momentum=None` (cumulative average, used by the BN calibration
of the parity fixtures), `track_running_stats=False` -- goes through the parent class unchanged."""
from __future__ import annotations

import torch
import torch.nn as nn


class _TrainKernelMixin:

  def forward(self, x: torch.Tensor) -> torch.Tensor:
    if not (self.training and x.is_cuda and x.dtype == torch.float32 and self.momentum is not None and self.track_running_stats and x.numel() > 0):
      return super().forward(x)
    from .. import ops
    self._check_input_dim(x)
    if self.num_batches_tracked is not None:
      self.num_batches_tracked.add_(1)
    y, mean, _, var_u = ops.batch_norm_train(x, self.weight, self.bias, float(self.eps))
    with torch.no_grad():  # torch's rule: running = (1 - momentum) * running + momentum * batch statistic (unbiased variance)
      self.running_mean.lerp_(mean.detach(), float(self.momentum))
      self.running_var.lerp_(var_u.detach(), float(self.momentum))
    return y


class BatchNorm2d(_TrainKernelMixin, nn.BatchNorm2d):
  pass


class BatchNorm3d(_TrainKernelMixin, nn.BatchNorm3d):
  pass
