"""Model initialisation / partial checkpoint loading -- same API as the reference models/initModel.py."""
from __future__ import annotations

import torch
import torch.nn as nn

from .sphere_conv import SphereConv

_INITS = {
    'kaiming_normal': lambda w: nn.init.kaiming_normal_(w, mode='fan_in', nonlinearity='leaky_relu'),
    'xavier_normal': nn.init.xavier_normal_,
    'kaiming_uniform': lambda w: nn.init.kaiming_uniform_(w, mode='fan_in', nonlinearity='leaky_relu'),
    'xavier_uniform': nn.init.xavier_uniform_,
    'normal': nn.init.normal_,
}


def initModelPara(model, initType):
  """Re-initialise conv / SphereConv / BN / Linear parameters (reference initModel.py:9-32)."""
  if initType is None or initType == 'default':
    return
  for m in model.modules():
    if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d, SphereConv)):
      if initType in _INITS:
        _INITS[initType](m.weight)
      if m.bias is not None:
        nn.init.constant_(m.bias, 0)
    elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
      nn.init.constant_(m.weight, 1)
      nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Linear):
      nn.init.normal_(m.weight, 0, 0.01)
      if m.bias is not None:
        nn.init.constant_(m.bias, 0)
  if hasattr(model, 'invalidate_plan'):
    model.invalidate_plan()


def loadStackHourglassOnly(model, savedDictPath):
  """Load every non-feature-extraction tensor of a PSMNet checkpoint (reference initModel.py:35-42)."""
  pretrained = torch.load(savedDictPath, map_location='cpu')['state_dict']
  current = model.state_dict()
  current.update({k: v for k, v in pretrained.items() if k in current and 'feature_extraction' not in k and 'forfilter1' not in k})
  model.load_state_dict(current)
