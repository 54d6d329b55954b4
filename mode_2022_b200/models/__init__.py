"""Same public names as the reference models/__init__.py:1-3."""
from .mode_disparity import ModeDisparity
from .initModel import initModelPara, loadStackHourglassOnly
from .mode_fusion import Baseline, ModeFusion
from .sphere_conv import SphereConv, sphere_conv

__all__ = ['ModeDisparity', 'initModelPara', 'loadStackHourglassOnly', 'Baseline', 'ModeFusion', 'SphereConv', 'sphere_conv']
