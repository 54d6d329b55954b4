from .geometry import cassini2Equirec, rotateCassini, depthViewTransWithConf, disp2depth, StageBoundary

__all__ = ['cassini2Equirec', 'rotateCassini', 'depthViewTransWithConf', 'disp2depth', 'StageBoundary']
