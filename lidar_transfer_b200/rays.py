"""Target-sensor beam pattern (host side, float64 -> float32 like the reference).

`create_rays` mirrors MultiSemLaserScan.create_rays (auxiliary/laserscan.py:1092-1119): same
name, arguments and result -- float32[H*W, 3], row-major, row 0 = fov_up, column 0 = yaw 180 deg,
endpoint-inclusive linspace (so column 0 and column W-1 are the same direction).  It is 131 072
trig evaluations per sensor and is computed once per target sensor, not per scan.
"""
import numpy as np


def create_rays(fov_up, fov_down, H, W):
  initial = 180.0  # "correct initial rotation of sensor", laserscan.py:1101
  yaw = np.linspace(0, 360, W) + initial
  yaw[yaw > 360] -= 360
  yaw = yaw / 180. * np.pi
  pitch = np.pi / 2 - np.linspace(fov_up, fov_down, H) / 180. * np.pi
  sp, cp = np.sin(pitch)[:, None], np.cos(pitch)[:, None]
  beams = np.empty((H, W, 3), np.float64)
  beams[..., 0] = sp * np.cos(-yaw)[None, :]
  beams[..., 1] = sp * np.sin(-yaw)[None, :]
  beams[..., 2] = cp * np.ones(W)[None, :]
  return np.ascontiguousarray(beams.reshape(H * W, 3).astype(np.float32))
