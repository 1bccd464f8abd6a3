"""TEST INFRASTRUCTURE ONLY — restatement of `transforms3d.euler.euler2mat`.

The reference's `rot2mat` (oatomobile/utils/carla.py:642-648) calls
`transforms3d.euler.euler2mat(roll, pitch, yaw)` from transforms3d==0.3.1 (pinned at
setup.py:57), a dependency that is NOT vendored under /root/reference and not installed
in this image.  Its published convention for the default `axes='sxyz'`: rotations about
the STATIC x, then y, then z axes, i.e. M = Rz(ak) @ Ry(aj) @ Rx(ai), column vectors.
This file restates that definition from elementary rotations (general static/rotating
axis strings), so `oracle/ref_shim.install_carla_stubs()` can run the reference's own
`local2world` / `world2local` on top of it to produce `tests/golden/geometry.npz`.
Known answers checked in tests/test_geometry.py (a 90 degree yaw maps x to y, etc.).
Nothing on the product path may import this module.
"""
import numpy as np

_AXIS = {"x": 0, "y": 1, "z": 2}


def _elementary(axis: int, angle: float) -> np.ndarray:
  c, s = np.cos(angle), np.sin(angle)
  m = np.eye(3)
  a, b = (axis + 1) % 3, (axis + 2) % 3
  m[a, a], m[a, b], m[b, a], m[b, b] = c, -s, s, c
  return m


def euler2mat(ai, aj, ak, axes="sxyz"):
  """Static ('s') axes: the three rotations are applied in order about fixed axes, so the
  matrices multiply right to left; rotating ('r') axes reverse the order of the angles."""
  frame, seq = axes[0], axes[1:]
  angles = (ai, aj, ak)
  if frame == "r":  # rotating frame abc == static frame cba with the angles reversed
    seq, angles = seq[::-1], angles[::-1]
  m = np.eye(3)
  for name, angle in zip(seq, angles):
    m = _elementary(_AXIS[name], angle) @ m
  return m
