"""Ego <-> world frame changes of the plan.

Mirrors `rot2mat`, `world2local`, `local2world` of oatomobile/utils/carla.py:642-700,
which `SetPointAgent.act` applies to the [N,3] plan returned by an agent's `__call__`
(oatomobile/baselines/base.py:128-135).  The reference builds the rotation with
`transforms3d.euler.euler2mat(roll, pitch, yaw).T` (transforms3d==0.3.1, static x-y-z
axes: R = Rz(yaw) Ry(pitch) Rx(roll)); this module writes that product out directly.
It is 3x3 host arithmetic on <= 40 points per tick, so it stays in NumPy float64 like
the reference.
"""
import numpy as np


def rot2mat(rotation: np.ndarray) -> np.ndarray:
  """utils/carla.py:642-648 — `rotation` = [pitch, yaw, roll] in degrees (the argument
  order of `carla.Rotation`, utils/carla.py:608-610) -> 3x3 world-to-ego matrix."""
  pitch, yaw, roll = (np.deg2rad(float(v)) for v in rotation)
  cr, sr = np.cos(roll), np.sin(roll)
  cp, sp = np.cos(pitch), np.sin(pitch)
  cy, sy = np.cos(yaw), np.sin(yaw)
  # Rz(yaw) @ Ry(pitch) @ Rx(roll), written out
  m = np.array([
      [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
      [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
      [-sp, cp * sr, cp * cr],
  ], dtype=np.float64)
  return m.T


def world2local(*, current_location: np.ndarray, current_rotation: np.ndarray,
                world_locations: np.ndarray) -> np.ndarray:
  """utils/carla.py:651-675 — [..., 3] world points -> ego frame (squeezed like the reference)."""
  current_location = np.asarray(current_location)
  current_rotation = np.asarray(current_rotation)
  world_locations = np.asarray(world_locations)
  assert current_location.shape == (3,)
  assert current_rotation.shape == (3,)
  assert len(world_locations.shape) < 3
  world_locations = np.atleast_2d(world_locations)
  R = rot2mat(current_rotation)
  return np.squeeze(np.dot(R, (world_locations - current_location).T).T)


def local2world(*, current_location: np.ndarray, current_rotation: np.ndarray,
                local_locations: np.ndarray) -> np.ndarray:
  """utils/carla.py:677-700 — [..., 3] ego-frame points -> world frame, always 2-D."""
  current_location = np.asarray(current_location)
  current_rotation = np.asarray(current_rotation)
  local_locations = np.asarray(local_locations)
  assert current_location.shape == (3,)
  assert current_rotation.shape == (3,)
  assert len(local_locations.shape) < 3
  local_locations = np.atleast_2d(local_locations)
  R_inv = np.linalg.inv(rot2mat(current_rotation))  # as written: inv(), not the transpose
  return np.dot(R_inv, local_locations.T).T + current_location
