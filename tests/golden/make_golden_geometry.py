"""Generates tests/golden/geometry.npz: outputs of the REAL reference's `local2world`,
`world2local` and `rot2mat` (oatomobile/utils/carla.py:642-700), run in the build container
through `oracle/ref_shim.install_carla_stubs()`.  transforms3d==0.3.1 (setup.py:57) is neither
vendored nor installed, so `euler2mat` underneath is the restatement in oracle/euler.py — the
reference code around it (argument order, the transpose, `np.linalg.inv`, `atleast_2d`,
`squeeze`) is the reference's own.  Run: `python tests/golden/make_golden_geometry.py`.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402


def main():
  cutil = ref_shim.install_carla_stubs()
  rng = np.random.default_rng(20260117)
  n = 12
  loc = rng.normal(0, 80, (n, 3))
  rot = np.stack([rng.uniform(-20, 20, n), rng.uniform(-180, 180, n), rng.uniform(-10, 10, n)], -1)
  rot[0] = 0.0                      # identity
  rot[1] = (0.0, 90.0, 0.0)         # pure yaw
  pts = rng.normal(0, 15, (n, 30, 3))
  pts[..., 2] = 0.0                 # plans carry a zero z column (rip/agent.py:149-151)
  out = dict(loc=loc, rot=rot, pts=pts)
  out["rot2mat"] = np.stack([cutil.rot2mat(r) for r in rot])
  out["world"] = np.stack([cutil.local2world(current_location=l, current_rotation=r, local_locations=p)
                           for l, r, p in zip(loc, rot, pts)])
  out["local"] = np.stack([cutil.world2local(current_location=l, current_rotation=r, world_locations=p)
                           for l, r, p in zip(loc, rot, pts)])
  out["world_single"] = cutil.local2world(current_location=loc[2], current_rotation=rot[2],
                                          local_locations=pts[2, 0])
  out["local_single"] = cutil.world2local(current_location=loc[2], current_rotation=rot[2],
                                          world_locations=pts[2, 0])
  np.savez_compressed(os.path.join(HERE, "geometry.npz"), **out)
  print("wrote geometry.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
  main()
