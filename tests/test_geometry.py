"""Ego <-> world transforms: `oatomobile_b200.geometry` against goldens made by
the reference's own functions (tests/golden/make_golden_geometry.py), the transforms3d
restatement in oracle/euler.py, known answers and the round trip."""
import os

import numpy as np
import pytest

from oatomobile_b200 import geometry as G

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "geometry.npz"))


def test_rot2mat_matches_reference_golden():
  for r, want in zip(GOLD["rot"], GOLD["rot2mat"]):
    np.testing.assert_allclose(G.rot2mat(r), want, rtol=0, atol=1e-15)


def test_local2world_world2local_match_reference_golden():
  for i in range(len(GOLD["loc"])):
    w = G.local2world(current_location=GOLD["loc"][i], current_rotation=GOLD["rot"][i],
                      local_locations=GOLD["pts"][i])
    l = G.world2local(current_location=GOLD["loc"][i], current_rotation=GOLD["rot"][i],
                      world_locations=GOLD["pts"][i])
    np.testing.assert_allclose(w, GOLD["world"][i], rtol=0, atol=1e-12)
    np.testing.assert_allclose(l, GOLD["local"][i], rtol=0, atol=1e-12)
  # single points: local2world keeps the atleast_2d shape, world2local squeezes (as written)
  w1 = G.local2world(current_location=GOLD["loc"][2], current_rotation=GOLD["rot"][2],
                     local_locations=GOLD["pts"][2, 0])
  l1 = G.world2local(current_location=GOLD["loc"][2], current_rotation=GOLD["rot"][2],
                     world_locations=GOLD["pts"][2, 0])
  assert w1.shape == GOLD["world_single"].shape == (1, 3)
  assert l1.shape == GOLD["local_single"].shape == (3,)
  np.testing.assert_allclose(w1, GOLD["world_single"], atol=1e-12)
  np.testing.assert_allclose(l1, GOLD["local_single"], atol=1e-12)


def test_known_answers_and_round_trip():
  # CARLA rotation = [pitch, yaw, roll] degrees; a +90 degree yaw turns ego-x into world-y
  w = G.local2world(current_location=np.zeros(3), current_rotation=np.array([0.0, 90.0, 0.0]),
                    local_locations=np.array([[1.0, 0.0, 0.0]]))
  np.testing.assert_allclose(w, [[0.0, 1.0, 0.0]], atol=1e-15)
  np.testing.assert_allclose(G.rot2mat(np.zeros(3)), np.eye(3), atol=0)
  rng = np.random.default_rng(3)
  for _ in range(8):
    loc, rot = rng.normal(0, 50, 3), rng.uniform(-180, 180, 3)
    pts = rng.normal(0, 20, (30, 3))
    R = G.rot2mat(rot)
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-14)
    back = G.world2local(current_location=loc, current_rotation=rot,
                         world_locations=G.local2world(current_location=loc, current_rotation=rot,
                                                       local_locations=pts))
    np.testing.assert_allclose(back, pts, atol=1e-11)


def test_euler_restatement_conventions():
  """The oracle's transforms3d restatement: static xyz == Rz Ry Rx; rotating zyx is the same matrix."""
  from oracle.euler import euler2mat
  a, b, c = 0.3, -0.7, 1.9
  rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
  ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
  rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
  np.testing.assert_allclose(euler2mat(a, b, c), rz @ ry @ rx, atol=1e-15)
  np.testing.assert_allclose(euler2mat(c, b, a, "rzyx"), rz @ ry @ rx, atol=1e-15)


@pytest.mark.parametrize("T", [2, 4, 5, 8, 10, 20, 40])
def test_interpolate_plan_equals_scipy_interp1d(T):
  """rip/agent.py:141-151 interpolates with `scipy.interpolate.interp1d(x=time_index, y=plan,
  axis=0)`; the product and the oracle use np.interp per column — same piecewise-linear values."""
  import scipy.interpolate
  from oatomobile_b200.agents import interpolate_plan
  from oracle import restatement as R
  rng = np.random.RandomState(T)
  plan = rng.randn(T, 2).astype(np.float32) * 10
  time_index = list(range(0, 40, 40 // T))
  xy = scipy.interpolate.interp1d(x=time_index, y=plan, axis=0)(np.arange(0, time_index[-1]))
  want = np.c_[xy, np.zeros((xy.shape[0], 1))]
  for fn in (interpolate_plan, R.interpolate_plan):
    got = fn(plan)
    assert got.shape == want.shape and got.dtype == np.float64
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("T", [3, 6, 7, 9])
def test_interpolate_plan_rejects_what_the_reference_rejects(T):
  """For T that does not divide 40 the reference's time index has more entries than the plan and
  interp1d raises ValueError; so does the product."""
  import scipy.interpolate
  from oatomobile_b200.agents import interpolate_plan
  plan = np.zeros((T, 2), np.float32)
  with pytest.raises(ValueError):
    scipy.interpolate.interp1d(x=list(range(0, 40, 40 // T)), y=plan, axis=0)
  with pytest.raises(ValueError):
    interpolate_plan(plan)
