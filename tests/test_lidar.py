"""LIDAR → BEV histogram row: oracle vs the reference golden
(CPU), CUDA kernel vs oracle and golden, bit-exact (integer work)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import restatement as R
from tests.helpers import GOLDEN_DIR


def _points():
  spec = importlib.util.spec_from_file_location(
      "make_golden", os.path.join(GOLDEN_DIR, "make_golden.py"))
  # only the generator of the synthetic cloud is needed; avoid importing the reference
  src = open(spec.origin).read()
  start = src.index("def lidar_points")
  end = src.index("def make_lidar")
  ns = {"np": np}
  exec(src[start:end], ns)
  return ns["lidar_points"]()


def test_oracle_matches_reference_golden():
  g = np.load(os.path.join(GOLDEN_DIR, "lidar_bev.npz"))["levels"]
  bev = R.lidar_bev(_points())
  assert bev.shape == (200, 200, 2) and bev.dtype == np.float32
  assert np.array_equal(bev, g.astype(np.float32) / np.float32(5.0))
  assert g.max() == 5  # the clip is exercised


@pytest.mark.gpu
def test_cuda_lidar_bev_bit_exact():
  from oatomobile_b200 import ops
  g = np.load(os.path.join(GOLDEN_DIR, "lidar_bev.npz"))["levels"]
  pts = _points()
  bev = ops.lidar_bev(torch.from_numpy(pts).cuda()).cpu().numpy()
  assert np.array_equal(bev, g.astype(np.float32) / np.float32(5.0))
  rng = np.random.RandomState(7)
  for n in (0, 1, 100000):
    p = (rng.randn(n, 3) * np.array([30.0, 30.0, 3.0]) + np.array([0.0, 0.0, -2.5])).astype(np.float32)
    out = ops.lidar_bev(torch.from_numpy(p).cuda().reshape(-1, 3)).cpu().numpy()
    assert np.array_equal(out, R.lidar_bev(p)), n
