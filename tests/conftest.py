import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
  try:
    import torch
    has_gpu = torch.cuda.is_available()
  except Exception:
    has_gpu = False
  if has_gpu:
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
  """Achieved parity errors (max per quantity, see tests/helpers.assert_close) of a GPU run go to
  gpurun_out/parity_achieved.json, so the numbers behind the bars can be committed under profiles/."""
  try:
    from tests import helpers
    if not helpers.ACHIEVED:
      return
    import json
    import torch
    if not torch.cuda.is_available():
      return
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_achieved.json"), "w") as f:
      json.dump({k: helpers.ACHIEVED[k] for k in sorted(helpers.ACHIEVED)}, f, indent=1)
  except Exception:
    pass
