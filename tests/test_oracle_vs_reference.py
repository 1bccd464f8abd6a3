"""Live comparison of the restatement (oracle/restatement.py) with the REAL reference, driven
through oracle/reference_arm.py on fresh seeds (not the committed goldens).  Runs wherever the
reference tree is present (/root/reference in the build container, baseline/_ref on a box that
received the snapshot); skipped elsewhere — tests/test_oracle_golden.py pins the oracle there."""
import pytest
import torch

from oracle import reference_arm as RA
from oracle import restatement as R
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
from tests.helpers import assert_close

pytestmark = pytest.mark.skipif(not RA.available(), reason="reference tree not present")


@pytest.mark.parametrize("T,C,E,B,K,algo,seed", [(4, 2, 2, 2, 16, "WCM", 11), (10, 4, 3, 2, 32, "MA", 12),
                                                 (10, 4, 2, 1, 64, "BCM", 13)])
def test_k_sample_scoring_restatement_equals_reference(T, C, E, B, K, algo, seed):
  inp = synthetic_inputs(B, C, K, T, seed=seed)
  sds = [synthetic_state_dict("dim", C, 900 + seed + m) for m in range(E)]
  models = RA.build_models(sds, T, C)
  args = (inp["lidar"], inp["velocity"], inp["is_at_traffic_light"], inp["traffic_light_state"],
          inp["x"], inp["goal"], 1.0, algo)
  ref = RA.rip_score(models, *args)
  with torch.no_grad():
    got = R.rip_score_from_inputs(sds, *args)
  for k in ("z", "y", "q", "s", "plan"):
    assert_close(got[k], ref[k].numpy(), 2e-5, k)
  # index: exact whenever the reference's own top-2 gap is above fp32 summation noise
  s_sorted, _ = torch.sort(ref["s"], dim=1)
  for b in range(B):
    if float(s_sorted[b, 1] - s_sorted[b, 0]) > 1e-4 * max(1.0, abs(float(s_sorted[b, 0]))):
      assert int(got["kstar"][b]) == int(ref["kstar"][b])
