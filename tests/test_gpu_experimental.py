"""EXPERIMENTAL GPU paths that have not been validated on a B200 yet (written after the
round's GPU budget was spent).  They are opt-in in the product and these tests only run with
OAT_EXPERIMENTAL=1, so that the regular `-m gpu` run covers validated code only."""
import os

import pytest
import torch

from oatomobile_b200.synthetic import synthetic_state_dict
from tests.helpers import TRAIN_CONFIGS, train_inputs

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("OAT_EXPERIMENTAL") != "1",
                                 reason="experimental path: set OAT_EXPERIMENTAL=1")]


@pytest.mark.parametrize("name", sorted(TRAIN_CONFIGS))
def test_graphed_training_step_equals_plain_launches(name):
  """`Trainer(use_cuda_graphs=True)`: three optimiser steps must leave the parameters,
  BatchNorm statistics and losses of the one-launch-per-kernel path (same RNG stream).
  First B200 run (profiles/r1_train_graph_experiment.log): the graphed step works and is 10 %
  faster, but exact equality of the losses failed for CIL — the row reductions go through
  atomics, so the last bits are order dependent; the bar here is therefore 1e-5 / 1e-4 and a
  plain-vs-plain control run tells a real difference from that noise."""
  from tests.test_gpu_train import batch_of, make_trainer
  cfg = TRAIN_CONFIGS[name]
  visual, scalars, target = train_inputs(cfg)
  results = []
  for graphs in (False, False, True):
    torch.manual_seed(1234)
    model, trainer, _ = make_trainer(cfg, use_cuda_graphs=graphs)
    batch = batch_of(cfg, visual, scalars)
    batch["player_future"] = torch.cat([target, torch.zeros_like(target[..., :1])], -1).cuda()
    losses = [trainer.train_step(batch).item() for _ in range(3)]
    results.append((losses, {k: v.clone() for k, v in model.state_dict().items()}))
  (l0, sd0), (lc, sdc), (l1, sd1) = results

  def worst(a, b):
    return max(float((a[k].double() - b[k].double()).abs().max() /
                     max(1.0, float(a[k].double().abs().max())))
               for k in a if a[k].is_floating_point())

  noise = worst(sd0, sdc)  # run-to-run difference of the plain path itself
  print("plain-vs-plain %.3e  graphed-vs-plain %.3e  losses %s %s %s" % (noise, worst(sd0, sd1), l0, lc, l1))
  for a, b in zip(l0, l1):
    assert abs(a - b) <= 1e-5 * max(1.0, abs(a))
  assert worst(sd0, sd1) <= max(1e-4, 10 * noise)
