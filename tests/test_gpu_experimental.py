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
  """`Trainer(use_cuda_graphs=True)`: three optimiser steps must leave exactly the parameters,
  BatchNorm statistics and losses of the one-launch-per-kernel path (same RNG stream)."""
  from tests.test_gpu_train import batch_of, make_trainer
  cfg = TRAIN_CONFIGS[name]
  visual, scalars, target = train_inputs(cfg)
  results = []
  for graphs in (False, True):
    torch.manual_seed(1234)
    model, trainer, _ = make_trainer(cfg, use_cuda_graphs=graphs)
    batch = batch_of(cfg, visual, scalars)
    batch["player_future"] = torch.cat([target, torch.zeros_like(target[..., :1])], -1).cuda()
    losses = [trainer.train_step(batch).item() for _ in range(3)]
    results.append((losses, {k: v.clone() for k, v in model.state_dict().items()}))
  (l0, sd0), (l1, sd1) = results
  assert l0 == l1
  for k in sd0:
    assert torch.equal(sd0[k], sd1[k]), k
