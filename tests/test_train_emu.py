"""Training step without a GPU.

1. Pins the oracle: the float64 restatement (`oracle.restatement.train_forward_backward`)
   against the golden vectors produced by the REAL reference's train_step bodies in
   float64 (tests/golden/make_golden_train.py).
2. Checks the arithmetic of every training kernel body: oatomobile_b200/csrc/train_*.h
   executed on the host (tests/emu) against that oracle, to 1e-4.

Gradients of a ReLU6 + BatchNorm network are discontinuous in the inputs: a pre-activation
within rounding distance of 0 or 6 takes a different linear piece in float32 than in
float64, which changes the gradient of a 4x4-resolution layer (64 rows per channel at B=4)
by percents and shifts everything upstream by ~0.5 %.  The float32 REFERENCE itself is
2e-2..7e-2 away from its own float64 run on these fixtures (`ref32_err` in the goldens).
The gradient check therefore evaluates the float64 oracle on the linear piece the
implementation took (the ReLU masks are read back from its activations) — everything else
about the oracle is unchanged — and holds the result to the 1e-4 bar; the unconstrained
comparison is kept with the float32 reference's own deviation as the yardstick."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from oatomobile_b200.synthetic import synthetic_state_dict
from tests.emu.driver import EmuTrainer
from tests.helpers import (TRAIN_CONFIGS, assert_close, dropout_mask, golden, grad_errors,
                           train_inputs)

def _reference64(cfg):
  sd = synthetic_state_dict(cfg["kind"], cfg["C"], cfg["wseed"])
  sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
  visual, scalars, target = train_inputs(cfg)
  mask = dropout_mask(cfg)
  out = R.train_forward_backward(sd64, cfg["kind"], visual.double(), scalars.double(), target.double(),
                                 None if mask is None else mask.double())
  return sd, (visual, scalars, target, mask), out


def yardstick(gold):
  """Worst deviation of the float32 reference from its float64 run, over tensors whose
  gradient does not vanish."""
  stats, err = gold["grad_stats"], gold["ref32_err"]
  return float(err[stats[:, 2] > 1e-6 * stats[:, 2].max()].max())


@pytest.mark.parametrize("name", sorted(TRAIN_CONFIGS))
def test_float64_restatement_reproduces_the_reference_train_step(name):
  cfg, gold = TRAIN_CONFIGS[name], golden(name)
  _, _, (loss, grads, buffers, aux) = _reference64(cfg)
  assert abs(loss.item() - float(gold["loss"])) < 1e-10 * abs(float(gold["loss"]))
  assert_close(aux, gold["aux"], tol=1e-9, what="z / predictions")
  names = [str(n) for n in gold["grad_names"]]
  assert sorted(names) == sorted(grads)
  top = gold["grad_stats"][:, 2].max()
  for i, k in enumerate(names):
    g = grads[k]
    got = np.array([g.sum().item(), g.norm().item(), g.abs().max().item()])
    scale = max(gold["grad_stats"][i, 2], 1e-6 * top)  # vanishing gradients: absolute
    assert np.all(np.abs(got - gold["grad_stats"][i]) <= 1e-7 * scale * max(1.0, g.numel()**0.5)), k
  for key in gold:
    if key.startswith("grad:"):
      assert_close(grads[key[5:]], gold[key], tol=1e-8, what=key)
    if key.startswith("buffer:"):
      assert_close(buffers[key[7:]], gold[key], tol=1e-10, what=key)


@pytest.mark.parametrize("name", sorted(TRAIN_CONFIGS))
def test_kernel_bodies_on_the_host_match_the_oracle(name):
  cfg, gold = TRAIN_CONFIGS[name], golden(name)
  sd, (visual, scalars, target, mask), (loss64, grads64, buffers64, aux64) = _reference64(cfg)
  emu = EmuTrainer(sd, cfg["kind"])
  loss, aux = emu.forward_backward(visual, scalars, target, mask)
  assert abs(loss.item() - float(gold["loss"])) < 2e-6 * abs(float(gold["loss"]))
  assert_close(aux, aux64, tol=1e-4, what="z / predictions")
  for k, v in buffers64.items():
    assert_close(emu.params[k], v, tol=1e-5, what=k)
  # (a) same linear piece: tight
  sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
  lossb, gradsb, _, _ = R.train_forward_backward(
      sd64, cfg["kind"], visual.double(), scalars.double(), target.double(),
      None if mask is None else mask.double(), branch=emu.branch(cfg["B"]))
  assert abs(lossb.item() - loss64.item()) < 1e-5 * abs(loss64.item())  # neighbouring pieces meet
  errs = grad_errors(emu.grads, gradsb)
  worst_key = max(errs, key=errs.get)
  assert errs[worst_key] <= 1e-4, "%s: %.2e" % (worst_key, errs[worst_key])
  # (b) unconstrained float64 truth: same order as the float32 reference's own deviation
  # (which flips happen is a lottery; one flip moves a 64-row channel by ~1/16)
  errs = grad_errors(emu.grads, grads64)
  worst_key = max(errs, key=errs.get)
  bar = max(0.25, 5.0 * yardstick(gold))
  assert errs[worst_key] <= bar, "%s: %.2e > %.2e" % (worst_key, errs[worst_key], bar)


@pytest.mark.parametrize("name", ["train_dim_T4_C2", "train_cil_T4_C2"])
def test_three_adam_steps_follow_the_reference_losses(name):
  cfg, gold = TRAIN_CONFIGS[name], golden(name)
  sd = synthetic_state_dict(cfg["kind"], cfg["C"], cfg["wseed"])
  visual, scalars, target = train_inputs(cfg)
  emu = EmuTrainer(sd, cfg["kind"])
  losses = []
  for _ in range(3):
    loss, _ = emu.forward_backward(visual, scalars, target)
    losses.append(loss.item())
    emu.adam(lr=1e-3)
  assert np.allclose(losses, gold["losses"], rtol=2e-3), (losses, gold["losses"])


def test_adam_body_matches_torch_optim_adam():
  g = torch.Generator().manual_seed(3)
  p0 = torch.randn(4097, generator=g)
  ref = torch.nn.Parameter(p0.clone())
  opt = torch.optim.Adam([ref], lr=3e-3, weight_decay=0.01)
  from tests.emu import driver
  p, m, v = p0.clone(), torch.zeros(4097), torch.zeros(4097)
  for step in range(1, 6):
    grad = torch.randn(4097, generator=g) * (10.0 if step == 2 else 0.1)
    ref.grad = grad.clone()
    torch.nn.utils.clip_grad_norm_([ref], 1.0)
    opt.step()
    driver.lib().emu_adam_step(driver._p(p), driver._p(grad), driver._p(m), driver._p(v), 4097, step,
                               3e-3, 0.9, 0.999, 1e-8, 0.01, 1.0, 1.0)
    assert_close(p, ref.detach(), tol=2e-6, what="step %d" % step)
