"""Generates tests/golden/train_{dim,cil}_T4_C2.npz by running the REAL reference's
`train_step` bodies (oatomobile/baselines/torch/dim/train.py:175-213,
cil/train.py:168-190) on the CPU through oracle/ref_shim.py, in `model.train()` mode.

Run in the build container only (`python tests/golden/make_golden_train.py`).  Weights
and inputs are regenerated from seeds by `oatomobile_b200.synthetic`.

The reference is run twice: in float64 (`model.double()`, the ground truth the tests
compare against) and in float32 as written.  Gradients of a ReLU6/BatchNorm network
over a handful of rows are badly conditioned — the float32 reference itself deviates
from its float64 run by up to ~1e-1 (relative to each tensor's largest entry) on these
inputs — so the float32 run's own deviation is stored next to the truth as the
yardstick (`ref32_err`).

Stored: loss (f64, f32), z / predictions, per-tensor gradient statistics (sum, L2 norm,
max |.|) for all 158 parameter tensors, a few complete gradient tensors, updated
BatchNorm running statistics of a few layers, and the losses of the two following
Adam steps (lr 1e-3) on the same batch.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402
from oracle import restatement as R  # noqa: E402
from oatomobile_b200.synthetic import synthetic_state_dict  # noqa: E402
from tests.helpers import TRAIN_CONFIGS, dropout_mask, train_inputs  # noqa: E402,F401

FULL_TENSORS = {
    "dim": ("_encoder._model.features.0.0.weight", "_encoder._model.features.17.conv.1.0.weight",
            "_encoder._model.features.9.conv.3.weight", "_encoder._model.classifier.1.bias",
            "_merger._model.0.weight", "_decoder._decoder.weight_ih", "_decoder._decoder.weight_hh",
            "_decoder._decoder.bias_hh", "_decoder._locscale._model.0.weight",
            "_decoder._locscale._model.2.weight", "_decoder._locscale._model.2.bias"),
    "cil": ("_encoder._model.features.0.0.weight", "_encoder._model.features.17.conv.1.0.weight",
            "_encoder._model.features.9.conv.3.weight", "_encoder._model.classifier.1.bias",
            "_merger._model.0.weight", "_decoder.weight_ih", "_decoder.weight_hh", "_decoder.bias_hh",
            "_output.weight", "_output.bias"),
}
BUFFERS = ("_encoder._model.features.0.1.running_mean", "_encoder._model.features.0.1.running_var",
           "_encoder._model.features.7.conv.1.1.running_var", "_encoder._model.features.18.1.running_mean",
           "_encoder._model.features.18.1.running_var")


def reference_model(cfg, dtype):
  make = ref_shim.make_imitative_model if cfg["kind"] == "dim" else ref_shim.make_behavioural_model
  model = make(T=cfg["T"], in_channels=cfg["C"], seed=0, randomize_bn=False)
  model.load_state_dict(synthetic_state_dict(cfg["kind"], cfg["C"], cfg["wseed"]), strict=True)
  model = model.to(dtype).train()
  if cfg["dropout_seed"] is None:
    model._encoder._model.classifier[0].p = 0.0
  return model


def reference_loss(model, cfg, visual, scalars, target):
  """The forward part of the reference `train_step` (no target noise: `target` is passed
  in already perturbed so that both precisions see the same numbers)."""
  batch = dict(visual_features=visual, velocity=scalars[:, 0:3], is_at_traffic_light=scalars[:, 3:4],
               traffic_light_state=scalars[:, 4:5])
  if cfg["dropout_seed"] is not None:
    torch.manual_seed(cfg["dropout_seed"])
  if cfg["kind"] == "dim":
    z = model._params(**batch)  # dim/train.py:192-197
    _, log_prob, logabsdet = model._decoder._inverse(y=target, z=z)
    return -torch.mean(log_prob - logabsdet, dim=0), z  # :200
  pred = model(mode=scalars[:, 5:6], **batch)  # cil/train.py:178
  loss = torch.nn.L1Loss(reduction="none")(pred, target)
  return torch.mean(torch.sum(loss, dim=[-2, -1]), dim=0), pred  # :180-182


def run(cfg, dtype, steps):
  visual, scalars, target = (t.to(dtype) for t in train_inputs(cfg))
  model = reference_model(cfg, dtype)
  opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.0)
  losses, first = [], None
  for s in range(steps):
    opt.zero_grad()
    loss, aux = reference_loss(model, cfg, visual, scalars, target)
    loss.backward()
    if s == 0:
      first = dict(loss=loss.detach().clone(), aux=aux.detach().clone(),
                   grads={k: p.grad.detach().clone() for k, p in model.named_parameters()},
                   buffers={k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k})
    losses.append(loss.item())
    opt.step()
  return first, losses


def make(name, cfg):
  f64, losses64 = run(cfg, torch.float64, steps=3)
  f32, _ = run(cfg, torch.float32, steps=1)
  names = list(f64["grads"].keys())
  stats = np.zeros((len(names), 3))
  err32 = np.zeros(len(names))
  for i, k in enumerate(names):
    g = f64["grads"][k]
    stats[i] = [g.sum().item(), g.norm().item(), g.abs().max().item()]
    err32[i] = (f32["grads"][k].double() - g).abs().max().item() / max(stats[i, 2], 1e-300)
  out = dict(loss=np.float64(f64["loss"].item()), loss32=np.float32(f32["loss"].item()),
             aux=f64["aux"].numpy(), losses=np.array(losses64), grad_stats=stats, ref32_err=err32,
             grad_names=np.array(names))
  for k in FULL_TENSORS[cfg["kind"]]:
    out["grad:" + k] = f64["grads"][k].numpy()
  for k in BUFFERS:
    out["buffer:" + k] = f64["buffers"][k].numpy()
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
  big = stats[:, 2] > 1e-6
  print(name, "loss %.9f (f32 %.9f) losses %s; float32 reference deviates by up to %.2e" %
        (out["loss"], out["loss32"], losses64, err32[big].max()))


if __name__ == "__main__":
  torch.set_num_threads(8)
  ref_shim.install()
  for name, cfg in TRAIN_CONFIGS.items():
    make(name, cfg)
