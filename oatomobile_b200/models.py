"""`ImitativeModel` and `BehaviouralModel` — drop-ins for
oatomobile/baselines/torch/dim/model.py:36-253 and cil/model.py:31-165.

Same constructor, methods, context keys, error behaviour and `state_dict` layout
as the reference; the arithmetic runs in the sm_100a kernels behind the C-ABI.
Inference semantics are `eval()` (BatchNorm running statistics, no dropout): the
reference agents never switch modes — here it is explicit.
"""
from typing import Mapping, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from oatomobile_b200 import _native as N
from oatomobile_b200 import ops
from oatomobile_b200 import transforms
from oatomobile_b200.networks import MLP, AutoregressiveFlow, MobileNetV2, _HandleCache

_CONTEXT_KEYS = ("visual_features", "velocity", "is_at_traffic_light", "traffic_light_state")


def _require(context: Mapping[str, torch.Tensor], keys: Sequence[str]) -> None:
  for key in keys:  # same message as dim/model.py:189-196 / cil/model.py:72-85
    if key not in context:
      raise ValueError("Missing `{}` keyword argument.".format(key))


def _scalars(context: Mapping[str, torch.Tensor], keys: Sequence[str]) -> torch.Tensor:
  # The reference concatenates these to the encoder output (dim/model.py:205-214);
  # here they are gathered into one [B,S] buffer the merger kernel reads.
  return torch.cat([context[k].reshape(context[k].shape[0], -1).float() for k in keys], dim=-1)


class _EncoderModel(nn.Module):
  """Shared plumbing: packed-weight cache, single-model ensemble, `transform`."""

  _KIND = N.KIND_DIM

  def _init_cache(self):
    object.__setattr__(self, "_cache", _HandleCache(self, self._KIND))
    object.__setattr__(self, "_ens", None)
    object.__setattr__(self, "_ens_of", None)

  def native_handle(self) -> N.ModelHandle:
    if getattr(self, "_cache", None) is None:
      self._init_cache()
    return self._cache.get()

  def _single_ensemble(self) -> N.EnsembleHandle:
    h = self.native_handle()
    if self._ens is None or self._ens_of is not h:
      object.__setattr__(self, "_ens", N.EnsembleHandle([h]))
      object.__setattr__(self, "_ens_of", h)
    return self._ens

  def transform(self, sample):
    """dim/model.py:221-253 / cil/model.py:129-165 — mutates and returns `sample`."""
    if "player_future" in sample:
      sample["player_future"] = transforms.downsample_target(
          player_future=sample["player_future"],
          num_timesteps_to_keep=self._output_shape[-2])
    if "lidar" in sample:
      sample["visual_features"] = sample.pop("lidar")
    if "visual_features" in sample:
      v = sample["visual_features"]
      if v.dim() == 4 and v.shape[-1] <= 8 < v.shape[1] and v.shape[1] == v.shape[2]:
        # extension: [B,H,W,C] as stored on disk / delivered by the simulator
        # (`DeviceCollator`) -> HWC->CHW fused into the resize kernel
        sample["visual_features"] = transforms.downsample_and_transpose_visual_features_hwc(v)
      else:
        sample["visual_features"] = transforms.downsample_and_transpose_visual_features(v)
    return sample


class ImitativeModel(_EncoderModel):
  """Deep imitative model: MobileNetV2 encoder → merger MLP → autoregressive flow."""

  _KIND = N.KIND_DIM

  def __init__(self, output_shape: Tuple[int, int] = (4, 2), in_channels: int = 2) -> None:
    """`in_channels` is an extension (the reference hard-codes 2, dim/model.py:53)."""
    super().__init__()
    self._output_shape = tuple(output_shape)
    self._encoder = MobileNetV2(num_classes=128, in_channels=in_channels)
    self._merger = MLP(input_size=128 + 3 + 1 + 1, output_sizes=[64, 64, 64],
                       activation_fn=nn.ReLU, dropout_rate=None, activate_final=True)
    self._decoder = AutoregressiveFlow(output_shape=self._output_shape, hidden_size=64)
    self._init_cache()

  def to(self, *args, **kwargs):
    """dim/model.py:70-74."""
    self = super().to(*args, **kwargs)
    self._decoder = self._decoder.to(*args, **kwargs)
    return self

  def _params(self, **context: torch.Tensor) -> torch.Tensor:
    """dim/model.py:173-219 → z [B,64] (stem → 17 MBConv blocks → pool → FC → merger)."""
    _require(context, _CONTEXT_KEYS)
    z = ops.encode(self._single_ensemble(), context["visual_features"],
                   _scalars(context, _CONTEXT_KEYS[1:]))
    return z[0]

  def _goal_likelihood(self, y: torch.Tensor, goal: torch.Tensor, **hyperparams) -> torch.Tensor:
    """dim/model.py:143-171 — batch-mean log-likelihood of y[:, -1] under the goal mixture."""
    epsilon = hyperparams.get("epsilon", 1.0)
    return ops.goal_likelihood(y, goal, epsilon)[1]

  def forward(self, num_steps: int, goal: Optional[torch.Tensor] = None, lr: float = 1e-1,
              epsilon: float = 1.0, x0: Optional[torch.Tensor] = None,
              **context: torch.Tensor) -> torch.Tensor:
    """dim/model.py:76-141 — Adam-on-latent MAP planner, one fused kernel launch.

    `x0` (extension, [1|B,T,2]) fixes the initial latent; by default it is one random
    base-distribution sample shared by the whole batch, as at dim/model.py:100-105."""
    if "visual_features" not in context:
      raise ValueError("Missing `visual_features` keyword argument.")
    batch_size = context["visual_features"].shape[0]
    z = self._params(**context)
    if x0 is None:
      x0 = torch.randn(1, *self._output_shape, device=z.device)
    x0 = x0.to(z.device, torch.float32).reshape(-1, *self._output_shape)
    if x0.shape[0] == 1:
      x0 = x0.repeat(batch_size, 1, 1)
    return ops.plan([self.native_handle()], z.unsqueeze(0), x0, num_steps=num_steps, lr=lr,
                    goal=goal, epsilon=epsilon, algorithm=None)[0]


class BehaviouralModel(_EncoderModel):
  """Conditional imitation learner: encoder → merger(+mode) → GRU residual roll-out."""

  _KIND = N.KIND_CIL

  def __init__(self, output_shape: Tuple[int, int] = (40, 2), in_channels: int = 2) -> None:
    super().__init__()
    self._output_shape = tuple(output_shape)
    self._encoder = MobileNetV2(num_classes=128, in_channels=in_channels)
    self._merger = MLP(input_size=128 + 3 + 1 + 1 + 1, output_sizes=[64, 64, 64],
                       activation_fn=nn.ReLU, dropout_rate=None, activate_final=True)
    self._decoder = nn.GRUCell(input_size=2, hidden_size=64)
    self._output = nn.Linear(in_features=64, out_features=self._output_shape[-1])
    self._init_cache()

  def forward(self, **context: torch.Tensor) -> torch.Tensor:
    """cil/model.py:68-127 → plan [B,T,2]."""
    keys = _CONTEXT_KEYS + ("mode",)
    _require(context, keys)
    z = ops.encode(self._single_ensemble(), context["visual_features"],
                   _scalars(context, keys[1:]))[0]
    return ops.cil_rollout(self.native_handle(), z, self._output_shape[0])

  def transform(self, sample):
    sample = super().transform(sample)
    if "mode" in sample:  # cil/model.py:161-163: drop the STOP command
      sample["mode"][sample["mode"] == 1.0] = 0.0
    return sample
