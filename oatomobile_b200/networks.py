"""Parameter containers with the reference's module tree and `state_dict` keys.

Mirrors oatomobile/torch/networks/{perception,mlp,sequence}.py: same class names,
constructor arguments and key layout (`_model.features.N.conv.i.j.weight`, ...), so
reference checkpoints load unchanged (`Checkpointer`, oatomobile/torch/savers.py).
The modules only *hold* weights; the arithmetic happens in the sm_100a kernels
(`oatomobile_b200/csrc`) reached through the C-ABI.  The leaf modules are stock
`torch.nn` layers so initialisation follows the reference (torchvision's
kaiming-normal/fan-out convs, N(0, 0.01) classifier, default GRUCell/Linear init).
"""
import math
from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributions as D
import torch.nn as nn

from oatomobile_b200 import _native as N
from oatomobile_b200 import ops

# (expand t, out c, repeats n, first stride s) — Sandler et al. 2018, table 2.
_MBV2_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
                 (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))


class _ConvBNReLU(nn.Sequential):

  def __init__(self, cin, cout, kernel_size=3, stride=1, groups=1):
    pad = (kernel_size - 1) // 2
    super().__init__(
        nn.Conv2d(cin, cout, kernel_size, stride, pad, groups=groups, bias=False),
        nn.BatchNorm2d(cout), nn.ReLU6(inplace=True))


class _InvertedResidual(nn.Module):

  def __init__(self, cin, cout, stride, expand_ratio):
    super().__init__()
    hidden = int(round(cin * expand_ratio))
    layers = []
    if expand_ratio != 1:
      layers.append(_ConvBNReLU(cin, hidden, kernel_size=1))
    layers.extend([
        _ConvBNReLU(hidden, hidden, stride=stride, groups=hidden),
        nn.Conv2d(hidden, cout, 1, 1, 0, bias=False),
        nn.BatchNorm2d(cout),
    ])
    self.conv = nn.Sequential(*layers)


class _MobileNetV2Body(nn.Module):
  """Key-compatible with torchvision `mobilenet_v2(num_classes=...)` (hub pin
  `pytorch/vision:v0.6.0`, oatomobile/torch/networks/perception.py:36-40)."""

  def __init__(self, num_classes):
    super().__init__()
    feats = [_ConvBNReLU(3, 32, stride=2)]
    cin = 32
    for t, c, n, s in _MBV2_SETTING:
      for i in range(n):
        feats.append(_InvertedResidual(cin, c, s if i == 0 else 1, t))
        cin = c
    feats.append(_ConvBNReLU(cin, 1280, kernel_size=1))
    self.features = nn.Sequential(*feats)
    self.classifier = nn.Sequential(nn.Dropout(0.2), nn.Linear(1280, num_classes))
    for m in self.modules():  # torchvision's initialisation
      if isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode="fan_out")
      elif isinstance(m, nn.BatchNorm2d):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)
      elif isinstance(m, nn.Linear):
        nn.init.normal_(m.weight, 0, 0.01)
        nn.init.zeros_(m.bias)


class MobileNetV2(nn.Module):
  """perception.py:25-55 — MobileNetV2 with a fresh `in_channels` stem conv."""

  def __init__(self, num_classes: int, in_channels: int = 3) -> None:
    super().__init__()
    self._model = _MobileNetV2Body(num_classes)
    # perception.py:43-51: the stem is replaced by a default-initialised Conv2d.
    self._model.features[0][0] = nn.Conv2d(in_channels, 32, kernel_size=3, stride=2, padding=1,
                                           bias=False)

    self._num_classes = num_classes
    object.__setattr__(self, "_cache", None)
    object.__setattr__(self, "_ens", None)

  def forward(self, x: torch.Tensor) -> torch.Tensor:
    """perception.py:53-55 — x [B,C,100,100] -> [B,num_classes], eval-mode arithmetic (folded
    BatchNorm running statistics, Dropout = identity) through `oat_encode_features`.  The CUDA
    encoder is specialised for the shapes of this path: 100x100 inputs and 128 classes."""
    if self._num_classes != 128:
      raise N.NativeLibraryError("the CUDA encoder is specialised for num_classes=128 "
                                 "(dim/model.py:53), got %d" % self._num_classes)
    if self._cache is None:
      object.__setattr__(self, "_cache", _HandleCache(self, N.KIND_ENCODER))
    h = self._cache.get()
    if self._ens is None or self._ens.models[0] is not h:
      object.__setattr__(self, "_ens", N.EnsembleHandle([h]))
    return ops.encode_features(self._ens, x)[0]


class MLP(nn.Module):
  """mlp.py:25-72 — Linear/activation stack with the reference's `_model.N` keys."""

  def __init__(self, input_size: int, output_sizes: Sequence[int],
               activation_fn: Callable[[], nn.Module] = nn.ReLU,
               dropout_rate: Optional[float] = None, activate_final: bool = False) -> None:
    super().__init__()
    layers = []
    sizes = [input_size] + list(output_sizes)
    for i in range(len(output_sizes) - 1):
      layers.append(nn.Linear(sizes[i], sizes[i + 1]))
      layers.append(activation_fn(inplace=True))
      if dropout_rate is not None:
        layers.append(nn.Dropout(p=dropout_rate, inplace=True))
    layers.append(nn.Linear(output_sizes[-2], output_sizes[-1]))
    if activate_final:
      layers.append(activation_fn(inplace=True))
    self._model = nn.Sequential(*layers)

    self._activate_final = activate_final
    if activation_fn is not nn.ReLU or dropout_rate is not None:
      self._unsupported = "activation_fn=%s, dropout_rate=%s" % (getattr(activation_fn, "__name__", activation_fn), dropout_rate)
    else:
      self._unsupported = None

  def forward(self, x: torch.Tensor) -> torch.Tensor:
    """mlp.py:70-72 — one fused launch (`oat_mlp_forward`) over the module's own parameters.
    Only the configuration this path uses has a kernel: ReLU activations, no dropout."""
    if self._unsupported:
      raise N.NativeLibraryError("MLP.forward: no CUDA kernel for %s (ReLU stacks without "
                                 "dropout only)" % self._unsupported)
    linears = [m for m in self._model if isinstance(m, nn.Linear)]
    return ops.mlp_forward([l.weight for l in linears], [l.bias for l in linears], x,
                           self._activate_final)


class _HandleCache:
  """Rebuilds the packed device weights when parameters or the device change."""

  def __init__(self, owner: nn.Module, kind: int):
    self._owner, self._kind = owner, kind
    self._key, self._handle = None, None

  def get(self) -> N.ModelHandle:
    tensors = list(self._owner.parameters()) + list(self._owner.buffers())
    dev = tensors[0].device
    key = (dev, tuple(t._version for t in tensors), tuple(t.data_ptr() for t in tensors))
    if key != self._key:
      if dev.type != "cuda":
        raise N.NativeLibraryError(
            "the model lives on %s: oatomobile_b200 runs on CUDA (sm_100a) only — call "
            "`.to('cuda')`; there is no CPU fallback." % dev)
      self._handle = N.ModelHandle(self._owner.state_dict(), self._kind, dev)
      self._key = key
    return self._handle

  def __deepcopy__(self, memo):  # device handles are never shared between copies
    return None

  def __reduce__(self):
    return (type(None), ())


class AutoregressiveFlow(nn.Module):
  """sequence.py:28-216 — GRU-conditioned affine autoregressive flow.

  Deviation (documented): the head is `MLP(hidden, [32, 4])`; the
  reference sizes it `[32, output_shape[0]]` (sequence.py:61), which equals 4 only
  for the default T=4 and crashes for every other T."""

  def __init__(self, output_shape: Tuple[int, int] = (4, 2), hidden_size: int = 64):
    super().__init__()
    if hidden_size != 64:
      raise ValueError("the fused flow kernel is specialised for hidden_size=64")
    if output_shape[-1] != 2:
      raise ValueError("the flow models 2-D waypoints (output_shape[-1] == 2)")
    self._output_shape = tuple(output_shape)
    d = self._output_shape[-2] * self._output_shape[-1]
    self._base_dist = D.MultivariateNormal(loc=torch.zeros(d), scale_tril=torch.eye(d))
    self._decoder = nn.GRUCell(input_size=2, hidden_size=hidden_size)
    self._locscale = MLP(input_size=hidden_size, output_sizes=[32, 4], activation_fn=nn.ReLU,
                         dropout_rate=None, activate_final=False)
    self._cache = None

  def to(self, *args, **kwargs):
    """sequence.py:67-74 — also moves the base distribution."""
    self = super().to(*args, **kwargs)
    self._base_dist = D.MultivariateNormal(
        loc=self._base_dist.mean.to(*args, **kwargs),
        scale_tril=self._base_dist.scale_tril.to(*args, **kwargs))
    return self

  def _handle(self) -> N.ModelHandle:
    if self._cache is None:
      object.__setattr__(self, "_cache", _HandleCache(self, N.KIND_FLOW))
    return self._cache.get()

  def forward(self, z: torch.Tensor) -> torch.Tensor:
    """sequence.py:76-93 — draw x ~ N(0, I) on the device and push it forward."""
    x = torch.randn(z.shape[0], *self._output_shape, device=z.device, dtype=torch.float32)
    return self._forward(x, z)[0]

  def _forward(self, x: torch.Tensor, z: torch.Tensor):
    """sequence.py:95-151 → (y [N,T,2], logabsdet [N])."""
    return ops.flow_forward(self._handle(), x, z)

  def _inverse(self, y: torch.Tensor, z: torch.Tensor):
    """sequence.py:153-216 → (x [N,T,2], log_prob [N], logabsdet [N])."""
    return ops.flow_inverse(self._handle(), y, z)
