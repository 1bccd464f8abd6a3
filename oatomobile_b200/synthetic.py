"""Seeded synthetic inputs and weights.

Generated on the CPU with `torch.Generator` so the CPU oracle and the GPU path
see identical bits; used by the tests, `__graft_entry__.smoke()` and `bench.py`.
BEV values follow the reference LIDAR histogram levels k/5, ~90% empty
(oatomobile/utils/carla.py:199-204).
"""
import math
from collections import OrderedDict
from typing import Dict

import torch

_MBV2_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
                 (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))


def synthetic_inputs(B: int, C: int, K: int, T: int, G: int = 10,
                     seed: int = 0) -> Dict[str, torch.Tensor]:
  """lidar [B,C,200,200], velocity [B,3], is_at_traffic_light [B,1],
  traffic_light_state [B,1], goal [B,G,2], x [B,K,T,2] (base-distribution noise)."""
  g = torch.Generator().manual_seed(seed)
  occ = torch.rand(B, C, 200, 200, generator=g) < 0.1
  lvl = torch.randint(1, 6, (B, C, 200, 200), generator=g).float() / 5.0
  lidar = torch.where(occ, lvl, torch.zeros(()))
  velocity = torch.randn(B, 3, generator=g) * 5.0
  is_at_tl = torch.randint(0, 2, (B, 1), generator=g).float()
  tl_state = torch.randint(0, 4, (B, 1), generator=g).float()
  ahead = torch.arange(1, G + 1).float().view(1, G, 1) * torch.tensor([2.0, 0.0]).view(1, 1, 2)
  goal = ahead + torch.randn(B, G, 2, generator=g) * 0.5
  gx = torch.Generator().manual_seed(seed + 1)
  x = torch.randn(B, K, T, 2, generator=gx)
  return dict(lidar=lidar, velocity=velocity, is_at_traffic_light=is_at_tl,
              traffic_light_state=tl_state, goal=goal, x=x)


def synthetic_state_dict(kind: str = "dim", in_channels: int = 2,
                         seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
  """A full reference-format state_dict (328 entries for "dim", 326 for "cil") with
  seeded, well-conditioned random values: fan-in scaled weights so activations stay
  O(1) through the 53 layers, non-trivial BatchNorm statistics/affine so folding
  bugs show, GRU/head weights large enough for non-degenerate flow dynamics.
  Independent of any module construction order → reproducible on every box."""
  assert kind in ("dim", "cil")
  g = torch.Generator().manual_seed(seed)
  sd = OrderedDict()

  def randn(*shape, std=1.0):
    return torch.randn(*shape, generator=g) * std

  def conv(name, cout, cin_per_group, k, gain=2.0):
    sd[name + ".weight"] = randn(cout, cin_per_group, k, k,
                                 std=math.sqrt(gain / (cin_per_group * k * k)))

  def bn(name, n):
    sd[name + ".weight"] = torch.rand(n, generator=g) + 0.5
    sd[name + ".bias"] = randn(n, std=0.1)
    sd[name + ".running_mean"] = randn(n, std=0.1)
    sd[name + ".running_var"] = torch.rand(n, generator=g) + 0.5
    sd[name + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.long)

  def linear(name, nout, nin, gain=1.0, bias_std=0.1):
    sd[name + ".weight"] = randn(nout, nin, std=math.sqrt(gain / nin))
    sd[name + ".bias"] = randn(nout, std=bias_std)

  f = "_encoder._model.features."
  conv(f + "0.0", 32, in_channels, 3)
  bn(f + "0.1", 32)
  cin, idx = 32, 1
  for t, c, n, s in _MBV2_SETTING:
    for _ in range(n):
      hid = cin * t
      p = f + "%d.conv." % idx
      j = 0
      if t != 1:
        conv(p + "0.0", hid, cin, 1)
        bn(p + "0.1", hid)
        j = 1
      conv(p + "%d.0" % j, hid, 1, 3)
      bn(p + "%d.1" % j, hid)
      conv(p + "%d" % (j + 1), c, hid, 1, gain=1.0)
      bn(p + "%d" % (j + 2), c)
      cin, idx = c, idx + 1
  conv(f + "18.0", 1280, cin, 1)
  bn(f + "18.1", 1280)
  linear("_encoder._model.classifier.1", 128, 1280, gain=1.0)
  nscal = 5 if kind == "dim" else 6
  linear("_merger._model.0", 64, 128 + nscal, gain=0.5)
  linear("_merger._model.2", 64, 64, gain=2.0)
  linear("_merger._model.4", 64, 64, gain=1.0)
  gru = "_decoder._decoder." if kind == "dim" else "_decoder."
  sd[gru + "weight_ih"] = (torch.rand(192, 2, generator=g) - 0.5) * 0.6
  sd[gru + "weight_hh"] = (torch.rand(192, 64, generator=g) - 0.5) * 0.5
  sd[gru + "bias_ih"] = (torch.rand(192, generator=g) - 0.5) * 0.25
  sd[gru + "bias_hh"] = (torch.rand(192, generator=g) - 0.5) * 0.25
  if kind == "dim":
    linear("_decoder._locscale._model.0", 32, 64, gain=2.0)
    linear("_decoder._locscale._model.2", 4, 32, gain=1.0, bias_std=0.3)
  else:
    linear("_output", 2, 64, gain=1.0)
  return sd
