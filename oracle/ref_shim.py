"""TEST INFRASTRUCTURE ONLY — import shim for the *real* reference (OATML/oatomobile).

Only usable where ``/root/reference`` exists (the build container, NOT the GPU
box).  It is used by ``tests/golden/make_golden*.py`` to generate the committed
golden vectors; ``tests/test_oracle_golden.py`` then pins the in-repo restatement
(``oracle/restatement.py``) against those vectors on every CPU run.

Nothing on the product path may import this module.

What the shim does:
  1. stubs ``gym`` / ``imageio`` / ``matplotlib`` (absent here, imported eagerly
     by ``oatomobile/__init__.py`` and ``oatomobile/core/rl.py:23-24``);
  2. puts ``/root/reference`` on ``sys.path``;
  3. replaces ``torch.hub.load`` (network) with the local torchvision
     ``mobilenet_v2`` constructor — the hub pin is
     ``oatomobile/torch/networks/perception.py:36-40``;
  4. ``fix_locscale(model)``: for T != 4 swaps ``_decoder._locscale`` for
     ``MLP(64, [32, 4])`` (``oatomobile/torch/networks/sequence.py:61`` sizes the
     head by T; the single sanctioned deviation);
  5. ``set_in_channels(model, C)``: rebuilds the stem conv exactly like
     ``perception.py:43-51`` does, for C != 2 (BASELINE.json uses C=4).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OAT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
  return os.path.isdir(os.path.join(REFERENCE_ROOT, "oatomobile"))


def _stub(name, **attrs):
  if name in sys.modules:
    return sys.modules[name]
  m = types.ModuleType(name)
  for k, v in attrs.items():
    setattr(m, k, v)
  sys.modules[name] = m
  return m


_installed = False


def install():
  """Makes ``import oatomobile`` work on CPU without CARLA/gym/network."""
  global _installed
  if _installed:
    return
  if not available():
    raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
  import torch
  import torchvision

  class _Env:  # gym.Env / gym.Wrapper stand-ins
    metadata = {}

    def __init__(self, *a, **k):
      pass

  class _Space:

    def __init__(self, *a, **k):
      pass

  spaces = _stub("gym.spaces", Box=_Space, Dict=_Space, Discrete=_Space,
                 Space=_Space, Tuple=_Space, MultiDiscrete=_Space)
  _stub("gym", Env=_Env, Wrapper=_Env, spaces=spaces, Space=_Space,
        ObservationWrapper=_Env, ActionWrapper=_Env, RewardWrapper=_Env)
  _stub("imageio")
  plt = _stub("matplotlib.pyplot")
  _stub("matplotlib", use=lambda *a, **k: None, pyplot=plt)
  class _AnyModule(types.ModuleType):  # attribute access yields dummy classes (type hints)

    def __getattr__(self, name):
      if name.startswith("__"):
        raise AttributeError(name)
      return type(name, (), {})

  for extra in ("tree", "wget", "pygame", "transforms3d", "transforms3d.euler", "skimage",
                "skimage.transform", "seaborn", "carla"):
    try:
      __import__(extra)
    except Exception:  # absent: harmless stub, never touched on the hot path
      if extra not in sys.modules:
        sys.modules[extra] = _AnyModule(extra)

  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)

  def _hub_load(github=None, model=None, *a, **k):
    assert model == "mobilenet_v2", model
    return torchvision.models.mobilenet_v2(*a, **k)

  torch.hub.load = _hub_load
  _installed = True


def install_carla_stubs():
  """For the reference's geometry helpers (oatomobile/utils/carla.py:642-700): a `carla`
  stub whose `Rotation(pitch, yaw, roll)` / `Location(x, y, z)` keep their arguments, and
  `transforms3d.euler.euler2mat` served by the restatement in oracle/euler.py (transforms3d
  0.3.1 is neither vendored in the reference nor installed here)."""
  install()
  from oracle import euler as _euler

  class _Rotation:

    def __init__(self, pitch=0.0, yaw=0.0, roll=0.0):
      self.pitch, self.yaw, self.roll = pitch, yaw, roll

  class _Location:

    def __init__(self, x=0.0, y=0.0, z=0.0):
      self.x, self.y, self.z = x, y, z

  carla = sys.modules["carla"]
  carla.Rotation, carla.Location = _Rotation, _Location
  t3d = sys.modules["transforms3d"]
  t3d.euler = types.ModuleType("transforms3d.euler")
  t3d.euler.euler2mat = _euler.euler2mat
  sys.modules["transforms3d.euler"] = t3d.euler
  import importlib
  cutil = importlib.import_module("oatomobile.utils.carla")
  cutil.carla, cutil.transforms3d = carla, t3d
  return cutil


def fix_locscale(model):
  """sequence.py:59-65 sizes the head by T; restore width 4 for T != 4."""
  from oatomobile.torch.networks.mlp import MLP
  import torch.nn as nn
  if model._output_shape[0] != 4:
    model._decoder._locscale = MLP(input_size=64, output_sizes=[32, 4],
                                   activation_fn=nn.ReLU, dropout_rate=None,
                                   activate_final=False)
  return model


def set_in_channels(model, in_channels):
  """Re-runs the stem swap of perception.py:43-51 for another channel count."""
  import torch.nn as nn
  if in_channels == 2:
    return model
  feats = model._encoder._model.features
  tmp = feats._modules['0']._modules['0']
  feats._modules['0']._modules['0'] = nn.Conv2d(
      in_channels=in_channels, out_channels=tmp.out_channels,
      kernel_size=tmp.kernel_size, stride=tmp.stride, padding=tmp.padding,
      bias=tmp.bias)
  return model


def make_imitative_model(T=4, in_channels=2, seed=0, randomize_bn=True):
  """Reference ``ImitativeModel`` with seeded weights."""
  install()
  import torch
  from oatomobile.baselines.torch.dim.model import ImitativeModel
  torch.manual_seed(seed)
  m = ImitativeModel(output_shape=(T, 2))
  fix_locscale(m)
  set_in_channels(m, in_channels)
  if randomize_bn:
    randomize_batchnorm(m, seed + 7919)
  return m.eval()


def make_behavioural_model(T=4, in_channels=2, seed=0, randomize_bn=True):
  install()
  import torch
  from oatomobile.baselines.torch.cil.model import BehaviouralModel
  torch.manual_seed(seed)
  m = BehaviouralModel(output_shape=(T, 2))
  set_in_channels(m, in_channels)
  if randomize_bn:
    randomize_batchnorm(m, seed + 7919)
  return m.eval()


def randomize_batchnorm(model, seed):
  """Non-trivial BN statistics/affine so BN-folding bugs show up in goldens."""
  import torch
  g = torch.Generator().manual_seed(seed)
  for mod in model.modules():
    if isinstance(mod, torch.nn.BatchNorm2d):
      n = mod.num_features
      with torch.no_grad():
        mod.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
        mod.running_var.copy_(torch.rand(n, generator=g) + 0.5)
        mod.weight.copy_(torch.rand(n, generator=g) + 0.5)
        mod.bias.copy_(torch.randn(n, generator=g) * 0.1)
  return model
