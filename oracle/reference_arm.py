"""TEST / BASELINE INFRASTRUCTURE ONLY — the K-sample RIP sample-and-score loop (SURVEY.md
§3.5, the BASELINE.json metric) driven through the REAL reference's own modules.

The reference (OATML/oatomobile) is a Python program: when its tree is present — the
build container's `/root/reference`, or the copy installed under `baseline/_ref/` that
travels to the GPU box — this module imports it through `oracle/ref_shim.py` and runs

  ImitativeModel.transform                  oatomobile/baselines/torch/dim/model.py:221-253
  ImitativeModel._params                    dim/model.py:173-219
  AutoregressiveFlow._forward / _inverse    oatomobile/torch/networks/sequence.py:95-216
  per-sample goal log-likelihood            the distribution objects of dim/model.py:163-169
                                            (`_goal_likelihood` itself returns the batch mean)
  WCM / BCM / MA as written                 oatomobile/baselines/torch/rip/agent.py:121-127

on the CPU in fp32 (`kind: "reference"` in bench.py's `cpu_baseline` / `--impl reference`).
The sanctioned deviations are those of SURVEY §0: the `_locscale` head is `MLP(64,[32,4])`
for T != 4 and `_forward/_inverse` are reached through `model._decoder` (the reference's
RIPAgent calls them on the model and raises AttributeError).  Also used by
`tests/test_oracle_vs_reference.py` to compare the restatement with the reference live.
Nothing under `oatomobile_b200/` may import this module.
"""
import os
from typing import Dict, Optional, Sequence

from oracle import ref_shim

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.environ.get("OAT_REFERENCE_ROOT"), "/root/reference",
               os.path.join(os.path.dirname(_HERE), "baseline", "_ref"))


def locate() -> Optional[str]:
  """Directory that contains the `oatomobile` package of the reference, or None."""
  for root in _CANDIDATES:
    if root and os.path.isdir(os.path.join(root, "oatomobile", "baselines", "torch")):
      return root
  return None


def available() -> bool:
  return locate() is not None


def _install():
  root = locate()
  if root is None:
    raise RuntimeError("reference tree not found (looked in %s)" % (_CANDIDATES,))
  ref_shim.REFERENCE_ROOT = root
  ref_shim.install()
  return root


def build_models(state_dicts: Sequence[dict], T: int, in_channels: int):
  """Reference `ImitativeModel`s (eval mode) loaded with reference-format state_dicts."""
  _install()
  models = []
  for sd in state_dicts:
    m = ref_shim.make_imitative_model(T=T, in_channels=in_channels, seed=0, randomize_bn=False)
    m.load_state_dict(sd, strict=True)
    models.append(m.eval())
  return models


def rip_score(models, lidar, velocity, is_at_traffic_light, traffic_light_state, x, goal=None,
              epsilon: float = 1.0, algorithm: str = "WCM") -> Dict[str, "object"]:
  """One full step of the metric on the reference's modules: lidar [B,C,200,200],
  x [B,K,T,2] -> z, y, q [E,B,K], s, kstar, plan."""
  import torch
  import torch.distributions as D
  assert algorithm in ("WCM", "MA", "BCM")  # rip/agent.py:43
  B, K, T, _ = x.shape
  with torch.no_grad():
    obs = models[0].transform({"lidar": lidar.clone()})
    ctx = dict(visual_features=obs["visual_features"], velocity=velocity,
               is_at_traffic_light=is_at_traffic_light, traffic_light_state=traffic_light_state)
    zs = [m._params(**ctx) for m in models]                               # rip/agent.py:93
    rep = lambda z: z.repeat_interleave(K, dim=0)
    y, _ = models[0]._decoder._forward(x.reshape(B * K, T, 2), rep(zs[0]))  # rip/agent.py:106
    q = torch.empty(len(models), B, K, device=x.device)
    for m, (model, z) in enumerate(zip(models, zs)):                      # rip/agent.py:109-112
      _, log_prob, logabsdet = model._decoder._inverse(y, rep(z))
      q[m] = (log_prob - logabsdet).view(B, K)
    y = y.view(B, K, T, 2)
    if goal is not None:
      g = goal.repeat_interleave(K, dim=0)                                # [B*K,G,2]
      dist = D.MixtureSameFamily(                                         # dim/model.py:163-169
          mixture_distribution=D.Categorical(probs=torch.ones(g.shape[:2], device=g.device)),
          component_distribution=D.Independent(
              D.Normal(loc=g, scale=torch.ones_like(g) * epsilon), reinterpreted_batch_ndims=1))
      q = q + dist.log_prob(y[:, :, -1, :].reshape(B * K, 2)).view(1, B, K)
    if algorithm == "WCM":                                                # rip/agent.py:121-127
      s, _ = torch.min(-q, dim=0)
    elif algorithm == "BCM":
      s, _ = torch.max(-q, dim=0)
    else:
      s = torch.mean(-q, dim=0)
    kstar = torch.argmin(s, dim=1)
    rows = torch.arange(B, device=x.device)
    plan = y[rows, kstar]
  return dict(z=torch.stack(zs), y=y, q=q, s=s, kstar=kstar, plan=plan,
              sbest=s[rows, kstar])
