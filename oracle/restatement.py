"""TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the OATomobile RIP/DIM hot path.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``oatomobile_b200``) never
imports anything under ``oracle/`` and fails loudly without its CUDA library.

Parity status: the reference ships no tests or golden vectors, so
the restatement is pinned against *outputs of the reference itself*, run in the
build container through ``oracle/ref_shim.py``; the vectors are committed under
``tests/golden/`` by ``tests/golden/make_golden.py`` and re-checked on every
``pytest -m "not gpu"`` run (``tests/test_oracle_golden.py``).  Where the
reference tree is present ``tests/test_oracle_vs_reference.py`` additionally
compares restatement and reference live on fresh seeds.

Everything is float32 on the CPU with plain ``torch`` tensor ops (the reference
is itself a PyTorch-CPU program, so this is also the honest CPU baseline).  Each
function cites the reference lines it follows (paths relative to the reference
root).  Third-party arithmetic that is not vendored in the reference:
torchvision ``mobilenet_v2`` (hub pin ``pytorch/vision:v0.6.0``,
``oatomobile/torch/networks/perception.py:36-40``) — restated here from its
published architecture (Sandler et al. 2018, table 2: (t,c,n,s) =
(1,16,1,1),(6,24,2,2),(6,32,3,2),(6,64,4,2),(6,96,3,1),(6,160,3,2),(6,320,1,1),
conv-BN(eps 1e-5)-ReLU6, linear bottlenecks, residual iff stride 1 and cin==cout,
1x1 320->1280, global average pool, Linear) and driven by the reference's
``state_dict`` key layout; ``torch.nn.GRUCell`` / ``softplus`` /
``MultivariateNormal.log_prob`` / ``MixtureSameFamily.log_prob`` / ``Adam`` —
restated from their documented closed forms.
"""
import math
from typing import Dict, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Mapping[str, Tensor]

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used by torchvision ConvBNReLU
LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------
# a1. transforms (oatomobile/torch/transforms.py:23-49, dim/model.py:221-253)
# ----------------------------------------------------------------------------
def downsample_target(player_future: Tensor, num_timesteps_to_keep: int) -> Tensor:
  """transforms.py:23-31 — stride-slice the target sequence."""
  T = player_future.shape[1]
  inc = T // num_timesteps_to_keep
  return player_future[:, 0::inc, :]


def bilinear_resize_align_corners(x: Tensor, out_h: int, out_w: int) -> Tensor:
  """transforms.py:34-44 — F.interpolate(bilinear, align_corners=True), restated.

  src = dst * (in-1)/(out-1); 4-tap lerp with the top/left index floor(src) and
  the +1 neighbour clamped to the last row/column (ATen upsample_bilinear2d).
  """
  B, C, H, W = x.shape
  sh = (H - 1) / (out_h - 1) if out_h > 1 else 0.0
  sw = (W - 1) / (out_w - 1) if out_w > 1 else 0.0
  # ATen computes the source coordinate in float32: scale (fp32) * dst index.
  sy = torch.arange(out_h, dtype=torch.float32) * torch.tensor(sh, dtype=torch.float32)
  sx = torch.arange(out_w, dtype=torch.float32) * torch.tensor(sw, dtype=torch.float32)
  y0 = sy.floor().long().clamp(max=H - 1)
  x0 = sx.floor().long().clamp(max=W - 1)
  y1 = (y0 + 1).clamp(max=H - 1)
  x1 = (x0 + 1).clamp(max=W - 1)
  ly = (sy - y0.float()).view(1, 1, out_h, 1)
  lx = (sx - x0.float()).view(1, 1, 1, out_w)
  hy, hx = 1.0 - ly, 1.0 - lx
  r0 = x[:, :, y0, :]
  r1 = x[:, :, y1, :]
  top = hx * r0[:, :, :, x0] + lx * r0[:, :, :, x1]
  bot = hx * r1[:, :, :, x0] + lx * r1[:, :, :, x1]
  return hy * top + ly * bot


def transform_visual(lidar: Tensor) -> Tensor:
  """dim/model.py:245-251 — resize to 100x100 then swap H and W."""
  return bilinear_resize_align_corners(lidar, 100, 100).transpose(2, 3)


# ----------------------------------------------------------------------------
# a2. MobileNetV2 encoder (perception.py:25-55 + torchvision mobilenet_v2, eval)
# ----------------------------------------------------------------------------
def _bn(x: Tensor, sd: StateDict, p: str, train: Optional[dict] = None) -> Tensor:
  """BatchNorm2d.  Eval mode (train is None): (x-mean)/sqrt(var+eps)*gamma+beta with the
  running statistics.  Training mode (`model.train()`, dim/train.py:223): batch mean and
  biased batch variance normalise; the updated running estimates (momentum 0.1, unbiased
  variance) are recorded in `train["buffers"]`."""
  g = sd[p + ".weight"].view(1, -1, 1, 1)
  b = sd[p + ".bias"].view(1, -1, 1, 1)
  if train is None:
    mean = sd[p + ".running_mean"].view(1, -1, 1, 1)
    var = sd[p + ".running_var"].view(1, -1, 1, 1)
    return (x - mean) / torch.sqrt(var + BN_EPS) * g + b
  train["unit"] = train.get("unit", 0) + 1  # conv+BN units in execution order, 1-based
  n = x.numel() // x.shape[1]
  mean = x.mean(dim=(0, 2, 3))
  var = ((x - mean.view(1, -1, 1, 1))**2).mean(dim=(0, 2, 3))
  with torch.no_grad():
    train["buffers"][p + ".running_mean"] = 0.9 * sd[p + ".running_mean"] + 0.1 * mean
    train["buffers"][p + ".running_var"] = 0.9 * sd[p + ".running_var"] + 0.1 * var * (n / max(n - 1, 1))
  return (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS) * g + b


def _relu6(x: Tensor, train: Optional[dict] = None) -> Tensor:
  """ReLU6.  With `train["branch"]` (gradient checks only) the piecewise-linear branch is
  not chosen by x but prescribed: branch[unit] = (pass-through mask, saturated-at-6 mask),
  so that the float64 oracle differentiates the same linear piece a float32 implementation
  took when a pre-activation sits within rounding distance of 0 or 6."""
  if train is None or train.get("branch") is None:
    return x.clamp(min=0.0, max=6.0)
  passthrough, saturated = train["branch"][train["unit"] - 1]
  return x * passthrough.to(x.dtype) + 6.0 * saturated.to(x.dtype)


def _conv_bn_relu6(x, sd, p, stride, groups, train=None):
  w = sd[p + ".0.weight"]
  pad = (w.shape[-1] - 1) // 2
  x = F.conv2d(x, w, None, stride=stride, padding=pad, groups=groups)
  return _relu6(_bn(x, sd, p + ".1", train), train)


# (expand t, out c, repeats n, first stride s) — Sandler et al. 2018, table 2.
MBV2_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
                (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))


def mbv2_block_table():
  """[(features index, cin, hidden, cout, stride, residual)] for blocks 1..17."""
  table, cin, idx = [], 32, 1
  for t, c, n, s in MBV2_SETTING:
    for i in range(n):
      stride = s if i == 0 else 1
      table.append((idx, cin, cin * t, c, stride, stride == 1 and cin == c))
      cin, idx = c, idx + 1
  return table


def mobilenet_v2_encode(sd: StateDict, x: Tensor, prefix: str = "_encoder._model.",
                        train: Optional[dict] = None) -> Tensor:
  """perception.py:53-55 → torchvision MobileNetV2.forward.  Eval mode (no dropout) by
  default; with `train` (a dict holding "buffers" and optionally "dropout_mask" [B,1280],
  already scaled by 1/(1-p)) the BatchNorms use batch statistics and the classifier's
  Dropout(0.2) multiplies the pooled features by the given mask.

  x: [B,C,100,100] → [B,128].
  """
  f = prefix + "features."
  x = _conv_bn_relu6(x, sd, f + "0", stride=2, groups=1, train=train)
  for idx, cin, hid, cout, stride, res in mbv2_block_table():
    p = f + "%d.conv" % idx
    h = x
    if hid != cin:  # expand 1x1
      h = _conv_bn_relu6(h, sd, p + ".0", stride=1, groups=1, train=train)
      dw, pj, pjbn = p + ".1", p + ".2", p + ".3"
    else:  # t == 1: no expand conv
      dw, pj, pjbn = p + ".0", p + ".1", p + ".2"
    h = _conv_bn_relu6(h, sd, dw, stride=stride, groups=hid, train=train)
    h = _bn(F.conv2d(h, sd[pj + ".weight"], None), sd, pjbn, train)  # linear bottleneck
    x = x + h if res else h
  x = _conv_bn_relu6(x, sd, f + "18", stride=1, groups=1, train=train)
  x = x.mean(dim=(2, 3))  # adaptive_avg_pool2d(1) + flatten
  if train is not None and train.get("dropout_mask") is not None:
    x = x * train["dropout_mask"]
  return F.linear(x, sd[prefix + "classifier.1.weight"], sd[prefix + "classifier.1.bias"])


# ----------------------------------------------------------------------------
# a3/a4. merger MLP and _params (mlp.py:49-68, dim/model.py:173-219)
# ----------------------------------------------------------------------------
def mlp3_relu(sd: StateDict, u: Tensor, prefix: str = "_merger._model.",
              train: Optional[dict] = None) -> Tensor:
  """MLP(.., [64,64,64], activate_final=True): Linear-ReLU x3 (mlp.py:49-66).  A
  prescribed branch (see `_relu6`) is read from train["branch"][52 + layer]."""
  for j, i in enumerate((0, 2, 4)):
    u = F.linear(u, sd[prefix + "%d.weight" % i], sd[prefix + "%d.bias" % i])
    if train is not None and train.get("branch") is not None:
      u = u * train["branch"][52 + j][0].to(u.dtype)
    else:
      u = F.relu(u)
  return u


def imitative_params(sd: StateDict, visual_features: Tensor, velocity: Tensor,
                     is_at_traffic_light: Tensor, traffic_light_state: Tensor,
                     train: Optional[dict] = None) -> Tensor:
  """dim/model.py:203-217 — z = merger(cat[encoder(v), vel, tl, tls]); [B,64]."""
  e = mobilenet_v2_encode(sd, visual_features, train=train)
  return mlp3_relu(sd, torch.cat([e, velocity, is_at_traffic_light, traffic_light_state], -1),
                   train=train)


# ----------------------------------------------------------------------------
# a5/a6. autoregressive flow (sequence.py:95-216)
# ----------------------------------------------------------------------------
def gru_cell(u: Tensor, h: Tensor, w_ih, w_hh, b_ih, b_hh) -> Tensor:
  """torch.nn.GRUCell, gate order (r, z, n) — used at sequence.py:128,188."""
  gi = F.linear(u, w_ih, b_ih)
  gh = F.linear(h, w_hh, b_hh)
  H = h.shape[-1]
  r = torch.sigmoid(gi[:, :H] + gh[:, :H])
  g = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
  n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
  return (1.0 - g) * n + g * h


def _flow_weights(sd: StateDict, prefix: str = "_decoder."):
  return dict(
      w_ih=sd[prefix + "_decoder.weight_ih"], w_hh=sd[prefix + "_decoder.weight_hh"],
      b_ih=sd[prefix + "_decoder.bias_ih"], b_hh=sd[prefix + "_decoder.bias_hh"],
      w1=sd[prefix + "_locscale._model.0.weight"], b1=sd[prefix + "_locscale._model.0.bias"],
      w2=sd[prefix + "_locscale._model.2.weight"], b2=sd[prefix + "_locscale._model.2.bias"])


def _locscale(h: Tensor, w) -> Tuple[Tensor, Tensor]:
  """sequence.py:131-133 — head MLP(64→32→4); scale = softplus(.)+1e-3."""
  o = F.linear(F.relu(F.linear(h, w["w1"], w["b1"])), w["w2"], w["b2"])
  return o[:, :2], F.softplus(o[:, 2:]) + 1e-3


def flow_forward(sd: StateDict, x: Tensor, z: Tensor) -> Tuple[Tensor, Tensor]:
  """sequence.py:95-151 — x [N,T,2], z [N,64] → y [N,T,2], logabsdet [N].

  logabsdet = sum_xy log|prod_t sigma| (product over T first, sequence.py:148-149).
  """
  w = _flow_weights(sd)
  h, y_prev = z, torch.zeros(z.shape[0], 2, dtype=z.dtype)
  ys, scales = [], []
  for t in range(x.shape[1]):
    h = gru_cell(y_prev, h, w["w_ih"], w["w_hh"], w["b_ih"], w["b_hh"])
    dloc, scale = _locscale(h, w)
    y_t = (y_prev + dloc) + scale * x[:, t, :]
    ys.append(y_t)
    scales.append(scale)
    y_prev = y_t
  y = torch.stack(ys, dim=1)
  s = torch.stack(scales, dim=1)
  return y, torch.log(torch.abs(torch.prod(s, dim=1))).sum(-1)


def flow_inverse(sd: StateDict, y: Tensor, z: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
  """sequence.py:153-216 — y [N,T,2], z [N,64] → x, log_prob [N], logabsdet [N].

  log_prob = MVN(0,I_2T).log_prob(x) = -0.5*||x||^2 - T*log(2*pi)  (:208);
  logabsdet = sum_t log|sigma_x*sigma_y|  (:211-214).
  """
  w = _flow_weights(sd)
  h, y_prev = z, torch.zeros(z.shape[0], 2, dtype=z.dtype)
  xs, scales = [], []
  for t in range(y.shape[1]):
    y_t = y[:, t, :]
    h = gru_cell(y_prev, h, w["w_ih"], w["w_hh"], w["b_ih"], w["b_hh"])
    dloc, scale = _locscale(h, w)
    xs.append((y_t - (y_prev + dloc)) / scale)
    scales.append(scale)
    y_prev = y_t
  x = torch.stack(xs, dim=1)
  s = torch.stack(scales, dim=1)
  T = x.shape[1]
  log_prob = -0.5 * (x.reshape(x.shape[0], -1)**2).sum(-1) - T * LOG_2PI
  logabsdet = torch.log(torch.abs(torch.prod(s, dim=-1))).sum(-1)
  return x, log_prob, logabsdet


# ----------------------------------------------------------------------------
# a8. goal likelihood (dim/model.py:143-171)
# ----------------------------------------------------------------------------
def goal_log_likelihood_rows(y_last: Tensor, goal: Tensor, epsilon: float = 1.0) -> Tensor:
  """Per-row mixture log-likelihood; dim/model.py:163-171 without the batch mean.

  y_last [..., 2], goal [..., G, 2] (broadcastable) →
  logsumexp_g(-||y-g||^2/(2 eps^2) - log(2 pi eps^2)) - log G.
  """
  G = goal.shape[-2]
  d = y_last.unsqueeze(-2) - goal
  # Independent(Normal).log_prob sums the two per-coordinate log-densities.
  comp = (-(d**2) / (2.0 * epsilon**2) - math.log(epsilon) - 0.5 * LOG_2PI).sum(-1)
  return torch.logsumexp(comp - math.log(G), dim=-1)


def goal_log_likelihood(y: Tensor, goal: Tensor, epsilon: float = 1.0) -> Tensor:
  """dim/model.py:143-171 — mean over the batch of the mixture log-prob of y[:,-1]."""
  return goal_log_likelihood_rows(y[:, -1, :], goal, epsilon).mean(0)


# ----------------------------------------------------------------------------
# a11 + §3.5. K-sample RIP sample-and-score (BASELINE.json metric)
# ----------------------------------------------------------------------------
def rip_aggregate(q: Tensor, algorithm: str) -> Tensor:
  """rip/agent.py:121-127 on per-sample scores q [E,B,K] → loss s [B,K].

  As written in the reference: "WCM" = min_m(-q), "BCM" = max_m(-q), else mean.
  """
  assert algorithm in ("WCM", "MA", "BCM")
  if algorithm == "WCM":
    return torch.min(-q, dim=0)[0]
  if algorithm == "BCM":
    return torch.max(-q, dim=0)[0]
  s = -q[0]
  for m in range(1, q.shape[0]):  # fixed order 0..E-1 (matters for bit-exact index)
    s = s + (-q[m])
  return s / q.shape[0]


def rip_sample_and_score(sds: Sequence[StateDict], zs: Sequence[Tensor], x: Tensor,
                         goal: Optional[Tensor] = None, epsilon: float = 1.0,
                         algorithm: str = "WCM") -> Dict[str, Tensor]:
  """SURVEY.md §3.5, assembled from rip/agent.py:106-127,137.

  zs[m] [B,64]; x [B,K,T,2].  Proposals y = f_0(x; z_0) (rip/agent.py:106);
  q[m,b,k] = log_prob_m - logabsdet_m (+ per-sample goal log-likelihood);
  s = aggregate; k* = argmin_k s (lowest index on ties); plan = y[b,k*].
  """
  B, K, T, _ = x.shape
  E = len(sds)
  rep = lambda z: z.unsqueeze(1).expand(B, K, z.shape[-1]).reshape(B * K, -1)
  y, _ = flow_forward(sds[0], x.reshape(B * K, T, 2), rep(zs[0]))
  q = torch.empty(E, B, K)
  for m in range(E):
    _, lp, lad = flow_inverse(sds[m], y, rep(zs[m]))
    q[m] = (lp - lad).view(B, K)
  y = y.view(B, K, T, 2)
  if goal is not None:
    q = q + goal_log_likelihood_rows(y[:, :, -1, :], goal.unsqueeze(1), epsilon).unsqueeze(0)
  s = rip_aggregate(q, algorithm)
  kstar = torch.argmin(s, dim=1)
  plan = y[torch.arange(B), kstar]
  return dict(y=y, q=q, s=s, kstar=kstar, plan=plan,
              sbest=s[torch.arange(B), kstar])


def rip_score_from_inputs(sds, lidar, velocity, is_at_traffic_light, traffic_light_state,
                          x, goal=None, epsilon=1.0, algorithm="WCM"):
  """Full path of the metric: transform → E×_params → sample-and-score."""
  vis = transform_visual(lidar)
  zs = [imitative_params(sd, vis, velocity, is_at_traffic_light, traffic_light_state)
        for sd in sds]
  out = rip_sample_and_score(sds, zs, x, goal, epsilon, algorithm)
  out["z"] = torch.stack(zs, 0)
  return out


# ----------------------------------------------------------------------------
# a9/a10. gradient-based planners as written (dim/model.py:97-141, rip/agent.py:84-137)
# ----------------------------------------------------------------------------
def _adam_step(x, grad, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
  """torch.optim.Adam (no weight decay, no amsgrad), one parameter tensor."""
  m.mul_(b1).add_(grad, alpha=1 - b1)
  v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
  bc1 = 1 - b1**step
  bc2 = 1 - b2**step
  denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
  return x - (lr / bc1) * (m / denom)


def planner(sds: Sequence[StateDict], zs: Sequence[Tensor], x0: Tensor, num_steps: int,
            lr: float, goal: Optional[Tensor], epsilon: float,
            algorithm: Optional[str], trace: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
  """Shared restatement of the two Adam-on-latent planners.

  dim/model.py:117-139 (single model, ``algorithm=None``) and rip/agent.py:102-137
  (ensemble).  Quirk kept: ``x_best`` is the *post-step* x (dim/model.py:133-137).
  Returns (plan y [B,T,2], x_best).  `trace` (a dict) receives the per-step losses and the
  first-step gradient, for tests that pin per-step quantities.
  """
  x = x0.clone()
  m = torch.zeros_like(x)
  v = torch.zeros_like(x)
  x_best = x.clone()
  loss_best = torch.tensor(1000.0)
  for step in range(1, num_steps + 1):
    xg = x.clone().requires_grad_(True)
    y, _ = flow_forward(sds[0], xg, zs[0])
    posts = []
    for sd, z in zip(sds, zs):
      _, lp, lad = flow_inverse(sd, y, z)
      post = torch.mean(lp - lad)
      if goal is not None:
        post = post + goal_log_likelihood(y, goal, epsilon)
      posts.append(post)
    P = torch.stack(posts, 0)
    if algorithm is None:
      loss = -P[0]
    elif algorithm == "WCM":
      loss = torch.min(-P, dim=0)[0]
    elif algorithm == "BCM":
      loss = torch.max(-P, dim=0)[0]
    else:
      loss = torch.mean(-P, dim=0)
    (grad,) = torch.autograd.grad(loss, xg)
    if trace is not None:
      trace.setdefault("losses", []).append(float(loss.detach()))
      if step == 1:
        trace["grad1"] = grad.detach().clone()
    x = _adam_step(x, grad, m, v, step, lr)
    if loss.detach() < loss_best:
      x_best = x.clone()
      loss_best = loss.detach().clone()
  with torch.no_grad():
    y, _ = flow_forward(sds[0], x_best, zs[0])
  return y, x_best


def imitative_forward(sd: StateDict, x0: Tensor, num_steps: int, goal=None, lr=1e-1,
                      epsilon=1.0, **context) -> Tensor:
  """dim/model.py:76-141 with the initial base sample ``x0`` passed in (:100-105)."""
  with torch.no_grad():
    z = imitative_params(sd, context["visual_features"], context["velocity"],
                         context["is_at_traffic_light"], context["traffic_light_state"])
  return planner([sd], [z], x0, num_steps, lr, goal, epsilon, None)[0]


def rip_plan(sds, zs, T, goal, num_steps=10, lr=1e-1, epsilon=1.0, algorithm="WCM"):
  """rip/agent.py:84-137 — x starts at the base mean (zeros [B,T,2]), E models."""
  x0 = torch.zeros(zs[0].shape[0], T, 2)
  return planner(sds, zs, x0, num_steps, lr, goal, epsilon, algorithm)[0]


# ----------------------------------------------------------------------------
# a13. BehaviouralModel.forward (cil/model.py:68-127)
# ----------------------------------------------------------------------------
def behavioural_forward(sd: StateDict, T: int, visual_features, velocity,
                        is_at_traffic_light, traffic_light_state, mode, train=None) -> Tensor:
  """cil/model.py:88-127 — encoder → merger(134) → T x {GRUCell; x += Linear(h)}."""
  e = mobilenet_v2_encode(sd, visual_features, train=train)
  h = mlp3_relu(sd, torch.cat([e, velocity, is_at_traffic_light, traffic_light_state, mode], -1),
                train=train)
  x = torch.zeros(h.shape[0], 2, dtype=h.dtype)
  ys = []
  for _ in range(T):
    h = gru_cell(x, h, sd["_decoder.weight_ih"], sd["_decoder.weight_hh"],
                 sd["_decoder.bias_ih"], sd["_decoder.bias_hh"])
    x = F.linear(h, sd["_output.weight"], sd["_output.bias"]) + x
    ys.append(x)
  return torch.stack(ys, dim=1)


# ----------------------------------------------------------------------------
# a14. training steps (dim/train.py:175-213, cil/train.py:168-190)
# ----------------------------------------------------------------------------
def _leaf_state_dict(sd: StateDict):
  """Floating-point entries as autograd leaves (buffers stay plain tensors)."""
  out = {}
  for k, v in sd.items():
    if not v.is_floating_point():
      out[k] = v
    elif "running_" in k:
      out[k] = v.detach().clone()
    else:
      out[k] = v.detach().clone().requires_grad_(True)
  return out


def train_forward_backward(sd: StateDict, kind: str, visual_features, scalars: Tensor,
                           target: Tensor, dropout_mask: Optional[Tensor] = None,
                           branch: Optional[dict] = None):
  """`train_step` up to (not including) `optimizer.step()`, in `model.train()` mode.

  kind "dim" (dim/train.py:190-203): z = _params(batch); loss = -mean(log_prob - logabsdet)
  of `target` (the perturbed player_future[..., :2]) under `_decoder._inverse`.
  kind "cil" (cil/train.py:176-184): loss = mean_b sum_{t,d} |forward(batch) - target|.
  scalars = [velocity(3) | is_at_traffic_light | traffic_light_state (| mode)].
  `branch` (gradient checks): {unit: (pass-through mask, saturated mask)} prescribing the
  ReLU6 / ReLU linear piece per activation — units 0..51 NCHW masks, 52..54 the merger.
  Returns (loss, grads {name: tensor}, new BatchNorm buffers {name: tensor}, z-or-pred)."""
  leaves = _leaf_state_dict(sd)
  train = {"buffers": {}, "dropout_mask": dropout_mask, "branch": branch}
  vel, tl, tls = scalars[:, 0:3], scalars[:, 3:4], scalars[:, 4:5]
  if kind == "dim":
    z = imitative_params(leaves, visual_features, vel, tl, tls, train=train)
    _, log_prob, logabsdet = flow_inverse(leaves, target, z)
    loss = -torch.mean(log_prob - logabsdet, dim=0)
    aux = z.detach()
  else:
    pred = behavioural_forward(leaves, target.shape[1], visual_features, vel, tl, tls,
                               scalars[:, 5:6], train=train)
    loss = torch.mean(torch.sum(torch.abs(pred - target), dim=[-2, -1]), dim=0)
    aux = pred.detach()
  names = [k for k, v in leaves.items() if v.is_floating_point() and v.requires_grad]
  grads = torch.autograd.grad(loss, [leaves[k] for k in names])
  return loss.detach(), dict(zip(names, grads)), train["buffers"], aux


def adam_update(param, grad, exp_avg, exp_avg_sq, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                weight_decay=0.0):
  """torch.optim.Adam single-tensor update (dim/train.py:115-119), returns the new
  (param, exp_avg, exp_avg_sq)."""
  if weight_decay != 0.0:
    grad = grad + weight_decay * param
  exp_avg = betas[0] * exp_avg + (1 - betas[0]) * grad
  exp_avg_sq = betas[1] * exp_avg_sq + (1 - betas[1]) * grad * grad
  denom = exp_avg_sq.sqrt() / math.sqrt(1 - betas[1]**step) + eps
  return param - (lr / (1 - betas[0]**step)) * exp_avg / denom, exp_avg, exp_avg_sq


def clip_coefficient(grads, max_norm: float) -> float:
  """torch.nn.utils.clip_grad_norm_ (dim/train.py:207-208): min(1, max_norm/(||g||+1e-6))."""
  total = math.sqrt(sum(float((g.double()**2).sum()) for g in grads))
  return min(1.0, max_norm / (total + 1e-6))


# ----------------------------------------------------------------------------
# a12. plan post-processing (rip/agent.py:141-151)
# ----------------------------------------------------------------------------
def interpolate_plan(plan):
  """rip/agent.py:141-151 — linear interp of [T,2] at integer times, append z=0."""
  import numpy as np
  T = plan.shape[0]
  inc = 40 // T
  t_idx = np.arange(0, 40, inc)[:T].astype(np.float64)
  tq = np.arange(0, int(t_idx[-1])).astype(np.float64)
  xy = np.stack([np.interp(tq, t_idx, plan[:, d].astype(np.float64)) for d in range(2)], -1)
  return np.c_[xy, np.zeros((xy.shape[0], 1))]


# ----------------------------------------------------------------------------
# (f)4. LIDAR point cloud -> BEV histogram (oatomobile/utils/carla.py:165-233)
# ----------------------------------------------------------------------------
def lidar_bev(points, pixels_per_meter: int = 2, hist_max_per_pixel: int = 5, meters_max: int = 50):
  """utils/carla.py:181-231 — split at z=-2.5 (<= below, >= above), np.histogramdd over
  linspace(-m, m+1, 2*m*ppm+1) edges, clip at hist_max, divide, stack → [200,200,2] f32."""
  import numpy as np
  points = np.asarray(points, dtype=np.float32).reshape(-1, 3)
  edges = np.linspace(-meters_max, meters_max + 1, meters_max * 2 * pixels_per_meter + 1)
  feats = []
  for part in (points[points[..., 2] <= -2.5], points[points[..., 2] >= -2.5]):
    hist = np.histogramdd(part[..., :2], bins=(edges, edges))[0]
    hist[hist > hist_max_per_pixel] = hist_max_per_pixel
    feats.append(hist / hist_max_per_pixel)
  return np.stack(feats, axis=-1).astype(np.float32)
