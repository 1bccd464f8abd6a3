"""Thin functional wrappers: torch CUDA tensors in/out, kernels from liboat_b200.so.

PyTorch is only the allocator/stream provider here; every number is produced by
the hand-written sm_100a kernels behind the C-ABI (include/oat_b200.h).
"""
from typing import Optional, Tuple

import torch

from oatomobile_b200 import _native as N


def transform_visual(lidar: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
  """oatomobile/torch/transforms.py:34-49 — [B,C,H,W] -> [B,C,100,100] (resize + H<->W).
  `out` (optional) is a preallocated contiguous float32 CUDA result buffer."""
  lidar = N.require_cuda_f32(lidar, "lidar")
  B, C, H, W = lidar.shape
  if out is None:
    out = torch.empty(B, C, 100, 100, device=lidar.device, dtype=torch.float32)
  elif (tuple(out.shape) != (B, C, 100, 100) or out.dtype != torch.float32 or
        out.device != lidar.device or not out.is_contiguous()):
    raise ValueError("`out` must be a contiguous float32 [B,C,100,100] tensor on the input's device")
  with torch.cuda.device(lidar.device):
    N.check(N.lib().oat_transform_visual(lidar.data_ptr(), B, C, H, W, out.data_ptr(),
                                         N.stream_ptr(lidar.device)))
  return out


def transform_visual_hwc(lidar: torch.Tensor) -> torch.Tensor:
  """[B,H,W,C] (simulator / on-disk layout) -> [B,C,100,100]: HWC->CHW, resize, H<->W fused."""
  lidar = N.require_cuda_f32(lidar, "lidar")
  B, H, W, C = lidar.shape
  out = torch.empty(B, C, 100, 100, device=lidar.device, dtype=torch.float32)
  with torch.cuda.device(lidar.device):
    N.check(N.lib().oat_transform_visual_hwc(lidar.data_ptr(), B, H, W, C, out.data_ptr(),
                                             N.stream_ptr(lidar.device)))
  return out


def encode(ens: N.EnsembleHandle, visual: torch.Tensor, scalars: torch.Tensor) -> torch.Tensor:
  """dim/model.py:173-219 for all E models: -> z [E,B,64]."""
  visual = N.require_cuda_f32(visual, "visual_features")
  scalars = N.require_cuda_f32(scalars, "scalars")
  B = visual.shape[0]
  if tuple(visual.shape[2:]) != (100, 100):
    raise ValueError("`visual_features` must be [B,C,100,100] (apply `model.transform` first), "
                     "got %s" % (tuple(visual.shape),))
  # the C call receives only B: a wrong channel count or scalar width would be read
  # mis-strided / out of bounds, so both are checked here against the packed model
  if visual.shape[1] != ens.in_channels:
    raise ValueError("`visual_features` has %d channels, the model's stem expects %d"
                     % (visual.shape[1], ens.in_channels))
  if tuple(scalars.shape) != (B, ens.scalars):
    raise ValueError("context scalars must concatenate to [B,%d], got %s"
                     % (ens.scalars, tuple(scalars.shape)))
  z = torch.empty(len(ens), B, 64, device=visual.device, dtype=torch.float32)
  with torch.cuda.device(visual.device):
    N.check(N.lib().oat_encode(ens.ptr, visual.data_ptr(), scalars.data_ptr(), B, z.data_ptr(),
                               N.stream_ptr(visual.device)))
  return z


def encode_features(ens: N.EnsembleHandle, visual: torch.Tensor) -> torch.Tensor:
  """perception.py:53-55 for all E models: visual [B,C,100,100] -> features [E,B,128]."""
  visual = N.require_cuda_f32(visual, "x")
  B = visual.shape[0]
  if visual.dim() != 4 or tuple(visual.shape[2:]) != (100, 100):
    raise ValueError("the CUDA encoder is specialised for [B,C,100,100] inputs (the size "
                     "`ImitativeModel.transform` produces), got %s" % (tuple(visual.shape),))
  if visual.shape[1] != ens.in_channels:
    raise ValueError("input has %d channels, the stem expects %d" % (visual.shape[1], ens.in_channels))
  feat = torch.empty(len(ens), B, 128, device=visual.device, dtype=torch.float32)
  with torch.cuda.device(visual.device):
    N.check(N.lib().oat_encode_features(ens.ptr, visual.data_ptr(), B, feat.data_ptr(),
                                        N.stream_ptr(visual.device)))
  return feat


def mlp_forward(weights, biases, x: torch.Tensor, activate_final: bool) -> torch.Tensor:
  """mlp.py:70-72 for a Linear/ReLU stack: `weights[l]` [out,in], `biases[l]` [out] (device)."""
  import ctypes
  x = N.require_cuda_f32(x, "x")
  lead = x.shape[:-1]
  x2 = x.reshape(-1, x.shape[-1])
  ws = [N.require_cuda_f32(w, "weight") for w in weights]
  bs = [None if b is None else N.require_cuda_f32(b, "bias") for b in biases]
  sizes = [ws[0].shape[1]] + [w.shape[0] for w in ws]
  if x2.shape[1] != sizes[0] or any(ws[l].shape[1] != sizes[l] for l in range(len(ws))):
    raise ValueError("MLP: input / layer widths do not chain: %s vs input %d" % (sizes, x2.shape[1]))
  out = torch.empty(x2.shape[0], sizes[-1], device=x.device, dtype=torch.float32)
  vp = ctypes.c_void_p
  w_arr = (vp * len(ws))(*[w.data_ptr() for w in ws])
  b_arr = (vp * len(ws))(*[None if b is None else b.data_ptr() for b in bs])
  s_arr = (ctypes.c_int32 * len(sizes))(*sizes)
  with torch.cuda.device(x.device):
    N.check(N.lib().oat_mlp_forward(w_arr, b_arr, s_arr, len(ws), 1 if activate_final else 0,
                                    x2.data_ptr(), x2.shape[0], out.data_ptr(), N.stream_ptr(x.device)))
  return out.reshape(*lead, sizes[-1])


# (h, c) of the activation after `blocks` inverted-residual blocks (0 = stem)
_PREFIX_SHAPES = [(50, 32), (50, 16), (25, 24), (25, 24), (13, 32), (13, 32), (13, 32), (7, 64),
                  (7, 64), (7, 64), (7, 64), (7, 96), (7, 96), (7, 96), (4, 160), (4, 160),
                  (4, 160), (4, 320)]


def encoder_prefix(ens: N.EnsembleHandle, visual: torch.Tensor, blocks: int) -> torch.Tensor:
  """TEST HOOK — the encoder's activation after `blocks` blocks: -> [E,B,h,h,c] (NHWC)."""
  visual = N.require_cuda_f32(visual, "visual_features")
  B = visual.shape[0]
  h, c = _PREFIX_SHAPES[blocks]
  out = torch.empty(len(ens), B, h, h, c, device=visual.device, dtype=torch.float32)
  with torch.cuda.device(visual.device):
    N.check(N.lib().oat_debug_encoder_prefix(ens.ptr, visual.data_ptr(), B, blocks, out.data_ptr(),
                                             N.stream_ptr(visual.device)))
  return out


def flow_forward(model: N.ModelHandle, x: torch.Tensor, z: torch.Tensor,
                 rows_per_z: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
  """sequence.py:95-151 — x [N,T,2], z [N/rows_per_z,64] -> y, logabsdet."""
  x = N.require_cuda_f32(x, "x")
  z = N.require_cuda_f32(z, "z")
  n, T, _ = x.shape
  if z.shape[0] * rows_per_z != n:
    raise ValueError("z has %d rows, expected %d" % (z.shape[0], n // max(rows_per_z, 1)))
  y = torch.empty_like(x)
  lad = torch.empty(n, device=x.device, dtype=torch.float32)
  with torch.cuda.device(x.device):
    N.check(N.lib().oat_flow_forward(model.ptr, x.data_ptr(), z.data_ptr(), n, T, rows_per_z,
                                     y.data_ptr(), lad.data_ptr(), N.stream_ptr(x.device)))
  return y, lad


def flow_inverse(model: N.ModelHandle, y: torch.Tensor, z: torch.Tensor, rows_per_z: int = 1,
                 want_x: bool = True):
  """sequence.py:153-216 — y [N,T,2] -> x, log_prob, logabsdet."""
  y = N.require_cuda_f32(y, "y")
  z = N.require_cuda_f32(z, "z")
  n, T, _ = y.shape
  if z.shape[0] * rows_per_z != n:
    raise ValueError("z has %d rows, expected %d" % (z.shape[0], n // max(rows_per_z, 1)))
  x = torch.empty_like(y) if want_x else None
  lp = torch.empty(n, device=y.device, dtype=torch.float32)
  lad = torch.empty(n, device=y.device, dtype=torch.float32)
  with torch.cuda.device(y.device):
    N.check(N.lib().oat_flow_inverse(model.ptr, y.data_ptr(), z.data_ptr(), n, T, rows_per_z,
                                     N.ptr(x), lp.data_ptr(), lad.data_ptr(),
                                     N.stream_ptr(y.device)))
  return x, lp, lad


def rip_sample_score(ens: N.EnsembleHandle, z: torch.Tensor, x: Optional[torch.Tensor],
                     goal: Optional[torch.Tensor], epsilon: float, proposal_idx: int = 0,
                     y: Optional[torch.Tensor] = None, q: Optional[torch.Tensor] = None):
  """z [E,B,64], x [B,K,T,2] -> y [B,K,T,2], q [E,B,K].

  With `proposal_idx = -1` the proposals `y` are an input (sharded ensembles)."""
  z = N.require_cuda_f32(z, "z")
  E, B, _ = z.shape
  if E != len(ens):
    raise ValueError("z has %d models, the ensemble %d" % (E, len(ens)))
  if proposal_idx >= 0:
    x = N.require_cuda_f32(x, "x")
    _, K, T, _ = x.shape
    if y is None:
      y = torch.empty_like(x)
  else:
    y = N.require_cuda_f32(y, "y")
    _, K, T, _ = y.shape
  if goal is not None:
    goal = N.require_cuda_f32(goal, "goal")
  G = 0 if goal is None else goal.shape[1]
  if q is None:
    q = torch.empty(E, B, K, device=z.device, dtype=torch.float32)
  with torch.cuda.device(z.device):
    N.check(N.lib().oat_rip_sample_score(ens.ptr, proposal_idx, z.data_ptr(), N.ptr(x),
                                         N.ptr(goal), G, float(epsilon), B, K, T, y.data_ptr(),
                                         q.data_ptr(), N.stream_ptr(z.device)))
  return y, q


def rip_aggregate(q: torch.Tensor, y: Optional[torch.Tensor], algorithm: str,
                  want_s: bool = False):
  """rip/agent.py:121-127,137 — q [E,B,K] -> (kstar int32 [B], sbest [B], plan [B,T,2], s)."""
  if algorithm not in N.ALGORITHMS:
    raise AssertionError("algorithm must be one of WCM, MA, BCM")
  q = N.require_cuda_f32(q, "q")
  E, B, K = q.shape
  T = 0
  plan = None
  if y is not None:
    y = N.require_cuda_f32(y, "y")
    T = y.shape[2]
    plan = torch.empty(B, T, 2, device=q.device, dtype=torch.float32)
  s = torch.empty(B, K, device=q.device, dtype=torch.float32) if want_s else None
  kstar = torch.empty(B, device=q.device, dtype=torch.int32)
  sbest = torch.empty(B, device=q.device, dtype=torch.float32)
  with torch.cuda.device(q.device):
    N.check(N.lib().oat_rip_aggregate(q.data_ptr(), E, B, K, N.ALGORITHMS[algorithm], N.ptr(y), T,
                                      N.ptr(s), kstar.data_ptr(), sbest.data_ptr(), N.ptr(plan),
                                      N.stream_ptr(q.device)))
  return kstar, sbest, plan, s


def cil_rollout(model: N.ModelHandle, z: torch.Tensor, T: int) -> torch.Tensor:
  """cil/model.py:106-127 — z [B,64] -> y [B,T,2]."""
  z = N.require_cuda_f32(z, "z")
  B = z.shape[0]
  y = torch.empty(B, T, 2, device=z.device, dtype=torch.float32)
  with torch.cuda.device(z.device):
    N.check(N.lib().oat_cil_rollout(model.ptr, z.data_ptr(), B, T, y.data_ptr(),
                                    N.stream_ptr(z.device)))
  return y


def goal_likelihood(y: torch.Tensor, goal: torch.Tensor, epsilon: float = 1.0):
  """dim/model.py:143-171 — y [B,T,2], goal [B,G,2] -> (per-row [B], batch mean [])."""
  y = N.require_cuda_f32(y, "y")
  goal = N.require_cuda_f32(goal, "goal")
  B, G = goal.shape[0], goal.shape[1]
  y_last = y[:, -1, :].contiguous()
  rows = torch.empty(B, device=y.device, dtype=torch.float32)
  mean = torch.empty((), device=y.device, dtype=torch.float32)
  with torch.cuda.device(y.device):
    N.check(N.lib().oat_goal_likelihood(y_last.data_ptr(), goal.data_ptr(), B, G, float(epsilon),
                                        rows.data_ptr(), mean.data_ptr(), N.stream_ptr(y.device)))
  return rows, mean


def plan(models, z: torch.Tensor, x0: torch.Tensor, num_steps: int, lr: float,
         goal: Optional[torch.Tensor], epsilon: float, algorithm: Optional[str],
         want_loss: bool = False):
  """The gradient-based MAP planner of dim/model.py:97-141 (algorithm=None, one model)
  and rip/agent.py:84-137 (algorithm in WCM|BCM|MA), one fused kernel launch.

  models: list of `_native.ModelHandle`; z [E,B,64]; x0 [B,T,2] initial latent.
  Returns (plan [B,T,2], x_best [B,T,2], losses [num_steps] | None)."""
  import ctypes
  z = N.require_cuda_f32(z, "z")
  x = N.require_cuda_f32(x0, "x0").clone()
  E, B, _ = z.shape
  T = x.shape[1]
  if E != len(models):
    raise ValueError("z has %d models, got %d handles" % (E, len(models)))
  algo = -1 if algorithm is None else N.ALGORITHMS[algorithm]
  if goal is not None:
    goal = N.require_cuda_f32(goal, "goal")
  G = 0 if goal is None else goal.shape[1]
  x_best = torch.empty_like(x)
  out = torch.empty_like(x)
  nws = int(N.lib().oat_plan_workspace_floats(B, E, T))
  ws = torch.empty(nws, device=x.device, dtype=torch.float32)
  losses = torch.empty(max(num_steps, 1), device=x.device, dtype=torch.float32) if want_loss else None
  arr = (ctypes.c_void_p * E)(*[m.ptr.value for m in models])
  with torch.cuda.device(x.device):
    N.check(N.lib().oat_plan(arr, E, algo, z.data_ptr(), N.ptr(goal), G, float(epsilon), B, T,
                             int(num_steps), float(lr), x.data_ptr(), x_best.data_ptr(),
                             out.data_ptr(), ws.data_ptr(), nws, N.ptr(losses),
                             N.stream_ptr(x.device)))
  return out, x_best, losses


def lidar_bev(points: torch.Tensor, pixels_per_meter: int = 2, hist_max_per_pixel: int = 5,
              meters_max: int = 50) -> torch.Tensor:
  """oatomobile/utils/carla.py:165-233 — points [N,3] -> BEV histogram [200,200,2]."""
  points = N.require_cuda_f32(points, "points")
  n = points.shape[0]
  counts = torch.empty(200 * 200 * 2, device=points.device, dtype=torch.int32)
  out = torch.empty(200, 200, 2, device=points.device, dtype=torch.float32)
  with torch.cuda.device(points.device):
    N.check(N.lib().oat_lidar_bev(points.data_ptr() if n else None, n, pixels_per_meter,
                                  hist_max_per_pixel, meters_max, counts.data_ptr(),
                                  out.data_ptr(), N.stream_ptr(points.device)))
  return out
