"""Training steps of `ImitativeModel` (DIM) and `BehaviouralModel` (CIL) on the GPU.

Mirrors the closures of oatomobile/baselines/torch/dim/train.py:115-119,175-260 and
cil/train.py:113-118,168-225 — `optim.Adam(model.parameters(), lr, weight_decay)`,
`train_step`, `evaluate_step` — as one object.  The arithmetic (training-mode
MobileNetV2 forward/backward with BatchNorm batch statistics, merger, flow NLL / L1
roll-out with analytic BPTT, Adam, gradient clipping) runs in hand-written CUDA behind
`oat_train_forward_backward` / `oat_adam_step`; PyTorch owns the storage, the RNG
(target noise, dropout mask) and the optional gradient all-reduce.

All parameters are re-homed as views of ONE flat device buffer (gradients and the Adam
moments likewise), so the model keeps its reference `state_dict`, a data-parallel step
is one NCCL all-reduce over one buffer, and Adam is one kernel.
"""
import ctypes
from typing import Mapping, Optional

import torch
import torch.distributed as dist

from oatomobile_b200 import _native as N
from oatomobile_b200 import ops
from oatomobile_b200.models import BehaviouralModel, ImitativeModel, _scalars

_DIM_KEYS = ("velocity", "is_at_traffic_light", "traffic_light_state")
_CIL_KEYS = _DIM_KEYS + ("mode",)


def _bump_versions(tensors):
  """The kernels write through raw pointers; tell autograd-version-keyed caches (the
  packed inference weights, `_HandleCache`) that the values changed."""
  inc = getattr(torch.autograd.graph, "increment_version", None)
  for t in tensors:
    if inc is not None:
      inc(t)
    else:
      with torch.no_grad():
        t.add_(0)


class Trainer:
  """`train_step` / `evaluate_step` of dim/train.py and cil/train.py for one model.

  Args:
    model: an `oatomobile_b200.ImitativeModel` or `BehaviouralModel` on a CUDA device.
    lr, weight_decay: as `--learning_rate` / `--weight_decay` (defaults 1e-3 / 0).
    clip_gradients: `--clip_gradients` (global norm 1.0, dim/train.py:207-208).
    noise_level: std of the target perturbation (dim/train.py:100,185-189; DIM only).
    group: optional `torch.distributed` process group → data-parallel replicas; the
      gradients are summed over the group and scaled by 1/world inside the Adam kernel
      (the reference has no DDP).  BatchNorm statistics stay per replica.
    use_cuda_graphs: replay the ~510 launches of `forward_backward` as one CUDA graph per batch
      shape (off by default; `bench.py`'s training workloads switch it on): the step's inputs
      are copied into static buffers, the RNG draws (target noise, dropout mask) and the Adam
      launch stay outside the graph.  Validated on B200 against the plain path with a
      plain-vs-plain noise control (`tests/test_gpu_train.py::test_graphed_training_step_
      equals_plain_launches`); DIM B=64: 9.10 -> 8.01 ms per step.
  """

  def __init__(self, model, lr: float = 1e-3, weight_decay: float = 0.0,
               clip_gradients: bool = False, noise_level: float = 1e-2,
               betas=(0.9, 0.999), eps: float = 1e-8,
               group: Optional["dist.ProcessGroup"] = None, use_cuda_graphs: bool = False):
    if isinstance(model, ImitativeModel):
      self._kind, self._keys = N.KIND_DIM, _DIM_KEYS
    elif isinstance(model, BehaviouralModel):
      self._kind, self._keys = N.KIND_CIL, _CIL_KEYS
    else:
      raise TypeError("Trainer needs an oatomobile_b200 ImitativeModel or BehaviouralModel")
    self.model = model
    self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, tuple(betas), eps
    self.clip_gradients, self.noise_level = clip_gradients, noise_level
    self.group = group
    self.world = dist.get_world_size(group) if group is not None else 1
    self.step_count = 0
    self._use_graphs = bool(use_cuda_graphs)
    self._graphs = {}  # (B, T, C, has_mask) -> (graph, static inputs, static outputs)
    params = list(model.parameters())
    if not params or not params[0].is_cuda:
      raise N.NativeLibraryError(
          "Trainer: the model lives on the CPU; oatomobile_b200 trains on CUDA (sm_100a) only — "
          "call `.to('cuda')` first. There is no CPU fallback.")
    self.device = params[0].device
    self._flatten()
    self._bind()
    if self.world > 1:
      self.sync_replicas()

  # -- storage ------------------------------------------------------------------
  def _flatten(self):
    """Re-homes every parameter (and its gradient) into flat fp32 buffers, 16-byte aligned."""
    named = list(self.model.named_parameters())
    offsets, total = [], 0
    for _, p in named:
      offsets.append(total)
      total += (p.numel() + 3) // 4 * 4
    self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
    self.flat_grad = torch.zeros_like(self.flat)
    self.exp_avg = torch.zeros_like(self.flat)
    self.exp_avg_sq = torch.zeros_like(self.flat)
    self._norm_ws = torch.zeros(1, dtype=torch.float64, device=self.device)
    with torch.no_grad():
      for (name, p), off in zip(named, offsets):
        view = self.flat[off:off + p.numel()].view(p.shape)
        view.copy_(p.detach().float())
        p.data = view
        p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)
    self._named = named

  def _bind(self):
    arr, keep = [], []
    params = dict(self._named)
    for name, t in self.model.state_dict(keep_vars=True).items():
      if not torch.is_floating_point(t):
        continue  # num_batches_tracked is advanced on the host side
      if name not in params:  # BatchNorm running statistics
        if not t.is_contiguous() or t.dtype != torch.float32:
          raise N.NativeLibraryError("buffer `%s` must be contiguous float32" % name)
      grad = params[name].grad.data_ptr() if name in params else None
      shape = (ctypes.c_int64 * 4)(*(list(t.shape) + [0] * (4 - t.dim())))
      arr.append(N.OatTrainTensor(name.encode(), t.data_ptr(), grad, t.dim(), shape))
      keep.append(t)
    c_arr = (N.OatTrainTensor * len(arr))(*arr)
    out = ctypes.c_void_p()
    index = self.device.index if self.device.index is not None else torch.cuda.current_device()
    with torch.cuda.device(index):
      N.check(N.lib().oat_trainer_create(c_arr, len(arr), self._kind, index,
                                         self.flat_grad.data_ptr(), self.flat_grad.numel(),
                                         ctypes.byref(out)))
    self._ptr, self._keep = out, keep
    self._bn_counters = [b for n, b in self.model.named_buffers() if n.endswith("num_batches_tracked")]
    self._bn_stats = [b for n, b in self.model.named_buffers() if "running_" in n]

  def sync_replicas(self, src: int = 0) -> None:
    """What DDP does at construction: every replica starts from rank `src`'s parameters and
    BatchNorm buffers (ranks built with different seeds or checkpoints would otherwise drift
    apart while still averaging gradients).  Also called after `load_state_dict`."""
    if self.world <= 1:
      return
    g = dist.get_global_rank(self.group, src) if self.group is not None else src
    dist.broadcast(self.flat, src=g, group=self.group)
    for b in self._bn_stats + self._bn_counters:
      dist.broadcast(b, src=g, group=self.group)
    for t in (self.exp_avg, self.exp_avg_sq):
      dist.broadcast(t, src=g, group=self.group)
    step = torch.tensor([self.step_count], device=self.device, dtype=torch.int64)
    dist.broadcast(step, src=g, group=self.group)
    self.step_count = int(step.item())
    _bump_versions([p for _, p in self._named] + self._bn_stats)

  # -- optimiser state (the reference checkpoints only the model, savers.py:39-55; resuming
  #    Adam needs its moments and step, so they are exposed in `torch.optim` form) ---------
  def state_dict(self):
    """{"step", "exp_avg", "exp_avg_sq"}: moments keyed by parameter name (clones)."""
    out = {"step": self.step_count, "exp_avg": {}, "exp_avg_sq": {}}
    off = 0
    for name, p in self._named:
      n = p.numel()
      out["exp_avg"][name] = self.exp_avg[off:off + n].view(p.shape).clone()
      out["exp_avg_sq"][name] = self.exp_avg_sq[off:off + n].view(p.shape).clone()
      off += (n + 3) // 4 * 4
    return out

  def load_state_dict(self, state) -> None:
    self.step_count = int(state["step"])
    off = 0
    with torch.no_grad():
      for name, p in self._named:
        n = p.numel()
        self.exp_avg[off:off + n].copy_(state["exp_avg"][name].reshape(-1))
        self.exp_avg_sq[off:off + n].copy_(state["exp_avg_sq"][name].reshape(-1))
        off += (n + 3) // 4 * 4

  def __del__(self):
    try:
      if getattr(self, "_ptr", None):
        N.lib().oat_trainer_destroy(self._ptr)
        self._ptr = None
    except Exception:
      pass

  # -- the reference closures ----------------------------------------------------
  def forward_backward(self, batch: Mapping[str, torch.Tensor], target: torch.Tensor,
                       dropout_mask: Optional[torch.Tensor] = "sample"):
    """Training-mode forward + loss + backward; fills `p.grad` of every parameter.
    `dropout_mask`: "sample" draws the classifier's Dropout(0.2) mask with the torch RNG,
    None disables dropout, a tensor [B,1280] (already scaled by 1/(1-p)) is used as is.
    Returns (loss [1], z [B,64] or CIL predictions [B,T,2])."""
    visual = N.require_cuda_f32(batch["visual_features"], "visual_features")
    scalars = N.require_cuda_f32(_scalars(batch, self._keys), "scalars")
    target = N.require_cuda_f32(target, "target")
    B, T = target.shape[0], target.shape[1]
    if visual.shape[0] != B or tuple(visual.shape[2:]) != (100, 100):
      raise ValueError("visual_features must be [B,C,100,100] (apply `model.transform` first)")
    in_channels = self.model._encoder._model.features[0][0].in_channels
    if visual.shape[1] != in_channels:
      raise ValueError("visual_features has %d channels, the model's stem expects %d"
                       % (visual.shape[1], in_channels))
    if tuple(scalars.shape) != (B, len(self._keys) + 2):
      raise ValueError("context scalars must concatenate to [B,%d] (%s), got %s"
                       % (len(self._keys) + 2, ", ".join(self._keys), tuple(scalars.shape)))
    if isinstance(dropout_mask, str):
      keep = 1.0 - self.model._encoder._model.classifier[0].p
      dropout_mask = torch.bernoulli(torch.full((B, 1280), keep, device=self.device)) / keep
    elif dropout_mask is not None:
      dropout_mask = N.require_cuda_f32(dropout_mask, "dropout_mask")
    if self._use_graphs:
      loss, z, pred = self._forward_backward_graphed(visual, scalars, target, dropout_mask)
    else:
      loss, z, pred = self._forward_backward_native(visual, scalars, target, dropout_mask)
    with torch.no_grad():
      for c in self._bn_counters:
        c += 1
    _bump_versions(self._bn_stats)  # running_mean / running_var were updated in place
    return loss, (pred if pred is not None else z)

  def _forward_backward_native(self, visual, scalars, target, dropout_mask):
    B, T = target.shape[0], target.shape[1]
    loss = torch.empty(1, device=self.device)
    z = torch.empty(B, 64, device=self.device)
    pred = torch.empty(B, T, 2, device=self.device) if self._kind == N.KIND_CIL else None
    with torch.cuda.device(self.device):
      N.check(N.lib().oat_train_forward_backward(
          self._ptr, visual.data_ptr(), scalars.data_ptr(), target.data_ptr(), N.ptr(dropout_mask),
          B, T, loss.data_ptr(), z.data_ptr(), N.ptr(pred), N.stream_ptr(self.device)))
    return loss, z, pred

  def _forward_backward_graphed(self, visual, scalars, target, dropout_mask):
    """One CUDA-graph replay per step.  The graph bakes in the static input /
    output buffers and the trainer's activation workspace (which only grows with the batch
    shape, hence the key).  Outputs are the graph's static tensors: valid until the next step."""
    key = (tuple(visual.shape), tuple(scalars.shape), tuple(target.shape), dropout_mask is not None)
    entry = self._graphs.get(key)
    if entry is None:
      static = [visual.clone(), scalars.clone(), target.clone(),
                None if dropout_mask is None else dropout_mask.clone()]
      # un-captured first: workspace reservation and lazy initialisation; gradients and the
      # BatchNorm running statistics it changes are recomputed / restored below
      stats = [b.clone() for b in self._bn_stats]
      self._forward_backward_native(*static)
      with torch.no_grad():
        for b, s in zip(self._bn_stats, stats):
          b.copy_(s)
      torch.cuda.current_stream(self.device).synchronize()
      graph = torch.cuda.CUDAGraph()
      with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        out = self._forward_backward_native(*static)
      entry = (graph, static, out)
      self._graphs.clear()  # the workspace may have been reallocated for this shape
      self._graphs[key] = entry
    graph, static, out = entry
    with torch.no_grad():
      for dst, src in zip(static, (visual, scalars, target, dropout_mask)):
        if dst is not None:
          dst.copy_(src)
    graph.replay()
    return out

  def activation(self, index: int) -> torch.Tensor:
    """Post-activation output [rows, channels] of conv+BN unit `index` (0..51, NHWC rows) or
    merger layer 52..54 from the last `forward_backward` (debug / gradient checks)."""
    rows, ch = ctypes.c_int64(), ctypes.c_int32()
    L = N.lib()
    N.check(L.oat_trainer_activation(self._ptr, index, None, ctypes.byref(rows), ctypes.byref(ch), None))
    out = torch.empty(rows.value, ch.value, device=self.device)
    with torch.cuda.device(self.device):
      N.check(L.oat_trainer_activation(self._ptr, index, out.data_ptr(), ctypes.byref(rows),
                                       ctypes.byref(ch), N.stream_ptr(self.device)))
    return out

  def optimizer_step(self) -> None:
    """(all-reduce +) optional clip + `Adam.step()` over the flat parameter buffer."""
    if self.world > 1:
      dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
    self.step_count += 1
    with torch.cuda.device(self.device):
      N.check(N.lib().oat_adam_step(
          self.flat.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(),
          self.exp_avg_sq.data_ptr(), self.flat.numel(), self.step_count, self.lr, self.betas[0],
          self.betas[1], self.eps, self.weight_decay, 1.0 / self.world,
          1.0 if self.clip_gradients else 0.0, self._norm_ws.data_ptr(),
          N.stream_ptr(self.device)))
    _bump_versions([p for _, p in self._named])

  def train_step(self, batch: Mapping[str, torch.Tensor], clip: Optional[bool] = None) -> torch.Tensor:
    """dim/train.py:175-213 / cil/train.py:168-190 → the scalar loss (device tensor)."""
    if clip is not None:
      self.clip_gradients = clip
    target = batch["player_future"][..., :2]
    if self._kind == N.KIND_DIM:  # perturb the target (dim/train.py:185-189)
      target = torch.normal(mean=target, std=torch.ones_like(target) * self.noise_level)
    loss, _ = self.forward_backward(batch, target)
    self.optimizer_step()
    return loss[0]

  def evaluate_step(self, batch: Mapping[str, torch.Tensor]) -> torch.Tensor:
    """dim/train.py:231-250 / cil/train.py:205-219 — eval-mode loss through the inference kernels."""
    target = N.require_cuda_f32(batch["player_future"][..., :2], "player_future")
    if self._kind == N.KIND_DIM:
      z = self.model._params(**batch)
      _, log_prob, logabsdet = self.model._decoder._inverse(y=target, z=z)
      return -torch.mean(log_prob - logabsdet, dim=0)
    pred = self.model(**batch)
    return torch.mean(torch.sum(torch.abs(pred - target), dim=[-2, -1]), dim=0)
