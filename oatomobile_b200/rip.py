"""K-sample RIP sample-and-score (the BASELINE.json metric), single- and multi-GPU.

Assembled from the reference's own primitives:
  z_m   = models[m]._params(**ctx)                      dim/model.py:173-219
  y     = models[0]._decoder._forward(x, z_0)           rip/agent.py:106,137
  q_m   = log_prob_m(y) - logabsdet_m(y) (+ goal ll)    rip/agent.py:109-119
  s     = min|max|mean_m(-q_m)   ("WCM"|"BCM"|"MA")     rip/agent.py:121-127
  k*    = argmin_k s ; plan = y[b, k*]

Multi-GPU: the ensemble is sharded in contiguous blocks of
E/R models per rank.  One tiny broadcast of z_0 lets every rank decode proposals with
the (replicated, 61 KB) decoder of model 0: each rank decodes 1/R of the scenes and the
slices are all-gathered over NVLink (rows are independent, so the gathered y is
bit-identical to a single-GPU decode; when the scene count does not divide, every rank
regenerates all of y instead).  Each rank then scores y under its own models and a
single all-gather of the per-model scores q (`[E_local,B,K]` fp32 per rank) follows.
Every rank aggregates the same gathered tensor, so `k*` and the plan are bit-identical
on all ranks.
"""
from typing import Dict, Optional, Sequence

import torch

from oatomobile_b200 import _native as N
from oatomobile_b200 import ops
from oatomobile_b200.models import ImitativeModel, _CONTEXT_KEYS, _require, _scalars


class RIPScorer:
  """Scores K sampled trajectories per scene under an ensemble of ImitativeModels."""

  _MAX_GRAPHS = 4          # cached encoder graphs (distinct input buffers) per scorer
  _MAX_GRAPH_MISSES = 12   # input pointers keep changing -> graphs are switched off

  def __init__(self, models: Sequence[ImitativeModel], algorithm: str = "WCM",
               group=None, proposal_model: Optional[ImitativeModel] = None,
               use_cuda_graphs: bool = False) -> None:
    """Args:
      models: the models owned by THIS rank (all E of them on one GPU).
      algorithm: "WCM" | "MA" | "BCM", semantics as written at rip/agent.py:121-127.
      group: optional torch.distributed process group the ensemble is sharded over
        (rank r owns global models [r*E_local, (r+1)*E_local)).
      proposal_model: on ranks that do not own global model 0, a replica of it (only
        its flow decoder is used) so proposals can be regenerated locally.
      use_cuda_graphs: replay the ~55 launches of the encoder stage as ONE CUDA graph per set
        of input buffers (captured on first use; callers that feed the same device buffers
        every step, like `HostRIPPipeline`, hit the cache).  The returned `z` is then the
        graph's static output buffer: valid until the next call with the same input buffers.
    """
    assert algorithm in ("WCM", "MA", "BCM")  # rip/agent.py:43
    self._algorithm = algorithm
    self._models = list(models)
    self._group = group
    self._rank, self._world = 0, 1
    if group is not None:
      import torch.distributed as dist
      self._rank, self._world = dist.get_rank(group), dist.get_world_size(group)
    self._proposal_model = proposal_model
    if self._world > 1 and self._rank != 0 and proposal_model is None:
      raise ValueError("ranks other than 0 need `proposal_model` (a replica of global model 0)")
    self._ens = None
    self._ens_key = None
    self.stage_events = None  # set to a list to record (name, cuda event) marks per call
    self._use_graphs = bool(use_cuda_graphs)
    self._graphs = {}          # key -> (graph, z, inputs kept alive, launches per replay)
    self._graph_misses = 0
    self._graph_max_batch = 0
    self._vis_buf = None
    self.replayed_launches = 0  # kernel launches executed through graph replays

  def _mark(self, name: str) -> None:
    if self.stage_events is not None:
      ev = torch.cuda.Event(enable_timing=True)
      ev.record()
      self.stage_events.append((name, ev))

  # ---- handles ---------------------------------------------------------------
  def _ensemble(self) -> N.EnsembleHandle:
    handles = [m.native_handle() for m in self._models]
    key = tuple(id(h) for h in handles)
    if key != self._ens_key:
      self._ens = N.EnsembleHandle(handles)
      self._ens_key = key
    return self._ens

  @property
  def num_local_models(self) -> int:
    return len(self._models)

  # ---- stages ----------------------------------------------------------------
  def encode(self, **context: torch.Tensor) -> torch.Tensor:
    """E_local x `_params` in grouped launches → z [E_local,B,64]."""
    _require(context, _CONTEXT_KEYS)
    if not self._use_graphs:
      return ops.encode(self._ensemble(), context["visual_features"],
                        _scalars(context, _CONTEXT_KEYS[1:]))
    return self._encode_graphed(context)

  def _encode_graphed(self, context) -> torch.Tensor:
    """The encoder stage as one CUDA-graph replay.  The graph bakes in the input pointers,
    the ensemble's activation workspace and the TMA descriptors, so it is keyed by the input
    buffers and dropped whenever a larger batch makes the workspace grow."""
    ens = self._ensemble()
    tensors = [N.require_cuda_f32(context[k], k) for k in _CONTEXT_KEYS]
    batch = tensors[0].shape[0]
    if batch > self._graph_max_batch:  # `oat_ensemble_reserve` reallocates: old graphs dangle
      self._graphs.clear()
      self._graph_max_batch = batch
    key = (id(ens),) + tuple((t.data_ptr(), tuple(t.shape)) for t in tensors)
    entry = self._graphs.get(key)
    if entry is None:
      self._graph_misses += 1
      ctx = dict(zip(_CONTEXT_KEYS, tensors))
      run = lambda: ops.encode(ens, ctx["visual_features"], _scalars(ctx, _CONTEXT_KEYS[1:]))
      z = run()  # un-captured first: workspace reservation, kernel attributes, lazy init
      if self._graph_misses > self._MAX_GRAPH_MISSES:
        self._use_graphs = False  # the caller does not reuse its buffers: plain launches
        self._graphs.clear()
        return z
      torch.cuda.current_stream(tensors[0].device).synchronize()
      graph = torch.cuda.CUDAGraph()
      before = N.launch_count()
      try:
        # thread_local: CUDA calls of other threads (e.g. the NCCL watchdog of a process
        # group) must not invalidate this capture
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
          z = run()
      except Exception:  # capture not possible here (e.g. a foreign capture in progress):
        self._use_graphs = False  # keep working with one launch per kernel
        self._graphs.clear()
        torch.cuda.synchronize(tensors[0].device)
        return run()
      launches = N.launch_count() - before
      while len(self._graphs) >= self._MAX_GRAPHS:
        self._graphs.pop(next(iter(self._graphs)))
      entry = (graph, z, tensors, launches)
      self._graphs[key] = entry
    entry[0].replay()
    self.replayed_launches += entry[3]
    return entry[1]

  def _visual_buffer(self, lidar: torch.Tensor) -> Optional[torch.Tensor]:
    """With CUDA graphs the resized grids live in one persistent buffer per scorer (a fresh
    tensor per call would change the encoder graph's input pointer every step)."""
    if not self._use_graphs:
      return None
    shape = (lidar.shape[0], lidar.shape[1], 100, 100)
    if self._vis_buf is None or tuple(self._vis_buf.shape) != shape or self._vis_buf.device != lidar.device:
      self._vis_buf = torch.empty(shape, device=lidar.device, dtype=torch.float32)
    return self._vis_buf

  def score(self, z: torch.Tensor, x: torch.Tensor, goal: Optional[torch.Tensor] = None,
            epsilon: float = 1.0, want_s: bool = False) -> Dict[str, torch.Tensor]:
    """z [E_local,B,64], x [B,K,T,2] → plan/kstar/sbest (+ y, q, s)."""
    ens = self._ensemble()
    self._mark("flow_begin")
    if self._world == 1:
      y, q = ops.rip_sample_score(ens, z, x, goal, epsilon, proposal_idx=0)
      self._mark("flow_end")
    else:
      import torch.distributed as dist
      # (1) z_0 from the owner of model 0 (64 floats per scene).
      z0 = z[0].contiguous() if self._rank == 0 else torch.empty_like(z[0])
      dist.broadcast(z0, src=dist.get_global_rank(self._group, 0), group=self._group)
      Bn, Kn, Tn = x.shape[0], x.shape[1], x.shape[2]
      if Bn % self._world == 0:
        # (2) proposals: 1/R of the scenes per rank through the replicated decoder of model 0,
        # all-gathered over NVLink.  Measured before (every rank but 0 decoding ALL rows, then
        # scoring): the flow stage of ranks != 0 cost 2x rank 0's at one model per rank.
        n = Bn // self._world
        lo = self._rank * n
        prop = self._models[0] if self._rank == 0 else self._proposal_model
        y_part, _ = ops.flow_forward(prop._decoder._handle(), x[lo:lo + n].reshape(-1, Tn, 2),
                                     z0[lo:lo + n], rows_per_z=Kn)
        y = torch.empty_like(x)
        dist.all_gather_into_tensor(y, y_part.view(n, Kn, Tn, 2).contiguous(), group=self._group)
        _, q = ops.rip_sample_score(ens, z, None, goal, epsilon, proposal_idx=-1, y=y)
      elif self._rank == 0:
        y, q = ops.rip_sample_score(ens, z, x, goal, epsilon, proposal_idx=0)
      else:
        # (2') identical proposals, regenerated locally from the replicated decoder.
        y, _ = ops.flow_forward(self._proposal_model._decoder._handle(),
                                x.reshape(-1, Tn, 2), z0, rows_per_z=Kn)
        y = y.view_as(x)
        _, q = ops.rip_sample_score(ens, z, None, goal, epsilon, proposal_idx=-1, y=y)
      self._mark("flow_end")
      # (3) the single data-path collective: all-gather of per-model scores.
      q_all = torch.empty(self._world * q.shape[0], q.shape[1], q.shape[2], device=q.device,
                          dtype=q.dtype)
      dist.all_gather_into_tensor(q_all, q.contiguous(), group=self._group)
      q = q_all
    self._mark("aggregate_begin")
    kstar, sbest, plan, s = ops.rip_aggregate(q, y, self._algorithm, want_s=want_s)
    self._mark("aggregate_end")
    out = dict(plan=plan, kstar=kstar, sbest=sbest, y=y, q=q)
    if want_s:
      out["s"] = s
    return out

  def __call__(self, x: torch.Tensor, goal: Optional[torch.Tensor] = None, epsilon: float = 1.0,
               want_s: bool = False, **context: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Full step of the metric on device-resident inputs.  `context` holds either
    `lidar` [B,C,200,200] (raw) or `visual_features` [B,C,100,100] (transformed)."""
    self._mark("step_begin")
    if "lidar" in context:
      context = dict(context)
      lidar = context.pop("lidar")
      B = lidar.shape[0]
      if self._world > 1 and B % self._world == 0:
        # every rank needs every scene's features: resize 1/R of the scenes here and
        # all-gather the (4x smaller) result over NVLink instead of R redundant resizes
        import torch.distributed as dist
        lo = self._rank * (B // self._world)
        part = ops.transform_visual(lidar[lo:lo + B // self._world])
        vis = self._visual_buffer(lidar)
        if vis is None:
          vis = torch.empty((B,) + tuple(part.shape[1:]), device=part.device, dtype=part.dtype)
        dist.all_gather_into_tensor(vis, part, group=self._group)
        context["visual_features"] = vis
      else:
        buf = self._visual_buffer(lidar)
        context["visual_features"] = (ops.transform_visual(lidar) if buf is None
                                      else ops.transform_visual(lidar, out=buf))
    self._mark("encode_begin")
    z = self.encode(**context)
    self._mark("encode_end")
    out = self.score(z, x, goal, epsilon, want_s)
    out["z"] = z
    return out


class HostRIPPipeline:
  """End-to-end call with HOST buffers: pinned-memory inputs are copied to the GPU,
  scored, and the selected plans copied back — what an agent loop outside the GPU
  would call once per batch of observations (rip/agent.py:71-74,139 do the same
  H2D/D2H per tick).  Scenes are independent, so the batch is cut into `chunks`
  slices: a copy stream uploads slice i+1 (double-buffered device inputs) while the
  compute stream scores slice i, hiding the 174 MB/step of PCIe traffic behind the
  kernels.  Device and pinned result buffers are allocated once."""

  INPUT_KEYS = ("lidar", "velocity", "is_at_traffic_light", "traffic_light_state", "goal", "x")

  def __init__(self, scorer: RIPScorer, device, chunks: int = 1) -> None:
    self._scorer = scorer
    self._device = torch.device(device)
    self._chunks = max(1, int(chunks))
    self._dev = [{}, {}]            # double-buffered device inputs
    self._host_out = {}
    self._copy_stream = torch.cuda.Stream(device=self._device)
    self._uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
    self.h2d_bytes = 0
    self.d2h_bytes = 0

  def _upload(self, host, lo, hi, slot):
    """Async H2D of scenes [lo,hi) into buffer `slot` on the copy stream.

    When the ensemble is sharded over R ranks every rank needs every scene, but each rank
    pulls only its 1/R slice over PCIe; `_assemble` then all-gathers the slices over NVLink."""
    h2d = 0
    R, r = self._scorer._world, self._scorer._rank
    shard = R > 1 and (hi - lo) % R == 0
    with torch.cuda.stream(self._copy_stream):
      self._copy_stream.wait_event(self._consumed[slot])  # previous user of the slot is done
      for k in self.INPUT_KEYS:
        src = host[k][lo:hi]
        buf = self._dev[slot].get(k)
        if buf is None or buf.shape != src.shape:
          buf = torch.empty(src.shape, dtype=torch.float32, device=self._device)
          self._dev[slot][k] = buf
        if shard:
          n = (hi - lo) // R
          buf[r * n:(r + 1) * n].copy_(src[r * n:(r + 1) * n], non_blocking=True)
          h2d += src[r * n:(r + 1) * n].numel() * 4
        else:
          buf.copy_(src, non_blocking=True)
          h2d += src.numel() * 4
      self._uploaded[slot].record(self._copy_stream)
    self._sharded = shard
    return h2d

  def _assemble(self, slot):
    """Sharded uploads: all-gather every input in place (compute stream, NCCL/NVLink)."""
    if not getattr(self, "_sharded", False):
      return
    import torch.distributed as dist
    R, r = self._scorer._world, self._scorer._rank
    for k in self.INPUT_KEYS:
      buf = self._dev[slot][k]
      n = buf.shape[0] // R
      dist.all_gather_into_tensor(buf, buf[r * n:(r + 1) * n].clone(), group=self._scorer._group)

  def __call__(self, host: Dict[str, torch.Tensor], epsilon: float = 1.0):
    B = host["lidar"].shape[0]
    n = min(self._chunks, B)
    bounds = [(i * B // n, (i + 1) * B // n) for i in range(n)]
    compute = torch.cuda.current_stream(self._device)
    for ev in self._consumed:
      ev.record(compute)
    T2 = host["x"].shape[2:]
    for k, shape, dtype in (("plan", (B,) + tuple(T2), torch.float32), ("kstar", (B,), torch.int32),
                            ("sbest", (B,), torch.float32)):
      hb = self._host_out.get(k)
      if hb is None or tuple(hb.shape) != shape:
        self._host_out[k] = torch.empty(shape, dtype=dtype, pin_memory=True)
    h2d = self._upload(host, bounds[0][0], bounds[0][1], 0)
    d2h = 0
    for i, (lo, hi) in enumerate(bounds):
      slot = i & 1
      if i + 1 < n:  # prefetch the next slice while this one is scored
        h2d += self._upload(host, bounds[i + 1][0], bounds[i + 1][1], slot ^ 1)
      compute.wait_event(self._uploaded[slot])
      self._assemble(slot)
      d = dict(self._dev[slot])
      x, goal = d.pop("x"), d.pop("goal")
      out = self._scorer(x=x, goal=goal, epsilon=epsilon, **d)
      self._consumed[slot].record(compute)
      for k in ("plan", "kstar", "sbest"):
        self._host_out[k][lo:hi].copy_(out[k], non_blocking=True)
        d2h += out[k].numel() * out[k].element_size()
    compute.synchronize()  # results are now valid on the host
    self.h2d_bytes, self.d2h_bytes = h2d, d2h
    return {k: self._host_out[k] for k in ("plan", "kstar", "sbest")}

  def stream(self, batches, epsilon: float = 1.0):
    """Streaming form for a continuous feed of batches (an agent fleet / a replay): yields
    the host results of batch i while batch i+1 is already uploaded AND enqueued (inputs
    double-buffered on the device, results double-buffered in pinned memory), so neither the
    PCIe copies nor the host-side launch work of the next batch are exposed.  Every batch is
    still copied H2D from pinned memory and its plans are read back D2H."""
    compute = torch.cuda.current_stream(self._device)
    for ev in self._consumed:
      ev.record(compute)
    it = iter(batches)

    def enqueue(batch, slot):
      """upload + score + async D2H of one batch; returns (results, done event)."""
      h2d = self._upload(batch, 0, batch["lidar"].shape[0], slot)
      compute.wait_event(self._uploaded[slot])
      self._assemble(slot)
      d = dict(self._dev[slot])
      x, goal = d.pop("x"), d.pop("goal")
      out = self._scorer(x=x, goal=goal, epsilon=epsilon, **d)
      self._consumed[slot].record(compute)
      d2h, res = 0, {}
      for k in ("plan", "kstar", "sbest"):
        t = out[k]
        hb = self._host_out.get((k, slot))
        if hb is None or hb.shape != t.shape:
          hb = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
          self._host_out[(k, slot)] = hb
        hb.copy_(t, non_blocking=True)
        d2h += t.numel() * t.element_size()
        res[k] = hb
      done = torch.cuda.Event()
      done.record(compute)
      self.h2d_bytes, self.d2h_bytes = h2d, d2h
      return res, done

    pending = None
    i = 0
    for batch in it:
      cur = enqueue(batch, i & 1)
      if pending is not None:
        pending[1].synchronize()  # batch i-1 is on the host; batch i keeps the GPU busy
        yield pending[0]
      pending, i = cur, i + 1
    if pending is not None:
      pending[1].synchronize()
      yield pending[0]
