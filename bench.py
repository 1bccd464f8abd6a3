#!/usr/bin/env python
"""bench.py — RIP trajectory samples scored/sec (ens=4, K=512, T=10) on B200.

One "step" = one pass of the hot path over one batch of synthetic scenes:
  transform (200x200 -> 100x100) -> E x encoder+merger -> proposals from model 0
  -> scores under all E models -> WCM aggregation -> argmin -> plan.

Workloads (`--workload`):
  rip        (default) BASELINE.json configs[2]: B=256 scenes/GPU, E=4, K=512, T=10, C=4; weak
             scaling, the ensemble sharded E/min(N,E) models per rank, N/E replica groups beyond
             E ranks.  At N=8 the line additionally carries a `cfg5` object: BASELINE configs[4]
             (E=8, K=2048, one model per GPU, B=256) measured in the same run.
  cfg5       BASELINE configs[4] as the line's own workload (any N dividing 8; strong scaling).
  train-dim  BASELINE configs[1]: DIM `train_step`, batch 64, 1 GPU (data parallel beyond).
  train-cil  BASELINE configs[3]: CIL `train_step`, batch 512 over the ranks (64/GPU at N=8).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path
  torchrun ... bench.py --gpus N ...                       # N > 1, one rank per GPU

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the launch stream with
inputs resident in HBM; `e2e` is the same metric through the public host-buffer API (pinned
H2D of every input + D2H of the results inside the timed region); `roofline` describes the
DOMINANT kernel family of the step: after the timed region the same step runs a few more
times with one launch per kernel while the library records a CUDA event behind every launch
(`oat_profile_begin/_end`), which gives each family's device time live in this run; its
algorithmic flops/bytes are the SURVEY.md §8(d) figures.  `cpu_baseline` / `--impl reference`
time the imported reference (`baseline/_ref` or /root/reference through oracle/reference_arm.py,
kind "reference") or, where no reference tree exists, the in-repo restatement (kind "port").
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

G_GOALS = 10
FLOW_FLOP_PER_ROW_STEP = 29696          # 2*(64*192 + 2*192 + 64*32 + 32*4)
METRIC = "RIP trajectory samples scored/sec (ens=4,K=512,T=10)"

RIP_WORKLOADS = {
    # scenes: per GPU (weak) or in total (strong)
    "rip": dict(E=4, K=512, T=10, C=4, scenes=256, scaling="weak", baseline_config=2,
                name="BASELINE configs[2]: RIPAgent WCM sample-and-score, 4-model ensemble, K=512, "
                     "T=10, 256 scenes/GPU of 200x200x4 BEV grids"),
    "cfg5": dict(E=8, K=2048, T=10, C=4, scenes=256, scaling="strong", baseline_config=4,
                 name="BASELINE configs[4]: RIP 8-model ensemble, K=2048, T=10, sharded "
                      "one-model-per-GPU, NCCL all-gather of per-model scores, 256 scenes of "
                      "200x200x4 BEV grids in total"),
}
TRAIN_WORKLOADS = {
    "train-dim": dict(kind="dim", batch=64, per_gpu=True, T=4, C=2, baseline_config=1,
                      name="BASELINE configs[1]: DIM (ImitativeModel) train_step, batch 64 per GPU, "
                           "synthetic 200x200x2 episodes, T=4"),
    "train-cil": dict(kind="cil", batch=512, per_gpu=False, T=4, C=2, baseline_config=3,
                      name="BASELINE configs[3]: BehaviouralModel (CIL) train_step, batch 512 split "
                           "over the ranks (data parallel, one gradient all-reduce), T=4"),
}

# ---- encoder layer table (torchvision MobileNetV2 as wrapped by perception.py:25-55) --------
_MBV2 = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
         (6, 320, 1, 1))


def encoder_layers(C):
  """[(kind, name, pixels_out, K, N)] per image; kind in stem|expand|dw|project|last|fc."""
  out = [("stem", "features.0", 2500, 9 * C, 32)]
  h, cin, idx = 50, 32, 1
  for t, c, n, s in _MBV2:
    for i in range(n):
      stride = s if i == 0 else 1
      hid, hout = cin * t, (h + stride - 1) // stride if stride == 2 else h
      if t != 1:
        out.append(("expand", "features.%d" % idx, h * h, cin, hid))
      out.append(("dw", "features.%d" % idx, hout * hout, 9, hid))
      out.append(("project", "features.%d" % idx, hout * hout, hid, c))
      h, cin, idx = hout, c, idx + 1
  out.append(("last", "features.18", h * h, cin, 1280))
  out.append(("fc", "classifier.1", 1, 1280, 128))
  return out


def layer_flops(l):
  kind, _, px, K, N = l
  return 2.0 * px * K * N  # depthwise: K = 9 taps, N = channels


def _peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    p = json.load(open(path))
    return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                source="measured (MEASURED_PEAKS.json)")
  return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
              source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
  """Samples SM clock / throttle reasons of one GPU during the timed region."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
    self._stop_evt = threading.Event()
    self.ok = False
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
      self.ok = True
    except Exception:
      self.ok = False

  def run(self):
    if not self.ok:
      return
    nv = self.nv
    names = {
        "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
        "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
        "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
        "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
    }
    while not self._stop_evt.is_set():
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for n, bit in names.items():
          if mask & bit:
            self.reasons.add(n)
      except Exception:
        pass
      time.sleep(0.02)

  def stop(self):
    self._stop_evt.set()
    if self.ok:
      self.join(timeout=2)
    return dict(sm_mhz=(statistics.median(self.samples) if self.samples else None),
                sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                samples=len(self.samples))


# ----------------------------------------------------------------------------------
# CPU arm: the reference's own PyTorch-CPU path (imported) or, without a reference tree,
# the oracle restatement — all usable host threads, explicit thread count.
# ----------------------------------------------------------------------------------
def usable_cores():
  """Cores this process may actually run on (affinity mask and cgroup CPU quota)."""
  n = os.cpu_count() or 1
  try:
    n = min(n, len(os.sched_getaffinity(0)))
  except Exception:
    pass
  try:
    quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
    if quota != "max":
      n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
  except Exception:
    pass
  return max(1, n)


def pick_cpu_threads(run_once):
  """PyTorch-CPU throughput on this path collapses when the intra-op pool is larger than what
  the small GRU/conv ops can use (measured: 128 threads = 400x slower than 8).  Try ascending
  pool sizes on one small call each and keep the fastest.  The count is set explicitly with
  `torch.set_num_threads`, so torchrun's OMP_NUM_THREADS=1 default does not starve the arm."""
  import torch
  cap = usable_cores()
  cands = sorted({c for c in (4, 8, 16, 32, 64, cap) if c <= cap}) or [1]
  best, best_t = cands[0], float("inf")
  for c in cands:
    torch.set_num_threads(c)
    run_once()  # warm the pool
    t0 = time.perf_counter()
    run_once()
    dt = time.perf_counter() - t0
    if dt < best_t:
      best, best_t = c, dt
    elif dt > 2.0 * best_t:
      break  # past the knee: larger pools only get slower
  torch.set_num_threads(best)
  return best


def make_cpu_rip_step(cfg, scenes):
  """Returns (step(), kind): one full step of the metric on `scenes` scenes on the CPU."""
  import torch
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  E, K, T, C = cfg["E"], cfg["K"], cfg["T"], cfg["C"]
  inp = synthetic_inputs(scenes, C, K, T, G_GOALS, seed=0)
  sds = [synthetic_state_dict("dim", C, 100 + m) for m in range(E)]
  args = [inp[k] for k in ("lidar", "velocity", "is_at_traffic_light", "traffic_light_state", "x", "goal")]
  kind = "port"
  try:
    from oracle import reference_arm as RA
    if RA.available():
      models = RA.build_models(sds, T, C)
      kind = "reference"
  except Exception as e:  # an unusable reference tree must not take the bench down
    sys.stderr.write("bench: reference tree not usable (%r); timing the restatement\n" % (e,))
    kind = "port"
  if kind == "reference":
    def step(n=scenes, k=K):
      return RA.rip_score(models[:E], *[a[:n] for a in args[:4]], args[4][:n, :k], args[5][:n], 1.0, "WCM")
  else:
    from oracle import restatement as R  # the CPU baseline IS the oracle

    def step(n=scenes, k=K):
      with torch.no_grad():
        return R.rip_score_from_inputs(sds, *[a[:n] for a in args[:4]], args[4][:n, :k], args[5][:n], 1.0, "WCM")
  return step, kind


def cpu_rip_run(cfg, steps, warmup, scenes):
  step, kind = make_cpu_rip_step(cfg, scenes)
  cores = pick_cpu_threads(lambda: step(1, 64))
  times = []
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    step()
    if i >= warmup:
      times.append(time.perf_counter() - t0)
  total = sum(times)
  return dict(value=scenes * cfg["K"] * len(times) / total, ms=1e3 * total / len(times),
              cores=cores, kind=kind, scenes=scenes)


def make_cpu_train_step(tcfg, batch):
  """The reference's train_step arithmetic on the CPU via the oracle restatement (torch
  autograd, training-mode BatchNorm, Adam) — kind "port"."""
  import torch
  from oracle import restatement as R
  from oatomobile_b200.synthetic import synthetic_state_dict
  from tests.helpers import train_inputs
  cfg = dict(kind=tcfg["kind"], T=tcfg["T"], C=tcfg["C"], B=batch, wseed=400, iseed=21)
  visual, scalars, target = train_inputs(cfg)
  state = {k: v.clone() for k, v in synthetic_state_dict(tcfg["kind"], tcfg["C"], 400).items()}
  moments = {}

  def step():
    loss, grads, bufs, _ = R.train_forward_backward(state, tcfg["kind"], visual, scalars, target)
    for k, g in grads.items():
      m, v = moments.get(k, (torch.zeros_like(g), torch.zeros_like(g)))
      state[k], m, v = R.adam_update(state[k], g, m, v, 1)
      moments[k] = (m, v)
    state.update(bufs)
  return step


def cpu_train_run(tcfg, steps, warmup, batch):
  step = make_cpu_train_step(tcfg, batch)
  cores = pick_cpu_threads(step)
  for _ in range(max(warmup - 2, 0)):
    step()
  t0 = time.perf_counter()
  for _ in range(steps):
    step()
  ms = 1e3 * (time.perf_counter() - t0) / steps
  return dict(value=batch / ms * 1e3, ms=ms, cores=cores, kind="port", batch=batch)


def run_reference_arm(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return  # only rank 0 runs the CPU arm; the others exit 0 without work
  warm = max(args.warmup, 1)
  if args.workload in TRAIN_WORKLOADS:
    tcfg = TRAIN_WORKLOADS[args.workload]
    batch = args.cpu_batch
    r = cpu_train_run(tcfg, min(args.steps, 5), min(warm, 2), batch)
    metric = train_metric(tcfg)
    sample = ("%d-sample batch per step (%s train_step: training-mode forward, loss, backward, Adam), "
              "torch CPU fp32 autograd through oracle/restatement.py, %d threads" %
              (batch, tcfg["kind"].upper(), r["cores"]))
    config = {"workload": tcfg["name"] + "; CPU arm on a %d-sample batch" % batch, "batch": batch,
              "T": tcfg["T"], "bev_channels": tcfg["C"]}
    steps = min(args.steps, 5)
  else:
    cfg = RIP_WORKLOADS[args.workload]
    scenes = args.cpu_scenes or cfg["scenes"]
    r = cpu_rip_run(cfg, args.steps, warm, scenes)
    metric = METRIC if args.workload == "rip" else rip_metric(cfg)
    sample = ("%d scenes x K=%d samples per step (E=%d, T=%d, C=%d), the %s on torch CPU fp32, %d threads"
              % (scenes, cfg["K"], cfg["E"], cfg["T"], cfg["C"],
                 "imported reference modules (oracle/reference_arm.py)" if r["kind"] == "reference"
                 else "in-repo restatement (oracle/restatement.py)", r["cores"]))
    config = {"workload": cfg["name"] + ("" if scenes == cfg["scenes"] else
                                         "; bounded sample of %d scenes per step" % scenes),
              "ensemble": cfg["E"], "K": cfg["K"], "T": cfg["T"], "bev_channels": cfg["C"],
              "scenes": scenes}
    steps = args.steps
  line = {
      "impl": "reference", "metric": metric, "value": r["value"], "unit": "samples/s",
      "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["ms"],
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic", "config": config,
      "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"],
                       "kind": r["kind"], "sample": sample},
      "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


def rip_metric(cfg):
  return "RIP trajectory samples scored/sec (ens=%d,K=%d,T=%d)" % (cfg["E"], cfg["K"], cfg["T"])


def train_metric(tcfg):
  return "%s training samples/sec (train_step: forward+loss+backward+Adam, T=%d)" % (
      tcfg["kind"].upper(), tcfg["T"])


def pin_to_gpu_numa_node(index):
  """Multi-rank runs: bind this process to the CPU cores NVML reports as local to its GPU before
  any pinned host buffer is allocated, so the e2e arm's H2D source pages live on the GPU's own
  NUMA node instead of all ranks pulling from one node.  Returns the number of cores, or None."""
  try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    words = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
    cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
    allowed = os.sched_getaffinity(0)
    cpus = (cpus & allowed) or allowed
    os.sched_setaffinity(0, cpus)
    return len(cpus)
  except Exception:
    return None


# ----------------------------------------------------------------------------------
# GPU arm: RIP sample-and-score
# ----------------------------------------------------------------------------------
class Dist:
  """Process-group plumbing shared by the workloads of one run."""

  def __init__(self, args):
    import torch
    import torch.distributed as dist
    self.world = int(os.environ.get("WORLD_SIZE", "1"))
    self.rank = int(os.environ.get("RANK", "0"))
    self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if self.world != args.gpus and self.world == 1 and args.gpus > 1:
      raise SystemExit("launch N>1 with torchrun (one rank per GPU)")
    torch.cuda.set_device(self.local_rank)
    self.dev = torch.device("cuda", self.local_rank)
    self.numa = pin_to_gpu_numa_node(self.local_rank) if self.world > 1 else None
    if self.world > 1:
      dist.init_process_group("nccl", device_id=self.dev)
    self.dist = dist

  def groups(self, gsize):
    """Splits the world into world/gsize groups of consecutive ranks; returns this rank's."""
    mine = None
    if self.world > 1:
      for g in range(self.world // gsize):
        pg = self.dist.new_group(list(range(g * gsize, (g + 1) * gsize)))
        if g == self.rank // gsize:
          mine = pg
    return mine

  def barrier(self):
    import torch
    if self.world > 1:
      self.dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(self, values):
    import torch
    t = torch.tensor(values, device=self.dev, dtype=torch.float64)
    if self.world > 1:
      self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
    return [float(v) for v in t]

  def all_true(self, flag):
    import torch
    t = torch.tensor([1.0 if flag else 0.0], device=self.dev)
    if self.world > 1:
      self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
    return bool(t.item() == 1.0)


def rip_family_work(cfg, e_local, scenes, rows_sample, rows_score, fusion):
  """Algorithmic work per step and rank of every kernel family: (flop, bytes, bound)."""
  C, K, T = cfg["C"], cfg["K"], cfg["T"]
  img = float(e_local * scenes)  # (model, image) pairs encoded by this rank
  L = encoder_layers(C)
  f = lambda pred: sum(layer_flops(l) for l in L if pred(l))
  fused_expand = {"features.2", "features.3", "features.4"} if fusion & 0b1110 else set()
  fused_f1 = bool(fusion & 16)
  # bit 5: expand + depthwise of features.5-17 inside the tcgen05 GEMM (depthwise epilogue)
  late = {"features.%d" % i for i in range(5, 18)} if fusion & 32 else set()
  fam = {}
  fam["transform"] = (0.0, scenes * C * (200 * 200 + 100 * 100) * 4.0, "hbm")
  fam["stem"] = (img * f(lambda l: l[0] == "stem"),
                 scenes * C * 100 * 100 * 4.0 + img * 2500 * 32 * 4.0, "hbm")
  fam["fused_dw_project"] = (img * f(lambda l: l[1] == "features.1" and fused_f1),
                             img * 2500 * (32 + 16) * 4.0, "hbm")
  fam["fused_expand_dw"] = (img * f(lambda l: l[1] in fused_expand and l[0] in ("expand", "dw")),
                            img * sum(2500 * 16 + 625 * 96 if n == "features.2" else
                                      625 * 24 + 625 * 144 if n == "features.3" else
                                      625 * 24 + 169 * 144 for n in fused_expand) * 4.0, "hbm")
  # block input (expand: pixels x K) + depthwise output (pixels_out x channels)
  fam["tc_expand_dw"] = (img * f(lambda l: l[1] in late and l[0] in ("expand", "dw")),
                         img * sum(l[2] * (l[3] if l[0] == "expand" else l[4]) for l in L
                                   if l[1] in late and l[0] in ("expand", "dw")) * 4.0, "tensor")
  pw = lambda l: (l[0] in ("expand", "project", "last", "fc") and
                  not (l[0] == "expand" and (l[1] in fused_expand or l[1] in late)) and
                  not (l[0] == "project" and l[1] == "features.1" and fused_f1))
  fam["tc_pw_gemm"] = (img * f(pw), img * sum((l[3] + l[4]) * l[2] for l in L if pw(l)) * 4.0, "tensor")
  dwl = lambda l: (l[0] == "dw" and l[1] not in fused_expand and l[1] not in late and
                   not (l[1] == "features.1" and fused_f1))
  fam["depthwise"] = (img * f(dwl), img * sum(2.0 * l[2] * l[4] for l in L if dwl(l)) * 4.0, "hbm")
  fam["pool"] = (0.0, img * (16 + 1) * 1280 * 4.0, "hbm")
  fam["merger"] = (img * 2.0 * (133 * 64 + 64 * 64 + 64 * 64), img * (133 + 64) * 4.0, "hbm")
  fam["flow_sample"] = (rows_sample * T * FLOW_FLOP_PER_ROW_STEP, rows_sample * (T * 16 + 4.0), "tensor")
  fam["flow_score"] = (rows_score * T * FLOW_FLOP_PER_ROW_STEP, rows_score * (T * 8 + 4.0), "tensor")
  fam["aggregate"] = (0.0, scenes * K * 4.0 * cfg["E"] + scenes * T * 8.0, "hbm")
  return fam


def build_roofline(prof, psteps, work, peaks, clocks):
  """`roofline` of the dominant family + the table of all families (ms per step, live)."""
  tensor_peak, hbm_peak = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
  traffic = {}
  tpath = os.path.join(ROOT, "profiles", "traffic.json")
  if os.path.exists(tpath):
    traffic = json.load(open(tpath))
  rows = {}
  for name, p in prof.items():
    if not name:
      continue
    ms = p["ms"] / psteps
    flop, byts, bound = work.get(name, (0.0, 0.0, "hbm"))
    row = {"ms_per_step": ms, "launches_per_step": p["launches"] / psteps, "bound": bound,
           "algorithmic_flop": flop, "algorithmic_bytes": byts}
    if ms > 0:
      row["tflops"] = flop / (ms * 1e-3) / 1e12
      row["gbs"] = byts / (ms * 1e-3) / 1e9
      row["frac"] = (row["tflops"] / tensor_peak) if bound == "tensor" else (row["gbs"] / hbm_peak)
    rows[name] = row
  if not rows:
    return None
  top = max((n for n in rows if not n.startswith("(")), key=lambda n: rows[n]["ms_per_step"])
  r = rows[top]
  tensor = r["bound"] == "tensor"
  t = traffic.get(top)
  out = {
      "kernel": top, "bound": r["bound"],
      "achieved": r.get("tflops" if tensor else "gbs", 0.0),
      "peak": tensor_peak if tensor else hbm_peak, "unit": "TFLOP/s" if tensor else "GB/s",
      "frac": r.get("frac", 0.0),
      "traffic": (t or {}).get("dram_bytes_per_step") if isinstance(t, dict) else t,
      "traffic_source": (t or {}).get("source") if isinstance(t, dict) else None,
      "peak_source": peaks["source"] + (", bf16 dense sustained" if tensor else ", copy bandwidth"),
      "ms_per_step": r["ms_per_step"], "launches_per_step": r["launches_per_step"],
      "algorithmic_flop_per_step": r["algorithmic_flop"], "algorithmic_bytes_per_step": r["algorithmic_bytes"],
      "how": "CUDA events recorded by the library behind every launch of %d extra un-graphed steps "
             "right after the timed region (oat_profile_begin/_end); dominant = largest share" % psteps,
      "kernels": rows,
  }
  if tensor:
    out["pipe"] = ("tcgen05.mma kind::tf32 with 3xTF32 error compensation: 3 TF32 MMAs per algorithmic "
                   "MMA at half the bf16 rate -> the ceiling of this formulation is peak/6")
    out["frac_of_3xtf32_ceiling"] = out["frac"] * 6.0
  sm_mhz = clocks.get("sm_mhz") or 0
  if sm_mhz:
    out["fp32_simt_peak_tflops"] = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
  return out


def run_rip(D, cfg, args, steps, warmup, measure_e2e=True, with_profile=True):
  """Builds the (sharded) ensemble for `cfg` and measures it.  Returns a dict of results
  (identical on every rank where it matters; rank 0 prints)."""
  import torch
  import oatomobile_b200 as ob
  from oatomobile_b200 import _native
  from oatomobile_b200.rip import HostRIPPipeline, RIPScorer
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  E, K, T, C = cfg["E"], cfg["K"], cfg["T"], cfg["C"]
  world, rank, dev = D.world, D.rank, D.dev

  gsize = min(world, E)                   # ranks sharing one ensemble
  assert E % gsize == 0 and world % gsize == 0, "world size must divide / be a multiple of E"
  e_local = E // gsize
  group_id, grank = rank // gsize, rank % gsize
  group = D.groups(gsize) if gsize > 1 else None
  n_groups = world // gsize
  if cfg["scaling"] == "weak":
    scenes = cfg["scenes"] * gsize        # scenes scored by this replica group per step
    total_scenes = cfg["scenes"] * world
  else:
    assert n_groups == 1, "strong-scaling workloads use one ensemble group"
    scenes = total_scenes = cfg["scenes"]

  sds = {m: synthetic_state_dict("dim", C, 100 + m) for m in range(E)}

  def make(m):
    model = ob.ImitativeModel(output_shape=(T, 2), in_channels=C)
    model.load_state_dict(sds[m], strict=True)
    return model.to(dev).eval()

  models = [make(m) for m in range(grank * e_local, (grank + 1) * e_local)]
  proposal = None if (grank == 0 or args.flow_sharding == "scenes") else make(0)
  scorer = RIPScorer(models, "WCM", group=group if gsize > 1 else None, proposal_model=proposal,
                     use_cuda_graphs=not args.no_cuda_graphs, flow_sharding=args.flow_sharding)

  inp = synthetic_inputs(scenes, C, K, T, G_GOALS, seed=group_id)
  host = {k: v.pin_memory() for k, v in inp.items()}
  d = {k: v.to(dev) for k, v in inp.items()}
  x, goal = d.pop("x"), d.pop("goal")

  def step():  # the metric's outputs: plan, kstar, sbest for every scene (no proposal/score tensors gathered)
    return scorer(x=x, goal=goal, epsilon=1.0, gather_details=False, **d)

  for _ in range(max(warmup, 3)):
    step()
  D.barrier()

  # ---- correctness inside the run: sharded result == one GPU holding the whole ensemble -----
  equal = None
  if gsize > 1:
    nchk = 8 if 8 % gsize == 0 else gsize
    sl = slice(0, nchk)
    full = RIPScorer([make(m) for m in range(E)], "WCM")
    a = scorer(x=x[sl], goal=goal[sl], epsilon=1.0, **{k: v[sl] for k, v in d.items()})
    b = full(x=x[sl], goal=goal[sl], epsilon=1.0, **{k: v[sl] for k, v in d.items()})
    n = nchk // gsize
    ls = slice(grank * n, (grank + 1) * n)
    c = scorer(x=x[ls], goal=goal[ls], epsilon=1.0, local_slice=True, **{k: v[ls] for k, v in d.items()})
    torch.cuda.synchronize()
    equal = D.all_true(all(torch.equal(a[k], b[k]) and torch.equal(c[k], b[k]) for k in ("q", "kstar", "plan")))
    del full, a, b, c
    for _ in range(2):
      step()  # the ensemble workspace / graphs are back at the full batch
    D.barrier()

  # ---- timed region: device-resident inputs (BEV grids larger than L2) ----------------------
  sampler = ClockSampler(D.local_rank)
  sampler.start()
  scorer.stage_events = []
  launches0 = _native.launch_count() + scorer.replayed_launches
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  D.barrier()
  ev0.record()
  for _ in range(steps):
    step()
  ev1.record()
  D.barrier()
  launches = _native.launch_count() + scorer.replayed_launches - launches0  # incl. graph-replayed kernels
  clocks = sampler.stop()
  elapsed_ms = ev0.elapsed_time(ev1)
  marks = scorer.stage_events
  scorer.stage_events = None

  def stage_ms(a, b):
    ea = [e for n, e in marks if n == a]
    eb = [e for n, e in marks if n == b]
    return sum(s.elapsed_time(t) for s, t in zip(ea, eb)) / max(len(ea), 1)

  stages = {"transform": stage_ms("step_begin", "encode_begin"),
            "encode": stage_ms("encode_begin", "encode_end"),
            "flow": stage_ms("flow_begin", "flow_end"),
            "aggregate": stage_ms("aggregate_begin", "aggregate_end")}

  # ---- per-family device time, live: a few more steps, one launch per kernel ----------------
  prof, psteps = {}, 3
  if with_profile:
    graphs = scorer._use_graphs
    scorer._use_graphs = False
    step()
    torch.cuda.synchronize()
    _native.profile_begin(dev)
    for _ in range(psteps):
      step()
    prof = _native.profile_end()
    scorer._use_graphs = graphs
    step()
    D.barrier()

  # ---- e2e: host (pinned) buffers through the public pipeline -------------------------------
  e2e = None
  if measure_e2e:
    pipe = HostRIPPipeline(scorer, dev, chunks=args.e2e_chunks)
    for _ in pipe.stream(host for _ in range(3)):
      pass
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    checksum = 0.0
    for res in pipe.stream(host for _ in range(steps)):
      checksum += float(res["plan"][0, 0, 0])  # the host really reads every step's result
    e1.record()
    D.barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t_host0))
    l0 = time.perf_counter()
    nb = max(steps // 4, 3)
    for _ in range(nb):  # latency form: one blocking call per step (no cross-step overlap)
      pipe(host)
    blocking_ms = 1e3 * (time.perf_counter() - l0) / nb
    e2e = dict(ms=e2e_ms, checksum=checksum, blocking_ms=blocking_ms,
               h2d=int(pipe.h2d_bytes), d2h=int(pipe.d2h_bytes))

  elapsed_ms, e2e_ms = D.max_over_ranks([elapsed_ms, e2e["ms"] if e2e else 0.0])
  fusion = 30
  try:
    fusion = int(scorer._ensemble().fusion())
  except Exception:
    pass
  if gsize > 1 and args.flow_sharding == "scenes" and scenes % gsize == 0:
    rows_sample = (scenes // gsize) * K          # every rank: the single-GPU flow stage on its scenes
    rows_score = (E - 1) * (scenes // gsize) * K
  elif gsize > 1:
    rows_sample = (scenes // gsize if scenes % gsize == 0 else scenes) * K
    rows_score = e_local * scenes * K
  else:
    rows_sample = scenes * K                 # the sampling pass also emits q[0]
    rows_score = (e_local - 1) * scenes * K
  work = rip_family_work(cfg, e_local, scenes, rows_sample, rows_score, fusion)
  return dict(cfg=cfg, value=total_scenes * K * steps / (elapsed_ms * 1e-3),
              ms_per_step=elapsed_ms / steps, stages=stages, clocks=clocks,
              launches=int(launches) * world, e2e=e2e,
              e2e_value=(total_scenes * K * steps / (e2e_ms * 1e-3)) if e2e else None,
              e2e_ms_per_step=(e2e_ms / steps) if e2e else None, prof=prof, psteps=psteps,
              work=work, equal=equal, scenes=scenes, total_scenes=total_scenes, e_local=e_local,
              gsize=gsize, n_groups=n_groups, steps=steps)


def gpu_eager_baseline(cfg, dev, scenes):
  """The number the hand-written kernels have to beat on the SAME GPU: the reference's own
  modules in stock PyTorch eager (cuDNN/cuBLAS/ATen, fp32, TF32 off) running the same step."""
  import torch
  from oracle import reference_arm as RA
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  if not RA.available():
    return {"unavailable": "no reference tree (baseline/_ref or /root/reference) on this box"}
  torch.backends.cuda.matmul.allow_tf32 = False
  torch.backends.cudnn.allow_tf32 = False
  E, K, T, C = cfg["E"], cfg["K"], cfg["T"], cfg["C"]
  sds = [synthetic_state_dict("dim", C, 100 + m) for m in range(E)]
  models = [m.to(dev) for m in RA.build_models(sds, T, C)]
  inp = {k: v.to(dev) for k, v in synthetic_inputs(scenes, C, K, T, G_GOALS, seed=0).items()}
  args = [inp[k] for k in ("lidar", "velocity", "is_at_traffic_light", "traffic_light_state", "x", "goal")]
  for _ in range(2):
    RA.rip_score(models, *args, 1.0, "WCM")
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  n = 3
  e0.record()
  for _ in range(n):
    out = RA.rip_score(models, *args, 1.0, "WCM")
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / n
  return {"value": scenes * K / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
          "what": "reference modules (ImitativeModel._params, AutoregressiveFlow._forward/_inverse) "
                  "in stock PyTorch eager on this GPU, fp32, TF32 off, same %d-scene step, %d timed "
                  "steps after 2 warm-ups" % (scenes, n), "kstar0": int(out["kstar"][0])}


def rip_line(D, r, args, peaks, warmup):
  cfg = r["cfg"]
  line = {
      "metric": METRIC if cfg is RIP_WORKLOADS["rip"] else rip_metric(cfg),
      "value": r["value"], "unit": "samples/s", "n_gpus": D.world,
      "steps": r["steps"], "warmup": max(warmup, 3), "ms_per_step": r["ms_per_step"],
      "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      "config": {
          "workload": cfg["name"], "ensemble": cfg["E"], "K": cfg["K"], "T": cfg["T"],
          "bev_channels": cfg["C"], "scenes_total": r["total_scenes"],
          "scenes_per_group": r["scenes"],
          "parallelism": "ensemble sharded %d model(s)/rank x %d replica group(s)%s" % (
              r["e_local"], r["n_groups"],
              "" if r["gsize"] == 1 else (
                  "; flow stage sharded by scenes behind ONE all-gather of the per-model latents z "
                  "(decoders replicated, 15 k parameters each)" if args.flow_sharding == "scenes" else
                  "; flow stage sharded by models: z_0 broadcast, proposal all-gather, all-gather of per-model scores")),
          "l2": "inputs larger than L2 (BEV grids %.0f MB + noise %.0f MB per group step)" %
                (r["scenes"] * cfg["C"] * 200 * 200 * 4 / 1e6, r["scenes"] * cfg["K"] * cfg["T"] * 8 / 1e6),
          "proposal_score": "q[0] is emitted by the sampling pass (bit-identical to a separate scoring "
                            "pass); flops counted = E passes",
          "launch": ("encoder stage replayed as one CUDA graph per input-buffer set"
                     if not args.no_cuda_graphs else "one launch per kernel"),
          "host_numa": ("each rank bound to the %s cores NVML lists as local to its GPU" % D.numa
                        if D.numa else "no binding"),
      },
      "stages_ms": r["stages"], "clocks": r["clocks"], "gpu_launches": r["launches"],
  }
  if r["equal"] is not None:
    line["sharded_equals_single"] = r["equal"]
  if r["e2e"]:
    e = r["e2e"]
    line["e2e"] = {"value": r["e2e_value"], "unit": "samples/s", "ms_per_step": r["e2e_ms_per_step"],
                   "h2d_bytes_per_step": e["h2d"] * D.world, "d2h_bytes_per_step": e["d2h"] * D.world,
                   "checksum": e["checksum"],
                   "pipeline": "streamed: pinned H2D of step i+1 (copy stream, double-buffered device "
                               "inputs) overlaps the kernels of step i; every step's plans are read back "
                               "and touched on the host; sharded ensembles: each rank uploads only its "
                               "1/R of the scenes, resizes them and all-gathers the 4x smaller features",
                   "blocking_call_ms": e["blocking_ms"]}
  rl = build_roofline(r["prof"], r["psteps"], r["work"], peaks, r["clocks"]) if r["prof"] else None
  if rl:
    rl["hbm_frac_of_step"] = ((r["scenes"] / r["gsize"]) * cfg["C"] * 200 * 200 * 4 +
                              r["scenes"] * cfg["K"] * (cfg["T"] * 16 + 4 * r["e_local"])) / \
                             (r["ms_per_step"] * 1e-3) / 1e9 / peaks["hbm_gbs"]
    line["roofline"] = rl
  return line


def run_gpu_rip(args):
  D = Dist(args)
  peaks = _peaks()
  cfg = RIP_WORKLOADS[args.workload]
  r = run_rip(D, cfg, args, args.steps, args.warmup)
  line = rip_line(D, r, args, peaks, args.warmup) if D.rank == 0 else None
  if args.workload == "rip" and D.world == 8 and not args.no_cfg5:
    # the 8-GPU configuration BASELINE.json names besides the metric's own (configs[4])
    r5 = run_rip(D, RIP_WORKLOADS["cfg5"], args, max(args.steps // 2, 5), max(args.warmup, 3))
    if D.rank == 0:
      l5 = rip_line(D, r5, args, peaks, args.warmup)
      line["cfg5"] = {k: l5[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "scaling", "config",
                                         "stages_ms", "gpu_launches", "sharded_equals_single", "e2e", "roofline")
                      if k in l5}
  if D.rank == 0:
    if D.world == 1 and not args.no_cpu_baseline:
      scenes = args.cpu_scenes or cfg["scenes"]
      c = cpu_rip_run(cfg, 2, 1, scenes)
      line["cpu_baseline"] = {
          "value": c["value"], "unit": "samples/s", "cores": c["cores"], "kind": c["kind"],
          "sample": "%d scenes x K=%d per step (same E/T/C%s), 2 timed steps after 1 warm-up, torch CPU "
                    "fp32 with %d threads; %.0f ms/step" % (scenes, cfg["K"], ", the full workload"
                                                            if scenes == cfg["scenes"] else "",
                                                            c["cores"], c["ms"])}
    if D.world == 1 and not args.no_eager_baseline:
      try:
        line["gpu_eager_baseline"] = gpu_eager_baseline(cfg, D.dev, cfg["scenes"])
      except Exception as e:
        line["gpu_eager_baseline"] = {"unavailable": repr(e)[:200]}
    print(json.dumps(line), flush=True)
  if D.world > 1:
    D.dist.destroy_process_group()


# ----------------------------------------------------------------------------------
# GPU arm: training steps (BASELINE configs[1] and [3])
# ----------------------------------------------------------------------------------
def run_gpu_train(args):
  import torch
  import oatomobile_b200 as ob
  from oatomobile_b200 import _native
  from oatomobile_b200.datasets import DeviceCollator
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  from oatomobile_b200.train import Trainer
  D = Dist(args)
  peaks = _peaks()
  tcfg = TRAIN_WORKLOADS[args.workload]
  kind, T, C = tcfg["kind"], tcfg["T"], tcfg["C"]
  world, rank, dev = D.world, D.rank, D.dev
  if tcfg["per_gpu"]:
    b_local, b_total = tcfg["batch"], tcfg["batch"] * world
  else:
    assert tcfg["batch"] % world == 0
    b_local, b_total = tcfg["batch"] // world, tcfg["batch"]
  group = D.groups(world) if world > 1 else None

  cls = ob.ImitativeModel if kind == "dim" else ob.BehaviouralModel
  model = cls(output_shape=(T, 2), in_channels=C)
  model.load_state_dict(synthetic_state_dict(kind, C, 400), strict=True)
  trainer = Trainer(model.to(dev), lr=1e-3, group=group, use_cuda_graphs=args.train_graphs)

  # synthetic episodes in the on-disk sample format (datasets/carla.py:238-325): HWC lidar,
  # 80-frame futures; the documented loop = collate -> model.transform -> train_step
  inp = synthetic_inputs(b_local, C, 1, T, seed=31 + rank)
  g = torch.Generator().manual_seed(77 + rank)
  future = torch.cumsum(torch.rand(b_local, 80, 3, generator=g) * 0.5, dim=1)
  samples = []
  for i in range(b_local):
    s = {"lidar": inp["lidar"][i].permute(1, 2, 0).contiguous().numpy(),
         "velocity": inp["velocity"][i].numpy(),
         "is_at_traffic_light": inp["is_at_traffic_light"][i].numpy(),
         "traffic_light_state": inp["traffic_light_state"][i].numpy(),
         "player_future": future[i].numpy()}
    if kind == "cil":
      s["mode"] = torch.randint(0, 4, (1,), generator=g).float().numpy()
    samples.append(s)
  collate = DeviceCollator(dev)
  resident = model.transform(collate(samples))
  resident = {k: v.clone() for k, v in resident.items()}

  warm = max(args.warmup, 3)
  for _ in range(warm):
    trainer.train_step(resident)
  D.barrier()
  sampler = ClockSampler(D.local_rank)
  sampler.start()
  l0 = _native.launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  D.barrier()
  ev0.record()
  for _ in range(args.steps):
    loss = trainer.train_step(resident)
  ev1.record()
  D.barrier()
  launches = _native.launch_count() - l0
  clocks = sampler.stop()
  elapsed_ms = ev0.elapsed_time(ev1)

  psteps = 2
  _native.profile_begin(dev)
  for _ in range(psteps):
    trainer.train_step(resident)
  prof = _native.profile_end()
  D.barrier()

  # e2e: host samples -> pinned staging -> H2D -> transform -> train_step -> loss read back
  h2d = sum(s[k].size * 4 for s in samples for k in s)
  for _ in range(2):
    float(trainer.train_step(model.transform(collate(samples))))
  D.barrier()
  t0 = time.perf_counter()
  checksum = 0.0
  for _ in range(args.steps):
    checksum += float(trainer.train_step(model.transform(collate(samples))))
  torch.cuda.synchronize()
  e2e_ms = 1e3 * (time.perf_counter() - t0)
  elapsed_ms, e2e_ms = D.max_over_ranks([elapsed_ms, e2e_ms])

  if rank == 0:
    ms = elapsed_ms / args.steps
    fam = {}
    top = None
    for name, p in prof.items():
      if not name:
        continue
      fam[name] = {"ms_per_step": p["ms"] / psteps, "launches_per_step": p["launches"] / psteps}
      if not name.startswith("(") and (top is None or fam[name]["ms_per_step"] > fam[top]["ms_per_step"]):
        top = name
    # algorithmic flops of one training step: forward + dX + dW of every conv = 3x forward
    fwd = sum(layer_flops(l) for l in encoder_layers(C))
    pw_fwd = sum(layer_flops(l) for l in encoder_layers(C) if l[0] in ("expand", "project", "last", "fc"))
    step_flop = 3.0 * fwd * b_local
    roofline = {
        "kernel": top, "bound": "tensor",
        "achieved": (3.0 * pw_fwd * b_local / (fam["simt_pw_gemm"]["ms_per_step"] * 1e-3) / 1e12)
                    if "simt_pw_gemm" in fam and top == "simt_pw_gemm" else step_flop / (ms * 1e-3) / 1e12,
        "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "traffic": None,
        "peak_source": peaks["source"] + ", bf16 dense sustained",
        "note": "the training step is FP32 SIMT (no tensor cores yet): achieved = algorithmic conv "
                "flops (forward + dX + dW = 3x forward) of the dominant family, or of the whole step "
                "when the dominant family is not a GEMM, over its live event time",
        "step_tflops": step_flop / (ms * 1e-3) / 1e12, "kernels": fam,
    }
    roofline["frac"] = roofline["achieved"] / roofline["peak"]
    line = {
        "metric": train_metric(tcfg), "value": b_total * args.steps / (elapsed_ms * 1e-3),
        "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak" if tcfg["per_gpu"] else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": tcfg["name"], "batch_total": b_total, "batch_per_gpu": b_local, "T": T,
                   "bev_channels": C,
                   "parallelism": "data parallel x%d, one NCCL all-reduce of the flat 9.7 MB gradient "
                                  "buffer per step" % world if world > 1 else "single GPU",
                   "l2": "activations of one step (>1 GB at B=64) exceed L2",
                   "launch": "forward+backward replayed as one CUDA graph" if args.train_graphs
                             else "one launch per kernel"},
        "clocks": clocks, "gpu_launches": int(launches) * world, "loss": float(loss),
        "e2e": {"value": b_total * args.steps / (e2e_ms * 1e-3), "unit": "samples/s",
                "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(h2d) * world,
                "d2h_bytes_per_step": 4 * world, "checksum": checksum,
                "pipeline": "per step: samples (on-disk format, HWC lidar) -> pinned staging -> async H2D "
                            "-> model.transform (fused HWC->CHW + resize kernel) -> train_step -> loss.item()"},
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
      c = cpu_train_run(tcfg, 2, 2, args.cpu_batch)
      line["cpu_baseline"] = {"value": c["value"], "unit": "samples/s", "cores": c["cores"], "kind": "port",
                              "sample": "%d-sample batch, 2 timed train steps through oracle/restatement.py "
                                        "(torch CPU autograd + Adam), %d threads; %.0f ms/step"
                                        % (c["batch"], c["cores"], c["ms"])}
    print(json.dumps(line), flush=True)
  if world > 1:
    D.dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--workload", default="rip", choices=sorted(RIP_WORKLOADS) + sorted(TRAIN_WORKLOADS))
  ap.add_argument("--cpu-scenes", type=int, default=0,
                  help="scenes per CPU step (0 = the workload's own scene count)")
  ap.add_argument("--cpu-batch", type=int, default=16, help="batch of the CPU training arm")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-eager-baseline", action="store_true")
  ap.add_argument("--no-cfg5", action="store_true", help="N=8: skip the extra BASELINE configs[4] measurement")
  ap.add_argument("--no-train-graphs", dest="train_graphs", action="store_false",
                  help="training workloads: one launch per kernel instead of replaying forward+backward as one "
                       "CUDA graph (Trainer(use_cuda_graphs=True))")
  ap.add_argument("--flow-sharding", default="scenes", choices=["scenes", "models"],
                  help="sharded ensembles: how the flow stage behind the model-sharded encoders is split "
                       "(oatomobile_b200/rip.py); 'models' = per-model scores all-gathered")
  ap.add_argument("--no-cuda-graphs", action="store_true",
                  help="launch the encoder's kernels one by one instead of replaying a CUDA graph")
  ap.add_argument("--e2e-chunks", type=int, default=1,
                  help="slices of the batch pipelined H2D-vs-compute in the e2e arm")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference_arm(args)
  elif args.workload in TRAIN_WORKLOADS:
    run_gpu_train(args)
  else:
    run_gpu_rip(args)


if __name__ == "__main__":
  main()
