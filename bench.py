#!/usr/bin/env python
"""bench.py — RIP trajectory samples scored/sec (ens=4, K=512, T=10) on B200.

One "step" = one pass of the hot path over one batch of synthetic scenes:
  transform (200x200 -> 100x100) -> E x encoder+merger -> proposals from model 0
  -> scores under all E models -> WCM aggregation -> argmin -> plan.
Workload at N=1: BASELINE.json configs[2] (B=256 scenes, E=4, K=512, T=10, C=4
BEV channels); weak scaling: 256 scenes per GPU, the ensemble sharded
E/min(N,E) models per rank, N/E replica groups beyond E ranks.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W    # CPU oracle port (reference arm)
  torchrun ... bench.py --gpus N ...                       # N > 1, one rank per GPU

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events on the launch
stream with inputs resident in HBM; `e2e` is the same metric through the public
host-buffer API (pinned H2D of every input + D2H of the plans inside the timed
region); `roofline` is for the dominant kernel pair (flow sample+score), its
duration measured live with CUDA events inside the timed steps; `cpu_baseline`
is the oracle (a PyTorch-CPU port of the reference path) on a bounded sample.
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

E_MODELS, K_SAMPLES, T_STEPS, C_BEV, B_PER_GPU, G_GOALS = 4, 512, 10, 4, 256, 10
FLOW_FLOP_PER_ROW_STEP = 29696          # 2*(64*192 + 2*192 + 64*32 + 32*4)
ENC_MFLOP_PER_IMAGE = 152.6             # C=4
METRIC = "RIP trajectory samples scored/sec (ens=4,K=512,T=10)"


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
# CPU arm: the oracle (PyTorch-CPU port of the reference path) on a bounded sample.
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
  """PyTorch-CPU throughput on this path collapses when the intra-op pool is larger
  than what the small GRU/conv ops can use (measured: 128 threads = 400x slower than
  8).  Try ascending pool sizes on one tiny call each and keep the fastest."""
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


def cpu_reference_run(steps, warmup, scenes):
  import torch
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  from oracle import restatement as R  # the CPU baseline IS the oracle
  inp = synthetic_inputs(scenes, C_BEV, K_SAMPLES, T_STEPS, G_GOALS, seed=0)
  sds = [synthetic_state_dict("dim", C_BEV, 100 + m) for m in range(E_MODELS)]

  def tiny():
    with torch.no_grad():
      R.rip_score_from_inputs(sds[:1], inp["lidar"][:1], inp["velocity"][:1],
                              inp["is_at_traffic_light"][:1], inp["traffic_light_state"][:1],
                              inp["x"][:1, :64], inp["goal"][:1], 1.0, "WCM")

  cores = pick_cpu_threads(tiny)
  times = []
  with torch.no_grad():
    for i in range(warmup + steps):
      t0 = time.perf_counter()
      R.rip_score_from_inputs(sds, inp["lidar"], inp["velocity"], inp["is_at_traffic_light"],
                              inp["traffic_light_state"], inp["x"], inp["goal"], 1.0, "WCM")
      dt = time.perf_counter() - t0
      if i >= warmup:
        times.append(dt)
  total = sum(times)
  value = scenes * K_SAMPLES * len(times) / total
  return value, 1e3 * total / len(times), cores


def run_reference_arm(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return  # only rank 0 runs the CPU arm; the others exit 0 without work
  scenes = args.cpu_scenes
  value, ms, cores = cpu_reference_run(args.steps, max(args.warmup, 1), scenes)
  sample = ("%d scenes x K=%d samples per step (E=%d, T=%d, C=%d), torch CPU fp32, %d threads"
            % (scenes, K_SAMPLES, E_MODELS, T_STEPS, C_BEV, cores))
  line = {
      "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
      "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
      "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f32", "data": "synthetic",
      "config": {"workload": "RIP WCM sample-and-score, E=4, K=512, T=10, C=4 BEV 200x200, "
                             "bounded sample of %d scenes per step" % scenes,
                 "ensemble": E_MODELS, "K": K_SAMPLES, "T": T_STEPS, "scenes": scenes},
      "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                       "sample": sample},
      "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------
def run_gpu_arm(args):
  import torch
  import torch.distributed as dist
  import oatomobile_b200 as ob
  from oatomobile_b200 import _native
  from oatomobile_b200.rip import HostRIPPipeline, RIPScorer
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if world != args.gpus:
    if world == 1 and args.gpus > 1:
      raise SystemExit("launch N>1 with torchrun (one rank per GPU)")
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)

  # ---- ensemble sharding: contiguous blocks of E/R models, N/E replica groups ----
  gsize = min(world, E_MODELS)            # ranks sharing one ensemble
  assert E_MODELS % gsize == 0 and world % gsize == 0
  e_local = E_MODELS // gsize
  group_id, grank = rank // gsize, rank % gsize
  group = None
  if world > 1:
    for g in range(world // gsize):
      pg = dist.new_group(list(range(g * gsize, (g + 1) * gsize)))
      if g == group_id:
        group = pg
  scenes = B_PER_GPU * gsize              # scenes scored by this replica group per step
  total_scenes = B_PER_GPU * world

  sds = {m: synthetic_state_dict("dim", C_BEV, 100 + m) for m in
         set(range(grank * e_local, (grank + 1) * e_local)) | {0}}

  def make(m):
    model = ob.ImitativeModel(output_shape=(T_STEPS, 2), in_channels=C_BEV)
    model.load_state_dict(sds[m], strict=True)
    return model.to(dev).eval()

  models = [make(m) for m in range(grank * e_local, (grank + 1) * e_local)]
  proposal = None if grank == 0 else make(0)
  scorer = RIPScorer(models, "WCM", group=group if gsize > 1 else None, proposal_model=proposal,
                     use_cuda_graphs=not args.no_cuda_graphs)

  inp = synthetic_inputs(scenes, C_BEV, K_SAMPLES, T_STEPS, G_GOALS, seed=group_id)
  host = {k: v.pin_memory() for k, v in inp.items()}
  d = {k: v.to(dev) for k, v in inp.items()}
  x, goal = d.pop("x"), d.pop("goal")

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def step():
    return scorer(x=x, goal=goal, epsilon=1.0, **d)

  for _ in range(max(args.warmup, 3)):
    step()
  barrier()

  # ---- timed region: device-resident inputs (164 MB of BEV grids > 126 MB L2) ----
  sampler = ClockSampler(local_rank)
  sampler.start()
  scorer.stage_events = []
  launches0 = _native.launch_count() + scorer.replayed_launches
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  ev0.record()
  for _ in range(args.steps):
    step()
  ev1.record()
  barrier()
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

  # ---- e2e: host (pinned) buffers through the public pipeline --------------------
  pipe = HostRIPPipeline(scorer, dev, chunks=args.e2e_chunks)
  for _ in pipe.stream(host for _ in range(3)):
    pass
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  t_host0 = time.perf_counter()
  checksum = 0.0
  for res in pipe.stream(host for _ in range(args.steps)):
    checksum += float(res["plan"][0, 0, 0])  # the host really reads every step's result
  e1.record()
  barrier()
  e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t_host0))
  # latency form: one blocking call per step (no cross-step overlap)
  l0 = time.perf_counter()
  for _ in range(max(args.steps // 4, 3)):
    res = pipe(host)
  e2e_blocking_ms = 1e3 * (time.perf_counter() - l0) / max(args.steps // 4, 3)

  # ---- max over ranks ------------------------------------------------------------
  t = torch.tensor([elapsed_ms, e2e_ms], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  elapsed_ms, e2e_ms = float(t[0]), float(t[1])
  ms_per_step = elapsed_ms / args.steps
  value = total_scenes * K_SAMPLES * args.steps / (elapsed_ms * 1e-3)
  e2e_value = total_scenes * K_SAMPLES * args.steps / (e2e_ms * 1e-3)

  if rank == 0:
    peaks = _peaks()
    # dominant kernel pair: flow sample + score launches on this rank
    rows = scenes * K_SAMPLES
    if gsize > 1 and scenes % gsize == 0:
      passes = e_local + 1.0 / gsize  # proposals decoded for 1/R of the scenes, then E_local scoring passes
    else:
      passes = e_local if grank == 0 else e_local + 1  # rank 0 scores model 0 while sampling
    flow_flop = passes * rows * T_STEPS * FLOW_FLOP_PER_ROW_STEP
    flow_tflops = flow_flop / (stages["flow"] * 1e-3) / 1e12 if stages["flow"] > 0 else 0.0
    sm_mhz = clocks.get("sm_mhz") or 0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12 if sm_mhz else None
    tensor_peak = peaks["bf16_tflops_sustained"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):  # dram bytes/launch from the last `ncu --set full` capture
      traffic = json.load(open(tpath)).get("flow_tc_kernel_pair_bytes")
    roofline = {
        "kernel": "oat::flow_tc2_kernel<0> + <1> (sample + score launches of one step)",
        "bound": "tensor", "achieved": flow_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
        "frac": flow_tflops / tensor_peak, "traffic": traffic,
        "peak_source": peaks["source"] + ", bf16 dense sustained",
        "pipe": "tcgen05.mma kind::tf32, 3xTF32 error compensation (3 TF32 MMAs per algorithmic "
                "MMA, TF32 = 1/2 bf16 rate -> ceiling of this formulation = peak/6), co-limited "
                "by the FP32/MUFU gate math of the GRU",
        "frac_of_3xtf32_ceiling": flow_tflops / (tensor_peak / 6.0),
        "fp32_simt_peak": fp32_peak,
        "x_fp32_simt_peak": (flow_tflops / fp32_peak) if fp32_peak else None,
        "algorithmic_flop_per_launch_pair": flow_flop,
        "algorithmic_bytes_per_launch_pair": rows * (T_STEPS * 16 + 4 * passes),
        "ms_per_launch_pair": stages["flow"],
        "encoder": {"ms": stages["encode"],
                    "tflops": e_local * scenes * ENC_MFLOP_PER_IMAGE * 1e6 /
                              (stages["encode"] * 1e-3) / 1e12 if stages["encode"] > 0 else 0.0},
        "hbm_frac_of_step": (scenes * (C_BEV * 200 * 200 * 4 + K_SAMPLES * (T_STEPS * 16 + 4 * e_local)))
                            / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
    }
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[2]: RIPAgent WCM sample-and-score, 4-model ensemble, "
                        "K=512, T=10, %d scenes/GPU of 200x200x4 BEV grids" % B_PER_GPU,
            "ensemble": E_MODELS, "K": K_SAMPLES, "T": T_STEPS, "bev_channels": C_BEV,
            "scenes_total": total_scenes, "scenes_per_gpu": B_PER_GPU,
            "parallelism": "ensemble sharded %d model(s)/rank x %d replica group(s)" %
                           (e_local, world // gsize),
            "l2": "inputs larger than L2 (BEV grids %.0f MB + noise %.0f MB per step)" %
                  (scenes * C_BEV * 200 * 200 * 4 / 1e6, scenes * K_SAMPLES * T_STEPS * 8 / 1e6),
            "proposal_score": "q[0] is emitted by the sampling pass (bit-identical to a separate "
                              "scoring pass); flops counted = E passes",
            "launch": ("encoder stage replayed as one CUDA graph per input-buffer set"
                       if not args.no_cuda_graphs else "one launch per kernel"),
        },
        "stages_ms": stages, "clocks": clocks, "gpu_launches": int(launches) * world,
        "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(pipe.h2d_bytes) * world,
                "d2h_bytes_per_step": int(pipe.d2h_bytes) * world, "checksum": checksum,
                "pipeline": "streamed: pinned H2D of step i+1 (copy stream, double-buffered device "
                            "inputs) overlaps the kernels of step i; every step's plans are read "
                            "back and touched on the host",
                "blocking_call_ms": e2e_blocking_ms},
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
      v, ms, cores = cpu_reference_run(3, 1, args.cpu_scenes)
      line["cpu_baseline"] = {
          "value": v, "unit": "samples/s", "cores": cores, "kind": "port",
          "sample": "%d scenes x K=%d per step (same E/T/C), 3 timed steps after 1 warm-up, "
                    "torch CPU fp32 with %d threads; %.0f ms/step" % (args.cpu_scenes, K_SAMPLES,
                                                                     cores, ms)}
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--cpu-scenes", type=int, default=16,
                  help="scenes per CPU-baseline step (bounded sample of the workload)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-cuda-graphs", action="store_true",
                  help="launch the encoder's kernels one by one instead of replaying a CUDA graph")
  ap.add_argument("--e2e-chunks", type=int, default=1,
                  help="slices of the batch pipelined H2D-vs-compute in the e2e arm")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference_arm(args)
  else:
    run_gpu_arm(args)


if __name__ == "__main__":
  main()
