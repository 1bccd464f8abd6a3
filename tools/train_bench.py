"""Training-step throughput (BASELINE configs cfg2 = DIM B64 and
cfg4 = CIL B512 over 8 GPUs = 64 per GPU), one GPU, next to the CPU oracle.

  python tools/train_bench.py [--steps 20] [--warmup 5] [--batch 64] [--no-cpu]

One step = H2D of nothing (inputs resident), training-mode forward + backward + Adam, i.e.
`Trainer.forward_backward` + `Trainer.optimizer_step` (the DIM target noise and the dropout
mask are drawn by the torch RNG inside the timed region, as in `train_step`).  Timed with
CUDA events on the launching stream.  The CPU leg runs the oracle restatement
(torch autograd, float32) + Adam on the host cores with the best thread count.
Prints one JSON line per model kind."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--batch", type=int, default=64)
  ap.add_argument("--no-cpu", action="store_true")
  ap.add_argument("--graphs", action="store_true", help="Trainer(use_cuda_graphs=True) (experimental)")
  a = ap.parse_args()
  import oatomobile_b200 as ob
  from oatomobile_b200 import _native as N
  from oatomobile_b200.synthetic import synthetic_state_dict
  from oatomobile_b200.train import Trainer
  from bench import pick_cpu_threads
  from tests.helpers import train_inputs
  dev = torch.device("cuda", 0)
  for kind in ("dim", "cil"):
    cfg = dict(kind=kind, T=4, C=2, B=a.batch, wseed=400, iseed=21)
    visual, scalars, target = train_inputs(cfg)
    sd = synthetic_state_dict(kind, 2, 400)
    cls = ob.ImitativeModel if kind == "dim" else ob.BehaviouralModel
    model = cls(output_shape=(4, 2), in_channels=2)
    model.load_state_dict(sd)
    trainer = Trainer(model.to(dev), lr=1e-3, use_cuda_graphs=a.graphs)
    batch = dict(visual_features=visual.to(dev), velocity=scalars[:, 0:3].to(dev),
                 is_at_traffic_light=scalars[:, 3:4].to(dev), traffic_light_state=scalars[:, 4:5].to(dev),
                 player_future=torch.cat([target, torch.zeros(a.batch, 4, 1)], -1).to(dev))
    if kind == "cil":
      batch["mode"] = scalars[:, 5:6].to(dev)
    losses = []
    for _ in range(a.warmup):
      losses.append(trainer.train_step(batch).item())
    l0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
      loss = trainer.train_step(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    line = {"metric": "training samples per second (%s train_step, B=%d, T=4, C=2)" % (kind.upper(), a.batch),
            "value": a.batch / ms * 1e3, "unit": "samples/s", "ms_per_step": ms, "steps": a.steps,
            "warmup": a.warmup, "n_gpus": 1, "dtype": "f32", "data": "synthetic",
            "gpu_launches": int(N.launch_count() - l0), "loss_first": losses[0] if losses else None,
            "loss_last": float(loss.item())}
    if not a.no_cpu:
      from oracle import restatement as R
      state = {k: v.clone() for k, v in sd.items()}
      moments = {}

      def cpu_step():
        loss, grads, bufs, _ = R.train_forward_backward(state, kind, visual, scalars, target)
        for k, g in grads.items():
          m, v = moments.get(k, (torch.zeros_like(g), torch.zeros_like(g)))
          state[k], m, v = R.adam_update(state[k], g, m, v, 1)
          moments[k] = (m, v)
        state.update(bufs)

      threads = pick_cpu_threads(cpu_step)
      t0 = time.perf_counter()
      n = 3
      for _ in range(n):
        cpu_step()
      cpu_ms = (time.perf_counter() - t0) / n * 1e3
      line["cpu_baseline"] = {"value": a.batch / cpu_ms * 1e3, "unit": "samples/s", "ms_per_step": cpu_ms,
                              "cores": threads, "kind": "port",
                              "sample": "%d steps of the same batch through oracle/restatement.py (torch CPU autograd)" % n}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
  main()
