"""Latency of the fused gradient planner (ImitativeModel.forward / RIPAgent body) on the GPU
box, next to the CPU oracle's autograd planner on the same inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oatomobile_b200 as ob
from oatomobile_b200 import ops
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
from oracle import restatement as R  # checker / CPU baseline only

dev = "cuda:0"
for (B, T, E, C, steps) in [(1, 4, 4, 2, 10), (4, 4, 1, 2, 10), (4, 10, 4, 4, 10), (64, 4, 4, 2, 10)]:
  inp = synthetic_inputs(B, C, 1, T, seed=1)
  sds = [synthetic_state_dict("dim", C, 10 + m) for m in range(E)]
  models = []
  for sd in sds:
    m = ob.ImitativeModel(output_shape=(T, 2), in_channels=C)
    m.load_state_dict(sd)
    models.append(m.to(dev).eval())
  obs = models[0].transform({"lidar": inp["lidar"].to(dev)})
  ctx = dict(visual_features=obs["visual_features"], velocity=inp["velocity"].to(dev),
             is_at_traffic_light=inp["is_at_traffic_light"].to(dev),
             traffic_light_state=inp["traffic_light_state"].to(dev))
  zs = torch.stack([m._params(**ctx) for m in models])
  x0 = torch.zeros(B, T, 2, device=dev)
  goal = inp["goal"].to(dev)
  algo = None if E == 1 else "WCM"
  hs = [m.native_handle() for m in models]
  f = lambda: ops.plan(hs, zs, x0, steps, 0.1, goal, 1.0, algo)
  for _ in range(3): f()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(20): f()
  torch.cuda.synchronize()
  gpu_ms = (time.perf_counter() - t0) / 20 * 1e3
  # full agent-style call: encoders + planner
  def full():
    z = torch.stack([m._params(**ctx) for m in models])
    return ops.plan(hs, z, x0, steps, 0.1, goal, 1.0, algo)
  for _ in range(3): full()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(20): full()
  torch.cuda.synchronize()
  full_ms = (time.perf_counter() - t0) / 20 * 1e3
  torch.set_num_threads(8)
  zc = [z.cpu() for z in zs]
  t0 = time.perf_counter()
  R.planner(sds, zc, x0.cpu(), steps, 0.1, inp["goal"], 1.0, algo)
  cpu_ms = (time.perf_counter() - t0) * 1e3
  print("PLANNER B=%d T=%d E=%d steps=%d: fused kernel %.3f ms | encoders+planner %.3f ms | "
        "CPU oracle planner only %.1f ms" % (B, T, E, steps, gpu_ms, full_ms, cpu_ms))
