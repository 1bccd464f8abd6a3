"""Times encoder and flow at the bench sizes with the library selected by OAT_B200_LIB and
reports parity of z against the FP32 SIMT path (GPU box helper for A/B-ing kernel variants)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oatomobile_b200 as ob
from oatomobile_b200 import _native, ops
from oatomobile_b200.rip import RIPScorer
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict

dev = "cuda:0"
B, C, E, K, T = 256, 4, 4, 512, 10
inp = synthetic_inputs(B, C, K, T, seed=0)
models = []
for m in range(E):
  mod = ob.ImitativeModel(output_shape=(T, 2), in_channels=C)
  mod.load_state_dict(synthetic_state_dict("dim", C, 100 + m))
  models.append(mod.to(dev).eval())
d = {k: v.to(dev) for k, v in inp.items()}
x, goal = d.pop("x"), d.pop("goal")
vis = ops.transform_visual(d.pop("lidar"))
ctx = dict(visual_features=vis, **d)

def timeit(fn, n=10):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n

if os.environ.get("OAT_FLOW_IMPL"):
  _native.set_flow_impl(os.environ["OAT_FLOW_IMPL"])
sc = RIPScorer(models, "WCM")
z = sc.encode(**ctx)
t_enc = timeit(lambda: sc.encode(**ctx))
t_flow = timeit(lambda: ops.rip_sample_score(sc._ensemble(), z, x, goal, 1.0))
_native.set_default_pw_impl("simt")
sc2 = RIPScorer(models, "WCM")
z_ref = sc2.encode(**ctx)
rel = ((z - z_ref).abs() / torch.clamp(torch.maximum(z.abs(), z_ref.abs()), min=1.0)).max().item()
y1, q1 = ops.rip_sample_score(sc._ensemble(), z, x, goal, 1.0)
_native.set_flow_impl("simt")
y0, q0 = ops.rip_sample_score(sc._ensemble(), z, x, goal, 1.0)
relq = ((q1 - q0).abs() / torch.clamp(torch.maximum(q1.abs(), q0.abs()), min=1.0)).max().item()
print("VARIANT %s flow=%s  encode %.3f ms  flow %.3f ms  z-vs-simt %.2e  q-vs-simt %.2e" %
      (os.path.basename(os.environ.get("OAT_B200_LIB", "default")),
       os.environ.get("OAT_FLOW_IMPL", "default"), t_enc, t_flow, rel, relq))
