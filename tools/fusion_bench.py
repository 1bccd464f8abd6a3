"""GPU box helper: times the encoder at the bench size for several fusion masks (see
`oat_ensemble_set_fusion`) in one process and reports parity of z against the unfused path.
`--once MASK` runs a single encode with that mask (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oatomobile_b200 as ob
from oatomobile_b200 import ops
from oatomobile_b200.rip import RIPScorer
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict

dev = "cuda:0"
B, C, E, K, T = 256, 4, 4, 1, 10
inp = synthetic_inputs(B, C, K, T, seed=0)
models = []
for m in range(E):
  mod = ob.ImitativeModel(output_shape=(T, 2), in_channels=C)
  mod.load_state_dict(synthetic_state_dict("dim", C, 100 + m))
  models.append(mod.to(dev).eval())
d = {k: v.to(dev) for k, v in inp.items()}
d.pop("x"), d.pop("goal")
vis = ops.transform_visual(d.pop("lidar"))
ctx = dict(visual_features=vis, **d)
sc = RIPScorer(models, "WCM")
ens = sc._ensemble()


def timeit(fn, n=10):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n


if "--once" in sys.argv:
  ens.set_fusion(int(sys.argv[sys.argv.index("--once") + 1]))
  sc.encode(**ctx)
  torch.cuda.synchronize()
  sys.exit(0)

ens.set_fusion(0)
z0 = sc.encode(**ctx)
for mask in (0, 1, 2, 4, 8, 14, 16, 30, 0):
  ens.set_fusion(mask)
  z = sc.encode(**ctx)
  rel = ((z - z0).abs() / torch.clamp(torch.maximum(z.abs(), z0.abs()), min=1.0)).max().item()
  t = timeit(lambda: sc.encode(**ctx))
  print("FUSION mask %2d splits %s: encode %.3f ms  z-vs-unfused %.2e" %
        (mask, os.environ.get("OAT_FUSED_SPLITS", "default"), t, rel), flush=True)
