"""Runs the flow sample+score stage alone (E=4, B=256, K=512, T=10) — target of ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oatomobile_b200 as ob
from oatomobile_b200 import _native as N, ops
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict

def main():
  reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
  E, B, C, K, T = 4, 256, 4, 512, 10
  dev = "cuda:0"
  models = []
  for m in range(E):
    model = ob.ImitativeModel(output_shape=(T, 2), in_channels=C)
    model.load_state_dict(synthetic_state_dict("dim", C, 100 + m), strict=True)
    models.append(model.to(dev).eval())
  ens = N.EnsembleHandle([m.native_handle() for m in models])
  inp = synthetic_inputs(B, C, K, T, seed=0)
  g = torch.Generator().manual_seed(3)
  z = (torch.randn(E, B, 64, generator=g) * 0.4).clamp(min=0).to(dev)
  x, goal = inp["x"].to(dev), inp["goal"].to(dev)
  for _ in range(reps):
    y, q = ops.rip_sample_score(ens, z, x, goal, 1.0, proposal_idx=0)
  torch.cuda.synchronize()
  print("ok", float(q.sum()))

main()
