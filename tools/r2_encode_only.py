"""Runs the encoder stage alone (E=4, B=256, C=4) a few times — the target of ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oatomobile_b200 as ob
from oatomobile_b200 import _native as N, ops
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict

def main():
  reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
  E, B, C = 4, 256, 4
  dev = "cuda:0"
  models = []
  for m in range(E):
    model = ob.ImitativeModel(output_shape=(10, 2), in_channels=C)
    model.load_state_dict(synthetic_state_dict("dim", C, 100 + m), strict=True)
    models.append(model.to(dev).eval())
  ens = N.EnsembleHandle([m.native_handle() for m in models])
  inp = synthetic_inputs(B, C, 1, 10, seed=0)
  vis = ops.transform_visual(inp["lidar"].to(dev))
  scal = torch.cat([inp["velocity"], inp["is_at_traffic_light"], inp["traffic_light_state"]], 1).to(dev)
  for _ in range(reps):
    z = ops.encode(ens, vis, scal)
  torch.cuda.synchronize()
  print("ok", float(z.abs().sum()))

main()
