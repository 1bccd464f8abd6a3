"""Tiny fused-front run for `compute-sanitizer --tool memcheck|racecheck` (GPU box helper):
features.0-4 of one model on one scene with every fused kernel switched on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oatomobile_b200 as ob
from oatomobile_b200 import _native as N, ops
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict

dev = "cuda:0"
C = 4
m = ob.ImitativeModel(output_shape=(4, 2), in_channels=C)
m.load_state_dict(synthetic_state_dict("dim", C, 50))
m = m.to(dev).eval()
ens = N.EnsembleHandle([m.native_handle()])
ens.set_fusion(15)
ens.set_fusion_tc(2)
vis = ops.transform_visual(synthetic_inputs(1, C, 1, 4, seed=3)["lidar"].to(dev))
out = ops.encoder_prefix(ens, vis, 4)
torch.cuda.synchronize()
print("ok fused prefix", float(out.sum()))
