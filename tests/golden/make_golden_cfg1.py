"""Generates tests/golden/cfg1_forward_B4_C4.npz: BASELINE.json configs[0] — the reference's own
CPU-runnable case, `ImitativeModel.forward` on a batch of 4 synthetic 200x200x4 BEV grids — by
running the REAL reference (`oatomobile/baselines/torch/dim/model.py:76-141`, 10 Adam steps, with
and without goals) on the CPU through oracle/ref_shim.py.  Run in the build container only."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict  # noqa: E402

CFG1 = dict(T=4, C=4, B=4, wseed=610, iseed=61)


def main():
  torch.set_num_threads(8)
  cfg = CFG1
  inp = synthetic_inputs(cfg["B"], cfg["C"], 1, cfg["T"], seed=cfg["iseed"])
  model = ref_shim.make_imitative_model(T=cfg["T"], in_channels=cfg["C"], seed=0, randomize_bn=False)
  model.load_state_dict(synthetic_state_dict("dim", cfg["C"], cfg["wseed"]), strict=True)
  model.eval()
  for p in model.parameters():
    p.requires_grad_(False)
  out = {}
  with torch.no_grad():
    vis = model.transform({"lidar": inp["lidar"].clone()})["visual_features"].contiguous()
  ctx = dict(visual_features=vis, velocity=inp["velocity"], is_at_traffic_light=inp["is_at_traffic_light"],
             traffic_light_state=inp["traffic_light_state"])
  with torch.no_grad():
    out["z"] = model._params(**ctx).numpy()
  torch.manual_seed(777)
  out["x0"] = model._decoder._base_dist.sample().view(1, cfg["T"], 2).numpy()  # dim/model.py:100-105
  torch.manual_seed(777)
  out["plan_goal"] = model.forward(num_steps=10, goal=inp["goal"], lr=1e-1, epsilon=1.0, **ctx).detach().numpy()
  torch.manual_seed(777)
  out["plan_nogoal"] = model.forward(num_steps=10, goal=None, lr=1e-1, epsilon=1.0, **ctx).detach().numpy()
  np.savez_compressed(os.path.join(HERE, "cfg1_forward_B4_C4.npz"), **out)
  print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
  main()
