"""Small end-to-end step for `compute-sanitizer --tool memcheck` (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oatomobile_b200 as ob
from oatomobile_b200 import _native
from oatomobile_b200.rip import RIPScorer
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict

which = sys.argv[1] if len(sys.argv) > 1 else "tcgen05x2"  # simt | tcgen05 | tcgen05x2 (library default)
_native.set_flow_impl(which)
_native.set_default_pw_impl("simt" if which == "simt" else "tcgen05")
dev = "cuda:0"
B, C, E, K, T = 3, 4, 2, 96, 10
inp = synthetic_inputs(B, C, K, T, seed=3)
models = []
for m in range(E):
  mod = ob.ImitativeModel(output_shape=(T, 2), in_channels=C)
  mod.load_state_dict(synthetic_state_dict("dim", C, 50 + m))
  models.append(mod.to(dev).eval())
scorer = RIPScorer(models, "WCM")
d = {k: v.to(dev) for k, v in inp.items()}
x, goal = d.pop("x"), d.pop("goal")
out = scorer(x=x, goal=goal, **d)
torch.cuda.synchronize()
print("ok", which, out["kstar"].tolist(), float(out["q"].sum()))
