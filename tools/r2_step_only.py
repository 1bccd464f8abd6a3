"""One full RIP step (transform -> encode -> flow -> aggregate; E=4, B=256, K=512, T=10, C=4), run
`reps` times without CUDA graphs — the target of the ncu launch-list / traffic captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oatomobile_b200 as ob
from oatomobile_b200.rip import RIPScorer
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
  scorer = RIPScorer(models, "WCM", use_cuda_graphs=False)
  inp = {k: v.to(dev) for k, v in synthetic_inputs(B, C, K, T, seed=0).items()}
  x, goal = inp.pop("x"), inp.pop("goal")
  for _ in range(reps):
    out = scorer(x=x, goal=goal, epsilon=1.0, **inp)
  torch.cuda.synchronize()
  print("ok", out["kstar"][:4].tolist())

main()
