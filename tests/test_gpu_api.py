"""GPU tests of the stand-alone L2 entry points the reference exports
(oatomobile/torch/networks/__init__.py:17-19): `MobileNetV2.forward`, `MLP.forward`,
`AutoregressiveFlow.forward(z)`, and of the BASELINE configs[4]
shape (E=8, K=2048) against the oracle."""
import pytest
import torch

from oracle import restatement as R
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
from tests.helpers import REL_TOL, assert_close, top2_gap

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("C", [2, 4])
def test_mobilenet_v2_forward_alone(C):
  """perception.py:53-55 — the encoder as its own module (keys `_model.features...`)."""
  from oatomobile_b200.networks import MobileNetV2
  sd = synthetic_state_dict("dim", C, 610 + C)
  enc_sd = {k[len("_encoder."):]: v for k, v in sd.items() if k.startswith("_encoder.")}
  net = MobileNetV2(num_classes=128, in_channels=C)
  net.load_state_dict(enc_sd, strict=True)
  net = net.to(DEV).eval()
  inp = synthetic_inputs(5, C, 1, 4, seed=33)
  with torch.no_grad():
    vis = R.transform_visual(inp["lidar"])
    want = R.mobilenet_v2_encode(sd, vis)
  got = net(vis.to(DEV))
  assert tuple(got.shape) == (5, 128)
  assert_close(got, want, REL_TOL, "MobileNetV2.forward")
  # the same features feed `_params`: encoder inside the model == encoder alone
  import oatomobile_b200 as ob
  model = ob.ImitativeModel(output_shape=(4, 2), in_channels=C)
  model.load_state_dict(sd, strict=True)
  model = model.to(DEV).eval()
  assert torch.equal(model._encoder(vis.to(DEV)), got)
  with pytest.raises(ValueError):
    net(torch.zeros(1, C, 64, 64, device=DEV))


def test_mlp_forward_alone():
  """mlp.py:25-72 — merger-shaped and head-shaped stacks against torch on the CPU."""
  from oatomobile_b200.networks import MLP
  torch.manual_seed(3)
  for sizes, final in (((133, [64, 64, 64]), True), ((64, [32, 4]), False), ((7, [5, 3]), False),
                       ((300, [1000, 17, 260]), True)):
    mlp = MLP(input_size=sizes[0], output_sizes=sizes[1], activate_final=final)
    x = torch.randn(9, sizes[0])
    with torch.no_grad():
      want = torch.nn.Sequential(*[m for m in mlp._model])(x.clone())
    got = mlp.to(DEV)(x.to(DEV))
    assert_close(got, want, 1e-5, "MLP.forward %s" % (sizes,))
  # leading dimensions are kept, like nn.Linear
  mlp = MLP(input_size=6, output_sizes=[8, 3]).to(DEV)
  assert tuple(mlp(torch.randn(2, 5, 6, device=DEV)).shape) == (2, 5, 3)
  from oatomobile_b200._native import NativeLibraryError
  with pytest.raises(NativeLibraryError):
    MLP(input_size=4, output_sizes=[4, 4], dropout_rate=0.5).to(DEV)(torch.zeros(1, 4, device=DEV))


@pytest.mark.parametrize("T", [4, 10])
def test_autoregressive_flow_forward_samples(T):
  """sequence.py:76-93 — `forward(z)` draws x ~ N(0, I) on the device and pushes it through
  `_forward`: shape, finiteness, and the draw is recovered by `_inverse` (round trip)."""
  from oatomobile_b200.networks import AutoregressiveFlow
  sd = synthetic_state_dict("dim", 2, 77)
  flow = AutoregressiveFlow(output_shape=(T, 2))
  flow.load_state_dict({k[len("_decoder."):]: v for k, v in sd.items() if k.startswith("_decoder.")},
                       strict=True)
  flow = flow.to(DEV)
  z = (torch.randn(300, 64, generator=torch.Generator().manual_seed(1)) * 0.4).clamp(min=0).to(DEV)
  torch.manual_seed(123)
  y = flow(z)
  assert tuple(y.shape) == (300, T, 2) and bool(torch.isfinite(y).all())
  torch.manual_seed(123)  # the same device draw, pushed through `_forward` explicitly
  x = torch.randn(300, T, 2, device=DEV)
  y2, _ = flow._forward(x, z)
  assert torch.equal(y, y2)
  xr, log_prob, logabsdet = flow._inverse(y, z)
  assert float((xr - x).abs().max()) < 1e-3
  want_lp = -0.5 * (x.double() ** 2).sum((1, 2)) - T * 1.8378770664093453
  assert_close(log_prob, want_lp, 1e-4, "forward(z) round-trip log_prob")
  torch.manual_seed(124)
  assert not torch.equal(flow(z), y)  # a fresh draw every call


def test_cfg5_shape_e8_k2048_against_oracle():
  """BASELINE configs[4]: 8 models, K=2048, T=10 (the flow/score path at that K and E)."""
  import oatomobile_b200 as ob
  from oatomobile_b200.rip import RIPScorer
  E, B, K, T, C = 8, 3, 2048, 10, 4
  inp = synthetic_inputs(B, C, K, T, seed=40)
  sds = [synthetic_state_dict("dim", C, 100 + m) for m in range(E)]
  models = []
  for sd in sds:
    m = ob.ImitativeModel(output_shape=(T, 2), in_channels=C)
    m.load_state_dict(sd, strict=True)
    models.append(m.to(DEV).eval())
  d = {k: v.to(DEV) for k, v in inp.items()}
  x, goal = d.pop("x"), d.pop("goal")
  for algo in ("WCM", "MA"):
    out = RIPScorer(models, algo)(x=x, goal=goal, want_s=True, **d)
    with torch.no_grad():
      ref = R.rip_score_from_inputs(sds, inp["lidar"], inp["velocity"], inp["is_at_traffic_light"],
                                    inp["traffic_light_state"], inp["x"], inp["goal"], 1.0, algo)
    assert tuple(out["q"].shape) == (E, B, K)
    for k in ("z", "y", "q", "s"):
      assert_close(out[k], ref[k], REL_TOL, "cfg5 %s %s" % (k, algo))
    assert torch.equal(out["s"].cpu(), R.rip_aggregate(out["q"].cpu(), algo))  # bit-exact on its own q
    assert torch.equal(out["kstar"].cpu().long(), torch.argmin(out["s"].cpu(), dim=1))
    gap = top2_gap(ref["s"])
    for b in range(B):
      kg, kr = int(out["kstar"][b]), int(ref["kstar"][b])
      assert float(ref["s"][b, kg] - ref["s"][b, kr]) <= 2 * REL_TOL * max(1.0, abs(float(ref["s"][b, kr])))
      if gap[b] > 2 * REL_TOL:
        assert kg == kr
