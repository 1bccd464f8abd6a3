"""Pins the CPU restatement (oracle/restatement.py) against the committed golden
vectors, which are outputs of the REAL reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from tests.helpers import GOLDEN_CONFIGS, assert_close, fixture, golden
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict

CPU_TOL = 2e-5  # CPU-vs-CPU fp32: summation order only


@pytest.mark.parametrize("name", sorted(GOLDEN_CONFIGS))
def test_sample_and_score_matches_reference(name):
  cfg, g = GOLDEN_CONFIGS[name], golden(name)
  inp, sds = fixture(cfg)
  B, K, T = cfg["B"], cfg["K"], cfg["T"]
  with torch.no_grad():
    vis = R.transform_visual(inp["lidar"])
    assert_close(vis, g["visual_features"], 1e-6, "visual_features")
    zs = [R.imitative_params(sd, vis, inp["velocity"], inp["is_at_traffic_light"],
                             inp["traffic_light_state"]) for sd in sds]
    assert_close(torch.stack(zs), g["z"], CPU_TOL, "z")
    rep = lambda z: z.repeat_interleave(K, dim=0)
    y, lad_f = R.flow_forward(sds[0], inp["x"].reshape(B * K, T, 2), rep(zs[0]))
    assert_close(y.view(B, K, T, 2), g["y"], CPU_TOL, "y")
    assert_close(lad_f.view(B, K), g["fwd_logabsdet"], CPU_TOL, "fwd logabsdet")
    xi, lp, lad = R.flow_inverse(sds[1], y, rep(zs[1]))
    assert_close(xi.view(B, K, T, 2), g["inv1_x"], CPU_TOL, "inverse x")
    assert_close(lp.view(B, K), g["inv1_log_prob"], CPU_TOL, "log_prob")
    assert_close(lad.view(B, K), g["inv1_logabsdet"], CPU_TOL, "logabsdet")
    gl = R.goal_log_likelihood_rows(y.view(B, K, T, 2)[:, :, -1], inp["goal"].unsqueeze(1), 1.0)
    assert_close(gl, g["goal_ll"], CPU_TOL, "goal ll")
    assert_close(R.goal_log_likelihood(y.view(B, K, T, 2)[:, 0], inp["goal"], 1.0),
                 g["goal_ll_batchmean_k0"], CPU_TOL, "goal ll batch mean")
    for algo in ("WCM", "BCM", "MA"):
      out = R.rip_sample_and_score(sds, zs, inp["x"], inp["goal"], 1.0, algo)
      assert_close(out["q"], g["q"], CPU_TOL, "q")
      assert_close(out["s"], g["s_" + algo], CPU_TOL, "s " + algo)
      assert np.array_equal(out["kstar"].numpy(), g["kstar_" + algo]), algo
      assert_close(out["plan"], g["plan_" + algo], CPU_TOL, "plan " + algo)
    # aggregation must be bit-exact on the reference's own q
    for algo in ("WCM", "BCM", "MA"):
      s = R.rip_aggregate(torch.from_numpy(g["q"]), algo)
      assert torch.equal(s, torch.from_numpy(g["s_" + algo])), algo


@pytest.mark.parametrize("name", sorted(GOLDEN_CONFIGS))
def test_gradient_planners_match_reference(name):
  cfg, g = GOLDEN_CONFIGS[name], golden(name)
  inp, sds = fixture(cfg)
  B, T = cfg["B"], cfg["T"]
  with torch.no_grad():
    vis = R.transform_visual(inp["lidar"])
  ctx = dict(visual_features=vis, velocity=inp["velocity"],
             is_at_traffic_light=inp["is_at_traffic_light"],
             traffic_light_state=inp["traffic_light_state"])
  x0 = torch.from_numpy(g["dim_forward_x0"]).repeat(B, 1, 1)
  y = R.imitative_forward(sds[0], x0, 10, goal=inp["goal"], lr=1e-1, epsilon=1.0, **ctx)
  assert_close(y, g["dim_forward_goal"], 1e-4, "ImitativeModel.forward (goal)")
  y = R.imitative_forward(sds[0], x0, 10, goal=None, lr=5e-2, epsilon=1.0, **ctx)
  assert_close(y, g["dim_forward_nogoal"], 1e-4, "ImitativeModel.forward (no goal)")
  # RIPAgent.__call__ on scene 0: planner + interpolation (rip/agent.py:84-151)
  one = {k: v[:1] for k, v in ctx.items()}
  with torch.no_grad():
    zs = [R.imitative_params(sd, **one) for sd in sds]
  for algo in ("WCM", "BCM", "MA"):
    plan = R.rip_plan(sds, zs, T, goal=inp["goal"][:1], num_steps=10, lr=1e-1, epsilon=1.0,
                      algorithm=algo)
    assert_close(R.interpolate_plan(plan[0].numpy()), g["rip_agent_" + algo], 1e-4,
                 "RIPAgent " + algo)
  x0 = torch.from_numpy(g["dim_agent_x0"])
  # DIMAgent forwards the whole observation, goal included (dim/agent.py:69-72)
  plan = R.planner([sds[0]], [zs[0]], x0, 20, 5e-2, inp["goal"][:1], 1.0, None)[0]
  assert_close(R.interpolate_plan(plan[0].numpy()), g["dim_agent"], 1e-4, "DIMAgent")


def test_behavioural_model_matches_reference():
  g = golden("cil_T4_C2")
  inp = synthetic_inputs(3, 2, 1, 4, seed=9)
  sd = synthetic_state_dict("cil", 2, 300)
  mode = torch.tensor([[0.0], [2.0], [3.0]])
  with torch.no_grad():
    vis = R.transform_visual(inp["lidar"])
    plan = R.behavioural_forward(sd, 4, vis, inp["velocity"], inp["is_at_traffic_light"],
                                 inp["traffic_light_state"], mode)
  assert_close(plan, g["plan"], CPU_TOL, "BehaviouralModel.forward")


def test_flow_round_trip_property():
  """_inverse(_forward(x)) == x (reference achieves ~1e-7)."""
  sd = synthetic_state_dict("dim", 2, 7)
  g = torch.Generator().manual_seed(3)
  x = torch.randn(64, 10, 2, generator=g)
  z = torch.randn(64, 64, generator=g).abs() * 0.3
  with torch.no_grad():
    y, lad_f = R.flow_forward(sd, x, z)
    xr, lp, lad_i = R.flow_inverse(sd, y, z)
  assert (xr - x).abs().max().item() < 2e-5
  assert_close(lad_f, lad_i, 1e-5, "forward/inverse logabsdet")


def test_cfg1_imitative_forward_b4_c4_matches_reference():
  """BASELINE.json configs[0] (the reference's own CPU-runnable case): `ImitativeModel.forward` on a
  batch of 4 synthetic 200x200x4 grids, 10 Adam steps — the restatement against the real reference's
  output (tests/golden/make_golden_cfg1.py)."""
  import os
  from tests.helpers import GOLDEN_DIR
  g = dict(np.load(os.path.join(GOLDEN_DIR, "cfg1_forward_B4_C4.npz")))
  T, C, B = 4, 4, 4
  inp = synthetic_inputs(B, C, 1, T, seed=61)
  sd = synthetic_state_dict("dim", C, 610)
  with torch.no_grad():
    vis = R.transform_visual(inp["lidar"])
    z = R.imitative_params(sd, vis, inp["velocity"], inp["is_at_traffic_light"], inp["traffic_light_state"])
  assert_close(z, g["z"], CPU_TOL, "cfg1 z")
  ctx = dict(visual_features=vis, velocity=inp["velocity"], is_at_traffic_light=inp["is_at_traffic_light"],
             traffic_light_state=inp["traffic_light_state"])
  x0 = torch.from_numpy(g["x0"]).repeat(B, 1, 1)
  y = R.imitative_forward(sd, x0, 10, goal=inp["goal"], lr=1e-1, epsilon=1.0, **ctx)
  assert_close(y, g["plan_goal"], 1e-4, "cfg1 ImitativeModel.forward (goal)")
  y = R.imitative_forward(sd, x0, 10, goal=None, lr=1e-1, epsilon=1.0, **ctx)
  assert_close(y, g["plan_nogoal"], 1e-4, "cfg1 ImitativeModel.forward (no goal)")
