"""GPU parity of the gradient-based planners and agents
against goldens produced by the REAL reference: ImitativeModel.forward,
RIPAgent.__call__, DIMAgent.__call__, CILAgent.__call__, _goal_likelihood."""
import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN_CONFIGS, assert_close, fixture, golden, observation

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# End point after 10-20 Adam steps: Adam normalises the gradient (step = lr * m / sqrt(v)), so a
# 1e-6 rounding difference in a near-zero gradient component moves x by a visible fraction of lr;
# round 1 therefore held the end point to 1e-3.  Measured on B200 the end points stay within 3.2e-5
# of the reference's (profiles/r2_parity_achieved.json), so since round 2 the end point is held to
# the north-star 1e-4 as well, next to the PER-STEP quantities (loss of every step, the first-step
# gradient through its Adam update) in `test_planner_per_step_quantities`.
PLAN_TOL = 1e-4
STEP_TOL = 1e-4


def _models(cfg, sds):
  import oatomobile_b200 as ob
  ms = []
  for sd in sds:
    m = ob.ImitativeModel(output_shape=(cfg["T"], 2), in_channels=cfg["C"])
    m.load_state_dict(sd, strict=True)
    ms.append(m.to(DEV).eval())
  return ms


def _bypass(cls, **attrs):
  agent = object.__new__(cls)  # SetPointAgent.__init__ needs CARLA
  for k, v in attrs.items():
    setattr(agent, k, v)
  return agent


@pytest.mark.parametrize("name", sorted(GOLDEN_CONFIGS))
def test_imitative_model_forward(name):
  cfg, g = GOLDEN_CONFIGS[name], golden(name)
  inp, sds = fixture(cfg)
  model = _models(cfg, sds[:1])[0]
  obs = model.transform({"lidar": inp["lidar"].to(DEV)})
  ctx = dict(visual_features=obs["visual_features"], velocity=inp["velocity"].to(DEV),
             is_at_traffic_light=inp["is_at_traffic_light"].to(DEV),
             traffic_light_state=inp["traffic_light_state"].to(DEV))
  x0 = torch.from_numpy(g["dim_forward_x0"])
  y = model(num_steps=10, goal=inp["goal"].to(DEV), lr=1e-1, epsilon=1.0, x0=x0, **ctx)
  assert_close(y, g["dim_forward_goal"], PLAN_TOL, "forward (goal)")
  y = model(num_steps=10, goal=None, lr=5e-2, epsilon=1.0, x0=x0, **ctx)
  assert_close(y, g["dim_forward_nogoal"], PLAN_TOL, "forward (no goal)")
  # _goal_likelihood on the reference's proposals
  yk0 = torch.from_numpy(g["y"][:, 0]).to(DEV)
  gl = model._goal_likelihood(y=yk0, goal=inp["goal"].to(DEV), epsilon=1.0)
  assert_close(gl, g["goal_ll_batchmean_k0"], 1e-5, "_goal_likelihood")


@pytest.mark.parametrize("name", sorted(GOLDEN_CONFIGS))
def test_agents_call(name):
  from oatomobile_b200.agents import DIMAgent, RIPAgent
  cfg, g = GOLDEN_CONFIGS[name], golden(name)
  inp, sds = fixture(cfg)
  models = _models(cfg, sds)
  for algo in ("WCM", "BCM", "MA"):
    agent = _bypass(RIPAgent, _algorithm=algo, _models=models, _device=torch.device(DEV))
    out = agent(observation(inp, 0))
    assert out.shape == g["rip_agent_" + algo].shape and out.dtype == np.float64
    assert_close(out, g["rip_agent_" + algo], PLAN_TOL, "RIPAgent " + algo)
  agent = _bypass(DIMAgent, _model=models[0], _device=torch.device(DEV))
  out = agent(observation(inp, 0), x0=torch.from_numpy(g["dim_agent_x0"]))
  assert_close(out, g["dim_agent"], PLAN_TOL, "DIMAgent")


def test_cil_agent_call():
  import oatomobile_b200 as ob
  from oatomobile_b200.agents import CILAgent
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  g = golden("cil_T4_C2")
  inp = synthetic_inputs(3, 2, 1, 4, seed=9)
  model = ob.BehaviouralModel(output_shape=(4, 2))
  model.load_state_dict(synthetic_state_dict("cil", 2, 300), strict=True)
  agent = _bypass(CILAgent, _model=model.to(DEV), _device=torch.device(DEV))
  out = agent(observation(inp, 0))
  assert_close(out, g["cil_agent"], 1e-4, "CILAgent")


def test_planner_per_step_quantities():
  """Per-step bars at 1e-4: the loss of each of the first steps and the first Adam update
  (x_1 = x_0 - lr * sign-like(grad_1), i.e. the first-step gradient) against the autograd oracle."""
  from oatomobile_b200 import ops
  from oracle import restatement as R
  cfg = GOLDEN_CONFIGS["dim_T4_C2"]
  inp, sds = fixture(cfg)
  models = _models(cfg, sds)
  g = torch.Generator().manual_seed(15)
  B, T, E = 3, cfg["T"], cfg["E"]
  zs = [(torch.randn(B, 64, generator=g) * 0.4).clamp(min=0) for _ in range(E)]
  goal = inp["goal"][:1].repeat(B, 1, 1) + torch.randn(B, 10, 2, generator=g) * 0.3
  x0 = torch.randn(B, T, 2, generator=g) * 0.5
  for algo in (None, "WCM", "BCM", "MA"):
    n = 1 if algo is None else E
    handles = [m.native_handle() for m in models[:n]]
    zt = torch.stack(zs[:n]).to(DEV)
    trace = {}
    R.planner(sds[:n], zs[:n], x0, 3, 0.1, goal, 1.0, algo, trace=trace)
    _, _, losses = ops.plan(handles, zt, x0.to(DEV), num_steps=3, lr=0.1, goal=goal.to(DEV),
                            epsilon=1.0, algorithm=algo, want_loss=True)
    assert_close(losses, torch.tensor(trace["losses"]), STEP_TOL, "planner loss trajectory (3 steps) %s" % algo)
    # one step: x_best is the post-step x (dim/model.py:133-137), i.e. x_0 - lr * g / (|g| + eps)
    ref1 = {}
    _, xb_ref = R.planner(sds[:n], zs[:n], x0, 1, 0.1, goal, 1.0, algo, trace=ref1)
    _, xb, l1 = ops.plan(handles, zt, x0.to(DEV), num_steps=1, lr=0.1, goal=goal.to(DEV), epsilon=1.0,
                         algorithm=algo, want_loss=True)
    assert_close(l1, torch.tensor(ref1["losses"]), 1e-5, "planner first loss %s" % algo)
    # Adam's first step is lr * g / (|g| + eps): compare where the oracle's gradient is not at
    # rounding level (a component within 1e-6 of zero may legitimately take the other sign)
    clear = (ref1["grad1"].abs() > 1e-5).to(DEV)
    assert bool(clear.float().mean() > 0.9)
    assert_close(torch.where(clear, xb, xb_ref.to(DEV)), xb_ref, STEP_TOL, "planner first Adam update %s" % algo)


def test_planner_matches_oracle_losses():
  """End points of the fused planner vs the autograd oracle, B > 1, every algorithm."""
  from oatomobile_b200 import ops
  from oracle import restatement as R
  cfg = GOLDEN_CONFIGS["dim_T4_C2"]
  inp, sds = fixture(cfg)
  models = _models(cfg, sds)
  g = torch.Generator().manual_seed(5)
  B, T, E = 3, cfg["T"], cfg["E"]
  zs = [(torch.randn(B, 64, generator=g) * 0.4).clamp(min=0) for _ in range(E)]
  goal = inp["goal"][:1].repeat(B, 1, 1) + torch.randn(B, 10, 2, generator=g) * 0.3
  x0 = torch.randn(B, T, 2, generator=g) * 0.5
  for algo in (None, "WCM", "BCM", "MA"):
    n = 1 if algo is None else E
    # 3 steps: with random (untrained, unrelated) z the min/max over models can flip
    # between near-tied models after a few steps, which is a discontinuity, not an error;
    # the 10-20 step trajectories are pinned by the reference goldens above.
    ref_y, ref_xb = R.planner(sds[:n], zs[:n], x0, 3, 0.1, goal, 1.0, algo)
    plan, xb, _ = ops.plan([m.native_handle() for m in models[:n]], torch.stack(zs[:n]).to(DEV),
                           x0.to(DEV), num_steps=3, lr=0.1, goal=goal.to(DEV), epsilon=1.0,
                           algorithm=algo)
    assert_close(xb, ref_xb, PLAN_TOL, "x_best %s" % algo)
    assert_close(plan, ref_y, PLAN_TOL, "plan %s" % algo)
  # zero steps: plan = f_0(x0)
  plan, xb, _ = ops.plan([models[0].native_handle()], zs[0][None].to(DEV), x0.to(DEV), 0, 0.1,
                         None, 1.0, None)
  with torch.no_grad():
    ref_y, _ = R.flow_forward(sds[0], x0, zs[0])
  assert_close(plan, ref_y, 1e-5, "zero-step plan")
