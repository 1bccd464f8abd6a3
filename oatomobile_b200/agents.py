"""`RIPAgent`, `DIMAgent`, `CILAgent` — mirrors of
oatomobile/baselines/torch/{rip,dim,cil}/agent.py on the CUDA path.

`__call__(observation) -> np.ndarray [N,3]` (ego-frame waypoints) has the reference's
semantics: the same observation pre-processing, the gradient-based planner as written
(`ops.plan`, one fused kernel) or the CIL roll-out, and the same linear interpolation.
`act()` (waypoints -> `carla.VehicleControl`) needs the CARLA PythonAPI PID controller
exactly like the reference's `SetPointAgent` (oatomobile/baselines/base.py:46-176) and
raises the same ImportError without it; it is simulator plumbing, outside the hot path.
"""
import copy
from typing import Any, Mapping, Optional, Sequence

import numpy as np
import torch

from oatomobile_b200 import geometry, ops
from oatomobile_b200.models import BehaviouralModel, ImitativeModel


def _prepare_observation(observation: Mapping[str, Any], device) -> Mapping[str, torch.Tensor]:
  """rip/agent.py:59-74: float32, batch dim, 2-D goals, lidar HWC -> CHW, to device.

  Only the modalities the models read are moved (the reference also uploads the
  unused camera images)."""
  keys = ("lidar", "velocity", "is_at_traffic_light", "traffic_light_state", "goal", "mode")
  out = {}
  for k in keys:
    if k not in observation:
      continue
    v = observation[k]
    if not isinstance(v, np.ndarray):
      v = np.atleast_1d(v)
    v = v[None, ...].astype(np.float32)
    if k == "goal":
      v = v[..., :2]
    if k == "lidar":
      v = np.transpose(v, (0, 3, 1, 2))
    out[k] = torch.from_numpy(np.ascontiguousarray(v)).to(device)
  return out


def interpolate_plan(plan: np.ndarray) -> np.ndarray:
  """rip/agent.py:141-151: linear interpolation of [T,2] onto integer time steps of a
  40-frame horizon, plus a zero z column → float64 [N,3]."""
  player_future_length = 40
  increments = player_future_length // plan.shape[0]
  time_index = np.arange(0, player_future_length, increments)
  if time_index.shape[0] != plan.shape[0]:
    # the reference hands both to scipy's interp1d, which rejects them (T must divide 40)
    raise ValueError("x and y arrays must be equal in length along interpolation axis "
                     "(%d time steps for a plan of %d waypoints)" % (time_index.shape[0], plan.shape[0]))
  query = np.arange(0, time_index[-1])
  xy = np.stack([np.interp(query, time_index, plan[:, d].astype(np.float64)) for d in range(2)],
                axis=-1)
  return np.c_[xy, np.zeros((xy.shape[0], 1))]


class SetPointAgent:
  """baselines/base.py:46-176 — setpoint agent driven by CARLA's PID controller."""

  def __init__(self, environment, *, setpoint_index: int = 5, replan_every_steps: int = 1,
               lateral_control_dict=None, longitudinal_control_dict=None,
               fixed_delta_seconds_between_setpoints: Optional[int] = None) -> None:
    try:
      from agents.navigation.controller import VehiclePIDController  # pylint: disable=import-error
    except ImportError:
      raise ImportError("Missing CARLA installation, "
                        "make sure the environment variable CARLA_ROOT is provided "
                        "and that the PythonAPI is `easy_install`ed")
    self._environment = environment
    self._vehicle = environment.simulator.hero
    self._world = self._vehicle.get_world()
    self._map = self._world.get_map()
    dt = self._world.get_settings().fixed_delta_seconds
    lat = dict(lateral_control_dict or {"K_P": 1.0, "K_D": 0.0, "K_I": 0.0}, dt=dt)
    lon = dict(longitudinal_control_dict or {"K_P": 1.0, "K_D": 0.0, "K_I": 1.0}, dt=dt)
    self._vehicle_controller = VehiclePIDController(vehicle=self._vehicle, args_lateral=lat,
                                                    args_longitudinal=lon)
    self._setpoint_index = setpoint_index
    self._replan_every_steps = replan_every_steps
    self._fixed_delta_seconds_between_setpoints = fixed_delta_seconds_between_setpoints or dt
    self._setpoints_buffer = None
    self._steps_counter = 0

  def act(self, observation, *args, **kwargs):
    """baselines/base.py:116-176 (replanning cadence, local->world, PID step).  The plan's
    frame change is `geometry.local2world` (utils/carla.py:677-700); only the waypoint lookup
    and the PID step touch the CARLA PythonAPI."""
    import carla  # pylint: disable=import-error  (the simulator's own API)
    loc, rot = observation["location"], observation["rotation"]
    if self._setpoints_buffer is None or self._steps_counter % self._replan_every_steps == 0:
      plan_ego = self(copy.deepcopy(observation), *args, **kwargs)
      self._setpoints_buffer = geometry.local2world(current_location=loc, current_rotation=rot,
                                                    local_locations=plan_ego)
    else:
      self._setpoints_buffer = self._setpoints_buffer[1:]
    self._steps_counter += 1
    speed = np.linalg.norm(np.diff(self._setpoints_buffer[:self._setpoint_index], axis=0),
                           axis=1).mean() / self._fixed_delta_seconds_between_setpoints
    setpoint = self._map.get_waypoint(
        carla.Location(*map(float, self._setpoints_buffer[self._setpoint_index])))
    if self._steps_counter <= 100:
      speed = 20.0 / 3.6
    return self._vehicle_controller.run_step(target_speed=speed * 3.6, waypoint=setpoint)


class RIPAgent(SetPointAgent):
  """rip/agent.py:30-151 — robust imitative planning over an ensemble."""

  def __init__(self, environment, *, algorithm: str, models: Sequence[ImitativeModel],
               **kwargs) -> None:
    assert algorithm in ("WCM", "MA", "BCM")
    self._algorithm = algorithm
    super().__init__(environment=environment, **kwargs)
    self._device = torch.device("cuda")
    self._models = [model.to(self._device) for model in models]

  def __call__(self, observation: Mapping[str, np.ndarray]) -> np.ndarray:
    obs = _prepare_observation(observation, self._device)
    obs = self._models[0].transform(obs)
    goal = obs.pop("goal")
    # rip/agent.py:78-80: hard-coded planner hyper-parameters
    lr, epsilon, num_steps = 1e-1, 1.0, 10
    zs = torch.stack([m._params(**obs) for m in self._models])
    batch = zs.shape[1]
    x0 = torch.zeros(batch, *self._models[0]._output_shape, device=self._device)  # base mean
    plan, _, _ = ops.plan([m.native_handle() for m in self._models], zs, x0, num_steps=num_steps,
                          lr=lr, goal=goal, epsilon=epsilon, algorithm=self._algorithm)
    return interpolate_plan(plan.detach().cpu().numpy()[0])


class DIMAgent(SetPointAgent):
  """dim/agent.py:28-84 — single deep imitative model."""

  def __init__(self, environment, *, model: ImitativeModel, **kwargs) -> None:
    super().__init__(environment=environment, **kwargs)
    self._device = torch.device("cuda")
    self._model = model.to(self._device)

  def __call__(self, observation: Mapping[str, np.ndarray], **kwargs) -> np.ndarray:
    obs = _prepare_observation(observation, self._device)
    obs = self._model.transform(obs)
    plan = self._model(num_steps=kwargs.get("num_steps", 20), epsilon=kwargs.get("epsilon", 1.0),
                       lr=kwargs.get("lr", 5e-2), x0=kwargs.get("x0"), **obs)
    return interpolate_plan(plan.detach().cpu().numpy()[0])


class CILAgent(SetPointAgent):
  """cil/agent.py:28-97 — conditional imitation learner."""

  def __init__(self, environment, *, model: BehaviouralModel, **kwargs) -> None:
    super().__init__(environment=environment, **kwargs)
    self._device = torch.device("cuda")
    self._model = model.to(self._device)

  def __call__(self, observation: Mapping[str, np.ndarray]) -> np.ndarray:
    # cil/agent.py:64-76: command from the last goal (thresholds as written)
    goal = np.asarray(observation["goal"], dtype=np.float32)[..., :2]
    x_t, y_t = goal[-1, :2]
    norm = np.linalg.norm([x_t, y_t])
    theta = np.degrees(np.arccos(x_t / (norm + 1e-3)))
    if norm < 3:
      mode = 1
    elif theta > 15:
      mode = 2
    elif theta <= 15:
      mode = 3
    else:
      mode = 0
    observation = dict(observation)
    observation["mode"] = np.float32(mode)
    obs = _prepare_observation(observation, self._device)
    obs["mode"] = obs["mode"].reshape(1, 1)
    obs.pop("goal", None)
    obs = self._model.transform(obs)
    plan = self._model(**obs)
    return interpolate_plan(plan.detach().cpu().numpy()[0])
