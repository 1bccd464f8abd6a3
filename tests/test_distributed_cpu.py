"""World-size-2 `gloo` test of the ensemble-sharding host logic (rip.RIPScorer with a
process group): the CUDA ops are replaced by oracle-backed CPU stand-ins, so what is
exercised is the collective plumbing — z_0 broadcast, locally regenerated proposals,
the single all-gather of q in global model order, identical aggregation on every
rank — against the single-process oracle result."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

E, B, K, T, C = 4, 3, 8, 4, 2


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  return port


class _Handle:

  def __init__(self, sd):
    self.sd = sd


class _FakeDecoder:

  def __init__(self, sd):
    self._h = _Handle(sd)

  def _handle(self):
    return self._h


class _FakeModel:
  """Stands in for ImitativeModel: only what RIPScorer touches."""

  def __init__(self, sd):
    self._h = _Handle(sd)
    self._decoder = _FakeDecoder(sd)

  def native_handle(self):
    return self._h


def _install_cpu_ops():
  from oracle import restatement as R
  from oatomobile_b200 import _native, ops

  class FakeEnsemble:

    def __init__(self, handles):
      self.models = list(handles)

    def __len__(self):
      return len(self.models)

  def encode(ens, visual, scalars):
    zs = [R.imitative_params(h.sd, visual, scalars[:, :3], scalars[:, 3:4], scalars[:, 4:5])
          for h in ens.models]
    return torch.stack(zs)

  def flow_forward(handle, x, z, rows_per_z=1):
    return R.flow_forward(handle.sd, x, z.repeat_interleave(rows_per_z, dim=0))

  def rip_sample_score(ens, z, x, goal, epsilon, proposal_idx=0, y=None, q=None):
    Bn = z.shape[1]
    if proposal_idx >= 0:
      Kn, Tn = x.shape[1], x.shape[2]
      y, _ = R.flow_forward(ens.models[proposal_idx].sd, x.reshape(Bn * Kn, Tn, 2),
                            z[proposal_idx].repeat_interleave(Kn, dim=0))
      y = y.view(Bn, Kn, Tn, 2)
    Kn, Tn = y.shape[1], y.shape[2]
    q = torch.empty(len(ens), Bn, Kn)
    for m, h in enumerate(ens.models):
      _, lp, lad = R.flow_inverse(h.sd, y.reshape(Bn * Kn, Tn, 2), z[m].repeat_interleave(Kn, dim=0))
      q[m] = (lp - lad).view(Bn, Kn)
    if goal is not None:
      q = q + R.goal_log_likelihood_rows(y[:, :, -1], goal.unsqueeze(1), epsilon).unsqueeze(0)
    return y, q

  def rip_aggregate(q, y, algorithm, want_s=False):
    s = R.rip_aggregate(q, algorithm)
    ks = torch.argmin(s, dim=1)
    idx = torch.arange(q.shape[1])
    return ks.int(), s[idx, ks], y[idx, ks], (s if want_s else None)

  _native.EnsembleHandle = FakeEnsemble
  ops.transform_visual = R.transform_visual
  ops.encode, ops.flow_forward = encode, flow_forward
  ops.rip_sample_score, ops.rip_aggregate = rip_sample_score, rip_aggregate


def _worker(rank, world, port, algo, out_path):
  sys.path.insert(0, ROOT)
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.set_num_threads(2)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  _install_cpu_ops()
  from oracle import restatement as R
  from oatomobile_b200.rip import RIPScorer
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  sds = [synthetic_state_dict("dim", C, 700 + m) for m in range(E)]
  inp = synthetic_inputs(B, C, K, T, seed=4)
  e_local = E // world
  mine = [_FakeModel(sds[m]) for m in range(rank * e_local, (rank + 1) * e_local)]
  group = dist.new_group(list(range(world)))
  scorer = RIPScorer(mine, algo, group=group, flow_sharding="models",
                     proposal_model=None if rank == 0 else _FakeModel(sds[0]))
  with torch.no_grad():
    # raw grids: exercises the sharded resize + all-gather of visual features too (B=3 is not
    # divisible by 2 ranks for "MA" -> replicated resize; B=4 for "WCM" -> sharded)
    if algo == "WCM":
      inp = synthetic_inputs(4, C, K, T, seed=4)
    out = scorer(x=inp["x"], goal=inp["goal"], epsilon=1.0, want_s=True, lidar=inp["lidar"],
                 velocity=inp["velocity"], is_at_traffic_light=inp["is_at_traffic_light"],
                 traffic_light_state=inp["traffic_light_state"])
    ref = R.rip_score_from_inputs(sds, inp["lidar"], inp["velocity"], inp["is_at_traffic_light"],
                                  inp["traffic_light_state"], inp["x"], inp["goal"], 1.0, algo)
  ok = (torch.equal(out["q"], ref["q"]) and torch.equal(out["s"], ref["s"]) and
        torch.equal(out["kstar"].long(), ref["kstar"]) and torch.equal(out["plan"], ref["plan"]))
  if algo == "WCM":
    # rank-local feed (`local_slice=True`, what HostRIPPipeline uploads): each rank passes only
    # its B/R scenes; grids are resized locally, visual features + packed context all-gathered
    n = inp["lidar"].shape[0] // world
    sl = slice(rank * n, (rank + 1) * n)
    with torch.no_grad():
      loc = scorer(x=inp["x"][sl], goal=inp["goal"][sl], epsilon=1.0, want_s=True, local_slice=True,
                   lidar=inp["lidar"][sl], velocity=inp["velocity"][sl],
                   is_at_traffic_light=inp["is_at_traffic_light"][sl],
                   traffic_light_state=inp["traffic_light_state"][sl])
    ok = ok and all(torch.equal(loc[k], out[k]) for k in ("q", "s", "kstar", "plan", "y", "z"))
    # flow_sharding="scenes": z all-gathered, every rank runs the full flow stage on its scenes
    # with replicas of all decoders (here: the fake handles of all E models), results gathered
    from oatomobile_b200 import _native

    def fake_replicas(self):
      self._flow_replicas = [_FakeDecoder(sd) for sd in sds]
      self._flow_ens = _native.EnsembleHandle([d._handle() for d in self._flow_replicas])

    RIPScorer._replicate_decoders = fake_replicas
    by_scene = RIPScorer(mine, algo, group=group, flow_sharding="scenes")
    with torch.no_grad():
      sc = by_scene(x=inp["x"], goal=inp["goal"], epsilon=1.0, want_s=True, lidar=inp["lidar"],
                    velocity=inp["velocity"], is_at_traffic_light=inp["is_at_traffic_light"],
                    traffic_light_state=inp["traffic_light_state"])
      sc_loc = by_scene(x=inp["x"][sl], goal=inp["goal"][sl], epsilon=1.0, want_s=True, local_slice=True,
                        gather_details=False, lidar=inp["lidar"][sl], velocity=inp["velocity"][sl],
                        is_at_traffic_light=inp["is_at_traffic_light"][sl],
                        traffic_light_state=inp["traffic_light_state"][sl])
    ok = ok and all(torch.equal(sc[k], out[k]) for k in ("q", "s", "kstar", "plan", "y"))
    ok = ok and all(torch.equal(sc_loc[k], out[k]) for k in ("kstar", "plan", "sbest"))
    ok = ok and torch.equal(sc_loc["q"], out["q"][:, sl])  # details stay rank-local on the hot path
  # every rank must hold the same selection
  ks = [torch.empty_like(out["kstar"]) for _ in range(world)]
  dist.all_gather(ks, out["kstar"])
  ok = ok and all(torch.equal(k, ks[0]) for k in ks)
  torch.save({"ok": bool(ok), "q_shape": tuple(out["q"].shape)}, out_path % rank)
  dist.destroy_process_group()


def _worker_groups(rank, world, port, out_path):
  """bench.py's layout beyond E ranks: world 4, E = 2 -> two replica groups of 2 ranks, each
  group scoring its own scenes; collectives stay inside the group."""
  sys.path.insert(0, ROOT)
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.set_num_threads(1)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  _install_cpu_ops()
  from oracle import restatement as R
  from oatomobile_b200.rip import RIPScorer
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  E2 = 2
  gsize = min(world, E2)
  group_id, grank = rank // gsize, rank % gsize
  group = None
  for g in range(world // gsize):
    pg = dist.new_group(list(range(g * gsize, (g + 1) * gsize)))
    if g == group_id:
      group = pg
  sds = [synthetic_state_dict("dim", C, 800 + m) for m in range(E2)]
  inp = synthetic_inputs(2, C, K, T, seed=10 + group_id)  # each group has its own scenes
  scorer = RIPScorer([_FakeModel(sds[grank])], "WCM", group=group, flow_sharding="models",
                     proposal_model=None if grank == 0 else _FakeModel(sds[0]))
  with torch.no_grad():
    vis = R.transform_visual(inp["lidar"])
    out = scorer(x=inp["x"], goal=inp["goal"], epsilon=1.0, visual_features=vis,
                 velocity=inp["velocity"], is_at_traffic_light=inp["is_at_traffic_light"],
                 traffic_light_state=inp["traffic_light_state"])
    ref = R.rip_score_from_inputs(sds, inp["lidar"], inp["velocity"], inp["is_at_traffic_light"],
                                  inp["traffic_light_state"], inp["x"], inp["goal"], 1.0, "WCM")
  ok = torch.equal(out["q"], ref["q"]) and torch.equal(out["kstar"].long(), ref["kstar"])
  t = torch.tensor([1.0 if ok else 0.0])
  dist.all_reduce(t, op=dist.ReduceOp.MIN)  # global collective still works next to the groups
  torch.save({"ok": bool(t.item() == 1.0)}, out_path % rank)
  dist.destroy_process_group()


def _worker_scenes4(rank, world, port, out_path):
  """bench.py's default layout at N = 4 = E: one encoder per rank, rank-local feed of B/4 scenes,
  `flow_sharding="scenes"` (z all-gathered, every rank decodes its own scenes with replicas of all
  four decoders), the small per-scene results all-gathered — for all three aggregations."""
  sys.path.insert(0, ROOT)
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.set_num_threads(1)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  _install_cpu_ops()
  from oracle import restatement as R
  from oatomobile_b200 import _native
  from oatomobile_b200.rip import RIPScorer
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  sds = [synthetic_state_dict("dim", C, 900 + m) for m in range(world)]
  Bs = 2 * world
  inp = synthetic_inputs(Bs, C, K, T, seed=21)
  group = dist.new_group(list(range(world)))

  def fake_replicas(self):
    self._flow_replicas = [_FakeDecoder(sd) for sd in sds]
    self._flow_ens = _native.EnsembleHandle([d._handle() for d in self._flow_replicas])

  RIPScorer._replicate_decoders = fake_replicas
  sl = slice(rank * 2, rank * 2 + 2)
  ok = True
  for algo in ("WCM", "MA", "BCM"):
    scorer = RIPScorer([_FakeModel(sds[rank])], algo, group=group)  # default sharding = scenes
    with torch.no_grad():
      out = scorer(x=inp["x"][sl], goal=inp["goal"][sl], epsilon=1.0, want_s=True, local_slice=True,
                   gather_details=False, lidar=inp["lidar"][sl], velocity=inp["velocity"][sl],
                   is_at_traffic_light=inp["is_at_traffic_light"][sl],
                   traffic_light_state=inp["traffic_light_state"][sl])
      ref = R.rip_score_from_inputs(sds, inp["lidar"], inp["velocity"], inp["is_at_traffic_light"],
                                    inp["traffic_light_state"], inp["x"], inp["goal"], 1.0, algo)
    ok = ok and torch.equal(out["kstar"].long(), ref["kstar"]) and torch.equal(out["plan"], ref["plan"])
    ok = ok and torch.equal(out["q"], ref["q"][:, sl])
    ok = ok and tuple(out["plan"].shape) == (Bs, T, 2)
  t = torch.tensor([1.0 if ok else 0.0])
  dist.all_reduce(t, op=dist.ReduceOp.MIN)
  torch.save({"ok": bool(t.item() == 1.0)}, out_path % rank)
  dist.destroy_process_group()


def test_scene_sharded_flow_world4_matches_single_process(tmp_path):
  world = 4
  out_path = str(tmp_path / "rank%d.pt")
  mp.spawn(_worker_scenes4, args=(world, _free_port(), out_path), nprocs=world, join=True)
  assert all(torch.load(out_path % r)["ok"] for r in range(world))


def test_two_replica_groups_world4(tmp_path):
  world = 4
  out_path = str(tmp_path / "rank%d.pt")
  mp.spawn(_worker_groups, args=(world, _free_port(), out_path), nprocs=world, join=True)
  assert all(torch.load(out_path % r)["ok"] for r in range(world))


@pytest.mark.parametrize("algo", ["WCM", "MA"])
def test_sharded_scorer_world2_matches_single_process(tmp_path, algo):
  world = 2
  port = _free_port()
  out_path = str(tmp_path / "rank%d.pt")
  mp.spawn(_worker, args=(world, port, algo, out_path), nprocs=world, join=True)
  for r in range(world):
    res = torch.load(out_path % r)
    assert res["ok"], "rank %d diverged from the single-process oracle" % r
    assert res["q_shape"] == (E, 4 if algo == "WCM" else B, K)


def test_non_zero_rank_needs_proposal_model():
  from oatomobile_b200.rip import RIPScorer

  class G:  # minimal stand-in: RIPScorer only queries rank/size at construction
    pass

  import torch.distributed as d
  orig = (d.get_rank, d.get_world_size)
  try:
    d.get_rank = lambda g=None: 1
    d.get_world_size = lambda g=None: 2
    with pytest.raises(ValueError):
      RIPScorer([], "WCM", group=G(), proposal_model=None, flow_sharding="models")
  finally:
    d.get_rank, d.get_world_size = orig
