"""Multi-GPU equivalence (needs >= 2 CUDA devices; skipped otherwise): the ensemble
sharded over 2 ranks with NCCL (z_0 broadcast + one all-gather of q) must reproduce the
single-GPU result — q bit-identical, k* and plan bit-identical on every rank."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E, B, K, T, C = 4, 6, 64, 10, 4


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out_path):
  sys.path.insert(0, ROOT)
  import torch.distributed as dist
  import oatomobile_b200 as ob
  from oatomobile_b200.rip import RIPScorer
  from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dev = torch.device("cuda", rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
  sds = [synthetic_state_dict("dim", C, 100 + m) for m in range(E)]
  inp = synthetic_inputs(B, C, K, T, seed=2)

  def make(m):
    model = ob.ImitativeModel(output_shape=(T, 2), in_channels=C)
    model.load_state_dict(sds[m], strict=True)
    return model.to(dev).eval()

  e_local = E // world
  mine = [make(m) for m in range(rank * e_local, (rank + 1) * e_local)]
  group = dist.new_group(list(range(world)))
  sharded = RIPScorer(mine, "WCM", group=group, proposal_model=None if rank == 0 else make(0))
  d = {k: v.to(dev) for k, v in inp.items()}
  x, goal = d.pop("x"), d.pop("goal")
  out = sharded(x=x, goal=goal, want_s=True, **d)
  single = RIPScorer([make(m) for m in range(E)], "WCM")
  ref = single(x=x, goal=goal, want_s=True, **d)
  torch.cuda.synchronize()
  ok = all(torch.equal(out[k], ref[k]) for k in ("q", "s", "kstar", "plan", "y"))
  torch.save({"ok": bool(ok)}, out_path % rank)
  dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_equals_single_gpu(tmp_path):
  import torch.multiprocessing as mp
  out_path = str(tmp_path / "r%d.pt")
  mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
  for r in range(2):
    assert torch.load(out_path % r)["ok"], "rank %d differs from the single-GPU result" % r
