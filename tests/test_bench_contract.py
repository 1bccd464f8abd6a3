"""bench.py's JSON-line contract on the arm that runs without a GPU (`--impl reference`): the keys
the driver reads, the reference-backed CPU arm where a reference tree is present, and that ranks
other than 0 stay silent under torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
  r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=env,
                     stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
  assert r.returncode == 0, r.stderr.decode()[-2000:]
  return r.stdout.decode().strip()


def test_reference_arm_line():
  out = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-scenes", "2"])
  line = json.loads(out.splitlines()[-1])
  assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
  assert line["metric"] == "RIP trajectory samples scored/sec (ens=4,K=512,T=10)"
  assert line["value"] > 0 and line["steps"] == 1 and line["gpu_launches"] == 0
  assert "workload" in line["config"] and line["config"]["K"] == 512 and line["config"]["ensemble"] == 4
  cb = line["cpu_baseline"]
  assert cb["value"] == line["value"] and cb["cores"] >= 1 and cb["sample"]
  from oracle import reference_arm
  assert cb["kind"] == ("reference" if reference_arm.available() else "port")
  e2e = line["e2e"]
  assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
  out = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-scenes", "2", "--gpus", "2"],
             env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
  assert out == ""


def test_training_reference_arm_line():
  out = _run(["--impl", "reference", "--workload", "train-dim", "--steps", "1", "--warmup", "1", "--cpu-batch", "2"])
  line = json.loads(out.splitlines()[-1])
  assert line["impl"] == "reference" and "DIM training samples/sec" in line["metric"]
  assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
