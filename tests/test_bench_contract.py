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


def test_algorithmic_flops_match_the_survey():
  """bench.py's layer table reproduces SURVEY.md §6 / §8(d): 149.7 MFLOP per image at C=2 (pointwise
  137.3, depthwise 9.3, stem 2.9, fc 0.3) and 152.6 at C=4; the flow figure is 29 696 flop per row-step."""
  sys.path.insert(0, ROOT)
  import bench
  for C, want in ((2, 149.7), (4, 152.6)):
    L = bench.encoder_layers(C)
    by = {}
    for l in L:
      by[l[0]] = by.get(l[0], 0.0) + bench.layer_flops(l) / 1e6
    total = sum(by.values())
    assert abs(total - want) < 0.15, (C, total, by)
    assert abs(by["expand"] + by["project"] + by["last"] - 137.3) < 0.2
    assert abs(by["dw"] - 9.3) < 0.1 and abs(by["fc"] - 0.33) < 0.01
  assert bench.FLOW_FLOP_PER_ROW_STEP == 2 * (64 * 192 + 2 * 192 + 64 * 32 + 32 * 4)
  # every family's work is defined and the GEMM family is the 31 launches DESIGN §5 names
  w = bench.rip_family_work(bench.RIP_WORKLOADS["rip"], 4, 256, 131072, 3 * 131072, 30)
  assert abs(w["tc_pw_gemm"][0] / 1e9 - 121.5) < 0.3 and w["flow_score"][0] == 3 * 131072 * 10 * 29696
