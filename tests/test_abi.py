"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/oat_b200.h declares, and fails loudly (no fallback) without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(ROOT, "include", "oat_b200.h")).read()
  return sorted(set(re.findall(r"OAT_API\s+[\w\s\*]+?\b(oat_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
  from oatomobile_b200 import _native
  declared = _declared_symbols()
  assert len(declared) >= 16
  assert sorted(_native.SYMBOLS) == declared
  lib = ctypes.CDLL(_native.LIB_PATH)
  for name in declared:
    assert hasattr(lib, name), name
  assert _native.lib().oat_abi_version() == 1


def test_no_cpu_fallback():
  """Without the CUDA path the product raises; it never computes on the CPU."""
  import oatomobile_b200 as ob
  from oatomobile_b200 import _native
  model = ob.ImitativeModel(output_shape=(4, 2))
  ctx = dict(visual_features=torch.zeros(1, 2, 100, 100), velocity=torch.zeros(1, 3),
             is_at_traffic_light=torch.zeros(1, 1), traffic_light_state=torch.zeros(1, 1))
  with pytest.raises(_native.NativeLibraryError):
    model._params(**ctx)
  with pytest.raises(_native.NativeLibraryError):
    model.transform({"lidar": torch.zeros(1, 2, 200, 200)})
  with pytest.raises(_native.NativeLibraryError):
    model._decoder._forward(torch.zeros(1, 4, 2), torch.zeros(1, 64))


def test_product_never_imports_the_oracle():
  pkg = os.path.join(ROOT, "oatomobile_b200")
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        text = open(os.path.join(dirpath, f)).read()
        # comments may mention the oracle; importing, including or executing it may not
        bad = re.search(r"(from|import)\s+oracle|oracle[./](restatement|ref_shim|_ref)|"
                        r"#include\s*[\"<].*oracle", text)
        assert bad is None, (os.path.join(dirpath, f), bad.group(0))
        if f.endswith(".py"):  # nor the reference package itself (agents.py did)
          ref = re.search(r"^\s*(from\s+oatomobile(\.|\s)|import\s+oatomobile(\.|\s|$))", text, re.M)
          assert ref is None, (os.path.join(dirpath, f), ref.group(0))


def test_error_behaviour_matches_reference():
  """ValueError messages of dim/model.py:95-96,189-196 and cil/model.py:72-85."""
  import oatomobile_b200 as ob
  model = ob.ImitativeModel()
  with pytest.raises(ValueError, match="Missing `visual_features` keyword argument."):
    model._params(velocity=torch.zeros(1, 3))
  with pytest.raises(ValueError, match="Missing `traffic_light_state` keyword argument."):
    model._params(visual_features=torch.zeros(1, 2, 100, 100), velocity=torch.zeros(1, 3),
                  is_at_traffic_light=torch.zeros(1, 1))
  with pytest.raises(ValueError, match="Missing `visual_features` keyword argument."):
    model.forward(num_steps=1)
  cil = ob.BehaviouralModel()
  with pytest.raises(ValueError, match="Missing `mode` keyword argument."):
    cil(visual_features=torch.zeros(1, 2, 100, 100), velocity=torch.zeros(1, 3),
        is_at_traffic_light=torch.zeros(1, 1), traffic_light_state=torch.zeros(1, 1))
  from oatomobile_b200.rip import RIPScorer
  with pytest.raises(AssertionError):
    RIPScorer([model], algorithm="XYZ")


def test_state_dict_layout():
  """328 / 326 entries with the reference's key names and shapes."""
  import oatomobile_b200 as ob
  from oatomobile_b200.synthetic import synthetic_state_dict
  m = ob.ImitativeModel(output_shape=(10, 2), in_channels=4)
  sd = m.state_dict()
  assert len(sd) == 328
  assert tuple(sd["_encoder._model.features.0.0.weight"].shape) == (32, 4, 3, 3)
  assert tuple(sd["_encoder._model.classifier.1.weight"].shape) == (128, 1280)
  assert tuple(sd["_merger._model.0.weight"].shape) == (64, 133)
  assert tuple(sd["_decoder._decoder.weight_hh"].shape) == (192, 64)
  assert tuple(sd["_decoder._locscale._model.2.weight"].shape) == (4, 32)
  assert sd["_encoder._model.features.1.conv.0.1.num_batches_tracked"].dtype == torch.int64
  ref = synthetic_state_dict("dim", 4, 0)
  assert list(ref.keys()) == list(sd.keys())
  assert all(ref[k].shape == sd[k].shape for k in ref)
  m.load_state_dict(ref, strict=True)
  c = ob.BehaviouralModel(output_shape=(4, 2))
  assert len(c.state_dict()) == 326
  c.load_state_dict(synthetic_state_dict("cil", 2, 0), strict=True)
  assert tuple(c.state_dict()["_merger._model.0.weight"].shape) == (64, 134)


def test_checkpoint_interchange_with_reference(tmp_path):
  """A checkpoint written by `Checkpointer` loads into the reference class and back
  (SURVEY.md §8 a16); skipped where the reference tree is absent."""
  import oatomobile_b200 as ob
  from oatomobile_b200.savers import Checkpointer
  from oatomobile_b200.synthetic import synthetic_state_dict
  model = ob.ImitativeModel(output_shape=(4, 2))
  model.load_state_dict(synthetic_state_dict("dim", 2, 5))
  path = Checkpointer(model, str(tmp_path)).save(epoch=3)
  assert path.endswith("model-3.pt")
  other = ob.ImitativeModel(output_shape=(4, 2))
  Checkpointer(other, str(tmp_path)).load(epoch=3)
  assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(),
                                               other.state_dict().values()))
  from oracle import ref_shim
  if ref_shim.available():
    ref = ref_shim.make_imitative_model(T=4, in_channels=2, seed=0, randomize_bn=False)
    ref.load_state_dict(torch.load(path), strict=True)          # ours -> reference
    torch.save(ref.state_dict(), str(tmp_path / "ref.pt"))
    other.load_state_dict(torch.load(str(tmp_path / "ref.pt")), strict=True)  # reference -> ours


def test_trainer_rejects_cpu_models():
  """The training step has no CPU implementation either."""
  import oatomobile_b200 as ob
  from oatomobile_b200._native import NativeLibraryError
  from oatomobile_b200.train import Trainer
  with pytest.raises(NativeLibraryError):
    Trainer(ob.ImitativeModel(output_shape=(4, 2)))


def test_reference_citations_resolve():
  """Every `file.py:LINE` citation of the reference in our docs, headers and sources names an
  existing reference file with that many lines (tools/check_citations.py).  Needs the reference
  checkout, so it only runs in the build container."""
  import importlib.util
  spec = importlib.util.spec_from_file_location(
      "check_citations", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "check_citations.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  if not os.path.isdir(mod.REF):
    pytest.skip("reference checkout not present")
  total, bad = mod.scan()
  assert total > 100
  assert not bad, "\n".join(bad)
