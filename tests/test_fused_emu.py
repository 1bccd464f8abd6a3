"""Fused encoder front (oatomobile_b200/csrc/fused_body.h) without a GPU: the CTA bodies of
the `front_kernel` / `expand_dw_kernel` executed on the host (tests/emu) against the oracle's
layer-by-layer MobileNetV2 (perception.py:53-55 + torchvision, eval mode).  Checks the row
ring, the top/bottom/left/right zero padding, the row splits and the work-item tilings for
any thread count — the arithmetic is plain fp32 FMA, so the bar is 2e-5."""
import pytest
import torch
import torch.nn.functional as F

from oracle import restatement as R
from oatomobile_b200.synthetic import synthetic_inputs, synthetic_state_dict
from tests.emu import fused_driver as FD

PFX = "_encoder._model.features."


def _prefix_activations(sd, visual, blocks=4):
  """Oracle activations: stem out, block-1 out, then for blocks 2..`blocks` (dw output, output)."""
  with torch.no_grad():
    x = R._conv_bn_relu6(visual, sd, PFX + "0", stride=2, groups=1)
    acts = {"stem": x}
    for idx, cin, hid, cout, stride, res in R.mbv2_block_table()[:blocks]:
      p = PFX + "%d.conv" % idx
      h = x
      if hid != cin:
        h = R._conv_bn_relu6(h, sd, p + ".0", stride=1, groups=1)
        dw, pj, pjbn = p + ".1", p + ".2", p + ".3"
      else:
        dw, pj, pjbn = p + ".0", p + ".1", p + ".2"
      h = R._conv_bn_relu6(h, sd, dw, stride=stride, groups=hid)
      acts["dw%d" % idx] = h
      h = R._bn(F.conv2d(h, sd[pj + ".weight"], None), sd, pjbn)
      x = x + h if res else h
      acts["out%d" % idx] = x
  return acts


def _err(a, b):
  return ((a - b).abs() / torch.clamp(torch.maximum(a.abs(), b.abs()), min=1.0)).max().item()


@pytest.fixture(scope="module")
def case():
  C = 4
  sd = synthetic_state_dict("dim", C, 7)
  inp = synthetic_inputs(2, C, 1, 4, seed=11)
  visual = R.transform_visual(inp["lidar"])
  return sd, visual, _prefix_activations(sd, visual)


@pytest.mark.parametrize("splits,threads", [(1, 256), (2, 256), (3, 96), (5, 33)])
def test_front_matches_oracle(case, splits, threads):
  sd, visual, acts = case
  got = FD.front(sd, visual, splits=splits, threads=threads)
  assert not torch.isnan(got).any()
  want = acts["out1"].permute(0, 2, 3, 1)
  assert _err(got, want) < 2e-5


@pytest.mark.parametrize("C", [1, 2, 3, 8])
def test_front_other_channel_counts(C):
  sd = synthetic_state_dict("dim", C, 3)
  visual = R.transform_visual(synthetic_inputs(1, C, 1, 4, seed=5)["lidar"])
  got = FD.front(sd, visual)
  want = _prefix_activations(sd, visual)["out1"].permute(0, 2, 3, 1)
  assert not torch.isnan(got).any()
  assert _err(got, want) < 2e-5


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("idx", [2, 3, 4])
@pytest.mark.parametrize("splits,threads", [(1, 256), (2, 512), (4, 64), (3, 37)])
def test_expand_dw_matches_oracle(case, idx, splits, threads, tc):
  sd, _, acts = case
  x = acts["out%d" % (idx - 1)].permute(0, 2, 3, 1).contiguous()
  got = FD.expand_dw(idx, sd, x, splits=splits, threads=threads, tc=tc)
  assert not torch.isnan(got).any()
  want = acts["dw%d" % idx].permute(0, 2, 3, 1)
  assert got.shape == want.shape
  assert _err(got, want) < 2e-5


@pytest.mark.parametrize("threads", [128, 37])
def test_dw_project_matches_oracle(case, threads):
  sd, _, acts = case
  x = acts["stem"].permute(0, 2, 3, 1).contiguous()
  got = FD.dw_project(sd, x, threads=threads)
  assert not torch.isnan(got).any()
  assert _err(got, acts["out1"].permute(0, 2, 3, 1)) < 2e-5
