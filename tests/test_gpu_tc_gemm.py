"""Unit parity of the tcgen05/TMA 3xTF32 pointwise GEMM against float64 matmul.

Shapes cover every (K, N) family of the encoder: K below one swizzle atom (16, 24),
K not a multiple of 32 (144), N not a multiple of 16 (24), multi-tile N (320..1280),
ragged M (tail tile), E > 1, ReLU6 and residual epilogues."""
import pytest
import torch

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [
    # (E, M, K, N, relu6, residual)
    (1, 128, 32, 16, 0, 0),
    (1, 128, 32, 256, 0, 0),
    (2, 300, 16, 96, 1, 0),
    (1, 1000, 24, 144, 1, 0),
    (3, 257, 144, 24, 0, 1),
    (2, 640, 96, 24, 0, 0),
    (1, 513, 192, 64, 0, 1),
    (2, 200, 384, 96, 0, 0),
    (1, 400, 576, 160, 0, 0),
    (2, 130, 160, 960, 1, 0),
    (1, 200, 64, 224, 0, 0),
    (1, 100, 32, 8, 0, 0),
    (1, 256, 960, 320, 0, 0),
    (1, 272, 320, 1280, 1, 0),
    (4, 64, 1280, 128, 0, 0),
    (4, 5000, 64, 384, 1, 0),
]


@pytest.mark.parametrize("E,M,K,N,relu6,residual", SHAPES)
def test_tc_gemm_matches_fp64(E, M, K, N, relu6, residual):
  from oatomobile_b200 import _native as Nat
  g = torch.Generator().manual_seed(E * 1000003 + M * 131 + K * 7 + N)
  A = torch.randn(E, M, K, generator=g)
  W = torch.randn(E, N, K, generator=g) / (K**0.5)
  bias = torch.randn(E, N, generator=g)
  R = torch.randn(E, M, N, generator=g) if residual else None
  ref = torch.einsum("emk,enk->emn", A.double(), W.double()) + bias.double().unsqueeze(1)
  if relu6:
    ref = ref.clamp(0.0, 6.0)
  if residual:
    ref = ref + R.double()
  Ad, Wd, bd = A.to(DEV), W.to(DEV), bias.to(DEV)
  Rd = R.to(DEV) if residual else None
  C = torch.full((E, M, N), float("nan"), device=DEV)
  with torch.cuda.device(DEV):
    Nat.check(Nat.lib().oat_debug_tc_gemm(Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr(),
                                          Nat.ptr(Rd), C.data_ptr(), M, K, N, E, relu6,
                                          Nat.stream_ptr(torch.device(DEV))))
  torch.cuda.synchronize()
  assert torch.isfinite(C).all(), "unwritten or non-finite outputs"
  # 3xTF32 (a_lo*w_lo dropped: 2^-22 per product) + chunked accumulation (tc_gemm.cu):
  # the tensor core's truncating accumulate stays below ~1.5e-6 per layer.
  assert_close(C, ref, 1e-5, "tc gemm E%d M%d K%d N%d" % (E, M, K, N))


@pytest.mark.parametrize("env", [{"OAT_TC_TS": "1"}, {"OAT_TC_STACK_K": "0"}, {"OAT_TC_WSPLIT": "1"},
                                 {"OAT_TC_DIRECT": "1"}, {"OAT_TC_DIRECT": "0", "OAT_TC_BN_SHALLOW": "0"}],
                         ids=lambda e: ",".join("%s=%s" % kv for kv in sorted(e.items())))
def test_evaluated_gemm_variants_stay_correct(env):
  """The GEMM's opt-in forms (tensor-memory A operand, stacked [W_hi;W_lo] two-MMA form, in-SM
  weight split, direct-store epilogue everywhere / nowhere) are selected by environment variables
  read once per process: every (K, N) family of this file is re-run against float64 in a
  subprocess per variant, so the code paths DESIGN.md §12 reports on cannot rot."""
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-k", "not variants", __file__],
                     env=dict(os.environ, **env), cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                     timeout=600)
  assert r.returncode == 0, r.stdout.decode()[-2000:]
