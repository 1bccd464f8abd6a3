"""Experiment: how does tcgen05 kind::tf32 accumulate?  Inputs are pre-truncated to
TF32 so every product is exact in fp32 and the low-part MMAs add exact zeros; what
remains is the tensor core's fp32 accumulation error over K/8 MMAs.  Prints max /
mean relative error and the signed bias (negative = magnitude shrinks = truncation)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oatomobile_b200 import _native as Nat

dev = torch.device("cuda:0")
def trunc(a):
  return (a.view(torch.int32) & ~0x1fff).view(torch.float32)

for K in (32, 64, 256, 960, 1280):
  for mode in ("tf32-exact inputs", "full fp32 inputs (3xTF32)"):
    g = torch.Generator().manual_seed(K)
    M, N, E = 512, 64, 1
    A = torch.randn(E, M, K, generator=g).abs()      # positive: accumulator grows steadily
    W = torch.randn(E, N, K, generator=g).abs() / K
    if mode.startswith("tf32"):
      A, W = trunc(A), trunc(W)
    bias = torch.zeros(E, N)
    ref = torch.einsum("emk,enk->emn", A.double(), W.double())
    C = torch.empty(E, M, N, device=dev)
    Ad, Wd, bd = A.to(dev), W.to(dev), bias.to(dev)
    Nat.check(Nat.lib().oat_debug_tc_gemm(Ad.data_ptr(), Wd.data_ptr(), bd.data_ptr(), None,
                                          C.data_ptr(), M, K, N, E, 0, Nat.stream_ptr(dev)))
    torch.cuda.synchronize()
    err = (C.cpu().double() - ref) / ref.abs()
    simt = (A.to(dev) @ W.to(dev).transpose(1, 2)).cpu().double()
    err_simt = (simt - ref) / ref.abs()
    print("K=%4d %-28s max %.2e mean|.| %.2e signed mean %+.2e   | fp32 cuBLAS: max %.2e signed %+.2e"
          % (K, mode, err.abs().max(), err.abs().mean(), err.mean(), err_simt.abs().max(), err_simt.mean()))
