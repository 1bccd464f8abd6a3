"""Timeline of the tcgen05 GEMM pipeline (debug build -DOAT_TC_TRACE, CTA 0): where do the
TMA producer, the splitter warps, the MMA issuer and the epilogue actually wait?

  python tools/tc_trace.py [M K N E]      (builds oatomobile_b200/variants/liboat_trace.so)
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "oatomobile_b200", "variants", "liboat_trace.so")

def build():
  from oatomobile_b200 import build as B
  os.makedirs(os.path.dirname(VAR), exist_ok=True)
  B.build(defines=["OAT_TC_TRACE"], out=VAR)

def main():
  if "--build-only" in sys.argv:
    build(); return
  if not os.path.exists(VAR):
    build()
  os.environ["OAT_B200_LIB"] = VAR
  import torch
  from oatomobile_b200 import _native as N
  args = [int(a) for a in sys.argv[1:] if a.isdigit()]
  M, K, Nn, E = args if len(args) == 4 else (4096, 960, 320, 4)
  L = N.lib()
  L.oat_debug_tc_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
  g = torch.Generator().manual_seed(0)
  A = torch.randn(E, M, K, generator=g).cuda()
  W = (torch.randn(E, Nn, K, generator=g) / K ** 0.5).cuda()
  b = torch.randn(E, Nn, generator=g).cuda()
  C = torch.empty(E, M, Nn, device="cuda")
  def run():
    N.check(L.oat_debug_tc_gemm(A.data_ptr(), W.data_ptr(), b.data_ptr(), None, C.data_ptr(), M, K, Nn, E, 0,
                                 N.stream_ptr(A.device)))
  run(); run()
  L.oat_debug_tc_trace_clear()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); run(); e1.record(); torch.cuda.synchronize()
  rows = 2048
  buf = (ctypes.c_longlong * (rows * 8))()
  assert L.oat_debug_tc_trace_read(buf, rows) == 0
  import numpy as np
  t = np.frombuffer(buf, dtype=np.int64).reshape(rows, 8).copy()
  used = int((t[:, 1] != 0).sum())
  t0 = t[0, 0]
  print("GEMM M=%d K=%d N=%d E=%d: %.1f us (incl. weight split kernel); CTA 0 ran %d k-blocks" %
        (M, K, Nn, E, 1e3 * e0.elapsed_time(e1), used))
  kb = (K + 31) // 32
  names = ["P:wait_empty", "P:tma_issue", "S:data_landed", "S:split_done", "M:operands_ready", "M:issued", "E:acc_ready", "E:done"]
  print("cycles relative to the first stamp; one row per k-block (first 2 tiles)")
  print("  it  " + "  ".join("%16s" % n for n in names[:6]))
  for it in range(min(used, 2 * kb + 4)):
    print("%4d  " % it + "  ".join("%16d" % (t[it, c] - t0) for c in range(6)))
  tiles = int((t[:, 6] != 0).sum())
  for lt in range(min(tiles, 6)):
    print("tile %d: acc ready at %d, epilogue done at %d" % (lt, t[lt, 6] - t0, t[lt, 7] - t0))
  d = t[:used]
  if used > 8:
    print("steady state (medians over k-blocks 4..):")
    print("  k-block period (MMA issue to MMA issue)   %8.0f" % np.median(np.diff(d[4:, 5])))
    print("  TMA issue -> data landed                  %8.0f" % np.median(d[4:, 2] - d[4:, 1]))
    print("  data landed -> split done                 %8.0f" % np.median(d[4:, 3] - d[4:, 2]))
    print("  split done -> MMA sees operands           %8.0f" % np.median(d[4:, 4] - d[4:, 3]))
    print("  MMA issue duration (12 MMAs + commit)     %8.0f" % np.median(d[4:, 5] - d[4:, 4]))
    print("  producer wait for a free slot             %8.0f" % np.median(d[4:, 1] - d[4:, 0]))
    print("  TMA issue lead over MMA issue of same it  %8.0f" % np.median(d[4:, 4] - d[4:, 1]))

main()
