"""Per-layer model of the pointwise GEMM family: tensor-pipe time of 3xTF32 vs the
shared-memory-traffic bound of the SS-form main loop, with the tile policy of tc_pw_gemm().

  python tools/gemm_model.py [E] [B]        (default 4 models x 256 images)

Per k-block (32 k of a 128-row tile) the stage moves through shared memory: TMA writes (16 KB of A +
2*BN*128 B of W_hi/W_lo), the split (16 KB read, 32 KB written), the operand reads of the 12 MMAs
(3 x 4 x 4 KB of A, 3 x 4 x BN*32 B of W); the staged epilogue adds 32 KB per 32-column slab.
128 B/clk of shared-memory bandwidth per SM, 148 SMs, 1.9 GHz; tensor pipe: 3 x 0.5*BN cycles per
K=8 slice (TF32 at half the bf16 rate)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
  E = int(sys.argv[1]) if len(sys.argv) > 1 else 4
  B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
  fused_expand = {"features.2", "features.3", "features.4"}
  tot_s = tot_p = 0.0
  print("%-13s %-8s %7s %5s %5s %4s %6s %3s %10s %9s" % ("layer", "kind", "M/model", "K", "N", "BN", "tiles", "kb",
                                                        "smem us", "pipe us"))
  for kind, name, px, K, N in bench.encoder_layers(4):
    if kind not in ("expand", "project", "last", "fc"):
      continue
    if (kind == "expand" and name in fused_expand) or (kind == "project" and name == "features.1"):
      continue  # inside the fused front kernels
    M = px * B
    C = min(math.ceil(K / 512), 3)
    bn_max = min((512 // (C + 1)) // 32 * 32, 160)
    if K <= 160 and C == 1:
      bn_max = min(bn_max, 128)          # two accumulator groups
    direct = K >= 192
    gran = 16 if direct else 32
    nt = math.ceil(N / bn_max)
    BN = math.ceil(math.ceil(N / nt) / gran) * gran
    tiles = math.ceil(M / 128) * nt * E
    kb = math.ceil(K / 32)
    kbytes = (16 + 2 * BN * 128 / 1024) + 48 + (48 + 3 * BN * 32 * 4 / 1024)
    cyc_smem = tiles * kb * kbytes * 1024 / 128 + (0 if direct else tiles * (BN / 32) * 32 * 1024 / 128)
    cyc_pipe = tiles * kb * 4 * 3 * 0.5 * BN
    t_s, t_p = cyc_smem / 148 / 1.9e9 * 1e6, cyc_pipe / 148 / 1.9e9 * 1e6
    tot_s += t_s
    tot_p += t_p
    print("%-13s %-8s %7d %5d %5d %4d %6d %3d %10.1f %9.1f" % (name, kind, M, K, N, BN, tiles, kb, t_s, t_p))
  print("total: shared-memory bound %.0f us, tensor-pipe bound %.0f us (measured family time: bench.py roofline)"
        % (tot_s, tot_p))


main()
