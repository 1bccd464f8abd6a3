// Pointwise-convolution GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA),
// FP32-accurate through 3xTF32 error compensation.  sm_100a only.
//
//   C[e][m][n] = act( sum_k A[e][m][k] * W[e][n][k] + bias[e][n] ) (+ R[e][m][n])
//
// Why 3xTF32: the parity bar is 1e-4 relative on z through 53 stacked layers; a
// single TF32 (or BF16) pass gives 2e-2 on z, BF16x3 sits exactly on the bar, the
// split  a = a_hi + a_lo (both exactly TF32-representable)  with
//   D += a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
// gives 9e-6 (measured by emulation against the fp32 oracle).
//
// Accumulation: tcgen05.mma truncates its fp32 accumulate (measured on B200: a bias
// of -2^-24 per MMA, profiles/r1_exp_tcgen05_tf32_accumulate_truncation.txt), i.e.
// -(K/8)*3*6e-8 if all three products go into one accumulator (2e-5 at K=960).  So
// (1) the correction products accumulate in their own TMEM columns (2^-11 times
// smaller, their truncation is invisible) and (2) the main product a_hi*w_hi is spread
// round-robin over `C` accumulators by k-block (C = ceil(K/512)); the epilogue adds
// the partial tiles in round-to-nearest fp32.  Error -> (K/8)*6e-8/C <= 4e-6.
//
// Round 2 (measured with tools/tc_trace.py, profiles/r2_gemm_pipeline_trace.txt):
//  * behind `if (lane == 0)` ptxas wrapped EVERY tcgen05.mma in an ELECT / BRA.U.ANY serialisation
//    loop; the issuer warp now runs converged and elects one lane with elect.sync;
//  * the k-block period of the main loop equals the shared-memory traffic of a stage (TMA writes
//    + split read/writes + the MMAs' operand reads: 4 KB of A and 32*N B of W per K=8 slice)
//    divided by 128 B/clk for BN <= 128, and the tensor-pipe time (12 MMAs x 81 cycles) at
//    BN = 160 -> tiles as wide as TMEM allows: C = ceil(K/512) instead of ceil(K/200);
//  * per-layer tile policy (tc_pw_gemm() below): deep K (project layers, K >= 192) store straight
//    from registers (no staging barriers, 16-column granules); shallow K (expand layers, a few
//    k-blocks per tile, store-bound) keep the swizzled staging + TMA store and take BN <= 128 so
//    that TWO accumulator groups fit TMEM and the epilogue of tile i overlaps tile i+1;
//  * evaluated and left OFF (environment switches, each covered by tests/test_gpu_tc_gemm.py):
//    OAT_TC_STACK_K (3xTF32 in two MMAs: a_hi x stacked [W_hi ; W_lo] with N = 2*BN, + a_lo x W_hi),
//    OAT_TC_TS (A operand in tensor memory, written by the splitters with tcgen05.st),
//    OAT_TC_WSPLIT (unsplit weights streamed from L2 and split in shared memory),
//    and the depthwise 3x3 epilogue of TcGemmProblem::dw_out (fusion bit 5).
//
// Structure (one persistent CTA per SM, 448 threads, warp-specialised):
//   (warp numbers for the default of 4 splitter warps, -DOAT_TC_SPLIT_WARPS=n shifts the last two)
//   warp 12  TMA producer: A tile [128 x 32] fp32 + W_hi/W_lo tiles [BN x 32] per
//            k-block, 128B-swizzled, completion on an mbarrier (expect_tx);
//   warps8-11 splitters: turn the raw A tile into a_hi (in place) and a_lo (second
//            buffer) in shared memory — the split is element-wise, so it is
//            layout-agnostic w.r.t. the TMA swizzle; fence.proxy.async hands the
//            tiles to the tensor core;
//   warp 13  MMA issuer: the warp walks the loop converged, one elect.sync lane issues 3
//            tcgen05.mma.kind::tf32 per 8-wide k-slice into TMEM accumulators (128 lanes x BN
//            columns fp32), tcgen05.commit releases smem stages / publishes the accumulator;
//   warps0-7 epilogue (two groups of 4, alternating 32-column slabs): tcgen05.ld 32x32b
//            (one output row per thread), RN sum of the
//            partial accumulators, folded-BN bias, ReLU6, residual; shallow K: 32-column slabs
//            are staged in 128B-swizzled shared memory and written with TMA stores
//            (cp.async.bulk.tensor, double-buffered), which also clip the M/N tails; deep K:
//            float4 stores straight from registers.  When TMEM allows, two accumulator groups
//            overlap the epilogue of tile i with the main loop of tile i+1.
#include <cuda.h>

#include <cstdlib>

#include <map>
#include <mutex>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace oat {
namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // 32 fp32 = 128 B = one swizzle atom row
#ifndef OAT_TC_SPLIT_WARPS
#define OAT_TC_SPLIT_WARPS 4
#endif
constexpr int TC_SPLIT_WARPS = OAT_TC_SPLIT_WARPS;       // warps 8 .. 8+TC_SPLIT_WARPS-1
constexpr int TC_SPLIT_THREADS = 32 * TC_SPLIT_WARPS;
constexpr int TC_TMA_WARP = 8 + TC_SPLIT_WARPS;
constexpr int TC_MMA_WARP = 9 + TC_SPLIT_WARPS;
constexpr int TC_THREADS = 32 * (10 + TC_SPLIT_WARPS);  // 8 epilogue + splitter warps + TMA + MMA
static_assert((TC_BM * TC_BK / 4) % TC_SPLIT_THREADS == 0, "splitters must tile the A tile");
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
constexpr int TC_MAX_STAGES = 5;
constexpr int TC_TMEM_COLS = 512;
#ifndef OAT_TC_BN_SHALLOW_DEFAULT
#define OAT_TC_BN_SHALLOW_DEFAULT 128
#endif
#ifndef OAT_TC_STACK_K_DEFAULT
#define OAT_TC_STACK_K_DEFAULT 1000000   // stacked two-MMA form: measured no faster than three MMAs (off)
#endif
#ifndef OAT_TC_TS_DEFAULT
#define OAT_TC_TS_DEFAULT 0
#endif

#ifdef OAT_TC_TRACE
// Debug build only (tools/tc_trace.py): CTA 0 records clock64() at the hand-over points of the
// TMA -> split -> MMA -> epilogue chain, one row of 8 stamps per k-block, to find what the
// pipeline actually waits for.  [it][0..1] producer (slot free, TMA issued), [2..3] splitter (data
// landed, split done), [4..5] MMA (operands ready, MMAs + commit issued), [6..7] epilogue per tile.
constexpr int kTraceRows = 2048;
__device__ long long g_tc_trace[kTraceRows * 8];
#define OAT_TRACE(row, col)                                                                   \
  do {                                                                                        \
    if (blockIdx.x == 0 && (row) < (uint32_t)kTraceRows) g_tc_trace[(row) * 8 + (col)] = clock64(); \
  } while (0)
#else
#define OAT_TRACE(row, col) do { } while (0)
#endif

struct TcArgs {
  CUtensorMap mapA;   // {K, M, E}, box {32, 128, 1}
  CUtensorMap mapWh;  // {K, N, E}, box {32, BN, 1}
  CUtensorMap mapWl;
  CUtensorMap mapC;   // {N, M, E}, box {32, 128, 1} (store)
  const float* bias;  // [E][N]
  const float* R;     // [E][M][N] or null
  float* C;           // [E][M][N]
  int M, K, N, E;
  int relu6;
  int BN, n_tiles, m_tiles, stages;
  int chunks;  // main accumulators per tile (k-blocks round-robin); +1 correction buffer
  int nbuf;    // 1 or 2 accumulator groups (tile double-buffering when TMEM allows)
  int wres;    // 1: the model's whole W_hi/W_lo stays resident in smem, stages hold only A
  int ts;      // 1: the A operand of the MMAs lives in TENSOR memory (tcgen05.mma "TS" form): the splitter
               // warps write a_hi / a_lo with tcgen05.st instead of st.shared, the MMAs read only W
               // from shared memory.  Measured (tools/tc_trace.py): the SS form is bound by shared-
               // memory bandwidth (TMA writes + split reads/writes + 2 x 4 KB of A per K slice)
  int a_ring;  // ts: number of A slots (64 TMEM columns each: a_hi | a_lo) behind the accumulators
  int acc_cols_per_buf;  // TMEM columns of one accumulator group
  int acc_cols;  // ts: TMEM columns taken by the accumulators (the A ring starts there)
  int stack;   // 1 (deep K): 3xTF32 in two MMAs, a_hi x [W_hi ; W_lo] (N = 2*BN) + a_lo x W_hi, accumulators
               //   C x [hi*hi | hi*lo] + [lo*hi];  0 (shallow K, epilogue-bound layers): three MMAs of width
               //   BN into C x [hi*hi] + [lo*hi + hi*lo] — one accumulator less for the epilogue to read
  int wsplit;  // 1: mapWh addresses the UNSPLIT weights; the splitter warps form W_hi / W_lo in smem
  int direct;  // 1: the epilogue stores straight from registers (no smem staging, no TMA store):
               // frees 64 KB for a third pipeline stage on the stage-starved late layers
  // depthwise-epilogue mode (dw != 0), see TcGemmProblem
  int dw;
  const float* dww[16];
  const float* dwb[16];
  float* D;          // [E][Mout][N]
  int hin, hout, stride;
  int G;             // whole images per M tile (splits == 1)
  int splits;        // 2: an M tile is one half of an image (13x13 inputs)
  int ho0;           // splits == 2: output rows [0, ho0) belong to half 0
  int nimg;          // images per model (B)
  int Mout;          // B * hout * hout
};

// First A row (within the model) of M tile `mt`.
__device__ __forceinline__ int tile_m0(const TcArgs& a, int mt) {
  if (!a.dw) return mt * TC_BM;
  if (a.splits == 1) return mt * a.G * a.hin * a.hin;
  const int img = mt >> 1, half = mt & 1;
  return img * a.hin * a.hin + (half ? (a.ho0 * a.stride - 1) * a.hin : 0);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)),
      "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
          reinterpret_cast<uint64_t>(map)),
      "r"(c0), "r"(c1), "r"(c2), "r"(src)
      : "memory");
}
// One lane of a converged warp (elect.sync): ptxas then knows the guarded region runs on exactly
// one lane and emits the tcgen05 / TMA instructions (uniform-datapath operands) straight; behind
// `if (lane == 0)` it wraps EVERY such instruction in an ELECT / BRA.U.ANY serialisation loop,
// which cost ~80 cycles per tcgen05.mma (measured with tools/tc_trace.py: the MMA issue of a
// k-block took as long as the MMAs themselves should).
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "     elect.sync %%rx|%%px, %2;\n"
      "@%%px mov.s32 %1, 1;\n"
      "     mov.s32 %0, %%rx;\n"
      "}\n"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFF));
  return pred;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in tensor memory (lane = row, one column per K element), B from shared memory.
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// Accumulator layout of one tile (TMEM columns): C chunks [a_hi*w_hi | a_hi*w_lo] of 2*BN columns each
// (one MMA with the stacked B operand [W_hi ; W_lo] writes both halves), then a_lo*w_hi (BN columns).
// Returns in v the 16 columns c..c+15 of  sum_j main_j  +  (sum_j hilo_j + lohi),  main chunks
// added in round-to-nearest fp32 first (see the header on the accumulate truncation).
__device__ __forceinline__ void acc_load16(uint32_t taddr, int c, int C, int BN, int used, int stack,
                                           float (&v)[16]) {
  float u[16];
  if (stack) {
    {
      uint32_t r0[16], r1[16], r2[16];
      tmem_ld16_nowait(taddr + (uint32_t)c, r0);
      tmem_ld16_nowait(taddr + (uint32_t)(BN + c), r1);
      tmem_ld16_nowait(taddr + (uint32_t)(2 * C * BN + c), r2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = __uint_as_float(r0[i]);
        u[i] = __uint_as_float(r1[i]) + __uint_as_float(r2[i]);
      }
    }
    for (int j = 1; j < used; ++j) {
      uint32_t r0[16], r1[16];
      tmem_ld16_nowait(taddr + (uint32_t)(2 * j * BN + c), r0);
      tmem_ld16_nowait(taddr + (uint32_t)(2 * j * BN + BN + c), r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] += __uint_as_float(r0[i]);
        u[i] += __uint_as_float(r1[i]);
      }
    }
  } else {  // C x [hi*hi] then one correction tile
    {
      uint32_t r0[16], r1[16];
      tmem_ld16_nowait(taddr + (uint32_t)c, r0);
      tmem_ld16_nowait(taddr + (uint32_t)(C * BN + c), r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(r0[i]); u[i] = __uint_as_float(r1[i]); }
    }
    for (int j = 1; j < used; ++j) {
      uint32_t r0[16];
      tmem_ld16_nowait(taddr + (uint32_t)(j * BN + c), r0);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(r0[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += u[i];
}

// K-major, 128B-swizzled operand tile (rows of 128 B, 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);  // start address  [0,14)
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // layout type: SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float relu6f(float v) { return fminf(fmaxf(v, 0.0f), 6.0f); }

__global__ void __launch_bounds__(TC_THREADS, 1) tc_pw_gemm_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[4 * TC_MAX_STAGES + 8];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = a.stages;
  const int BN = a.BN;
  const uint32_t w_bytes = (uint32_t)BN * 128u;
  const int k_blocks_all = (a.K + TC_BK - 1) / TC_BK;
  // resident-W mode: [W_hi k-blocks][W_lo k-blocks] first, then A-only stages
  const uint32_t wres_bytes = a.wres ? 2u * (uint32_t)k_blocks_all * w_bytes : 0u;
  const uint32_t a_stage_bytes = a.ts ? (uint32_t)TC_A_BYTES : 2u * TC_A_BYTES;  // ts: raw tile only
  const uint32_t stage_bytes = a_stage_bytes + (a.wres ? 0u : 2u * w_bytes);
  auto sA = [&](int s) { return smem_base + wres_bytes + (uint32_t)s * stage_bytes; };
  auto sAl = [&](int s) { return sA(s) + TC_A_BYTES; };
  auto sWh = [&](int s) { return sA(s) + a_stage_bytes; };
  auto sWl = [&](int s) { return sWh(s) + w_bytes; };
  const uint32_t stage_out = smem_base + wres_bytes + (uint32_t)S * stage_bytes;  // 2 x 16 KB store staging
  auto rWh = [&](int kb) { return smem_base + (uint32_t)kb * w_bytes; };
  auto rWl = [&](int kb) { return smem_base + (uint32_t)(k_blocks_all + kb) * w_bytes; };
  const uint32_t bar_wfull = smem_u32(&bars[4 * TC_MAX_STAGES + 4]);   // resident W landed (per model)
  const uint32_t bar_wfree = smem_u32(&bars[4 * TC_MAX_STAGES + 5]);   // MMAs of the model retired
  auto bar_full = [&](int s) { return smem_u32(&bars[s]); };
  auto bar_split = [&](int s) { return smem_u32(&bars[TC_MAX_STAGES + s]); };
  auto bar_empty = [&](int s) { return smem_u32(&bars[2 * TC_MAX_STAGES + s]); };
  auto bar_tfull = [&](int i) { return smem_u32(&bars[3 * TC_MAX_STAGES + i]); };
  auto bar_tempty = [&](int i) { return smem_u32(&bars[3 * TC_MAX_STAGES + 2 + i]); };
  auto bar_afree = [&](int j) { return smem_u32(&bars[4 * TC_MAX_STAGES + 6 + j]); };  // ts: A slot j drained

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_split(s), TC_SPLIT_THREADS);
      mbar_init(bar_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull(i), 1);
      mbar_init(bar_tempty(i), 256);
    }
    mbar_init(bar_wfull, 1);
    mbar_init(bar_wfree, 1);
    mbar_init(bar_afree(0), 1);
    mbar_init(bar_afree(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_MMA_WARP) {  // TMEM allocation is a warp-wide operation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == TC_TMA_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapWh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapWl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.mapC)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;

  const int k_blocks = (a.K + TC_BK - 1) / TC_BK;
  const int tiles_per_model = a.m_tiles * a.n_tiles;
  const int num_tiles = tiles_per_model * a.E;

  if (warp == TC_TMA_WARP) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0, wgen = 0;
      int ring_s = 0;
      uint32_t ring_ph = 0;
      int cur_e = -1;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int e = t / tiles_per_model, r = t % tiles_per_model;
        const int m0 = tile_m0(a, r / a.n_tiles), n0 = (r % a.n_tiles) * BN;
        if (a.wres && e != cur_e) {  // (re)load the whole weight matrix of model e once
          if (cur_e >= 0) mbar_wait(bar_wfree, (wgen - 1) & 1);  // previous model's MMAs retired
          mbar_expect_tx(bar_wfull, 2u * (uint32_t)k_blocks * w_bytes);
          for (int kb = 0; kb < k_blocks; ++kb) {
            tma_load_3d(rWh(kb), &a.mapWh, bar_wfull, kb * TC_BK, 0, e);
            tma_load_3d(rWl(kb), &a.mapWl, bar_wfull, kb * TC_BK, 0, e);
          }
          cur_e = e;
          ++wgen;
        }
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = ring_s;            // it % S and (it / S) & 1 as running counters
          const uint32_t ph = ring_ph;
          if (++ring_s == S) { ring_s = 0; ring_ph ^= 1u; }
          OAT_TRACE(it, 0);
          mbar_wait(bar_empty(s), ph ^ 1);
          OAT_TRACE(it, 1);
          if (a.wres) {
            mbar_expect_tx(bar_full(s), TC_A_BYTES);
            tma_load_3d(sA(s), &a.mapA, bar_full(s), kb * TC_BK, m0, e);
          } else if (a.wsplit) {
            mbar_expect_tx(bar_full(s), TC_A_BYTES + w_bytes);
            tma_load_3d(sA(s), &a.mapA, bar_full(s), kb * TC_BK, m0, e);
            tma_load_3d(sWh(s), &a.mapWh, bar_full(s), kb * TC_BK, n0, e);
          } else {
            mbar_expect_tx(bar_full(s), TC_A_BYTES + 2u * w_bytes);
            tma_load_3d(sA(s), &a.mapA, bar_full(s), kb * TC_BK, m0, e);
            tma_load_3d(sWh(s), &a.mapWh, bar_full(s), kb * TC_BK, n0, e);
            tma_load_3d(sWl(s), &a.mapWl, bar_full(s), kb * TC_BK, n0, e);
          }
        }
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (uniform control flow); one elected lane issues.
    {
      // kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = BN
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      // N = 2*BN: the stacked B operand [W_hi ; W_lo] (the two tiles are adjacent in a stage)
      const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);
      uint32_t it = 0, lt = 0, wgen = 0;
      int ring_s = 0, ring_a = 0;
      uint32_t ring_ph = 0;
      int cur_e = -1;
      const int C = a.chunks;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
        if (a.wres) {
          const int e = t / tiles_per_model;
          if (e != cur_e) {
            if (cur_e >= 0 && elect_one_sync()) umma_commit(bar_wfree);  // all MMAs that read the old weights
            __syncwarp();
            mbar_wait(bar_wfull, wgen & 1);
            cur_e = e;
            ++wgen;
          }
        }
        const int ab = (a.nbuf == 2) ? (int)(lt & 1) : 0;
        const uint32_t aph = (a.nbuf == 2) ? ((lt >> 1) & 1) : (lt & 1);
        mbar_wait(bar_tempty(ab), aph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t grp = tmem_base + (uint32_t)(ab * a.acc_cols_per_buf);
        const uint32_t d_lohi = grp + (uint32_t)((a.stack ? 2 * C : C) * BN);  // stack 0: lo*hi + hi*lo
        int kc = 0;  // kb % C
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = ring_s;
          const uint32_t ph = ring_ph;
          if (++ring_s == S) { ring_s = 0; ring_ph ^= 1u; }
          const int ja = ring_a;           // ts: it % a_ring
          if (++ring_a == a.a_ring) ring_a = 0;
          const int kcc = kc;
          if (++kc == C) kc = 0;
          mbar_wait(bar_split(s), ph);
          if (lane == 0) OAT_TRACE(it, 4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t dAh = make_desc_sw128(sA(s)), dAl = make_desc_sw128(sAl(s));
          const uint64_t dWh = make_desc_sw128(a.wres ? rWh(kb) : sWh(s));
          const int kleft = a.K - kb * TC_BK;
          const int slices = kleft >= TC_BK ? 4 : (kleft + 7) / 8;  // zero-filled k-slices skipped
          const uint32_t d_hi = grp + (uint32_t)(kcc * (a.stack ? 2 : 1) * BN);  // [a_hi*w_hi (| a_hi*w_lo)]
          const uint64_t dWl = make_desc_sw128(sWl(s));
          if (a.ts) {
            const uint32_t ta_hi = tmem_base + (uint32_t)(a.acc_cols + ja * 64), ta_lo = ta_hi + 32u;
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                if (ks < slices) {
                  const uint64_t off = (uint64_t)(ks * 2);
                  umma_tf32_ts(d_lohi, ta_lo + (uint32_t)(ks * 8), dWh + off, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
                  umma_tf32_ts(d_hi, ta_hi + (uint32_t)(ks * 8), dWh + off, idesc2, (kb < C && ks == 0) ? 0u : 1u);
                }
              }
              umma_commit(bar_empty(s));   // W tiles (and the raw A tile) of the stage are free
              umma_commit(bar_afree(ja));  // the A slot in tensor memory may be overwritten
            }
          } else if (!a.stack) {
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                if (ks < slices) {
                  const uint64_t off = (uint64_t)(ks * 2);
                  umma_tf32(d_lohi, dAl + off, dWh + off, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
                  umma_tf32(d_lohi, dAh + off, dWl + off, idesc, 1u);
                  umma_tf32(d_hi, dAh + off, dWh + off, idesc, (kb < C && ks == 0) ? 0u : 1u);
                }
              }
              umma_commit(bar_empty(s));
            }
          } else if (elect_one_sync()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < slices) {
                const uint64_t off = (uint64_t)(ks * 2);  // 8 tf32 = 32 B = 2 x 16 B along K
                // 3xTF32 in TWO instructions: a_hi meets the stacked [W_hi ; W_lo] once (the MMAs are
                // bound by their shared-memory operand reads at these N: one read of a_hi less)
                umma_tf32(d_lohi, dAl + off, dWh + off, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
                umma_tf32(d_hi, dAh + off, dWh + off, idesc2, (kb < C && ks == 0) ? 0u : 1u);
              }
            }
            umma_commit(bar_empty(s));  // smem stage reusable once these MMAs retire
          }
          __syncwarp();
          if (lane == 0) OAT_TRACE(it, 5);
        }
        if (elect_one_sync()) umma_commit(bar_tfull(ab));   // accumulators complete
        __syncwarp();
      }
    }
  } else if (warp >= 8) {
    // ===================== splitters =====================
    const int st = threadIdx.x - 256;
    uint32_t it = 0;
    int ring_s = 0, ring_a = 0;
    uint32_t ring_ph = 0, ring_aph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = ring_s;
        const uint32_t ph = ring_ph;
        if (++ring_s == S) { ring_s = 0; ring_ph ^= 1u; }
        mbar_wait(bar_full(s), ph);
        if (st == 0) OAT_TRACE(it, 2);
        if (a.ts) {
          // thread = row of the tile = TMEM lane (warp 8+q owns lanes 32q..32q+31): read the row's
          // 128 B from the swizzled raw tile, write a_hi | a_lo into the A slot with tcgen05.st
          const int ja = ring_a;
          const uint32_t aph = ring_aph;   // parity of the slot's previous use
          if (++ring_a == a.a_ring) { ring_a = 0; ring_aph ^= 1u; }
          mbar_wait(bar_afree(ja), aph ^ 1u);  // MMAs that read the slot's previous contents retired
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t rowaddr = sA(s) + (uint32_t)st * 128u;
          const uint32_t tdst = tmem_base + ((uint32_t)((st >> 5) * 32) << 16) + (uint32_t)(a.acc_cols + ja * 64);
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // 16 K elements (64 B) at a time
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const uint32_t chunk = (uint32_t)(h * 4 + cc);
              uint32_t x, y, z, w;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(x), "=r"(y), "=r"(z), "=r"(w)
                           : "r"(rowaddr + ((chunk ^ ((uint32_t)st & 7u)) << 4)));
              const uint32_t v4[4] = {x, y, z, w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t hb = v4[q] & 0xffffe000u;
                hi[cc * 4 + q] = hb;
                lo[cc * 4 + q] = __float_as_uint(__uint_as_float(v4[q]) - __uint_as_float(hb)) & 0xffffe000u;
              }
            }
            tmem_st16(tdst + (uint32_t)(h * 16), hi);
            tmem_st16(tdst + 32u + (uint32_t)(h * 16), lo);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(bar_split(s));
          if (st == 0) OAT_TRACE(it, 3);
          continue;
        }
        const uint32_t pa = sA(s), pl = sAl(s);
        // All loads of a pass are issued before the first use: with one load -> split -> store
        // chain per 16 bytes (round 1) every chunk paid the full shared-memory latency and the
        // split of a 16 KB tile took ~970 cycles, more than its MMAs (tools/tc_trace.py).
        auto split_pass = [](uint32_t hi_base, uint32_t lo_base, int first, int st_) {
          constexpr int NB = 8;
          uint32_t v[NB][4];
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const uint32_t off = (uint32_t)(st_ + (first + i) * TC_SPLIT_THREADS) * 16u;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3])
                         : "r"(hi_base + off));
          }
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const uint32_t off = (uint32_t)(st_ + (first + i) * TC_SPLIT_THREADS) * 16u;
            uint32_t h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              h[q] = v[i][q] & 0xffffe000u;
              l[q] = __float_as_uint(__uint_as_float(v[i][q]) - __uint_as_float(h[q])) & 0xffffe000u;
            }
#if !defined(OAT_TC_NO_HI_STORE)
            // (-DOAT_TC_NO_HI_STORE: experiment that leaves the raw fp32 tile as the hi operand,
            // relying on kind::tf32 ignoring the 13 low mantissa bits)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(hi_base + off), "r"(h[0]), "r"(h[1]),
                         "r"(h[2]), "r"(h[3])
                         : "memory");
#endif
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(lo_base + off), "r"(l[0]), "r"(l[1]),
                         "r"(l[2]), "r"(l[3])
                         : "memory");
          }
        };
        static_assert(TC_A_BYTES / 16 / TC_SPLIT_THREADS == 8, "one pass of 8 chunks per thread covers the A tile");
        split_pass(pa, pl, 0, st);
        if (a.wsplit) {  // the weight tile [BN rows x 128 B], BN % 32 == 0: BN/16 chunks per thread
          const uint32_t ph_ = sWh(s), pl_ = sWl(s);
          const int chunks_w = (int)(w_bytes / 16u) / TC_SPLIT_THREADS;
          int done = 0;
          for (; done + 8 <= chunks_w; done += 8) split_pass(ph_, pl_, done, st);
          for (; done < chunks_w; ++done) {  // tail: one chunk at a time
            const uint32_t off = (uint32_t)(st + done * TC_SPLIT_THREADS) * 16u;
            uint32_t x, y, z, w;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(ph_ + off));
            const uint32_t hx = x & 0xffffe000u, hy = y & 0xffffe000u, hz = z & 0xffffe000u, hw = w & 0xffffe000u;
            const uint32_t lx = __float_as_uint(__uint_as_float(x) - __uint_as_float(hx)) & 0xffffe000u;
            const uint32_t ly = __float_as_uint(__uint_as_float(y) - __uint_as_float(hy)) & 0xffffe000u;
            const uint32_t lz = __float_as_uint(__uint_as_float(z) - __uint_as_float(hz)) & 0xffffe000u;
            const uint32_t lw = __float_as_uint(__uint_as_float(w) - __uint_as_float(hw)) & 0xffffe000u;
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ph_ + off), "r"(hx), "r"(hy), "r"(hz), "r"(hw) : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(pl_ + off), "r"(lx), "r"(ly), "r"(lz), "r"(lw) : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic -> async proxy
        mbar_arrive(bar_split(s));
        if (st == 0) OAT_TRACE(it, 3);
      }
    }
  } else {
    // ===== epilogue (warps 0-7: group = warp/4 takes slabs of its parity, TMEM lanes 32*(warp%4)..) =====
    uint32_t lt = 0;
    const int C = a.chunks;
    const int used = k_blocks < C ? k_blocks : C;  // main accumulators actually written
    const int grp = warp >> 2, qd = warp & 3;      // slab parity, TMEM lane quadrant
    const int et = threadIdx.x & 127;              // thread index inside the group
    // two staging tiles per group (ping-pong): the TMA store of slab i drains while slab i+1
    // is being assembled; measured before: with one tile the store drain (~1 us) serialised
    // every slab and bounded all memory-bound layers.
    const uint32_t sbuf0 = stage_out + (uint32_t)(2 * grp) * (uint32_t)TC_A_BYTES;
    uint32_t slab_it = 0;
    const int bar_id = 2 + grp;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
      const int e = t / tiles_per_model, r = t % tiles_per_model;
      const int mt = r / a.n_tiles;
      const int m0 = tile_m0(a, mt), n0 = (r % a.n_tiles) * BN;
      const int ab = (a.nbuf == 2) ? (int)(lt & 1) : 0;
      const uint32_t aph = (a.nbuf == 2) ? ((lt >> 1) & 1) : (lt & 1);
      mbar_wait(bar_tfull(ab), aph);
      if (threadIdx.x == 0) OAT_TRACE(lt, 6);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (a.dw) {
        // ===== fused depthwise 3x3: slab -> smem (expand output) -> 3x3 window -> global =====
        // geometry of this M tile: images [img0, img0+nim), input rows from iy0, output rows [oy0, oy1)
        int img0, nim, iy0, oy0, oy1;
        if (a.splits == 1) {
          img0 = mt * a.G;
          nim = a.nimg - img0 < a.G ? a.nimg - img0 : a.G;
          iy0 = 0; oy0 = 0; oy1 = a.hout;
        } else {
          img0 = mt >> 1; nim = 1;
          if (mt & 1) { iy0 = a.ho0 * a.stride - 1; oy0 = a.ho0; oy1 = a.hout; }
          else { iy0 = 0; oy0 = 0; oy1 = a.ho0; }
        }
        const int hin = a.hin, hout = a.hout, st = a.stride;
        const int in_img = (a.splits == 1) ? hin * hin : 0;   // tile rows between images
        const int opi = (oy1 - oy0) * hout;                   // output pixels per image in this tile
        const int P = nim * opi;
        const int rrow_i = qd * 32 + lane;                    // row inside the tile = TMEM lane
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(ab * a.acc_cols_per_buf);
        const float* __restrict__ bias = a.bias + (int64_t)e * a.N;
        const int c4 = et & 7;                                // float4 column of this thread in a slab
        const int64_t out_base = ((int64_t)e * a.Mout + (int64_t)img0 * hout * hout + (int64_t)oy0 * hout) * a.N;
        const float inv_opi = 1.0f / (float)opi, inv_hout = 1.0f / (float)hout;
        // N % 32 == 0: slabs are all-or-nothing, and the ones beyond N come last.  `slab_it` only
        // counts slabs that are processed, so the two staging tiles strictly alternate.
        for (int c0 = grp * 32; c0 < BN && n0 + c0 < a.N; c0 += 64, ++slab_it) {
          const uint32_t sbuf = sbuf0 + (slab_it & 1) * (uint32_t)TC_A_BYTES;
          const int nch = n0 + c0 + 4 * c4;
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            const int c = c0 + hc * 16;
            const int n = n0 + c;
            float4 bq[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(bias + n + 4 * j));
            float v[16];
            acc_load16(taddr, c, C, BN, used, a.stack, v);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b = bq[j >> 2];
              const float4 o = make_float4(relu6f(v[j] + b.x), relu6f(v[j + 1] + b.y), relu6f(v[j + 2] + b.z),
                                           relu6f(v[j + 3] + b.w));
              const uint32_t chunk = (uint32_t)(hc * 4 + (j >> 2));
              const uint32_t dst = sbuf + (uint32_t)rrow_i * 128u + ((chunk ^ (uint32_t)(rrow_i & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
            }
          }
          // one barrier per slab: the two staging tiles alternate, and a thread can only refill
          // a tile after passing the NEXT slab's barrier, i.e. after every reader has left it
          // depthwise taps of this thread's 4 channels: in flight across the barrier
          float4 kw[9];
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) kw[tp] = __ldg(reinterpret_cast<const float4*>(a.dww[e] + (int64_t)tp * a.N + nch));
          const float4 kb = __ldg(reinterpret_cast<const float4*>(a.dwb[e] + nch));
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          float* __restrict__ dst_g = a.D + out_base + nch;
          for (int o = et >> 3; o < P; o += 16) {
            // exact for these small integers: (n + 0.5) / d truncated
            const int g = __float2int_rz(((float)o + 0.5f) * inv_opi), rem = o - g * opi;
            const int oyr = __float2int_rz(((float)rem + 0.5f) * inv_hout);
            const int oy = oy0 + oyr, ox = rem - oyr * hout;
            const int rbase = g * in_img - iy0 * hin;          // tile row of (image g, iy = 0, ix = 0)
            float4 acc = kb;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int iy = oy * st - 1 + dy;
              if (iy < 0 || iy >= hin) continue;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const int ix = ox * st - 1 + dx;
                if (ix < 0 || ix >= hin) continue;
                const uint32_t rr = (uint32_t)(rbase + iy * hin + ix);
                float4 x;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                             : "r"(sbuf + rr * 128u + (((uint32_t)c4 ^ (rr & 7u)) << 4)));
                const float4 k = kw[dy * 3 + dx];
                acc.x = fmaf(x.x, k.x, acc.x); acc.y = fmaf(x.y, k.y, acc.y);
                acc.z = fmaf(x.z, k.z, acc.z); acc.w = fmaf(x.w, k.w, acc.w);
              }
            }
            acc.x = relu6f(acc.x); acc.y = relu6f(acc.y); acc.z = relu6f(acc.z); acc.w = relu6f(acc.w);
            // output pixel index within the tile's output range: image g, row (oy - oy0), column ox
            *reinterpret_cast<float4*>(dst_g + ((int64_t)g * hout * hout + rem) * a.N) = acc;
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar_tempty(ab));
      if (threadIdx.x == 0) OAT_TRACE(lt, 7);
        continue;
      }
      const int rrow_i = qd * 32 + lane;           // row inside the tile = TMEM lane
      const int row = m0 + rrow_i;
      const bool row_ok = row < a.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(ab * a.acc_cols_per_buf);
      const float* __restrict__ bias = a.bias + (int64_t)e * a.N;
      const float* __restrict__ rrow = (a.R && row_ok) ? a.R + ((int64_t)e * a.M + row) * a.N : nullptr;
      if (a.direct) {
        // registers -> global: one output row per thread, 64 contiguous bytes per 16-column chunk
        float* __restrict__ crow = a.C + ((int64_t)e * a.M + row) * a.N;
        for (int c = grp * 16; c < BN; c += 32) {
          const int n = n0 + c;
          if (n >= a.N) break;
          float4 bq[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            bq[j] = (n + 4 * j < a.N) ? __ldg(reinterpret_cast<const float4*>(bias + n + 4 * j))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
          float v[16];
            acc_load16(taddr, c, C, BN, used, a.stack, v);
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              if (n + j >= a.N) break;
              const float4 b = bq[j >> 2];
              float4 o = make_float4(v[j] + b.x, v[j + 1] + b.y, v[j + 2] + b.z, v[j + 3] + b.w);
              if (a.relu6) { o.x = relu6f(o.x); o.y = relu6f(o.y); o.z = relu6f(o.z); o.w = relu6f(o.w); }
              if (rrow) {
                const float4 rr = __ldg(reinterpret_cast<const float4*>(rrow + n + j));
                o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
              }
              *reinterpret_cast<float4*>(crow + n + j) = o;
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar_tempty(ab));
      if (threadIdx.x == 0) OAT_TRACE(lt, 7);
        continue;
      }
      for (int c0 = grp * 32; c0 < BN; c0 += 64, ++slab_it) {
        const uint32_t sbuf = sbuf0 + (slab_it & 1) * (uint32_t)TC_A_BYTES;
        // the TMA store that read this staging tile two slabs ago must have drained
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          const int c = c0 + hc * 16;
          // main[0] and the correction tile are fetched together (one wait), bias in flight
          const int n = n0 + c;
          float4 bq[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            bq[j] = (n + 4 * j < a.N) ? __ldg(reinterpret_cast<const float4*>(bias + n + 4 * j))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
          float v[16];
            acc_load16(taddr, c, C, BN, used, a.stack, v);
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (n + j < a.N) {  // N is a multiple of 4: float4 groups are all-or-nothing
              const float4 b = bq[j >> 2];
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              if (a.relu6) { o.x = relu6f(o.x); o.y = relu6f(o.y); o.z = relu6f(o.z); o.w = relu6f(o.w); }
              if (rrow) {
                const float4 rr = __ldg(reinterpret_cast<const float4*>(rrow + n + j));
                o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
              }
            }
            // staging tile [128 rows][32 floats], 128B swizzle: 16B chunk index ^= row % 8
            const uint32_t chunk = (uint32_t)(hc * 4 + (j >> 2));
            const uint32_t dst = sbuf + (uint32_t)rrow_i * 128u + ((chunk ^ (uint32_t)(rrow_i & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o.x), "f"(o.y),
                         "f"(o.z), "f"(o.w)
                         : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (et == 0) {
          if (n0 + c0 < a.N) tma_store_3d(&a.mapC, sbuf, n0 + c0, m0, e);
          // one bulk group per slab, even when the slab lies beyond N and nothing is stored:
          // `wait_group.read 1` above counts groups, so the ping-pong stays in step.
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_tempty(ab));
      if (threadIdx.x == 0) OAT_TRACE(lt, 7);
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TC_TMEM_COLS)
                 : "memory");
  }
}

// ---- w -> (w_hi, w_lo): both exactly TF32-representable, w_hi + w_lo = w to 2^-22 ----
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi,
                                  float* __restrict__ lo, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t x = __float_as_uint(w[i]);
  const uint32_t h = x & 0xffffe000u;
  hi[i] = __uint_as_float(h);
  lo[i] = __uint_as_float(__float_as_uint(w[i] - __uint_as_float(h)) & 0xffffe000u);
}

// folded [K][N] weights -> TF32-split [N][K] (K-major B operand of the UMMA)
__global__ void pack_split_kernel(const float* __restrict__ w_kn, int K, int N,
                                  float* __restrict__ hi_nk, float* __restrict__ lo_nk,
                                  float* __restrict__ raw_nk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  const int n = idx / K, k = idx % K;
  const float w = w_kn[(int64_t)k * N + n];
  const uint32_t h = __float_as_uint(w) & 0xffffe000u;
  hi_nk[idx] = __uint_as_float(h);
  lo_nk[idx] = __uint_as_float(__float_as_uint(w - __uint_as_float(h)) & 0xffffe000u);
  if (raw_nk) raw_nk[idx] = w;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 3-D fp32 tensor {K (contiguous), rows, E}, box {32, box_rows, 1}, 128B swizzle, zero OOB fill.
int make_map(CUtensorMap* map, const float* base, int64_t K, int64_t rows, int64_t E,
             int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)E};
  cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)K * 4 * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return 0;
}

}  // namespace

#ifdef OAT_TC_TRACE
extern "C" __attribute__((visibility("default"))) int oat_debug_tc_trace_read(long long* out, int rows) {
  if (rows > kTraceRows) rows = kTraceRows;
  return cudaMemcpyFromSymbol(out, g_tc_trace, (size_t)rows * 8 * sizeof(long long)) == cudaSuccess ? 0 : 1;
}
extern "C" __attribute__((visibility("default"))) int oat_debug_tc_trace_clear() {
  void* p = nullptr;
  if (cudaGetSymbolAddress(&p, g_tc_trace) != cudaSuccess) return 1;
  return cudaMemset(p, 0, sizeof(long long) * kTraceRows * 8) == cudaSuccess ? 0 : 1;
}
#endif

int tc_split_weights(const float* w, float* hi, float* lo, int64_t n, cudaStream_t stream) {
  if (n <= 0) return 0;
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, hi, lo, n);
  OAT_LAUNCH_CHECK();
  return 0;
}

int tc_pack_weights(const float* w_kn, int K, int N, float* hi_nk, float* lo_nk, float* raw_nk,
                    cudaStream_t stream) {
  pack_split_kernel<<<(K * N + 255) / 256, 256, 0, stream>>>(w_kn, K, N, hi_nk, lo_nk, raw_nk);
  OAT_LAUNCH_CHECK();
  return 0;
}

bool tc_dw_epilogue_supported(int hin, int stride, int N) {
  if (N % 32 != 0 || (stride != 1 && stride != 2)) return false;
  if (hin * hin <= TC_BM) return hin >= 2;
  // two halves: half 0 needs input rows [0, (ho0-1)*s+2), half 1 rows [ho0*s-1, hin)
  const int hout = (hin + stride - 1) / stride, ho0 = (hout + 1) / 2;
  int r0 = (ho0 - 1) * stride + 2;
  if (r0 > hin) r0 = hin;
  const int r1 = hin - (ho0 * stride - 1);
  return ho0 < hout && r0 * hin <= TC_BM && r1 * hin <= TC_BM;
}

int tc_pw_gemm(const TcGemmProblem& p, cudaStream_t stream) {
  if (p.M <= 0 || p.E <= 0) return 0;
  if (p.K % 4 != 0 || p.N % 4 != 0) return fail("tc_pw_gemm: K and N must be multiples of 4");
  TcArgs a;
  a.bias = p.bias; a.R = p.R; a.C = p.C;
  a.M = p.M; a.K = p.K; a.N = p.N; a.E = p.E; a.relu6 = p.relu6;
  // main-product accumulators: k-blocks go round-robin over C of them, which keeps the truncation
  // bias of the tensor core's fp32 accumulate, (K/8)*2^-24/C per output, below 4e-6
  a.chunks = (p.K + 511) / 512;
  if (a.chunks > 3) a.chunks = 3;
  static const int stack_k = []() { const char* e = getenv("OAT_TC_STACK_K"); return e ? atoi(e) : OAT_TC_STACK_K_DEFAULT; }();
  static const int ts_env = []() { const char* e = getenv("OAT_TC_TS"); return e ? atoi(e) : -1; }();
  a.ts = (ts_env >= 0 ? ts_env != 0 : OAT_TC_TS_DEFAULT) ? 1 : 0;
  a.stack = (p.K >= stack_k || a.ts) ? 1 : 0;  // the tensor-memory A operand is written for the stacked form
  const int sets = a.stack ? 2 * a.chunks + 1 : a.chunks + 1;
  a.a_ring = 2;
  const int tmem_for_acc = TC_TMEM_COLS - (a.ts ? a.a_ring * 64 : 0);
  int bn_max = ((tmem_for_acc / sets) / 32) * 32;      // (2C+1)*BN (+ A ring) <= 512 TMEM columns
#ifndef OAT_TC_BN_CAP
#define OAT_TC_BN_CAP 128                              // the stacked MMA has N = 2*BN <= 256
#endif
  if (bn_max > (a.stack ? OAT_TC_BN_CAP : 160)) bn_max = a.stack ? OAT_TC_BN_CAP : 160;
  // Shallow-K layers (the expand convolutions: 1-5 k-blocks per tile) spend most of a tile in the
  // epilogue; tiles narrow enough for TWO accumulator groups (2 * 3 * BN <= 512) let the epilogue
  // of tile i overlap the main loop of tile i+1.  OAT_TC_BN_SHALLOW=<cols> overrides (0: off).
  static const int shallow_env = []() { const char* e = getenv("OAT_TC_BN_SHALLOW"); return e ? atoi(e) : -1; }();
  const int shallow = shallow_env >= 0 ? shallow_env : OAT_TC_BN_SHALLOW_DEFAULT;
  static const int shallow_k = []() { const char* e = getenv("OAT_TC_BN_SHALLOW_K"); return e ? atoi(e) : 160; }();
  if (shallow > 0 && p.K <= shallow_k && a.chunks == 1 && bn_max > shallow) bn_max = shallow;
  // direct-store epilogue for the deep-K layers (the project convolutions, K >= 192: few output
  // bytes per flop): 16-column granules, so N = 160 / 320 tile without waste (2 x 80, 4 x 80), no
  // staging barriers.  The shallow-K layers (expand convolutions: output-store bound) keep the
  // swizzled staging + TMA store, which measured 25 % faster there (profiles/r2_gemm_ab.md).
  static const int direct_env = []() { const char* e = getenv("OAT_TC_DIRECT"); return e ? atoi(e) : -1; }();
  static const int direct_k = []() { const char* e = getenv("OAT_TC_DIRECT_K"); return e ? atoi(e) : 192; }();
  a.direct = (p.dw_out == nullptr && (direct_env >= 0 ? direct_env != 0 : p.K >= direct_k)) ? 1 : 0;
  const int gran = a.direct ? 16 : 32;
  a.n_tiles = (p.N + bn_max - 1) / bn_max;
  const int per = (p.N + a.n_tiles - 1) / a.n_tiles;
  a.BN = ((per + gran - 1) / gran) * gran;
  a.nbuf = (2 * sets * a.BN <= tmem_for_acc) ? 2 : 1;
  a.acc_cols_per_buf = sets * a.BN;
  a.acc_cols = a.nbuf * sets * a.BN;
  a.m_tiles = (p.M + TC_BM - 1) / TC_BM;
  a.dw = p.dw_out != nullptr;
  a.D = p.dw_out; a.hin = p.hin; a.hout = p.hout; a.stride = p.stride; a.nimg = p.B;
  a.G = 1; a.splits = 1; a.ho0 = 0; a.Mout = p.B * p.hout * p.hout;
  for (int i = 0; i < 16; ++i) { a.dww[i] = p.dw_w[i]; a.dwb[i] = p.dw_b[i]; }
  if (a.dw) {
    if (!tc_dw_epilogue_supported(p.hin, p.stride, p.N) || p.M != p.B * p.hin * p.hin || p.E > 16 ||
        p.hout != (p.hin + p.stride - 1) / p.stride || !p.relu6)
      return fail("tc_pw_gemm: unsupported shape for the depthwise epilogue");
    if (p.hin * p.hin <= TC_BM) {
      a.G = TC_BM / (p.hin * p.hin);
      a.m_tiles = (p.B + a.G - 1) / a.G;
    } else {  // 13x13: two halves per image, each with the halo rows its outputs need
      a.splits = 2;
      a.ho0 = (p.hout + 1) / 2;
      a.m_tiles = 2 * p.B;
    }
  }
  const int kb_all = (p.K + TC_BK - 1) / TC_BK;
  const int wres_bytes = 2 * kb_all * a.BN * 128;
  a.wres = 0;
#ifdef OAT_TC_WRES
#error "resident-W mode keeps W_hi and W_lo apart; the stacked [W_hi ; W_lo] operand needs them adjacent"
  // small weight matrices (early, memory-bound layers): keep W resident, stream only A
  if (a.n_tiles == 1 && wres_bytes <= 64 * 1024) a.wres = 1;
#endif
  const int stage_bytes = (a.ts ? 1 : 2) * TC_A_BYTES + (a.wres ? 0 : 2 * a.BN * 128);
  // Deep-K layers are bound by the TMA -> split -> MMA latency chain with only two 72 KB stages in
  // flight: storing straight from registers frees the 64 KB of store staging for a third stage.
  // Shallow-K layers (memory-bound, one or two k-blocks per tile) keep the TMA-store epilogue.
  const int staging = (a.direct ? 0 : 4) * TC_A_BYTES;
  a.stages = (216 * 1024 - staging - (a.wres ? wres_bytes : 0)) / stage_bytes;
  if (a.stages > TC_MAX_STAGES) a.stages = TC_MAX_STAGES;
  if (a.stages < 2) return fail("tc_pw_gemm: tile does not fit in shared memory");
  if (int rc = make_map(&a.mapA, p.A, p.K, p.M, p.E, TC_BM)) return rc;
  static const int wsplit_env = []() { const char* e = getenv("OAT_TC_WSPLIT"); return e ? atoi(e) : -1; }();
  // default off: measured neutral-to-slower (the kernel is bound by shared-memory bandwidth, not by
  // L2 -> SM bytes); not combined with the tensor-memory A operand (its splitter only handles A)
  a.wsplit = (p.Wr != nullptr && !a.wres && !a.ts && wsplit_env > 0) ? 1 : 0;
  if (int rc = make_map(&a.mapWh, a.wsplit ? p.Wr : p.Wh, p.K, p.N, p.E, a.BN)) return rc;
  if (int rc = make_map(&a.mapWl, p.Wl, p.K, p.N, p.E, a.BN)) return rc;
  if (a.dw) a.mapC = a.mapA;  // no TMA store in the depthwise-epilogue mode
  else if (int rc = make_map(&a.mapC, p.C, p.N, p.M, p.E, TC_BM)) return rc;
  const int smem = a.stages * stage_bytes + (a.wres ? wres_bytes : 0) + staging + 1024;
  static int configured[64] = {0};
  int dev = 0;
  OAT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    OAT_CUDA(cudaFuncSetAttribute(tc_pw_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  222 * 1024));
    if (dev >= 0 && dev < 64) configured[dev] = 222 * 1024;
  }
  static int num_sms[64] = {0};
  if (dev >= 0 && dev < 64 && num_sms[dev] == 0)
    OAT_CUDA(cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  const int sms = (dev >= 0 && dev < 64) ? num_sms[dev] : 148;
  const int tiles = a.m_tiles * a.n_tiles * a.E;
  const int grid = tiles < sms ? tiles : sms;
  tc_pw_gemm_kernel<<<grid, TC_THREADS, smem, stream>>>(a);
  OAT_LAUNCHED(a.dw ? "tc_expand_dw" : "tc_pw_gemm");
  return 0;
}

}  // namespace oat
