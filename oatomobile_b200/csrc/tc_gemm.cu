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
// (1) the two correction products accumulate in their own TMEM buffer (2^-11 times
// smaller, its truncation is invisible) and (2) the main product a_hi*w_hi is spread
// round-robin over `C` accumulators by k-block (C = ceil(K/200)); the epilogue adds
// the C+1 partial tiles in round-to-nearest fp32.  Error -> (K/8)*6e-8/C <= 1.5e-6.
//
// Structure (one persistent CTA per SM, 448 threads, warp-specialised):
//   (warp numbers for the default of 4 splitter warps, -DOAT_TC_SPLIT_WARPS=n shifts the last two)
//   warp 12  TMA producer: A tile [128 x 32] fp32 + W_hi/W_lo tiles [BN x 32] per
//            k-block, 128B-swizzled, completion on an mbarrier (expect_tx);
//   warps8-11 splitters: turn the raw A tile into a_hi (in place) and a_lo (second
//            buffer) in shared memory — the split is element-wise, so it is
//            layout-agnostic w.r.t. the TMA swizzle; fence.proxy.async hands the
//            tiles to the tensor core;
//   warp 13  MMA issuer: one elected thread issues 3 tcgen05.mma.kind::tf32 per
//            8-wide k-slice into a TMEM accumulator (128 lanes x BN columns fp32),
//            tcgen05.commit releases smem stages / publishes the accumulator;
//   warps0-7 epilogue (two groups of 4, alternating 32-column slabs): tcgen05.ld 32x32b
//            (one output row per thread), RN sum of the
//            partial accumulators, folded-BN bias, ReLU6, residual; 32-column slabs
//            are staged in 128B-swizzled shared memory and written with TMA stores
//            (cp.async.bulk.tensor, double-buffered), which also clip the M/N tails.
//            When TMEM allows, two accumulator groups overlap the epilogue of tile i
//            with the main loop of tile i+1.
#include <cuda.h>

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
};

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
  __shared__ uint64_t bars[4 * TC_MAX_STAGES + 6];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = a.stages;
  const int BN = a.BN;
  const uint32_t w_bytes = (uint32_t)BN * 128u;
  const int k_blocks_all = (a.K + TC_BK - 1) / TC_BK;
  // resident-W mode: [W_hi k-blocks][W_lo k-blocks] first, then A-only stages
  const uint32_t wres_bytes = a.wres ? 2u * (uint32_t)k_blocks_all * w_bytes : 0u;
  const uint32_t stage_bytes = 2u * TC_A_BYTES + (a.wres ? 0u : 2u * w_bytes);
  auto sA = [&](int s) { return smem_base + wres_bytes + (uint32_t)s * stage_bytes; };
  auto sAl = [&](int s) { return sA(s) + TC_A_BYTES; };
  auto sWh = [&](int s) { return sA(s) + 2u * TC_A_BYTES; };
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
      int cur_e = -1;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int e = t / tiles_per_model, r = t % tiles_per_model;
        const int m0 = (r / a.n_tiles) * TC_BM, n0 = (r % a.n_tiles) * BN;
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
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(bar_empty(s), ph ^ 1);
          if (a.wres) {
            mbar_expect_tx(bar_full(s), TC_A_BYTES);
            tma_load_3d(sA(s), &a.mapA, bar_full(s), kb * TC_BK, m0, e);
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
    if (lane == 0) {
      // kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = BN
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      uint32_t it = 0, lt = 0, wgen = 0;
      int cur_e = -1;
      const int C = a.chunks;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
        if (a.wres) {
          const int e = t / tiles_per_model;
          if (e != cur_e) {
            if (cur_e >= 0) umma_commit(bar_wfree);  // all MMAs that read the old weights
            mbar_wait(bar_wfull, wgen & 1);
            cur_e = e;
            ++wgen;
          }
        }
        const int ab = (a.nbuf == 2) ? (int)(lt & 1) : 0;
        const uint32_t aph = (a.nbuf == 2) ? ((lt >> 1) & 1) : (lt & 1);
        mbar_wait(bar_tempty(ab), aph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t grp = tmem_base + (uint32_t)(ab * (C + 1) * BN);
        const uint32_t d_corr = grp + (uint32_t)(C * BN);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(bar_split(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t dAh = make_desc_sw128(sA(s)), dAl = make_desc_sw128(sAl(s));
          const uint64_t dWh = make_desc_sw128(a.wres ? rWh(kb) : sWh(s));
          const uint64_t dWl = make_desc_sw128(a.wres ? rWl(kb) : sWl(s));
          const int kleft = a.K - kb * TC_BK;
          const int slices = kleft >= TC_BK ? 4 : (kleft + 7) / 8;  // zero-filled k-slices skipped
          const uint32_t d_main = grp + (uint32_t)((kb % C) * BN);
          for (int ks = 0; ks < slices; ++ks) {
            const uint64_t off = (uint64_t)(ks * 2);  // 8 tf32 = 32 B = 2 x 16 B along K
            umma_tf32(d_corr, dAl + off, dWh + off, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
            umma_tf32(d_corr, dAh + off, dWl + off, idesc, 1u);
            umma_tf32(d_main, dAh + off, dWh + off, idesc, (kb < C && ks == 0) ? 0u : 1u);
          }
          umma_commit(bar_empty(s));  // smem stage reusable once these MMAs retire
        }
        umma_commit(bar_tfull(ab));   // accumulators complete
      }
    }
  } else if (warp >= 8) {
    // ===================== splitters =====================
    const int st = threadIdx.x - 256;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(bar_full(s), ph);
        const uint32_t pa = sA(s), pl = sAl(s);
#pragma unroll
        for (int i = 0; i < TC_A_BYTES / 16 / TC_SPLIT_THREADS; ++i) {
          const uint32_t off = (uint32_t)(st + i * TC_SPLIT_THREADS) * 16u;
          uint32_t x, y, z, w;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(x), "=r"(y), "=r"(z), "=r"(w)
                       : "r"(pa + off));
          const uint32_t hx = x & 0xffffe000u, hy = y & 0xffffe000u, hz = z & 0xffffe000u,
                         hw = w & 0xffffe000u;
          const uint32_t lx = __float_as_uint(__uint_as_float(x) - __uint_as_float(hx)) & 0xffffe000u;
          const uint32_t ly = __float_as_uint(__uint_as_float(y) - __uint_as_float(hy)) & 0xffffe000u;
          const uint32_t lz = __float_as_uint(__uint_as_float(z) - __uint_as_float(hz)) & 0xffffe000u;
          const uint32_t lw = __float_as_uint(__uint_as_float(w) - __uint_as_float(hw)) & 0xffffe000u;
#if !defined(OAT_TC_NO_HI_STORE)
          // (-DOAT_TC_NO_HI_STORE: experiment that leaves the raw fp32 tile as the hi operand,
          // relying on kind::tf32 ignoring the 13 low mantissa bits)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(pa + off), "r"(hx), "r"(hy),
                       "r"(hz), "r"(hw)
                       : "memory");
#endif
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(pl + off), "r"(lx), "r"(ly),
                       "r"(lz), "r"(lw)
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic -> async proxy
        mbar_arrive(bar_split(s));
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
      const int m0 = (r / a.n_tiles) * TC_BM, n0 = (r % a.n_tiles) * BN;
      const int ab = (a.nbuf == 2) ? (int)(lt & 1) : 0;
      const uint32_t aph = (a.nbuf == 2) ? ((lt >> 1) & 1) : (lt & 1);
      mbar_wait(bar_tfull(ab), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int rrow_i = qd * 32 + lane;           // row inside the tile = TMEM lane
      const int row = m0 + rrow_i;
      const bool row_ok = row < a.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(ab * (C + 1) * BN);
      const float* __restrict__ bias = a.bias + (int64_t)e * a.N;
      const float* __restrict__ rrow = (a.R && row_ok) ? a.R + ((int64_t)e * a.M + row) * a.N : nullptr;
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
          float v[16], u[16];
          {
            uint32_t r0[16], r1[16];
            tmem_ld16_nowait(taddr + (uint32_t)c, r0);  // warp-wide: executed by all 32 lanes
            tmem_ld16_nowait(taddr + (uint32_t)(C * BN + c), r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(r0[i]); u[i] = __uint_as_float(r1[i]); }
          }
          for (int j = 1; j < used; ++j) {    // partial main products, round-to-nearest adds
            float w[16];
            tmem_ld16(taddr + (uint32_t)(j * BN + c), w);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w[i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += u[i];
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
                                  float* __restrict__ hi_nk, float* __restrict__ lo_nk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  const int n = idx / K, k = idx % K;
  const float w = w_kn[(int64_t)k * N + n];
  const uint32_t h = __float_as_uint(w) & 0xffffe000u;
  hi_nk[idx] = __uint_as_float(h);
  lo_nk[idx] = __uint_as_float(__float_as_uint(w - __uint_as_float(h)) & 0xffffe000u);
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

int tc_split_weights(const float* w, float* hi, float* lo, int64_t n, cudaStream_t stream) {
  if (n <= 0) return 0;
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, hi, lo, n);
  OAT_LAUNCH_CHECK();
  return 0;
}

int tc_pack_weights(const float* w_kn, int K, int N, float* hi_nk, float* lo_nk,
                    cudaStream_t stream) {
  pack_split_kernel<<<(K * N + 255) / 256, 256, 0, stream>>>(w_kn, K, N, hi_nk, lo_nk);
  OAT_LAUNCH_CHECK();
  return 0;
}

int tc_pw_gemm(const TcGemmProblem& p, cudaStream_t stream) {
  if (p.M <= 0 || p.E <= 0) return 0;
  if (p.K % 4 != 0 || p.N % 4 != 0) return fail("tc_pw_gemm: K and N must be multiples of 4");
  TcArgs a;
  a.bias = p.bias; a.R = p.R; a.C = p.C;
  a.M = p.M; a.K = p.K; a.N = p.N; a.E = p.E; a.relu6 = p.relu6;
  // main-product accumulators: keep the truncation bias (K/8)*2^-24/C below ~1.5e-6
  a.chunks = (p.K + 199) / 200;
  if (a.chunks > 7) a.chunks = 7;
  const int sets = a.chunks + 1;                       // + correction buffer
  int bn_max = ((TC_TMEM_COLS / sets) / 32) * 32;  // (C+1)*BN <= 512 TMEM columns
#ifndef OAT_TC_BN_CAP
#define OAT_TC_BN_CAP 160
#endif
  if (bn_max > OAT_TC_BN_CAP) bn_max = OAT_TC_BN_CAP;  // >= 2 smem stages + store staging must fit
  a.n_tiles = (p.N + bn_max - 1) / bn_max;
  const int per = (p.N + a.n_tiles - 1) / a.n_tiles;
  a.BN = ((per + 31) / 32) * 32;  // whole 32-column store slabs
  a.nbuf = (2 * sets * a.BN <= TC_TMEM_COLS) ? 2 : 1;
  a.m_tiles = (p.M + TC_BM - 1) / TC_BM;
  const int kb_all = (p.K + TC_BK - 1) / TC_BK;
  const int wres_bytes = 2 * kb_all * a.BN * 128;
  a.wres = 0;
#ifdef OAT_TC_WRES
  // small weight matrices (early, memory-bound layers): keep W resident, stream only A
  if (a.n_tiles == 1 && wres_bytes <= 64 * 1024) a.wres = 1;
#endif
  const int stage_bytes = 2 * TC_A_BYTES + (a.wres ? 0 : 2 * a.BN * 128);
  a.stages = (216 * 1024 - 4 * TC_A_BYTES - (a.wres ? wres_bytes : 0)) / stage_bytes;  // 4 store-staging tiles
  if (a.stages > TC_MAX_STAGES) a.stages = TC_MAX_STAGES;
  if (a.stages < 2) return fail("tc_pw_gemm: tile does not fit in shared memory");
  if (int rc = make_map(&a.mapA, p.A, p.K, p.M, p.E, TC_BM)) return rc;
  if (int rc = make_map(&a.mapWh, p.Wh, p.K, p.N, p.E, a.BN)) return rc;
  if (int rc = make_map(&a.mapWl, p.Wl, p.K, p.N, p.E, a.BN)) return rc;
  if (int rc = make_map(&a.mapC, p.C, p.N, p.M, p.E, TC_BM)) return rc;
  const int smem = a.stages * stage_bytes + (a.wres ? wres_bytes : 0) + 4 * TC_A_BYTES + 1024;
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
  OAT_LAUNCHED("tc_pw_gemm");
  return 0;
}

}  // namespace oat
