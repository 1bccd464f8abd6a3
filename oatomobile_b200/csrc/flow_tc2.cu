// Autoregressive flow on tcgen05, two row tiles per CTA time-sharing ONE state buffer.
//
// Variant of flow_tc.cu (same math, image and launch interface): a CTA owns 256 rows as two
// 128-row tiles A and B.  The state tile h (hi/lo, 64 KB) is only needed in shared memory
// from the moment a tile publishes h_t until its two GEMMs (head of step t, recurrent of
// step t+1) have read it; during the long gate phase the state lives in registers.  So A and
// B take turns on the SAME 64 KB buffer: while A runs its gate math, B's GEMMs run, and vice
// versa — the tensor pipe and the SIMT/MUFU pipes overlap without extra shared memory.
// (Original header follows.)
//
// Autoregressive flow on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same math as flow.cu (oatomobile/torch/networks/sequence.py:95-216) but the two
// GEMM-shaped pieces of every step run as 3xTF32 tcgen05.mma:
//   gh[128,192] = h[128,64] . W_hh^T      (83 % of the step's flops)
//   a [128, 32] = h'[128,64] . W_1^T      (head layer 0)
// One CTA = 128 rows (UMMA M = 128).  W_hh/W_1 (hi and lo parts, 112 KB) stay in
// shared memory for all T steps in the canonical K-major SWIZZLE_128B layout (the
// image is pre-swizzled on the host at pack time); the GRU state h is re-split into
// hi/lo and written back into the same layout by the gate threads each step, so no
// [N,64..192] intermediate ever leaves the SM.  Accumulators live in TMEM (192 + 32
// columns); thread (row, half) reads its 3 x 32 gate pre-activations with
// tcgen05.ld, applies the GRU non-linearities and the affine flow update.
// Truncation of the tensor core's accumulate is kept at 8 MMAs per accumulator by
// issuing the 16 small correction MMAs (h_lo.W_hi, h_hi.W_lo) before the 8 main ones.
//
// Roles: warps 0-7 compute (warp w: TMEM lanes 32*(w%4).., hidden units 32*(w/4)..),
// warp 8 = TMEM allocator + single-thread MMA issuer.  The recurrent MMA of step t+1
// is issued right after h_t is published, so it overlaps the head/flow update of t.
#include <cstring>

#include "common.cuh"

namespace oat {
namespace {

constexpr int TR = 128;                 // rows per tile (two tiles per CTA)
constexpr int TTHREADS = 544;           // 2 x 8 compute warps + 1 MMA warp
constexpr int kH_BYTES = TR * 128;      // one k-block (32 fp32) of the state tile: 16 KB
constexpr int kW_BYTES = 192 * 128;     // one k-block of W_hh: 24 KB
constexpr int kW1_BYTES = 32 * 128;     // one k-block of W_1: 4 KB
constexpr float kLog2Pi = 1.8378770664093453f;

// shared-memory image offsets (bytes); the weight part mirrors the packed global image
constexpr int OFF_WHI = 0;
constexpr int OFF_WLO = OFF_WHI + 2 * kW_BYTES;
constexpr int OFF_W1HI = OFF_WLO + 2 * kW_BYTES;
constexpr int OFF_W1LO = OFF_W1HI + 2 * kW1_BYTES;
constexpr int OFF_GATE = OFF_W1LO + 2 * kW1_BYTES;   // [64][12] floats
constexpr int OFF_B1 = OFF_GATE + 64 * 12 * 4;       // [32]
constexpr int OFF_W2 = OFF_B1 + 32 * 4;              // [4][32]
constexpr int OFF_B2 = OFF_W2 + 128 * 4;             // [4]
constexpr int kImageBytes = OFF_B2 + 16;             // = kFlowTcFloats * 4
static_assert(kImageBytes == kFlowTcFloats * 4, "flow TC image size mismatch");
constexpr int OFF_HHI = ((kImageBytes + 1023) / 1024) * 1024;
constexpr int OFF_HLO = OFF_HHI + 2 * kH_BYTES;
constexpr int OFF_YPREV = OFF_HLO + 2 * kH_BYTES;    // [128][2]
constexpr int OFF_OHALF = OFF_YPREV + 2 * TR * 2 * 4;  // [256][4] partial head outputs of half 1
constexpr int OFF_IO = OFF_OHALF + 2 * TR * 4 * 4;    // [256][2T]

struct FlowTc2Args {
  PtrTable images;  // per-model pre-swizzled TC weight image
  const float* in;
  const float* z;
  int64_t z_model_stride;
  float* out;
  float* logprob;
  float* logabsdet;
  float* q;
  int64_t out_model_stride;
  const float* goal;
  int G;
  float inv_two_eps2, log_norm;
  int64_t N;
  int T;
  int rows_per_z;
  int skip_model;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
// One lane of a converged warp (elect.sync): behind `if (lane == 0)` ptxas wraps every tcgen05
// instruction (uniform-datapath operands) in an ELECT / BRA.U.ANY serialisation loop — measured in
// the GEMM (tools/tc_trace.py) at ~80 cycles per MMA, which is what the N = 32 head MMAs cost.
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
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
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ float sigmoid_fast(float v) {
  return __fdividef(1.0f, 1.0f + __expf(-v));
}
__device__ __forceinline__ float tanh_fast(float v) {
  return __fdividef(2.0f, 1.0f + __expf(-2.0f * v)) - 1.0f;
}
__device__ __forceinline__ float softplus_ref(float v) {
  return v > 20.0f ? v : log1pf(expf(v));
}
__device__ __forceinline__ uint32_t tf32_hi(float v) { return __float_as_uint(v) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float v, uint32_t hi) {
  return __float_as_uint(v - __uint_as_float(hi)) & 0xffffe000u;
}

// MODE 0: sample (x -> y, also scores); MODE 1: score (y -> x).
template <int MODE>
__global__ void __launch_bounds__(TTHREADS, 1) flow_tc2_kernel(const __grid_constant__ FlowTc2Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  const int model = blockIdx.y;
  if (model == a.skip_model) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const int T = a.T, T2 = 2 * a.T;
  const int64_t row0 = (int64_t)blockIdx.x * (2 * TR);
  const int rows_here = (int)min((int64_t)(2 * TR), a.N - row0);
  // per tile tl: h published (256 arrivals) | recurrent acc ready | head acc ready | GEMM window over
  auto bar_h_of = [&](int tl) { return smem_u32(&bars[4 * tl + 0]); };
  auto bar_d_of = [&](int tl) { return smem_u32(&bars[4 * tl + 1]); };
  auto bar_d2_of = [&](int tl) { return smem_u32(&bars[4 * tl + 2]); };
  auto bar_w_of = [&](int tl) { return smem_u32(&bars[4 * tl + 3]); };

  if (tid == 0) {
    for (int tl = 0; tl < 2; ++tl) {
      mbar_init(bar_h_of(tl), 256);
      mbar_init(bar_d_of(tl), 1);
      mbar_init(bar_d2_of(tl), 1);
      mbar_init(bar_w_of(tl), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- stage the pre-swizzled weight image and the x / y tile -----------------------
  {
    const float4* src = reinterpret_cast<const float4*>(a.images.p[model]);
    float4* dst = reinterpret_cast<float4*>(sptr);
    for (int i = tid; i < kImageBytes / 16; i += TTHREADS) dst[i] = __ldg(src + i);
    float* io = reinterpret_cast<float*>(sptr + OFF_IO);
    const float* in = a.in + row0 * T2;
    const int n = rows_here * T2;
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
    if (vec) {
      for (int i = tid; i < n / 4; i += TTHREADS)
        reinterpret_cast<float4*>(io)[i] = __ldg(reinterpret_cast<const float4*>(in) + i);
    } else {
      for (int i = tid; i < n; i += TTHREADS) io[i] = __ldg(in + i);
    }
    for (int i = n + tid; i < 2 * TR * T2; i += TTHREADS) io[i] = 0.0f;
    float* yprev = reinterpret_cast<float*>(sptr + OFF_YPREV);
    for (int i = tid; i < 2 * TR * 2; i += TTHREADS) yprev[i] = 0.0f;  // y_{-1} = 0
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // weight image -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 16) {
    // ===================== MMA issuer =====================
    // The whole warp walks the event loop (uniform control flow); one elected lane issues.
    {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // weight image -> async proxy
      const uint32_t idesc_hh = (1u << 4) | (2u << 7) | (2u << 10) | ((192u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_hd = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      // issues the 24 MMAs of one GEMM and commits them to `bar` (and to `bar2` if non-zero)
      auto issue = [&](uint32_t d_tmem, uint32_t w_hi, uint32_t w_lo, uint32_t w_kb_bytes,
                       uint32_t idesc, uint32_t bar, uint32_t bar2) {
        if (elect_one_sync()) {
          // corrections first (accumulator still ~2^-11 of its final size), main product last
          uint32_t acc = 0;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t dHh = make_desc_sw128(sbase + OFF_HHI + kb * kH_BYTES);
            const uint64_t dHl = make_desc_sw128(sbase + OFF_HLO + kb * kH_BYTES);
            const uint64_t dWh = make_desc_sw128(w_hi + kb * w_kb_bytes);
            const uint64_t dWl = make_desc_sw128(w_lo + kb * w_kb_bytes);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_tf32(d_tmem, dHl + 2 * ks, dWh + 2 * ks, idesc, acc);
              acc = 1;
              umma_tf32(d_tmem, dHh + 2 * ks, dWl + 2 * ks, idesc, 1u);
            }
          }
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t dHh = make_desc_sw128(sbase + OFF_HHI + kb * kH_BYTES);
            const uint64_t dWh = make_desc_sw128(w_hi + kb * w_kb_bytes);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_tf32(d_tmem, dHh + 2 * ks, dWh + 2 * ks, idesc, 1u);
          }
          umma_commit(bar);
          if (bar2) umma_commit(bar2);
        }
        __syncwarp();
      };
      // Per tile: event 0 = h_0 published -> recurrent GEMM of step 0; event t+1 = h_t
      // published -> head GEMM of step t, then the recurrent GEMM of step t+1.  After the
      // last MMA of an event `bar_w` is committed: the shared state buffer may change hands.
      int ev[2] = {0, 0};
      while (ev[0] <= T || ev[1] <= T) {
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
          if (ev[tl] > T) continue;
          // uniform decision: every lane must have seen the phase complete
          if (!__all_sync(0xffffffffu, mbar_try_wait(bar_h_of(tl), (uint32_t)(ev[tl] & 1)))) continue;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d0 = tmem + (uint32_t)(tl * 256);
          if (ev[tl] == 0) {
            issue(d0, sbase + OFF_WHI, sbase + OFF_WLO, kW_BYTES, idesc_hh, bar_d_of(tl), bar_w_of(tl));
          } else {
            const int t = ev[tl] - 1;
            if (t + 1 < T) {
              issue(d0 + 192, sbase + OFF_W1HI, sbase + OFF_W1LO, kW1_BYTES, idesc_hd, bar_d2_of(tl), 0u);
              issue(d0, sbase + OFF_WHI, sbase + OFF_WLO, kW_BYTES, idesc_hh, bar_d_of(tl), bar_w_of(tl));
            } else {
              issue(d0 + 192, sbase + OFF_W1HI, sbase + OFF_W1LO, kW1_BYTES, idesc_hd, bar_d2_of(tl), bar_w_of(tl));
            }
          }
          ++ev[tl];
        }
      }
    }
  } else {
    // ===================== compute warps =====================
    const int tl = warp >> 3;                        // tile A (0) or B (1)
    const int q = warp & 3, hf = (warp >> 2) & 1;
    const int hrow = q * 32 + lane;                  // row inside the tile = TMEM lane = H row
    const int row = tl * TR + hrow;                  // row inside the CTA (io / yprev index)
    const uint32_t bar_h = bar_h_of(tl), bar_d = bar_d_of(tl), bar_d2 = bar_d2_of(tl);
    const uint32_t bar_w_other = bar_w_of(tl ^ 1);
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * 256);
    const float* gate = reinterpret_cast<const float*>(sptr + OFF_GATE);
    const float* B1 = reinterpret_cast<const float*>(sptr + OFF_B1);
    const float* W2 = reinterpret_cast<const float*>(sptr + OFF_W2);
    const float* B2 = reinterpret_cast<const float*>(sptr + OFF_B2);
    float* yprev = reinterpret_cast<float*>(sptr + OFF_YPREV);
    float* io = reinterpret_cast<float*>(sptr + OFF_IO);
    float* ohalf = reinterpret_cast<float*>(sptr + OFF_OHALF);
    const uint32_t hhi_row = sbase + OFF_HHI + hf * kH_BYTES + hrow * 128;
    const uint32_t hlo_row = sbase + OFF_HLO + hf * kH_BYTES + hrow * 128;
    int pub_event = 0;

    auto publish_h = [&](const float (&h)[32]) {
      // take the shared state buffer: the other tile's GEMMs of its previous turn are done
      if (tl == 1) mbar_wait(bar_w_other, (uint32_t)(pub_event & 1));
      else if (pub_event > 0) mbar_wait(bar_w_other, (uint32_t)((pub_event - 1) & 1));
      ++pub_event;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t off = (uint32_t)((c ^ (hrow & 7)) << 4);  // SWIZZLE_128B: chunk ^= row % 8
        const uint32_t h0 = tf32_hi(h[4 * c]), h1 = tf32_hi(h[4 * c + 1]),
                       h2 = tf32_hi(h[4 * c + 2]), h3 = tf32_hi(h[4 * c + 3]);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(hhi_row + off), "r"(h0),
                     "r"(h1), "r"(h2), "r"(h3)
                     : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(hlo_row + off),
                     "r"(tf32_lo(h[4 * c], h0)), "r"(tf32_lo(h[4 * c + 1], h1)),
                     "r"(tf32_lo(h[4 * c + 2], h2)), "r"(tf32_lo(h[4 * c + 3], h3))
                     : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_h);
    };

    // h_0 = z[row / rows_per_z], this thread's 32 hidden units
    float h[32];
    {
      const float* zb = a.z + (int64_t)model * a.z_model_stride;
      if (row < rows_here) {
        const float4* zr = reinterpret_cast<const float4*>(zb + ((row0 + row) / a.rows_per_z) * kHidden + hf * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = __ldg(zr + c);
          h[4 * c] = v.x; h[4 * c + 1] = v.y; h[4 * c + 2] = v.z; h[4 * c + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) h[i] = 0.0f;
      }
    }
    publish_h(h);

    float sumsq = 0.0f, sumlog = 0.0f, goal_ll = 0.0f;
    for (int t = 0; t < T; ++t) {
      // ---- gates: r|z|n pre-activations from TMEM columns [g*64 + hf*32, +32) -----------
      mbar_wait(bar_d, (uint32_t)(t & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float2 yp = *reinterpret_cast<const float2*>(yprev + 2 * row);
#pragma unroll
      for (int p8 = 0; p8 < 4; ++p8) {  // 8 hidden units per pass keeps the live set small
        uint32_t ar[8], az[8], an[8];
        const uint32_t col = (uint32_t)(hf * 32 + p8 * 8);
        tmem_ld8_nowait(trow + col, ar);
        tmem_ld8_nowait(trow + 64 + col, az);
        tmem_ld8_nowait(trow + 128 + col, an);
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = hf * 32 + p8 * 8 + u;
          const float4 g0 = *reinterpret_cast<const float4*>(gate + j * 12);      // wr0 wr1 br wz0
          const float4 g1 = *reinterpret_cast<const float4*>(gate + j * 12 + 4);  // wz1 bz wn0 wn1
          const float2 g2 = *reinterpret_cast<const float2*>(gate + j * 12 + 8);  // bin bhn
          const float ir = fmaf(g0.y, yp.y, fmaf(g0.x, yp.x, g0.z));
          const float iz = fmaf(g1.x, yp.y, fmaf(g0.w, yp.x, g1.y));
          const float in_ = fmaf(g1.w, yp.y, fmaf(g1.z, yp.x, g2.x));
          const float rr = sigmoid_fast(ir + __uint_as_float(ar[u]));
          const float gg = sigmoid_fast(iz + __uint_as_float(az[u]));
          const float nn = tanh_fast(fmaf(rr, __uint_as_float(an[u]) + g2.y, in_));
          const int hi = p8 * 8 + u;
          h[hi] = fmaf(gg, h[hi] - nn, nn);
        }
      }
      publish_h(h);  // -> head MMA of step t and recurrent MMA of step t+1

      // ---- head layer 1: both halves reduce 16 of the 32 head units each; half 1 hands its
      // partial sums over through shared memory, half 0 finishes the row -------------------
      mbar_wait(bar_d2, (uint32_t)(t & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
      {
        uint32_t d[16];
        tmem_ld16_nowait(trow + 192 + hf * 16, d);
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = hf * 16 + u;
          const float av = fmaxf(__uint_as_float(d[u]) + B1[j], 0.0f);
          o0 = fmaf(av, W2[j], o0);
          o1 = fmaf(av, W2[32 + j], o1);
          o2 = fmaf(av, W2[64 + j], o2);
          o3 = fmaf(av, W2[96 + j], o3);
        }
      }
      if (hf == 1) *reinterpret_cast<float4*>(ohalf + 4 * row) = make_float4(o0, o1, o2, o3);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync %0, 256;" ::"r"(1 + tl) : "memory");  // partial sums visible
      if (hf == 0) {
        {
          const float4 p = *reinterpret_cast<const float4*>(ohalf + 4 * row);
          o0 += p.x + B2[0]; o1 += p.y + B2[1]; o2 += p.z + B2[2]; o3 += p.w + B2[3];
        }
        const float mu0 = yp.x + o0, mu1 = yp.y + o1;
        const float s0 = softplus_ref(o2) + 1e-3f, s1 = softplus_ref(o3) + 1e-3f;
        float2 v = *reinterpret_cast<const float2*>(io + row * T2 + 2 * t);
        float y0, y1, x0, x1;
        if (MODE == 0) {
          y0 = fmaf(s0, v.x, mu0);
          y1 = fmaf(s1, v.y, mu1);
          x0 = (y0 - mu0) / s0;
          x1 = (y1 - mu1) / s1;
          *reinterpret_cast<float2*>(io + row * T2 + 2 * t) = make_float2(y0, y1);
        } else {
          y0 = v.x; y1 = v.y;
          x0 = (y0 - mu0) / s0;
          x1 = (y1 - mu1) / s1;
          if (a.out != nullptr) *reinterpret_cast<float2*>(io + row * T2 + 2 * t) = make_float2(x0, x1);
        }
        sumsq = fmaf(x0, x0, sumsq);
        sumsq = fmaf(x1, x1, sumsq);
        sumlog += logf(s0);
        sumlog += logf(s1);
        *reinterpret_cast<float2*>(yprev + 2 * row) = make_float2(y0, y1);
        if (a.goal != nullptr && t == T - 1 && row < rows_here) {
          const float* g = a.goal + ((row0 + row) / a.rows_per_z) * a.G * 2;
          float mx = -INFINITY;
          for (int i = 0; i < a.G; ++i) {
            const float d0 = y0 - __ldg(g + 2 * i), d1 = y1 - __ldg(g + 2 * i + 1);
            mx = fmaxf(mx, -(d0 * d0 + d1 * d1) * a.inv_two_eps2);
          }
          float se = 0.0f;
          for (int i = 0; i < a.G; ++i) {
            const float d0 = y0 - __ldg(g + 2 * i), d1 = y1 - __ldg(g + 2 * i + 1);
            se += expf(-(d0 * d0 + d1 * d1) * a.inv_two_eps2 - mx);
          }
          goal_ll = mx + logf(se) + a.log_norm;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + tl) : "memory");  // y_t visible to the tile's 8 warps
    }

    if (hf == 0 && row < rows_here) {
      const float lp = -0.5f * sumsq - (float)T * kLog2Pi;
      const int64_t idx = (int64_t)model * a.out_model_stride + row0 + row;
      if (a.logprob != nullptr) a.logprob[idx] = lp;
      if (a.logabsdet != nullptr) a.logabsdet[idx] = sumlog;
      if (a.q != nullptr) a.q[idx] = (lp - sumlog) + goal_ll;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (a.out != nullptr) {
    const float* io = reinterpret_cast<const float*>(sptr + OFF_IO);
    float* dst = a.out + row0 * T2;
    const int n = rows_here * T2;
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    if (vec) {
      for (int i = tid; i < n / 4; i += TTHREADS)
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(io)[i];
    } else {
      for (int i = tid; i < n; i += TTHREADS) dst[i] = io[i];
    }
  }
  if (warp == 16) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)
                 : "memory");
  }
}

template <int MODE>
int launch_tc2_mode(const FlowTc2Args& fa, int num_models, cudaStream_t stream) {
  const size_t smem = (size_t)OFF_IO + (size_t)2 * TR * 2 * fa.T * sizeof(float) + 1024;
  if (smem > 227 * 1024) return fail("flow_tc: T too large for one CTA's shared memory");
  static size_t configured[64] = {0};
  int dev = 0;
  OAT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    OAT_CUDA(cudaFuncSetAttribute(flow_tc2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  dim3 grid((unsigned)((fa.N + 2 * TR - 1) / (2 * TR)), (unsigned)num_models);
  flow_tc2_kernel<MODE><<<grid, TTHREADS, smem, stream>>>(fa);
  OAT_LAUNCHED(MODE == 0 ? "flow_sample" : "flow_score");
  return 0;
}

}  // namespace

int launch_flow_tc2(const FlowLaunch& a, cudaStream_t stream) {
  if (a.N <= 0 || a.T <= 0) return 0;
  if (a.mode != 0 && a.mode != 1) return fail("flow_tc2: only sample/score modes");
  FlowTc2Args fa;
  fa.images = a.weights;
  fa.in = a.in; fa.z = a.z; fa.z_model_stride = a.z_model_stride;
  fa.out = a.out; fa.logprob = a.logprob; fa.logabsdet = a.logabsdet; fa.q = a.q;
  fa.out_model_stride = a.out_model_stride;
  fa.goal = a.goal; fa.G = a.G;
  const double eps = a.epsilon;
  fa.inv_two_eps2 = (float)(1.0 / (2.0 * eps * eps));
  fa.log_norm = a.goal ? (float)(-log(2.0 * 3.14159265358979323846 * eps * eps) - log((double)a.G)) : 0.0f;
  fa.N = a.N; fa.T = a.T; fa.rows_per_z = a.rows_per_z; fa.skip_model = a.skip_model;
  return a.mode == 0 ? launch_tc2_mode<0>(fa, a.num_models, stream)
                     : launch_tc2_mode<1>(fa, a.num_models, stream);
}

}  // namespace oat
