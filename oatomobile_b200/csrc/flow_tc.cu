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

constexpr int TR = 128;                 // rows per CTA
constexpr int TTHREADS = 288;           // 8 compute warps + 1 MMA warp
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
constexpr int OFF_IO = OFF_YPREV + TR * 2 * 4;       // [128][2T]

struct FlowTcArgs {
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
__global__ void __launch_bounds__(TTHREADS, 1) flow_tc_kernel(const __grid_constant__ FlowTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[3];
  __shared__ uint32_t tmem_slot;
  const int model = blockIdx.y;
  if (model == a.skip_model) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sptr = smem_raw + (sbase - smem_u32(smem_raw));
  const int T = a.T, T2 = 2 * a.T;
  const int64_t row0 = (int64_t)blockIdx.x * TR;
  const int rows_here = (int)min((int64_t)TR, a.N - row0);
  const uint32_t bar_h = smem_u32(&bars[0]);   // h_t published (256 arrivals)
  const uint32_t bar_d = smem_u32(&bars[1]);   // recurrent accumulator ready (commit)
  const uint32_t bar_d2 = smem_u32(&bars[2]);  // head accumulator ready (commit)

  if (tid == 0) {
    mbar_init(bar_h, 256);
    mbar_init(bar_d, 1);
    mbar_init(bar_d2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_slot)),
                 "r"(256)
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
    for (int i = n + tid; i < TR * T2; i += TTHREADS) io[i] = 0.0f;
    float* yprev = reinterpret_cast<float*>(sptr + OFF_YPREV);
    for (int i = tid; i < TR * 2; i += TTHREADS) yprev[i] = 0.0f;  // y_{-1} = 0
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // weight image -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 8) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // weight image -> async proxy
      const uint32_t idesc_hh = (1u << 4) | (2u << 7) | (2u << 10) | ((192u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_hd = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      auto issue = [&](uint32_t d_tmem, uint32_t w_hi, uint32_t w_lo, uint32_t w_kb_bytes,
                       uint32_t idesc) {
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
      };
      // event e = 0: h_0 = z published;  e = t+1: h_t published
      mbar_wait(bar_h, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue(tmem, sbase + OFF_WHI, sbase + OFF_WLO, kW_BYTES, idesc_hh);
      umma_commit(bar_d);
      for (int t = 0; t < T; ++t) {
        mbar_wait(bar_h, (uint32_t)((t + 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue(tmem + 192, sbase + OFF_W1HI, sbase + OFF_W1LO, kW1_BYTES, idesc_hd);
        umma_commit(bar_d2);
        if (t + 1 < T) {  // next step's recurrent product overlaps this step's head/flow update
          issue(tmem, sbase + OFF_WHI, sbase + OFF_WLO, kW_BYTES, idesc_hh);
          umma_commit(bar_d);
        }
      }
    }
  } else {
    // ===================== compute warps =====================
    const int q = warp & 3, hf = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const float* gate = reinterpret_cast<const float*>(sptr + OFF_GATE);
    const float* B1 = reinterpret_cast<const float*>(sptr + OFF_B1);
    const float* W2 = reinterpret_cast<const float*>(sptr + OFF_W2);
    const float* B2 = reinterpret_cast<const float*>(sptr + OFF_B2);
    float* yprev = reinterpret_cast<float*>(sptr + OFF_YPREV);
    float* io = reinterpret_cast<float*>(sptr + OFF_IO);
    const uint32_t hhi_row = sbase + OFF_HHI + hf * kH_BYTES + row * 128;
    const uint32_t hlo_row = sbase + OFF_HLO + hf * kH_BYTES + row * 128;

    auto publish_h = [&](const float (&h)[32]) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t off = (uint32_t)((c ^ (row & 7)) << 4);  // SWIZZLE_128B: chunk ^= row % 8
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
      for (int half16 = 0; half16 < 2; ++half16) {
        uint32_t ar[16], az[16], an[16];
        const uint32_t col = (uint32_t)(hf * 32 + half16 * 16);
        tmem_ld16_nowait(trow + col, ar);
        tmem_ld16_nowait(trow + 64 + col, az);
        tmem_ld16_nowait(trow + 128 + col, an);
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = hf * 32 + half16 * 16 + u;
          const float4 g0 = *reinterpret_cast<const float4*>(gate + j * 12);      // wr0 wr1 br wz0
          const float4 g1 = *reinterpret_cast<const float4*>(gate + j * 12 + 4);  // wz1 bz wn0 wn1
          const float2 g2 = *reinterpret_cast<const float2*>(gate + j * 12 + 8);  // bin bhn
          const float ir = fmaf(g0.y, yp.y, fmaf(g0.x, yp.x, g0.z));
          const float iz = fmaf(g1.x, yp.y, fmaf(g0.w, yp.x, g1.y));
          const float in_ = fmaf(g1.w, yp.y, fmaf(g1.z, yp.x, g2.x));
#ifdef OAT_FLOW_SHARED_RCP
          // r = 1/A, g = 1/B from ONE reciprocal: inv = 1/(A*B); r = inv*B, g = inv*A
          // (exponent clamped at 40 so A*B stays finite; sigmoid(-40) = 4e-18 either way)
          const float ea = 1.0f + __expf(fminf(-(ir + __uint_as_float(ar[u])), 40.0f));
          const float eb = 1.0f + __expf(fminf(-(iz + __uint_as_float(az[u])), 40.0f));
          const float inv = __fdividef(1.0f, ea * eb);
          const float rr = inv * eb;
          const float gg = inv * ea;
#else
          const float rr = sigmoid_fast(ir + __uint_as_float(ar[u]));
          const float gg = sigmoid_fast(iz + __uint_as_float(az[u]));
#endif
          const float nn = tanh_fast(fmaf(rr, __uint_as_float(an[u]) + g2.y, in_));
          const int hi = half16 * 16 + u;
          h[hi] = fmaf(gg, h[hi] - nn, nn);
        }
      }
      publish_h(h);  // -> head MMA of step t and recurrent MMA of step t+1

      // ---- head + affine flow update: the 4 warps of half 0 own one row each ------------
      if (hf == 0) {
        mbar_wait(bar_d2, (uint32_t)(t & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float o0 = B2[0], o1 = B2[1], o2 = B2[2], o3 = B2[3];
#pragma unroll
        for (int half16 = 0; half16 < 2; ++half16) {
          uint32_t d[16];
          tmem_ld16_nowait(trow + 192 + half16 * 16, d);
          tmem_wait_ld();
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = half16 * 16 + u;
            const float av = fmaxf(__uint_as_float(d[u]) + B1[j], 0.0f);
            o0 = fmaf(av, W2[j], o0);
            o1 = fmaf(av, W2[32 + j], o1);
            o2 = fmaf(av, W2[64 + j], o2);
            o3 = fmaf(av, W2[96 + j], o3);
          }
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
      asm volatile("bar.sync 1, 256;" ::: "memory");  // y_t visible to all 8 compute warps
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
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256)
                 : "memory");
  }
}

template <int MODE>
int launch_tc_mode(const FlowTcArgs& fa, int num_models, cudaStream_t stream) {
  const size_t smem = (size_t)OFF_IO + (size_t)TR * 2 * fa.T * sizeof(float) + 1024;
  if (smem > 227 * 1024) return fail("flow_tc: T too large for one CTA's shared memory");
  static size_t configured[64] = {0};
  int dev = 0;
  OAT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    OAT_CUDA(cudaFuncSetAttribute(flow_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  dim3 grid((unsigned)((fa.N + TR - 1) / TR), (unsigned)num_models);
  flow_tc_kernel<MODE><<<grid, TTHREADS, smem, stream>>>(fa);
  OAT_LAUNCHED(MODE == 0 ? "flow_sample" : "flow_score");
  return 0;
}

}  // namespace

// Host-side construction of the per-model image (called from oat_model_create):
// `fw` is the kFlow* layout of common.cuh; `img` receives kFlowTcFloats floats.
void pack_flow_tc_image(const float* fw, float* img) {
  auto put = [&](int byte_off, float v) { img[byte_off / 4] = v; };
  auto split = [](float w, float* hi, float* lo) {
    uint32_t x;
    memcpy(&x, &w, 4);
    uint32_t h = x & 0xffffe000u;
    memcpy(hi, &h, 4);
    float l = w - *hi;
    uint32_t lx;
    memcpy(&lx, &l, 4);
    lx &= 0xffffe000u;
    memcpy(lo, &lx, 4);
  };
  // B operand, K-major SWIZZLE_128B: element (n, k) of k-block kb = k / 32 lives at
  //   kb * rows*128 + n*128 + (((k%32)/4) ^ (n%8))*16 + (k%4)*4
  auto b_off = [](int rows, int n, int k) {
    const int kb = k >> 5, kk = k & 31;
    return kb * rows * 128 + n * 128 + ((((kk >> 2) ^ (n & 7))) << 4) + (kk & 3) * 4;
  };
  for (int n = 0; n < 192; ++n)
    for (int k = 0; k < 64; ++k) {
      float hi, lo;
      split(fw[kFlowWhh + k * 192 + n], &hi, &lo);  // whhT[k][n] = W_hh[n][k]
      put(OFF_WHI + b_off(192, n, k), hi);
      put(OFF_WLO + b_off(192, n, k), lo);
    }
  for (int n = 0; n < 32; ++n)
    for (int k = 0; k < 64; ++k) {
      float hi, lo;
      split(fw[kFlowW1T + k * 32 + n], &hi, &lo);
      put(OFF_W1HI + b_off(32, n, k), hi);
      put(OFF_W1LO + b_off(32, n, k), lo);
    }
  for (int j = 0; j < 64; ++j) {
    float* g = img + OFF_GATE / 4 + j * 12;
    const float* wih0 = fw + kFlowWihT;        // [192] weights of y_x
    const float* wih1 = fw + kFlowWihT + 192;  // [192] weights of y_y
    g[0] = wih0[j];        g[1] = wih1[j];        g[2] = fw[kFlowBih + j] + fw[kFlowBhh + j];
    g[3] = wih0[64 + j];   g[4] = wih1[64 + j];   g[5] = fw[kFlowBih + 64 + j] + fw[kFlowBhh + 64 + j];
    g[6] = wih0[128 + j];  g[7] = wih1[128 + j];  g[8] = fw[kFlowBih + 128 + j];
    g[9] = fw[kFlowBhh + 128 + j];
    g[10] = 0.0f; g[11] = 0.0f;
  }
  for (int j = 0; j < 32; ++j) img[OFF_B1 / 4 + j] = fw[kFlowB1 + j];
  for (int i = 0; i < 128; ++i) img[OFF_W2 / 4 + i] = fw[kFlowW2 + i];
  for (int i = 0; i < 4; ++i) img[OFF_B2 / 4 + i] = fw[kFlowB2 + i];
}

int launch_flow_tc(const FlowLaunch& a, cudaStream_t stream) {
  if (a.N <= 0 || a.T <= 0) return 0;
  if (a.mode != 0 && a.mode != 1) return fail("flow_tc: only sample/score modes");
  FlowTcArgs fa;
  fa.images = a.weights;
  fa.in = a.in; fa.z = a.z; fa.z_model_stride = a.z_model_stride;
  fa.out = a.out; fa.logprob = a.logprob; fa.logabsdet = a.logabsdet; fa.q = a.q;
  fa.out_model_stride = a.out_model_stride;
  fa.goal = a.goal; fa.G = a.G;
  const double eps = a.epsilon;
  fa.inv_two_eps2 = (float)(1.0 / (2.0 * eps * eps));
  fa.log_norm = a.goal ? (float)(-log(2.0 * 3.14159265358979323846 * eps * eps) - log((double)a.G)) : 0.0f;
  fa.N = a.N; fa.T = a.T; fa.rows_per_z = a.rows_per_z; fa.skip_model = a.skip_model;
  return a.mode == 0 ? launch_tc_mode<0>(fa, a.num_models, stream)
                     : launch_tc_mode<1>(fa, a.num_models, stream);
}

}  // namespace oat
