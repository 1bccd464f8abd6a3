// Fused autoregressive-flow kernels (sm_100a): the metric's unit of work.
//
// Replaces oatomobile/torch/networks/sequence.py:95-216 (`_forward`, `_inverse`)
// and the GRU roll-out of oatomobile/baselines/torch/cil/model.py:106-127.
//
// One CTA owns a tile of FR=64 consecutive rows (samples) and walks all T steps
// with the GRU state, the head activations and the x/y tile resident in shared
// memory; the model's 15 268 fp32 weights (61 KB) are staged once per CTA.  Per
// step the [64,64]x[64,192] recurrent product runs as a register-tiled FP32 GEMM
// (each thread: 4 rows x 4 hidden units x 3 gates = 48 accumulators), followed by
// the gate non-linearities, the 64->32->4 head and the affine flow update.
// HBM traffic is only the x/y tile (coalesced float4, staged through smem) and
// one score per row; no [N,64..192] intermediate ever reaches global memory.
#include "common.cuh"

namespace oat {
namespace {

constexpr int FR = 64;         // rows per CTA
constexpr int FTHREADS = 256;  // 16 row-groups x 16 unit-groups
constexpr int HS = 68;         // padded row stride (floats) of the h tiles
constexpr int AS = 36;         // padded row stride of the head activations
constexpr float kLog2Pi = 1.8378770664093453f;

struct FlowArgs {
  PtrTable weights;
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
  float inv_two_eps2;  // 1 / (2 eps^2)
  float log_norm;      // -log(2 pi eps^2) - log G
  int64_t N;
  int T;
  int rows_per_z;
  int skip_model;
};

__device__ __forceinline__ float sigmoid_fast(float v) {
  // 1/(1+e^-v): ex2.approx + rcp.approx, |abs err| < 3e-7 over the real line.
  return __fdividef(1.0f, 1.0f + __expf(-v));
}
__device__ __forceinline__ float tanh_fast(float v) {
  return __fdividef(2.0f, 1.0f + __expf(-2.0f * v)) - 1.0f;
}
__device__ __forceinline__ float softplus_ref(float v) {
  // torch.nn.functional.softplus(beta=1, threshold=20)
  return v > 20.0f ? v : log1pf(expf(v));
}

// MODE 0: sample  y_t = mu + sigma * x_t   (sequence.py:136)   — also scores.
// MODE 1: score   x_t = (y_t - mu) / sigma (sequence.py:196).
// MODE 2: CIL     y_t = y_{t-1} + W_out h  (cil/model.py:118-124).
template <int MODE>
__global__ void __launch_bounds__(FTHREADS, 2) flow_kernel(const __grid_constant__ FlowArgs a) {
  extern __shared__ float4 smem4[];
  float* sm = reinterpret_cast<float*>(smem4);
  float* Whh = sm + kFlowWhh;
  float* W1T = sm + kFlowW1T;
  float* WihT = sm + kFlowWihT;
  float* Bih = sm + kFlowBih;
  float* Bhh = sm + kFlowBhh;
  float* B1 = sm + kFlowB1;
  float* W2 = sm + kFlowW2;
  float* B2 = sm + kFlowB2;
  float* hb0 = sm + kFlowFloats;
  float* hb1 = hb0 + FR * HS;
  float* yprev = hb1 + FR * HS;  // [FR][2]
  float* io = yprev + FR * 2;    // [FR][T*2]

  const int tid = threadIdx.x;
  const int model = blockIdx.y;
  if (model == a.skip_model) return;  // already scored by the sampling pass
  const int64_t row0 = (int64_t)blockIdx.x * FR;
  const int T = a.T;
  const int T2 = 2 * T;
  const int rows_here = (int)min((int64_t)FR, a.N - row0);

  // ---- stage weights (float4, coalesced, L2-resident after the first CTA) ----
  {
    const float4* src = reinterpret_cast<const float4*>(a.weights.p[model]);
    float4* dst = reinterpret_cast<float4*>(sm);
    for (int i = tid; i < kFlowFloats / 4; i += FTHREADS) dst[i] = __ldg(src + i);
  }
  // ---- h_0 = z[row / rows_per_z]  (sequence.py:119-128: the GRU state starts at z)
  {
    const float* zb = a.z + (int64_t)model * a.z_model_stride;
    for (int i = tid; i < FR * kHidden; i += FTHREADS) {
      const int r = i >> 6, k = i & 63;
      float v = 0.0f;
      if (r < rows_here) v = __ldg(zb + ((row0 + r) / a.rows_per_z) * kHidden + k);
      hb0[r * HS + k] = v;
    }
  }
  if (tid < FR * 2) yprev[tid] = 0.0f;  // y_{-1} = 0 (sequence.py:119-122)
  // ---- stage the x (sample) / y (score) tile ---------------------------------
  if (MODE != 2) {
    const float* src = a.in + row0 * T2;
    const int n = rows_here * T2;
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(io);
      for (int i = tid; i < n / 4; i += FTHREADS) d4[i] = __ldg(s4 + i);
    } else {
      for (int i = tid; i < n; i += FTHREADS) io[i] = __ldg(src + i);
    }
    for (int i = n + tid; i < FR * T2; i += FTHREADS) io[i] = 0.0f;
  }
  __syncthreads();

  const int ug = tid & 15;  // unit group: hidden units 4ug..4ug+3
  const int rg = tid >> 4;  // row group : rows 4rg..4rg+3
  const int hrow = tid >> 2;  // head stage: row
  const int oi = tid & 3;     // head stage: output index (0,1 = dloc; 2,3 = scale)
  float sumsq = 0.0f, sumlog = 0.0f, goal_ll = 0.0f;
  int cur = 0;

  for (int t = 0; t < T; ++t) {
    float* hc = cur ? hb1 : hb0;
    float* hn = cur ? hb0 : hb1;

    // ---- recurrent GEMM: acc[g][r][u] = b_hh + sum_k h[r][k] * W_hh^T[k][g*64+4ug+u]
    float acc[3][4][4];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4 b = *reinterpret_cast<const float4*>(Bhh + g * 64 + 4 * ug);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[g][r][0] = b.x; acc[g][r][1] = b.y; acc[g][r][2] = b.z; acc[g][r][3] = b.w;
      }
    }
#pragma unroll 2
    for (int k4 = 0; k4 < 16; ++k4) {
      float hv[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(hc + (4 * rg + r) * HS + 4 * k4);
        hv[r][0] = v.x; hv[r][1] = v.y; hv[r][2] = v.z; hv[r][3] = v.w;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float* wrow = Whh + (4 * k4 + kk) * 192 + 4 * ug;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const float4 w = *reinterpret_cast<const float4*>(wrow + g * 64);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            acc[g][r][0] = fmaf(hv[r][kk], w.x, acc[g][r][0]);
            acc[g][r][1] = fmaf(hv[r][kk], w.y, acc[g][r][1]);
            acc[g][r][2] = fmaf(hv[r][kk], w.z, acc[g][r][2]);
            acc[g][r][3] = fmaf(hv[r][kk], w.w, acc[g][r][3]);
          }
        }
      }
    }
    __syncthreads();  // y_{t-1} of the previous step's flow stage is visible

    // ---- gates (torch.nn.GRUCell, order r|z|n) and h' = n + g (h - n) ----------
    {
      float wi[3][2][4], bi[3][4];
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const float4 b = *reinterpret_cast<const float4*>(Bih + g * 64 + 4 * ug);
        const float4 w0 = *reinterpret_cast<const float4*>(WihT + g * 64 + 4 * ug);
        const float4 w1 = *reinterpret_cast<const float4*>(WihT + 192 + g * 64 + 4 * ug);
        bi[g][0] = b.x; bi[g][1] = b.y; bi[g][2] = b.z; bi[g][3] = b.w;
        wi[g][0][0] = w0.x; wi[g][0][1] = w0.y; wi[g][0][2] = w0.z; wi[g][0][3] = w0.w;
        wi[g][1][0] = w1.x; wi[g][1][1] = w1.y; wi[g][1][2] = w1.z; wi[g][1][3] = w1.w;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = 4 * rg + r;
        const float2 yp = *reinterpret_cast<const float2*>(yprev + 2 * row);
        const float4 ho = *reinterpret_cast<const float4*>(hc + row * HS + 4 * ug);
        const float hold[4] = {ho.x, ho.y, ho.z, ho.w};
        float hnew[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float ir = fmaf(wi[0][1][u], yp.y, fmaf(wi[0][0][u], yp.x, bi[0][u]));
          const float iz = fmaf(wi[1][1][u], yp.y, fmaf(wi[1][0][u], yp.x, bi[1][u]));
          const float in = fmaf(wi[2][1][u], yp.y, fmaf(wi[2][0][u], yp.x, bi[2][u]));
          const float rr = sigmoid_fast(ir + acc[0][r][u]);
          const float gg = sigmoid_fast(iz + acc[1][r][u]);
          const float nn = tanh_fast(fmaf(rr, acc[2][r][u], in));
          hnew[u] = fmaf(gg, hold[u] - nn, nn);
        }
        *reinterpret_cast<float4*>(hn + row * HS + 4 * ug) =
            make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
      }
    }
    __syncthreads();  // h_t complete

    float o;  // head output `oi` of row `hrow`
    if (MODE != 2) {
      // ---- head layer 0: a = relu(h W1^T + b1), thread = 4 rows x 2 columns -----
      float ha[4][2];
      {
        const float2 b = *reinterpret_cast<const float2*>(B1 + 2 * ug);
#pragma unroll
        for (int r = 0; r < 4; ++r) { ha[r][0] = b.x; ha[r][1] = b.y; }
      }
#pragma unroll 4
      for (int k4 = 0; k4 < 16; ++k4) {
        float hv[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 v = *reinterpret_cast<const float4*>(hn + (4 * rg + r) * HS + 4 * k4);
          hv[r][0] = v.x; hv[r][1] = v.y; hv[r][2] = v.z; hv[r][3] = v.w;
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float2 w = *reinterpret_cast<const float2*>(W1T + (4 * k4 + kk) * 32 + 2 * ug);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            ha[r][0] = fmaf(hv[r][kk], w.x, ha[r][0]);
            ha[r][1] = fmaf(hv[r][kk], w.y, ha[r][1]);
          }
        }
      }
      float* as = hc;  // h_{t-1} is dead after the barrier above: reuse as [FR][AS]
#pragma unroll
      for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float2*>(as + (4 * rg + r) * AS + 2 * ug) =
            make_float2(fmaxf(ha[r][0], 0.0f), fmaxf(ha[r][1], 0.0f));
      __syncthreads();
      // ---- head layer 1: o = a W2^T + b2  (sequence.py:131) ----------------------
      o = B2[oi];
      const float* arow = as + hrow * AS;
      const float* wrow = W2 + oi * 32;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 av = *reinterpret_cast<const float4*>(arow + 4 * j4);
        const float4 wv = *reinterpret_cast<const float4*>(wrow + 4 * j4);
        o = fmaf(av.x, wv.x, o); o = fmaf(av.y, wv.y, o);
        o = fmaf(av.z, wv.z, o); o = fmaf(av.w, wv.w, o);
      }
    } else {
      // ---- CIL: dx = W_out h + b_out (cil/model.py:121), W_out in the w1T slot ---
      o = 0.0f;
      if (oi < 2) {
        o = B2[oi];
        const float* hrow_p = hn + hrow * HS;
        const float* wrow = W1T + oi * 64;
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
          const float4 hv = *reinterpret_cast<const float4*>(hrow_p + 4 * k4);
          const float4 wv = *reinterpret_cast<const float4*>(wrow + 4 * k4);
          o = fmaf(hv.x, wv.x, o); o = fmaf(hv.y, wv.y, o);
          o = fmaf(hv.z, wv.z, o); o = fmaf(hv.w, wv.w, o);
        }
      }
    }
    const float oscale = __shfl_down_sync(0xffffffffu, o, 2);

    // ---- affine flow update, lanes oi = 0 (x coordinate) and 1 (y coordinate) ---
    float ycoord = 0.0f;
    if (oi < 2) {
      const float ypv = yprev[2 * hrow + oi];
      if (MODE == 2) {
        ycoord = ypv + o;  // x = dx + x
        io[hrow * T2 + 2 * t + oi] = ycoord;
      } else {
        const float mu = ypv + o;                           // y_{t-1} + dloc
        const float sigma = softplus_ref(oscale) + 1e-3f;   // sequence.py:133
        const float v = io[hrow * T2 + 2 * t + oi];
        float xb;
        if (MODE == 0) {
          ycoord = fmaf(sigma, v, mu);                      // sequence.py:136
          io[hrow * T2 + 2 * t + oi] = ycoord;
          xb = (ycoord - mu) / sigma;  // what `_inverse` recovers from the rounded y
        } else {
          ycoord = v;
          xb = (v - mu) / sigma;                            // sequence.py:196
          if (a.out != nullptr) io[hrow * T2 + 2 * t + oi] = xb;
        }
        sumsq = fmaf(xb, xb, sumsq);
        sumlog += logf(sigma);
      }
      yprev[2 * hrow + oi] = ycoord;
    }
    if (MODE != 2 && a.goal != nullptr && t == T - 1) {
      // per-sample goal log-likelihood (dim/model.py:163-171 without the batch mean)
      const float y0 = __shfl_sync(0xffffffffu, ycoord, (tid & 31) & ~3);
      const float y1 = __shfl_sync(0xffffffffu, ycoord, ((tid & 31) & ~3) | 1);
      if (oi == 0 && hrow < rows_here) {
        const float* g = a.goal + ((row0 + hrow) / a.rows_per_z) * a.G * 2;
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
    }
    cur ^= 1;
  }

  // ---- per-row results ----------------------------------------------------------
  if (MODE != 2) {
    const float sq = sumsq + __shfl_xor_sync(0xffffffffu, sumsq, 1);
    const float sl = sumlog + __shfl_xor_sync(0xffffffffu, sumlog, 1);
    if (oi == 0 && hrow < rows_here) {
      const float lp = -0.5f * sq - (float)T * kLog2Pi;  // MVN(0,I).log_prob, sequence.py:208
      const int64_t idx = (int64_t)model * a.out_model_stride + row0 + hrow;
      if (a.logprob != nullptr) a.logprob[idx] = lp;
      if (a.logabsdet != nullptr) a.logabsdet[idx] = sl;
      if (a.q != nullptr) a.q[idx] = (lp - sl) + goal_ll;
    }
  }
  if (a.out != nullptr) {
    __syncthreads();
    float* dst = a.out + row0 * T2;
    const int n = rows_here * T2;
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(io);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int i = tid; i < n / 4; i += FTHREADS) d4[i] = s4[i];
    } else {
      for (int i = tid; i < n; i += FTHREADS) dst[i] = io[i];
    }
  }
}

size_t flow_smem_bytes(int T) {
  return sizeof(float) * (size_t)(kFlowFloats + 2 * FR * HS + FR * 2 + FR * 2 * T);
}

template <int MODE>
int launch_mode(const FlowArgs& fa, int num_models, cudaStream_t stream) {
  const size_t smem = flow_smem_bytes(fa.T);
  if (smem > 227 * 1024) return fail("flow: T too large for one CTA's shared memory");
  static size_t configured[64] = {0};  // per device
  int dev = 0;
  OAT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    OAT_CUDA(cudaFuncSetAttribute(flow_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  dim3 grid((unsigned)((fa.N + FR - 1) / FR), (unsigned)num_models);
  flow_kernel<MODE><<<grid, FTHREADS, smem, stream>>>(fa);
  OAT_LAUNCHED(MODE == 0 ? "flow_sample" : MODE == 1 ? "flow_score" : "cil_rollout");
  return 0;
}

}  // namespace

int launch_flow(const FlowLaunch& a, cudaStream_t stream) {
  if (a.N <= 0 || a.T <= 0) return 0;
  if (a.num_models < 1 || a.num_models > kMaxModels) return fail("flow: bad model count");
  if (a.rows_per_z < 1) return fail("flow: rows_per_z must be >= 1");
  FlowArgs fa;
  fa.weights = a.weights;
  fa.in = a.in;
  fa.z = a.z;
  fa.z_model_stride = a.z_model_stride;
  fa.out = a.out;
  fa.logprob = a.logprob;
  fa.logabsdet = a.logabsdet;
  fa.q = a.q;
  fa.out_model_stride = a.out_model_stride;
  fa.goal = a.goal;
  fa.G = a.G;
  const double eps = a.epsilon;
  fa.inv_two_eps2 = (float)(1.0 / (2.0 * eps * eps));
  fa.log_norm = a.goal ? (float)(-log(2.0 * 3.14159265358979323846 * eps * eps) - log((double)a.G))
                       : 0.0f;
  fa.N = a.N;
  fa.T = a.T;
  fa.rows_per_z = a.rows_per_z;
  fa.skip_model = a.skip_model;
  switch (a.mode) {
    case 0: return launch_mode<0>(fa, a.num_models, stream);
    case 1: return launch_mode<1>(fa, a.num_models, stream);
    case 2: return launch_mode<2>(fa, a.num_models, stream);
  }
  return fail("flow: bad mode");
}

}  // namespace oat
