// Ensemble aggregation + plan selection (sm_100a).
//
// Replaces oatomobile/baselines/torch/rip/agent.py:121-127 (stack -> min/max/mean
// over models) and the plan decode of :137 for the K-sample formulation:
//   s[b,k] = min_m(-q) | max_m(-q) | mean_m(-q);  k* = argmin_k s;  plan = y[b,k*].
// HBM-bound: reads E*4 bytes per sample once (coalesced along k), writes 4 (s).
// One CTA per scene b; warp-shuffle (value,index) reduction, lowest index on ties
// like torch.argmin; MA sums models in index order 0..E-1 then divides by E so
// every rank of a sharded ensemble reproduces the same bits.
#include "common.cuh"

namespace oat {
namespace {

constexpr int ATHREADS = 256;

__device__ __forceinline__ void argmin_combine(float& v, int& k, float v2, int k2) {
  if (v2 < v || (v2 == v && k2 < k)) { v = v2; k = k2; }
}

__global__ void __launch_bounds__(ATHREADS) aggregate_kernel(
    const float* __restrict__ q, int E, int B, int K, int algo, const float* __restrict__ y,
    int T, float* __restrict__ s, int32_t* __restrict__ kstar, float* __restrict__ sbest,
    float* __restrict__ plan) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t BK = (int64_t)B * K;
  const float* qb = q + (int64_t)b * K;
  float best = INFINITY;
  int bestk = 0x7fffffff;
  for (int k = tid; k < K; k += ATHREADS) {
    float v = -__ldg(qb + k);
    if (algo == OAT_ALGO_WCM) {
      for (int m = 1; m < E; ++m) v = fminf(v, -__ldg(qb + m * BK + k));
    } else if (algo == OAT_ALGO_BCM) {
      for (int m = 1; m < E; ++m) v = fmaxf(v, -__ldg(qb + m * BK + k));
    } else {
      for (int m = 1; m < E; ++m) v = v + (-__ldg(qb + m * BK + k));
      v = v / (float)E;
    }
    if (s != nullptr) s[(int64_t)b * K + k] = v;
    if (v < best) { best = v; bestk = k; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, best, off);
    const int k2 = __shfl_xor_sync(0xffffffffu, bestk, off);
    argmin_combine(best, bestk, v2, k2);
  }
  __shared__ float sv[ATHREADS / 32];
  __shared__ int sk[ATHREADS / 32];
  __shared__ int kfinal;
  if ((tid & 31) == 0) { sv[tid >> 5] = best; sk[tid >> 5] = bestk; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < ATHREADS / 32; ++w) argmin_combine(best, bestk, sv[w], sk[w]);
    if (bestk == 0x7fffffff) bestk = 0;  // all-NaN / empty row
    kfinal = bestk;
    kstar[b] = bestk;
    if (sbest != nullptr) sbest[b] = best;
  }
  __syncthreads();
  if (plan != nullptr && y != nullptr) {
    const float* src = y + ((int64_t)b * K + kfinal) * (2 * T);
    for (int i = tid; i < 2 * T; i += ATHREADS) plan[(int64_t)b * 2 * T + i] = __ldg(src + i);
  }
}

// ImitativeModel._goal_likelihood (dim/model.py:143-171): per-row mixture log-likelihood
// of y[:, -1] under N(goal_g, eps^2 I) with uniform weights, and its batch mean.
__global__ void __launch_bounds__(256) goal_likelihood_kernel(
    const float* __restrict__ y_last, const float* __restrict__ goal, int B, int G,
    float inv_two_eps2, float log_norm, float* __restrict__ rows, float* __restrict__ mean) {
  __shared__ float part[8];
  float local = 0.0f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float y0 = y_last[2 * b], y1 = y_last[2 * b + 1];
    const float* g = goal + (int64_t)b * G * 2;
    float mx = -INFINITY;
    for (int i = 0; i < G; ++i) {
      const float d0 = y0 - g[2 * i], d1 = y1 - g[2 * i + 1];
      mx = fmaxf(mx, -(d0 * d0 + d1 * d1) * inv_two_eps2);
    }
    float se = 0.0f;
    for (int i = 0; i < G; ++i) {
      const float d0 = y0 - g[2 * i], d1 = y1 - g[2 * i + 1];
      se += expf(-(d0 * d0 + d1 * d1) * inv_two_eps2 - mx);
    }
    const float ll = mx + logf(se) + log_norm;
    if (rows) rows[b] = ll;
    local += ll;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0 && mean) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += part[w];
    *mean = t / (float)B;
  }
}

}  // namespace

int launch_goal_likelihood(const float* y_last, const float* goal, int B, int G, float epsilon,
                           float* rows, float* mean, cudaStream_t stream) {
  const double eps = epsilon;
  goal_likelihood_kernel<<<1, 256, 0, stream>>>(
      y_last, goal, B, G, (float)(1.0 / (2.0 * eps * eps)),
      (float)(-log(2.0 * 3.14159265358979323846 * eps * eps) - log((double)G)), rows, mean);
  OAT_LAUNCHED("goal_likelihood");
  return 0;
}

int launch_aggregate(const float* q, int E, int B, int K, int algo, const float* y, int T,
                     float* s, int32_t* kstar, float* sbest, float* plan,
                     cudaStream_t stream) {
  if (B <= 0) return 0;
  if (E < 1 || K < 1) return fail("aggregate: need E >= 1 and K >= 1");
  if (algo < OAT_ALGO_WCM || algo > OAT_ALGO_MA) return fail("aggregate: unknown algorithm");
  aggregate_kernel<<<B, ATHREADS, 0, stream>>>(q, E, B, K, algo, y, T, s, kstar, sbest, plan);
  OAT_LAUNCHED("aggregate");
  return 0;
}

}  // namespace oat
