// Stand-alone `MLP.forward` (oatomobile/torch/networks/mlp.py:25-72): a stack of
// Linear(+ReLU) layers evaluated in ONE launch, activations kept in shared memory.
//
// The weights are read where the module keeps them — PyTorch layout W[out][in], fp32,
// device memory — so nothing is packed or cached.  One CTA per input row; a warp owns an
// output neuron at a time: its 32 lanes stride over the input vector (coalesced reads of
// the weight row), the partial sums meet in a shuffle reduction.  Inside the fused path the
// merger MLP runs in `merger_kernel` (encoder.cu); this entry point is for callers that use
// the `MLP` class on its own (the reference exports it, networks/__init__.py:17-19).
#include "common.cuh"

namespace oat {
namespace {

constexpr int kMlpMaxLayers = 8;
constexpr int kMlpMaxWidth = 4096;
constexpr int kMlpThreads = 256;

struct MlpArgs {
  const float* w[kMlpMaxLayers];
  const float* b[kMlpMaxLayers];
  int size[kMlpMaxLayers + 1];  // size[0] = input width, size[l + 1] = width of layer l
  int layers;
  int activate_final;
  const float* x;  // [B][size[0]]
  float* out;      // [B][size[layers]]
  int width;       // largest layer width (shared-memory row length)
};

__global__ void __launch_bounds__(kMlpThreads) mlp_kernel(const __grid_constant__ MlpArgs a) {
  extern __shared__ float act[];  // two rows of a.width floats
  float* cur = act;
  float* nxt = act + a.width;
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < a.size[0]; i += kMlpThreads) cur[i] = a.x[(int64_t)row * a.size[0] + i];
  __syncthreads();
  for (int l = 0; l < a.layers; ++l) {
    const int K = a.size[l], N = a.size[l + 1];
    const bool relu = l + 1 < a.layers || a.activate_final;
    const bool last = l + 1 == a.layers;
    for (int n = warp; n < N; n += kMlpThreads / 32) {
      const float* __restrict__ wr = a.w[l] + (int64_t)n * K;
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s = fmaf(__ldg(wr + k), cur[k], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) {
        s += a.b[l] ? __ldg(a.b[l] + n) : 0.f;
        if (relu) s = fmaxf(s, 0.f);
        if (last) a.out[(int64_t)row * N + n] = s;
        else nxt[n] = s;
      }
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

}  // namespace

int launch_mlp(const float* const* w, const float* const* b, const int* sizes, int layers,
               int activate_final, const float* x, int B, float* out, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (layers < 1 || layers > kMlpMaxLayers) return fail("oat_mlp_forward: 1 <= layers <= 8");
  MlpArgs a;
  a.layers = layers; a.activate_final = activate_final; a.x = x; a.out = out; a.width = 0;
  for (int l = 0; l <= layers; ++l) {
    if (sizes[l] < 1 || sizes[l] > kMlpMaxWidth) return fail("oat_mlp_forward: layer width must be in [1, 4096]");
    a.size[l] = sizes[l];
    if (sizes[l] > a.width) a.width = sizes[l];
  }
  for (int l = 0; l < layers; ++l) {
    if (!w[l]) return fail("oat_mlp_forward: null weight");
    a.w[l] = w[l];
    a.b[l] = b ? b[l] : nullptr;
  }
  mlp_kernel<<<B, kMlpThreads, 2 * a.width * sizeof(float), stream>>>(a);
  OAT_LAUNCHED("mlp");
  return 0;
}

}  // namespace oat
