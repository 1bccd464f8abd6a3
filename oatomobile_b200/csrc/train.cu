// Training step on the GPU (dim/train.py:175-213, cil/train.py:168-190).
//
// The arithmetic lives in train_functors.h (work-item bodies) and train_impl.h (the
// sequence of launches).  This file supplies the CUDA backend — every functor becomes a
// grid-stride kernel on the caller's stream — and the C-ABI entry points.
//
// FP32 SIMT: the pointwise forward and input-gradient products run on the shared-memory tiled
// GEMM of the inference path (encoder.cu), the weight gradients and everything else as
// register-tiled work items straight from global memory, reductions through atomics.  Accuracy matters more than speed here (the BatchNorm
// backward over a handful of rows is badly conditioned — see DESIGN.md §9); the
// tensor-core / fused-block treatment the inference path received comes next.
#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "train_impl.h"

namespace oat {
namespace {

template <class F>
__global__ void __launch_bounds__(128) run_functor(F f, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

// One batch row per thread with a long serial body: spread the rows over as many SMs as
// possible instead of packing them into one block.
template <class F>
struct BlockSize {
  static constexpr int value = 128;
};
template <>
struct BlockSize<train::DimNllStep> {
  static constexpr int value = 32;
};
template <>
struct BlockSize<train::CilL1Step> {
  static constexpr int value = 32;
};

// ---------------------------------------------------------------------------------------------
// Cooperative form of the DIM decoder pass (train_functors.h: DimNllStep): ONE 64-thread CTA per
// batch row instead of one thread — thread j owns hidden unit j in the forward recurrence and
// column k = j of the transposed products in the reverse sweep.  Every sum is accumulated in the
// SAME order as in the work-item functor (which stays the host-emulated reference), so the two
// produce identical bits; at B = 64 the pass drops from 0.96 ms (64 threads on the whole chip)
// to tens of microseconds.
// ---------------------------------------------------------------------------------------------
// gru_forward for hidden unit j (h, us in shared memory); returns h'_j, records the step
__device__ __forceinline__ float coop_gru_forward(const train::DecParams& p, const float* us, const float* h,
                                                   float* rec, int j) {
  using namespace train;
  float ar = p.bhh[j], ag = p.bhh[64 + j], an = p.bhh[128 + j], ar2 = 0.0f, ag2 = 0.0f, an2 = 0.0f;
  const float *wr = p.whh + j * 64, *wg = p.whh + (64 + j) * 64, *wn = p.whh + (128 + j) * 64;
#pragma unroll
  for (int k = 0; k < 64; k += 4) {
    const F4 a = ld4(wr + k), bb = ld4(wg + k), c = ld4(wn + k);
    ar = fmaf(a.x, h[k], ar); ar2 = fmaf(a.y, h[k + 1], ar2);
    ar = fmaf(a.z, h[k + 2], ar); ar2 = fmaf(a.w, h[k + 3], ar2);
    ag = fmaf(bb.x, h[k], ag); ag2 = fmaf(bb.y, h[k + 1], ag2);
    ag = fmaf(bb.z, h[k + 2], ag); ag2 = fmaf(bb.w, h[k + 3], ag2);
    an = fmaf(c.x, h[k], an); an2 = fmaf(c.y, h[k + 1], an2);
    an = fmaf(c.z, h[k + 2], an); an2 = fmaf(c.w, h[k + 3], an2);
  }
  ar += ar2; ag += ag2; an += an2;
  const float u0 = us[0], u1 = us[1];
  const float ir = fmaf(p.wih[j * 2 + 1], u1, fmaf(p.wih[j * 2], u0, p.bih[j]));
  const float ig = fmaf(p.wih[(64 + j) * 2 + 1], u1, fmaf(p.wih[(64 + j) * 2], u0, p.bih[64 + j]));
  const float in = fmaf(p.wih[(128 + j) * 2 + 1], u1, fmaf(p.wih[(128 + j) * 2], u0, p.bih[128 + j]));
  const float r = sigmoidf_(ir + ar), g = sigmoidf_(ig + ag);
  const float n = tanhf(fmaf(r, an, in));
  const float hc = fmaf(g, h[j] - n, n);
  rec[kRecH + j] = h[j]; rec[kRecR + j] = r; rec[kRecG + j] = g; rec[kRecN + j] = n;
  rec[kRecHn + j] = an; rec[kRecHcur + j] = hc;
  return hc;
}

// gru_backward: thread j turns d = dh[j] into the gate gradients of unit j (recorded, and left in
// ghs / gis for the transposed products), then returns dh_prev[j] accumulated in the functor's
// order: for unit u ascending: (u == j: + d_j g_j), then the three gates' W_hh[row][j] terms.
// Contains two barriers; ghs / gis hold the step's d(W_hh h) / d(W_ih u) gradients afterwards.
__device__ __forceinline__ float coop_gru_backward(const train::DecParams& p, float* rec, float d, float* ghs,
                                                    float* gis, int j) {
  using namespace train;
  const float hp = rec[kRecH + j], r = rec[kRecR + j], g = rec[kRecG + j], n = rec[kRecN + j],
              an = rec[kRecHn + j];
  const float dn_pre = d * (1.0f - g) * (1.0f - n * n);
  const float dg_pre = d * (hp - n) * g * (1.0f - g);
  const float dr_pre = dn_pre * an * r * (1.0f - r);
  const float dhn = dn_pre * r;
  rec[kRecDgi + j] = dr_pre; rec[kRecDgi + 64 + j] = dg_pre; rec[kRecDgi + 128 + j] = dn_pre;
  rec[kRecDgh + j] = dr_pre; rec[kRecDgh + 64 + j] = dg_pre; rec[kRecDgh + 128 + j] = dhn;
  __syncthreads();  // previous readers of ghs / gis are done
  ghs[j] = dr_pre; ghs[64 + j] = dg_pre; ghs[128 + j] = dhn;
  gis[j] = dr_pre; gis[64 + j] = dg_pre; gis[128 + j] = dn_pre;
  __syncthreads();
  float acc = 0.0f;
  for (int u = 0; u < 64; ++u) {
    if (u == j) acc = fmaf(d, g, acc);
#pragma unroll
    for (int q = 0; q < 3; ++q) acc = fmaf(ghs[q * 64 + u], p.whh[(q * 64 + u) * 64 + j], acc);
  }
  return acc;
}

__global__ void __launch_bounds__(64) dim_nll_coop_kernel(train::DimNllStep f) {
  using namespace train;
  const int b = blockIdx.x, j = threadIdx.x;
  const DecParams& p = f.p;
  const int T = f.T;
  __shared__ float h[64], hn[64], a1s[32], os[4], ghs[192], gis[192], das[32], us[2];
  h[j] = f.z[(int64_t)b * 64 + j];
  float* rec0 = f.scratch + (int64_t)b * T * kDecRecord;
  const float* yb = f.y + (int64_t)b * T * 2;
  const float invB = 1.0f / (float)f.B;
  float row_loss = (float)T * 1.8378770664093453f;  // thread 0 only
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    float* rec = rec0 + t * kDecRecord;
    float* misc = rec + kRecMisc;
    if (j < 2) {
      const float u = t > 0 ? yb[(t - 1) * 2 + j] : 0.0f;
      us[j] = u;
      misc[6 + j] = u;
    }
    __syncthreads();
    hn[j] = coop_gru_forward(p, us, h, rec, j);
    __syncthreads();
    if (j < 32) {  // head hidden layer
      float acc = p.b1[j], acc2 = 0.0f;
#pragma unroll
      for (int k = 0; k < 64; k += 4) {
        const F4 wv = ld4(p.w1 + j * 64 + k);
        acc = fmaf(wv.x, hn[k], acc); acc2 = fmaf(wv.y, hn[k + 1], acc2);
        acc = fmaf(wv.z, hn[k + 2], acc); acc2 = fmaf(wv.w, hn[k + 3], acc2);
      }
      const float a = fmaxf(acc + acc2, 0.0f);
      a1s[j] = a;
      rec[kRecA1 + j] = a;
    }
    __syncthreads();
    if (j < 4) {
      float acc = p.b2[j];
      for (int q = 0; q < 32; ++q) acc = fmaf(p.w2[j * 32 + q], a1s[q], acc);
      os[j] = acc;
    }
    __syncthreads();
    if (j == 0) {
      for (int d = 0; d < 2; ++d) {
        const float mu = us[d] + os[d];
        const float sraw = os[2 + d];
        const float sp = sraw > 20.0f ? sraw : log1pf(expf(sraw));
        const float sigma = sp + 1e-3f;
        const float x = (yb[t * 2 + d] - mu) / sigma;
        misc[d] = x; misc[2 + d] = sigma; misc[4 + d] = sraw;
        row_loss += 0.5f * x * x + logf(sigma);
      }
    }
    h[j] = hn[j];
    __syncthreads();
  }
  if (j == 0) atomicAdd(f.loss, (double)row_loss);

  float dh = 0.0f;  // dh[j]: only thread j ever touches it
  for (int t = T - 1; t >= 0; --t) {
    float* rec = rec0 + t * kDecRecord;
    const float* misc = rec + kRecMisc;
    float dout[4];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const float x = misc[d], sigma = misc[2 + d], sraw = misc[4 + d];
      const float dx = x * invB;
      dout[d] = -dx / sigma;
      const float dsigma = (invB - dx * x) / sigma;
      dout[2 + d] = dsigma * (sraw > 20.0f ? 1.0f : sigmoidf_(sraw));
    }
    if (j < 4) rec[kRecDout + j] = dout[j];
    __syncthreads();  // das of the previous step has been consumed
    if (j < 32) {
      float da = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) da = fmaf(dout[i], p.w2[i * 32 + j], da);
      if (!(rec[kRecA1 + j] > 0.0f)) da = 0.0f;
      rec[kRecDa1 + j] = da;
      das[j] = da;
    }
    __syncthreads();
    for (int q = 0; q < 32; ++q) {  // dh[k] += sum_j da[j] * W1[j][k], j ascending (zero terms add nothing)
      const float da = das[q];
      if (da != 0.0f) dh = fmaf(da, p.w1[q * 64 + j], dh);
    }
    dh = coop_gru_backward(p, rec, dh, ghs, gis, j);  // inputs are data: du is not needed
  }
  f.gz[(int64_t)b * 64 + j] = dh;
}

// Same for the CIL roll-out (train_functors.h: CilL1Step): x_{t-1} feeds both the residual and the
// GRU input, so the reverse sweep also needs du (a 192-term chain, done by thread 0 in the
// functor's order).
__global__ void __launch_bounds__(64) cil_l1_coop_kernel(train::CilL1Step f) {
  using namespace train;
  const int b = blockIdx.x, j = threadIdx.x;
  const DecParams& p = f.p;
  const int T = f.T;
  __shared__ float h[64], hn[64], ghs[192], gis[192], xs[2], dxs[2];
  h[j] = f.z[(int64_t)b * 64 + j];
  float* rec0 = f.scratch + (int64_t)b * T * kDecRecord;
  const float* yb = f.y + (int64_t)b * T * 2;
  const float invB = 1.0f / (float)f.B;
  float row_loss = 0.0f;  // thread 0 only
  if (j < 2) xs[j] = 0.0f;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    float* rec = rec0 + t * kDecRecord;
    float* misc = rec + kRecMisc;
    if (j < 2) misc[j] = xs[j];
    hn[j] = coop_gru_forward(p, xs, h, rec, j);
    __syncthreads();
    if (j == 0) {
      for (int d = 0; d < 2; ++d) {
        float acc = p.b1[d];
        for (int k = 0; k < 64; ++k) acc = fmaf(p.w1[d * 64 + k], hn[k], acc);
        const float x = xs[d] + acc;
        const float e = x - yb[t * 2 + d];
        misc[2 + d] = e > 0.0f ? 1.0f : (e < 0.0f ? -1.0f : 0.0f);
        row_loss += fabsf(e);
        if (f.pred) f.pred[((int64_t)b * T + t) * 2 + d] = x;
        xs[d] = x;
      }
    }
    h[j] = hn[j];
    __syncthreads();
  }
  if (j == 0) atomicAdd(f.loss, (double)row_loss);

  float dh = 0.0f;
  if (j < 2) dxs[j] = 0.0f;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    float* rec = rec0 + t * kDecRecord;
    const float* misc = rec + kRecMisc;
    if (j == 0) {
      for (int d = 0; d < 2; ++d) {
        dxs[d] += misc[2 + d] * invB;  // x_t = x_{t-1} + W_o h_t + b_o
        rec[kRecDout + d] = dxs[d];
      }
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 2; ++d) dh = fmaf(dxs[d], p.w1[d * 64 + j], dh);
    dh = coop_gru_backward(p, rec, dh, ghs, gis, j);
    if (j == 0) {  // du: the functor's chain over unit u ascending, gates r|z|n inside
      float du0 = 0.0f, du1 = 0.0f;
      for (int u = 0; u < 64; ++u) {
        for (int q = 0; q < 3; ++q) {
          const int row = q * 64 + u;
          du0 = fmaf(gis[row], p.wih[row * 2], du0);
          du1 = fmaf(gis[row], p.wih[row * 2 + 1], du1);
        }
      }
      dxs[0] += du0;  // x_{t-1} also feeds the GRU input of step t
      dxs[1] += du1;
    }
    __syncthreads();
  }
  f.gz[(int64_t)b * 64 + j] = dh;
}

// W [N][K] (reference layout, changes every step) -> W^T [K][N] for the tiled GEMM's B operand
__global__ void __launch_bounds__(256) transpose_nk_kernel(const float* __restrict__ w, float* __restrict__ wt,
                                                          int N, int K) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + threadIdx.x;
    tile[r][threadIdx.x] = (n < N && k < K) ? w[(int64_t)n * K + k] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + threadIdx.x;
    if (k < K && n < N) wt[(int64_t)k * N + n] = tile[threadIdx.x][r];
  }
}

struct CudaBackend {
  cudaStream_t stream = nullptr;
  cudaError_t status = cudaSuccess;
  // Pointwise forward / input-gradient products on the shared-memory tiled FP32 GEMM of the
  // inference path (encoder.cu: pw_gemm_kernel) instead of the work-item functors that read
  // both operands straight from global memory.  OAT_TRAIN_TILED=0 selects the functors.
  static constexpr int kMaxWt = 1280 * 320;  // largest pointwise weight (features.18)
  float* wt = nullptr;     // transposed-weight scratch
  float* zeros = nullptr;  // zero bias
  int tiled = -1;
  CudaBackend() = default;
  CudaBackend(const CudaBackend&) = delete;
  CudaBackend& operator=(const CudaBackend&) = delete;
  ~CudaBackend() {
    if (wt) cudaFree(wt);
    if (zeros) cudaFree(zeros);
  }

  bool tiled_ready() {
    if (tiled < 0) {
      const char* e = getenv("OAT_TRAIN_TILED");
#ifndef OAT_TRAIN_TILED_DEFAULT
#define OAT_TRAIN_TILED_DEFAULT 1
#endif
      tiled = e ? (atoi(e) != 0) : OAT_TRAIN_TILED_DEFAULT;
      if (tiled) {
        wt = static_cast<float*>(alloc((size_t)kMaxWt * sizeof(float)));
        zeros = static_cast<float*>(alloc(1280 * sizeof(float)));
        if (!wt || !zeros) tiled = 0;
      }
    }
    return tiled == 1;
  }
  // R[m][n] = sum_k A[m][k] W[n][k]
  bool pw_forward(const float* a, const float* w, float* r, int64_t M, int N, int K) {
    if (!tiled_ready() || K % 8 != 0 || N % 4 != 0 || (int64_t)N * K > kMaxWt || N > 1280 ||
        M > 0x7fffffff)
      return false;
    transpose_nk_kernel<<<dim3((K + 31) / 32, (N + 31) / 32), dim3(32, 8), 0, stream>>>(w, wt, N, K);
    g_launch_count++;
    if (g_profile_on) profile_mark("train_transpose", stream);
    note(cudaGetLastError());
    if (simt_pw_gemm(a, wt, zeros, nullptr, r, (int)M, K, N, 0, stream) != 0) note(cudaErrorUnknown);
    return true;
  }
  // dA[m][k] (+)= sum_n G[m][n] W[n][k]: W [N][K] is already the [K'][N'] operand (K' = N, N' = K)
  bool pw_backward_x(const float* g, const float* w, float* da, int64_t M, int N, int K, int accumulate) {
    if (!tiled_ready() || N % 8 != 0 || K % 4 != 0 || K > 1280 || M > 0x7fffffff) return false;
    if (simt_pw_gemm(g, w, zeros, accumulate ? da : nullptr, da, (int)M, N, K, 0, stream) != 0)
      note(cudaErrorUnknown);
    return true;
  }

  void* alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    if (cudaMemset(p, 0, bytes ? bytes : 1) != cudaSuccess) {
      cudaFree(p);
      return nullptr;
    }
    return p;
  }
  void free(void* p) { cudaFree(p); }
  void zero(void* p, size_t bytes) { note(cudaMemsetAsync(p, 0, bytes, stream)); }
  void note(cudaError_t e) {
    if (e != cudaSuccess && status == cudaSuccess) status = e;
  }
  // DIM decoder pass: cooperative kernel, one CTA per batch row (OAT_TRAIN_COOP=0: the work-item functor)
  void run(int64_t n, const train::DimNllStep& f) {
    static const int coop = []() { const char* e = getenv("OAT_TRAIN_COOP"); return e ? atoi(e) : 1; }();
    if (n <= 0) return;
    if (!coop) { run<train::DimNllStep>(n, f); return; }
    dim_nll_coop_kernel<<<(unsigned)n, 64, 0, stream>>>(f);
    g_launch_count++;
    if (g_profile_on) profile_mark("DimNllStep", stream);
    note(cudaGetLastError());
  }
  void run(int64_t n, const train::CilL1Step& f) {
    static const int coop = []() { const char* e = getenv("OAT_TRAIN_COOP"); return e ? atoi(e) : 1; }();
    if (n <= 0) return;
    if (!coop) { run<train::CilL1Step>(n, f); return; }
    cil_l1_coop_kernel<<<(unsigned)n, 64, 0, stream>>>(f);
    g_launch_count++;
    if (g_profile_on) profile_mark("CilL1Step", stream);
    note(cudaGetLastError());
  }
  template <class F>
  static const char* functor_tag() { return __PRETTY_FUNCTION__; }  // "... [with F = oat::train::X]"
  template <class F>
  void run(int64_t n, const F& f) {
    if (n <= 0) return;
    constexpr int kBlock = BlockSize<F>::value;
    int64_t blocks = (n + kBlock - 1) / kBlock;
    const int64_t cap = 148 * 32;
    if (blocks > cap) blocks = cap;
    run_functor<F><<<(unsigned)blocks, kBlock, 0, stream>>>(f, n);
    g_launch_count++;
    if (g_profile_on) profile_mark(functor_tag<F>(), stream);
    note(cudaGetLastError());
  }
};

}  // namespace
}  // namespace oat

struct OatTrainer {
  oat::train::TrainerT<oat::CudaBackend> impl;
  int device = 0;
};

using namespace oat;

static int trainer_device_check(int want, const char* who) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess) return fail(std::string(who) + ": no CUDA device");
  if (cur != want) return fail(std::string(who) + ": current CUDA device differs from the trainer's device");
  return 0;
}

extern "C" {

int oat_trainer_create(const OatTrainTensor* tensors, int32_t num_tensors, int32_t kind,
                       int32_t device, float* grad_flat, int64_t grad_flat_floats,
                       OatTrainer** out) {
  if (!tensors || !out) return fail("oat_trainer_create: null argument");
  if (kind != OAT_KIND_DIM && kind != OAT_KIND_CIL) return fail("oat_trainer_create: bad kind");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("oat_trainer_create: no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("oat_trainer_create: bad device index");
  OatTrainer* t = new OatTrainer();
  t->device = device;
  for (int i = 0; i < num_tensors; ++i) {
    const OatTrainTensor& s = tensors[i];
    if (!s.name) continue;
    train::TensorRef r;
    r.p = static_cast<float*>(s.param);
    r.g = static_cast<float*>(s.grad);
    for (int d = 0; d < s.ndim && d < 4; ++d) r.shape.push_back(s.shape[d]);
    if (((uintptr_t)r.p | (uintptr_t)r.g) & 15) {
      delete t;
      return fail(std::string("oat_trainer_create: `") + s.name + "` is not 16-byte aligned");
    }
    t->impl.sd[s.name] = r;
  }
  t->impl.grad_flat = grad_flat;
  t->impl.grad_flat_floats = grad_flat ? grad_flat_floats : 0;
  if (!t->impl.init(kind == OAT_KIND_CIL ? 1 : 0)) {
    const std::string e = t->impl.err;
    delete t;
    return fail("oat_trainer_create: " + e);
  }
  *out = t;
  return 0;
}

int oat_trainer_destroy(OatTrainer* trainer) {
  delete trainer;
  return 0;
}

int oat_train_forward_backward(OatTrainer* trainer, const float* visual, const float* scalars,
                               const float* target, const float* dropout_mask, int32_t B,
                               int32_t T, float* loss, float* z, float* pred, void* stream) {
  if (g_profile_on) profile_mark("(host gap before call)", (cudaStream_t)stream);
  if (!trainer || !visual || !scalars || !target || !loss)
    return fail("oat_train_forward_backward: null argument");
  if (B <= 0 || T <= 0) return fail("oat_train_forward_backward: empty batch");
  if (int rc = trainer_device_check(trainer->device, "oat_train_forward_backward")) return rc;
  auto& impl = trainer->impl;
  impl.bk.stream = (cudaStream_t)stream;
  impl.bk.status = cudaSuccess;
  if (B > impl.cap_B || T > impl.cap_T) OAT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (!impl.forward_backward(visual, scalars, target, dropout_mask, B, T, loss, z, pred))
    return fail("oat_train_forward_backward: " + impl.err);
  if (impl.bk.status != cudaSuccess)
    return fail(std::string("oat_train_forward_backward: ") + cudaGetErrorString(impl.bk.status));
  return 0;
}

int oat_trainer_activation(const OatTrainer* trainer, int32_t index, float* out, int64_t* rows,
                           int32_t* channels, void* stream) {
  if (!trainer || !rows || !channels) return fail("oat_trainer_activation: null argument");
  const float* data = nullptr;
  int ch = 0;
  if (!trainer->impl.activation(index, &data, rows, &ch))
    return fail("oat_trainer_activation: bad index, or no training step has run yet");
  *channels = ch;
  if (out)
    OAT_CUDA(cudaMemcpyAsync(out, data, (size_t)*rows * ch * sizeof(float), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
  return 0;
}

int oat_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  int32_t step, float lr, float beta1, float beta2, float eps,
                  float weight_decay, float grad_scale, float clip_norm, double* norm_ws,
                  void* stream) {
  if (n <= 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return fail("oat_adam_step: null argument");
  if (step < 1) return fail("oat_adam_step: step counts from 1");
  if (clip_norm > 0.0f && !norm_ws) return fail("oat_adam_step: clipping needs norm_ws");
  CudaBackend bk;
  bk.stream = (cudaStream_t)stream;
  const double* sumsq = nullptr;
  if (clip_norm > 0.0f) {
    bk.zero(norm_ws, sizeof(double));
    bk.run((n + 1023) / 1024, train::SumSquares{grad, norm_ws, n, grad_scale});
    sumsq = norm_ws;
  }
  const float bias1 = 1.0f - (float)std::pow((double)beta1, (double)step);
  const float bias2_sqrt = (float)std::sqrt(1.0 - std::pow((double)beta2, (double)step));
  bk.run(n, train::AdamStep{param, grad, exp_avg, exp_avg_sq, sumsq, lr, beta1, beta2, eps,
                            weight_decay, grad_scale, clip_norm, bias1, bias2_sqrt});
  if (bk.status != cudaSuccess) return fail(std::string("oat_adam_step: ") + cudaGetErrorString(bk.status));
  return 0;
}

}  // extern "C"
