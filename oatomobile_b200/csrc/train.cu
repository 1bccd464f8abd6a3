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
