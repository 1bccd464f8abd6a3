// TEST TOOLING — host execution of the training-step work-item bodies.
//
// Compiles oatomobile_b200/csrc/train_functors.h + train_impl.h (the exact code the CUDA
// product runs inside grid-stride kernels) with g++ and a backend whose "launch" is a
// plain loop, so `-m "not gpu"` tests can compare every gradient with the reference's
// autograd without a GPU.  Never imported, linked or loaded by the oatomobile_b200 package.
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/oat_b200.h"
#include "../../oatomobile_b200/csrc/train_impl.h"

namespace {

struct HostBackend {
  void* alloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
  void free(void* p) { ::free(p); }
  void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
  template <class F>
  void run(int64_t n, const F& f) {
    for (int64_t i = 0; i < n; ++i) f(i);
  }
  // no tiled GEMM on the host: the work-item functors run
  bool pw_forward(const float*, const float*, float*, int64_t, int, int) { return false; }
  bool pw_backward_x(const float*, const float*, float*, int64_t, int, int, int) { return false; }
};
using Trainer = oat::train::TrainerT<HostBackend>;
std::string g_err;

}  // namespace

extern "C" {

const char* emu_last_error() { return g_err.c_str(); }

void* emu_trainer_create(const OatTrainTensor* tensors, int32_t n, int32_t kind) {
  Trainer* t = new Trainer();
  for (int i = 0; i < n; ++i) {
    oat::train::TensorRef r;
    r.p = static_cast<float*>(tensors[i].param);
    r.g = static_cast<float*>(tensors[i].grad);
    for (int d = 0; d < tensors[i].ndim && d < 4; ++d) r.shape.push_back(tensors[i].shape[d]);
    t->sd[tensors[i].name] = r;
  }
  if (!t->init(kind)) {
    g_err = t->err;
    delete t;
    return nullptr;
  }
  return t;
}

void emu_trainer_destroy(void* t) { delete static_cast<Trainer*>(t); }

int emu_forward_backward(void* tp, const float* visual, const float* scalars, const float* target,
                         const float* mask, int32_t B, int32_t T, float* loss, float* z, float* pred) {
  Trainer* t = static_cast<Trainer*>(tp);
  if (!t->forward_backward(visual, scalars, target, mask, B, T, loss, z, pred)) {
    g_err = t->err;
    return 1;
  }
  return 0;
}

int emu_activation(void* tp, int32_t index, const float** data, int64_t* rows, int32_t* channels) {
  int ch = 0;
  if (!static_cast<Trainer*>(tp)->activation(index, data, rows, &ch)) return 1;
  *channels = ch;
  return 0;
}

int emu_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr,
                  float beta1, float beta2, float eps, float wd, float grad_scale, float clip_norm) {
  using namespace oat::train;
  HostBackend bk;
  double sumsq = 0.0;
  if (clip_norm > 0.0f) bk.run((n + 1023) / 1024, SumSquares{g, &sumsq, n, grad_scale});
  const float b1 = 1.0f - (float)pow((double)beta1, step);
  const float b2s = (float)sqrt(1.0 - pow((double)beta2, step));
  bk.run(n, AdamStep{p, g, m, v, clip_norm > 0.0f ? &sumsq : nullptr, lr, beta1, beta2, eps, wd,
                     grad_scale, clip_norm, b1, b2s});
  return 0;
}
}
