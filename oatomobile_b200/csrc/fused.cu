// CUDA instantiation of the fused encoder-front bodies (fused_body.h): one CTA per
// (model, image, row split), 256 threads, the wide intermediates in shared memory.
//
// Replaces, for the first blocks of torchvision's MobileNetV2 as wrapped by
// oatomobile/torch/networks/perception.py:25-55, the separate stem / depthwise / pointwise
// launches of encoder.cu, whose 6x expanded activations otherwise make a round trip through
// HBM (blocks 1-4 are ~45 % of the encoder's time for ~4 % of its algorithmic traffic).
#include <cstdlib>

#include "common.cuh"
#include "fused.cuh"
#include "fused_body.h"

namespace oat {
namespace {

constexpr int kFusedThreads = 256;

struct CudaExec {
  float* sm;
  __device__ __forceinline__ float* smem() const { return sm; }
  __device__ __forceinline__ int nthreads() const { return kFusedThreads; }
  template <class F>
  __device__ __forceinline__ void phase(F f) {
    f((int)threadIdx.x);
    __syncthreads();
  }
  __device__ __forceinline__ void async16(float* dst, const float* src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  }
  __device__ __forceinline__ void async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
};

template <class Body>
__global__ void __launch_bounds__(kFusedThreads, 2) expand_dw_kernel(const __grid_constant__ fused::ExpandDwArgs a) {
  extern __shared__ __align__(16) float fused_smem[];
  CudaExec x{fused_smem};
  Body::run(x, a, (int)blockIdx.x);
}

__global__ void __launch_bounds__(kFusedThreads, 2) front_kernel(const __grid_constant__ fused::FrontArgs a) {
  extern __shared__ __align__(16) float fused_smem[];
  CudaExec x{fused_smem};
  fused::FrontBody::run(x, a, (int)blockIdx.x);
}

// opt in to > 48 KB of dynamic shared memory, once per (kernel, device)
template <class K>
int allow_smem(K kernel, int bytes, int* configured) {
  if (bytes > 227 * 1024) return fail("fused encoder kernel: shared-memory tile does not fit");
  int dev = 0;
  OAT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || configured[dev] < bytes) {
    OAT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev >= 0 && dev < 64) configured[dev] = bytes;
  }
  return 0;
}

// CTAs per image: default 2 (tail balance vs. the rows recomputed at a split boundary);
// OAT_FUSED_SPLITS overrides it for tuning runs.
int splits_for(int units) {
  static const int env = []() {
    const char* e = getenv("OAT_FUSED_SPLITS");
    return e ? atoi(e) : 0;
  }();
  int s = env > 0 ? env : 2;
  if (s > units) s = units;
  return s < 1 ? 1 : s;
}

fused::Weights table(const PtrTable& t) {
  fused::Weights w;
  for (int i = 0; i < fused::kMaxModels; ++i) w.p[i] = i < kMaxModels ? t.p[i] : nullptr;
  return w;
}

template <class Body>
int launch_body(const FusedBlockLaunch& l, cudaStream_t stream) {
  fused::ExpandDwArgs a;
  a.we = table(l.we); a.be = table(l.be); a.wd = table(l.wd); a.bd = table(l.bd);
  a.in = l.in; a.out = l.out; a.B = l.B;
  a.splits = splits_for(Body::GROUPS);
  const int smem = Body::kSmemFloats * (int)sizeof(float);
  static int configured[64] = {0};
  if (int rc = allow_smem(expand_dw_kernel<Body>, smem, configured)) return rc;
  const int64_t ctas = (int64_t)l.E * l.B * a.splits;
  expand_dw_kernel<Body><<<(unsigned)ctas, kFusedThreads, smem, stream>>>(a);
  OAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool fused_block_supported(int cin, int hid, int stride, int hin) {
  return (cin == 16 && hid == 96 && stride == 2 && hin == 50) ||
         (cin == 24 && hid == 144 && stride == 1 && hin == 25) ||
         (cin == 24 && hid == 144 && stride == 2 && hin == 25);
}

int launch_fused_expand_dw(const FusedBlockLaunch& l, cudaStream_t stream) {
  if (l.E <= 0 || l.B <= 0) return 0;
  //                                    CIN  HID  S  HIN OR  TP NSEG
  if (l.cin == 16 && l.hid == 96 && l.stride == 2 && l.hin == 50)
    return launch_body<fused::ExpandDwBody<16, 96, 2, 50, 1, 10, 10>>(l, stream);
  if (l.cin == 24 && l.hid == 144 && l.stride == 1 && l.hin == 25)
    return launch_body<fused::ExpandDwBody<24, 144, 1, 25, 2, 8, 7>>(l, stream);
  if (l.cin == 24 && l.hid == 144 && l.stride == 2 && l.hin == 25)
    return launch_body<fused::ExpandDwBody<24, 144, 2, 25, 1, 8, 7>>(l, stream);
  return fail("launch_fused_expand_dw: unsupported block shape");
}

int launch_fused_front(const FusedFrontLaunch& l, cudaStream_t stream) {
  if (l.E <= 0 || l.B <= 0) return 0;
  if (l.C < 1 || l.C > 8) return fail("launch_fused_front: 1 <= in_channels <= 8");
  fused::FrontArgs a;
  a.ws = table(l.ws); a.bs = table(l.bs); a.wd = table(l.wd); a.bd = table(l.bd);
  a.wp = table(l.wp); a.bp = table(l.bp);
  a.vis = l.visual; a.out = l.out; a.B = l.B; a.C = l.C;
  a.splits = splits_for(fused::FrontBody::PAIRS);
  const int smem = fused::FrontBody::smem_floats(l.C) * (int)sizeof(float);
  static int configured[64] = {0};
  if (int rc = allow_smem(front_kernel, smem, configured)) return rc;
  const int64_t ctas = (int64_t)l.E * l.B * a.splits;
  front_kernel<<<(unsigned)ctas, kFusedThreads, smem, stream>>>(a);
  OAT_LAUNCH_CHECK();
  return 0;
}

}  // namespace oat
