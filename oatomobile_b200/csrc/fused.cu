// CUDA instantiation of the fused encoder-front bodies (fused_body.h): the executors
// (`CudaExec` 256 threads, `SmallExec` 128 threads, `PipeExec` producer/consumer halves with
// tcgen05 + TMEM + mbarriers), the kernels and their launchers.
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

// ---------------------------------------------------------------------------------------
// Tensor-core executor: 3xTF32 tcgen05 GEMMs, accumulator in TMEM (see tc_gemm.cu for the
// precision argument and the operand layout; K <= 72 here, so the fp32-accumulate truncation
// of tcgen05.mma is below 1e-6 with the small correction products issued first).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
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
// One lane of a converged warp (elect.sync); see tc_gemm.cu: behind `if (threadIdx.x == 0)` every
// tcgen05.mma is wrapped in an ELECT / BRA.U.ANY serialisation loop by ptxas.
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
// K-major, 128B-swizzled operand tile (rows of 128 B, 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

#ifndef OAT_PIPE_THREADS
#define OAT_PIPE_THREADS 512
#endif
constexpr int kPipeThreads = OAT_PIPE_THREADS;   // first half of the warps: producer, second half: consumer
constexpr int kHalf = kPipeThreads / 2;

// Executor of the pipelined tensor-core bodies (fused_body.h, "Executor interface").
struct PipeExec {
  static constexpr bool kConcurrent = true;
  float* sm;        // 1024-byte aligned dynamic shared memory
  uint32_t tmem;    // TMEM base address (lane 0, first allocated column)
  uint32_t bars;    // shared address of: [0,1] GEMM complete, [2,3] ready, [4,5] done (8 B each)
  uint32_t par0, par1;  // phase of GEMM barrier 0 / 1 the next epilogue waits for (producer only)
  __device__ __forceinline__ float* smem() const { return sm; }
  __device__ __forceinline__ bool is_producer() const { return threadIdx.x < kHalf; }
  __device__ __forceinline__ int p_threads() const { return kHalf; }
  __device__ __forceinline__ int c_threads() const { return kHalf; }
  __device__ __forceinline__ void p_barrier() const { asm volatile("bar.sync 1, %0;" ::"n"(kHalf) : "memory"); }
  template <class F>
  __device__ __forceinline__ void all_phase(F f) {
    __syncthreads();
    f((int)threadIdx.x, kPipeThreads);
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ void p_phase(F f) {
    f((int)threadIdx.x);
    p_barrier();
  }
  template <class F>
  __device__ __forceinline__ void p_phase_nosync(F f) {
    f((int)threadIdx.x);
  }
  template <class F>
  __device__ __forceinline__ void c_run(F f) {
    f((int)threadIdx.x - kHalf);
  }
  __device__ __forceinline__ void signal_ready(uint32_t i) { mbar_arrive(bars + 8u * (2u + (i & 1u))); }
  __device__ __forceinline__ void wait_ready(uint32_t i) { mbar_wait(bars + 8u * (2u + (i & 1u)), (i >> 1) & 1u); }
  __device__ __forceinline__ void signal_done(uint32_t i) { mbar_arrive(bars + 8u * (4u + (i & 1u))); }
  __device__ __forceinline__ void wait_done(uint32_t i) { mbar_wait(bars + 8u * (4u + (i & 1u)), (i >> 1) & 1u); }
  __device__ __forceinline__ void async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
  }
  __device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  __device__ __forceinline__ void async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
  __device__ __forceinline__ void async_wait_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

  // v = 4 consecutive-k elements of operand row `row`; hi tile at `tile`, lo tile `rows`*128 B behind
  __device__ __forceinline__ void op_store4(float* tile, int rows, int row, int k, fused::F4 v) {
    const uint32_t x = __float_as_uint(v.x), y = __float_as_uint(v.y), z = __float_as_uint(v.z),
                   w = __float_as_uint(v.w);
    const uint32_t hx = x & 0xffffe000u, hy = y & 0xffffe000u, hz = z & 0xffffe000u, hw = w & 0xffffe000u;
    const uint32_t lx = __float_as_uint(v.x - __uint_as_float(hx)) & 0xffffe000u;
    const uint32_t ly = __float_as_uint(v.y - __uint_as_float(hy)) & 0xffffe000u;
    const uint32_t lz = __float_as_uint(v.z - __uint_as_float(hz)) & 0xffffe000u;
    const uint32_t lw = __float_as_uint(v.w - __uint_as_float(hw)) & 0xffffe000u;
    const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)k >> 2) ^ ((uint32_t)row & 7u)) << 4);
    const uint32_t hi = smem_u32(tile) + off, lo = hi + (uint32_t)rows * 128u;
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(hi), "r"(hx), "r"(hy), "r"(hz), "r"(hw) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(lo), "r"(lx), "r"(ly), "r"(lz), "r"(lw) : "memory");
  }

  // producer half: D[acc .. acc+N) = A B^T  (A: 128 rows, B: N rows, K-major, ceil(K/32)
  // k-blocks; K % 8 == 0).  Asynchronous: completion (barrier `buf`) is observed by epilogue().
  __device__ __forceinline__ void mma(int buf, int acc, int N, const float* a_tile, const float* b_tile, int K) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy tile writes -> tensor core
    p_barrier();
    if (threadIdx.x < 32 && elect_one_sync()) {  // warp 0 arrives converged from the barrier
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(fused::kTileRows >> 4) << 24);
      const uint32_t d = tmem + (uint32_t)acc;
      const uint32_t a0 = smem_u32(a_tile), b0 = smem_u32(b_tile);
      const uint32_t a_kb = (uint32_t)fused::kATileFloats * 4u, b_kb = (uint32_t)fused::b_tile_floats(N) * 4u;
      const uint32_t a_lo = (uint32_t)fused::kTileRows * 128u, b_lo = (uint32_t)N * 128u;
      const int kbs = (K + 31) / 32;
      uint32_t accum = 0;
      for (int kb = 0; kb < kbs; ++kb) {  // correction products first (2^-11 of the main one)
        const int left = K - 32 * kb, slices = left >= 32 ? 4 : left / 8;
        const uint64_t dAh = make_desc_sw128(a0 + kb * a_kb), dAl = make_desc_sw128(a0 + kb * a_kb + a_lo);
        const uint64_t dBh = make_desc_sw128(b0 + kb * b_kb), dBl = make_desc_sw128(b0 + kb * b_kb + b_lo);
        for (int ks = 0; ks < slices; ++ks) {
          const uint64_t off = (uint64_t)(2 * ks);  // 8 tf32 = 32 B = 2 x 16 B along K
          umma_tf32(d, dAl + off, dBh + off, idesc, accum);
          accum = 1;
          umma_tf32(d, dAh + off, dBl + off, idesc, 1u);
        }
      }
      for (int kb = 0; kb < kbs; ++kb) {
        const int left = K - 32 * kb, slices = left >= 32 ? 4 : left / 8;
        const uint64_t dAh = make_desc_sw128(a0 + kb * a_kb), dBh = make_desc_sw128(b0 + kb * b_kb);
        for (int ks = 0; ks < slices; ++ks) umma_tf32(d, dAh + (uint64_t)(2 * ks), dBh + (uint64_t)(2 * ks), idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       bars + 8u * (uint32_t)buf)
                   : "memory");
    }
  }

  // producer half: emit(row, c4, F4) for accumulator row = 32*(warp%4)+lane and every
  // (kHalf/128)-th 16-column chunk; a producer barrier must follow before the accumulator or
  // the operand tile is reused (the bodies' collect() ends with one)
  template <class Emit>
  __device__ __forceinline__ void epilogue(int buf, int acc, int N, Emit emit) {
    if (buf == 0) {
      mbar_wait(bars, par0);
      par0 ^= 1u;
    } else {
      mbar_wait(bars + 8u, par1);
      par1 ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
    const int quad = warp & 3, grp = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc;
    for (int c = grp; c < N / 16; c += kHalf / 128) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
            "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
            "=r"(r[14]), "=r"(r[15])
          : "r"(taddr + (uint32_t)(16 * c)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 4; ++j)
        emit(row, 4 * c + j,
             fused::F4{__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])});
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
};

template <class Body>
__global__ void __launch_bounds__(kPipeThreads, 1) expand_dw_tc_kernel(const __grid_constant__ fused::ExpandDwArgs a) {
  extern __shared__ __align__(16) uint8_t tc_smem_raw[];
  __shared__ uint64_t bars[6];
  __shared__ uint32_t tmem_slot;
  const int warp = (int)threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    for (int i = 2; i < 6; ++i) mbar_init(smem_u32(&bars[i]), kHalf);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // TMEM allocation is a warp-wide operation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(Body::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // 1024-byte alignment by an offset INTO the shared array (not an integer round trip): the
  // compiler keeps the shared address space, so the bodies' loads/stores are LDS/STS and not
  // generic LD/ST (measured: generic accesses made this kernel latency-bound on the long scoreboard)
  const uint32_t pad = (1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u;
  PipeExec x{reinterpret_cast<float*>(tc_smem_raw + pad), tmem_slot, smem_u32(&bars[0]), 0u, 0u};
  Body::run(x, a, (int)blockIdx.x, (int)gridDim.x);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(x.tmem), "r"(Body::kTmemCols) : "memory");
  }
}

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

constexpr int kDwProjectThreads = 128;
struct SmallExec {  // same interface as CudaExec for a 128-thread CTA
  float* sm;
  __device__ __forceinline__ float* smem() const { return sm; }
  __device__ __forceinline__ int nthreads() const { return kDwProjectThreads; }
  template <class F>
  __device__ __forceinline__ void phase(F f) {
    f((int)threadIdx.x);
    __syncthreads();
  }
};

__global__ void __launch_bounds__(kDwProjectThreads, 6) dw_project_kernel(const __grid_constant__ fused::DwProjectArgs a) {
  __shared__ __align__(16) float dwp_smem[fused::DwProjectBody::kSmemFloats];
  SmallExec x{dwp_smem};
  fused::DwProjectBody::run(x, a, (int)blockIdx.x);
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

// Work units per image.  Measured on B200 (profiles/r1_fused_*): the FP32 kernels are fastest
// with one CTA per image (no rows recomputed at split boundaries), the persistent tensor-core
// kernel with two units per image (tail balance).  OAT_FUSED_SPLITS overrides both for tuning.
int splits_for(int units, int dflt) {
  static const int env = []() {
    const char* e = getenv("OAT_FUSED_SPLITS");
    return e ? atoi(e) : 0;
  }();
  int s = env > 0 ? env : dflt;
  if (s > units) s = units;
  return s < 1 ? 1 : s;
}

fused::Weights table(const PtrTable& t) {
  fused::Weights w;
  for (int i = 0; i < fused::kMaxModels; ++i) w.p[i] = i < kMaxModels ? t.p[i] : nullptr;
  return w;
}

template <class Body>
int launch_tc_body(const FusedBlockLaunch& l, cudaStream_t stream) {
  fused::ExpandDwArgs a;
  a.we = table(l.we); a.be = table(l.be); a.wd = table(l.wd); a.bd = table(l.bd);
  a.in = l.in; a.out = l.out; a.B = l.B;
  a.splits = splits_for(Body::GROUPS, 2);
  const int64_t units = (int64_t)l.E * l.B * a.splits;
  if (units > 0x7fffffff) return fail("fused encoder kernel: batch too large");
  a.units = (int)units;
  const int smem = Body::kSmemFloats * (int)sizeof(float) + 1024;  // + alignment slack
  static int configured[64] = {0};
  if (int rc = allow_smem(expand_dw_tc_kernel<Body>, smem, configured)) return rc;
  static int num_sms[64] = {0};
  int dev = 0;
  OAT_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && num_sms[dev] == 0)
    OAT_CUDA(cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  const int sms = (dev >= 0 && dev < 64) ? num_sms[dev] : 148;
  // persistent: one CTA per SM walks the units with a stride (weights reloaded per model only)
  const int grid = a.units < sms ? a.units : sms;
  expand_dw_tc_kernel<Body><<<grid, kPipeThreads, smem, stream>>>(a);
  OAT_LAUNCHED("fused_expand_dw");
  return 0;
}

template <class Body>
int launch_body(const FusedBlockLaunch& l, cudaStream_t stream) {
  fused::ExpandDwArgs a;
  a.we = table(l.we); a.be = table(l.be); a.wd = table(l.wd); a.bd = table(l.bd);
  a.in = l.in; a.out = l.out; a.B = l.B;
  a.splits = splits_for(Body::GROUPS, 1);
  a.units = l.E * l.B * a.splits;
  const int smem = Body::kSmemFloats * (int)sizeof(float);
  static int configured[64] = {0};
  if (int rc = allow_smem(expand_dw_kernel<Body>, smem, configured)) return rc;
  const int64_t ctas = (int64_t)l.E * l.B * a.splits;
  expand_dw_kernel<Body><<<(unsigned)ctas, kFusedThreads, smem, stream>>>(a);
  OAT_LAUNCHED("fused_expand_dw");
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
  if (l.tensor_cores == 2 || (l.tensor_cores == 1 && l.hin == 50)) {
    // tensor_cores == 1 ("auto"): the pipelined tcgen05 kernel where it measured faster than the
    // FP32 one (features.2: 50x50 input, 100 pixels per UMMA); == 2 forces it for every shape.
    //                                          CIN  HID  S  HIN OR NSEG
    if (l.cin == 16 && l.hid == 96 && l.stride == 2 && l.hin == 50)
      return launch_tc_body<fused::ExpandDwPipeBody<16, 96, 2, 50, 1, 10>>(l, stream);
    if (l.cin == 24 && l.hid == 144 && l.stride == 1 && l.hin == 25)
      return launch_tc_body<fused::ExpandDwPipeBody<24, 144, 1, 25, 2, 7>>(l, stream);
    if (l.cin == 24 && l.hid == 144 && l.stride == 2 && l.hin == 25)
      return launch_tc_body<fused::ExpandDwPipeBody<24, 144, 2, 25, 1, 7>>(l, stream);
    return fail("launch_fused_expand_dw: unsupported block shape");
  }
  //                                    CIN  HID  S  HIN OR  TP NSEG
  if (l.cin == 16 && l.hid == 96 && l.stride == 2 && l.hin == 50)
    return launch_body<fused::ExpandDwBody<16, 96, 2, 50, 1, 10, 10>>(l, stream);
  if (l.cin == 24 && l.hid == 144 && l.stride == 1 && l.hin == 25)
    return launch_body<fused::ExpandDwBody<24, 144, 1, 25, 2, 8, 7>>(l, stream);
  if (l.cin == 24 && l.hid == 144 && l.stride == 2 && l.hin == 25)
    return launch_body<fused::ExpandDwBody<24, 144, 2, 25, 1, 8, 7>>(l, stream);
  return fail("launch_fused_expand_dw: unsupported block shape");
}

int launch_fused_dw_project(const FusedDwProjectLaunch& l, cudaStream_t stream) {
  if (l.E <= 0 || l.B <= 0) return 0;
  fused::DwProjectArgs a;
  a.wd = table(l.wd); a.bd = table(l.bd); a.wp = table(l.wp); a.bp = table(l.bp);
  a.in = l.in; a.out = l.out; a.B = l.B;
  const int64_t ctas = (int64_t)l.E * l.B * fused::DwProjectBody::PAIRS;
  if (ctas > 0x7fffffff) return fail("fused encoder kernel: batch too large");
  dw_project_kernel<<<(unsigned)ctas, kDwProjectThreads, 0, stream>>>(a);
  OAT_LAUNCHED("fused_dw_project");
  return 0;
}

int launch_fused_front(const FusedFrontLaunch& l, cudaStream_t stream) {
  if (l.E <= 0 || l.B <= 0) return 0;
  if (l.C < 1 || l.C > 8) return fail("launch_fused_front: 1 <= in_channels <= 8");
  fused::FrontArgs a;
  a.ws = table(l.ws); a.bs = table(l.bs); a.wd = table(l.wd); a.bd = table(l.bd);
  a.wp = table(l.wp); a.bp = table(l.bp);
  a.vis = l.visual; a.out = l.out; a.B = l.B; a.C = l.C;
  a.splits = splits_for(fused::FrontBody::PAIRS, 1);
  const int smem = fused::FrontBody::smem_floats(l.C) * (int)sizeof(float);
  static int configured[64] = {0};
  if (int rc = allow_smem(front_kernel, smem, configured)) return rc;
  const int64_t ctas = (int64_t)l.E * l.B * a.splits;
  front_kernel<<<(unsigned)ctas, kFusedThreads, smem, stream>>>(a);
  OAT_LAUNCHED("fused_front");
  return 0;
}

}  // namespace oat
