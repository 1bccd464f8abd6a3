// BEV encoder kernels (sm_100a), FP32 SIMT path.
//
// Replaces oatomobile/torch/transforms.py:34-49 (bilinear resize + transpose),
// oatomobile/torch/networks/perception.py:53-55 (torchvision MobileNetV2 forward,
// eval mode, BatchNorm folded at pack time) and the merger MLP of
// oatomobile/baselines/torch/dim/model.py:203-217.
//
// Activations are NHWC fp32, grouped over the E models of the ensemble on
// gridDim.z/.y so one launch serves every model ([E][B][H][W][C]); the 34
// pointwise convolutions are [M=B*H*W, K] x [K, N] GEMMs with a fused
// bias(+BN) / ReLU6 / residual epilogue.
#include "common.cuh"
#include "fused.cuh"
#include "tc_gemm.cuh"

namespace oat {
namespace {

__device__ __forceinline__ float relu6f(float v) { return fminf(fmaxf(v, 0.0f), 6.0f); }

// ---------------------------------------------------------------------------
// a1. bilinear 2x-ish resize (align_corners=True) fused with the H<->W swap.
// out[b,c,i,j] = R[b,c,j,i],  R[p,q] = bilinear(src, p*(H-1)/(OH-1), q*(W-1)/(OW-1)).
// ---------------------------------------------------------------------------
template <bool HWC>
__global__ void __launch_bounds__(256) transform_visual_kernel(
    const float* __restrict__ lidar, int BC, int C, int H, int W, float* __restrict__ out,
    float scale_h, float scale_w) {
  constexpr int O = 100;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)BC * O * O) return;
  const int j = (int)(idx % O);          // output column  = resized row index p
  const int i = (int)((idx / O) % O);    // output row     = resized column index q
  const int64_t bc = idx / (O * O);
  const int p = j, qq = i;
  const float sy = scale_h * (float)p, sx = scale_w * (float)qq;
  const int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float hy = 1.0f - ly, hx = 1.0f - lx;
  // NCHW: plane (b,c) is contiguous; HWC (the on-disk / simulator layout): channel-strided
  const int64_t b = bc / C;
  const int c = (int)(bc % C);
  const float* src = HWC ? lidar + b * H * W * C + c : lidar + bc * H * W;
  const int es = HWC ? C : 1;
  const float v00 = __ldg(src + (int64_t)(y0 * W + x0) * es), v01 = __ldg(src + (int64_t)(y0 * W + x1) * es);
  const float v10 = __ldg(src + (int64_t)(y1 * W + x0) * es), v11 = __ldg(src + (int64_t)(y1 * W + x1) * es);
  // same association as ATen: hy*(hx*v00 + lx*v01) + ly*(hx*v10 + lx*v11)
  out[idx] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

// Tiled form for NCHW sources: the kernel above reads the source transposed (consecutive
// threads walk source ROWS), 104 us for 164 MB at B=256, C=4 = a quarter of the HBM rate.  Here a
// CTA stages the source window of a 25x25 output tile in shared memory with row-contiguous loads
// and reads it transposed from there.  Same arithmetic, bit-identical results.
constexpr int TV_T = 25;   // output tile edge (100 = 4 tiles)
constexpr int TV_S = 56;   // largest staged source window edge (scale <= ~2.1)

__global__ void __launch_bounds__(256) transform_visual_tiled_kernel(
    const float* __restrict__ lidar, int H, int W, float* __restrict__ out, float scale_h,
    float scale_w) {
  constexpr int O = 100;
  __shared__ float tile[TV_S][TV_S + 1];
  const int plane = blockIdx.z;
  const int j0 = blockIdx.x * TV_T, i0 = blockIdx.y * TV_T;   // j <-> source row p, i <-> source column q
  const int yb = min((int)(scale_h * (float)j0), H - 1), xb = min((int)(scale_w * (float)i0), W - 1);
  const int ye = min(min((int)(scale_h * (float)(j0 + TV_T - 1)), H - 1) + 1, H - 1);
  const int xe = min(min((int)(scale_w * (float)(i0 + TV_T - 1)), W - 1) + 1, W - 1);
  const int ny = ye - yb + 1, nx = xe - xb + 1;
  const float* __restrict__ src = lidar + (int64_t)plane * H * W;
  for (int t = threadIdx.x; t < ny * nx; t += 256) {
    const int r = t / nx, c = t - r * nx;
    tile[r][c] = __ldg(src + (int64_t)(yb + r) * W + xb + c);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < TV_T * TV_T; t += 256) {
    const int di = t / TV_T, dj = t - di * TV_T;
    const int i = i0 + di, j = j0 + dj;
    const float sy = scale_h * (float)j, sx = scale_w * (float)i;
    const int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const float v00 = tile[y0 - yb][x0 - xb], v01 = tile[y0 - yb][x1 - xb];
    const float v10 = tile[y1 - yb][x0 - xb], v11 = tile[y1 - yb][x1 - xb];
    out[((int64_t)plane * O + i) * O + j] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  }
}

// ---------------------------------------------------------------------------
// stem: 3x3 stride-2 pad-1 conv C->32 + folded BN + ReLU6 (features.0).
// in NCHW [B,C,100,100] (shared by all models) -> out [E][B][50][50][32].
// ---------------------------------------------------------------------------
constexpr int STEM_THREADS = 128;
constexpr int STEM_MAXC = 8;

__global__ void __launch_bounds__(STEM_THREADS) stem_kernel(
    const __grid_constant__ PtrTable w, const __grid_constant__ PtrTable bias, const float* __restrict__ vis, int B, int C,
    float* __restrict__ out) {
  __shared__ __align__(16) float ws[9 * STEM_MAXC * 32];
  __shared__ __align__(16) float bs[32];
  const int model = blockIdx.y;
  for (int i = threadIdx.x; i < 9 * C * 32; i += STEM_THREADS) ws[i] = __ldg(w.p[model] + i);
  if (threadIdx.x < 32) bs[threadIdx.x] = __ldg(bias.p[model] + threadIdx.x);
  __syncthreads();
#if !defined(OAT_STEM_1PX)
  // One thread = two horizontally adjacent output pixels (every weight float4 feeds 8 FMAs;
  // -DOAT_STEM_1PX selects the one-pixel form).
  // Loads are unconditional on clamped coordinates and masked by a 0/1 factor.
  const int64_t pair = (int64_t)blockIdx.x * STEM_THREADS + threadIdx.x;
  if (pair >= (int64_t)B * 1250) return;
  const int ow = 2 * (int)(pair % 25), oh = (int)((pair / 25) % 50);
  const int64_t b = pair / 1250;
  float acc0[32], acc1[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { acc0[i] = bs[i]; acc1[i] = bs[i]; }
  for (int c = 0; c < C; ++c) {
    const float* plane = vis + (b * C + c) * 10000;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = 2 * oh - 1 + kh;
      const int ihc = min(max(ih, 0), 99);
      const float rmask = (ih == ihc) ? 1.0f : 0.0f;
      float v[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int iw = 2 * ow - 1 + j;
        const int iwc = min(max(iw, 0), 99);
        v[j] = __ldg(plane + ihc * 100 + iwc) * ((iw == iwc) ? rmask : 0.0f);
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4* wr = reinterpret_cast<const float4*>(ws + ((kh * 3 + kw) * C + c) * 32);
        const float va = v[kw], vb = v[kw + 2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 wv = wr[j];
          acc0[4 * j + 0] = fmaf(va, wv.x, acc0[4 * j + 0]);
          acc0[4 * j + 1] = fmaf(va, wv.y, acc0[4 * j + 1]);
          acc0[4 * j + 2] = fmaf(va, wv.z, acc0[4 * j + 2]);
          acc0[4 * j + 3] = fmaf(va, wv.w, acc0[4 * j + 3]);
          acc1[4 * j + 0] = fmaf(vb, wv.x, acc1[4 * j + 0]);
          acc1[4 * j + 1] = fmaf(vb, wv.y, acc1[4 * j + 1]);
          acc1[4 * j + 2] = fmaf(vb, wv.z, acc1[4 * j + 2]);
          acc1[4 * j + 3] = fmaf(vb, wv.w, acc1[4 * j + 3]);
        }
      }
    }
  }
  const int64_t pix = b * 2500 + (int64_t)oh * 50 + ow;
  float4* dst = reinterpret_cast<float4*>(out + ((int64_t)model * B * 2500 + pix) * 32);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dst[j] = make_float4(relu6f(acc0[4 * j]), relu6f(acc0[4 * j + 1]), relu6f(acc0[4 * j + 2]),
                         relu6f(acc0[4 * j + 3]));
    dst[8 + j] = make_float4(relu6f(acc1[4 * j]), relu6f(acc1[4 * j + 1]), relu6f(acc1[4 * j + 2]),
                             relu6f(acc1[4 * j + 3]));
  }
}
#else
  const int64_t pix = (int64_t)blockIdx.x * STEM_THREADS + threadIdx.x;
  if (pix >= (int64_t)B * 2500) return;
  const int ow = (int)(pix % 50), oh = (int)((pix / 50) % 50);
  const int64_t b = pix / 2500;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = bs[i];
  for (int c = 0; c < C; ++c) {
    const float* plane = vis + (b * C + c) * 10000;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = 2 * oh - 1 + kh;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = 2 * ow - 1 + kw;
        float v = 0.0f;
        if (ih >= 0 && ih < 100 && iw >= 0 && iw < 100) v = __ldg(plane + ih * 100 + iw);
        const float4* wr = reinterpret_cast<const float4*>(ws + ((kh * 3 + kw) * C + c) * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 wv = wr[j];
          acc[4 * j + 0] = fmaf(v, wv.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(v, wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(v, wv.w, acc[4 * j + 3]);
        }
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(out + ((int64_t)model * B * 2500 + pix) * 32);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    dst[j] = make_float4(relu6f(acc[4 * j]), relu6f(acc[4 * j + 1]), relu6f(acc[4 * j + 2]),
                         relu6f(acc[4 * j + 3]));
}
#endif

// ---------------------------------------------------------------------------
// pointwise (1x1) convolution = GEMM  C[M,N] = act(A[M,K] W[K,N] + bias) (+ R).
// 256 threads, BK = 8, register-prefetched double-buffered smem tiles.
// ---------------------------------------------------------------------------
struct PwArgs {
  PtrTable w, bias;
  const float* A;
  float* C;
  const float* R;        // residual or null
  int64_t a_stride;      // floats between models in A (0 = shared input)
  int64_t c_stride;      // floats between models in C / R
  int M, K, N;
  int relu6;
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) pw_gemm_kernel(const __grid_constant__ PwArgs a) {
  constexpr int BK = 8;
  constexpr int NTX = BN / TN;            // threads along n
  constexpr int AS_LD = BM + 4;           // padded leading dim of the transposed A tile
  constexpr int A_F4 = BM * BK / 4 / 256; // float4 loads of A per thread
  static_assert((BM / TM) * NTX == 256, "tile/thread mismatch");
  static_assert(A_F4 >= 1, "A tile too small");
  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int model = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const float* __restrict__ A = a.A + (int64_t)model * a.a_stride;
  const float* __restrict__ W = a.w.p[model];
  const int M = a.M, K = a.K, N = a.N;

  const int tx = tid % NTX, ty = tid / NTX;

  float4 areg[A_F4];
  float4 breg = make_float4(0.f, 0.f, 0.f, 0.f);
  const int b_k = tid / (BN / 4), b_n4 = tid % (BN / 4);
  const bool b_active = tid < BK * BN / 4;

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      const int f = tid + i * 256;
      const int row = f / 2, kq = f % 2;
      const int gm = m0 + row;
      areg[i] = (gm < M) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)gm * K + k0 + 4 * kq))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (b_active) {
      const int gn = n0 + 4 * b_n4;
      breg = (gn < N) ? __ldg(reinterpret_cast<const float4*>(W + (int64_t)(k0 + b_k) * N + gn))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
      const int f = tid + i * 256;
      const int row = f / 2, kq = f % 2;
      As[buf][4 * kq + 0][row] = areg[i].x;
      As[buf][4 * kq + 1][row] = areg[i].y;
      As[buf][4 * kq + 2][row] = areg[i].z;
      As[buf][4 * kq + 3][row] = areg[i].w;
    }
    if (b_active) *reinterpret_cast<float4*>(&Bs[buf][b_k][4 * b_n4]) = breg;
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  const int nk = K / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float af[TM], bf[TN];
#pragma unroll
      for (int i = 0; i < TM / 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4 * i]);
        af[4 * i] = v.x; af[4 * i + 1] = v.y; af[4 * i + 2] = v.z; af[4 * i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN / 4; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + 4 * j]);
        bf[4 * j] = v.x; bf[4 * j + 1] = v.y; bf[4 * j + 2] = v.z; bf[4 * j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: folded-BN bias, ReLU6, residual ------------------------------
  const float* __restrict__ bias = a.bias.p[model];
  float* __restrict__ C = a.C + (int64_t)model * a.c_stride;
  const float* __restrict__ R = a.R ? a.R + (int64_t)model * a.c_stride : nullptr;
#pragma unroll
  for (int j = 0; j < TN / 4; ++j) {
    const int gn = n0 + tx * TN + 4 * j;
    if (gn >= N) continue;
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + gn));
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int gm = m0 + ty * TM + i;
      if (gm >= M) continue;
      float4 v = make_float4(acc[i][4 * j] + bv.x, acc[i][4 * j + 1] + bv.y,
                             acc[i][4 * j + 2] + bv.z, acc[i][4 * j + 3] + bv.w);
      if (a.relu6) { v.x = relu6f(v.x); v.y = relu6f(v.y); v.z = relu6f(v.z); v.w = relu6f(v.w); }
      if (R) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(R + (int64_t)gm * N + gn));
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      *reinterpret_cast<float4*>(C + (int64_t)gm * N + gn) = v;
    }
  }
}

int launch_pw(const PwArgs& a, int E, cudaStream_t stream, const TcLayer* tc = nullptr) {
  if (tc != nullptr) {  // tcgen05 / TMA path (3xTF32)
    TcGemmProblem p;
    p.A = a.A; p.Wh = tc->wh; p.Wl = tc->wl; p.Wr = tc->wr; p.bias = tc->bias; p.R = a.R; p.C = a.C;
    p.M = a.M; p.K = a.K; p.N = a.N; p.E = E; p.relu6 = a.relu6;
    return tc_pw_gemm(p, stream);
  }
  if (a.K % 8 != 0 || a.N % 4 != 0) return fail("pw_gemm: K must be a multiple of 8, N of 4");
  if (a.N <= 16) {
    dim3 grid((a.M + 255) / 256, (a.N + 15) / 16, E);
    pw_gemm_kernel<256, 16, 4, 4><<<grid, 256, 0, stream>>>(a);
  } else if (a.N <= 32 || a.N == 96) {
    dim3 grid((a.M + 127) / 128, (a.N + 31) / 32, E);
    pw_gemm_kernel<128, 32, 4, 4><<<grid, 256, 0, stream>>>(a);
  } else {
    dim3 grid((a.M + 127) / 128, (a.N + 63) / 64, E);
    pw_gemm_kernel<128, 64, 8, 4><<<grid, 256, 0, stream>>>(a);
  }
  OAT_LAUNCHED("simt_pw_gemm");
  return 0;
}

// ---------------------------------------------------------------------------
// depthwise 3x3 (pad 1, stride 1|2) + folded BN + ReLU6, NHWC.
// One thread = 4 channels (float4) x one whole output row: the 9 taps stay in
// registers and the 3x3 window slides along W, so each output costs 3 (stride 1)
// or 6 (stride 2) 16-byte loads instead of 9 + 9 weight loads.  HBM-bound.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float4 fma4(const float4 v, const float4 k, const float4 acc) {
  return make_float4(fmaf(v.x, k.x, acc.x), fmaf(v.y, k.y, acc.y), fmaf(v.z, k.z, acc.z),
                     fmaf(v.w, k.w, acc.w));
}

#ifndef OAT_DW_THREADS
#define OAT_DW_THREADS 128
#endif
template <int STRIDE>
__global__ void __launch_bounds__(OAT_DW_THREADS) dw_kernel(const __grid_constant__ PtrTable w,
                                                 const __grid_constant__ PtrTable bias,
                                                 const float* __restrict__ in,
                                                 float* __restrict__ out, int B, int Hin,
                                                 int Hout, int C) {
  const int model = blockIdx.y;
  const int C4 = C >> 2;
  const int64_t total = (int64_t)B * Hout * C4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  const int oh = (int)((idx / C4) % Hout);
  const int64_t b = idx / ((int64_t)C4 * Hout);
  const float* __restrict__ wm = w.p[model] + 4 * c4;
  float4 k[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) k[t] = __ldg(reinterpret_cast<const float4*>(wm + t * C));
  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias.p[model] + 4 * c4));
  const float* __restrict__ src = in + ((int64_t)model * B + b) * Hin * Hin * C + 4 * c4;
  float* __restrict__ dst = out + (((int64_t)model * B + b) * Hout + oh) * Hout * C + 4 * c4;
  const int ih0 = oh * STRIDE - 1;
  const bool rv[3] = {ih0 >= 0, true, ih0 + 2 < Hin};
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  auto ldcol = [&](int iw, float4 (&col)[3]) {
    const bool cv = iw >= 0 && iw < Hin;
#pragma unroll
    for (int r = 0; r < 3; ++r)
      col[r] = (cv && rv[r])
                   ? __ldg(reinterpret_cast<const float4*>(src + ((int64_t)(ih0 + r) * Hin + iw) * C))
                   : zero;
  };
  auto emit = [&](int ow, const float4 (&A)[3], const float4 (&Bc)[3], const float4 (&Cc)[3]) {
    float4 acc = bv;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      acc = fma4(A[r], k[3 * r + 0], acc);
      acc = fma4(Bc[r], k[3 * r + 1], acc);
      acc = fma4(Cc[r], k[3 * r + 2], acc);
    }
    acc.x = relu6f(acc.x); acc.y = relu6f(acc.y); acc.z = relu6f(acc.z); acc.w = relu6f(acc.w);
    *reinterpret_cast<float4*>(dst + (int64_t)ow * C) = acc;
  };
  // Two outputs per iteration: all 6 (stride 1) / 12 (stride 2) column loads of the pair are
  // issued before the first FMA, so each thread keeps them in flight together.
  float4 q0[3], q1[3], q2[3], q3[3], q4[3];
  if (STRIDE == 1) {
    ldcol(-1, q0);
    ldcol(0, q1);
    for (int ow = 0; ow < Hout; ow += 2) {
      ldcol(ow + 1, q2);
      ldcol(ow + 2, q3);
      emit(ow, q0, q1, q2);
      if (ow + 1 < Hout) emit(ow + 1, q1, q2, q3);
#pragma unroll
      for (int r = 0; r < 3; ++r) { q0[r] = q2[r]; q1[r] = q3[r]; }
    }
  } else {
    ldcol(-1, q0);
    for (int ow = 0; ow < Hout; ow += 2) {
      ldcol(2 * ow, q1);
      ldcol(2 * ow + 1, q2);
      ldcol(2 * ow + 2, q3);
      ldcol(2 * ow + 3, q4);
      emit(ow, q0, q1, q2);
      if (ow + 1 < Hout) emit(ow + 1, q2, q3, q4);
#pragma unroll
      for (int r = 0; r < 3; ++r) q0[r] = q4[r];
    }
  }
}

#if !defined(OAT_DW_1ROW)
// Stride 1: one thread = 4 channels x TWO output rows; the 4 input rows are loaded once per
// column (4 loads for 2 outputs instead of 6).  -DOAT_DW_1ROW selects the one-row kernel.
__global__ void __launch_bounds__(OAT_DW_THREADS) dw2_kernel(const __grid_constant__ PtrTable w,
                                                            const __grid_constant__ PtrTable bias,
                                                            const float* __restrict__ in,
                                                            float* __restrict__ out, int B, int H,
                                                            int C) {
  const int model = blockIdx.y;
  const int C4 = C >> 2;
  const int HP = (H + 1) >> 1;  // row pairs
  const int64_t total = (int64_t)B * HP * C4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  const int oh = 2 * (int)((idx / C4) % HP);
  const int64_t b = idx / ((int64_t)C4 * HP);
  const float* __restrict__ wm = w.p[model] + 4 * c4;
  float4 k[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) k[t] = __ldg(reinterpret_cast<const float4*>(wm + t * C));
  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias.p[model] + 4 * c4));
  const float* __restrict__ src = in + ((int64_t)model * B + b) * H * H * C + 4 * c4;
  float* __restrict__ dst = out + (((int64_t)model * B + b) * H + oh) * H * C + 4 * c4;
  const bool second = oh + 1 < H;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  auto ldcol = [&](int iw, float4 (&col)[4]) {
    const bool cv = iw >= 0 && iw < H;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ih = oh - 1 + r;
      col[r] = (cv && ih >= 0 && ih < H)
                   ? __ldg(reinterpret_cast<const float4*>(src + ((int64_t)ih * H + iw) * C))
                   : zero;
    }
  };
  float4 L[4], M[4], R[4];
  ldcol(-1, L);
  ldcol(0, M);
  for (int ow = 0; ow < H; ++ow) {
    ldcol(ow + 1, R);
    float4 a0 = bv, a1 = bv;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      a0 = fma4(L[r], k[3 * r + 0], a0); a0 = fma4(M[r], k[3 * r + 1], a0); a0 = fma4(R[r], k[3 * r + 2], a0);
      a1 = fma4(L[r + 1], k[3 * r + 0], a1); a1 = fma4(M[r + 1], k[3 * r + 1], a1); a1 = fma4(R[r + 1], k[3 * r + 2], a1);
    }
    a0.x = relu6f(a0.x); a0.y = relu6f(a0.y); a0.z = relu6f(a0.z); a0.w = relu6f(a0.w);
    *reinterpret_cast<float4*>(dst + (int64_t)ow * C) = a0;
    if (second) {
      a1.x = relu6f(a1.x); a1.y = relu6f(a1.y); a1.z = relu6f(a1.z); a1.w = relu6f(a1.w);
      *reinterpret_cast<float4*>(dst + ((int64_t)H + ow) * C) = a1;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) { L[r] = M[r]; M[r] = R[r]; }
  }
}
#endif

// ---------------------------------------------------------------------------
// global average pool over P pixels: [E*B][P][C] -> [E*B][C].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ in,
                                                   float* __restrict__ out, int64_t rows,
                                                   int P, int C) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * C) return;
  const int c = (int)(idx % C);
  const int64_t r = idx / C;
  const float* src = in + r * P * C + c;
  float s = 0.0f;
  for (int p = 0; p < P; ++p) s += __ldg(src + (int64_t)p * C);
  out[idx] = s / (float)P;
}

// ---------------------------------------------------------------------------
// merger MLP: z = relu(W2 relu(W1 relu(W0 [feat, scalars] + b0) + b1) + b2)
// (dim/model.py:205-217, mlp.py:49-66).  One CTA of 64 threads per (model,row).
// ---------------------------------------------------------------------------
struct MergerArgs {
  PtrTable w0, b0, w1, b1, w2, b2;  // transposed weights [in][64]
  const float* feat;     // [E][B][128]
  const float* scalars;  // [B][S]
  int B, S;
  float* z;              // [E][B][64]
};

__global__ void __launch_bounds__(64) merger_kernel(const __grid_constant__ MergerArgs a) {
  __shared__ float u[OAT_ENC_FEATURES + 8];
  __shared__ float v[64];
  const int model = blockIdx.y, b = blockIdx.x, j = threadIdx.x;
  const int in0 = OAT_ENC_FEATURES + a.S;
  const float* f = a.feat + ((int64_t)model * a.B + b) * OAT_ENC_FEATURES;
  u[j] = __ldg(f + j);
  u[j + 64] = __ldg(f + j + 64);
  if (j < a.S) u[OAT_ENC_FEATURES + j] = __ldg(a.scalars + (int64_t)b * a.S + j);
  __syncthreads();
  float acc = __ldg(a.b0.p[model] + j);
  const float* w = a.w0.p[model];
  for (int i = 0; i < in0; ++i) acc = fmaf(u[i], __ldg(w + i * 64 + j), acc);
  v[j] = fmaxf(acc, 0.0f);
  __syncthreads();
  acc = __ldg(a.b1.p[model] + j);
  w = a.w1.p[model];
  for (int i = 0; i < 64; ++i) acc = fmaf(v[i], __ldg(w + i * 64 + j), acc);
  __syncthreads();
  u[j] = fmaxf(acc, 0.0f);
  __syncthreads();
  acc = __ldg(a.b2.p[model] + j);
  w = a.w2.p[model];
  for (int i = 0; i < 64; ++i) acc = fmaf(u[i], __ldg(w + i * 64 + j), acc);
  a.z[((int64_t)model * a.B + b) * 64 + j] = fmaxf(acc, 0.0f);
}

}  // namespace

// FP32 shared-memory tiled GEMM for one model: C[M][N] = A[M][K] W[K][N] + bias (+ R), used by
// the training step (train.cu) for the pointwise forward and input-gradient products.
int simt_pw_gemm(const float* A, const float* W_kn, const float* bias, const float* R, float* C,
                 int M, int K, int N, int relu6, cudaStream_t stream) {
  if (M <= 0) return 0;
  PwArgs a;
  for (int e = 0; e < kMaxModels; ++e) { a.w.p[e] = nullptr; a.bias.p[e] = nullptr; }
  a.w.p[0] = W_kn; a.bias.p[0] = bias;
  a.A = A; a.C = C; a.R = R; a.a_stride = 0; a.c_stride = 0;
  a.M = M; a.K = K; a.N = N; a.relu6 = relu6;
  return launch_pw(a, 1, stream, nullptr);
}

int launch_transform_visual(const float* lidar, int B, int C, int H, int W, float* visual,
                            cudaStream_t stream, bool hwc) {
  if (B <= 0) return 0;
  const int64_t total = (int64_t)B * C * 100 * 100;
  // ATen: scale = (in-1)/(out-1) evaluated in float for align_corners=True
  const float sh = (float)(H - 1) / 99.0f, sw = (float)(W - 1) / 99.0f;
  if (hwc)
    transform_visual_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        lidar, B * C, C, H, W, visual, sh, sw);
  else if (sh * (TV_T - 1) + 3.0f <= (float)TV_S && sw * (TV_T - 1) + 3.0f <= (float)TV_S &&
           (int64_t)B * C <= 65535)
    transform_visual_tiled_kernel<<<dim3(100 / TV_T, 100 / TV_T, (unsigned)(B * C)), 256, 0, stream>>>(
        lidar, H, W, visual, sh, sw);
  else
    transform_visual_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        lidar, B * C, C, H, W, visual, sh, sw);
  OAT_LAUNCHED("transform");
  return 0;
}

#ifndef OAT_ENC_GROUP_DEFAULT
#define OAT_ENC_GROUP_DEFAULT 0
#endif
static int encoder_forward_group(OatEnsemble* ens, const float* visual, const float* scalars, int B,
                                 float* z, cudaStream_t stream, int stop_after_blocks, float* prefix_out,
                                 int prefix_e0);

// The ensemble can be walked in groups of `OAT_ENC_GROUP` models (default: all at once), each group
// on its OWN stream (fork / join with events, capturable into a CUDA graph): every kernel of the
// network then covers fewer models, so (1) the activations between consecutive kernels shrink
// towards the L2 and (2) the persistent kernels of one group fill the SMs the other group's
// kernel leaves idle in its last wave (e.g. the 128 tiles of a 960->160 project on 148 SMs).
struct GroupStreams {
  cudaStream_t s[kMaxModels] = {nullptr};
  cudaEvent_t fork = nullptr, join[kMaxModels] = {nullptr};
};
static GroupStreams* group_streams(int device) {
  static GroupStreams per_device[64];
  static bool made[64] = {false};
  if (device < 0 || device >= 64) return nullptr;
  GroupStreams* g = &per_device[device];
  if (!made[device]) {
    if (cudaEventCreateWithFlags(&g->fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    for (int i = 1; i < kMaxModels; ++i) {
      if (cudaStreamCreateWithFlags(&g->s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&g->join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    made[device] = true;
  }
  return g;
}

int encoder_forward(OatEnsemble* ens, const float* visual, const float* scalars, int B,
                    float* z, cudaStream_t stream, int stop_after_blocks, float* prefix_out) {
  const int E = (int)ens->models.size();
  static const int group_env = []() { const char* e = getenv("OAT_ENC_GROUP"); return e ? atoi(e) : OAT_ENC_GROUP_DEFAULT; }();
  static const int serial_env = []() { const char* e = getenv("OAT_ENC_GROUP_SERIAL"); return e ? atoi(e) : 0; }();
  const int G = (group_env > 0 && group_env < E) ? group_env : E;
  if (G >= E) return encoder_forward_group(ens, visual, scalars, B, z, stream, stop_after_blocks, prefix_out, 0);
  // concurrent groups need disjoint workspace slices; the profile's event chain needs one stream
  GroupStreams* gs = (serial_env || g_profile_on) ? nullptr : group_streams(ens->device);
  if (gs) OAT_CUDA(cudaEventRecord(gs->fork, stream));
  const size_t kA = 50 * 50 * 32, kH1 = 50 * 50 * 96, kH2 = 25 * 25 * 144;  // per (model, image), see oat_ensemble_reserve
  int gi = 0;
  for (int e0 = 0; e0 < E; e0 += G, ++gi) {
    const int n = E - e0 < G ? E - e0 : G;
    OatEnsemble sub;  // a view: the group's models, tensor-core weights and workspace slice
    sub.models.assign(ens->models.begin() + e0, ens->models.begin() + e0 + n);
    sub.tc = ens->tc;
    for (TcLayer& t : sub.tc) {
      const size_t kn = (size_t)t.K * t.N;
      t.wh += kn * e0; t.wl += kn * e0; t.wr += kn * e0; t.bias += (size_t)t.N * e0;
    }
    sub.pw_impl = ens->pw_impl; sub.fuse = ens->fuse; sub.fuse_tc = ens->fuse_tc; sub.device = ens->device;
    sub.reserved_batch = ens->reserved_batch;
    const size_t eb = (size_t)e0 * ens->reserved_batch;
    sub.bufA = ens->bufA + eb * kA; sub.bufB = ens->bufB + eb * kA;
    sub.bufH1 = ens->bufH1 + eb * kH1; sub.bufH2 = ens->bufH2 + eb * kH2;
    sub.pooled = ens->pooled + eb * 1280; sub.feat = ens->feat + eb * 128;
    cudaStream_t gstream = (gs && gi > 0) ? gs->s[gi] : stream;
    if (gs && gi > 0) OAT_CUDA(cudaStreamWaitEvent(gstream, gs->fork, 0));
    const int rc = encoder_forward_group(&sub, visual, scalars, B, z ? z + (size_t)e0 * B * kHidden : nullptr,
                                         gstream, stop_after_blocks, prefix_out, e0);
    if (rc) return rc;
    if (gs && gi > 0) {
      OAT_CUDA(cudaEventRecord(gs->join[gi], gstream));
      OAT_CUDA(cudaStreamWaitEvent(stream, gs->join[gi], 0));
    }
  }
  return 0;
}

static int encoder_forward_group(OatEnsemble* ens, const float* visual, const float* scalars, int B,
                                 float* z, cudaStream_t stream, int stop_after_blocks, float* prefix_out,
                                 int prefix_e0) {
  const int E = (int)ens->models.size();
  const OatModel* m0 = ens->models[0];
  const int C = m0->in_channels;
  if (C > STEM_MAXC) return fail("encoder: in_channels > 8 unsupported");
  auto table = [&](auto getter) {
    PtrTable t;
    for (int e = 0; e < kMaxModels; ++e) t.p[e] = nullptr;
    for (int e = 0; e < E; ++e) t.p[e] = getter(ens->models[e]);
    return t;
  };

  float* x = ens->bufA;
  float* y = ens->bufB;
  size_t li = 0;  // index into ens->tc (pointwise layers in execution order)
  size_t first_block = 0;
  if (ens->fuse & 1) {
    // features.0 + features.1 in one kernel (fused.cu): visual -> bufA [E][B][50][50][16]
    FusedFrontLaunch f;
    f.ws = table([](const OatModel* m) { return m->stem.w; });
    f.bs = table([](const OatModel* m) { return m->stem.b; });
    f.wd = table([](const OatModel* m) { return m->blocks[0].dw.w; });
    f.bd = table([](const OatModel* m) { return m->blocks[0].dw.b; });
    f.wp = table([](const OatModel* m) { return m->blocks[0].project.w; });
    f.bp = table([](const OatModel* m) { return m->blocks[0].project.b; });
    f.visual = visual; f.out = x; f.E = E; f.B = B; f.C = C;
    if (int rc = launch_fused_front(f, stream)) return rc;
    first_block = 1;
    li = 1;  // the project layer of block 1
  } else {
    // stem -> bufA [E][B][50][50][32]
    PtrTable w = table([](const OatModel* m) { return m->stem.w; });
    PtrTable b = table([](const OatModel* m) { return m->stem.b; });
#if !defined(OAT_STEM_1PX)
    dim3 grid((unsigned)(((int64_t)B * 1250 + STEM_THREADS - 1) / STEM_THREADS), E);
#else
    dim3 grid((unsigned)(((int64_t)B * 2500 + STEM_THREADS - 1) / STEM_THREADS), E);
#endif
    stem_kernel<<<grid, STEM_THREADS, 0, stream>>>(w, b, visual, B, C, ens->bufA);
    OAT_LAUNCHED("stem");
  }
  auto prefix_done = [&](size_t blocks_done, const float* act, int h, int c) -> int {
    if (stop_after_blocks < 0 || (size_t)stop_after_blocks != blocks_done) return 0;
    OAT_CUDA(cudaMemcpyAsync(prefix_out + (size_t)prefix_e0 * B * h * h * c, act,
                             (size_t)E * B * h * h * c * sizeof(float),
                             cudaMemcpyDeviceToDevice, stream));
    return -1;
  };
  if (first_block == 1) {
    if (int rc = prefix_done(1, x, 50, 16)) return rc < 0 ? 0 : rc;
  } else {
    if (int rc = prefix_done(0, x, 50, 32)) return rc < 0 ? 0 : rc;
  }
  auto tc = [&]() -> const TcLayer* {
    const TcLayer* t = ens->pw_impl == 1 ? &ens->tc[li] : nullptr;
    ++li;
    return t;
  };
  for (size_t bi = first_block; bi < m0->blocks.size(); ++bi) {
    const BlockW& blk = m0->blocks[bi];
    const int Min = B * blk.hin * blk.hin, Mout = B * blk.hout * blk.hout;
    const float* dw_in = x;
    if (bi == 0 && (ens->fuse & 16) && blk.hid == 32 && blk.cout == 16 && blk.hin == 50 &&
        blk.stride == 1) {
      // features.1 depthwise + project in one kernel: the 50x50x32 depthwise output never leaves the SM
      FusedDwProjectLaunch f;
      f.wd = table([](const OatModel* m) { return m->blocks[0].dw.w; });
      f.bd = table([](const OatModel* m) { return m->blocks[0].dw.b; });
      f.wp = table([](const OatModel* m) { return m->blocks[0].project.w; });
      f.bp = table([](const OatModel* m) { return m->blocks[0].project.b; });
      f.in = x; f.out = y; f.E = E; f.B = B;
      if (int rc = launch_fused_dw_project(f, stream)) return rc;
      ++li;  // the project layer's tensor-core copy is not used
      float* t = x; x = y; y = t;
      if (int rc = prefix_done(1, x, blk.hout, blk.cout)) return rc < 0 ? 0 : rc;
      continue;
    }
    bool dw_fused = false;
    const bool fuse_block = bi >= 1 && bi <= 3 && ((ens->fuse >> bi) & 1) &&
                            fused_block_supported(blk.cin, blk.hid, blk.stride, blk.hin);
    if (fuse_block) {  // expand + depthwise in one kernel: the 6x tensor stays in shared memory
      FusedBlockLaunch f;
      f.we = table([bi](const OatModel* m) { return m->blocks[bi].expand.w; });
      f.be = table([bi](const OatModel* m) { return m->blocks[bi].expand.b; });
      f.wd = table([bi](const OatModel* m) { return m->blocks[bi].dw.w; });
      f.bd = table([bi](const OatModel* m) { return m->blocks[bi].dw.b; });
      f.in = x; f.out = ens->bufH2; f.E = E; f.B = B;
      f.cin = blk.cin; f.hid = blk.hid; f.stride = blk.stride; f.hin = blk.hin;
      f.tensor_cores = ens->pw_impl == 1 ? ens->fuse_tc : 0;
      if (int rc = launch_fused_expand_dw(f, stream)) return rc;
      ++li;  // the expand layer's tensor-core copy is not used
    } else if (blk.hid != blk.cin && (ens->fuse & 32) && ens->pw_impl == 1 &&
               tc_dw_epilogue_supported(blk.hin, blk.stride, blk.hid)) {
      // expand 1x1 + depthwise 3x3 in ONE tensor-core kernel (features.5-17): the M tiles hold whole
      // images, so the depthwise window slides over the slab the GEMM epilogue has just staged in
      // shared memory (tc_gemm.cu); the 6x expanded tensor is never written
      const TcLayer* t = &ens->tc[li];
      ++li;
      TcGemmProblem p;
      p.A = x; p.Wh = t->wh; p.Wl = t->wl; p.Wr = t->wr; p.bias = t->bias; p.R = nullptr; p.C = nullptr;
      p.M = Min; p.K = blk.cin; p.N = blk.hid; p.E = E; p.relu6 = 1;
      p.dw_out = ens->bufH2; p.B = B; p.hin = blk.hin; p.hout = blk.hout; p.stride = blk.stride;
      for (int e = 0; e < E; ++e) {
        p.dw_w[e] = ens->models[e]->blocks[bi].dw.w;
        p.dw_b[e] = ens->models[e]->blocks[bi].dw.b;
      }
      if (int rc = tc_pw_gemm(p, stream)) return rc;
      dw_fused = true;
    } else if (blk.hid != blk.cin) {  // expand 1x1 + BN + ReLU6
      PwArgs a;
      a.w = table([bi](const OatModel* m) { return m->blocks[bi].expand.w; });
      a.bias = table([bi](const OatModel* m) { return m->blocks[bi].expand.b; });
      a.A = x; a.C = ens->bufH1; a.R = nullptr;
      a.a_stride = (int64_t)Min * blk.cin; a.c_stride = (int64_t)Min * blk.hid;
      a.M = Min; a.K = blk.cin; a.N = blk.hid; a.relu6 = 1;
      if (int rc = launch_pw(a, E, stream, tc())) return rc;
      dw_in = ens->bufH1;
    }
    if (!fuse_block && !dw_fused) {  // depthwise 3x3 + BN + ReLU6
      PtrTable w = table([bi](const OatModel* m) { return m->blocks[bi].dw.w; });
      PtrTable b = table([bi](const OatModel* m) { return m->blocks[bi].dw.b; });
      const int64_t total = (int64_t)B * blk.hout * (blk.hid / 4);
      dim3 grid((unsigned)((total + OAT_DW_THREADS - 1) / OAT_DW_THREADS), E);
#if !defined(OAT_DW_1ROW)
      if (blk.stride == 1) {
        const int64_t tot2 = (int64_t)B * ((blk.hout + 1) / 2) * (blk.hid / 4);
        dim3 grid2((unsigned)((tot2 + OAT_DW_THREADS - 1) / OAT_DW_THREADS), E);
        dw2_kernel<<<grid2, OAT_DW_THREADS, 0, stream>>>(w, b, dw_in, ens->bufH2, B, blk.hin, blk.hid);
      } else
#else
      if (blk.stride == 1)
        dw_kernel<1><<<grid, OAT_DW_THREADS, 0, stream>>>(w, b, dw_in, ens->bufH2, B, blk.hin, blk.hout, blk.hid);
      else
#endif
        dw_kernel<2><<<grid, OAT_DW_THREADS, 0, stream>>>(w, b, dw_in, ens->bufH2, B, blk.hin, blk.hout, blk.hid);
      OAT_LAUNCHED("depthwise");
    }
    {  // project 1x1 + BN (linear) + residual
      PwArgs a;
      a.w = table([bi](const OatModel* m) { return m->blocks[bi].project.w; });
      a.bias = table([bi](const OatModel* m) { return m->blocks[bi].project.b; });
      a.A = ens->bufH2; a.C = y; a.R = blk.residual ? x : nullptr;
      a.a_stride = (int64_t)Mout * blk.hid; a.c_stride = (int64_t)Mout * blk.cout;
      a.M = Mout; a.K = blk.hid; a.N = blk.cout; a.relu6 = 0;
      if (int rc = launch_pw(a, E, stream, tc())) return rc;
    }
    float* t = x; x = y; y = t;
    if (int rc = prefix_done(bi + 1, x, blk.hout, blk.cout)) return rc < 0 ? 0 : rc;
  }
  const int hl = m0->blocks.back().hout;  // 4
  const int P = hl * hl;
  {  // features.18: 1x1 320->1280 + BN + ReLU6 -> bufH1 [E][B*P][1280]
    PwArgs a;
    a.w = table([](const OatModel* m) { return m->last.w; });
    a.bias = table([](const OatModel* m) { return m->last.b; });
    a.A = x; a.C = ens->bufH1; a.R = nullptr;
    a.a_stride = (int64_t)B * P * 320; a.c_stride = (int64_t)B * P * 1280;
    a.M = B * P; a.K = 320; a.N = 1280; a.relu6 = 1;
    if (int rc = launch_pw(a, E, stream, tc())) return rc;
  }
  {  // global average pool -> pooled [E][B][1280]
    const int64_t rows = (int64_t)E * B;
    pool_kernel<<<(unsigned)((rows * 1280 + 255) / 256), 256, 0, stream>>>(ens->bufH1, ens->pooled,
                                                                          rows, P, 1280);
    OAT_LAUNCHED("pool");
  }
  {  // classifier.1: Linear 1280 -> 128 (Dropout is identity in eval)
    PwArgs a;
    a.w = table([](const OatModel* m) { return m->fc.w; });
    a.bias = table([](const OatModel* m) { return m->fc.b; });
    a.A = ens->pooled; a.C = ens->feat; a.R = nullptr;
    a.a_stride = (int64_t)B * 1280; a.c_stride = (int64_t)B * 128;
    a.M = B; a.K = 1280; a.N = 128; a.relu6 = 0;
    if (int rc = launch_pw(a, E, stream, tc())) return rc;
  }
  if (stop_after_blocks == 18) {  // `MobileNetV2.forward` alone: the 128 features, no merger
    OAT_CUDA(cudaMemcpyAsync(prefix_out + (size_t)prefix_e0 * B * 128, ens->feat,
                             (size_t)E * B * 128 * sizeof(float),
                             cudaMemcpyDeviceToDevice, stream));
    return 0;
  }
  if (m0->kind == OAT_KIND_ENCODER) return fail("encoder: a MobileNetV2-only model has no merger");
  {  // merger MLP -> z [E][B][64]
    MergerArgs a;
    a.w0 = table([](const OatModel* m) { return m->merger[0].w; });
    a.b0 = table([](const OatModel* m) { return m->merger[0].b; });
    a.w1 = table([](const OatModel* m) { return m->merger[1].w; });
    a.b1 = table([](const OatModel* m) { return m->merger[1].b; });
    a.w2 = table([](const OatModel* m) { return m->merger[2].w; });
    a.b2 = table([](const OatModel* m) { return m->merger[2].b; });
    a.feat = ens->feat; a.scalars = scalars; a.B = B; a.S = m0->scalars; a.z = z;
    merger_kernel<<<dim3(B, E), 64, 0, stream>>>(a);
    OAT_LAUNCHED("merger");
  }
  return 0;
}

}  // namespace oat
