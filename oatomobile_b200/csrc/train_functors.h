// Work-item bodies of the training step (dim/train.py:175-213,
// cil/train.py:168-190): training-mode MobileNetV2 forward (BatchNorm batch statistics,
// torchvision `mobilenetv2.py`), its hand-derived backward, the merger MLP, the
// flow NLL / CIL L1 decoders with analytic back-propagation through time, and Adam.
//
// Every body is a functor over a flat work-item index `gid`; there is no shared
// memory and no intra-block synchronisation, reductions go through atomics.  The
// CUDA product wraps each functor in a grid-stride kernel (train.cu).  The same
// header also compiles as plain C++ so the unit tests can execute the bodies on
// the host (tests/emu) — that build is test tooling and is never loaded by the package.
//
// Layouts: activations NHWC `[B][H][W][C]` (rows m = (b,y,x)); parameters exactly as
// the reference `state_dict` stores them (conv `[Cout][Cin][kh][kw]`, Linear
// `[out][in]`, GRUCell `[3H][in]`), because they change every step.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OAT_HD __host__ __device__ __forceinline__
#else
#define OAT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define OAT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define OAT_ATOMIC_ADD(p, v) (*(p) += (v))
#endif
#if defined(__CUDACC__)
#define OAT_UNROLL _Pragma("unroll")
#else
#define OAT_UNROLL
#endif

namespace oat {
namespace train {

struct alignas(16) F4 {
  float x, y, z, w;
};
OAT_HD F4 ld4(const float* p) { return *reinterpret_cast<const F4*>(p); }
OAT_HD void st4(float* p, F4 v) { *reinterpret_cast<F4*>(p) = v; }
OAT_HD float& at(F4& v, int i) { return (&v.x)[i]; }
OAT_HD float at(const F4& v, int i) { return (&v.x)[i]; }
OAT_HD float relu6f(float v) { return fminf(fmaxf(v, 0.0f), 6.0f); }
OAT_HD float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// Reductions over rows are split into chunks of `rows` rows per work item (chosen by the
// launcher so that every reduction exposes ~150k work items) and combined with atomics.
// Same-address atomics from many SMs serialise (~45 ns each, measured), so every such
// accumulator exists in kReplicas copies (chunk c adds into copy c % kReplicas) that the
// per-channel / per-element finalising work item sums and clears.
constexpr int kReplicas = 32;

// ---------------------------------------------------------------------------------
// Stem: 3x3 stride-2 pad-1 convolution, NCHW image -> NHWC rows (perception.py:43-51)
// ---------------------------------------------------------------------------------
struct StemFwd {  // gid over B*Ho*Wo*32
  const float* x;  // [B][C][H][W]
  const float* w;  // [32][C][3][3]
  float* r;        // [B][Ho][Wo][32]
  int B, C, H, W, Ho, Wo;
  OAT_HD void operator()(int64_t gid) const {
    const int co = (int)(gid & 31);
    int64_t m = gid >> 5;
    const int ox = (int)(m % Wo);
    const int oy = (int)((m / Wo) % Ho);
    const int b = (int)(m / ((int64_t)Wo * Ho));
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) {
      const float* xp = x + ((int64_t)b * C + c) * H * W;
      const float* wp = w + (co * C + c) * 9;
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * 2 - 1 + kx;
          if (ix < 0 || ix >= W) continue;
          acc = fmaf(xp[iy * W + ix], wp[ky * 3 + kx], acc);
        }
      }
    }
    r[gid] = acc;
  }
};

struct StemBwdW {  // gid over chunks*32; chunk = `rows` output pixels; one output channel each
  const float* x;
  const float* g;  // dR [B*Ho*Wo][32]
  float* gw;       // replicas [kReplicas][32*C*9], pre-zeroed (summed by ReduceReplicas)
  int B, C, H, W, Ho, Wo, rows;
  OAT_HD void operator()(int64_t gid) const {
    const int co = (int)(gid & 31);
    const int64_t chunk = gid >> 5;
    const int64_t M = (int64_t)B * Ho * Wo;
    const int64_t m0 = chunk * rows;
    const int64_t m1 = m0 + rows < M ? m0 + rows : M;
    float* dst = gw + (chunk % kReplicas) * (int64_t)(32 * C * 9) + co * C * 9;
    for (int c = 0; c < C; ++c) {
      float acc[9] = {};
      for (int64_t m = m0; m < m1; ++m) {
        const int ox = (int)(m % Wo);
        const int oy = (int)((m / Wo) % Ho);
        const int b = (int)(m / ((int64_t)Wo * Ho));
        const float gv = g[m * 32 + co];
        const float* xp = x + ((int64_t)b * C + c) * H * W;
        OAT_UNROLL
        for (int ky = 0; ky < 3; ++ky) {
          const int iy = oy * 2 - 1 + ky;
          OAT_UNROLL
          for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            const bool in = iy >= 0 && iy < H && ix >= 0 && ix < W;
            acc[ky * 3 + kx] = fmaf(gv, in ? xp[iy * W + ix] : 0.0f, acc[ky * 3 + kx]);
          }
        }
      }
      for (int t = 0; t < 9; ++t) OAT_ATOMIC_ADD(dst + c * 9 + t, acc[t]);
    }
  }
};

// ---------------------------------------------------------------------------------
// Pointwise (1x1) convolutions as GEMMs over rows; 4x4 register tiles.
// Requires N % 4 == 0 and K % 4 == 0 (true for every MobileNetV2 layer).
// ---------------------------------------------------------------------------------
struct PwFwd {  // R[m][n] = sum_k A[m][k] W[n][k];  gid over ceil(M/4)*(N/4), n-tile fastest
  const float* a;
  const float* w;
  float* r;
  int64_t M;
  int N, K;
  OAT_HD void operator()(int64_t gid) const {
    const int nt = N >> 2;
    const int n0 = (int)(gid % nt) << 2;
    const int64_t m0 = (gid / nt) << 2;
    float acc[4][4] = {};
    const float* ap[4];
    for (int i = 0; i < 4; ++i) ap[i] = a + (m0 + i < M ? m0 + i : M - 1) * K;
    for (int k = 0; k < K; k += 4) {
      F4 av[4], wv[4];
      for (int i = 0; i < 4; ++i) av[i] = ld4(ap[i] + k);
      for (int j = 0; j < 4; ++j) wv[j] = ld4(w + (int64_t)(n0 + j) * K + k);
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(av[i].x, wv[j].x, acc[i][j]);
          acc[i][j] = fmaf(av[i].y, wv[j].y, acc[i][j]);
          acc[i][j] = fmaf(av[i].z, wv[j].z, acc[i][j]);
          acc[i][j] = fmaf(av[i].w, wv[j].w, acc[i][j]);
        }
    }
    for (int i = 0; i < 4; ++i)
      if (m0 + i < M) st4(r + (m0 + i) * N + n0, F4{acc[i][0], acc[i][1], acc[i][2], acc[i][3]});
  }
};

struct PwBwdX {  // dA[m][k] (+)= sum_n G[m][n] W[n][k];  gid over ceil(M/4)*(K/4), k-tile fastest
  const float* g;
  const float* w;
  float* da;
  int64_t M;
  int N, K;
  int accumulate;
  OAT_HD void operator()(int64_t gid) const {
    const int kt = K >> 2;
    const int k0 = (int)(gid % kt) << 2;
    const int64_t m0 = (gid / kt) << 2;
    float acc[4][4] = {};
    const float* gp[4];
    for (int i = 0; i < 4; ++i) gp[i] = g + (m0 + i < M ? m0 + i : M - 1) * N;
    for (int n = 0; n < N; n += 4) {
      F4 gv[4], wv[4];
      for (int i = 0; i < 4; ++i) gv[i] = ld4(gp[i] + n);
      for (int j = 0; j < 4; ++j) wv[j] = ld4(w + (int64_t)(n + j) * K + k0);
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          const float gij = at(gv[i], j);
          acc[i][0] = fmaf(gij, wv[j].x, acc[i][0]);
          acc[i][1] = fmaf(gij, wv[j].y, acc[i][1]);
          acc[i][2] = fmaf(gij, wv[j].z, acc[i][2]);
          acc[i][3] = fmaf(gij, wv[j].w, acc[i][3]);
        }
    }
    for (int i = 0; i < 4; ++i) {
      if (m0 + i >= M) continue;
      float* p = da + (m0 + i) * K + k0;
      F4 o{acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
      if (accumulate) {
        const F4 old = ld4(p);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      st4(p, o);
    }
  }
};

struct PwBwdW {  // gW[n][k] += sum_{m in chunk} G[m][n] A[m][k];  gid over chunks*(N/4)*(K/4)
  const float* g;
  const float* a;
  float* gw;  // pre-zeroed
  int64_t M;
  int N, K, rows;
  OAT_HD void operator()(int64_t gid) const {
    const int kt = K >> 2, nt = N >> 2;
    const int k0 = (int)(gid % kt) << 2;
    const int n0 = (int)((gid / kt) % nt) << 2;
    const int64_t chunk = gid / ((int64_t)kt * nt);
    const int64_t m0 = chunk * rows;
    const int64_t m1 = m0 + rows < M ? m0 + rows : M;
    float acc[4][4] = {};
    for (int64_t m = m0; m < m1; ++m) {
      const F4 gv = ld4(g + m * N + n0);
      const F4 av = ld4(a + m * K + k0);
      for (int i = 0; i < 4; ++i) {
        const float gi = at(gv, i);
        acc[i][0] = fmaf(gi, av.x, acc[i][0]);
        acc[i][1] = fmaf(gi, av.y, acc[i][1]);
        acc[i][2] = fmaf(gi, av.z, acc[i][2]);
        acc[i][3] = fmaf(gi, av.w, acc[i][3]);
      }
    }
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) OAT_ATOMIC_ADD(gw + (int64_t)(n0 + i) * K + k0 + j, acc[i][j]);
  }
};

// ---------------------------------------------------------------------------------
// Depthwise 3x3, pad 1, stride 1 or 2; weights [C][1][3][3]
// ---------------------------------------------------------------------------------
struct DwFwd {  // gid over B*Ho*Wo*(C/4)
  const float* a;  // [B][H][W][C]
  const float* w;
  float* r;  // [B][Ho][Wo][C]
  int B, H, W, Ho, Wo, C, stride;
  OAT_HD void operator()(int64_t gid) const {
    const int ct = C >> 2;
    const int c0 = (int)(gid % ct) << 2;
    int64_t m = gid / ct;
    const int ox = (int)(m % Wo);
    const int oy = (int)((m / Wo) % Ho);
    const int b = (int)(m / ((int64_t)Wo * Ho));
    F4 acc{0, 0, 0, 0};
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * stride - 1 + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * stride - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        const F4 v = ld4(a + (((int64_t)b * H + iy) * W + ix) * C + c0);
        const int t = ky * 3 + kx;
        acc.x = fmaf(v.x, w[(c0 + 0) * 9 + t], acc.x);
        acc.y = fmaf(v.y, w[(c0 + 1) * 9 + t], acc.y);
        acc.z = fmaf(v.z, w[(c0 + 2) * 9 + t], acc.z);
        acc.w = fmaf(v.w, w[(c0 + 3) * 9 + t], acc.w);
      }
    }
    st4(r + m * C + c0, acc);
  }
};

struct DwBwdX {  // gid over B*H*W*(C/4); dA written (the expanded tensor has one consumer)
  const float* g;  // dR [B][Ho][Wo][C]
  const float* w;
  float* da;  // [B][H][W][C]
  int B, H, W, Ho, Wo, C, stride;
  OAT_HD void operator()(int64_t gid) const {
    const int ct = C >> 2;
    const int c0 = (int)(gid % ct) << 2;
    int64_t m = gid / ct;
    const int ix = (int)(m % W);
    const int iy = (int)((m / W) % H);
    const int b = (int)(m / ((int64_t)W * H));
    F4 acc{0, 0, 0, 0};
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = iy + 1 - ky;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = ix + 1 - kx;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        const F4 v = ld4(g + (((int64_t)b * Ho + oy) * Wo + ox) * C + c0);
        const int t = ky * 3 + kx;
        acc.x = fmaf(v.x, w[(c0 + 0) * 9 + t], acc.x);
        acc.y = fmaf(v.y, w[(c0 + 1) * 9 + t], acc.y);
        acc.z = fmaf(v.z, w[(c0 + 2) * 9 + t], acc.z);
        acc.w = fmaf(v.w, w[(c0 + 3) * 9 + t], acc.w);
      }
    }
    st4(da + m * C + c0, acc);
  }
};

struct DwBwdW {  // gid over chunks*(C/4); chunk = `rows` output pixels
  const float* g;
  const float* a;
  float* gw;  // replicas [kReplicas][C*9], pre-zeroed (summed by ReduceReplicas)
  int B, H, W, Ho, Wo, C, stride, rows;
  OAT_HD void operator()(int64_t gid) const {
    const int ct = C >> 2;
    const int c0 = (int)(gid % ct) << 2;
    const int64_t chunk = gid / ct;
    const int64_t M = (int64_t)B * Ho * Wo;
    const int64_t m0 = chunk * rows;
    const int64_t m1 = m0 + rows < M ? m0 + rows : M;
    float acc[4][9] = {};
    for (int64_t m = m0; m < m1; ++m) {
      const int ox = (int)(m % Wo);
      const int oy = (int)((m / Wo) % Ho);
      const int b = (int)(m / ((int64_t)Wo * Ho));
      const F4 gv = ld4(g + m * C + c0);
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * stride - 1 + ky;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * stride - 1 + kx;
          if (ix < 0 || ix >= W) continue;
          const F4 v = ld4(a + (((int64_t)b * H + iy) * W + ix) * C + c0);
          const int t = ky * 3 + kx;
          acc[0][t] = fmaf(gv.x, v.x, acc[0][t]);
          acc[1][t] = fmaf(gv.y, v.y, acc[1][t]);
          acc[2][t] = fmaf(gv.z, v.z, acc[2][t]);
          acc[3][t] = fmaf(gv.w, v.w, acc[3][t]);
        }
      }
    }
    float* dst = gw + (chunk % kReplicas) * (int64_t)(C * 9);
    for (int i = 0; i < 4; ++i)
      for (int t = 0; t < 9; ++t) OAT_ATOMIC_ADD(dst + (c0 + i) * 9 + t, acc[i][t]);
  }
};

struct ReduceReplicas {  // gid over n: out[i] = sum_r rep[r][i]; clears the replicas
  float* rep;
  float* out;
  int64_t n;
  OAT_HD void operator()(int64_t i) const {
    float v[kReplicas];
    OAT_UNROLL
    for (int r = 0; r < kReplicas; ++r) v[r] = rep[r * n + i];
    float s = 0.0f;
    OAT_UNROLL
    for (int r = 0; r < kReplicas; ++r) {
      s += v[r];
      rep[r * n + i] = 0.0f;
    }
    out[i] = s;
  }
};

// ---------------------------------------------------------------------------------
// BatchNorm2d in training mode (eps 1e-5, momentum 0.1, biased variance for the
// normalisation, unbiased for the running estimate — torch.nn.BatchNorm2d)
// ---------------------------------------------------------------------------------
struct BnStats {  // gid over chunks*(C/4): acc[c] += sum R, acc[C+c] += sum R^2 (double)
  const float* r;
  double* acc;  // [kReplicas][2C], pre-zeroed
  int64_t M;
  int C, rows;
  OAT_HD void operator()(int64_t gid) const {
    const int ct = C >> 2;
    const int c0 = (int)(gid % ct) << 2;
    const int64_t m0 = (gid / ct) * rows;
    const int64_t m1 = m0 + rows < M ? m0 + rows : M;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    for (int64_t m = m0; m < m1; ++m) {
      const F4 v = ld4(r + m * C + c0);
      for (int i = 0; i < 4; ++i) {
        const double x = (double)at(v, i);
        s1[i] += x;
        s2[i] += x * x;
      }
    }
    double* dst = acc + ((gid / ct) % kReplicas) * 2 * C;
    for (int i = 0; i < 4; ++i) {
      OAT_ATOMIC_ADD(dst + c0 + i, s1[i]);
      OAT_ATOMIC_ADD(dst + C + c0 + i, s2[i]);
    }
  }
};

OAT_HD double drain_replicas(double* acc, int64_t index, int64_t stride) {
  double v[kReplicas];
  OAT_UNROLL
  for (int r = 0; r < kReplicas; ++r) v[r] = acc[r * stride + index];  // independent loads first
  double s = 0.0;
  OAT_UNROLL
  for (int r = 0; r < kReplicas; ++r) {
    s += v[r];
    acc[r * stride + index] = 0.0;
  }
  return s;
}

struct BnFinalize {  // gid over C: batch mean / 1/sqrt(var+eps), running estimates, clears acc
  double* acc;
  float *mean, *invstd, *running_mean, *running_var;
  int64_t M;
  int C;
  OAT_HD void operator()(int64_t c) const {
    const double mu = drain_replicas(acc, c, 2 * C) / (double)M;
    double var = drain_replicas(acc, C + c, 2 * C) / (double)M - mu * mu;  // sums are exact in double
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    invstd[c] = (float)(1.0 / sqrt(var + 1e-5));
    const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
    running_mean[c] = (float)(0.9 * (double)running_mean[c] + 0.1 * mu);
    running_var[c] = (float)(0.9 * (double)running_var[c] + 0.1 * unbiased);
  }
};

struct BnApply {  // gid over M*(C/4): A = act(gamma*(R-mean)*invstd + beta) (+ skip)
  const float* r;
  const float *mean, *invstd, *gamma, *beta;
  const float* skip;  // residual input or null
  float* a;
  int C, relu6;
  OAT_HD void operator()(int64_t gid) const {
    const int ct = C >> 2;
    const int c0 = (int)(gid % ct) << 2;
    const int64_t off = (gid / ct) * C + c0;
    const F4 v = ld4(r + off), mu = ld4(mean + c0), is = ld4(invstd + c0), ga = ld4(gamma + c0),
             be = ld4(beta + c0);
    F4 o;
    for (int i = 0; i < 4; ++i) {
      float p = fmaf((at(v, i) - at(mu, i)) * at(is, i), at(ga, i), at(be, i));
      if (relu6) p = relu6f(p);
      at(o, i) = p;
    }
    if (skip) {
      const F4 s = ld4(skip + off);
      o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
    }
    st4(a + off, o);
  }
};

struct BnBwdReduce {  // gid over chunks*(C/4): acc[c] += sum dP, acc[C+c] += sum dP*xhat
  const float* r;
  const float* g;  // dA
  const float *mean, *invstd, *gamma, *beta;
  double* acc;  // [kReplicas][2C], pre-zeroed
  int64_t M;
  int C, relu6, rows;
  OAT_HD void operator()(int64_t gid) const {
    const int ct = C >> 2;
    const int c0 = (int)(gid % ct) << 2;
    const int64_t m0 = (gid / ct) * rows;
    const int64_t m1 = m0 + rows < M ? m0 + rows : M;
    const F4 mu = ld4(mean + c0), is = ld4(invstd + c0), ga = ld4(gamma + c0), be = ld4(beta + c0);
    F4 s1{0, 0, 0, 0}, s2{0, 0, 0, 0};
    for (int64_t m = m0; m < m1; ++m) {
      const F4 v = ld4(r + m * C + c0), d = ld4(g + m * C + c0);
      for (int i = 0; i < 4; ++i) {
        const float xh = (at(v, i) - at(mu, i)) * at(is, i);
        float dp = at(d, i);
        if (relu6) {
          const float p = fmaf(xh, at(ga, i), at(be, i));
          if (!(p > 0.0f && p < 6.0f)) dp = 0.0f;
        }
        at(s1, i) += dp;
        at(s2, i) = fmaf(dp, xh, at(s2, i));
      }
    }
    double* dst = acc + ((gid / ct) % kReplicas) * 2 * C;
    for (int i = 0; i < 4; ++i) {
      OAT_ATOMIC_ADD(dst + c0 + i, (double)at(s1, i));
      OAT_ATOMIC_ADD(dst + C + c0 + i, (double)at(s2, i));
    }
  }
};

struct BnBwdParams {  // gid over C
  double* acc;
  float *ggamma, *gbeta;        // parameter gradients
  float *mean_dp, *mean_dpxh;   // per-channel means used by BnBwdDx
  int64_t M;
  int C;
  OAT_HD void operator()(int64_t c) const {
    const double s1 = drain_replicas(acc, c, 2 * C), s2 = drain_replicas(acc, C + c, 2 * C);
    gbeta[c] = (float)s1;
    ggamma[c] = (float)s2;
    mean_dp[c] = (float)(s1 / (double)M);
    mean_dpxh[c] = (float)(s2 / (double)M);
  }
};

struct BnBwdDx {  // gid over M*(C/4): G <- dR in place; optionally forwards dA to the skip branch
  const float* r;
  float* g;
  const float *mean, *invstd, *gamma, *beta, *mean_dp, *mean_dpxh;
  float* gskip;  // gradient buffer of the residual source (written) or null
  int C, relu6;
  OAT_HD void operator()(int64_t gid) const {
    const int ct = C >> 2;
    const int c0 = (int)(gid % ct) << 2;
    const int64_t off = (gid / ct) * C + c0;
    const F4 v = ld4(r + off), d = ld4(g + off);
    const F4 mu = ld4(mean + c0), is = ld4(invstd + c0), ga = ld4(gamma + c0), be = ld4(beta + c0),
             m1 = ld4(mean_dp + c0), m2 = ld4(mean_dpxh + c0);
    if (gskip) st4(gskip + off, d);
    F4 o;
    for (int i = 0; i < 4; ++i) {
      const float xh = (at(v, i) - at(mu, i)) * at(is, i);
      float dp = at(d, i);
      if (relu6) {
        const float p = fmaf(xh, at(ga, i), at(be, i));
        if (!(p > 0.0f && p < 6.0f)) dp = 0.0f;
      }
      at(o, i) = at(ga, i) * at(is, i) * (dp - at(m1, i) - xh * at(m2, i));
    }
    st4(g + off, o);
  }
};

// ---------------------------------------------------------------------------------
// Global average pool (+ the classifier's Dropout mask, pre-scaled by 1/(1-p))
// ---------------------------------------------------------------------------------
struct PoolFwd {  // gid over B*(C/4)
  const float* a;     // [B][HW][C]
  const float* mask;  // [B][C] or null
  float* pooled;      // [B][C]
  int HW, C;
  OAT_HD void operator()(int64_t gid) const {
    const int ct = C >> 2;
    const int c0 = (int)(gid % ct) << 2;
    const int64_t b = gid / ct;
    F4 s{0, 0, 0, 0};
    for (int i = 0; i < HW; ++i) {
      const F4 v = ld4(a + (b * HW + i) * C + c0);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float inv = 1.0f / (float)HW;
    F4 mk{1, 1, 1, 1};
    if (mask) mk = ld4(mask + b * C + c0);
    st4(pooled + b * C + c0, F4{s.x * inv * mk.x, s.y * inv * mk.y, s.z * inv * mk.z, s.w * inv * mk.w});
  }
};

struct PoolBwd {  // gid over B*HW*(C/4)
  const float* gpooled;
  const float* mask;
  float* g;
  int HW, C;
  OAT_HD void operator()(int64_t gid) const {
    const int ct = C >> 2;
    const int c0 = (int)(gid % ct) << 2;
    const int64_t m = gid / ct;
    const int64_t b = m / HW;
    F4 v = ld4(gpooled + b * C + c0);
    const float inv = 1.0f / (float)HW;
    F4 mk{1, 1, 1, 1};
    if (mask) mk = ld4(mask + b * C + c0);
    st4(g + m * C + c0, F4{v.x * inv * mk.x, v.y * inv * mk.y, v.z * inv * mk.z, v.w * inv * mk.w});
  }
};

// ---------------------------------------------------------------------------------
// Small dense layers (classifier 1280->128, merger 133/134->64->64->64); M = batch rows
// ---------------------------------------------------------------------------------
struct LinearFwd {  // gid over M*N
  const float* x;
  const float *w, *b;
  float* y;
  int N, K, ldx, ldy, relu;
  OAT_HD void operator()(int64_t gid) const {
    const int n = (int)(gid % N);
    const int64_t m = gid / N;
    const float* xp = x + m * ldx;
    const float* wp = w + (int64_t)n * K;
    float acc = b[n];
    for (int k = 0; k < K; ++k) acc = fmaf(xp[k], wp[k], acc);
    if (relu) acc = fmaxf(acc, 0.0f);
    y[m * ldy + n] = acc;
  }
};

struct LinearBwdX {  // gid over M*K: dX[m][k] = sum_n dY[m][n]*(Y>0) W[n][k]
  const float *gy, *y;  // y used for the ReLU mask when relu != 0
  const float* w;
  float* gx;
  int N, K, ldy, ldgx, relu;
  OAT_HD void operator()(int64_t gid) const {
    const int k = (int)(gid % K);
    const int64_t m = gid / K;
    float acc = 0.0f;
    for (int n = 0; n < N; ++n) {
      float d = gy[m * ldy + n];
      if (relu && !(y[m * ldy + n] > 0.0f)) d = 0.0f;
      acc = fmaf(d, w[(int64_t)n * K + k], acc);
    }
    gx[m * ldgx + k] = acc;
  }
};

struct LinearBwdW {  // gid over N*(K+1); column K is the bias
  const float *gy, *y, *x;
  float *gw, *gb;
  int64_t M;
  int N, K, ldy, ldx, relu;
  OAT_HD void operator()(int64_t gid) const {
    const int k = (int)(gid % (K + 1));
    const int n = (int)(gid / (K + 1));
    float acc = 0.0f;
    for (int64_t m = 0; m < M; ++m) {
      float d = gy[m * ldy + n];
      if (relu && !(y[m * ldy + n] > 0.0f)) d = 0.0f;
      acc = fmaf(d, k < K ? x[m * ldx + k] : 1.0f, acc);
    }
    if (k < K) gw[(int64_t)n * K + k] = acc;
    else gb[n] = acc;
  }
};

struct CopyCols {  // dst[m][off + j] = src[m][j];  gid over M*S
  const float* src;
  float* dst;
  int S, ld, off;
  OAT_HD void operator()(int64_t gid) const {
    const int j = (int)(gid % S);
    const int64_t m = gid / S;
    dst[m * ld + off + j] = src[m * S + j];
  }
};

// ---------------------------------------------------------------------------------
// Decoders.  Pass 1: one work item per batch row runs the recurrence forward and the
// reverse sweep, leaving per-(row, step) records in a scratch buffer [B][T][kDecRecord].
// Pass 2: the parameter gradients are reductions of outer products over those records,
// one work item per output element, summed in a fixed order (deterministic, no atomics).
// GRUCell (torch gate order r|z|n):  r = s(Wir u + bir + Whr h + bhr), g = s(...z...),
// n = tanh(Win u + bin + r * (Whn h + bhn)),  h' = (1-g) n + g h.
// ---------------------------------------------------------------------------------
// record layout (floats)
constexpr int kRecH = 0;       // h_prev [64]
constexpr int kRecR = 64;      // r [64]
constexpr int kRecG = 128;     // g (update gate) [64]
constexpr int kRecN = 192;     // n [64]
constexpr int kRecHn = 256;    // Whn h + bhn [64]
constexpr int kRecA1 = 320;    // DIM head hidden layer, post-ReLU [32]
constexpr int kRecMisc = 352;  // DIM: x0 x1 sigma0 sigma1 sraw0 sraw1 u0 u1; CIL: u0 u1 sign0 sign1
constexpr int kRecHcur = 360;  // h' of this step [64]
constexpr int kRecDgi = 424;   // d loss / d (W_ih u + b_ih) [192]
constexpr int kRecDgh = 616;   // d loss / d (W_hh h + b_hh) [192]
constexpr int kRecDa1 = 808;   // DIM: d loss / d (head pre-activation), ReLU-masked [32]
constexpr int kRecDout = 840;  // DIM: d loss / d o [4]; CIL: d loss / d (W_o h + b_o) [2]
constexpr int kDecRecord = 848;

struct DecParams {
  const float *wih, *whh, *bih, *bhh;  // [192][2], [192][64], [192], [192]
  const float *w1, *b1, *w2, *b2;      // DIM head [32][64],[32],[4][32],[4]; CIL: w1=[2][64], b1=[2]
  float *gwih, *gwhh, *gbih, *gbhh, *gw1, *gb1, *gw2, *gb2;
};

OAT_HD void gru_forward(const DecParams& p, const float* u, const float* h, float* rec) {
  for (int j = 0; j < 64; ++j) {
    // two partial sums per gate and float4 weight loads: six independent FMA chains
    float ar = p.bhh[j], ag = p.bhh[64 + j], an = p.bhh[128 + j], ar2 = 0.0f, ag2 = 0.0f, an2 = 0.0f;
    const float *wr = p.whh + j * 64, *wg = p.whh + (64 + j) * 64, *wn = p.whh + (128 + j) * 64;
    OAT_UNROLL
    for (int k = 0; k < 64; k += 4) {
      const F4 a = ld4(wr + k), b = ld4(wg + k), c = ld4(wn + k);
      ar = fmaf(a.x, h[k], ar); ar2 = fmaf(a.y, h[k + 1], ar2);
      ar = fmaf(a.z, h[k + 2], ar); ar2 = fmaf(a.w, h[k + 3], ar2);
      ag = fmaf(b.x, h[k], ag); ag2 = fmaf(b.y, h[k + 1], ag2);
      ag = fmaf(b.z, h[k + 2], ag); ag2 = fmaf(b.w, h[k + 3], ag2);
      an = fmaf(c.x, h[k], an); an2 = fmaf(c.y, h[k + 1], an2);
      an = fmaf(c.z, h[k + 2], an); an2 = fmaf(c.w, h[k + 3], an2);
    }
    ar += ar2;
    ag += ag2;
    an += an2;
    const float ir = fmaf(p.wih[j * 2 + 1], u[1], fmaf(p.wih[j * 2], u[0], p.bih[j]));
    const float ig = fmaf(p.wih[(64 + j) * 2 + 1], u[1], fmaf(p.wih[(64 + j) * 2], u[0], p.bih[64 + j]));
    const float in = fmaf(p.wih[(128 + j) * 2 + 1], u[1], fmaf(p.wih[(128 + j) * 2], u[0], p.bih[128 + j]));
    const float r = sigmoidf_(ir + ar), g = sigmoidf_(ig + ag);
    const float n = tanhf(fmaf(r, an, in));
    rec[kRecH + j] = h[j];
    rec[kRecR + j] = r;
    rec[kRecG + j] = g;
    rec[kRecN + j] = n;
    rec[kRecHn + j] = an;
    rec[kRecHcur + j] = fmaf(g, h[j] - n, n);  // (1-g) n + g h
  }
}

// Consumes dh (gradient wrt the step's output state), stores the gate gradients in the
// record, overwrites dh with the gradient wrt the previous state, returns d loss / d u.
OAT_HD void gru_backward(const DecParams& p, float* rec, float* dh, float* du) {
  float dh_prev[64];
  for (int k = 0; k < 64; ++k) dh_prev[k] = 0.0f;
  du[0] = du[1] = 0.0f;
  for (int j = 0; j < 64; ++j) {
    const float h = rec[kRecH + j], r = rec[kRecR + j], g = rec[kRecG + j], n = rec[kRecN + j],
                an = rec[kRecHn + j];
    const float d = dh[j];
    const float dn_pre = d * (1.0f - g) * (1.0f - n * n);
    const float dg_pre = d * (h - n) * g * (1.0f - g);
    const float dr_pre = dn_pre * an * r * (1.0f - r);
    const float dhn = dn_pre * r;  // gradient wrt (Whn h + bhn)
    dh_prev[j] = fmaf(d, g, dh_prev[j]);
    const float gi[3] = {dr_pre, dg_pre, dn_pre};
    const float gh[3] = {dr_pre, dg_pre, dhn};
    for (int q = 0; q < 3; ++q) {
      const int row = q * 64 + j;
      rec[kRecDgi + row] = gi[q];
      rec[kRecDgh + row] = gh[q];
      du[0] = fmaf(gi[q], p.wih[row * 2], du[0]);
      du[1] = fmaf(gi[q], p.wih[row * 2 + 1], du[1]);
      const float* wrow = p.whh + row * 64;
      OAT_UNROLL
      for (int k = 0; k < 64; k += 4) {
        const F4 wv = ld4(wrow + k);
        dh_prev[k] = fmaf(gh[q], wv.x, dh_prev[k]);
        dh_prev[k + 1] = fmaf(gh[q], wv.y, dh_prev[k + 1]);
        dh_prev[k + 2] = fmaf(gh[q], wv.z, dh_prev[k + 2]);
        dh_prev[k + 3] = fmaf(gh[q], wv.w, dh_prev[k + 3]);
      }
    }
  }
  for (int k = 0; k < 64; ++k) dh[k] = dh_prev[k];
}

struct DimNllStep {  // gid over B: loss = -mean_b(log_prob - logabsdet) (dim/train.py:196-199)
  DecParams p;
  const float* z;  // [B][64]
  const float* y;  // [B][T][2] (perturbed targets)
  float* scratch;  // [B][T][kDecRecord]
  float* gz;       // [B][64]
  double* loss;    // accumulates the batch SUM of row losses; pre-zeroed
  int B, T;
  OAT_HD void operator()(int64_t b) const {
    float h[64];
    for (int k = 0; k < 64; ++k) h[k] = z[b * 64 + k];
    float* rec0 = scratch + b * T * kDecRecord;
    const float* yb = y + b * T * 2;
    const float invB = 1.0f / (float)B;
    float row_loss = (float)T * 1.8378770664093453f;  // T * log(2 pi)
    for (int t = 0; t < T; ++t) {
      float* rec = rec0 + t * kDecRecord;
      float* misc = rec + kRecMisc;
      float u[2] = {0.0f, 0.0f};
      if (t > 0) { u[0] = yb[(t - 1) * 2]; u[1] = yb[(t - 1) * 2 + 1]; }
      misc[6] = u[0];
      misc[7] = u[1];
      gru_forward(p, u, h, rec);
      const float* hn = rec + kRecHcur;
      float* a1 = rec + kRecA1;
      for (int j = 0; j < 32; ++j) {
        float acc = p.b1[j], acc2 = 0.0f;
        OAT_UNROLL
        for (int k = 0; k < 64; k += 4) {
          const F4 wv = ld4(p.w1 + j * 64 + k);
          acc = fmaf(wv.x, hn[k], acc); acc2 = fmaf(wv.y, hn[k + 1], acc2);
          acc = fmaf(wv.z, hn[k + 2], acc); acc2 = fmaf(wv.w, hn[k + 3], acc2);
        }
        a1[j] = fmaxf(acc + acc2, 0.0f);
      }
      float o[4];
      for (int i = 0; i < 4; ++i) {
        float acc = p.b2[i];
        for (int j = 0; j < 32; ++j) acc = fmaf(p.w2[i * 32 + j], a1[j], acc);
        o[i] = acc;
      }
      for (int d = 0; d < 2; ++d) {
        const float mu = u[d] + o[d];
        const float sraw = o[2 + d];
        const float sp = sraw > 20.0f ? sraw : log1pf(expf(sraw));  // F.softplus (threshold 20)
        const float sigma = sp + 1e-3f;
        const float x = (yb[t * 2 + d] - mu) / sigma;
        misc[d] = x;
        misc[2 + d] = sigma;
        misc[4 + d] = sraw;
        row_loss += 0.5f * x * x + logf(sigma);
      }
      for (int k = 0; k < 64; ++k) h[k] = hn[k];
    }
    OAT_ATOMIC_ADD(loss, (double)row_loss);

    float dh[64];
    for (int k = 0; k < 64; ++k) dh[k] = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
      float* rec = rec0 + t * kDecRecord;
      const float* a1 = rec + kRecA1;
      const float* misc = rec + kRecMisc;
      float dout[4];
      for (int d = 0; d < 2; ++d) {
        const float x = misc[d], sigma = misc[2 + d], sraw = misc[4 + d];
        const float dx = x * invB;
        dout[d] = -dx / sigma;                         // d/d mu
        const float dsigma = (invB - dx * x) / sigma;  // log sigma term + x = (y-mu)/sigma
        dout[2 + d] = dsigma * (sraw > 20.0f ? 1.0f : sigmoidf_(sraw));
      }
      for (int i = 0; i < 4; ++i) rec[kRecDout + i] = dout[i];
      for (int j = 0; j < 32; ++j) {
        float da = 0.0f;
        for (int i = 0; i < 4; ++i) da = fmaf(dout[i], p.w2[i * 32 + j], da);
        if (!(a1[j] > 0.0f)) da = 0.0f;
        rec[kRecDa1 + j] = da;
        if (da != 0.0f) {
          OAT_UNROLL
          for (int k = 0; k < 64; k += 4) {
            const F4 wv = ld4(p.w1 + j * 64 + k);
            dh[k] = fmaf(da, wv.x, dh[k]);
            dh[k + 1] = fmaf(da, wv.y, dh[k + 1]);
            dh[k + 2] = fmaf(da, wv.z, dh[k + 2]);
            dh[k + 3] = fmaf(da, wv.w, dh[k + 3]);
          }
        }
      }
      float du[2];
      gru_backward(p, rec, dh, du);  // inputs are data: du is discarded
    }
    for (int k = 0; k < 64; ++k) gz[b * 64 + k] = dh[k];
  }
};

struct CilL1Step {  // gid over B: loss = mean_b sum_{t,d} |x_t - y_t| (cil/train.py:178-182)
  DecParams p;     // w1 = _output.weight [2][64], b1 = _output.bias [2]
  const float* z;
  const float* y;  // targets [B][T][2]
  float* scratch;  // [B][T][kDecRecord]
  float* gz;
  float* pred;     // [B][T][2] or null
  double* loss;
  int B, T;
  OAT_HD void operator()(int64_t b) const {
    float h[64];
    for (int k = 0; k < 64; ++k) h[k] = z[b * 64 + k];
    float* rec0 = scratch + b * T * kDecRecord;
    const float* yb = y + b * T * 2;
    const float invB = 1.0f / (float)B;
    float x[2] = {0.0f, 0.0f};
    float row_loss = 0.0f;
    for (int t = 0; t < T; ++t) {
      float* rec = rec0 + t * kDecRecord;
      float* misc = rec + kRecMisc;
      misc[0] = x[0];
      misc[1] = x[1];
      gru_forward(p, x, h, rec);
      const float* hn = rec + kRecHcur;
      for (int d = 0; d < 2; ++d) {
        float acc = p.b1[d];
        for (int k = 0; k < 64; ++k) acc = fmaf(p.w1[d * 64 + k], hn[k], acc);
        x[d] += acc;
        const float e = x[d] - yb[t * 2 + d];
        misc[2 + d] = e > 0.0f ? 1.0f : (e < 0.0f ? -1.0f : 0.0f);
        row_loss += fabsf(e);
        if (pred) pred[(b * T + t) * 2 + d] = x[d];
      }
      for (int k = 0; k < 64; ++k) h[k] = hn[k];
    }
    OAT_ATOMIC_ADD(loss, (double)row_loss);

    float dh[64];
    for (int k = 0; k < 64; ++k) dh[k] = 0.0f;
    float dx[2] = {0.0f, 0.0f};  // gradient wrt x_t flowing back from later steps
    for (int t = T - 1; t >= 0; --t) {
      float* rec = rec0 + t * kDecRecord;
      const float* misc = rec + kRecMisc;
      for (int d = 0; d < 2; ++d) {
        dx[d] += misc[2 + d] * invB;  // x_t = x_{t-1} + W_o h_t + b_o
        rec[kRecDout + d] = dx[d];
        for (int k = 0; k < 64; ++k) dh[k] = fmaf(dx[d], p.w1[d * 64 + k], dh[k]);
      }
      float du[2];
      gru_backward(p, rec, dh, du);
      dx[0] += du[0];  // x_{t-1} also feeds the GRU input of step t
      dx[1] += du[1];
    }
    for (int k = 0; k < 64; ++k) gz[b * 64 + k] = dh[k];
  }
};

// ---- pass 2: parameter gradients = sums of outer products over the B*T records ----
struct DecGradWhh {  // gid over 192*64: gW_hh[row][k] = sum dgh[row] * h_prev[k]
  const float* scratch;
  float* gwhh;
  int64_t records;  // B*T
  OAT_HD void operator()(int64_t gid) const {
    const int k = (int)(gid & 63), row = (int)(gid >> 6);
    float acc = 0.0f;
    for (int64_t i = 0; i < records; ++i) {
      const float* rec = scratch + i * kDecRecord;
      acc = fmaf(rec[kRecDgh + row], rec[kRecH + k], acc);
    }
    gwhh[gid] = acc;
  }
};

struct DecGradIh {  // gid over 192*4: gW_ih[row][0..1], gb_ih[row], gb_hh[row]
  const float* scratch;
  float *gwih, *gbih, *gbhh;
  int64_t records;
  int u_off;  // where the step's GRU input lives in the record
  OAT_HD void operator()(int64_t gid) const {
    const int col = (int)(gid & 3), row = (int)(gid >> 2);
    float acc = 0.0f;
    for (int64_t i = 0; i < records; ++i) {
      const float* rec = scratch + i * kDecRecord;
      if (col < 2) acc = fmaf(rec[kRecDgi + row], rec[u_off + col], acc);
      else acc += col == 2 ? rec[kRecDgi + row] : rec[kRecDgh + row];
    }
    if (col < 2) gwih[row * 2 + col] = acc;
    else if (col == 2) gbih[row] = acc;
    else gbhh[row] = acc;
  }
};

struct DecGradHead {  // gid over n1*65 + n2*33 (n2 = 0 for CIL)
  // DIM: gW1[j][k] = sum da1[j] hcur[k], gb1 (n1 = 32); gW2[i][j] = sum dout[i] a1[j], gb2 (n2 = 4)
  // CIL: gW_o[d][k] = sum dx[d] hcur[k], gb_o (n1 = 2, the "da1" slot is kRecDout)
  const float* scratch;
  float *gw1, *gb1, *gw2, *gb2;
  int64_t records;
  int n1, n2, d1_off;
  OAT_HD void operator()(int64_t gid) const {
    float acc = 0.0f;
    if (gid < (int64_t)n1 * 65) {
      const int k = (int)(gid % 65), j = (int)(gid / 65);
      for (int64_t i = 0; i < records; ++i) {
        const float* rec = scratch + i * kDecRecord;
        acc = fmaf(rec[d1_off + j], k < 64 ? rec[kRecHcur + k] : 1.0f, acc);
      }
      if (k < 64) gw1[j * 64 + k] = acc;
      else gb1[j] = acc;
    } else {
      const int64_t e = gid - (int64_t)n1 * 65;
      const int j = (int)(e % 33), i2 = (int)(e / 33);
      for (int64_t i = 0; i < records; ++i) {
        const float* rec = scratch + i * kDecRecord;
        acc = fmaf(rec[kRecDout + i2], j < 32 ? rec[kRecA1 + j] : 1.0f, acc);
      }
      if (j < 32) gw2[i2 * 32 + j] = acc;
      else gb2[i2] = acc;
    }
  }
};

struct LossFinalize {  // gid over 1
  double* acc;
  float* loss;
  int B;
  OAT_HD void operator()(int64_t) const {
    *loss = (float)(*acc / (double)B);
    *acc = 0.0;
  }
};

// ---------------------------------------------------------------------------------
// Optimiser: torch.optim.Adam (L2 weight decay folded into the gradient), with the
// optional global-norm clip of torch.nn.utils.clip_grad_norm_ (dim/train.py:207-208)
// ---------------------------------------------------------------------------------
struct SumSquares {  // gid over ceil(n/1024)
  const float* g;
  double* acc;
  int64_t n;
  float scale;
  OAT_HD void operator()(int64_t gid) const {
    const int64_t i0 = gid * 1024;
    const int64_t i1 = i0 + 1024 < n ? i0 + 1024 : n;
    float s = 0.0f;
    for (int64_t i = i0; i < i1; ++i) {
      const float v = g[i] * scale;
      s = fmaf(v, v, s);
    }
    OAT_ATOMIC_ADD(acc, (double)s);
  }
};

struct AdamStep {  // gid over n
  float* p;
  const float* g;
  float *m, *v;
  const double* sumsq;  // null: no clipping
  float lr, beta1, beta2, eps, weight_decay, grad_scale, clip_norm;
  float bias1, bias2_sqrt;  // 1 - beta1^t, sqrt(1 - beta2^t)
  OAT_HD void operator()(int64_t i) const {
    float coef = grad_scale;
    if (sumsq) {
      const float c = clip_norm / ((float)sqrt(*sumsq) + 1e-6f);
      if (c < 1.0f) coef *= c;
    }
    float gi = g[i] * coef;
    if (weight_decay != 0.0f) gi = fmaf(weight_decay, p[i], gi);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias2_sqrt + eps;
    p[i] -= (lr / bias1) * (mi / denom);
  }
};

}  // namespace train
}  // namespace oat
