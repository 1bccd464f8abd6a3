// Fused front of the BEV encoder: CTA bodies that keep the wide intermediate tensors of
// MobileNetV2's first blocks (torchvision, as wrapped by perception.py:25-55) in shared memory
// instead of HBM.  BatchNorms are folded at pack time (api.cu).
//
//   DwProjectBody     features.1: depthwise 3x3 (+BN+ReLU6) -> project 32->16 (+BN); one pair of
//                     output rows per small CTA, the depthwise output only in shared memory.
//   ExpandDwBody      expand 1x1 (cin -> 6 cin, +BN+ReLU6) -> depthwise 3x3 (stride 1|2,
//                     +BN+ReLU6) of one inverted-residual block, FP32: walks the image top to
//                     bottom; the 6x expanded tensor lives only as a ring of a few image rows.
//   ExpandDwPipeBody  the same block with the expand GEMM on tcgen05 (3xTF32) and the CTA split
//                     into a producer and a consumer half (see its header below).
//   FrontBody         features.0 (3x3 s2 conv C->32 as im2col + GEMM) + features.1 in one
//                     kernel (correct, not faster than separate launches; off by default).
//   The project 1x1 of features.2-4 stays on the tcgen05 GEMM (tc_gemm.cu).
//
// Ring bodies, per iteration: stage the next input rows (cp.async, issued ahead), run the
// pointwise convolution of the new rows into the row ring, slide the 3x3 depthwise window
// over the ring.
//
// A body is written against an executor `X` that provides shared memory, `phase(f)` =
// "run f(tid) for every thread, then barrier", asynchronous copies and (pipelined body) the
// tensor-core / hand-over primitives.  fused.cu instantiates it with the CUDA executors (the
// product); tests/emu instantiates it with a host loop so the `-m "not gpu"` suite can check
// the tiling, ring and padding logic against the oracle without a GPU (test tooling, never
// loaded by the package).
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OAT_FHD __host__ __device__ __forceinline__
#define OAT_FUNROLL _Pragma("unroll")
#else
#define OAT_FHD inline
#define OAT_FUNROLL
#endif

namespace oat {
namespace fused {

constexpr int kMaxModels = 16;

struct alignas(16) F4 {
  float x, y, z, w;
};
OAT_FHD F4 ld4(const float* p) { return *reinterpret_cast<const F4*>(p); }
OAT_FHD void st4(float* p, F4 v) { *reinterpret_cast<F4*>(p) = v; }
OAT_FHD F4 zero4() { return F4{0.0f, 0.0f, 0.0f, 0.0f}; }
OAT_FHD float relu6f(float v) { return fminf(fmaxf(v, 0.0f), 6.0f); }
OAT_FHD F4 relu6_4(F4 v) { return F4{relu6f(v.x), relu6f(v.y), relu6f(v.z), relu6f(v.w)}; }
OAT_FHD F4 fma4(F4 a, F4 b, F4 c) {
  return F4{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w)};
}
OAT_FHD float gload(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
OAT_FHD F4 gload4(const float* p) {
#if defined(__CUDA_ARCH__)
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return F4{v.x, v.y, v.z, v.w};
#else
  return ld4(p);
#endif
}
OAT_FHD void gstore4(float* p, F4 v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
#else
  st4(p, v);
#endif
}

struct Weights {  // per-model device pointers, passed by value
  const float* p[kMaxModels];
};

// ---------------------------------------------------------------------------------------
// out(p, c4) = bias[c4] + sum_k A[p][k] * W[k][c4]   for rows p < P, 4-channel groups c4 < N4.
// A: shared memory, row stride lda floats (lda % 8 == 4: conflict-free 16-byte row reads);
// W: shared memory [K][ldw]; K % 4 == 0.  Work item = (row group, c4), c4 fastest: the
// lanes of a warp read consecutive weight float4s and mostly one A row (broadcast).  A
// thread owns TP rows pg, pg+PG, pg+2PG, ... so every weight float4 feeds 4*TP FMAs.
// ---------------------------------------------------------------------------------------
template <int TP, class Emit>
OAT_FHD void gemm_rows(const float* A, int lda, const float* W, int ldw, const float* bias,
                       int P, int K, int N4, int tid, int nthreads, Emit emit) {
  const int PG = (P + TP - 1) / TP;
  for (int item = tid; item < PG * N4; item += nthreads) {
    const int pg = item / N4, c4 = item - pg * N4;
    const F4 b = ld4(bias + 4 * c4);
    F4 acc[TP];
    const float* ap[TP];
    OAT_FUNROLL
    for (int j = 0; j < TP; ++j) {
      acc[j] = b;
      const int p = pg + j * PG;
      ap[j] = A + (p < P ? p : P - 1) * lda;  // tail rows recompute the last row, never stored
    }
    const float* wp = W + 4 * c4;
    for (int k = 0; k < K; k += 4) {
      const F4 w0 = ld4(wp + (k + 0) * ldw), w1 = ld4(wp + (k + 1) * ldw);
      const F4 w2 = ld4(wp + (k + 2) * ldw), w3 = ld4(wp + (k + 3) * ldw);
      OAT_FUNROLL
      for (int j = 0; j < TP; ++j) {
        const F4 a = ld4(ap[j] + k);
        F4 c = acc[j];
        c.x = fmaf(a.x, w0.x, c.x); c.y = fmaf(a.x, w0.y, c.y); c.z = fmaf(a.x, w0.z, c.z); c.w = fmaf(a.x, w0.w, c.w);
        c.x = fmaf(a.y, w1.x, c.x); c.y = fmaf(a.y, w1.y, c.y); c.z = fmaf(a.y, w1.z, c.z); c.w = fmaf(a.y, w1.w, c.w);
        c.x = fmaf(a.z, w2.x, c.x); c.y = fmaf(a.z, w2.y, c.y); c.z = fmaf(a.z, w2.z, c.z); c.w = fmaf(a.z, w2.w, c.w);
        c.x = fmaf(a.w, w3.x, c.x); c.y = fmaf(a.w, w3.y, c.y); c.z = fmaf(a.w, w3.z, c.z); c.w = fmaf(a.w, w3.w, c.w);
        acc[j] = c;
      }
    }
    OAT_FUNROLL
    for (int j = 0; j < TP; ++j) {
      const int p = pg + j * PG;
      if (p < P) emit(p, c4, acc[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Depthwise 3x3 (pad 1, stride S) + bias + ReLU6 over a ring of RING image rows in shared
// memory.  One call produces nsub * ORD output rows: sub-pass `sub` reads the NW = S*(ORD-1)+3
// input rows starting at ir0 + sub*S*ORD and emits output rows sub*ORD .. sub*ORD+ORD-1.
// ring: [RING slots][win][ld] floats, the input row `ir` sits in slot (ir mod RING); rows
// outside the image hold zeros (the caller fills them), columns outside the image read as
// zeros here.  Work item = (sub-pass, column segment, c4), c4 fastest; the 3-column window
// slides along the segment in registers.  wd: [9][hid] (tap-major), bd: [hid], shared memory.
// emit(o, oc, c4, v): o-th output row of the call, column oc.
// ---------------------------------------------------------------------------------------
// GLOBAL = true: `ring` is the whole image in global memory ([RING = hin rows][win][ld]); rows
// outside it read as zeros and loads go through the read-only path.
// WSMEM = true: the 9 taps are re-read from shared memory at every use instead of living in 36
// registers (for small CTAs whose occupancy is register-bound).
template <int S, int ORD, int RING, bool GLOBAL = false, bool WSMEM = false, class Emit>
OAT_FHD void dw_rows(const float* ring, int ld, int win, int wout, const float* wd, const float* bd,
                     int hid, int ir0, int nsub, int nseg, int tid, int nthreads, Emit emit) {
  constexpr int NW = S * (ORD - 1) + 3;
  const int N4 = hid >> 2;
  for (int item = tid; item < nsub * nseg * N4; item += nthreads) {
    const int c4 = item % N4, seg = (item / N4) % nseg, sub = item / (N4 * nseg);
    const int c_lo = (seg * wout) / nseg, c_hi = ((seg + 1) * wout) / nseg;
    if (c_lo >= c_hi) continue;
    F4 k[WSMEM ? 1 : 9];
    if (!WSMEM) {
      OAT_FUNROLL
      for (int t = 0; t < (WSMEM ? 1 : 9); ++t) k[t] = ld4(wd + t * hid + 4 * c4);
    }
    const float* kp = wd + 4 * c4;
#define OAT_DW_TAP(t) (WSMEM ? ld4(kp + (t) * hid) : k[WSMEM ? 0 : (t)])
    const F4 bv = ld4(bd + 4 * c4);
    const float* rowp[NW];
    bool rowok[NW];
    OAT_FUNROLL
    for (int r = 0; r < NW; ++r) {
      const int ir = ir0 + sub * S * ORD + r;
      const int slot = (ir % RING + RING) % RING;
      rowok[r] = !GLOBAL || (ir >= 0 && ir < RING);
      rowp[r] = ring + (slot * win) * ld + 4 * c4;
    }
    F4 c0[NW], c1[NW], c2[NW];
    auto ldcol = [&](int ic, F4(&col)[NW]) {
      const bool ok = ic >= 0 && ic < win;
      OAT_FUNROLL
      for (int r = 0; r < NW; ++r)
        col[r] = (ok && rowok[r]) ? (GLOBAL ? gload4(rowp[r] + ic * ld) : ld4(rowp[r] + ic * ld)) : zero4();
    };
    ldcol(S * c_lo - 1, c0);
    ldcol(S * c_lo, c1);
    for (int oc = c_lo; oc < c_hi; ++oc) {
      ldcol(S * oc + 1, c2);
      OAT_FUNROLL
      for (int o = 0; o < ORD; ++o) {
        F4 part[3];  // one short FMA chain per kernel row (instruction-level parallelism)
        OAT_FUNROLL
        for (int kh = 0; kh < 3; ++kh) {
          F4 t = kh == 0 ? bv : zero4();
          t = fma4(c0[S * o + kh], OAT_DW_TAP(3 * kh + 0), t);
          t = fma4(c1[S * o + kh], OAT_DW_TAP(3 * kh + 1), t);
          t = fma4(c2[S * o + kh], OAT_DW_TAP(3 * kh + 2), t);
          part[kh] = t;
        }
        const F4 acc{(part[0].x + part[1].x) + part[2].x, (part[0].y + part[1].y) + part[2].y,
                     (part[0].z + part[1].z) + part[2].z, (part[0].w + part[1].w) + part[2].w};
        emit(sub * ORD + o, oc, c4, relu6_4(acc));
      }
      if (S == 1) {
        OAT_FUNROLL
        for (int r = 0; r < NW; ++r) { c0[r] = c1[r]; c1[r] = c2[r]; }
      } else {
        OAT_FUNROLL
        for (int r = 0; r < NW; ++r) c0[r] = c2[r];
        if (oc + 1 < c_hi) ldcol(S * (oc + 1), c1);
      }
    }
  }
#undef OAT_DW_TAP
}

// =======================================================================================
// ExpandDwBody<CIN, HID, S, HIN, OR, TP, NSEG>
//   in  [E][B][HIN][HIN][CIN]  (NHWC, output of the previous block)
//   out [E][B][HOUT][HOUT][HID] = relu6(dw3x3_S(relu6(in W_e + b_e)) + b_d)
// CTA = (model, image, split): `splits` CTAs share an image, each walks a contiguous range
// of output-row groups (OR rows per group) top to bottom.
// =======================================================================================
struct ExpandDwArgs {
  Weights we, be, wd, bd;  // expand [CIN][HID] + [HID]; depthwise [9][HID] + [HID]
  const float* in;
  float* out;
  int B;       // images per model
  int splits;  // work units per image (contiguous ranges of output-row groups)
  int units;   // E * B * splits (persistent kernels walk them with a stride)
};

template <int CIN_, int HID_, int S_, int HIN_, int OR_, int TP_, int NSEG_>
struct ExpandDwBody {
  static constexpr int CIN = CIN_, HID = HID_, S = S_, HIN = HIN_, OR = OR_, TP = TP_, NSEG = NSEG_;
  static constexpr int HOUT = (HIN - 1) / S + 1;
  static constexpr int NR = S * (OR - 1) + 3;   // ring rows
  static constexpr int NEW = S * OR;            // new input rows per iteration
  static constexpr int PRIME = NR - NEW;        // rows of the priming pass (3 - S)
  static constexpr int LDX = CIN + 4;           // staged input pixel stride
  static constexpr int LDE = HID + 4;           // ring pixel stride
  static constexpr int GROUPS = (HOUT + OR - 1) / OR;
  static_assert(PRIME >= 1 && PRIME <= NEW, "stride 1 needs OR >= 2");
  static_assert(CIN % 4 == 0 && HID % 4 == 0, "channel counts must be multiples of 4");
  // shared-memory map (floats)
  static constexpr int kWe = 0;
  static constexpr int kBe = kWe + CIN * HID;
  static constexpr int kWd = kBe + HID;
  static constexpr int kBd = kWd + 9 * HID;
  static constexpr int kX = kBd + HID;                    // 2 staging buffers
  static constexpr int kXFloats = NEW * HIN * LDX;
  static constexpr int kRing = kX + 2 * kXFloats;
  static constexpr int kSmemFloats = kRing + NR * HIN * LDE;

  template <class X>
  OAT_FHD static void run(X& x, const ExpandDwArgs& a, int cta) {
    float* sm = x.smem();
    const int nt = x.nthreads();
    const int split = cta % a.splits;
    const int img = cta / a.splits;  // model * B + b
    const int model = img / a.B;
    const float* in = a.in + (int64_t)img * HIN * HIN * CIN;
    float* out = a.out + (int64_t)img * HOUT * HOUT * HID;
    const int g0 = (split * GROUPS) / a.splits, g1 = ((split + 1) * GROUPS) / a.splits;
    if (g0 >= g1) return;
    float* We = sm + kWe;
    float* Be = sm + kBe;
    float* Wd = sm + kWd;
    float* Bd = sm + kBd;
    float* ring = sm + kRing;

    // input rows [lo, lo+n) clipped to the image -> staging buffer `buf` (pixel-major, LDX)
    auto stage = [&](int tid, int buf, int lo, int n) {
      const int vlo = lo < 0 ? 0 : lo;
      const int vhi = lo + n > HIN ? HIN : lo + n;
      if (vhi <= vlo) return;
      float* dst = sm + kX + buf * kXFloats;
      const float* src = in + (int64_t)vlo * HIN * CIN;
      const int chunks = (vhi - vlo) * HIN * (CIN / 4);
      for (int i = tid; i < chunks; i += nt) {
        const int px = i / (CIN / 4), q = i - px * (CIN / 4);
        x.async16(dst + px * LDX + 4 * q, src + px * CIN + 4 * q);
      }
    };
    // expand rows [lo, lo+n) from staging buffer `buf` into the ring (zeros outside the image)
    auto expand = [&](int tid, int buf, int lo, int n) {
      const int vlo = lo < 0 ? 0 : lo;
      const int vhi = lo + n > HIN ? HIN : lo + n;
      for (int ir = lo; ir < lo + n; ++ir) {
        if (ir >= 0 && ir < HIN) continue;
        float* row = ring + (((ir % NR) + NR) % NR) * HIN * LDE;
        for (int i = tid; i < HIN * LDE / 4; i += nt) st4(row + 4 * i, zero4());
      }
      if (vhi <= vlo) return;
      const float* Xs = sm + kX + buf * kXFloats;
      gemm_rows<TP>(Xs, LDX, We, HID, Be, (vhi - vlo) * HIN, CIN, HID / 4, tid, nt,
                    [&](int p, int c4, F4 v) {
                      const int r = p / HIN, col = p - r * HIN;
                      const int slot = (vlo + r) % NR;
                      st4(ring + (slot * HIN + col) * LDE + 4 * c4, relu6_4(v));
                    });
    };

    const int ir_first = S * g0 * OR - 1;  // first ring row of the first group
    x.phase([&](int tid) {
      const float *we = a.we.p[model], *be = a.be.p[model], *wd = a.wd.p[model], *bd = a.bd.p[model];
      for (int i = tid; i < CIN * HID / 4; i += nt) st4(We + 4 * i, gload4(we + 4 * i));
      for (int i = tid; i < HID / 4; i += nt) st4(Be + 4 * i, gload4(be + 4 * i));
      for (int i = tid; i < 9 * HID / 4; i += nt) st4(Wd + 4 * i, gload4(wd + 4 * i));
      for (int i = tid; i < HID / 4; i += nt) st4(Bd + 4 * i, gload4(bd + 4 * i));
      stage(tid, 0, ir_first, PRIME);
      stage(tid, 1, ir_first + PRIME, NEW);
      x.async_wait();
    });
    x.phase([&](int tid) { expand(tid, 0, ir_first, PRIME); });
    for (int g = g0; g < g1; ++g) {
      const int buf = (g - g0 + 1) & 1;
      const int ir0 = S * g * OR - 1;
      x.phase([&](int tid) {
        if (g + 1 < g1) stage(tid, buf ^ 1, ir0 + NEW + PRIME, NEW);  // rows of group g+1
        expand(tid, buf, ir0 + PRIME, NEW);
      });
      x.phase([&](int tid) {
        dw_rows<S, OR, NR>(ring, LDE, HIN, HOUT, Wd, Bd, HID, ir0, 1, NSEG, tid, nt,
                       [&](int o, int oc, int c4, F4 v) {
                         const int orow = g * OR + o;
                         if (orow < HOUT) gstore4(out + ((int64_t)orow * HOUT + oc) * HID + 4 * c4, v);
                       });
        x.async_wait();
      });
    }
  }
};

// =======================================================================================
// FrontBody: visual NCHW [B][C][100][100] (shared by the models) ->
//            out [E][B][50][50][16] = features.1(features.0(visual))
// =======================================================================================
struct FrontArgs {
  Weights ws, bs;  // stem [9*C][32] (k = (kh*3+kw)*C + c) + [32]
  Weights wd, bd;  // features.1 depthwise [9][32] + [32]
  Weights wp, bp;  // features.1 project [32][16] + [16]
  const float* vis;
  float* out;
  int B, C, splits;
};

struct FrontBody {
  static constexpr int HI = 100, HS = 50;      // input / stem-output size
  static constexpr int NR = 4, NEW = 2;        // ring rows, new stem rows per iteration
  static constexpr int XR = 2 * NEW + 1;       // input rows staged per iteration
  static constexpr int LDS_ = 36;              // ring / depthwise-output pixel stride (32 + 4)
  static constexpr int PAIRS = HS / 2;
  OAT_FHD static int kp(int C) { return (9 * C + 3) & ~3; }   // im2col depth, padded
  OAT_FHD static int ldi(int C) { return kp(C) + 4; }
  // shared-memory map (floats)
  OAT_FHD static int oWs(int) { return 0; }
  OAT_FHD static int oBs(int C) { return kp(C) * 32; }
  OAT_FHD static int oWd(int C) { return oBs(C) + 32; }
  OAT_FHD static int oBd(int C) { return oWd(C) + 9 * 32; }
  OAT_FHD static int oWp(int C) { return oBd(C) + 32; }
  OAT_FHD static int oBp(int C) { return oWp(C) + 32 * 16; }
  OAT_FHD static int oXI(int C) { return oBp(C) + 16; }                   // 2 x [C][XR][100]
  OAT_FHD static int oIM(int C) { return oXI(C) + 2 * C * XR * HI; }      // [NEW*50][ldi]
  OAT_FHD static int oRing(int C) { return oIM(C) + NEW * HS * ldi(C); }  // [NR][50][36]
  OAT_FHD static int oD(int C) { return oRing(C) + NR * HS * LDS_; }      // [2*50][36]
  OAT_FHD static int smem_floats(int C) { return oD(C) + 2 * HS * LDS_; }

  template <class X>
  OAT_FHD static void run(X& x, const FrontArgs& a, int cta) {
    float* sm = x.smem();
    const int nt = x.nthreads();
    const int C = a.C, KP = kp(C), LDI = ldi(C);
    const int split = cta % a.splits;
    const int img = cta / a.splits;  // model * B + b
    const int model = img / a.B, b = img - model * a.B;
    const float* vis = a.vis + (int64_t)b * C * HI * HI;
    float* out = a.out + (int64_t)img * HS * HS * 16;
    const int p0 = (split * PAIRS) / a.splits, p1 = ((split + 1) * PAIRS) / a.splits;
    if (p0 >= p1) return;
    float* Ws = sm + oWs(C);
    float* Bs = sm + oBs(C);
    float* Wd = sm + oWd(C);
    float* Bd = sm + oBd(C);
    float* Wp = sm + oWp(C);
    float* Bp = sm + oBp(C);
    float* IM = sm + oIM(C);
    float* ring = sm + oRing(C);
    float* D = sm + oD(C);

    // input rows 2*sr-1 .. 2*sr+3 of every channel for the stem rows [sr, sr+2)
    auto stage = [&](int tid, int buf, int sr) {
      float* dst = sm + oXI(C) + buf * C * XR * HI;
      const int lo = 2 * sr - 1;
      for (int i = tid; i < C * XR * (HI / 4); i += nt) {
        const int q = i % (HI / 4), r = (i / (HI / 4)) % XR, c = i / ((HI / 4) * XR);
        const int ir = lo + r;
        if (ir < 0 || ir >= HI) continue;
        x.async16(dst + (c * XR + r) * HI + 4 * q, vis + ((int64_t)c * HI + ir) * HI + 4 * q);
      }
    };
    // im2col of the stem rows [sr, sr+2): IM[(r*50+ow)][(kh*3+kw)*C + c]
    auto im2col = [&](int tid, int buf, int sr) {
      const float* XI = sm + oXI(C) + buf * C * XR * HI;
      for (int i = tid; i < NEW * HS * 9; i += nt) {
        const int t = i % 9, p = i / 9;
        const int kh = t / 3, kw = t - 3 * kh;
        const int r = p / HS, ow = p - r * HS;
        const int ir = 2 * (sr + r) - 1 + kh, ic = 2 * ow - 1 + kw;
        const bool ok = ir >= 0 && ir < HI && ic >= 0 && ic < HI;
        const int xr = 2 * r + kh;
        float* dst = IM + p * LDI + t * C;
        for (int c = 0; c < C; ++c) dst[c] = ok ? XI[(c * XR + xr) * HI + ic] : 0.0f;
      }
    };
    // stem GEMM of the rows [sr, sr+2) into the ring (rows outside the image: zeros)
    auto stem = [&](int tid, int sr) {
      gemm_rows<4>(IM, LDI, Ws, 32, Bs, NEW * HS, KP, 8, tid, nt, [&](int p, int c4, F4 v) {
        const int r = p / HS, ow = p - r * HS;
        const int row = sr + r;
        const bool ok = row >= 0 && row < HS;
        const int slot = ((row % NR) + NR) % NR;
        st4(ring + (slot * HS + ow) * LDS_ + 4 * c4, ok ? relu6_4(v) : zero4());
      });
    };

    const int sr_first = 2 * p0 - 1;  // first ring row of the first output pair
    x.phase([&](int tid) {
      const float *ws = a.ws.p[model], *bs = a.bs.p[model], *wd = a.wd.p[model], *bd = a.bd.p[model];
      const float *wp = a.wp.p[model], *bp = a.bp.p[model];
      for (int i = tid; i < KP * 32; i += nt) Ws[i] = i < 9 * C * 32 ? gload(ws + i) : 0.0f;
      for (int i = tid; i < 32; i += nt) { Bs[i] = gload(bs + i); Bd[i] = gload(bd + i); }
      for (int i = tid; i < 9 * 32; i += nt) Wd[i] = gload(wd + i);
      for (int i = tid; i < 32 * 16; i += nt) Wp[i] = gload(wp + i);
      for (int i = tid; i < 16; i += nt) Bp[i] = gload(bp + i);
      for (int i = tid; i < NEW * HS * LDI; i += nt) IM[i] = 0.0f;  // incl. the padded k columns
      stage(tid, 0, sr_first);
      x.async_wait();
    });
    // priming pass: stem rows sr_first, sr_first+1
    x.phase([&](int tid) {
      stage(tid, 1, sr_first + 2);
      im2col(tid, 0, sr_first);
    });
    x.phase([&](int tid) {
      stem(tid, sr_first);
      x.async_wait();
    });
    for (int p = p0; p < p1; ++p) {
      const int buf = (p - p0 + 1) & 1;
      const int sr = 2 * p + 1;  // new stem rows sr, sr+1 complete the window 2p-1 .. 2p+2
      x.phase([&](int tid) {
        if (p + 1 < p1) stage(tid, buf ^ 1, sr + 2);
        im2col(tid, buf, sr);
      });
      x.phase([&](int tid) { stem(tid, sr); });
      x.phase([&](int tid) {
        dw_rows<1, 2, NR>(ring, LDS_, HS, HS, Wd, Bd, 32, 2 * p - 1, 1, 16, tid, nt,
                      [&](int o, int oc, int c4, F4 v) { st4(D + (o * HS + oc) * LDS_ + 4 * c4, v); });
      });
      x.phase([&](int tid) {
        gemm_rows<2>(D, LDS_, Wp, 16, Bp, 2 * HS, 32, 4, tid, nt, [&](int q, int c4, F4 v) {
          gstore4(out + ((int64_t)(2 * p) * HS + q) * 16 + 4 * c4, v);
        });
        x.async_wait();
      });
    }
  }
};


// =======================================================================================
// DwProjectBody: features.1 after the stem — depthwise 3x3 (32 ch, stride 1, +BN+ReLU6) and
// project 1x1 32->16 (+BN) in one kernel.  in [E][B][50][50][32] -> out [E][B][50][50][16].
// Small CTAs (one pair of output rows each, 16 KB of shared memory) instead of a ring: the
// depthwise window slides over the input rows in global memory / L2, its output goes to shared
// memory, the project GEMM reads it from there.  Many CTAs per SM hide the two barriers.
// =======================================================================================
struct DwProjectArgs {
  Weights wd, bd;  // depthwise [9][32] + [32]
  Weights wp, bp;  // project [32][16] + [16]
  const float* in;
  float* out;
  int B;
};

struct DwProjectBody {
  static constexpr int H = 50, CH = 32, CO = 16, LDD = CH + 4;
  static constexpr int PAIRS = H / 2;  // CTAs per image
  static constexpr int kWd = 0;
  static constexpr int kBd = kWd + 9 * CH;
  static constexpr int kWp = kBd + CH;
  static constexpr int kBp = kWp + CH * CO;
  static constexpr int kD = kBp + CO;              // [2*50][36]
  static constexpr int kSmemFloats = kD + 2 * H * LDD;

  template <class X>
  OAT_FHD static void run(X& x, const DwProjectArgs& a, int cta) {
    float* sm = x.smem();
    const int nt = x.nthreads();
    const int pair = cta % PAIRS;
    const int img = cta / PAIRS;  // model * B + b
    const int model = img / a.B;
    const float* in = a.in + (int64_t)img * H * H * CH;
    float* out = a.out + ((int64_t)img * H + 2 * pair) * H * CO;
    float* Wd = sm + kWd;
    float* Bd = sm + kBd;
    float* Wp = sm + kWp;
    float* Bp = sm + kBp;
    float* D = sm + kD;
    x.phase([&](int tid) {
      const float *wd = a.wd.p[model], *bd = a.bd.p[model], *wp = a.wp.p[model], *bp = a.bp.p[model];
      for (int i = tid; i < 9 * CH / 4; i += nt) st4(Wd + 4 * i, gload4(wd + 4 * i));
      for (int i = tid; i < CH / 4; i += nt) st4(Bd + 4 * i, gload4(bd + 4 * i));
      for (int i = tid; i < CH * CO / 4; i += nt) st4(Wp + 4 * i, gload4(wp + 4 * i));
      for (int i = tid; i < CO / 4; i += nt) st4(Bp + 4 * i, gload4(bp + 4 * i));
    });
    x.phase([&](int tid) {
      dw_rows<1, 2, H, true, true>(in, CH, H, H, Wd, Bd, CH, 2 * pair - 1, 1, 16, tid, nt,
                             [&](int o, int oc, int c4, F4 v) { st4(D + (o * H + oc) * LDD + 4 * c4, v); });
    });
    x.phase([&](int tid) {
      gemm_rows<4>(D, LDD, Wp, CO, Bp, 2 * H, CH, CO / 4, tid, nt, [&](int q, int c4, F4 v) {
        gstore4(out + (int64_t)q * CO + 4 * c4, v);
      });
    });
  }
};

// =======================================================================================
// Tensor-core, pipelined variant (the default on the GPU).
//
// Same walk over the image, but (1) the pointwise convolution of the new rows is ONE M=128
// tcgen05 GEMM per row group (3xTF32, accumulator in TMEM) and (2) the CTA is split into a
// PRODUCER half (stages input rows, writes the UMMA operand, issues the GEMM, moves the
// accumulator through bias + ReLU6 into the row ring) and a CONSUMER half (slides the 3x3
// depthwise window over the ring and stores the block's depthwise output), coupled by
// ready/done barriers per row group, so the latency-bound producer chain of group g+1 hides
// behind the FMA-bound depthwise pass of group g.  The ring holds NR + NEW rows: the window
// of the group being consumed plus the new rows of the next one.  CTAs are persistent: each
// walks work units (model, image, row split) with a stride and reloads weights only when
// the model changes.
//
// Executor interface beyond smem()/async16():
//   kConcurrent                 true: the two halves run concurrently (CUDA); false: one
//                               sequential thread of control (host), producer one group ahead
//   all_phase(f)                barrier; f(tid, nthreads) on every thread of the CTA; barrier
//   is_producer(), p_threads(), p_phase(f), p_phase_nosync(f)   producer half (own barrier)
//   c_threads(), c_run(f)       consumer half
//   op_store4(tile, rows, row, k, v)   4 consecutive-k elements of an operand row: A tiles
//                               [128 rows][32 k], weight tiles [N rows][32 k] per k-block (device:
//                               TF32 hi/lo split, K-major 128B-swizzled UMMA layout, lo tile
//                               right behind the hi tile)
//   mma(buf, acc, N, a, b, K)   producer: D[acc .. acc+N) (128 x N) = A B^T, asynchronous; buf 0|1
//                               names the completion barrier (two GEMMs may be in flight)
//   epilogue(buf, acc, N, emit) producer: waits for that GEMM; emit(row, c4, F4) per accumulator
//                               row and 4-column group
//   async_commit(), async_wait_all(), async_wait_but_last()   cp.async group bookkeeping
//   signal_ready(i) / wait_ready(i), signal_done(i) / wait_done(i)   group i handed over / retired
// =======================================================================================
constexpr int kTileRows = 128;                      // UMMA M
constexpr int kATileFloats = 2 * kTileRows * 32;    // hi + lo halves of one A k-block
OAT_FHD constexpr int b_tile_floats(int n) { return 2 * n * 32; }

template <int CIN_, int HID_, int S_, int HIN_, int OR_, int NSEG_>
struct ExpandDwPipeBody {
  static constexpr int CIN = CIN_, HID = HID_, S = S_, HIN = HIN_, OR = OR_, NSEG = NSEG_;
  static constexpr int HOUT = (HIN - 1) / S + 1;
  static constexpr int NR = S * (OR - 1) + 3;   // rows of one depthwise window
  static constexpr int NEW = S * OR;            // new input rows per group
  static constexpr int PRIME = NR - NEW;        // rows of the priming pass (3 - S)
  static constexpr int RC = NR + NEW;           // ring capacity (rows)
  static constexpr int LDX = CIN + 4;
  static constexpr int LDE = HID + 4;
  static constexpr int GROUPS = (HOUT + OR - 1) / OR;
  static_assert(PRIME >= 1 && PRIME <= NEW, "stride 1 needs OR >= 2");
  static_assert(NEW * HIN <= kTileRows, "one group must fit one M=128 tile");
  static_assert(CIN % 8 == 0 && CIN <= 32, "one k-block of TF32 k-slices");
  static_assert(HID % 16 == 0 && HID <= 256, "UMMA N");
  // shared-memory map (floats); operand tiles are 1024-byte aligned
  static constexpr int kA = 0;                                 // 2 A tiles (double-buffered GEMM)
  static constexpr int kB = kA + 2 * kATileFloats;
  static constexpr int kBe = kB + b_tile_floats(HID);
  static constexpr int kWd = kBe + HID;
  static constexpr int kBd = kWd + 9 * HID;
  static constexpr int kX = kBd + HID;                         // 3 staging buffers (prefetch distance 2)
  static constexpr int kXFloats = NEW * HIN * LDX;
  static constexpr int kRing = kX + 3 * kXFloats;
  static constexpr int kSmemFloats = kRing + RC * HIN * LDE;
  static constexpr int kTmemCols = 2 * HID <= 128 ? 128 : (2 * HID <= 256 ? 256 : 512);

  template <class X>
  OAT_FHD static void run(X& x, const ExpandDwArgs& a, int first_unit, int unit_stride) {
    float* sm = x.smem();
    float* Bt = sm + kB;
    float* Be = sm + kBe;
    float* Wd = sm + kWd;
    float* Bd = sm + kBd;
    float* ring = sm + kRing;
    const int pt = x.p_threads(), ct = x.c_threads();
    int cur_model = -1;
    uint32_t it = 0;  // row groups handed over so far (the same count in both halves)
    for (int unit = first_unit; unit < a.units; unit += unit_stride) {
      const int split = unit % a.splits;
      const int img = unit / a.splits;  // model * B + b
      const int model = img / a.B;
      const int g0 = (split * GROUPS) / a.splits, g1 = ((split + 1) * GROUPS) / a.splits;
      if (g0 >= g1) continue;
      const int n = g1 - g0;
      if (model != cur_model) {
        x.all_phase([&](int tid, int nt) {
          const float *we = a.we.p[model], *be = a.be.p[model], *wd = a.wd.p[model], *bd = a.bd.p[model];
          for (int i = tid; i < (CIN / 4) * HID; i += nt) {  // W_e [CIN][HID] -> weight tile rows n
            const int nn = i % HID, k4 = i / HID;
            const F4 v{gload(we + (4 * k4 + 0) * HID + nn), gload(we + (4 * k4 + 1) * HID + nn),
                       gload(we + (4 * k4 + 2) * HID + nn), gload(we + (4 * k4 + 3) * HID + nn)};
            x.op_store4(Bt, HID, nn, 4 * k4, v);
          }
          for (int i = tid; i < HID / 4; i += nt) st4(Be + 4 * i, gload4(be + 4 * i));
          for (int i = tid; i < 9 * HID / 4; i += nt) st4(Wd + 4 * i, gload4(wd + 4 * i));
          for (int i = tid; i < HID / 4; i += nt) st4(Bd + 4 * i, gload4(bd + 4 * i));
        });
        cur_model = model;
      }
      const float* in = a.in + (int64_t)img * HIN * HIN * CIN;
      float* out = a.out + (int64_t)img * HOUT * HOUT * HID;

      // Row chunk j of the unit: j = -1 the priming rows, j >= 0 the new rows of group g0+j.
      // Its staging buffer is (j+1) % 3, its A tile / accumulator / GEMM barrier (j+1) & 1.
      const int ir_first = S * g0 * OR - 1;
      auto chunk_lo = [&](int j) { return j < 0 ? ir_first : ir_first + PRIME + j * NEW; };
      auto chunk_n = [&](int j) { return j < 0 ? PRIME : NEW; };
      auto valid_pixels = [&](int j, int* vlo_out) {
        const int lo = chunk_lo(j), hi = lo + chunk_n(j);
        const int vlo = lo < 0 ? 0 : lo;
        const int vhi = hi > HIN ? HIN : hi;
        *vlo_out = vlo;
        return vhi > vlo ? (vhi - vlo) * HIN : 0;
      };
      // ---- producer pieces -------------------------------------------------------------
      auto stage = [&](int tid, int j) {  // input rows of chunk j -> staging buffer (cp.async)
        int vlo;
        const int np = valid_pixels(j, &vlo);
        float* dst = sm + kX + ((j + 1) % 3) * kXFloats;
        const float* src = in + (int64_t)vlo * HIN * CIN;
        for (int i = tid; i < np * (CIN / 4); i += pt) {
          const int px = i / (CIN / 4), q = i - px * (CIN / 4);
          x.async16(dst + px * LDX + 4 * q, src + px * CIN + 4 * q);
        }
        x.async_commit();
      };
      auto issue = [&](int j) {  // staged pixels -> A tile, expand GEMM issued (asynchronous)
        int vlo;
        const int np = valid_pixels(j, &vlo);
        if (np == 0) return;
        const float* Xs = sm + kX + ((j + 1) % 3) * kXFloats;
        float* At = sm + kA + ((j + 1) & 1) * kATileFloats;
        x.p_phase_nosync([&](int tid) {
          for (int i = tid; i < np * (CIN / 4); i += pt) {
            const int px = i / (CIN / 4), q = i - px * (CIN / 4);
            x.op_store4(At, kTileRows, px, 4 * q, ld4(Xs + px * LDX + 4 * q));
          }
        });
        x.mma((j + 1) & 1, ((j + 1) & 1) * HID, HID, At, Bt, CIN);
      };
      auto collect = [&](int j) {  // accumulator -> ring rows (bias + ReLU6), zeros outside the image
        int vlo;
        const int np = valid_pixels(j, &vlo);
        if (np > 0)
          x.epilogue((j + 1) & 1, ((j + 1) & 1) * HID, HID, [&](int p, int c4, F4 v) {
            if (p >= np) return;
            const int r = p / HIN, col = p - r * HIN;
            const int slot = (vlo + r) % RC;
            const F4 b = ld4(Be + 4 * c4);
            st4(ring + (slot * HIN + col) * LDE + 4 * c4,
                relu6_4(F4{v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w}));
          });
        x.p_phase([&](int tid) {
          const int lo = chunk_lo(j);
          for (int ir = lo; ir < lo + chunk_n(j); ++ir) {
            if (ir >= 0 && ir < HIN) continue;
            float* row = ring + (((ir % RC) + RC) % RC) * HIN * LDE;
            for (int i = tid; i < HIN * LDE / 4; i += pt) st4(row + 4 * i, zero4());
          }
        });
      };
      auto begin = [&]() {  // stage chunks -1, 0, 1; GEMMs of -1 and 0 in flight; ring primed
        x.p_phase([&](int tid) {
          stage(tid, -1);
          stage(tid, 0);
          if (n > 1) stage(tid, 1);
          x.async_wait_all();
        });
        issue(-1);
        issue(0);
        collect(-1);
      };
      auto produce = [&](int j, uint32_t i) {  // finishes group g0+j; keeps chunk j+1's GEMM in flight
#if defined(OAT_ABL_NO_PRODUCE)  // timing ablation: the producer half only hands groups over
        if (X::kConcurrent && j >= 2) x.wait_done(i - 2);
        return;
#endif
        if (j + 1 < n) {
          x.p_phase([&](int tid) {
            if (j + 2 < n) {
              stage(tid, j + 2);
              x.async_wait_but_last();  // chunk j+1 has landed, chunk j+2 may be in flight
            } else {
              x.async_wait_all();
            }
          });
          issue(j + 1);
        }
        // the rows written next replace the oldest rows of the window of group i-2
        if (X::kConcurrent && j >= 2) x.wait_done(i - 2);
        collect(j);
      };
      // ---- consumer piece ----------------------------------------------------------------
      auto depthwise = [&](int tid, int g) {
#if defined(OAT_ABL_NO_DW)  // timing ablation: the consumer half only retires groups
        return;
#endif
        dw_rows<S, OR, RC>(ring, LDE, HIN, HOUT, Wd, Bd, HID, S * g * OR - 1, 1, NSEG, tid, ct,
                           [&](int o, int oc, int c4, F4 v) {
                             const int orow = g * OR + o;
                             if (orow < HOUT) gstore4(out + ((int64_t)orow * HOUT + oc) * HID + 4 * c4, v);
                           });
      };

      if (X::kConcurrent) {
        if (x.is_producer()) {
          if (it > 0) x.wait_done(it - 1);  // the previous unit's rows are retired
          begin();
          for (int j = 0; j < n; ++j) {
            produce(j, it + (uint32_t)j);
            x.signal_ready(it + (uint32_t)j);
          }
        } else {
          for (int j = 0; j < n; ++j) {
            x.wait_ready(it + (uint32_t)j);
            x.c_run([&](int tid) { depthwise(tid, g0 + j); });
            x.signal_done(it + (uint32_t)j);
          }
        }
      } else {  // sequential executor: the producer runs one group ahead, as it may on the GPU
        begin();
        produce(0, it);
        for (int j = 0; j < n; ++j) {
          if (j + 1 < n) produce(j + 1, it + (uint32_t)(j + 1));
          x.c_run([&](int tid) { depthwise(tid, g0 + j); });
        }
      }
      it += (uint32_t)n;
    }
  }
};

}  // namespace fused
}  // namespace oat
