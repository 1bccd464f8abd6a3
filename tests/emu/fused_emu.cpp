// TEST TOOLING — host execution of the fused encoder-front CTA bodies.
//
// Compiles oatomobile_b200/csrc/fused_body.h (the exact code the CUDA product runs inside
// fused.cu's kernels) with g++ and an executor whose "phase" is a plain loop over thread
// ids and whose asynchronous copies complete immediately, so the `-m "not gpu"` tests can
// check the tiling / ring / padding logic against the oracle without a GPU.  Never
// imported, linked or loaded by the oatomobile_b200 package.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../oatomobile_b200/csrc/fused_body.h"

namespace {

using namespace oat::fused;

struct HostExec {
  std::vector<float> buf;
  int nt;
  HostExec(int floats, int threads) : buf((size_t)floats + 4, 0.0f), nt(threads) {
    // poison so that reads of never-written shared memory show up as NaN in the outputs
    for (auto& v : buf) v = __builtin_nanf("");
  }
  float* smem() {  // 16-byte aligned
    uintptr_t p = reinterpret_cast<uintptr_t>(buf.data());
    return reinterpret_cast<float*>((p + 15) & ~uintptr_t(15));
  }
  int nthreads() const { return nt; }
  template <class F>
  void phase(F f) {
    for (int t = 0; t < nt; ++t) f(t);
  }
  void async16(float* dst, const float* src) { memcpy(dst, src, 16); }
  void async_wait() {}
  void async_commit() {}
  void async_wait_all() {}
  void async_wait_but_last() {}

  // ---- pipelined tensor-core interface as ONE sequential thread of control: the producer
  // half has `nt` threads, the consumer half `nt` threads; operand tiles are [rows][32] per
  // k-block (the lo halves stay unused), the GEMM runs when the accumulator is collected.
  static constexpr bool kConcurrent = false;
  bool is_producer() const { return true; }
  int p_threads() const { return nt; }
  int c_threads() const { return nt; }
  template <class F>
  void all_phase(F f) {
    for (int t = 0; t < 2 * nt; ++t) f(t, 2 * nt);
  }
  template <class F>
  void p_phase(F f) { phase(f); }
  template <class F>
  void p_phase_nosync(F f) { phase(f); }
  template <class F>
  void c_run(F f) { phase(f); }
  void signal_ready(uint32_t) {}
  void wait_ready(uint32_t) {}
  void signal_done(uint32_t) {}
  void wait_done(uint32_t) {}
  void op_store4(float* tile, int /*rows*/, int row, int k, F4 v) { st4(tile + row * 32 + k, v); }
  struct Pending { const float *a = nullptr, *b = nullptr; int k = 0, n = 0, acc = -1; } pend[2];
  void mma(int buf, int acc, int N, const float* a, const float* b, int K) {
    if (pend[buf].acc != -1) abort();  // the barrier of this buffer is still armed
    pend[buf].a = a; pend[buf].b = b; pend[buf].k = K; pend[buf].n = N; pend[buf].acc = acc;
  }
  template <class Emit>
  void epilogue(int buf, int acc, int N, Emit emit) {
    const Pending p = pend[buf];
    if (acc != p.acc || N != p.n) abort();  // collected something that was not issued
    for (int row = 0; row < kTileRows; ++row)
      for (int c4 = 0; c4 < N / 4; ++c4) {
        float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int k = 0; k < p.k; ++k) {
          const int kb = k / 32, kk = k % 32;
          const float av = p.a[kb * kATileFloats + row * 32 + kk];
          for (int j = 0; j < 4; ++j)
            o[j] = fmaf(av, p.b[kb * b_tile_floats(N) + (4 * c4 + j) * 32 + kk], o[j]);
        }
        emit(row, c4, F4{o[0], o[1], o[2], o[3]});
      }
    pend[buf].acc = -1;
  }
};

Weights one(const float* p) {
  Weights w;
  for (int i = 0; i < kMaxModels; ++i) w.p[i] = nullptr;
  w.p[0] = p;
  return w;
}

// persistent walk: `ctas` executors share the units with a stride, like the CUDA grid
template <class Body>
int run_pipe(const float* in, int B, const float* we, const float* be, const float* wd,
             const float* bd, float* out, int splits, int threads, int ctas) {
  ExpandDwArgs a;
  a.we = one(we); a.be = one(be); a.wd = one(wd); a.bd = one(bd);
  a.in = in; a.out = out; a.B = B; a.splits = splits; a.units = B * splits;
  for (int cta = 0; cta < ctas; ++cta) {
    HostExec x(Body::kSmemFloats, threads);
    Body::run(x, a, cta, ctas);
  }
  return 0;
}

template <class Body>
int run_block(const float* in, int B, const float* we, const float* be, const float* wd,
              const float* bd, float* out, int splits, int threads) {
  ExpandDwArgs a;
  a.we = one(we); a.be = one(be); a.wd = one(wd); a.bd = one(bd);
  a.in = in; a.out = out; a.B = B; a.splits = splits; a.units = B * splits;
  for (int cta = 0; cta < B * splits; ++cta) {
    HostExec x(Body::kSmemFloats, threads);
    Body::run(x, a, cta);
  }
  return 0;
}

}  // namespace

extern "C" {

// pipelined tensor-core bodies (host GEMM, sequential producer/consumer); cfg as below
int emu_expand_dw_tc(int cfg, const float* in, int B, const float* we, const float* be,
                     const float* wd, const float* bd, float* out, int splits, int threads, int ctas) {
  switch (cfg) {
    case 2: return run_pipe<ExpandDwPipeBody<16, 96, 2, 50, 1, 10>>(in, B, we, be, wd, bd, out, splits, threads, ctas);
    case 3: return run_pipe<ExpandDwPipeBody<24, 144, 1, 25, 2, 7>>(in, B, we, be, wd, bd, out, splits, threads, ctas);
    case 4: return run_pipe<ExpandDwPipeBody<24, 144, 2, 25, 1, 7>>(in, B, we, be, wd, bd, out, splits, threads, ctas);
  }
  return 1;
}

// cfg: 2, 3, 4 = the block of features.<cfg> (the shapes fused.cu instantiates)
int emu_expand_dw(int cfg, const float* in, int B, const float* we, const float* be,
                  const float* wd, const float* bd, float* out, int splits, int threads) {
  switch (cfg) {
    case 2: return run_block<ExpandDwBody<16, 96, 2, 50, 1, 10, 10>>(in, B, we, be, wd, bd, out, splits, threads);
    case 3: return run_block<ExpandDwBody<24, 144, 1, 25, 2, 8, 7>>(in, B, we, be, wd, bd, out, splits, threads);
    case 4: return run_block<ExpandDwBody<24, 144, 2, 25, 1, 8, 7>>(in, B, we, be, wd, bd, out, splits, threads);
  }
  return 1;
}

int emu_dw_project(const float* in, int B, const float* wd, const float* bd, const float* wp,
                   const float* bp, float* out, int threads) {
  DwProjectArgs a;
  a.wd = one(wd); a.bd = one(bd); a.wp = one(wp); a.bp = one(bp);
  a.in = in; a.out = out; a.B = B;
  for (int cta = 0; cta < B * DwProjectBody::PAIRS; ++cta) {
    HostExec x(DwProjectBody::kSmemFloats, threads);
    DwProjectBody::run(x, a, cta);
  }
  return 0;
}

int emu_front(const float* vis, int B, int C, const float* ws, const float* bs, const float* wd,
              const float* bd, const float* wp, const float* bp, float* out, int splits, int threads) {
  FrontArgs a;
  a.ws = one(ws); a.bs = one(bs); a.wd = one(wd); a.bd = one(bd); a.wp = one(wp); a.bp = one(bp);
  a.vis = vis; a.out = out; a.B = B; a.C = C; a.splits = splits;
  for (int cta = 0; cta < B * splits; ++cta) {
    HostExec x(FrontBody::smem_floats(C), threads);
    FrontBody::run(x, a, cta);
  }
  return 0;
}

}  // extern "C"
