// Orchestration of one training step (forward in training mode, backward, loss) over the
// work-item bodies of train_functors.h.  Templated on a backend that provides memory and
// "run functor F over n work items": train.cu instantiates it with grid-stride CUDA
// kernels on a stream (the product); tests/emu instantiates it with a host loop.
//
// Reference: dim/train.py:175-213 (`train_step`: z = model._params(batch); NLL of the
// perturbed targets under the flow; backward; Adam) and cil/train.py:168-190 (L1 on the
// roll-out).  `model.train()` semantics: BatchNorm uses batch statistics and updates its
// running estimates, the classifier Dropout mask is supplied by the caller.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "train_functors.h"

namespace oat {
namespace train {

struct TensorRef {
  float* p = nullptr;  // parameter / buffer storage
  float* g = nullptr;  // gradient storage (null for buffers)
  std::vector<int64_t> shape;
};

enum UnitType { kStem = 0, kPointwise = 1, kDepthwise = 2 };

struct Unit {  // convolution + BatchNorm (+ ReLU6) (+ residual add)
  int type = kPointwise;
  int cin = 0, cout = 0, stride = 1, hin = 0, hout = 0, relu6 = 1;
  int skip_from = -1;       // unit whose output is added after the BN (-1: none)
  int input_has_skip = 0;   // the input's gradient buffer already holds the skip gradient
  TensorRef w, gamma, beta, rmean, rvar;
  float *R = nullptr, *A = nullptr, *G = nullptr;          // [B*hout*hout][cout]
  float *mean = nullptr, *invstd = nullptr, *mdp = nullptr, *mdpxh = nullptr;  // [cout]
  double* acc = nullptr;                                   // [2*cout]
};

// (expand t, out c, repeats n, first stride s) — Sandler et al. 2018, table 2.
static const int kMbv2[7][4] = {{1, 16, 1, 1}, {6, 24, 2, 2},  {6, 32, 3, 2}, {6, 64, 4, 2},
                                {6, 96, 3, 1}, {6, 160, 3, 2}, {6, 320, 1, 1}};

template <class BK>
struct TrainerT {
  BK bk;
  int kind = 0;  // 0 DIM (flow NLL), 1 CIL (L1 roll-out)
  int C = 0, S = 5, H = 100;
  std::map<std::string, TensorRef> sd;
  std::vector<Unit> units;
  TensorRef fc_w, fc_b, mg_w[3], mg_b[3];
  TensorRef wih, whh, bih, bhh, h1w, h1b, h2w, h2b;
  float* grad_flat = nullptr;
  int64_t grad_flat_floats = 0;
  std::string err;

  int cap_B = 0, cap_T = 0;
  std::vector<void*> allocs;
  float *pooled = nullptr, *gpooled = nullptr, *U = nullptr, *gU = nullptr;
  float *H1 = nullptr, *H2 = nullptr, *Z = nullptr, *gH1 = nullptr, *gH2 = nullptr, *gZ = nullptr;
  float* scratch = nullptr;
  float* wrep = nullptr;  // kReplicas copies of the largest depthwise / stem weight gradient
  double* loss_acc = nullptr;
  int64_t work_items = 0;  // functor launches of the last step (gpu_launches accounting)

  ~TrainerT() { release(); }
  void release() {
    for (void* p : allocs) bk.free(p);
    allocs.clear();
    cap_B = cap_T = 0;
  }

  bool get(const std::string& name, std::vector<int64_t> shape, TensorRef* out, bool need_grad) {
    auto it = sd.find(name);
    if (it == sd.end()) {
      err = "missing tensor `" + name + "`";
      return false;
    }
    if (it->second.shape != shape) {
      err = "tensor `" + name + "` has an unexpected shape";
      return false;
    }
    if (!it->second.p || (need_grad && !it->second.g)) {
      err = "tensor `" + name + "` has no device storage" + (need_grad ? " for its gradient" : "");
      return false;
    }
    *out = it->second;
    return true;
  }
  bool bn(const std::string& p, int c, Unit* u) {
    return get(p + ".weight", {c}, &u->gamma, true) && get(p + ".bias", {c}, &u->beta, true) &&
           get(p + ".running_mean", {c}, &u->rmean, false) && get(p + ".running_var", {c}, &u->rvar, false);
  }

  // Builds the unit list from the reference key layout (torchvision MobileNetV2).
  bool init(int kind_) {
    kind = kind_;
    S = kind == 1 ? 6 : 5;
    const std::string f = "_encoder._model.features.";
    auto st = sd.find(f + "0.0.weight");
    if (st == sd.end() || st->second.shape.size() != 4) {
      err = "missing tensor `" + f + "0.0.weight`";
      return false;
    }
    C = (int)st->second.shape[1];
    if (C < 1 || C > 8) {
      err = "stem conv must be [32,C,3,3] with 1 <= C <= 8";
      return false;
    }
    Unit stem;
    stem.type = kStem;
    stem.cin = C; stem.cout = 32; stem.stride = 2; stem.hin = H; stem.hout = (H - 1) / 2 + 1;
    if (!get(f + "0.0.weight", {32, C, 3, 3}, &stem.w, true) || !bn(f + "0.1", 32, &stem)) return false;
    units.push_back(stem);
    int cin = 32, idx = 1, h = stem.hout;
    for (int s = 0; s < 7; ++s) {
      for (int i = 0; i < kMbv2[s][2]; ++i, ++idx) {
        const int t = kMbv2[s][0], cout = kMbv2[s][1], stride = i == 0 ? kMbv2[s][3] : 1;
        const int hid = cin * t, hout = (h - 1) / stride + 1;
        const bool residual = stride == 1 && cin == cout;
        const int block_in = (int)units.size() - 1;
        const std::string p = f + std::to_string(idx) + ".conv.";
        int j = 0;
        if (t != 1) {
          Unit e;
          e.type = kPointwise; e.cin = cin; e.cout = hid; e.hin = e.hout = h;
          e.input_has_skip = residual ? 1 : 0;
          if (!get(p + "0.0.weight", {hid, cin, 1, 1}, &e.w, true) || !bn(p + "0.1", hid, &e)) return false;
          units.push_back(e);
          j = 1;
        }
        Unit d;
        d.type = kDepthwise; d.cin = d.cout = hid; d.stride = stride; d.hin = h; d.hout = hout;
        const std::string dj = p + std::to_string(j);
        if (!get(dj + ".0.weight", {hid, 1, 3, 3}, &d.w, true) || !bn(dj + ".1", hid, &d)) return false;
        units.push_back(d);
        Unit pr;
        pr.type = kPointwise; pr.cin = hid; pr.cout = cout; pr.hin = pr.hout = hout; pr.relu6 = 0;
        pr.skip_from = residual ? block_in : -1;
        if (!get(p + std::to_string(j + 1) + ".weight", {cout, hid, 1, 1}, &pr.w, true) ||
            !bn(p + std::to_string(j + 2), cout, &pr))
          return false;
        units.push_back(pr);
        cin = cout;
        h = hout;
      }
    }
    Unit last;
    last.type = kPointwise; last.cin = cin; last.cout = 1280; last.hin = last.hout = h;
    if (!get(f + "18.0.weight", {1280, cin, 1, 1}, &last.w, true) || !bn(f + "18.1", 1280, &last)) return false;
    units.push_back(last);

    const std::string cl = "_encoder._model.classifier.1.";
    if (!get(cl + "weight", {128, 1280}, &fc_w, true) || !get(cl + "bias", {128}, &fc_b, true)) return false;
    const int in0 = 128 + S;
    const int mi[3] = {0, 2, 4};
    for (int l = 0; l < 3; ++l) {
      const std::string m = "_merger._model." + std::to_string(mi[l]) + ".";
      if (!get(m + "weight", {64, l == 0 ? in0 : 64}, &mg_w[l], true) || !get(m + "bias", {64}, &mg_b[l], true))
        return false;
    }
    const std::string g = kind == 0 ? "_decoder._decoder." : "_decoder.";
    if (!get(g + "weight_ih", {192, 2}, &wih, true) || !get(g + "weight_hh", {192, 64}, &whh, true) ||
        !get(g + "bias_ih", {192}, &bih, true) || !get(g + "bias_hh", {192}, &bhh, true))
      return false;
    if (kind == 0) {
      const std::string hd = "_decoder._locscale._model.";
      if (!get(hd + "0.weight", {32, 64}, &h1w, true) || !get(hd + "0.bias", {32}, &h1b, true) ||
          !get(hd + "2.weight", {4, 32}, &h2w, true) || !get(hd + "2.bias", {4}, &h2b, true)) {
        err += " (the flow head must be MLP(64,[32,4]); sequence.py:61 sizes it by T)";
        return false;
      }
    } else {
      if (!get("_output.weight", {2, 64}, &h1w, true) || !get("_output.bias", {2}, &h1b, true)) return false;
    }
    return true;
  }

  template <class T>
  T* alloc(size_t n) {
    void* p = bk.alloc(n * sizeof(T));
    if (p) allocs.push_back(p);
    return static_cast<T*>(p);
  }

  bool reserve(int B, int T) {
    if (B <= cap_B && T <= cap_T) return true;
    release();
    bool ok = true;
    for (Unit& u : units) {
      const size_t n = (size_t)B * u.hout * u.hout * u.cout;
      ok = ok && (u.R = alloc<float>(n)) && (u.A = alloc<float>(n)) && (u.G = alloc<float>(n));
      ok = ok && (u.mean = alloc<float>(u.cout)) && (u.invstd = alloc<float>(u.cout)) &&
           (u.mdp = alloc<float>(u.cout)) && (u.mdpxh = alloc<float>(u.cout)) &&
           (u.acc = alloc<double>(2 * (size_t)u.cout * kReplicas));
      if (!ok) break;
    }
    const int in0 = 128 + S;
    ok = ok && (pooled = alloc<float>((size_t)B * 1280)) && (gpooled = alloc<float>((size_t)B * 1280)) &&
         (U = alloc<float>((size_t)B * in0)) && (gU = alloc<float>((size_t)B * in0)) &&
         (H1 = alloc<float>((size_t)B * 64)) && (H2 = alloc<float>((size_t)B * 64)) &&
         (Z = alloc<float>((size_t)B * 64)) && (gH1 = alloc<float>((size_t)B * 64)) &&
         (gH2 = alloc<float>((size_t)B * 64)) && (gZ = alloc<float>((size_t)B * 64)) &&
         (scratch = alloc<float>((size_t)B * T * kDecRecord)) && (loss_acc = alloc<double>(1)) &&
         (wrep = alloc<float>((size_t)kReplicas * 960 * 9));
    if (!ok) {
      release();
      err = "out of device memory reserving the training workspace";
      return false;
    }
    cap_B = B;
    cap_T = T;
    return true;
  }

  // Post-activation outputs of the last step, for gradient checks that need the ReLU
  // masks the forward pass actually took: index 0..51 = encoder units ([B*h*h][cout]),
  // 52..54 = the three merger layers ([B][64]).
  int last_B = 0;
  bool activation(int index, const float** data, int64_t* rows, int* channels) const {
    const int n = (int)units.size();
    if (last_B <= 0 || index < 0 || index >= n + 3) return false;
    if (index < n) {
      const Unit& u = units[index];
      *data = u.A;
      *rows = (int64_t)last_B * u.hout * u.hout;
      *channels = u.cout;
    } else {
      const float* m[3] = {H1, H2, Z};
      *data = m[index - n];
      *rows = last_B;
      *channels = 64;
    }
    return true;
  }

  template <class F>
  void run(int64_t n, const F& f) {
    bk.run(n, f);
    ++work_items;
  }
  static int64_t chunks(int64_t M, int rows) { return (M + rows - 1) / rows; }
  // Rows per work item of a row reduction with `per_chunk` work items per chunk: aim at
  // ~150k work items in total, between `lo` and `hi` rows each.
  static int pick_rows(int64_t M, int64_t per_chunk, int lo, int hi) {
    int64_t r = (M * per_chunk + 150000 - 1) / 150000;
    if (r < lo) r = lo;
    if (r > hi) r = hi;
    return (int)r;
  }

  void batchnorm_forward(Unit& u, int64_t M) {
    const int c = u.cout;
    const int rows = pick_rows(M, c / 4, 16, 128);
    run(chunks(M, rows) * (c / 4), BnStats{u.R, u.acc, M, c, rows});
    run(c, BnFinalize{u.acc, u.mean, u.invstd, u.rmean.p, u.rvar.p, M, c});
    const float* skip = u.skip_from >= 0 ? units[u.skip_from].A : nullptr;
    run(M * (c / 4), BnApply{u.R, u.mean, u.invstd, u.gamma.p, u.beta.p, skip, u.A, c, u.relu6});
  }

  // visual [B][C][100][100], scalars [B][S], target [B][T][2], mask [B][1280] or null.
  // Writes every parameter gradient, the scalar loss, optionally z [B][64] and (CIL) the
  // predictions [B][T][2].
  bool forward_backward(const float* visual, const float* scalars, const float* target,
                        const float* mask, int B, int T, float* loss, float* z_out, float* pred_out) {
    if (!reserve(B, T)) return false;
    work_items = 0;
    last_B = B;
    if (grad_flat) bk.zero(grad_flat, (size_t)grad_flat_floats * sizeof(float));
    else
      for (auto& kv : sd)
        if (kv.second.g) {
          size_t n = 1;
          for (int64_t d : kv.second.shape) n *= (size_t)d;
          bk.zero(kv.second.g, n * sizeof(float));
        }

    // ---- encoder forward (training mode)
    for (size_t i = 0; i < units.size(); ++i) {
      Unit& u = units[i];
      const int64_t M = (int64_t)B * u.hout * u.hout;
      const float* in = i == 0 ? visual : units[i - 1].A;
      if (u.type == kStem)
        run(M * 32, StemFwd{in, u.w.p, u.R, B, C, u.hin, u.hin, u.hout, u.hout});
      else if (u.type == kPointwise) {
        // a backend may supply a shared-memory tiled GEMM (CUDA); the functor is the fallback
        if (!bk.pw_forward(in, u.w.p, u.R, M, u.cout, u.cin))
          run(chunks(M, 4) * (u.cout / 4), PwFwd{in, u.w.p, u.R, M, u.cout, u.cin});
      }
      else
        run(M * (u.cout / 4), DwFwd{in, u.w.p, u.R, B, u.hin, u.hin, u.hout, u.hout, u.cout, u.stride});
      batchnorm_forward(u, M);
    }
    Unit& last = units.back();
    const int HW = last.hout * last.hout;
    const int in0 = 128 + S;
    run((int64_t)B * (1280 / 4), PoolFwd{last.A, mask, pooled, HW, 1280});
    run((int64_t)B * 128, LinearFwd{pooled, fc_w.p, fc_b.p, U, 128, 1280, 1280, in0, 0});
    run((int64_t)B * S, CopyCols{scalars, U, S, in0, 128});
    run((int64_t)B * 64, LinearFwd{U, mg_w[0].p, mg_b[0].p, H1, 64, in0, in0, 64, 1});
    run((int64_t)B * 64, LinearFwd{H1, mg_w[1].p, mg_b[1].p, H2, 64, 64, 64, 64, 1});
    run((int64_t)B * 64, LinearFwd{H2, mg_w[2].p, mg_b[2].p, Z, 64, 64, 64, 64, 1});
    if (z_out) run((int64_t)B * 64, CopyCols{Z, z_out, 64, 64, 0});

    // ---- decoder: loss, gradient wrt z, decoder parameter gradients
    DecParams dp{wih.p, whh.p, bih.p, bhh.p, h1w.p, h1b.p, h2w.p, h2b.p,
                 wih.g, whh.g, bih.g, bhh.g, h1w.g, h1b.g, h2w.g, h2b.g};
    if (kind == 0) run(B, DimNllStep{dp, Z, target, scratch, gZ, loss_acc, B, T});
    else run(B, CilL1Step{dp, Z, target, scratch, gZ, pred_out, loss_acc, B, T});
    run(1, LossFinalize{loss_acc, loss, B});
    const int64_t records = (int64_t)B * T;
    run(192 * 64, DecGradWhh{scratch, whh.g, records});
    run(192 * 4, DecGradIh{scratch, wih.g, bih.g, bhh.g, records, kind == 0 ? kRecMisc + 6 : kRecMisc});
    if (kind == 0) run(32 * 65 + 4 * 33, DecGradHead{scratch, h1w.g, h1b.g, h2w.g, h2b.g, records, 32, 4, kRecDa1});
    else run(2 * 65, DecGradHead{scratch, h1w.g, h1b.g, nullptr, nullptr, records, 2, 0, kRecDout});

    // ---- merger + classifier backward
    run(64 * (64 + 1), LinearBwdW{gZ, Z, H2, mg_w[2].g, mg_b[2].g, B, 64, 64, 64, 64, 1});
    run((int64_t)B * 64, LinearBwdX{gZ, Z, mg_w[2].p, gH2, 64, 64, 64, 64, 1});
    run(64 * (64 + 1), LinearBwdW{gH2, H2, H1, mg_w[1].g, mg_b[1].g, B, 64, 64, 64, 64, 1});
    run((int64_t)B * 64, LinearBwdX{gH2, H2, mg_w[1].p, gH1, 64, 64, 64, 64, 1});
    run(64 * (in0 + 1), LinearBwdW{gH1, H1, U, mg_w[0].g, mg_b[0].g, B, 64, in0, 64, in0, 1});
    run((int64_t)B * in0, LinearBwdX{gH1, H1, mg_w[0].p, gU, 64, in0, 64, in0, 1});
    run(128 * (1280 + 1), LinearBwdW{gU, nullptr, pooled, fc_w.g, fc_b.g, B, 128, 1280, in0, 1280, 0});
    run((int64_t)B * 1280, LinearBwdX{gU, nullptr, fc_w.p, gpooled, 128, 1280, in0, 1280, 0});
    run((int64_t)B * HW * (1280 / 4), PoolBwd{gpooled, mask, last.G, HW, 1280});

    // ---- encoder backward
    for (int i = (int)units.size() - 1; i >= 0; --i) {
      Unit& u = units[i];
      const int64_t M = (int64_t)B * u.hout * u.hout;
      const int c = u.cout;
      const int srows = pick_rows(M, c / 4, 16, 128);
      run(chunks(M, srows) * (c / 4),
          BnBwdReduce{u.R, u.G, u.mean, u.invstd, u.gamma.p, u.beta.p, u.acc, M, c, u.relu6, srows});
      run(c, BnBwdParams{u.acc, u.gamma.g, u.beta.g, u.mdp, u.mdpxh, M, c});
      float* gskip = u.skip_from >= 0 ? units[u.skip_from].G : nullptr;
      run(M * (c / 4), BnBwdDx{u.R, u.G, u.mean, u.invstd, u.gamma.p, u.beta.p, u.mdp, u.mdpxh, gskip, c, u.relu6});
      if (u.type == kStem) {
        const int rows = pick_rows(M, 32, 16, 256);
        run(chunks(M, rows) * 32, StemBwdW{visual, u.G, wrep, B, C, u.hin, u.hin, u.hout, u.hout, rows});
        run(32 * C * 9, ReduceReplicas{wrep, u.w.g, 32 * C * 9});
        continue;
      }
      Unit& src = units[i - 1];
      if (u.type == kPointwise) {
        const int rows = pick_rows(M, (int64_t)(u.cout / 4) * (u.cin / 4), 8, 256);
        run(chunks(M, rows) * (u.cout / 4) * (u.cin / 4), PwBwdW{u.G, src.A, u.w.g, M, u.cout, u.cin, rows});
        if (!bk.pw_backward_x(u.G, u.w.p, src.G, M, u.cout, u.cin, u.input_has_skip))
          run(chunks(M, 4) * (u.cin / 4), PwBwdX{u.G, u.w.p, src.G, M, u.cout, u.cin, u.input_has_skip});
      } else {
        const int64_t Min = (int64_t)B * u.hin * u.hin;
        const int rows = pick_rows(M, c / 4, 16, 128);
        run(chunks(M, rows) * (c / 4), DwBwdW{u.G, src.A, wrep, B, u.hin, u.hin, u.hout, u.hout, c, u.stride, rows});
        run(c * 9, ReduceReplicas{wrep, u.w.g, c * 9});
        run(Min * (c / 4), DwBwdX{u.G, u.w.p, src.G, B, u.hin, u.hin, u.hout, u.hout, c, u.stride});
      }
    }
    return true;
  }
};

}  // namespace train
}  // namespace oat
