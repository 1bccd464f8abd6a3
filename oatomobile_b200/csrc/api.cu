// C-ABI entry points (include/oat_b200.h) and weight packing.
//
// `oat_model_create` consumes the reference `state_dict` verbatim (key names of
// oatomobile/baselines/torch/dim/model.py:53-68 + torchvision MobileNetV2),
// folds each eval-mode BatchNorm into the convolution in front of it (in double
// precision), transposes every matrix into the K-major layout the kernels read
// and uploads one arena per model.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace oat {

static thread_local std::string g_error;
int64_t g_launch_count = 0;
#ifndef OAT_FLOW_DEFAULT_IMPL
#define OAT_FLOW_DEFAULT_IMPL 2
#endif
int g_flow_impl = OAT_FLOW_DEFAULT_IMPL;

bool g_profile_on = false;
namespace {
struct ProfileMark {
  const char* tag;
  cudaEvent_t ev;
};
std::vector<ProfileMark> g_profile_marks;
std::mutex g_profile_mutex;
}  // namespace
void profile_mark(const char* tag, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, stream);
  g_profile_marks.push_back({tag, ev});
}

void set_error(const std::string& msg) { g_error = msg; }
int fail(const std::string& msg) {
  g_error = msg;
  return 1;
}

namespace {

struct HostTensor {
  const float* data;
  std::vector<int64_t> shape;
};

struct Packer {
  std::map<std::string, HostTensor> sd;
  std::vector<float> arena;
  std::string err;

  const HostTensor* get(const std::string& name, std::initializer_list<int64_t> shape = {}) {
    auto it = sd.find(name);
    if (it == sd.end()) {
      if (err.empty()) err = "state_dict is missing `" + name + "`";
      return nullptr;
    }
    if (shape.size() != 0) {
      std::vector<int64_t> want(shape);
      if (it->second.shape != want) {
        if (err.empty()) err = "state_dict entry `" + name + "` has an unexpected shape";
        return nullptr;
      }
    }
    return &it->second;
  }
  size_t alloc(size_t n) {  // 128-byte aligned sub-allocation
    size_t off = (arena.size() + 31) & ~size_t(31);
    arena.resize(off + n, 0.0f);
    return off;
  }
  // Folded BN: scale = gamma / sqrt(var + eps), shift = beta - mean * scale.
  bool bn(const std::string& p, int64_t n, std::vector<double>* scale, std::vector<double>* shift) {
    const HostTensor *g = get(p + ".weight", {n}), *b = get(p + ".bias", {n}),
                     *m = get(p + ".running_mean", {n}), *v = get(p + ".running_var", {n});
    if (!g || !b || !m || !v) return false;
    scale->resize(n);
    shift->resize(n);
    for (int64_t i = 0; i < n; ++i) {
      const double s = (double)g->data[i] / std::sqrt((double)v->data[i] + 1e-5);
      (*scale)[i] = s;
      (*shift)[i] = (double)b->data[i] - (double)m->data[i] * s;
    }
    return true;
  }
};

struct Off {
  size_t w, b;
};

// 1x1 conv [N,K,1,1] + BN -> w[K][N], b[N]
bool pack_pw(Packer& P, const std::string& conv, const std::string& bnp, int64_t K, int64_t N,
             Off* o) {
  const HostTensor* w = P.get(conv + ".weight", {N, K, 1, 1});
  std::vector<double> sc, sh;
  if (!w || !P.bn(bnp, N, &sc, &sh)) return false;
  o->w = P.alloc(K * N);
  o->b = P.alloc(N);
  for (int64_t n = 0; n < N; ++n) {
    for (int64_t k = 0; k < K; ++k) P.arena[o->w + k * N + n] = (float)((double)w->data[n * K + k] * sc[n]);
    P.arena[o->b + n] = (float)sh[n];
  }
  return true;
}

// depthwise 3x3 [C,1,3,3] + BN -> w[9][C], b[C]
bool pack_dw(Packer& P, const std::string& conv, const std::string& bnp, int64_t C, Off* o) {
  const HostTensor* w = P.get(conv + ".weight", {C, 1, 3, 3});
  std::vector<double> sc, sh;
  if (!w || !P.bn(bnp, C, &sc, &sh)) return false;
  o->w = P.alloc(9 * C);
  o->b = P.alloc(C);
  for (int64_t c = 0; c < C; ++c) {
    for (int t = 0; t < 9; ++t) P.arena[o->w + t * C + c] = (float)((double)w->data[c * 9 + t] * sc[c]);
    P.arena[o->b + c] = (float)sh[c];
  }
  return true;
}

// Linear [N,K] (+bias) -> w[K][N], b[N]
bool pack_linear(Packer& P, const std::string& p, int64_t K, int64_t N, Off* o) {
  const HostTensor *w = P.get(p + ".weight", {N, K}), *b = P.get(p + ".bias", {N});
  if (!w || !b) return false;
  o->w = P.alloc(K * N);
  o->b = P.alloc(N);
  for (int64_t n = 0; n < N; ++n) {
    for (int64_t k = 0; k < K; ++k) P.arena[o->w + k * N + n] = w->data[n * K + k];
    P.arena[o->b + n] = b->data[n];
  }
  return true;
}

// (expand t, out c, repeats n, first stride s) — Sandler et al. 2018, table 2.
const int kSetting[7][4] = {{1, 16, 1, 1}, {6, 24, 2, 2},  {6, 32, 3, 2}, {6, 64, 4, 2},
                            {6, 96, 3, 1}, {6, 160, 3, 2}, {6, 320, 1, 1}};

}  // namespace
}  // namespace oat

using namespace oat;

extern "C" {

const char* oat_last_error(void) { return g_error.c_str(); }
int oat_abi_version(void) { return OAT_ABI_VERSION; }
int64_t oat_launch_count(void) { return g_launch_count; }

int oat_profile_begin(void* stream) {
  if (g_profile_on) return fail("oat_profile_begin: a profile is already open");
  {
    std::lock_guard<std::mutex> lock(g_profile_mutex);
    g_profile_marks.clear();
  }
  g_profile_on = true;
  profile_mark("", (cudaStream_t)stream);  // origin of the first interval
  return 0;
}

int oat_profile_end(char* json, int64_t capacity) {
  if (!g_profile_on) return fail("oat_profile_end: no profile is open");
  g_profile_on = false;
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  std::map<std::string, std::pair<double, int64_t>> acc;  // tag -> (ms, launches)
  cudaError_t err = cudaSuccess;
  if (!g_profile_marks.empty()) err = cudaEventSynchronize(g_profile_marks.back().ev);
  for (size_t i = 1; i < g_profile_marks.size() && err == cudaSuccess; ++i) {
    float ms = 0.f;
    err = cudaEventElapsedTime(&ms, g_profile_marks[i - 1].ev, g_profile_marks[i].ev);
    auto& a = acc[g_profile_marks[i].tag];
    a.first += ms;
    a.second += 1;
  }
  for (auto& m : g_profile_marks) cudaEventDestroy(m.ev);
  g_profile_marks.clear();
  if (err != cudaSuccess) return fail(std::string("oat_profile_end: ") + cudaGetErrorString(err));
  std::string out = "{";
  for (auto it = acc.begin(); it != acc.end(); ++it) {
    if (it != acc.begin()) out += ", ";
    char buf[512];
    std::string tag = it->first;  // training functors: keep the type name of "[with F = ...]"
    const size_t w = tag.find("F = ");
    if (w != std::string::npos) {
      tag = tag.substr(w + 4);
      const size_t e = tag.find_first_of(";]");
      if (e != std::string::npos) tag = tag.substr(0, e);
      const size_t c = tag.rfind("::");
      if (c != std::string::npos) tag = tag.substr(c + 2);
    }
    for (auto& ch : tag) if (ch == '"' || ch == '\\') ch = '_';
    if (tag.size() > 200) tag.resize(200);
    snprintf(buf, sizeof(buf), "\"%s\": {\"ms\": %.6f, \"launches\": %lld}", tag.c_str(),
             it->second.first, (long long)it->second.second);
    out += buf;
  }
  out += "}";
  if (!json || capacity < (int64_t)out.size() + 1) return fail("oat_profile_end: buffer too small");
  memcpy(json, out.c_str(), out.size() + 1);
  return 0;
}

int oat_model_create(const OatTensor* tensors, int32_t num_tensors, int32_t kind, int32_t device,
                     OatModel** out) {
  if (!tensors || !out) return fail("oat_model_create: null argument");
  if (kind != OAT_KIND_DIM && kind != OAT_KIND_CIL && kind != OAT_KIND_FLOW && kind != OAT_KIND_ENCODER)
    return fail("oat_model_create: bad kind");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("oat_model_create: no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("oat_model_create: bad device index");

  Packer P;
  for (int i = 0; i < num_tensors; ++i) {
    const OatTensor& t = tensors[i];
    if (!t.name || !t.h_data) continue;
    HostTensor h;
    h.data = static_cast<const float*>(t.h_data);
    for (int d = 0; d < t.ndim && d < 4; ++d) h.shape.push_back(t.shape[d]);
    P.sd[t.name] = h;
  }
  // a stand-alone `MobileNetV2` module (perception.py:25-55) carries its keys without the
  // `_encoder.` prefix and has neither merger nor decoder
  const std::string enc = kind == OAT_KIND_ENCODER ? "_model." : "_encoder._model.";
  const std::string f = enc + "features.";

  // ---- stem: features.0 = conv3x3 s2 (C->32) + BN + ReLU6 (perception.py:43-51)
  const bool has_encoder = kind != OAT_KIND_FLOW;
  auto stem_it = P.sd.find(f + "0.0.weight");
  if (has_encoder && (stem_it == P.sd.end() || stem_it->second.shape.size() != 4))
    return fail("state_dict is missing `" + f + "0.0.weight`");
  const int64_t C = has_encoder ? stem_it->second.shape[1] : 0;
  if (has_encoder && (stem_it->second.shape != std::vector<int64_t>({32, C, 3, 3}) || C < 1 || C > 8))
    return fail("stem conv must be [32,C,3,3] with 1 <= C <= 8");
  Off stem{0, 0};
  if (has_encoder) {
    std::vector<double> sc, sh;
    if (!P.bn(f + "0.1", 32, &sc, &sh)) return fail(P.err);
    const float* w = stem_it->second.data;
    stem.w = P.alloc(9 * C * 32);
    stem.b = P.alloc(32);
    for (int64_t co = 0; co < 32; ++co) {
      for (int64_t c = 0; c < C; ++c)
        for (int t = 0; t < 9; ++t)
          P.arena[stem.w + (t * C + c) * 32 + co] = (float)((double)w[(co * C + c) * 9 + t] * sc[co]);
      P.arena[stem.b + co] = (float)sh[co];
    }
  }

  // ---- 17 inverted-residual blocks
  struct BlockOff {
    int cin, hid, cout, stride, residual, hin, hout;
    Off expand, dw, project;
  };
  std::vector<BlockOff> blocks;
  if (has_encoder) {
    int cin = 32, idx = 1, h = 50;
    for (int s = 0; s < 7; ++s) {
      for (int i = 0; i < kSetting[s][2]; ++i, ++idx) {
        BlockOff b;
        b.cin = cin;
        b.hid = cin * kSetting[s][0];
        b.cout = kSetting[s][1];
        b.stride = (i == 0) ? kSetting[s][3] : 1;
        b.residual = (b.stride == 1 && b.cin == b.cout) ? 1 : 0;
        b.hin = h;
        b.hout = (h - 1) / b.stride + 1;  // k=3, pad=1
        const std::string p = f + std::to_string(idx) + ".conv";
        bool ok = true;
        if (b.hid != b.cin) {
          ok = ok && pack_pw(P, p + ".0.0", p + ".0.1", b.cin, b.hid, &b.expand);
          ok = ok && pack_dw(P, p + ".1.0", p + ".1.1", b.hid, &b.dw);
          ok = ok && pack_pw(P, p + ".2", p + ".3", b.hid, b.cout, &b.project);
        } else {
          b.expand = Off{0, 0};
          ok = ok && pack_dw(P, p + ".0.0", p + ".0.1", b.hid, &b.dw);
          ok = ok && pack_pw(P, p + ".1", p + ".2", b.hid, b.cout, &b.project);
        }
        if (!ok) return fail(P.err);
        blocks.push_back(b);
        cin = b.cout;
        h = b.hout;
      }
    }
  }
  Off last{0, 0}, fc{0, 0}, mg[3] = {{0, 0}, {0, 0}, {0, 0}}, flow{0, 0};
  const int S = (kind == OAT_KIND_CIL) ? 6 : 5;
  if (has_encoder) {
    if (!pack_pw(P, f + "18.0", f + "18.1", 320, 1280, &last)) return fail(P.err);
    if (!pack_linear(P, enc + "classifier.1", 1280, OAT_ENC_FEATURES, &fc)) return fail(P.err);
    if (kind != OAT_KIND_ENCODER) {
      if (!pack_linear(P, "_merger._model.0", OAT_ENC_FEATURES + S, 64, &mg[0])) return fail(P.err);
      if (!pack_linear(P, "_merger._model.2", 64, 64, &mg[1])) return fail(P.err);
      if (!pack_linear(P, "_merger._model.4", 64, 64, &mg[2])) return fail(P.err);
    }
  }

  // ---- GRU + head (sequence.py:53-65) / CIL GRU + output (cil/model.py:60-66)
  if (kind != OAT_KIND_ENCODER) {
    const std::string g = (kind == OAT_KIND_DIM) ? "_decoder._decoder." : "_decoder.";  // FLOW == CIL key
    const HostTensor *wih = P.get(g + "weight_ih", {192, 2}), *whh = P.get(g + "weight_hh", {192, 64}),
                     *bih = P.get(g + "bias_ih", {192}), *bhh = P.get(g + "bias_hh", {192});
    if (!wih || !whh || !bih || !bhh) return fail(P.err);
    flow.w = P.alloc(kFlowPlanFloats);
    float* fw = nullptr;
    auto F = [&]() { return P.arena.data() + flow.w; };
    fw = F();
    for (int i = 0; i < 192 * 64; ++i) fw[kFlowWhhRaw + i] = whh->data[i];
    for (int j = 0; j < 192; ++j) {
      for (int k = 0; k < 64; ++k) fw[kFlowWhh + k * 192 + j] = whh->data[j * 64 + k];
      fw[kFlowWihT + j] = wih->data[j * 2 + 0];
      fw[kFlowWihT + 192 + j] = wih->data[j * 2 + 1];
      fw[kFlowBih + j] = bih->data[j];
      fw[kFlowBhh + j] = bhh->data[j];
    }
    if (kind != OAT_KIND_CIL) {
      const std::string h = (kind == OAT_KIND_DIM) ? "_decoder._locscale._model." : "_locscale._model.";
      const HostTensor *w1 = P.get(h + "0.weight", {32, 64}), *b1 = P.get(h + "0.bias", {32}),
                       *w2 = P.get(h + "2.weight", {4, 32}), *b2 = P.get(h + "2.bias", {4});
      if (!w1 || !b1 || !w2 || !b2)
        return fail(P.err + " (the flow head must be MLP(64,[32,4]); sequence.py:61 sizes it by T)");
      for (int i = 0; i < 32 * 64; ++i) fw[kFlowW1Raw + i] = w1->data[i];
      for (int j = 0; j < 32; ++j) {
        for (int k = 0; k < 64; ++k) fw[kFlowW1T + k * 32 + j] = w1->data[j * 64 + k];
        fw[kFlowB1 + j] = b1->data[j];
      }
      for (int i = 0; i < 128; ++i) fw[kFlowW2 + i] = w2->data[i];
      for (int i = 0; i < 4; ++i) fw[kFlowB2 + i] = b2->data[i];
    } else {
      const HostTensor *wo = P.get("_output.weight", {2, 64}), *bo = P.get("_output.bias", {2});
      if (!wo || !bo) return fail(P.err);
      for (int i = 0; i < 128; ++i) fw[kFlowW1T + i] = wo->data[i];
      fw[kFlowB2 + 0] = bo->data[0];
      fw[kFlowB2 + 1] = bo->data[1];
    }
  }

  Off flow_tc{0, 0};
  if (kind != OAT_KIND_CIL && kind != OAT_KIND_ENCODER) {
    flow_tc.w = P.alloc(kFlowTcFloats);
    pack_flow_tc_image(P.arena.data() + flow.w, P.arena.data() + flow_tc.w);
  }

  // ---- upload
  OatModel* m = new OatModel();
  m->kind = kind;
  m->device = device;
  m->in_channels = (int)C;
  m->scalars = S;
  m->arena_floats = P.arena.size();
  int prev = 0;
  cudaGetDevice(&prev);
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&m->arena, m->arena_floats * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMemcpy(m->arena, P.arena.data(), m->arena_floats * sizeof(float), cudaMemcpyHostToDevice);
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    if (m->arena) cudaFree(m->arena);
    delete m;
    return fail(std::string("oat_model_create: ") + cudaGetErrorString(e));
  }
  auto dev = [&](const Off& o) {
    ConvW c;
    c.w = m->arena + o.w;
    c.b = m->arena + o.b;
    return c;
  };
  m->stem = dev(stem);
  for (const BlockOff& b : blocks) {
    BlockW bw;
    bw.cin = b.cin; bw.hid = b.hid; bw.cout = b.cout; bw.stride = b.stride;
    bw.residual = b.residual; bw.hin = b.hin; bw.hout = b.hout;
    if (b.hid != b.cin) bw.expand = dev(b.expand);
    bw.dw = dev(b.dw);
    bw.project = dev(b.project);
    m->blocks.push_back(bw);
  }
  m->last = dev(last);
  m->fc = dev(fc);
  for (int i = 0; i < 3; ++i) m->merger[i] = dev(mg[i]);
  m->flow = kind != OAT_KIND_ENCODER ? m->arena + flow.w : nullptr;
  m->flow_tc = (kind != OAT_KIND_CIL && kind != OAT_KIND_ENCODER) ? m->arena + flow_tc.w : nullptr;
  *out = m;
  return 0;
}

int oat_model_destroy(OatModel* model) {
  if (!model) return 0;
  if (model->arena) cudaFree(model->arena);
  delete model;
  return 0;
}

int oat_model_in_channels(const OatModel* model) { return model ? model->in_channels : -1; }

int oat_ensemble_create(OatModel* const* models, int32_t num_models, OatEnsemble** out) {
  if (!models || !out || num_models < 1) return fail("oat_ensemble_create: need >= 1 model");
  if (num_models > kMaxModels) return fail("oat_ensemble_create: too many models for one GPU (max 16)");
  OatEnsemble* e = new OatEnsemble();
  for (int i = 0; i < num_models; ++i) {
    if (!models[i]) { delete e; return fail("oat_ensemble_create: null model"); }
    if (models[i]->device != models[0]->device || models[i]->kind != models[0]->kind ||
        models[i]->in_channels != models[0]->in_channels) {
      delete e;
      return fail("oat_ensemble_create: models must share device, kind and in_channels");
    }
    e->models.push_back(models[i]);
  }
  e->device = models[0]->device;
  if (models[0]->kind == OAT_KIND_FLOW) {
    // decoders only (replicas of every model's AutoregressiveFlow on a rank that shards the flow
    // stage by scenes): usable with oat_rip_sample_score, no encoder arena, no workspace
    *out = e;
    return 0;
  }
  // Default: depthwise+project of features.1 and expand+depthwise of features.2-4 fused
  // (measured 4.47 -> 4.21 ms per encode at B=256, E=4); the fused features.0+1 kernel is
  // correct but not faster than the separate launches, so bit 0 stays off.
#ifndef OAT_FUSE_DEFAULT
#define OAT_FUSE_DEFAULT 30
#endif
  e->fuse = OAT_FUSE_DEFAULT;
  if (const char* env = getenv("OAT_FUSE")) e->fuse = atoi(env) & 63;
  if (const char* env = getenv("OAT_FUSE_TC")) e->fuse_tc = atoi(env) < 0 ? 0 : (atoi(env) > 2 ? 2 : atoi(env));
  // ---- tensor-core copies of every pointwise layer: [E][N][K], TF32 hi/lo split ----
  {
    const OatModel* m0 = models[0];
    struct L { int K, N; };
    std::vector<L> layers;
    for (const BlockW& b : m0->blocks) {
      if (b.hid != b.cin) layers.push_back({b.cin, b.hid});
      layers.push_back({b.hid, b.cout});
    }
    layers.push_back({320, 1280});
    layers.push_back({1280, OAT_ENC_FEATURES});
    size_t total = 0;
    for (const L& l : layers) total += (size_t)num_models * (3 * (size_t)l.K * l.N + l.N);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    cudaError_t err = cudaMalloc(&e->tc_arena, total * sizeof(float));
    if (err != cudaSuccess) {
      cudaSetDevice(prev);
      delete e;
      return fail(std::string("oat_ensemble_create: ") + cudaGetErrorString(err));
    }
    float* p = e->tc_arena;
    int rc = 0;
    for (size_t li = 0; li < layers.size() && rc == 0; ++li) {
      TcLayer t;
      t.K = layers[li].K; t.N = layers[li].N;
      const size_t kn = (size_t)t.K * t.N;
      t.wh = p; p += kn * num_models;
      t.wl = p; p += kn * num_models;
      t.wr = p; p += kn * num_models;
      t.bias = p; p += (size_t)t.N * num_models;
      for (int mi = 0; mi < num_models && rc == 0; ++mi) {
        const OatModel* m = models[mi];
        // walk the same order as `layers`
        const ConvW* cw = nullptr;
        size_t idx = 0;
        for (const BlockW& b : m->blocks) {
          if (b.hid != b.cin) { if (idx == li) cw = &b.expand; ++idx; }
          if (idx == li) cw = &b.project;
          ++idx;
          if (cw) break;
        }
        if (!cw) cw = (li == layers.size() - 2) ? &m->last : &m->fc;
        rc = tc_pack_weights(cw->w, t.K, t.N, t.wh + kn * mi, t.wl + kn * mi, t.wr + kn * mi, 0);
        if (rc == 0 && cudaMemcpy(t.bias + (size_t)t.N * mi, cw->b, t.N * sizeof(float),
                                  cudaMemcpyDeviceToDevice) != cudaSuccess)
          rc = fail("oat_ensemble_create: bias copy failed");
      }
      e->tc.push_back(t);
    }
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) rc = fail("oat_ensemble_create: pack failed");
    cudaSetDevice(prev);
    if (rc != 0) {
      cudaFree(e->tc_arena);
      delete e;
      return rc;
    }
  }
  *out = e;
  return 0;
}

int oat_ensemble_set_pw_impl(OatEnsemble* ens, int32_t impl) {
  if (!ens) return fail("oat_ensemble_set_pw_impl: null ensemble");
  if (impl != 0 && impl != 1) return fail("oat_ensemble_set_pw_impl: impl must be 0 (simt) or 1 (tcgen05)");
  ens->pw_impl = impl;
  return 0;
}

int oat_ensemble_set_fusion(OatEnsemble* ens, int32_t mask) {
  if (!ens) return fail("oat_ensemble_set_fusion: null ensemble");
  if (mask < 0 || mask > 63) return fail("oat_ensemble_set_fusion: mask must be in [0, 63]");
  ens->fuse = mask;
  return 0;
}

int oat_ensemble_set_fusion_tc(OatEnsemble* ens, int32_t mode) {
  if (!ens) return fail("oat_ensemble_set_fusion_tc: null ensemble");
  if (mode < 0 || mode > 2) return fail("oat_ensemble_set_fusion_tc: mode must be 0, 1 or 2");
  ens->fuse_tc = mode;
  return 0;
}

int oat_ensemble_get_fusion(const OatEnsemble* ens) { return ens ? ens->fuse : -1; }

int oat_debug_tc_gemm(const float* A, const float* W, const float* bias, const float* R, float* C,
                      int32_t M, int32_t K, int32_t N, int32_t E, int32_t relu6, void* stream) {
  if (!A || !W || !bias || !C) return fail("oat_debug_tc_gemm: null argument");
  const int64_t n = (int64_t)E * N * K;
  float* tmp = nullptr;
  OAT_CUDA(cudaMalloc(&tmp, 2 * n * sizeof(float)));
  int rc = tc_split_weights(W, tmp, tmp + n, n, (cudaStream_t)stream);
  if (rc == 0) {
    TcGemmProblem p;
    p.A = A; p.Wh = tmp; p.Wl = tmp + n; p.bias = bias; p.R = R; p.C = C;
    p.M = M; p.K = K; p.N = N; p.E = E; p.relu6 = relu6;
    rc = tc_pw_gemm(p, (cudaStream_t)stream);
  }
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(tmp);
  if (rc == 0 && e != cudaSuccess) return fail(std::string("oat_debug_tc_gemm: ") + cudaGetErrorString(e));
  return rc;
}

int oat_ensemble_destroy(OatEnsemble* ens) {
  if (!ens) return 0;
  if (ens->ws) cudaFree(ens->ws);
  if (ens->tc_arena) cudaFree(ens->tc_arena);
  delete ens;
  return 0;
}

int oat_ensemble_reserve(OatEnsemble* ens, int32_t batch) {
  if (!ens) return fail("oat_ensemble_reserve: null ensemble");
  if (batch <= ens->reserved_batch) return 0;
  // per (model, image) floats: block in/out ping-pong, expanded, depthwise, pooled, feat
  const size_t kA = 50 * 50 * 32, kH1 = 50 * 50 * 96, kH2 = 25 * 25 * 144;
  const size_t per = 2 * kA + kH1 + kH2 + 1280 + 128;
  const size_t EB = ens->models.size() * (size_t)batch;
  int prev = 0;
  cudaGetDevice(&prev);
  OAT_CUDA(cudaSetDevice(ens->device));
  if (ens->ws) {
    OAT_CUDA(cudaDeviceSynchronize());
    OAT_CUDA(cudaFree(ens->ws));
    ens->ws = nullptr;
  }
  cudaError_t e = cudaMalloc(&ens->ws, per * EB * sizeof(float));
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    ens->reserved_batch = 0;
    return fail(std::string("oat_ensemble_reserve: ") + cudaGetErrorString(e));
  }
  float* p = ens->ws;
  ens->bufA = p; p += kA * EB;
  ens->bufB = p; p += kA * EB;
  ens->bufH1 = p; p += kH1 * EB;
  ens->bufH2 = p; p += kH2 * EB;
  ens->pooled = p; p += 1280 * EB;
  ens->feat = p;
  ens->reserved_batch = batch;
  return 0;
}

static int check_device(int want, const char* who) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess) return fail(std::string(who) + ": no CUDA device");
  if (cur != want)
    return fail(std::string(who) + ": current CUDA device differs from the model's device");
  return 0;
}

int oat_transform_visual(const float* lidar, int32_t B, int32_t C, int32_t H, int32_t W,
                         float* visual, void* stream) {
  if (B <= 0 || C <= 0) return 0;
  if (!lidar || !visual) return fail("oat_transform_visual: null pointer");
  if (g_profile_on) profile_mark("(host gap before call)", (cudaStream_t)stream);
  if (H < 2 || W < 2) return fail("oat_transform_visual: input must be at least 2x2");
  return launch_transform_visual(lidar, B, C, H, W, visual, (cudaStream_t)stream);
}

int oat_transform_visual_hwc(const float* lidar, int32_t B, int32_t H, int32_t W, int32_t C,
                             float* visual, void* stream) {
  if (B <= 0 || C <= 0) return 0;
  if (!lidar || !visual) return fail("oat_transform_visual_hwc: null pointer");
  if (H < 2 || W < 2) return fail("oat_transform_visual_hwc: input must be at least 2x2");
  return launch_transform_visual(lidar, B, C, H, W, visual, (cudaStream_t)stream, true);
}

int oat_encode(OatEnsemble* ens, const float* visual, const float* scalars, int32_t B, float* z,
               void* stream) {
  if (B <= 0) return 0;
  if (!ens || !visual || !scalars || !z) return fail("oat_encode: null argument");
  if (ens->models[0]->kind == OAT_KIND_FLOW) return fail("oat_encode: a decoder-only ensemble has no encoder");
  if (g_profile_on) profile_mark("(host gap before call)", (cudaStream_t)stream);
  if (int rc = check_device(ens->device, "oat_encode")) return rc;
  if (int rc = oat_ensemble_reserve(ens, B)) return rc;
  return encoder_forward(ens, visual, scalars, B, z, (cudaStream_t)stream);
}

int oat_encode_features(OatEnsemble* ens, const float* visual, int32_t B, float* features,
                        void* stream) {
  if (B <= 0) return 0;
  if (!ens || !visual || !features) return fail("oat_encode_features: null argument");
  if (ens->models[0]->kind == OAT_KIND_FLOW) return fail("oat_encode_features: a decoder-only ensemble has no encoder");
  if (int rc = check_device(ens->device, "oat_encode_features")) return rc;
  if (int rc = oat_ensemble_reserve(ens, B)) return rc;
  // stop_after_blocks = 18: the whole MobileNetV2 (features + pool + classifier), no merger
  return encoder_forward(ens, visual, nullptr, B, nullptr, (cudaStream_t)stream, 18, features);
}

int oat_mlp_forward(const float* const* weights, const float* const* biases, const int32_t* sizes,
                    int32_t num_layers, int32_t activate_final, const float* x, int32_t B, float* out,
                    void* stream) {
  if (B <= 0) return 0;
  if (!weights || !sizes || !x || !out) return fail("oat_mlp_forward: null argument");
  return launch_mlp(weights, biases, sizes, num_layers, activate_final, x, B, out, (cudaStream_t)stream);
}

int oat_debug_encoder_prefix(OatEnsemble* ens, const float* visual, int32_t B, int32_t blocks,
                             float* out, void* stream) {
  if (B <= 0) return 0;
  if (!ens || !visual || !out) return fail("oat_debug_encoder_prefix: null argument");
  if (blocks < 0 || blocks > 17) return fail("oat_debug_encoder_prefix: blocks must be in [0, 17]");
  if (blocks == 0 && (ens->fuse & 1))
    return fail("oat_debug_encoder_prefix: the fused front has no stem-only output");
  if (int rc = check_device(ens->device, "oat_debug_encoder_prefix")) return rc;
  if (int rc = oat_ensemble_reserve(ens, B)) return rc;
  return encoder_forward(ens, visual, nullptr, B, nullptr, (cudaStream_t)stream, blocks, out);
}

static PtrTable one_model(const OatModel* m) {
  PtrTable t;
  for (int i = 0; i < kMaxModels; ++i) t.p[i] = nullptr;
  t.p[0] = (g_flow_impl >= 1 && m->flow_tc) ? m->flow_tc : m->flow;
  return t;
}

static int dispatch_flow(const FlowLaunch& a, cudaStream_t stream) {
  // The tensor-core kernels keep the whole [rows, 2T] x/y tile in shared memory next to the
  // 118 KB weight image and the state tile: two-tile form up to T = 16, one-tile form up to
  // T = 40; longer horizons (CIL uses up to 40, anything beyond is exotic) take the SIMT kernel.
  if (g_flow_impl >= 1 && a.mode != 2) {
    if (g_flow_impl == 2 && a.T <= 16) return launch_flow_tc2(a, stream);
    if (a.T <= 40) return launch_flow_tc(a, stream);
  }
  return launch_flow(a, stream);
}

int oat_set_flow_impl(int32_t impl) {
  if (impl < 0 || impl > 2)
    return fail("oat_set_flow_impl: impl must be 0 (simt), 1 (tcgen05) or 2 (tcgen05, two tiles/CTA)");
  g_flow_impl = impl;
  return 0;
}

int oat_flow_forward(const OatModel* model, const float* x, const float* z, int64_t N, int32_t T,
                     int32_t rows_per_z, float* y, float* logabsdet, void* stream) {
  if (N <= 0) return 0;
  if (!model || !x || !z || !y) return fail("oat_flow_forward: null argument");
  if (g_profile_on) profile_mark("(host gap before call)", (cudaStream_t)stream);
  if (model->kind == OAT_KIND_CIL) return fail("oat_flow_forward: not a flow model");
  if (int rc = check_device(model->device, "oat_flow_forward")) return rc;
  FlowLaunch a{};
  a.mode = 0; a.num_models = 1; a.weights = one_model(model);
  a.in = x; a.z = z; a.z_model_stride = 0; a.out = y;
  a.logprob = nullptr; a.logabsdet = logabsdet; a.q = nullptr; a.out_model_stride = 0;
  a.goal = nullptr; a.G = 0; a.epsilon = 1.0f; a.N = N; a.T = T; a.rows_per_z = rows_per_z;
  a.skip_model = -1;
  return dispatch_flow(a, (cudaStream_t)stream);
}

int oat_flow_inverse(const OatModel* model, const float* y, const float* z, int64_t N, int32_t T,
                     int32_t rows_per_z, float* x, float* log_prob, float* logabsdet,
                     void* stream) {
  if (N <= 0) return 0;
  if (!model || !y || !z) return fail("oat_flow_inverse: null argument");
  if (g_profile_on) profile_mark("(host gap before call)", (cudaStream_t)stream);
  if (model->kind == OAT_KIND_CIL) return fail("oat_flow_inverse: not a flow model");
  if (int rc = check_device(model->device, "oat_flow_inverse")) return rc;
  FlowLaunch a{};
  a.mode = 1; a.num_models = 1; a.weights = one_model(model);
  a.in = y; a.z = z; a.z_model_stride = 0; a.out = x;
  a.logprob = log_prob; a.logabsdet = logabsdet; a.q = nullptr; a.out_model_stride = 0;
  a.goal = nullptr; a.G = 0; a.epsilon = 1.0f; a.N = N; a.T = T; a.rows_per_z = rows_per_z;
  a.skip_model = -1;
  return dispatch_flow(a, (cudaStream_t)stream);
}

int oat_rip_sample_score(OatEnsemble* ens, int32_t proposal_idx, const float* z, const float* x,
                         const float* goal, int32_t G, float epsilon, int32_t B, int32_t K,
                         int32_t T, float* y, float* q, void* stream) {
  if ((int64_t)B * K <= 0) return 0;
  if (!ens || !z || !y || !q) return fail("oat_rip_sample_score: null argument");
  if (g_profile_on) profile_mark("(host gap before call)", (cudaStream_t)stream);
  const int E = (int)ens->models.size();
  if (proposal_idx >= E) return fail("oat_rip_sample_score: proposal_idx out of range");
  if (proposal_idx >= 0 && !x) return fail("oat_rip_sample_score: x is required to sample");
  if (goal && G < 1) return fail("oat_rip_sample_score: G must be >= 1 when goal is given");
  if (epsilon <= 0.0f) return fail("oat_rip_sample_score: epsilon must be positive");
  if (ens->models[0]->kind != OAT_KIND_DIM && ens->models[0]->kind != OAT_KIND_FLOW)
    return fail("oat_rip_sample_score: not ImitativeModels / AutoregressiveFlows");
  if (int rc = check_device(ens->device, "oat_rip_sample_score")) return rc;
  const int64_t N = (int64_t)B * K;
  if (N <= 0) return 0;
  FlowLaunch a{};
  a.goal = goal; a.G = G; a.epsilon = epsilon; a.N = N; a.T = T; a.rows_per_z = K;
  a.logprob = nullptr; a.logabsdet = nullptr;
  if (proposal_idx >= 0) {
    // proposals y = f_p(x; z_p) (rip/agent.py:106); the same pass also emits q[p].
    a.mode = 0; a.num_models = 1; a.weights = one_model(ens->models[proposal_idx]);
    a.in = x; a.z = z + (int64_t)proposal_idx * B * kHidden; a.z_model_stride = 0;
    a.out = y; a.q = q + (int64_t)proposal_idx * N; a.out_model_stride = 0; a.skip_model = -1;
    if (int rc = dispatch_flow(a, (cudaStream_t)stream)) return rc;
    if (E == 1) return 0;
  }
  // scores under every (other) local model (rip/agent.py:109-119)
  a.mode = 1; a.num_models = E;
  for (int i = 0; i < kMaxModels; ++i)
    a.weights.p[i] = i < E ? (g_flow_impl >= 1 ? ens->models[i]->flow_tc : ens->models[i]->flow) : nullptr;
  a.in = y; a.z = z; a.z_model_stride = (int64_t)B * kHidden;
  a.out = nullptr; a.q = q; a.out_model_stride = N; a.skip_model = proposal_idx;
  return dispatch_flow(a, (cudaStream_t)stream);
}

int oat_rip_aggregate(const float* q, int32_t E, int32_t B, int32_t K, int32_t algo,
                      const float* y, int32_t T, float* s, int32_t* kstar, float* sbest,
                      float* plan, void* stream) {
  if (B <= 0) return 0;
  if (!q || !kstar) return fail("oat_rip_aggregate: null argument");
  if (g_profile_on) profile_mark("(host gap before call)", (cudaStream_t)stream);
  if (plan && !y) return fail("oat_rip_aggregate: plan requested without y");
  return launch_aggregate(q, E, B, K, algo, y, T, s, kstar, sbest, plan, (cudaStream_t)stream);
}

int oat_plan(OatModel* const* models, int32_t num_models, int32_t algo, const float* z,
             const float* goal, int32_t G, float epsilon, int32_t B, int32_t T, int32_t num_steps,
             float lr, float* x, float* x_best, float* plan, float* workspace,
             int64_t workspace_floats, float* loss_out, void* stream) {
  if (B <= 0 || T <= 0) return 0;
  if (!models || !z || !x || !x_best || !plan || !workspace) return fail("oat_plan: null argument");
  if (num_models < 1 || num_models > kMaxModels) return fail("oat_plan: bad ensemble size");
  if (algo < -1 || algo > OAT_ALGO_MA) return fail("oat_plan: unknown algorithm");
  if (algo == -1 && num_models != 1) return fail("oat_plan: single-model planning takes one model");
  if (goal && G < 1) return fail("oat_plan: G must be >= 1 when goal is given");
  if (epsilon <= 0.0f) return fail("oat_plan: epsilon must be positive");
  if (num_steps < 0) return fail("oat_plan: num_steps must be >= 0");
  if (workspace_floats < oat_plan_workspace_floats(B, num_models, T))
    return fail("oat_plan: workspace too small (see oat_plan_workspace_floats)");
  PlanLaunch p;
  for (int i = 0; i < kMaxModels; ++i) p.w.p[i] = nullptr;
  for (int i = 0; i < num_models; ++i) {
    if (!models[i] || models[i]->kind == OAT_KIND_CIL) return fail("oat_plan: models must be flows");
    if (int rc = check_device(models[i]->device, "oat_plan")) return rc;
    p.w.p[i] = models[i]->flow;
  }
  p.E = num_models; p.algo = algo; p.z = z; p.goal = goal; p.G = G; p.epsilon = epsilon;
  p.B = B; p.T = T; p.num_steps = num_steps; p.lr = lr;
  // workspace: [sync (4 floats)] [post 2*E*B] [adam 4*B*T] [scratch]
  cudaStream_t st = (cudaStream_t)stream;
  const size_t head = 4 + (size_t)2 * num_models * B + (size_t)4 * B * T;
  OAT_CUDA(cudaMemsetAsync(workspace, 0, head * sizeof(float), st));
  p.sync = reinterpret_cast<unsigned int*>(workspace);
  p.post = workspace + 4;
  p.adam = p.post + (size_t)2 * num_models * B;
  p.scratch = p.adam + (size_t)4 * B * T;
  p.x = x; p.x_best = x_best; p.plan = plan; p.loss_out = loss_out;
  OAT_CUDA(cudaMemcpyAsync(x_best, x, (size_t)B * T * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return launch_plan(p, st);
}

int64_t oat_plan_workspace_floats(int32_t B, int32_t num_models, int32_t T) {
  return (int64_t)(4 + (size_t)2 * num_models * B + (size_t)4 * B * T +
                   plan_scratch_floats(B, num_models, T));
}

int oat_lidar_bev(const float* points, int64_t num_points, int32_t pixels_per_meter,
                  int32_t hist_max_per_pixel, int32_t meters_max, uint32_t* counts, float* bev,
                  void* stream) {
  if (!counts || !bev || (num_points > 0 && !points)) return fail("oat_lidar_bev: null argument");
  if (num_points < 0) return fail("oat_lidar_bev: negative point count");
  return launch_lidar_bev(points, num_points, pixels_per_meter, hist_max_per_pixel, meters_max, counts,
                          bev, (cudaStream_t)stream);
}

int oat_goal_likelihood(const float* y_last, const float* goal, int32_t B, int32_t G, float epsilon,
                        float* rows, float* mean, void* stream) {
  if (B <= 0) return 0;
  if (!y_last || !goal || (!rows && !mean)) return fail("oat_goal_likelihood: null argument");
  if (G < 1 || epsilon <= 0.0f) return fail("oat_goal_likelihood: need G >= 1 and epsilon > 0");
  return launch_goal_likelihood(y_last, goal, B, G, epsilon, rows, mean, (cudaStream_t)stream);
}

int oat_cil_rollout(const OatModel* model, const float* z, int32_t B, int32_t T, float* y,
                    void* stream) {
  if (B <= 0 || T <= 0) return 0;
  if (!model || !z || !y) return fail("oat_cil_rollout: null argument");
  if (model->kind != OAT_KIND_CIL) return fail("oat_cil_rollout: not a BehaviouralModel");
  if (int rc = check_device(model->device, "oat_cil_rollout")) return rc;
  FlowLaunch a{};
  a.mode = 2; a.num_models = 1; a.weights = one_model(model);
  a.in = nullptr; a.z = z; a.z_model_stride = 0; a.out = y;
  a.N = B; a.T = T; a.rows_per_z = 1; a.skip_model = -1;
  return launch_flow(a, (cudaStream_t)stream);
}

}  // extern "C"
