// Shared declarations for the B200-native OATomobile RIP/DIM hot path (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/oat_b200.h"

namespace oat {

constexpr int kHidden = OAT_HIDDEN;  // 64
constexpr int kMaxModels = 16;       // models per grouped launch

// ---- error plumbing ---------------------------------------------------------
void set_error(const std::string& msg);
int fail(const std::string& msg);
extern int64_t g_launch_count;

#define OAT_CUDA(expr)                                                          \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess)                                                      \
      return ::oat::fail(std::string(#expr) + ": " + cudaGetErrorString(_e));   \
  } while (0)

#define OAT_LAUNCH_CHECK()                                                      \
  do {                                                                          \
    ::oat::g_launch_count++;                                                    \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess)                                                      \
      return ::oat::fail(std::string("kernel launch: ") + cudaGetErrorString(_e)); \
  } while (0)

// Per-kernel-family timing for bench.py's roofline (api.cu: oat_profile_begin/_end): while a
// profile is open every tagged launch is followed by a cudaEventRecord on its stream; the
// time between consecutive events is attributed to the launch in between.  No-op otherwise.
extern bool g_profile_on;
void profile_mark(const char* tag, cudaStream_t stream);
#define OAT_LAUNCHED(tag)                                                       \
  do {                                                                          \
    OAT_LAUNCH_CHECK();                                                         \
    if (::oat::g_profile_on) ::oat::profile_mark(tag, stream);                  \
  } while (0)

// ---- packed flow (GRU + head) weights, one contiguous device buffer per model
// Layout (floats), identical to the shared-memory image the flow kernel uses:
//   whhT [64][192]  W_hh transposed (k-major; columns = gate r|z|n x unit)
//   w1T  [64][32]   head layer 0 transposed      (CIL: first 128 = W_out [2][64])
//   wihT [2][192]   W_ih transposed
//   bih  [192] | bhh [192] | b1 [32] | w2 [4][32] | b2 [4]   (CIL: b2[0..1] = b_out)
constexpr int kFlowWhh = 0;
constexpr int kFlowW1T = kFlowWhh + 64 * 192;
constexpr int kFlowWihT = kFlowW1T + 64 * 32;
constexpr int kFlowBih = kFlowWihT + 2 * 192;
constexpr int kFlowBhh = kFlowBih + 192;
constexpr int kFlowB1 = kFlowBhh + 192;
constexpr int kFlowW2 = kFlowB1 + 32;
constexpr int kFlowB2 = kFlowW2 + 4 * 32;
constexpr int kFlowFloats = kFlowB2 + 4;  // 15268 (multiple of 4)
static_assert(kFlowFloats % 4 == 0, "flow weight image must be float4-copyable");
// Planner image (planner.cu) = the flow image above followed by the untransposed
// matrices the backward pass reads: W_hh [192][64], W_1 [32][64].
constexpr int kFlowWhhRaw = kFlowFloats;
constexpr int kFlowW1Raw = kFlowWhhRaw + 192 * 64;
constexpr int kFlowPlanFloats = kFlowW1Raw + 32 * 64;
// Tensor-core flow image (flow_tc.cu): W_hh / W_1 split into TF32 hi/lo parts and
// pre-swizzled into the K-major SWIZZLE_128B UMMA layout, + gate/head parameters.
constexpr int kFlowTcFloats = 29604;

// ---- encoder layer descriptors (device pointers into one weight arena) ------
struct ConvW {
  const float* w = nullptr;  // pointwise: [K][N]; depthwise: [9][C]; stem: [9*Cin][32]
  const float* b = nullptr;  // folded BN bias [N]
};

struct BlockW {  // one inverted-residual block (Sandler et al. 2018)
  int cin, hid, cout, stride, residual;
  int hin, hout;  // spatial size in / out (square)
  ConvW expand, dw, project;
};

}  // namespace oat

struct OatModel {
  int kind = 0;
  int device = 0;
  int in_channels = 2;
  int scalars = 5;  // 5 (DIM) or 6 (CIL): width of the vector inputs to the merger
  float* arena = nullptr;  // all packed weights, one cudaMalloc
  size_t arena_floats = 0;
  oat::ConvW stem;                  // [9*C][32]
  std::vector<oat::BlockW> blocks;  // 17
  oat::ConvW last;                  // 320 -> 1280
  oat::ConvW fc;                    // 1280 -> 128
  oat::ConvW merger[3];             // [in][64] transposed
  const float* flow = nullptr;      // kFlowFloats
  const float* flow_tc = nullptr;   // kFlowTcFloats (null for CIL)
};

namespace oat {
struct TcLayer {  // one pointwise layer of the whole ensemble, tensor-core layout
  float* wh = nullptr;    // [E][N][K] TF32-exact high parts
  float* wl = nullptr;    // [E][N][K] low parts
  float* wr = nullptr;    // [E][N][K] unsplit folded weights (split in shared memory by the GEMM)
  float* bias = nullptr;  // [E][N]
  int K = 0, N = 0;
};
}  // namespace oat

struct OatEnsemble {
  std::vector<OatModel*> models;
  std::vector<oat::TcLayer> tc;  // pointwise layers in execution order (+ last, fc)
  float* tc_arena = nullptr;
  int pw_impl = 1;               // 1 = tcgen05 3xTF32 (default), 0 = FP32 SIMT
  int fuse = 0;                  // bit 0: features.0+1 in one kernel; bits 1-3: expand+depthwise
                                 // of features.2-4 in one kernel; bit 4: depthwise+project of
                                 // features.1 in one kernel (ignored with bit 0) (fused.cu)
  int fuse_tc = 1;               // fused expand GEMM with pw_impl == 1: 1 auto, 2 tcgen05 always, 0 FP32
  int device = 0;
  int reserved_batch = 0;
  float* ws = nullptr;  // activation workspace
  float *bufA = nullptr, *bufB = nullptr, *bufH1 = nullptr, *bufH2 = nullptr;
  float *pooled = nullptr, *feat = nullptr;
};

namespace oat {

// Pointer tables passed by value to grouped (per-model) launches.
struct PtrTable {
  const float* p[kMaxModels];
};

// ---- kernel launchers (defined in the .cu files) -----------------------------
struct FlowLaunch {
  int mode;  // 0 sample (x->y), 1 score (y->x), 2 CIL roll-out
  int num_models;
  PtrTable weights;           // packed flow weights per model
  const float* in;            // x (mode 0) or y (mode 1); unused in mode 2
  const float* z;             // [num_models][NZ][64]
  int64_t z_model_stride;     // floats between models in z
  float* out;                 // y (mode 0/2) or x (mode 1, may be null)
  float* logprob;             // [num_models][N] or null
  float* logabsdet;           // [num_models][N] or null
  float* q;                   // [num_models][N] or null: logprob - logabsdet (+ goal)
  int64_t out_model_stride;   // floats between models in logprob/logabsdet/q
  const float* goal;          // [NZ][G][2] or null
  int G;
  float epsilon;
  int64_t N;
  int T;
  int rows_per_z;
  int skip_model;             // grid.y index that exits immediately (-1: none)
};
int launch_flow(const FlowLaunch& a, cudaStream_t stream);
int launch_flow_tc(const FlowLaunch& a, cudaStream_t stream);
int launch_flow_tc2(const FlowLaunch& a, cudaStream_t stream);  // two tiles per CTA (flow_tc2.cu)

struct PlanLaunch {
  PtrTable w;  // per-model planner images (kFlowPlanFloats)
  int E, algo;  // algo -1: single model (ImitativeModel.forward)
  const float* z;
  const float* goal;
  int G;
  float epsilon;
  int B, T, num_steps;
  float lr;
  float* scratch;
  float* post;
  unsigned int* sync;
  float* x;
  float* x_best;
  float* plan;
  float* adam;
  float* loss_out;
};
int launch_plan(const PlanLaunch& p, cudaStream_t stream);
size_t plan_scratch_floats(int B, int E, int T);
int launch_lidar_bev(const float* points, int64_t n, int pixels_per_meter, int hist_max,
                     int meters_max, unsigned int* counts, float* out, cudaStream_t stream);
int launch_goal_likelihood(const float* y_last, const float* goal, int B, int G, float epsilon,
                           float* rows, float* mean, cudaStream_t stream);  // weights = flow_tc images
void pack_flow_tc_image(const float* flow_weights, float* image);
extern int g_flow_impl;  // 0 = FP32 SIMT, 1 = tcgen05 one tile/CTA, 2 = tcgen05 two tiles/CTA

int launch_aggregate(const float* q, int E, int B, int K, int algo, const float* y, int T,
                     float* s, int32_t* kstar, float* sbest, float* plan,
                     cudaStream_t stream);

int launch_mlp(const float* const* w, const float* const* b, const int* sizes, int layers,
               int activate_final, const float* x, int B, float* out, cudaStream_t stream);
int simt_pw_gemm(const float* A, const float* W_kn, const float* bias, const float* R, float* C,
                 int M, int K, int N, int relu6, cudaStream_t stream);
int launch_transform_visual(const float* lidar, int B, int C, int H, int W, float* visual,
                            cudaStream_t stream, bool hwc = false);

// stop_after_blocks >= 0 (debug): stop after that many inverted-residual blocks (0 = after the
// stem, unless the fused front already contains block 1) and copy the activation
// [E][B][h][h][c] to prefix_out.
int encoder_forward(OatEnsemble* ens, const float* visual, const float* scalars, int B,
                    float* z, cudaStream_t stream, int stop_after_blocks = -1,
                    float* prefix_out = nullptr);

}  // namespace oat
