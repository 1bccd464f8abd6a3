// Gradient-based MAP planner in ONE kernel launch (sm_100a).
//
// Replaces the Adam-on-latent loops of
//   ImitativeModel.forward   oatomobile/baselines/torch/dim/model.py:97-141
//   RIPAgent.__call__        oatomobile/baselines/torch/rip/agent.py:84-137
// which the reference runs as ~10^3 tiny autograd kernels per call.  Here every Adam
// step does: proposals y = f_0(x; z_0) -> per-model posteriors q_m = mean_b(log_prob -
// logabsdet) + goal likelihood -> WCM | BCM | MA | single-model loss -> analytic
// back-propagation through all selected flows (BPTT over the GRU, head and affine
// coupling) -> torch.optim.Adam update of x -> best-x bookkeeping (the post-step x,
// as written in the reference), and finally plan = f_0(x_best).
// One CTA per scene (grid-stride over scenes); the only cross-scene coupling — the
// batch-mean posteriors that pick the model / the loss — goes through a grid barrier
// (cooperative launch).  Latency-bound by construction (B is 1 in the agents).
#include "common.cuh"

namespace oat {
namespace {

constexpr int PTHREADS = 256;
constexpr int ACT = 368;  // floats saved per (scene, model, step) for the backward pass
// layout of one activation record
constexpr int A_HPREV = 0, A_R = 64, A_G = 128, A_N = 192, A_HN = 256, A_PRE = 320, A_O = 352,
              A_MU = 356, A_S = 358, A_XP = 360, A_Y = 362;
constexpr float kLog2Pi = 1.8378770664093453f;

struct PlanArgs {
  PtrTable w;  // per-model plan image: kFlow* layout + raw W_hh [192][64] + raw W_1 [32][64]
  int E, algo;  // algo: -1 single model (DIM), else OAT_ALGO_*
  const float* z;      // [E][B][64]
  const float* goal;   // [B][G][2] or null
  int G;
  float inv_eps2, log_norm;
  int B, T, num_steps;
  float lr;
  float* scratch;      // [B][E][T][ACT]
  float* post;         // [2][E][B]   (double-buffered across Adam steps)
  unsigned int* sync;  // grid barrier counter (zeroed by the caller)
  float* x;            // [B][T][2] in: initial latent, out: final latent
  float* x_best;       // [B][T][2]
  float* plan;         // [B][T][2]
  float* adam;         // [B][2][T][2] first/second moments (zeroed by the caller)
  float* loss_out;     // [num_steps] or null
};

__device__ __forceinline__ float sigmoid_acc(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float softplus_ref(float v) { return v > 20.0f ? v : log1pf(expf(v)); }

struct Smem {
  float h[64], hnew[64], ga[192], pre[32], a[32], o[4];
  float dh[64], dgate[192], din[192], dpre[32], dout[4], du[2];
  float u[2], yv[2];
  float gy[2 * 64];  // dL/dy_t accumulators, T <= 64
  float red[8];
};

__device__ void grid_barrier(unsigned int* counter, unsigned int nblocks, unsigned int& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = (gen + 1u) * nblocks;
    while (atomicAdd(counter, 0u) < target) {
    }
    __threadfence();
  }
  ++gen;
  __syncthreads();
}

// One GRU + head step of model `w` at hidden state S.h with input S.u; leaves h_t in
// S.h, head outputs in S.o, and (optionally) saves the activation record.
__device__ void flow_step(const float* __restrict__ w, Smem& S, float* __restrict__ rec) {
  const int j = threadIdx.x;
  if (j < 192) {
    float acc = 0.0f;
    const float* col = w + kFlowWhh + j;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) acc = fmaf(S.h[k], __ldg(col + k * 192), acc);
    const float ih = __ldg(w + kFlowBih + j) + __ldg(w + kFlowWihT + j) * S.u[0] +
                     __ldg(w + kFlowWihT + 192 + j) * S.u[1];
    const float hh = acc + __ldg(w + kFlowBhh + j);
    S.ga[j] = (j < 128) ? (ih + hh) : hh;       // r|z: full pre-activation; n: W_hn h + b_hn
    if (j >= 128) S.dgate[j - 128] = ih;        // (scratch) input part of the n gate
  }
  __syncthreads();
  if (j < 64) {
    const float r = sigmoid_acc(S.ga[j]);
    const float g = sigmoid_acc(S.ga[64 + j]);
    const float hn = S.ga[128 + j];
    const float n = tanhf(S.dgate[j] + r * hn);
    const float hp = S.h[j];
    S.hnew[j] = (1.0f - g) * n + g * hp;
    if (rec) { rec[A_HPREV + j] = hp; rec[A_R + j] = r; rec[A_G + j] = g; rec[A_N + j] = n; rec[A_HN + j] = hn; }
  }
  __syncthreads();
  if (j < 64) S.h[j] = S.hnew[j];
  if (j < 32) {
    float acc = __ldg(w + kFlowB1 + j);
#pragma unroll 8
    for (int k = 0; k < 64; ++k) acc = fmaf(S.hnew[k], __ldg(w + kFlowW1T + k * 32 + j), acc);
    S.pre[j] = acc;
    S.a[j] = fmaxf(acc, 0.0f);
    if (rec) rec[A_PRE + j] = acc;
  }
  __syncthreads();
  if (j < 4) {
    float acc = __ldg(w + kFlowB2 + j);
    for (int i = 0; i < 32; ++i) acc = fmaf(S.a[i], __ldg(w + kFlowW2 + j * 32 + i), acc);
    S.o[j] = acc;
    if (rec) rec[A_O + j] = acc;
  }
  __syncthreads();
}

// Back-propagates (dmu, dsig) of step t through head + GRU of model `w`, with the
// hidden-state gradient carried in S.dh; returns d/du (u = y_{t-1}) in S.du.
__device__ void flow_step_backward(const float* __restrict__ w, Smem& S,
                                   const float* __restrict__ rec, float dmu0, float dmu1,
                                   float ds0, float ds1) {
  const int j = threadIdx.x;
  if (j < 4) {
    const float o = rec[A_O + j];
    S.dout[j] = (j == 0) ? dmu0 : (j == 1) ? dmu1 : ((j == 2) ? ds0 : ds1) * sigmoid_acc(o);
  }
  __syncthreads();
  if (j < 32) {
    float da = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) da = fmaf(__ldg(w + kFlowW2 + i * 32 + j), S.dout[i], da);
    S.dpre[j] = rec[A_PRE + j] > 0.0f ? da : 0.0f;
  }
  __syncthreads();
  if (j < 64) {
    float acc = S.dh[j];
    const float* w1 = w + kFlowW1Raw + j;  // W_1 [32][64]
#pragma unroll 8
    for (int i = 0; i < 32; ++i) acc = fmaf(__ldg(w1 + i * 64), S.dpre[i], acc);
    // GRU element-wise backward
    const float r = rec[A_R + j], g = rec[A_G + j], n = rec[A_N + j], hn = rec[A_HN + j],
                hp = rec[A_HPREV + j];
    const float dn = acc * (1.0f - g);
    const float dg = acc * (hp - n);
    const float dan = dn * (1.0f - n * n);
    const float daz = dg * g * (1.0f - g);
    const float dar = dan * hn * r * (1.0f - r);
    S.dgate[j] = dar; S.dgate[64 + j] = daz; S.dgate[128 + j] = dan * r;  // through W_hh
    S.din[j] = dar;   S.din[64 + j] = daz;   S.din[128 + j] = dan;        // through W_ih
    S.hnew[j] = acc * g;                                                   // direct path to h_{t-1}
  }
  __syncthreads();
  if (j < 64) {
    float acc = S.hnew[j];
    const float* wr = w + kFlowWhhRaw + j;  // W_hh [192][64]
#pragma unroll 8
    for (int i = 0; i < 192; ++i) acc = fmaf(__ldg(wr + i * 64), S.dgate[i], acc);
    S.dh[j] = acc;
  } else if (j < 128) {  // two warps reduce d/du = W_ih^T din
    const int c = (j - 64) >> 5, lane = j & 31;
    float acc = 0.0f;
    for (int i = lane; i < 192; i += 32) acc = fmaf(__ldg(w + kFlowWihT + c * 192 + i), S.din[i], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) S.du[c] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(PTHREADS) plan_kernel(const __grid_constant__ PlanArgs a) {
  __shared__ Smem S;
  const int tid = threadIdx.x;
  const int E = a.E, B = a.B, T = a.T;
  unsigned int gen = 0;
  float loss_best = 1000.0f;  // dim/model.py:115, rip/agent.py:100

  for (int step = 1; step <= a.num_steps + 1; ++step) {
    const bool final_pass = step == a.num_steps + 1;  // plan = f_0(x_best)
    float* post = a.post + (size_t)(step & 1) * E * B;
    // ================= forward: proposals + per-model posteriors =================
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      const float* xin = (final_pass ? a.x_best : a.x) + (size_t)b * T * 2;
      const float* goal = a.goal ? a.goal + (size_t)b * a.G * 2 : nullptr;
      for (int m = 0; m < (final_pass ? 1 : E); ++m) {
        const float* w = a.w.p[m];
        if (tid < 64) S.h[tid] = __ldg(a.z + ((size_t)m * B + b) * 64 + tid);
        if (tid < 2) S.u[tid] = 0.0f;
        float sumsq = 0.0f, sumlog = 0.0f;  // meaningful on tid < 2
        __syncthreads();
        for (int t = 0; t < T; ++t) {
          float* rec = a.scratch + (((size_t)b * E + m) * T + t) * ACT;
          flow_step(w, S, final_pass ? nullptr : rec);
          if (tid < 2) {
            const float mu = S.u[tid] + S.o[tid];
            const float s = softplus_ref(S.o[2 + tid]) + 1e-3f;
            float y;
            if (m == 0) {
              y = mu + s * xin[2 * t + tid];                 // sequence.py:136
              a.scratch[(((size_t)b * E) * T + t) * ACT + A_Y + tid] = y;
              if (final_pass) a.plan[((size_t)b * T + t) * 2 + tid] = y;
            } else {
              y = a.scratch[(((size_t)b * E) * T + t) * ACT + A_Y + tid];
            }
            const float xp = (y - mu) / s;                   // sequence.py:196
            if (!final_pass) { rec[A_MU + tid] = mu; rec[A_S + tid] = s; rec[A_XP + tid] = xp; }
            sumsq = fmaf(xp, xp, sumsq);
            sumlog += logf(s);
            S.yv[tid] = y;
          }
          __syncthreads();
          if (tid < 2) S.u[tid] = S.yv[tid];                 // teacher forcing: u_{t+1} = y_t
          __syncthreads();
        }
        if (tid < 2) { S.red[tid] = sumsq; S.red[2 + tid] = sumlog; }
        __syncthreads();
        if (!final_pass && tid == 0) {
          const float sq = S.red[0] + S.red[1];
          const float sl = S.red[2] + S.red[3];
          float p = (-0.5f * sq - (float)T * kLog2Pi) - sl;
          if (goal) {  // dim/model.py:163-171, per scene
            float mx = -INFINITY;
            for (int i = 0; i < a.G; ++i) {
              const float d0 = S.yv[0] - goal[2 * i], d1 = S.yv[1] - goal[2 * i + 1];
              mx = fmaxf(mx, -0.5f * (d0 * d0 + d1 * d1) * a.inv_eps2);
            }
            float se = 0.0f;
            for (int i = 0; i < a.G; ++i) {
              const float d0 = S.yv[0] - goal[2 * i], d1 = S.yv[1] - goal[2 * i + 1];
              se += expf(-0.5f * (d0 * d0 + d1 * d1) * a.inv_eps2 - mx);
            }
            p += mx + logf(se) + a.log_norm;
          }
          post[(size_t)m * B + b] = p;
        }
        __syncthreads();
      }
    }
    if (final_pass) break;
    grid_barrier(a.sync, gridDim.x, gen);

    // ================= loss and model weights (rip/agent.py:121-127) =================
    float wsel[kMaxModels];
    float loss;
    {
      float negP[kMaxModels];
      for (int m = 0; m < E; ++m) {
        float acc = 0.0f;
        for (int b = 0; b < B; ++b) acc += post[(size_t)m * B + b];
        negP[m] = -(acc / (float)B);
        wsel[m] = 0.0f;
      }
      if (a.algo < 0) { loss = negP[0]; wsel[0] = 1.0f; }
      else if (a.algo == OAT_ALGO_MA) {
        loss = 0.0f;
        for (int m = 0; m < E; ++m) { loss += negP[m]; wsel[m] = 1.0f / (float)E; }
        loss /= (float)E;
      } else {
        int sel = 0;
        for (int m = 1; m < E; ++m)
          if (a.algo == OAT_ALGO_WCM ? (negP[m] < negP[sel]) : (negP[m] > negP[sel])) sel = m;
        loss = negP[sel];
        wsel[sel] = 1.0f;
      }
    }
    if (blockIdx.x == 0 && tid == 0 && a.loss_out) a.loss_out[step - 1] = loss;

    // ================= backward (BPTT through every selected flow) + Adam =================
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      const float* xin = a.x + (size_t)b * T * 2;
      const float* goal = a.goal ? a.goal + (size_t)b * a.G * 2 : nullptr;
      for (int i = tid; i < 2 * T; i += PTHREADS) S.gy[i] = 0.0f;
      __syncthreads();
      // goal term: d/dy_{T-1} of -sum_m w_m/B * goal_ll  (identical for every model)
      if (goal && tid < 2) {
        float wsum = 0.0f;
        for (int m = 0; m < E; ++m) wsum += wsel[m];
        const float* rec = a.scratch + (((size_t)b * E) * T + (T - 1)) * ACT;
        const float y0 = rec[A_Y], y1 = rec[A_Y + 1];
        float mx = -INFINITY;
        for (int i = 0; i < a.G; ++i) {
          const float d0 = y0 - goal[2 * i], d1 = y1 - goal[2 * i + 1];
          mx = fmaxf(mx, -0.5f * (d0 * d0 + d1 * d1) * a.inv_eps2);
        }
        float se = 0.0f, gr = 0.0f;
        for (int i = 0; i < a.G; ++i) {
          const float d0 = y0 - goal[2 * i], d1 = y1 - goal[2 * i + 1];
          const float e = expf(-0.5f * (d0 * d0 + d1 * d1) * a.inv_eps2 - mx);
          se += e;
          gr += e * (-(tid == 0 ? d0 : d1) * a.inv_eps2);
        }
        S.gy[2 * (T - 1) + tid] += -(wsum / (float)B) * (gr / se);
      }
      __syncthreads();
      // models != 0 first (they feed dL/dy), then model 0 merged with the f_0 graph
      for (int pass = 0; pass < E; ++pass) {
        const int m = (pass + 1) % E;  // order 1, 2, ..., E-1, 0
        const float kappa = -wsel[m] / (float)B;  // dL/d post_{m,b}
        if (m != 0 && kappa == 0.0f) continue;
        const float* w = a.w.p[m];
        if (tid < 64) S.dh[tid] = 0.0f;
        __syncthreads();
        for (int t = T - 1; t >= 0; --t) {
          const float* rec = a.scratch + (((size_t)b * E + m) * T + t) * ACT;
          float dmu[2], dsg[2];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float xp = rec[A_XP + c], s = rec[A_S + c];
            dmu[c] = kappa * xp / s;
            dsg[c] = kappa * (xp * xp - 1.0f) / s;
          }
          if (m == 0) {  // + the f_0 graph: y_t = mu_t + s_t x_t with total dL/dy_t known now
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const float dy = S.gy[2 * t + c] + kappa * (-rec[A_XP + c] / rec[A_S + c]);
              dmu[c] += dy;
              dsg[c] += dy * xin[2 * t + c];
            }
          }
          __syncthreads();
          if (m == 0 && tid < 2) {
            // total dL/dy_t is final here -> dL/dx_t = dL/dy_t * s_t; stash it in gy
            const float dy = S.gy[2 * t + tid] + kappa * (-rec[A_XP + tid] / rec[A_S + tid]);
            S.gy[2 * t + tid] = dy * rec[A_S + tid];
          } else if (m != 0 && tid < 2) {
            S.gy[2 * t + tid] += kappa * (-rec[A_XP + tid] / rec[A_S + tid]);  // direct d x'/d y_t
          }
          flow_step_backward(w, S, rec, dmu[0], dmu[1], dsg[0], dsg[1]);
          if (t > 0 && tid < 2) S.gy[2 * (t - 1) + tid] += S.du[tid] + dmu[tid];  // u_t = y_{t-1}
          __syncthreads();
        }
      }
      // ---- torch.optim.Adam (betas 0.9/0.999, eps 1e-8) on x, then best-x bookkeeping ----
      if (tid < 2 * T) {
        const float g = S.gy[tid];
        float* mom = a.adam + (size_t)b * 4 * T;
        const float m1 = 0.9f * mom[tid] + 0.1f * g;
        const float v1 = 0.999f * mom[2 * T + tid] + 0.001f * g * g;
        mom[tid] = m1;
        mom[2 * T + tid] = v1;
        const float bc1 = 1.0f - powf(0.9f, (float)step);
        const float bc2 = 1.0f - powf(0.999f, (float)step);
        const float denom = sqrtf(v1) / sqrtf(bc2) + 1e-8f;
        const float xn = a.x[(size_t)b * 2 * T + tid] - (a.lr / bc1) * (m1 / denom);
        a.x[(size_t)b * 2 * T + tid] = xn;
        if (loss < loss_best) a.x_best[(size_t)b * 2 * T + tid] = xn;  // post-step x, as written
      }
      __syncthreads();
    }
    if (loss < loss_best) loss_best = loss;
    grid_barrier(a.sync, gridDim.x, gen);  // x / x_best visible before the next forward
  }
}

}  // namespace

int launch_plan(const PlanLaunch& p, cudaStream_t stream) {
  if (p.B <= 0 || p.T <= 0) return 0;
  if (p.T > 64) return fail("oat_plan: T must be <= 64");
  if (p.E < 1 || p.E > kMaxModels) return fail("oat_plan: bad ensemble size");
  PlanArgs a;
  a.w = p.w; a.E = p.E; a.algo = p.algo; a.z = p.z; a.goal = p.goal; a.G = p.G;
  const double eps = p.epsilon;
  a.inv_eps2 = (float)(1.0 / (eps * eps));
  a.log_norm = p.goal ? (float)(-log(2.0 * 3.14159265358979323846 * eps * eps) - log((double)p.G)) : 0.0f;
  a.B = p.B; a.T = p.T; a.num_steps = p.num_steps; a.lr = p.lr;
  a.scratch = p.scratch; a.post = p.post; a.sync = p.sync; a.x = p.x; a.x_best = p.x_best;
  a.plan = p.plan; a.adam = p.adam; a.loss_out = p.loss_out;
  int dev = 0, sms = 0;
  OAT_CUDA(cudaGetDevice(&dev));
  OAT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.B < sms ? p.B : sms;
  void* params[] = {(void*)&a};
  OAT_CUDA(cudaLaunchCooperativeKernel((const void*)plan_kernel, dim3(grid), dim3(PTHREADS), params, 0,
                                       stream));
  g_launch_count++;
  return 0;
}

size_t plan_scratch_floats(int B, int E, int T) { return (size_t)B * E * T * ACT; }

}  // namespace oat
