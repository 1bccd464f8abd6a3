// tcgen05/TMA 3xTF32 pointwise-convolution GEMM (see tc_gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oat {

struct TcGemmProblem {
  const float* A;    // [E][M][K] activations (NHWC rows), K contiguous
  const float* Wh;   // [E][N][K] TF32-exact high parts of the folded weights
  const float* Wl;   // [E][N][K] TF32-exact low parts
  const float* Wr = nullptr;  // [E][N][K] the unsplit folded weights (optional): when given, only
                              // they are streamed from L2 and the hi/lo parts are formed in shared
                              // memory by the splitter warps (same formulas, bit-identical operands);
                              // measured: the deep-K layers are bound by L2 -> SM bytes, 36 % fewer this way
  const float* bias; // [E][N]
  const float* R;    // [E][M][N] residual or null
  float* C;          // [E][M][N]
  int M, K, N, E;
  int relu6;
  // Optional fused depthwise 3x3 epilogue (expand 1x1 -> BN -> ReLU6 -> depthwise 3x3 -> BN ->
  // ReLU6 of one inverted-residual block in ONE kernel): an M tile holds whole images (or one
  // half of a 13x13 image, with the halo rows it needs), so every 3x3 neighbourhood lies inside
  // the 32-column slab the epilogue has just staged in shared memory; the 6x expanded tensor
  // never reaches HBM.  dw_out != null selects the mode: A is [E][B*hin*hin][K] (NHWC rows),
  // C is unused, dw_out is [E][B*hout*hout][N]; N % 32 == 0.
  float* dw_out = nullptr;
  const float* dw_w[16] = {nullptr};  // per model: [9][N] folded depthwise weights
  const float* dw_b[16] = {nullptr};  // per model: [N] folded BN bias
  int B = 0, hin = 0, hout = 0, stride = 1;
};

// shapes the depthwise epilogue supports (whole images, or halves of a 13x13 image, in <= 128 rows)
bool tc_dw_epilogue_supported(int hin, int stride, int N);

int tc_pw_gemm(const TcGemmProblem& p, cudaStream_t stream);
int tc_pack_weights(const float* w_kn, int K, int N, float* hi_nk, float* lo_nk, float* raw_nk,
                    cudaStream_t stream);
int tc_split_weights(const float* w, float* hi, float* lo, int64_t n, cudaStream_t stream);

}  // namespace oat
