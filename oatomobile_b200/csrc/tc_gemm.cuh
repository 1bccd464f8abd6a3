// tcgen05/TMA 3xTF32 pointwise-convolution GEMM (see tc_gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oat {

struct TcGemmProblem {
  const float* A;    // [E][M][K] activations (NHWC rows), K contiguous
  const float* Wh;   // [E][N][K] TF32-exact high parts of the folded weights
  const float* Wl;   // [E][N][K] TF32-exact low parts
  const float* bias; // [E][N]
  const float* R;    // [E][M][N] residual or null
  float* C;          // [E][M][N]
  int M, K, N, E;
  int relu6;
};

int tc_pw_gemm(const TcGemmProblem& p, cudaStream_t stream);
int tc_pack_weights(const float* w_kn, int K, int N, float* hi_nk, float* lo_nk,
                    cudaStream_t stream);
int tc_split_weights(const float* w, float* hi, float* lo, int64_t n, cudaStream_t stream);

}  // namespace oat
