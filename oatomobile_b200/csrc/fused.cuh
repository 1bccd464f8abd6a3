// Launchers of the fused encoder-front kernels (fused.cu / fused_body.h).
#pragma once
#include "common.cuh"

namespace oat {

struct FusedFrontLaunch {  // features.0 + features.1 in one kernel
  PtrTable ws, bs;         // stem [9*C][32], [32]
  PtrTable wd, bd;         // features.1 depthwise [9][32], [32]
  PtrTable wp, bp;         // features.1 project [32][16], [16]
  const float* visual;     // [B][C][100][100], shared by the models
  float* out;              // [E][B][50][50][16]
  int E, B, C;
};
int launch_fused_front(const FusedFrontLaunch& l, cudaStream_t stream);

struct FusedDwProjectLaunch {  // features.1: depthwise 3x3 + project 32->16 in one kernel
  PtrTable wd, bd;             // depthwise [9][32], [32]
  PtrTable wp, bp;             // project [32][16], [16]
  const float* in;             // [E][B][50][50][32] (stem output)
  float* out;                  // [E][B][50][50][16]
  int E, B;
};
int launch_fused_dw_project(const FusedDwProjectLaunch& l, cudaStream_t stream);

struct FusedBlockLaunch {  // expand 1x1 + depthwise 3x3 of one inverted-residual block
  PtrTable we, be;         // expand [cin][hid], [hid]
  PtrTable wd, bd;         // depthwise [9][hid], [hid]
  const float* in;         // [E][B][hin][hin][cin]
  float* out;              // [E][B][hout][hout][hid]
  int E, B, cin, hid, stride, hin;
  int tensor_cores;        // expand GEMM: 0 FP32 FMA kernel; 1 auto (tcgen05 3xTF32 pipelined kernel
                           // where it is the faster one); 2 tcgen05 for every shape
};
bool fused_block_supported(int cin, int hid, int stride, int hin);
int launch_fused_expand_dw(const FusedBlockLaunch& l, cudaStream_t stream);

}  // namespace oat
