// simt.h — argument blocks of the portable dp4a kernels (see simt.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mixdq {

struct SimtGemmArgs {
  const int8_t* A; int64_t lda; const int8_t* W; int K;      // W: [N][K] (or packed [N][K/2] if w4)
  const int8_t* A1; int64_t lda1; const int8_t* W1; int K1;  // optional second operand pair (split)
  const float* scale; const float* bias0;                    // static: scale[N], bias0[N]; dyn: w_scale, wsum
  const float* a_scale; const float* a_zp;                   // dyn scalars or nullptr
  const float* scale1; const float* bias0_1;                 // split second half
  const __half* bias;
  __half* D; int64_t ldd;
  int M, N;
  int w4;
  int32_t* acc_out;
};

struct SimtConvArgs {
  const int8_t* x; int64_t x_cpitch; const int8_t* w;
  const float* scale; const float* wsum_krs; const float* bias0_k; const float* zp;
  const __half* bias; __half* y;
  int N, H, W, C, K, R, S, stride, pad, P, Q;
  int32_t* acc_out;
};

int simt_gemm_launch(const SimtGemmArgs& g, cudaStream_t st);
int simt_conv_launch(const SimtConvArgs& c, cudaStream_t st);

}  // namespace mixdq
