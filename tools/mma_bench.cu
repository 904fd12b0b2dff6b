// mma_bench.cu — cycles per tcgen05.mma kind::i8 for several operand modes / shapes (timing only,
// operands are garbage). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mixdq_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace mixdq;

__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode 0: SS (A,B in smem); mode 1: TS (A in TMEM)
__global__ void bench(int mode, int M, int N, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_i8(M, N);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
    // warm
    for (int i = 0; i < 8; ++i) {
      if (mode == 0) umma_i8(tm, umma_desc_sw128(a_addr + (i & 3) * 32), umma_desc_sw128(b_addr + (i & 3) * 32), idesc, 1);
      else umma_i8_ts(tm, tm + 256 + (i & 3) * 8, umma_desc_sw128(b_addr + (i & 3) * 32), idesc, 1);
    }
    umma_commit(&bar); mbar_wait(&bar, 0);
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      if (mode == 0) umma_i8(tm, umma_desc_sw128(a_addr + (i & 3) * 32), umma_desc_sw128(b_addr + (i & 3) * 32), idesc, 1);
      else umma_i8_ts(tm, tm + 256 + (i & 3) * 8, umma_desc_sw128(b_addr + (i & 3) * 32), idesc, 1);
    }
    long long t1 = clock64();
    umma_commit(&bar); mbar_wait(&bar, 1);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 256;
  for (int mode = 0; mode < 2; ++mode)
    for (int M : {128, 64})
      for (int N : {16, 32, 64, 128, 256}) {
        if (M == 64 && N % 8) continue;
        bench<<<1, 128, 100 * 1024>>>(mode, M, N, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mode=%s M=%d N=%d: issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", mode ? "TS" : "SS", M, N,
               (double)h[0] / reps, (double)h[1] / reps, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
