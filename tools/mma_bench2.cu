// mma_bench2.cu — is the ~97-cycle tcgen05.mma kind::i8 floor an accumulate-dependency latency, an
// issue limit, or specific to kind::i8? Variants: number of distinct accumulators cycled through,
// and the MMA kind (i8 / f8f6f4 / f16). Timing only, operands are garbage.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mixdq_b200/csrc tools/mma_bench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace mixdq;

template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else if (KIND == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

__host__ __device__ constexpr uint32_t idesc_for(int kind, int m, int n) {
  // kind 0: i8 (c=S32=2, a=b=1 signed); kind 1: f8f6f4 (c=F32=1, a=b=0 E4M3); kind 2: f16 (c=F32=1, a=b=0 F16)
  return (kind == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : (1u << 4)) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

template <int KIND>
__global__ void bench(int M, int N, int nacc, int reps, long long* out) {
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
    const uint32_t idesc = idesc_for(KIND, M, N);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
    const int acc_stride = 512 / nacc;
    uint64_t ad[4], bd[4];
    uint32_t acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ad[i] = umma_desc_sw128(a_addr + i * 32);
      bd[i] = umma_desc_sw128(b_addr + i * 32);
      acc[i] = tm + (i % nacc) * acc_stride;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) umma<KIND>(acc[i & 3], ad[i & 3], bd[i & 3], idesc);
    umma_commit(&bar); mbar_wait(&bar, 0);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < reps; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) umma<KIND>(acc[j & 3], ad[j & 3], bd[j & 3], idesc);
    }
    long long t1 = clock64();
    umma_commit(&bar); mbar_wait(&bar, 1);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int KIND>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(bench<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 256;
  for (int nacc : {1, 2, 4})
    for (int N : {32, 64, 128, 256}) {
      if (N * nacc > 512) continue;
      bench<KIND><<<1, 128, 100 * 1024>>>(128, N, nacc, reps, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2] = {0, 0};
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("kind=%s M=128 N=%d nacc=%d: issue %.1f cyc/mma, complete %.1f cyc/mma (%s)\n", name, N, nacc,
             (double)h[0] / reps, (double)h[1] / reps, cudaGetErrorString(e));
      if (e != cudaSuccess) exit(1);
    }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<0>("i8", d);
  run<1>("f8f6f4", d);
  run<2>("f16", d);
  return 0;
}
