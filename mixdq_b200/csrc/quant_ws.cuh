// quant_ws.cuh — device workspace and grid-wide reductions shared by the dynamic-quantisation
// kernels (quant.cu: plain tensor; fused_quant.cu: LayerNorm / GEGLU / GroupNorm producers).
//
// All of these kernels run with every CTA co-resident (grid <= SM count x resident CTAs per SM), so
// a flag-based grid barrier is safe: each CTA publishes its partial result, the last one to arrive
// finishes the reduction and raises a flag, the others spin on it with acquire loads. The workspace
// is zero-initialised ONCE by the caller; every kernel leaves it zeroed again on exit. Kernels that
// share a workspace must be stream-ordered (they are: one workspace per stream on the host side).
#pragma once
#include "common.cuh"

namespace mixdq {

constexpr int kMaxPartials = 1024;   // CTAs of one launch
constexpr int kMaxStatGroups = 4096; // (image, group) pairs of one GroupNorm launch

struct DynWs {
  // ---- min/max barrier ----
  unsigned int counter;   // arrivals
  unsigned int flag;      // raised by the last arriver once scale/zp are published
  unsigned int done;      // CTAs that have consumed the flag (last one resets the workspace)
  unsigned int pad0;
  // ---- GroupNorm statistics barrier ----
  unsigned int counter2;
  unsigned int flag2;
  unsigned int done2;
  unsigned int pad1;
  float2 partial[kMaxPartials];
  // fixed-point (integer => order-independent, deterministic) sum / sum of squares per (n, group)
  unsigned long long gsum[2 * kMaxStatGroups];
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void spin_until_set(const unsigned int* flag) {
  unsigned int spins = 0;
  while (ld_acquire_u32(flag) == 0u) {
    __nanosleep(32);
    if (++spins > (1u << 24)) __trap();   // protocol bug: fail instead of hanging the device
  }
}

// qdiff asymmetric 8-bit min-max parameters (base_quantizer.py:155-190), fp32:
//   delta = max((x_max - x_min) / 255, 1e-6),  z = rint(-x_min / delta)
__device__ __forceinline__ void qdiff_params(float mn, float mx, float& delta, float& z) {
  delta = __fdiv_rn(__fsub_rn(mx, mn), 255.0f);
  if (delta < 1e-6f) delta = 1e-6f;
  z = rintf(__fdiv_rn(-mn, delta));
}

// Grid-wide min/max -> (delta, z). Called by ALL threads of every CTA (blockDim.x = NT, a multiple
// of 32, <= 1024) with the thread's partial min (<= 0) and max (>= 0). On return every thread
// holds delta and the UNSHIFTED zero point z in [0, 255]; *scale_out = delta, *zp_out = z - 128.
template <int NT>
__device__ __forceinline__ void grid_minmax_params(DynWs* __restrict__ ws, float mn, float mx,
                                                   float* __restrict__ scale_out,
                                                   float* __restrict__ zp_out, float& delta,
                                                   float& z) {
  constexpr int NW = NT / 32;
  __shared__ float smn[NW], smx[NW];
  __shared__ float s_delta, s_z;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    mn = lane < NW ? smn[lane] : 0.0f;
    mx = lane < NW ? smx[lane] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    unsigned int last = 0;
    if (lane == 0) {
      ws->partial[blockIdx.x] = make_float2(mn, mx);
      __threadfence();
      last = (atomicAdd(&ws->counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      mn = 0.0f; mx = 0.0f;
      for (int i = lane; i < static_cast<int>(gridDim.x); i += 32) {
        const float2 v = __ldcg(&ws->partial[i]);
        mn = fminf(mn, v.x);
        mx = fmaxf(mx, v.y);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      if (lane == 0) {
        float d, zz;
        qdiff_params(mn, mx, d, zz);
        *scale_out = d;
        *zp_out = zz - 128.0f;
        __threadfence();
        st_release_u32(&ws->flag, 1u);
      }
    }
    if (lane == 0) {
      spin_until_set(&ws->flag);
      s_delta = __ldcg(scale_out);
      s_z = __ldcg(zp_out) + 128.0f;
      // the last CTA to consume the flag leaves the workspace ready for the next call
      if (atomicAdd(&ws->done, 1u) == gridDim.x - 1) {
        ws->counter = 0; ws->done = 0;
        __threadfence();
        st_release_u32(&ws->flag, 0u);
      }
    }
  }
  __syncthreads();
  delta = s_delta;
  z = s_z;
}

// one qdiff code: clamp(rint(x / delta) + z, 0, 255) - 128, fp32 true division
__device__ __forceinline__ int qdiff_code(float x, float delta, float z) {
  float r = __fadd_rn(rintf(__fdiv_rn(x, delta)), z);
  r = fminf(fmaxf(r, 0.0f), 255.0f);
  return static_cast<int>(r) - 128;
}

// 8 halves (one 16-byte vector) -> 8 codes (one 8-byte vector)
__device__ __forceinline__ uint2 qdiff_vec8(const int4& raw, float delta, float z) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
  int q[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h2[i]);
    q[2 * i] = qdiff_code(f.x, delta, z);
    q[2 * i + 1] = qdiff_code(f.y, delta, z);
  }
  uint2 out;
  out.x = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((q[3] & 0xff) << 24);
  out.y = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((q[7] & 0xff) << 24);
  return out;
}

__device__ __forceinline__ void minmax_vec8(const int4& raw, float& mn, float& mx) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h2[j]);
    mn = fminf(mn, fminf(f.x, f.y));
    mx = fmaxf(mx, fmaxf(f.x, f.y));
  }
}

}  // namespace mixdq
