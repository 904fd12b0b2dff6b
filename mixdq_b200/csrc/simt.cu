// simt.cu — portable dp4a kernels for the shapes the tcgen05 path does not take
// (K or C not a multiple of 16, N not a multiple of 8 — e.g. SDXL conv_in C=4 / conv_out K=4 —
// strided convolutions, exotic padding) and as an independent on-device cross-check of the
// tensor-core kernels in the parity tests. Same integer arithmetic, same epilogue order.
#include "common.cuh"
#include "simt.h"

namespace mixdq {

__device__ __forceinline__ int dot_i8(const int8_t* __restrict__ a, const int8_t* __restrict__ b,
                                      int k, bool vec16) {
  int acc = 0;
  if (vec16) {
    const int4* a4 = reinterpret_cast<const int4*>(a);
    const int4* b4 = reinterpret_cast<const int4*>(b);
    for (int i = 0; i < (k >> 4); ++i) {
      const int4 x = __ldg(a4 + i), y = __ldg(b4 + i);
      acc = __dp4a(x.x, y.x, acc);
      acc = __dp4a(x.y, y.y, acc);
      acc = __dp4a(x.z, y.z, acc);
      acc = __dp4a(x.w, y.w, acc);
    }
  } else {
    const int* a1 = reinterpret_cast<const int*>(a);
    const int* b1 = reinterpret_cast<const int*>(b);
    for (int i = 0; i < (k >> 2); ++i) acc = __dp4a(__ldg(a1 + i), __ldg(b1 + i), acc);
  }
  return acc;
}

__device__ __forceinline__ int8_t nib_hi(uint8_t b) { return static_cast<int8_t>(b) >> 4; }
__device__ __forceinline__ int8_t nib_lo(uint8_t b) {
  return static_cast<int8_t>(static_cast<int8_t>(b << 4) >> 4);
}

__global__ void __launch_bounds__(256)
simt_gemm_kernel(SimtGemmArgs g) {
  const int n = blockIdx.y * 32 + (threadIdx.x & 31);
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= g.M || n >= g.N) return;
  int acc = 0, acc1 = 0;
  if (g.w4) {
    // packed signed 4-bit weights, even k in the high nibble
    const int8_t* a = g.A + static_cast<int64_t>(m) * g.lda;
    const uint8_t* w = reinterpret_cast<const uint8_t*>(g.W) + static_cast<int64_t>(n) * (g.K >> 1);
    for (int k = 0; k < g.K; k += 2) {
      const uint8_t b = __ldg(w + (k >> 1));
      acc += static_cast<int>(a[k]) * nib_hi(b) + static_cast<int>(a[k + 1]) * nib_lo(b);
    }
  } else {
    const bool v16 = (g.K % 16 == 0) && (g.lda % 16 == 0) &&
                     ((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.W)) & 15) == 0;
    acc = dot_i8(g.A + static_cast<int64_t>(m) * g.lda, g.W + static_cast<int64_t>(n) * g.K, g.K, v16);
    if (g.A1 != nullptr) {
      const bool v16b = (g.K1 % 16 == 0) && (g.lda1 % 16 == 0) &&
                        ((reinterpret_cast<uintptr_t>(g.A1) | reinterpret_cast<uintptr_t>(g.W1)) & 15) == 0;
      acc1 = dot_i8(g.A1 + static_cast<int64_t>(m) * g.lda1, g.W1 + static_cast<int64_t>(n) * g.K1,
                    g.K1, v16b);
    }
  }
  if (g.acc_out) g.acc_out[static_cast<int64_t>(m) * g.N + n] = acc;
  float sc, b0;
  if (g.a_scale) {
    sc = __fmul_rn(__ldg(g.scale + n), __ldg(g.a_scale));
    b0 = __fmul_rn(__ldg(g.bias0 + n), __ldg(g.a_zp));
  } else {
    sc = __ldg(g.scale + n);
    b0 = __ldg(g.bias0 + n);
  }
  float f = dequant_f32(acc, b0, sc);
  if (g.bias) f = __fadd_rn(f, __half2float(g.bias[n]));
  __half h = __float2half_rn(f);
  if (g.A1 != nullptr) {
    const float f1 = dequant_f32(acc1, __ldg(g.bias0_1 + n), __ldg(g.scale1 + n));
    h = __float2half_rn(__fadd_rn(__half2float(h), __half2float(__float2half_rn(f1))));
  }
  g.D[static_cast<int64_t>(m) * g.ldd + n] = h;
}

// Few output channels (SDXL conv_out: K = 4, C = 320): one WARP per output pixel, the lanes split
// the (tap, 16-channel chunk) items of the reduction and shuffle-add their INT32 partials (exact,
// order-independent); lane k < K finishes channel k with the same epilogue as below. The weights
// (K x R x S x C bytes, <= 46 KB) are staged in shared memory once per CTA and the CTAs walk the
// pixels with a grid stride: the first form fetched K sixteen-byte weight vectors per item from
// global memory (448 us per batch-64 launch for 2 MB of output). KMAX = 4 halves the shuffles.
template <int KMAX>
__global__ void __launch_bounds__(256)
simt_conv_smallk_kernel(SimtConvArgs c) {
  extern __shared__ int4 sk_w[];                       // [K][R*S][C/16]
  const int chunks = c.C >> 4;
  const int taps = c.R * c.S;
  const int items = taps * chunks;
  for (int i = threadIdx.x; i < c.K * items; i += blockDim.x)
    sk_w[i] = __ldg(reinterpret_cast<const int4*>(c.w) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned int npix = static_cast<unsigned int>(c.N) * c.P * c.Q;
  for (unsigned int pix = blockIdx.x * 8u + (threadIdx.x >> 5); pix < npix; pix += gridDim.x * 8u) {
    const int q = pix % c.Q;
    const int p = (pix / c.Q) % c.P;
    const int n = pix / (static_cast<unsigned int>(c.Q) * c.P);
    const int h0 = p * c.stride - c.pad, w0 = q * c.stride - c.pad;
    int acc[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) acc[k] = 0;
    // three items per lane and round: their activation loads are in flight together
    constexpr int UI = 3;
    for (int it0 = lane; it0 < items; it0 += 32 * UI) {
      int4 x[UI];
      bool ok[UI];
#pragma unroll
      for (int u = 0; u < UI; ++u) {
        const int it = it0 + 32 * u;
        const int tap = it / chunks, ch = it - tap * chunks;
        const int r = tap / c.S, s = tap - r * c.S;
        const int h = h0 + r, w = w0 + s;
        ok[u] = it < items && h >= 0 && h < c.H && w >= 0 && w < c.W;
        if (ok[u])
          x[u] = __ldg(reinterpret_cast<const int4*>(
                           c.x + ((static_cast<int64_t>(n) * c.H + h) * c.W + w) * c.x_cpitch) + ch);
      }
#pragma unroll
      for (int u = 0; u < UI; ++u) {
        if (!ok[u]) continue;
        const int it = it0 + 32 * u;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < c.K) {
            const int4 y = sk_w[k * items + it];
            acc[k] = __dp4a(x[u].x, y.x, acc[k]);
            acc[k] = __dp4a(x[u].y, y.y, acc[k]);
            acc[k] = __dp4a(x[u].z, y.z, acc[k]);
            acc[k] = __dp4a(x[u].w, y.w, acc[k]);
          }
        }
      }
    }
    int mine = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      int v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == k) mine = v;
    }
    if (lane >= c.K) continue;
    const int k = lane;
    float wacc = 0.f;
    if (c.wsum_krs) {
      for (int r = 0; r < c.R; ++r) {
        const int h = h0 + r;
        if (h < 0 || h >= c.H) continue;
        for (int s = 0; s < c.S; ++s) {
          const int w = w0 + s;
          if (w < 0 || w >= c.W) continue;
          wacc = __fadd_rn(wacc, __ldg(c.wsum_krs + (static_cast<int64_t>(k) * c.R + r) * c.S + s));
        }
      }
    }
    if (c.acc_out) c.acc_out[static_cast<int64_t>(pix) * c.K + k] = mine;
    const float b0 = c.wsum_krs ? __fmul_rn(wacc, __ldg(c.zp)) : __ldg(c.bias0_k + k);
    float f = dequant_f32(mine, b0, __ldg(c.scale + k));
    if (c.bias) f = __fadd_rn(f, __half2float(c.bias[k]));
    c.y[static_cast<int64_t>(pix) * c.K + k] = __float2half_rn(f);
  }
}

__global__ void __launch_bounds__(256)
simt_conv_kernel(SimtConvArgs c) {
  const int k = blockIdx.y * 32 + (threadIdx.x & 31);
  const unsigned int pix = blockIdx.x * 8u + (threadIdx.x >> 5);     // launch checks npix < 2^31
  const unsigned int npix = static_cast<unsigned int>(c.N) * c.P * c.Q;
  if (k >= c.K || pix >= npix) return;
  const int q = pix % c.Q;
  const int p = (pix / c.Q) % c.P;
  const int n = pix / (static_cast<unsigned int>(c.Q) * c.P);
  const int h0 = p * c.stride - c.pad, w0 = q * c.stride - c.pad;
  const bool v16 = (c.C % 16 == 0) && (c.x_cpitch % 16 == 0) &&
                   ((reinterpret_cast<uintptr_t>(c.x) | reinterpret_cast<uintptr_t>(c.w)) & 15) == 0;
  int acc = 0;
  float wacc = 0.f;
  for (int r = 0; r < c.R; ++r) {
    const int h = h0 + r;
    if (h < 0 || h >= c.H) continue;
    for (int s = 0; s < c.S; ++s) {
      const int w = w0 + s;
      if (w < 0 || w >= c.W) continue;
      const int8_t* xp = c.x + ((static_cast<int64_t>(n) * c.H + h) * c.W + w) * c.x_cpitch;
      const int8_t* wp = c.w + ((static_cast<int64_t>(k) * c.R + r) * c.S + s) * c.C;
      acc += dot_i8(xp, wp, c.C, v16);
      if (c.wsum_krs) wacc = __fadd_rn(wacc, __ldg(c.wsum_krs + (static_cast<int64_t>(k) * c.R + r) * c.S + s));
    }
  }
  if (c.acc_out) c.acc_out[static_cast<int64_t>(pix) * c.K + k] = acc;
  // zero-point propagation: float(acc_w) * zp (conv_act_zero_point_propagate.cu:35-49)
  const float b0 = c.wsum_krs ? __fmul_rn(wacc, __ldg(c.zp)) : __ldg(c.bias0_k + k);
  float f = dequant_f32(acc, b0, __ldg(c.scale + k));
  if (c.bias) f = __fadd_rn(f, __half2float(c.bias[k]));
  c.y[static_cast<int64_t>(pix) * c.K + k] = __float2half_rn(f);
}

// Four input channels (conv_in of both UNets: C = 4, 3x3, pad 1, stride 1): a pixel's channels are
// ONE 32-bit word, a tap of one output channel is one dp4a. A thread computes 8 consecutive output
// channels of kC4Px consecutive pixels of a row: the 3 x (kC4Px + 2) input words it needs are
// broadcast loads, the 9 x 8 weight words and border sums come from a [tap][k] copy in shared
// memory (two conflict-free LDS.128 per tap, reused for all its pixels), results leave as 16-byte
// stores. Same integers, same fp32 border-sum order (valid taps, r then s ascending) and the same
// epilogue as simt_conv_kernel, which spent 2 ms per batch-64 launch on this layer (one thread per
// output with two scalar loads per tap) against ~30 us of output traffic.
constexpr int kC4Px = 4;
__global__ void __launch_bounds__(256)
simt_conv_c4_kernel(SimtConvArgs c, int kgroups, int slots) {
  extern __shared__ int c4_smem[];
  int* s_w = c4_smem;                                              // [9][K] packed 4-channel words
  float* s_ws = reinterpret_cast<float*>(c4_smem + 9 * c.K);       // [9][K] border sums
  for (int i = threadIdx.x; i < 9 * c.K; i += blockDim.x) {
    const int tap = i / c.K, k = i - tap * c.K;
    s_w[i] = __ldg(reinterpret_cast<const int*>(c.w) + k * 9 + tap);
    s_ws[i] = __ldg(c.wsum_krs + k * 9 + tap);
  }
  __syncthreads();
  const int kg = threadIdx.x % kgroups, slot = threadIdx.x / kgroups;
  if (slot >= slots) return;
  const int qgroups = (c.Q + kC4Px - 1) / kC4Px;
  const long long items = static_cast<long long>(c.N) * c.P * qgroups;
  // the grid is capped at a few CTAs per SM: the 9 x K weight words are staged once per CTA
  for (long long item = static_cast<long long>(blockIdx.x) * slots + slot; item < items;
       item += static_cast<long long>(gridDim.x) * slots) {   // (n, p, q-group)
  const int qg = static_cast<int>(item % qgroups);
  const int p = static_cast<int>((item / qgroups) % c.P);
  const int n = static_cast<int>(item / (static_cast<long long>(qgroups) * c.P));
  const int q0 = qg * kC4Px, k0 = kg * 8;
  const int* x32 = reinterpret_cast<const int*>(c.x) + static_cast<int64_t>(n) * c.H * c.W;
  int xin[3][kC4Px + 2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int h = p - 1 + r;
#pragma unroll
    for (int j = 0; j < kC4Px + 2; ++j) {
      const int w = q0 - 1 + j;
      xin[r][j] = (h >= 0 && h < c.H && w >= 0 && w < c.W) ? __ldg(x32 + h * c.W + w) : 0;
    }
  }
  int acc[kC4Px][8];
  float wacc[kC4Px][8];
#pragma unroll
  for (int i = 0; i < kC4Px; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[i][j] = 0; wacc[i][j] = 0.f; }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const bool hv = (p - 1 + r) >= 0 && (p - 1 + r) < c.H;
#pragma unroll
    for (int s2 = 0; s2 < 3; ++s2) {
      const int tap = r * 3 + s2;
      const int4 w0 = *reinterpret_cast<const int4*>(s_w + tap * c.K + k0);
      const int4 w1 = *reinterpret_cast<const int4*>(s_w + tap * c.K + k0 + 4);
      const float4 f0 = *reinterpret_cast<const float4*>(s_ws + tap * c.K + k0);
      const float4 f1 = *reinterpret_cast<const float4*>(s_ws + tap * c.K + k0 + 4);
      const int wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
      for (int i = 0; i < kC4Px; ++i) {
        const int w = q0 + i - 1 + s2;
        const bool valid = hv && w >= 0 && w < c.W;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = __dp4a(xin[r][i + s2], wv[j], acc[i][j]);
          if (valid) wacc[i][j] = __fadd_rn(wacc[i][j], fv[j]);
        }
      }
    }
  }
  const float zp = __ldg(c.zp);
  float sc[8], bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(c.scale + k0 + j);
    bs[j] = c.bias ? __half2float(c.bias[k0 + j]) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < kC4Px; ++i) {
    const int q = q0 + i;
    if (q >= c.Q) break;
    const int64_t pix = (static_cast<int64_t>(n) * c.P + p) * c.Q + q;
    __align__(16) __half o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f = dequant_f32(acc[i][j], __fmul_rn(wacc[i][j], zp), sc[j]);
      if (c.bias) f = __fadd_rn(f, bs[j]);
      o[j] = __float2half_rn(f);
    }
    if (c.acc_out) {
#pragma unroll
      for (int j = 0; j < 8; ++j) c.acc_out[pix * c.K + k0 + j] = acc[i][j];
    }
    *reinterpret_cast<uint4*>(c.y + pix * c.K + k0) = *reinterpret_cast<const uint4*>(o);
  }
  }
}

int simt_gemm_launch(const SimtGemmArgs& g, cudaStream_t st) {
  dim3 grid((g.M + 7) / 8, (g.N + 31) / 32);
  if (grid.y > 65535) return -1;
  simt_gemm_kernel<<<grid, 256, 0, st>>>(g);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int simt_conv_launch(const SimtConvArgs& c, cudaStream_t st) {
  const int64_t npix = static_cast<int64_t>(c.N) * c.P * c.Q;
  dim3 grid(static_cast<unsigned>((npix + 7) / 8), (c.K + 31) / 32);
  if (grid.y > 65535 || npix > 2147483647LL) return -1;
  const bool v16 = (c.C % 16 == 0) && (c.x_cpitch % 16 == 0) &&
                   ((reinterpret_cast<uintptr_t>(c.x) | reinterpret_cast<uintptr_t>(c.w)) & 15) == 0;
  const bool c4 = c.C == 4 && c.x_cpitch == 4 && c.R == 3 && c.S == 3 && c.stride == 1 && c.pad == 1 &&
                  c.wsum_krs != nullptr && c.K % 8 == 0 && c.K / 8 <= 256 && c.K <= 1024 &&
                  ((reinterpret_cast<uintptr_t>(c.x) | reinterpret_cast<uintptr_t>(c.w)) & 3) == 0 &&
                  (reinterpret_cast<uintptr_t>(c.y) & 15) == 0;
  if (c4) {
    const int kgroups = c.K / 8, slots = 256 / kgroups;
    const long long items = static_cast<long long>(c.N) * c.P * ((c.Q + kC4Px - 1) / kC4Px);
    long long blocks = (items + slots - 1) / slots;
    if (blocks > 148 * 8) blocks = 148 * 8;
    simt_conv_c4_kernel<<<static_cast<unsigned>(blocks), 256, 9 * c.K * 8, st>>>(c, kgroups, slots);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
  }
  const size_t wbytes = static_cast<size_t>(c.K) * c.R * c.S * c.C;
  if (c.K <= 8 && v16 && wbytes <= 46 * 1024) {
    unsigned int blocks = grid.x < 148u * 8u ? grid.x : 148u * 8u;
    if (c.K <= 4) simt_conv_smallk_kernel<4><<<blocks, 256, wbytes, st>>>(c);
    else simt_conv_smallk_kernel<8><<<blocks, 256, wbytes, st>>>(c);
  } else
    simt_conv_kernel<<<grid, 256, 0, st>>>(c);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace mixdq
