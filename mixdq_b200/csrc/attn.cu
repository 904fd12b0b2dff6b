// attn.cu — cross-attention (encoder context of <= 96 tokens, head dim 64) of the fused
// transformer block, with the min/max partials of its output published for the single-pass
// quantiser of attn2.to_out (quant2.cu): replaces the library SDPA call (cuDNN: 5.7 us at
// batch 1, 20-40 CTAs) AND the separate min/max pass over its output (2.5 us) by one kernel.
//
// The reference leaves attention to stock PyTorch (diffusers Attention -> F.scaled_dot_product_
// attention on fp16); this is the same math with fp32 scores / softmax / accumulation and ONE
// rounding to fp16 at the end (SDPA's flash kernels round P to fp16 before P·V), i.e. at least as
// accurate; parity is a tolerance check against SDPA in tests/test_gpu_fused.py.
//
// (A first CUDA-core version — one warp per query, lanes over keys — measured 17 us per launch
// inside the UNet graph: 5 M warp-instructions per launch, and the next GEMM, resident early
// through programmatic dependent launch, leaves room for only two of its 27 KB CTAs per SM.)
#include "common.cuh"
#include "quant_ws.cuh"
#include "../../include/mixdq_b200.h"

namespace mixdq {

constexpr int kAttnThreads = 128;      // 4 warps x 16 queries
constexpr int kAttnMaxKeys = 96;
constexpr int kAttnQB = 64;            // queries per CTA
constexpr int kKPitch = 72;            // halves per K row in smem (144 B: fragment loads hit 32 banks)
constexpr int kVtPitch = 104;          // halves per V^T row (208 B)

__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                             uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float x, float y) {
  const __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One CTA = one (batch, head, block of 64 queries); a warp owns 16 queries. S = Q K^T and O = P V
// run on the tensor cores through mma.sync m16n8k16 (fp16 operands, fp32 accumulation; attention
// is a few hundred MFLOP per step, nowhere near the tcgen05 path's league — what matters here is
// instruction count and latency): S accumulators (16 x 96) stay in registers, softmax is done on
// them with quad shuffles, and they are re-packed in place as the A fragments of P V (the
// accumulator layout of one k16 pair of n8 tiles IS the A layout). K is staged row-major, V
// transposed, both with pitches that make the fragment loads bank-conflict free.
template <int NT8>   // 8-key tiles of S kept per warp: ceil(Lk / 16) * 2  (10 for 77 keys, max 12)
__global__ void __launch_bounds__(kAttnThreads)
cross_attn_d64_kernel(const __half* __restrict__ q, int64_t ldq, int64_t bsq,
                      const __half* __restrict__ k, int64_t ldk, int64_t bsk,
                      const __half* __restrict__ v, int64_t ldv, int64_t bsv,
                      __half* __restrict__ out, int T, int Lk, int H, float scale,
                      DynWs* __restrict__ ws, int vec16) {
  __shared__ __align__(16) __half s_k[kAttnMaxKeys * kKPitch];     // [key][d]
  __shared__ __align__(16) __half s_vt[64 * kVtPitch];             // [d][key]
  __shared__ float s_mn[kAttnThreads / 32], s_mx[kAttnThreads / 32];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * kAttnQB + warp * 16;
  constexpr int NKEY = NT8 * 8;                // keys covered by the S tiles (multiple of 16)
  // zero the key padding [Lk, NKEY) so that P (= 0 there) never multiplies garbage (NaN bits)
  for (int i = threadIdx.x; i < (NKEY - Lk) * 64; i += kAttnThreads) {
    const int key = Lk + i / 64, d = i % 64;
    s_k[key * kKPitch + d] = __float2half(0.f);
    s_vt[d * kVtPitch + key] = __float2half(0.f);
  }
  pdl_wait();
  // ---- stage K (row-major) and V (transposed) of this (batch, head) ----
  const __half* kb = k + b * bsk + h * 64;
  const __half* vb = v + b * bsv + h * 64;
  if (vec16) {
    constexpr int MAXIT = (kAttnMaxKeys * 8 + kAttnThreads - 1) / kAttnThreads;   // 6
    uint4 kk[MAXIT], vv[MAXIT];
#pragma unroll
    for (int u = 0; u < MAXIT; ++u) {
      const int i = threadIdx.x + u * kAttnThreads;
      if (i < Lk * 8) {
        kk[u] = __ldcg(reinterpret_cast<const uint4*>(kb + (i >> 3) * ldk + 8 * (i & 7)));
        vv[u] = __ldcg(reinterpret_cast<const uint4*>(vb + (i >> 3) * ldv + 8 * (i & 7)));
      }
    }
#pragma unroll
    for (int u = 0; u < MAXIT; ++u) {
      const int i = threadIdx.x + u * kAttnThreads;
      if (i < Lk * 8) {
        const int key = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(&s_k[key * kKPitch + 8 * ch]) = kk[u];
        const __half* vh = reinterpret_cast<const __half*>(&vv[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s_vt[(8 * ch + j) * kVtPitch + key] = vh[j];
      }
    }
  } else {
    for (int i = threadIdx.x; i < Lk * 32; i += kAttnThreads) {
      const int key = i >> 5, dp = i & 31;
      *reinterpret_cast<__half2*>(&s_k[key * kKPitch + 2 * dp]) =
          *reinterpret_cast<const __half2*>(kb + key * ldk + 2 * dp);
      const __half2 vv = *reinterpret_cast<const __half2*>(vb + key * ldv + 2 * dp);
      s_vt[(2 * dp) * kVtPitch + key] = __low2half(vv);
      s_vt[(2 * dp + 1) * kVtPitch + key] = __high2half(vv);
    }
  }
  // ---- Q fragments straight from global: rows t0+g and t0+g+8, 4 k-steps of 16 dims ----
  uint32_t qa[4][4];
  {
    const int r0 = t0 + g, r1 = t0 + g + 8;
    const __half* q0 = q + b * bsq + static_cast<int64_t>(r0) * ldq + h * 64 + 2 * tg;
    const __half* q1 = q + b * bsq + static_cast<int64_t>(r1) * ldq + h * 64 + 2 * tg;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa[ks][0] = (r0 < T) ? __ldcg(reinterpret_cast<const uint32_t*>(q0 + 16 * ks)) : 0u;
      qa[ks][1] = (r1 < T) ? __ldcg(reinterpret_cast<const uint32_t*>(q1 + 16 * ks)) : 0u;
      qa[ks][2] = (r0 < T) ? __ldcg(reinterpret_cast<const uint32_t*>(q0 + 16 * ks + 8)) : 0u;
      qa[ks][3] = (r1 < T) ? __ldcg(reinterpret_cast<const uint32_t*>(q1 + 16 * ks + 8)) : 0u;
    }
  }
  __syncthreads();
  float mn = 0.f, mx = 0.f;
  if (t0 < T) {
    // ---- S = Q K^T : NT8 tiles of 16 x 8 ----
    float sacc[NT8][4];
#pragma unroll
    for (int n = 0; n < NT8; ++n) {
      sacc[n][0] = sacc[n][1] = sacc[n][2] = sacc[n][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const __half* kp = &s_k[(8 * n + g) * kKPitch + 16 * ks + 2 * tg];
        mma_m16n8k16(sacc[n], qa[ks], *reinterpret_cast<const uint32_t*>(kp),
                     *reinterpret_cast<const uint32_t*>(kp + 8));
      }
    }
    // ---- softmax over the keys (rows g and g+8 of this warp's 16 queries) ----
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = 8 * n + 2 * tg + (e & 1);
        sacc[n][e] = (key < Lk) ? sacc[n][e] * scale : -INFINITY;
      }
      m0 = fmaxf(m0, fmaxf(sacc[n][0], sacc[n][1]));
      m1 = fmaxf(m1, fmaxf(sacc[n][2], sacc[n][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT8; ++n) {
      sacc[n][0] = expf(sacc[n][0] - m0); sacc[n][1] = expf(sacc[n][1] - m0);
      sacc[n][2] = expf(sacc[n][2] - m1); sacc[n][3] = expf(sacc[n][3] - m1);
      l0 += sacc[n][0] + sacc[n][1];
      l1 += sacc[n][2] + sacc[n][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- O = P V : P (fp16) re-packed from the S accumulators, 8 tiles of 16 x 8 dims ----
    float oacc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f;
#pragma unroll
    for (int kk2 = 0; kk2 < NT8 / 2; ++kk2) {           // 16 keys per step
      uint32_t pa[4];
      pa[0] = pack_half2(sacc[2 * kk2][0], sacc[2 * kk2][1]);
      pa[1] = pack_half2(sacc[2 * kk2][2], sacc[2 * kk2][3]);
      pa[2] = pack_half2(sacc[2 * kk2 + 1][0], sacc[2 * kk2 + 1][1]);
      pa[3] = pack_half2(sacc[2 * kk2 + 1][2], sacc[2 * kk2 + 1][3]);
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const __half* vp = &s_vt[(8 * n + g) * kVtPitch + 16 * kk2 + 2 * tg];
        mma_m16n8k16(oacc[n], pa, *reinterpret_cast<const uint32_t*>(vp),
                     *reinterpret_cast<const uint32_t*>(vp + 8));
      }
    }
    // ---- normalise, round once, store, min / max ----
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int C = H * 64;
    const int r0 = t0 + g, r1 = t0 + g + 8;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const __half2 a = __floats2half2_rn(oacc[n][0] * i0, oacc[n][1] * i0);
      const __half2 c = __floats2half2_rn(oacc[n][2] * i1, oacc[n][3] * i1);
      if (r0 < T) {
        *reinterpret_cast<__half2*>(out + (static_cast<int64_t>(b) * T + r0) * C + h * 64 + 8 * n + 2 * tg) = a;
        const float2 f = __half22float2(a);
        mn = fminf(mn, fminf(f.x, f.y)); mx = fmaxf(mx, fmaxf(f.x, f.y));
      }
      if (r1 < T) {
        *reinterpret_cast<__half2*>(out + (static_cast<int64_t>(b) * T + r1) * C + h * 64 + 8 * n + 2 * tg) = c;
        const float2 f = __half22float2(c);
        mn = fminf(mn, fminf(f.x, f.y)); mx = fmaxf(mx, fmaxf(f.x, f.y));
      }
    }
  }
  // ---- this CTA's min / max partial (quant2.cu protocol) ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < kAttnThreads / 32; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    ws->partial[cta] = make_float2(mn, mx);
  }
}

}  // namespace mixdq

using namespace mixdq;

static inline bool al4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; }

template <int NT8>
static cudaError_t launch_attn(dim3 grid, cudaStream_t st, const __half* q, int64_t ldq, int64_t bsq,
                               const __half* k, int64_t ldk, int64_t bsk, const __half* v,
                               int64_t ldv, int64_t bsv, __half* out, int T, int Lk, int H,
                               float scale, DynWs* ws, int vec16) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kAttnThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, cross_attn_d64_kernel<NT8>, q, ldq, bsq, k, ldk, bsk, v, ldv, bsv,
                            out, T, Lk, H, scale, ws, vec16);
}

extern "C" int mixdq_cross_attn_d64_f16(const mixdq_half_t* q, int64_t ldq, int64_t bsq,
                                        const mixdq_half_t* k, int64_t ldk, int64_t bsk,
                                        const mixdq_half_t* v, int64_t ldv, int64_t bsv,
                                        mixdq_half_t* out, int B, int T, int Lk, int H,
                                        float scale, void* ws, mixdq_stream_t stream) {
  if (!q || !k || !v || !out || !ws || B <= 0 || T <= 0 || Lk <= 0 || H <= 0)
    return MIXDQ_ERR_INVALID_ARG;
  if (!al4(q) || !al4(k) || !al4(v) || !al4(out) || (ldq & 1) || (ldk & 1) || (ldv & 1) ||
      (bsq & 1) || (bsk & 1) || (bsv & 1))
    return MIXDQ_ERR_ALIGNMENT;
  if (Lk > kAttnMaxKeys || H > 65535 || B > 65535) return MIXDQ_ERR_UNSUPPORTED;
  const int64_t ctas = static_cast<int64_t>((T + kAttnQB - 1) / kAttnQB) * H * B;
  if (ctas > kMaxPartials) return MIXDQ_ERR_UNSUPPORTED;
  dim3 grid((T + kAttnQB - 1) / kAttnQB, H, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* qh = reinterpret_cast<const __half*>(q);
  const __half* kh = reinterpret_cast<const __half*>(k);
  const __half* vh = reinterpret_cast<const __half*>(v);
  __half* oh = reinterpret_cast<__half*>(out);
  DynWs* w = static_cast<DynWs*>(ws);
  // 16-byte staging loads need 16-byte aligned K / V rows
  const int vec16 = ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                    (ldk % 8 == 0) && (ldv % 8 == 0) && (bsk % 8 == 0) && (bsv % 8 == 0);
  cudaError_t e;
  const int nt8 = ((Lk + 15) / 16) * 2;       // S tiles of 8 keys, in pairs (k16 steps of P V)
  switch (nt8) {
    case 2: e = launch_attn<2>(grid, st, qh, ldq, bsq, kh, ldk, bsk, vh, ldv, bsv, oh, T, Lk, H, scale, w, vec16); break;
    case 4: e = launch_attn<4>(grid, st, qh, ldq, bsq, kh, ldk, bsk, vh, ldv, bsv, oh, T, Lk, H, scale, w, vec16); break;
    case 6: e = launch_attn<6>(grid, st, qh, ldq, bsq, kh, ldk, bsk, vh, ldv, bsv, oh, T, Lk, H, scale, w, vec16); break;
    case 8: e = launch_attn<8>(grid, st, qh, ldq, bsq, kh, ldk, bsk, vh, ldv, bsv, oh, T, Lk, H, scale, w, vec16); break;
    case 10: e = launch_attn<10>(grid, st, qh, ldq, bsq, kh, ldk, bsk, vh, ldv, bsv, oh, T, Lk, H, scale, w, vec16); break;
    default: e = launch_attn<12>(grid, st, qh, ldq, bsq, kh, ldk, bsk, vh, ldv, bsv, oh, T, Lk, H, scale, w, vec16); break;
  }
  if (e != cudaSuccess) return MIXDQ_ERR_CUDA;
  partial_count_slot(ws) = static_cast<int>(ctas);
  return MIXDQ_OK;
}
