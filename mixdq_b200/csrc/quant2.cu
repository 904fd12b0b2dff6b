// quant2.cu — dynamic activation quantisation (A10) as SHORT kernels chained by programmatic
// dependent launch, instead of one kernel with a grid barrier in the middle.
//
// Measured inside the batch-1 UNet graph on B200 (tools/quant_phase.py, %globaltimer stamps): the
// single-kernel quantisers spent 2.6-3.6 us in the grid barrier alone (store + release fence +
// acquire spin + reload, all dependent L2 round trips) and ~2 us in each fully unrolled 50-60 KB
// code phase that a CTA executes exactly once (instruction fetch, not arithmetic), while a kernel
// boundary under programmatic dependent launch costs ~1.0 us. Hence:
//
//   pass 1  (minmax_rows_kernel | ln_minmax_kernel | gn_apply in fused_quant.cu | the GEGLU
//           epilogue of tc_i8_kernel) produce the fp16 values (if any op is fused); every CTA
//           stores ITS min / max into DynWs::partial[cta] — plain stores, no atomics, nothing to
//           reset;
//   pass 2  quant_rows_premm_kernel: every CTA reduces the producer's partials (a few KB from
//           L2, issued together with its first data load), quantises and writes int8.
//
// Every kernel here is a few KB of code: loops are not unrolled beyond what memory-level
// parallelism needs, min/max runs on packed halves (HMNMX2, exact), and the one-in-500 exact
// division of the rounding fix-up lives in a single out-of-line function.
#include "common.cuh"
#include "quant_ws.cuh"
#include "../../include/mixdq_b200.h"

namespace mixdq {

// CTA size of the min/max and quantise passes. Measured on the batch-1 step (ms/step): 128 threads
// 8.17, 256: 7.66, 512: 7.52, 1024: 7.90; two vectors per thread in the quantise pass: 7.77.
constexpr int kQ2Threads = 512;

// rare path of qdiff_round_quot, kept out of line so the hot loop stays small
__device__ __noinline__ float exact_round_quot(float x, float delta) {
  return rintf(__fdiv_rn(x, delta));
}

// 8 halves -> 8 codes in ~7 instructions per element (it was 21: FRND / F2I / clamp / shift / mask
// per element made the quantise pass issue-bound at batch >= 8, ncu: 250 warp instructions per
// vector against 126 MB of traffic per launch):
//   u = fma(x, 1/delta, 1.5 * 2^23)        -> the low mantissa bits of u hold k = rint(x / delta)
//                                             (exact product, ONE rounding; |x / delta| <= 255)
//   d = fma(x, 1/delta, -(u - 1.5 * 2^23)) -> distance of the exact product to k
//   the exact IEEE division only when some |d| of the vector exceeds 0.4999: the reference rounds
//   the correctly rounded fp32 QUOTIENT (torch.round(x / delta), base_quantizer.py:186), which can
//   differ from rint(x * (1/delta)) only that close to a rounding boundary (see qdiff_round_quot
//   in quant_ws.cuh; ~6 % of the warps take the out-of-line path);
//   code - shift = sat_s8(k + z - shift): integer add on the bit pattern of u, clamped by the
//   saturating cvt.pack (8 bit: [lo, hi] = [-128, 127] IS the s8 range; 4 bit: explicit min / max).
__device__ __forceinline__ uint2 quant8_compact(const int4& raw, float delta, float inv, float z,
                                                float qmax = 255.0f, int shift = 128) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
  constexpr float kMagic = 12582912.0f;    // 1.5 * 2^23, bit pattern 0x4B400000
  float x[8];
  int k[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h2[i]);
    x[2 * i] = f.x;
    x[2 * i + 1] = f.y;
  }
  float worst = 0.0f;                      // max |x / delta - rint| over the vector
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float u = __fmaf_rn(x[i], inv, kMagic);
    k[i] = __float_as_int(u) - 0x4B400000;
    worst = fmaxf(worst, fabsf(__fmaf_rn(x[i], inv, -__fsub_rn(u, kMagic))));
  }
  if (worst > 0.4999f) {                   // some element needs the correctly rounded quotient
    // the arrays are ROTATED so that the loop body only touches element 0 (static register
    // indexing, one call site); only flagged elements pay for the division
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      int e = k[0];
      if (fabsf(__fmaf_rn(x[0], inv, -static_cast<float>(e))) > 0.4999f)
        e = static_cast<int>(exact_round_quot(x[0], delta));
      const float x0 = x[0];
#pragma unroll
      for (int j = 0; j < 7; ++j) { x[j] = x[j + 1]; k[j] = k[j + 1]; }
      x[7] = x0;
      k[7] = e;
    }
  }
  const int zs = static_cast<int>(z) - shift;                   // z is integer-valued
  int c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = k[i] + zs;
  if (!(shift == 128 && qmax == 255.0f)) {
    const int lo = -shift, hi = static_cast<int>(qmax) - shift;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = min(max(c[i], lo), hi);
  }
  uint2 out;
  out.x = pack_sat_s8(c[1], c[0], pack_sat_s8(c[3], c[2], 0u));
  out.y = pack_sat_s8(c[5], c[4], pack_sat_s8(c[7], c[6], 0u));
  return out;
}

// Touch an address BEFORE the programmatic-dependency wait: the line may still be stale (the
// value is never used), but the translation and the L2 lookup are warm when the real ld.cg
// follows the wait (measured: the first dependent load after a kernel boundary costs ~0.6 us).
__device__ __forceinline__ void warm(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// packed-half running min / max of one 16-byte vector (exact: no rounding in min/max)
__device__ __forceinline__ void hminmax8(const int4& raw, __half2& mn, __half2& mx) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mn = __hmin2(mn, h2[i]);
    mx = __hmax2(mx, h2[i]);
  }
}

// CTA-wide reduction of the packed running min / max; thread 0 stores the CTA's partial.
template <int NT>
__device__ __forceinline__ void publish_partial(DynWs* __restrict__ ws, __half2 mn2, __half2 mx2) {
  constexpr int NW = NT / 32;
  __shared__ float s_mn[NW], s_mx[NW];
  float mn = fminf(__low2float(mn2), __high2float(mn2));
  float mx = fmaxf(__low2float(mx2), __high2float(mx2));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (NW > 1) {
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (warp == 0) {
      mn = lane < NW ? s_mn[lane] : 0.0f;
      mx = lane < NW ? s_mx[lane] : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
    }
  }
  // qdiff clamps x_min <= 0 <= x_max (base_quantizer.py:155-158)
  if (threadIdx.x == 0) ws->partial[blockIdx.x] = make_float2(fminf(mn, 0.0f), fmaxf(mx, 0.0f));
}

// ---------------------------------------------------------------------------------------------
// pass 1, plain tensor: min / max of a row-pitched fp16 view [M][8*nchunks] (pitch ldx halves)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQ2Threads)
minmax_rows_kernel(const __half* __restrict__ x, int64_t ldx, unsigned int nchunks,
                   unsigned int items, DynWs* __restrict__ ws) {
  QDbg dbg;
  dbg.begin(ws);
  pdl_launch_dependents();
  {
    const unsigned int it0 = blockIdx.x * kQ2Threads + threadIdx.x;
    if (it0 < items) {
      const unsigned int r0 = it0 / nchunks;
      warm(reinterpret_cast<const int4*>(x + static_cast<int64_t>(r0) * ldx) + (it0 - r0 * nchunks));
    }
  }
  pdl_wait();
  dbg.waited(ws);
  __half2 mn = __float2half2_rn(0.0f), mx = mn;
  const unsigned int stride = gridDim.x * kQ2Threads;
#pragma unroll 2
  for (unsigned int it = blockIdx.x * kQ2Threads + threadIdx.x; it < items; it += stride) {
    const unsigned int r = it / nchunks;
    const unsigned int c = it - r * nchunks;
    hminmax8(__ldcg(reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx) + c), mn, mx);
  }
  dbg.stamp(2);
  publish_partial<kQ2Threads>(ws, mn, mx);
  dbg.stamp(3);
  dbg.end(ws);
}

// ---------------------------------------------------------------------------------------------
// pass 2: quantise a row-pitched fp16 view with the min / max of the producer's `nparts` partials
// -> dense int8. `zero_words` (optional): u64 words block 0 clears for the next producer
// (GroupNorm statistics accumulators).
// ---------------------------------------------------------------------------------------------
// U = 16-byte vectors a thread keeps in flight: 1 for the batch-1 tensors (one vector per thread, the
// smallest code), 4 for tensors of several waves (measured at batch 8: with one load in flight per
// thread and 2 resident CTAs per SM the pass ran at ~1 TB/s, 10 us per launch).
// qmax / shift: code range [0, qmax] stored as code - shift (255 / 128 for 8 bit, 15 / 0 for the
// 4-bit activation layers).
template <int U>
__global__ void __launch_bounds__(kQ2Threads, U > 1 ? 2 : 1)
quant_rows_premm_kernel(const __half* __restrict__ x, int64_t ldx, unsigned int nchunks,
                        unsigned int items, int8_t* __restrict__ q, DynWs* __restrict__ ws,
                        int nparts, float* __restrict__ scale_out, float* __restrict__ zp_out,
                        unsigned long long* __restrict__ zero_words, int zero_n, float qmax,
                        int shift) {
  __shared__ float s_mn[kQ2Threads / 32], s_mx[kQ2Threads / 32];
  QDbg dbg;
  dbg.begin(ws);
  // dependents may launch right away. Triggering later was tried (after the dependency wait, after
  // the first loads, after the parameters, after the quantise loop): same-box A/B runs gave
  // 7.38-7.50 ms/step for this placement against 7.54 / 7.57 / 7.93 / 8.15.
  pdl_launch_dependents();
  const unsigned int stride = gridDim.x * kQ2Threads;
  unsigned int it = blockIdx.x * kQ2Threads + threadIdx.x;
  auto vec_ptr = [&](unsigned int i) {
    const unsigned int r = i / nchunks;
    return reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx) + (i - r * nchunks);
  };
  if (it < items) warm(vec_ptr(it));
  if (static_cast<int>(threadIdx.x) < nparts) warm(&ws->partial[threadIdx.x]);
  pdl_wait();
  dbg.waited(ws);
  // first data vector(s) and the partials travel together
  int4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    v[u] = make_int4(0, 0, 0, 0);
    const unsigned int i = it + u * stride;
    if (i < items) v[u] = __ldcg(vec_ptr(i));
  }
  float mn = 0.0f, mx = 0.0f;
#pragma unroll 1
  for (int i = threadIdx.x; i < nparts; i += kQ2Threads) {
    const float2 p = __ldcg(&ws->partial[i]);
    mn = fminf(mn, p.x);
    mx = fmaxf(mx, p.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kQ2Threads / 32; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
  float delta, z;
  qdiff_params(mn, mx, delta, z, qmax);
  const float inv = __frcp_rn(delta);
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) { *scale_out = delta; *zp_out = z - static_cast<float>(shift); }
#pragma unroll 1
    for (int i = threadIdx.x; i < zero_n; i += kQ2Threads) zero_words[i] = 0ull;
  }
  dbg.stamp(2);
  uint2* qv = reinterpret_cast<uint2*>(q);
  if (U == 1) {
#pragma unroll 1
    while (it < items) {                     // one inlined copy of the quantiser
      const uint2 codes = quant8_compact(v[0], delta, inv, z, qmax, shift);
      const unsigned int nxt = it + stride;
      if (nxt < items) v[0] = __ldcg(vec_ptr(nxt));
      qv[it] = codes;
      it = nxt;
    }
  } else {
    // batches of U vectors per thread, all U loads in flight together; two CTAs per SM (64
    // registers) overlap one CTA's conversion with the other's loads.
    const unsigned long long step = static_cast<unsigned long long>(U) * stride;
    unsigned long long base = it;
#pragma unroll 1
    while (base < items) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned long long i = base + static_cast<unsigned long long>(u) * stride;
        if (i < items) qv[i] = quant8_compact(v[u], delta, inv, z, qmax, shift);
      }
      base += step;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned long long k = base + static_cast<unsigned long long>(u) * stride;
        if (k < items) v[u] = __ldcg(vec_ptr(static_cast<unsigned int>(k)));
      }
    }
  }
  dbg.stamp(3);
  dbg.end(ws);
}

// ---------------------------------------------------------------------------------------------
// pass 1, LayerNorm: y = half(gamma * (rstd * (x - mean)) + beta) -> fp16 [M][C] + min / max.
// One row per warp at a time; 2 warps per CTA for the 256-token blocks of the batch-1 step (128
// CTAs), 8 warps and a row loop (<= 512 CTAs = partials) for larger M.
// PyTorch: statistics in fp32, biased variance (same restatement as fused_quant.cu::ln_row).
// ---------------------------------------------------------------------------------------------
// ROWS independent rows per warp and loop iteration: ROWS x MAXCH sixteen-byte loads in flight per
// lane (narrow rows — C = 320: 40 chunks, 1.25 per lane — left the big-batch launches at
// ~2.7 TB/s with one row at a time). Rows never interact, so the results do not depend on ROWS.
template <int MAXCH, int NT, int ROWS = 1>
__global__ void __launch_bounds__(NT)
ln_minmax_kernel(const __half* __restrict__ x, int64_t ldx, int M, int C,
                 const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps,
                 __half* __restrict__ y, DynWs* __restrict__ ws, int8_t* __restrict__ qs,
                 const float* __restrict__ s_inv, const float* __restrict__ s_zp) {
  // qs != nullptr: STATIC activation scales — the normalised fp16 values are quantised right here
  // with the consumer's checkpoint parameters (no fp16 round trip, no min/max, no second kernel)
  QDbg dbg;
  dbg.begin(ws);
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = C >> 3;
  // gamma / beta do not depend on the producer: fetch them before the dependency wait
  int4 gr[MAXCH], br[MAXCH];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      gr[i] = __ldg(reinterpret_cast<const int4*>(gamma) + c);
      br[i] = __ldg(reinterpret_cast<const int4*>(beta) + c);
    }
  }
  {
    const int r0 = blockIdx.x * (NT / 32) + warp;
    if (r0 < M) {
#pragma unroll
      for (int i = 0; i < MAXCH; ++i)
        if (lane + 32 * i < nchunks)
          warm(reinterpret_cast<const int4*>(x + static_cast<int64_t>(r0) * ldx) + lane + 32 * i);
    }
  }
  const float q_inv = qs ? __ldg(s_inv) : 0.0f, q_zp = qs ? __ldg(s_zp) : 0.0f;   // checkpoint constants
  pdl_wait();
  dbg.waited(ws);
  __half2 mn = __float2half2_rn(0.0f), mx = mn;
  const int rstep = gridDim.x * (NT / 32);
#pragma unroll 1
  for (int rb = blockIdx.x * (NT / 32) + warp; rb < M; rb += ROWS * rstep) {
    int4 raw[ROWS][MAXCH];
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
      const int r = rb + u * rstep;
      const int4* xrow = reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx);
#pragma unroll
      for (int i = 0; i < MAXCH; ++i) {
        const int c = lane + 32 * i;
        if (r < M && c < nchunks) raw[u][i] = __ldcg(xrow + c);
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
    const int r = rb + u * rstep;
    if (r >= M) break;                       // warp-uniform
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      if (lane + 32 * i < nchunks) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[u][i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          sum += f.x;
          sum += f.y;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / static_cast<float>(C);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      if (lane + 32 * i < nchunks) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[u][i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          const float d0 = f.x - mean, d1 = f.y - mean;
          ss += d0 * d0;
          ss += d1 * d1;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / static_cast<float>(C) + eps);
    int4* yrow = reinterpret_cast<int4*>(y + static_cast<int64_t>(r) * C);
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[u][i]);
        const __half2* g2 = reinterpret_cast<const __half2*>(&gr[i]);
        const __half2* b2 = reinterpret_cast<const __half2*>(&br[i]);
        int4 out;
        __half2* o2 = reinterpret_cast<__half2*>(&out);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          const float2 g = __half22float2(g2[j]);
          const float2 b = __half22float2(b2[j]);
          o2[j] = __floats2half2_rn(fmaf(g.x, rstd * (f.x - mean), b.x),
                                    fmaf(g.y, rstd * (f.y - mean), b.y));
        }
        if (qs != nullptr) {
          reinterpret_cast<uint2*>(qs + static_cast<int64_t>(r) * C)[c] = static_quant8(out, q_inv, q_zp);
        } else {
          hminmax8(out, mn, mx);
          yrow[c] = out;
        }
      }
    }
    }
  }
  dbg.stamp(2);
  if (qs == nullptr) publish_partial<NT>(ws, mn, mx);
  dbg.stamp(3);
  dbg.end(ws);
}

// one CTA: (min, max) of the `nparts` partials a min/max pass left in the workspace
__global__ void __launch_bounds__(256)
minmax_finish_kernel(const DynWs* __restrict__ ws, int nparts, float* __restrict__ out) {
  __shared__ float s_mn[8], s_mx[8];
  pdl_wait();
  float mn = 0.0f, mx = 0.0f;
  for (int i = threadIdx.x; i < nparts; i += 256) {
    const float2 p = __ldcg(&ws->partial[i]);
    mn = fminf(mn, p.x);
    mx = fmaxf(mx, p.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
    out[0] = mn;
    out[1] = mx;
  }
}

static inline int grid_for2(int64_t items, int per_block, int max_blocks) {
  int64_t g = (items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

}  // namespace mixdq

using namespace mixdq;

// ---- internal entry points used by quant.cu / fused_quant.cu ---------------------------------
// All return MIXDQ_ERR_UNSUPPORTED for tensors of >= 2^31 16-byte vectors (callers fall back to
// the single-kernel path).
static const int64_t kMaxItems = (1ll << 31) - 1;


// pass 2 launch. One vector per thread while the tensor fits ~2.5 waves of the 2 resident
// 512-thread CTAs per SM (every batch-1 tensor); beyond that, four vectors in flight per thread on
// a resident grid of 2 CTAs per SM: every extra wave of the one-vector form pays the whole
// dependent chain again (partials -> CTA reduction -> first load, ~2 us; measured at batch 8: 3
// waves = 7.7 us for a 2.6 M-element tensor).
static int launch_pass2(const __half* x, int64_t ldx, unsigned int nchunks, unsigned int n,
                        int8_t* q, void* ws, int nparts, float* scale_out, float* zp_out,
                        unsigned long long* zero_words, int zero_n, int n_bits, cudaStream_t st) {
  const float qmax = n_bits == 4 ? 15.0f : 255.0f;
  const int shift = n_bits == 4 ? 0 : 128;
  cudaError_t e;
  if (n > 148u * 5u * kQ2Threads / 2u) {               // > 2.5 waves of the one-vector form
    const int g2 = 148 * 2;                              // the whole resident grid: balanced SMs
    e = launch_pdl(quant_rows_premm_kernel<4>, g2, kQ2Threads, 0, st, x, ldx, nchunks, n, q,
                   static_cast<DynWs*>(ws), nparts, scale_out, zp_out, zero_words, zero_n, qmax,
                   shift);
  } else {
    const int g2 = grid_for2(n, kQ2Threads, 148 * 8);
    e = launch_pdl(quant_rows_premm_kernel<1>, g2, kQ2Threads, 0, st, x, ldx, nchunks, n, q,
                   static_cast<DynWs*>(ws), nparts, scale_out, zp_out, zero_words, zero_n, qmax,
                   shift);
  }
  return e == cudaSuccess ? MIXDQ_OK : MIXDQ_ERR_CUDA;
}

// A10 of a row-pitched view. cols % 8 == 0, 16-byte aligned rows. n_bits = 8 (codes - 128) or 4
// (codes 0..15 as they are: the 4-bit activation layers of the act_7.xx bit configs).
int mixdq_q2_rows(const __half* x, int64_t ldx, int64_t M, int cols, int8_t* q, float* scale_out,
                  float* zp_out, void* ws, cudaStream_t st, int n_bits) {
  const int64_t items = M * (cols >> 3);
  if (items > kMaxItems) return MIXDQ_ERR_UNSUPPORTED;
  // dense: one long row (no division result other than 0)
  const unsigned int nchunks = (ldx == cols) ? static_cast<unsigned int>(items)
                                             : static_cast<unsigned int>(cols >> 3);
  const unsigned int n = static_cast<unsigned int>(items);
  const int g1 = grid_for2(items, kQ2Threads, 148 * 4);
  if (launch_pdl(minmax_rows_kernel, g1, kQ2Threads, 0, st, x, ldx, nchunks, n,
                 static_cast<DynWs*>(ws)) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return launch_pass2(x, ldx, nchunks, n, q, ws, g1, scale_out, zp_out, nullptr, 0, n_bits, st);
}

// pass 2 alone on a dense tensor: the producer's CTAs stored `nparts` partials into ws->partial
int mixdq_q2_premm(const __half* x, int64_t numel, int8_t* q, float* scale_out, float* zp_out,
                   void* ws, int nparts, unsigned long long* zero_words, int zero_n,
                   cudaStream_t st) {
  const int64_t items = numel >> 3;
  if (items > kMaxItems || nparts < 1 || nparts > kMaxPartials) return MIXDQ_ERR_UNSUPPORTED;
  const unsigned int n = static_cast<unsigned int>(items);
  return launch_pass2(x, 0, n, n, q, ws, nparts, scale_out, zp_out, zero_words, zero_n, 8, st);
}

// LayerNorm -> fp16 y (caller's buffer) + per-CTA min/max, then pass 2 on y
int mixdq_q2_ln(const __half* x, int64_t ldx, int M, int C, const __half* gamma,
                const __half* beta, float eps, int8_t* q, __half* y, float* scale_out,
                float* zp_out, void* ws, cudaStream_t st, int8_t* qs, const float* s_inv,
                const float* s_zp) {
  if (static_cast<int64_t>(M) * (C >> 3) > kMaxItems) return MIXDQ_ERR_UNSUPPORTED;
  DynWs* w = static_cast<DynWs*>(ws);
  cudaError_t e;
  int g1;
  if (y == nullptr && qs == nullptr) return MIXDQ_ERR_UNSUPPORTED;   // the two-pass form needs the scratch
  if (M <= 512) {
    g1 = (M + 1) / 2;
    e = (C <= 5 * 256)
            ? launch_pdl(ln_minmax_kernel<5, 64>, g1, 64, 0, st, x, ldx, M, C, gamma, beta, eps, y, w, qs, s_inv, s_zp)
            : launch_pdl(ln_minmax_kernel<8, 64>, g1, 64, 0, st, x, ldx, M, C, gamma, beta, eps, y, w, qs, s_inv, s_zp);
  } else {
    g1 = (M + 7) / 8;
    if (g1 > 512) g1 = 512;
    // several rows in flight per warp once every warp has more than one row to walk
    const bool multi = M >= 2 * 512 * 8;
    if (multi && C <= 2 * 256)
      e = launch_pdl(ln_minmax_kernel<2, 256, 4>, g1, 256, 0, st, x, ldx, M, C, gamma, beta, eps, y, w, qs, s_inv, s_zp);
    else if (multi && C <= 3 * 256)
      e = launch_pdl(ln_minmax_kernel<3, 256, 2>, g1, 256, 0, st, x, ldx, M, C, gamma, beta, eps, y, w, qs, s_inv, s_zp);
    else
      e = (C <= 5 * 256)
            ? launch_pdl(ln_minmax_kernel<5, 256>, g1, 256, 0, st, x, ldx, M, C, gamma, beta, eps, y, w, qs, s_inv, s_zp)
            : launch_pdl(ln_minmax_kernel<8, 256>, g1, 256, 0, st, x, ldx, M, C, gamma, beta, eps, y, w, qs, s_inv, s_zp);
  }
  if (e != cudaSuccess) return MIXDQ_ERR_CUDA;
  if (q == nullptr) return MIXDQ_OK;       // LayerNorm only / LayerNorm + static quantisation
  return mixdq_q2_premm(y, static_cast<int64_t>(M) * C, q, scale_out, zp_out, ws, g1, nullptr, 0, st);
}

// (min(0, min x), max(0, max x)) of a dense fp16 tensor -> out[2] (PTQ calibration: the clamped
// range of base_quantizer.py:155-158). numel % 8 == 0.
extern "C" int mixdq_minmax_f16(const mixdq_half_t* x, int64_t numel, float* out, void* ws,
                                mixdq_stream_t stream) {
  if (numel <= 0 || !x || !out || !ws) return MIXDQ_ERR_INVALID_ARG;
  if ((numel & 7) || (reinterpret_cast<uintptr_t>(x) & 15)) return MIXDQ_ERR_ALIGNMENT;
  const int64_t items = numel >> 3;
  if (items > kMaxItems) return MIXDQ_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned int n = static_cast<unsigned int>(items);
  const int g1 = grid_for2(items, kQ2Threads, 148 * 4);
  if (launch_pdl(minmax_rows_kernel, g1, kQ2Threads, 0, st, reinterpret_cast<const __half*>(x),
                 static_cast<int64_t>(0), n, n, static_cast<DynWs*>(ws)) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  if (launch_pdl(minmax_finish_kernel, 1, 256, 0, st, static_cast<const DynWs*>(ws), g1, out) !=
      cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return MIXDQ_OK;
}
