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

// 8 halves -> 8 codes, compact: fast reciprocal path for all, exact fix-up only when any element
// of the vector sits within 1e-4 of a rounding boundary (see qdiff_round_quot in quant_ws.cuh;
// 1e-4 > the 5e-5 error bound, and keeps the out-of-line path to ~5 % of the warps)
__device__ __forceinline__ uint2 quant8_compact(const int4& raw, float delta, float inv, float z) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
  float x[8], r[8];
  unsigned int near = 0u;                  // bit i: element i needs the exact quotient
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h2[i]);
    x[2 * i] = f.x;
    x[2 * i + 1] = f.y;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float t = __fmul_rn(x[i], inv);
    r[i] = rintf(t);
    near |= (fabsf(__fsub_rn(t, r[i])) > 0.4999f ? 1u : 0u) << i;   // |t - x/delta| < 5e-5
  }
  if (near != 0u) {
    // the arrays are ROTATED so that the loop body only touches element 0 (static register
    // indexing, one call site); only flagged elements pay for the division
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      float e = r[0];
      if (near & 1u) e = exact_round_quot(x[0], delta);
      near >>= 1;
      const float x0 = x[0];
#pragma unroll
      for (int j = 0; j < 7; ++j) { x[j] = x[j + 1]; r[j] = r[j + 1]; }
      x[7] = x0;
      r[7] = e;
    }
  }
  uint32_t w[2] = {0u, 0u};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float c = __fadd_rn(r[i], z);
    c = fminf(fmaxf(c, 0.0f), 255.0f);
    const uint32_t b = static_cast<uint32_t>(static_cast<int>(c) - 128) & 0xffu;
    w[i >> 2] |= b << (8 * (i & 3));
  }
  return make_uint2(w[0], w[1]);
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
__global__ void __launch_bounds__(kQ2Threads)
quant_rows_premm_kernel(const __half* __restrict__ x, int64_t ldx, unsigned int nchunks,
                        unsigned int items, int8_t* __restrict__ q, DynWs* __restrict__ ws,
                        int nparts, float* __restrict__ scale_out, float* __restrict__ zp_out,
                        unsigned long long* __restrict__ zero_words, int zero_n) {
  __shared__ float s_mn[kQ2Threads / 32], s_mx[kQ2Threads / 32];
  QDbg dbg;
  dbg.begin(ws);
  // dependents may launch right away. Triggering later was tried (after the dependency wait, after
  // the first loads, after the parameters, after the quantise loop): same-box A/B runs gave
  // 7.38-7.50 ms/step for this placement against 7.54 / 7.57 / 7.93 / 8.15.
  pdl_launch_dependents();
  const unsigned int stride = gridDim.x * kQ2Threads;
  unsigned int it = blockIdx.x * kQ2Threads + threadIdx.x;
  if (it < items) {
    const unsigned int r0 = it / nchunks;
    warm(reinterpret_cast<const int4*>(x + static_cast<int64_t>(r0) * ldx) + (it - r0 * nchunks));
  }
  if (static_cast<int>(threadIdx.x) < nparts) warm(&ws->partial[threadIdx.x]);
  pdl_wait();
  dbg.waited(ws);
  // first data vector and the partials travel together
  int4 v = make_int4(0, 0, 0, 0);
  if (it < items) {
    const unsigned int r = it / nchunks;
    v = __ldcg(reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx) + (it - r * nchunks));
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
  qdiff_params(mn, mx, delta, z);
  const float inv = __frcp_rn(delta);
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) { *scale_out = delta; *zp_out = z - 128.0f; }
#pragma unroll 1
    for (int i = threadIdx.x; i < zero_n; i += kQ2Threads) zero_words[i] = 0ull;
  }
  dbg.stamp(2);
  uint2* qv = reinterpret_cast<uint2*>(q);
#pragma unroll 1
  while (it < items) {                     // one inlined copy of the quantiser
    const uint2 codes = quant8_compact(v, delta, inv, z);
    const unsigned int nxt = it + stride;
    if (nxt < items) {
      const unsigned int r = nxt / nchunks;
      v = __ldcg(reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx) + (nxt - r * nchunks));
    }
    qv[it] = codes;
    it = nxt;
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
template <int MAXCH, int NT>
__global__ void __launch_bounds__(NT)
ln_minmax_kernel(const __half* __restrict__ x, int64_t ldx, int M, int C,
                 const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps,
                 __half* __restrict__ y, DynWs* __restrict__ ws) {
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
  pdl_wait();
  dbg.waited(ws);
  __half2 mn = __float2half2_rn(0.0f), mx = mn;
#pragma unroll 1
  for (int r = blockIdx.x * (NT / 32) + warp; r < M; r += gridDim.x * (NT / 32)) {
    const int4* xrow = reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx);
    int4 raw[MAXCH];
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) raw[i] = __ldcg(xrow + c);
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      if (lane + 32 * i < nchunks) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[i]);
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
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[i]);
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
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[i]);
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
        hminmax8(out, mn, mx);
        yrow[c] = out;
      }
    }
  }
  dbg.stamp(2);
  publish_partial<NT>(ws, mn, mx);
  dbg.stamp(3);
  dbg.end(ws);
}

// ---------------------------------------------------------------------------------------------
// One-kernel variants for tensors whose values fit the REGISTERS of one co-resident grid (every
// transformer-block tensor of the batch-1 step): pass 1 and pass 2 above joined by a grid barrier
// made of ONE 8-byte store per CTA and plain polling loads — no fence, no atomic, no counter:
//   * min and max come from fp16 values, so their fp32 patterns have 13 zero low mantissa bits
//     each: 26 bits of room for a TAG = the launch's epoch (1 .. 2^26-1, never 0, so untagged
//     partials of the two-pass producers never match);
//   * a CTA publishes {min|tag_lo, max|tag_hi} with a single 64-bit store (single-copy atomic: a
//     reader that sees the tag sees the values), and thread i of every CTA polls slot i until it
//     carries the tag;
//   * the epoch is read after the dependency wait (the previous launch is complete) and bumped
//     by CTA 0 after it has passed the barrier (every CTA has read it by then).
// A release-increment / acquire-spin counter barrier measured 2.3-2.5 us here (the release fence
// dominates); this one is bounded by one store-to-load L2 round trip.
// ---------------------------------------------------------------------------------------------
constexpr unsigned int kTagMask = (1u << 13) - 1u;

__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All threads call it with the CTA's packed running min / max and the launch's tag (read from
// ws->epoch after the dependency wait); returns (delta, z) of the tensor.
template <int NT>
__device__ __forceinline__ void lean_grid_params(DynWs* __restrict__ ws, unsigned int epoch,
                                                 __half2 mn2, __half2 mx2,
                                                 float* __restrict__ scale_out,
                                                 float* __restrict__ zp_out, float& delta,
                                                 float& z) {
  constexpr int NW = NT / 32;
  __shared__ float s_mn[NW], s_mx[NW];
  const unsigned int tag = epoch % ((1u << 26) - 1u) + 1u;     // 1 .. 2^26-1
  float mn = fminf(__low2float(mn2), __high2float(mn2));
  float mx = fmaxf(__low2float(mx2), __high2float(mx2));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  unsigned long long* slots = reinterpret_cast<unsigned long long*>(ws->partial);
  const unsigned int tag_lo = tag & kTagMask, tag_hi = (tag >> 13) & kTagMask;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < NW; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
    // qdiff clamps x_min <= 0 <= x_max (base_quantizer.py:155-158)
    const unsigned int a = __float_as_uint(fminf(mn, 0.0f)) | tag_lo;
    const unsigned int b = __float_as_uint(fmaxf(mx, 0.0f)) | tag_hi;
    st_relaxed_u64(&slots[blockIdx.x], (static_cast<unsigned long long>(b) << 32) | a);
  }
  // thread i waits for CTA i's partial
  const unsigned int G = gridDim.x;
  mn = 0.0f; mx = 0.0f;
#pragma unroll 1
  for (unsigned int i = threadIdx.x; i < G; i += NT) {
    unsigned long long v;
    unsigned int spins = 0;
    do {
      v = ld_relaxed_u64(&slots[i]);
      if (++spins > (1u << 26)) __trap();     // protocol bug: fail instead of hanging the device
    } while ((static_cast<unsigned int>(v) & kTagMask) != tag_lo ||
             (static_cast<unsigned int>(v >> 32) & kTagMask) != tag_hi);
    mn = fminf(mn, __uint_as_float(static_cast<unsigned int>(v) & ~kTagMask));
    mx = fmaxf(mx, __uint_as_float(static_cast<unsigned int>(v >> 32) & ~kTagMask));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __syncthreads();                             // s_mn / s_mx are reused
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < NW; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
  qdiff_params(mn, mx, delta, z);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *scale_out = delta;
    *zp_out = z - 128.0f;
    // next launch's epoch (this CTA is past the barrier, so every CTA has read the current one)
    ws->epoch = epoch + 1u;
  }
}

__device__ __forceinline__ unsigned int lean_epoch(const DynWs* ws) {
  return *reinterpret_cast<const volatile unsigned int*>(&ws->epoch);
}

// plain tensor, row-pitched view -> dense int8; each thread owns <= 2 vectors (registers)
__global__ void __launch_bounds__(kQ2Threads)
quant_lean_kernel(const __half* __restrict__ x, int64_t ldx, unsigned int nchunks,
                  unsigned int items, int8_t* __restrict__ q, DynWs* __restrict__ ws,
                  float* __restrict__ scale_out, float* __restrict__ zp_out) {
  QDbg dbg;
  dbg.begin(ws);
  pdl_launch_dependents();
  pdl_wait();
  dbg.waited(ws);
  const unsigned int epoch = lean_epoch(ws);
  const unsigned int i0 = blockIdx.x * kQ2Threads + threadIdx.x;
  const unsigned int i1 = i0 + gridDim.x * kQ2Threads;
  int4 v0 = make_int4(0, 0, 0, 0), v1 = v0;
  if (i0 < items) {
    const unsigned int r = i0 / nchunks;
    v0 = __ldcg(reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx) + (i0 - r * nchunks));
  }
  if (i1 < items) {
    const unsigned int r = i1 / nchunks;
    v1 = __ldcg(reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx) + (i1 - r * nchunks));
  }
  __half2 mn = __float2half2_rn(0.0f), mx = mn;
  hminmax8(v0, mn, mx);     // zero-filled vectors do not move min <= 0 <= max
  hminmax8(v1, mn, mx);
  dbg.stamp(2);
  float delta, z;
  lean_grid_params<kQ2Threads>(ws, epoch, mn, mx, scale_out, zp_out, delta, z);
  dbg.stamp(3);
  const float inv = __frcp_rn(delta);
  uint2* qv = reinterpret_cast<uint2*>(q);
#pragma unroll 1
  for (int u = 0; u < 2; ++u) {            // one inlined copy of the quantiser
    const unsigned int it = u ? i1 : i0;
    if (it < items) qv[it] = quant8_compact(u ? v1 : v0, delta, inv, z);
  }
  dbg.end(ws);
}

// LayerNorm -> int8, one row per warp, the row's fp16 output stays in registers
template <int MAXCH, int NT>
__global__ void __launch_bounds__(NT)
ln_quant_lean_kernel(const __half* __restrict__ x, int64_t ldx, int M, int C,
                     const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps,
                     int8_t* __restrict__ q, __half* __restrict__ y_out, DynWs* __restrict__ ws,
                     float* __restrict__ scale_out, float* __restrict__ zp_out) {
  QDbg dbg;
  dbg.begin(ws);
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = C >> 3;
  const int r = blockIdx.x * (NT / 32) + warp;
  int4 gr[MAXCH], br[MAXCH];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      gr[i] = __ldg(reinterpret_cast<const int4*>(gamma) + c);
      br[i] = __ldg(reinterpret_cast<const int4*>(beta) + c);
    }
  }
  pdl_wait();
  dbg.waited(ws);
  const unsigned int epoch = lean_epoch(ws);
  __half2 mn = __float2half2_rn(0.0f), mx = mn;
  int4 out[MAXCH];
  if (r < M) {
    const int4* xrow = reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx);
    int4 raw[MAXCH];
#pragma unroll
    for (int i = 0; i < MAXCH; ++i)
      if (lane + 32 * i < nchunks) raw[i] = __ldcg(xrow + lane + 32 * i);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      if (lane + 32 * i < nchunks) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[i]);
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
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[i]);
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
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[i]);
        const __half2* g2 = reinterpret_cast<const __half2*>(&gr[i]);
        const __half2* b2 = reinterpret_cast<const __half2*>(&br[i]);
        __half2* o2 = reinterpret_cast<__half2*>(&out[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          const float2 g = __half22float2(g2[j]);
          const float2 b = __half22float2(b2[j]);
          o2[j] = __floats2half2_rn(fmaf(g.x, rstd * (f.x - mean), b.x),
                                    fmaf(g.y, rstd * (f.y - mean), b.y));
        }
        hminmax8(out[i], mn, mx);
        if (y_out) reinterpret_cast<int4*>(y_out + static_cast<int64_t>(r) * C)[c] = out[i];
      }
    }
  }
  dbg.stamp(2);
  float delta, z;
  lean_grid_params<NT>(ws, epoch, mn, mx, scale_out, zp_out, delta, z);
  dbg.stamp(3);
  if (r < M) {
    const float inv = __frcp_rn(delta);
    uint2* qrow = reinterpret_cast<uint2*>(q + static_cast<int64_t>(r) * C);
#pragma unroll 1
    for (int i = 0; i < MAXCH; ++i) {
      // rotate so that the loop body only touches out[0] (one inlined copy of the quantiser)
      const int c = lane + 32 * i;
      if (c < nchunks) qrow[c] = quant8_compact(out[0], delta, inv, z);
#pragma unroll
      for (int j = 0; j + 1 < MAXCH; ++j) out[j] = out[j + 1];
    }
  }
  dbg.end(ws);
}

// ---------------------------------------------------------------------------------------------
// ONE-CLUSTER variants (mode 3): the whole tensor in the registers of one thread-block cluster of
// 16 (or 8) CTAs x 512 threads; min/max meet through distributed shared memory and the hardware
// cluster barrier (~0.2 us), so a quantisation is a single short kernel with no global
// synchronisation at all. The first attempt at this (quant.cu / fused_quant.cu, 50 KB kernels with
// an IEEE division per element) was issue-bound on 16 SMs (7.5-10 us); with the compact quantiser
// (~12 instructions per element) the same 0.33 M elements are ~1 us of issue time on 16 SMs.
// ---------------------------------------------------------------------------------------------
constexpr int kClThreads = 512;
constexpr int kClVec = 6;            // 16-byte vectors per thread

__global__ void __launch_bounds__(kClThreads)
quant_cluster_kernel(const __half* __restrict__ x, int64_t ldx, unsigned int nchunks,
                     unsigned int items, int8_t* __restrict__ q, float* __restrict__ scale_out,
                     float* __restrict__ zp_out) {
  pdl_launch_dependents();
  cluster_enter();
  pdl_wait();
  const unsigned int stride = gridDim.x * kClThreads;
  const unsigned int i0 = blockIdx.x * kClThreads + threadIdx.x;
  int4 v[kClVec];
  float mn = 0.f, mx = 0.f;
  {
    __half2 mn2 = __float2half2_rn(0.0f), mx2 = mn2;
#pragma unroll
    for (int u = 0; u < kClVec; ++u) {
      const unsigned int it = i0 + u * stride;
      v[u] = make_int4(0, 0, 0, 0);
      if (it < items) {
        const unsigned int r = it / nchunks;
        v[u] = __ldcg(reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx) + (it - r * nchunks));
      }
    }
#pragma unroll
    for (int u = 0; u < kClVec; ++u) hminmax8(v[u], mn2, mx2);   // zero vectors keep min <= 0 <= max
    mn = fminf(__low2float(mn2), __high2float(mn2));
    mx = fmaxf(__low2float(mx2), __high2float(mx2));
  }
  float delta, z;
  cluster_minmax_params<kClThreads>(mn, mx, scale_out, zp_out, delta, z);
  const float inv = __frcp_rn(delta);
  uint2* qv = reinterpret_cast<uint2*>(q);
#pragma unroll 1
  for (int u = 0; u < kClVec; ++u) {       // rotate: one inlined copy of the quantiser
    const unsigned int it = i0 + u * stride;
    if (it < items) qv[it] = quant8_compact(v[0], delta, inv, z);
#pragma unroll
    for (int j = 0; j + 1 < kClVec; ++j) v[j] = v[j + 1];
  }
}

// LayerNorm -> int8, one row per warp, 16 rows per CTA
template <int MAXCH>
__global__ void __launch_bounds__(kClThreads)
ln_quant_cluster_kernel(const __half* __restrict__ x, int64_t ldx, int M, int C,
                        const __half* __restrict__ gamma, const __half* __restrict__ beta,
                        float eps, int8_t* __restrict__ q, __half* __restrict__ y_out,
                        float* __restrict__ scale_out, float* __restrict__ zp_out) {
  pdl_launch_dependents();
  cluster_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = C >> 3;
  const int r = blockIdx.x * (kClThreads / 32) + warp;
  pdl_wait();
  __half2 mn2 = __float2half2_rn(0.0f), mx2 = mn2;
  int4 out[MAXCH];
  if (r < M) {
    const int4* xrow = reinterpret_cast<const int4*>(x + static_cast<int64_t>(r) * ldx);
    int4 raw[MAXCH];
#pragma unroll
    for (int i = 0; i < MAXCH; ++i)
      if (lane + 32 * i < nchunks) raw[i] = __ldcg(xrow + lane + 32 * i);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      if (lane + 32 * i < nchunks) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[i]);
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
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[i]);
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
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        const int4 gr = __ldg(reinterpret_cast<const int4*>(gamma) + c);
        const int4 br = __ldg(reinterpret_cast<const int4*>(beta) + c);
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[i]);
        const __half2* g2 = reinterpret_cast<const __half2*>(&gr);
        const __half2* b2 = reinterpret_cast<const __half2*>(&br);
        __half2* o2 = reinterpret_cast<__half2*>(&out[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          const float2 g = __half22float2(g2[j]);
          const float2 b = __half22float2(b2[j]);
          o2[j] = __floats2half2_rn(fmaf(g.x, rstd * (f.x - mean), b.x),
                                    fmaf(g.y, rstd * (f.y - mean), b.y));
        }
        hminmax8(out[i], mn2, mx2);
        if (y_out) reinterpret_cast<int4*>(y_out + static_cast<int64_t>(r) * C)[c] = out[i];
      }
    }
  }
  float delta, z;
  cluster_minmax_params<kClThreads>(fminf(__low2float(mn2), __high2float(mn2)),
                                    fmaxf(__low2float(mx2), __high2float(mx2)), scale_out, zp_out,
                                    delta, z);
  if (r < M) {
    const float inv = __frcp_rn(delta);
    uint2* qrow = reinterpret_cast<uint2*>(q + static_cast<int64_t>(r) * C);
#pragma unroll 1
    for (int i = 0; i < MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) qrow[c] = quant8_compact(out[0], delta, inv, z);
#pragma unroll
      for (int j = 0; j + 1 < MAXCH; ++j) out[j] = out[j + 1];
    }
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

int mixdq_quant_mode();   // quant.cu: 0 single-kernel (old), 1 two-pass only, 2 lean one-kernel too

// A10 of a row-pitched view. cols % 8 == 0, 16-byte aligned rows.
int mixdq_q2_rows(const __half* x, int64_t ldx, int64_t M, int cols, int8_t* q, float* scale_out,
                  float* zp_out, void* ws, cudaStream_t st) {
  const int64_t items = M * (cols >> 3);
  if (items > kMaxItems) return MIXDQ_ERR_UNSUPPORTED;
  // dense: one long row (no division result other than 0)
  const unsigned int nchunks = (ldx == cols) ? static_cast<unsigned int>(items)
                                             : static_cast<unsigned int>(cols >> 3);
  const unsigned int n = static_cast<unsigned int>(items);
  // one-cluster form: the tensor fits the registers of 16 (8) CTAs x 512 threads x 6 vectors
  if (mixdq_quant_mode() == 3) {
    static int ncl = -1;
    if (ncl < 0) ncl = max_cluster_ctas(quant_cluster_kernel, kClThreads, 0);
    if (ncl > 0 && items <= static_cast<int64_t>(ncl) * kClThreads * kClVec) {
      int g = static_cast<int>((items + kClThreads * kClVec - 1) / (kClThreads * kClVec));
      // use the whole cluster: more SMs, fewer vectors per thread
      g = ncl;
      if (launch_cluster_pdl(quant_cluster_kernel, g, kClThreads, 0, st, x, ldx, nchunks, n, q,
                             scale_out, zp_out) != cudaSuccess)
        return MIXDQ_ERR_CUDA;
      return MIXDQ_OK;
    }
  }
  // register-resident one-kernel form: <= 2 vectors per thread on a co-resident grid
  // (256-thread CTAs without shared memory: 4 per SM is always resident)
  if (mixdq_quant_mode() == 2 && items <= static_cast<int64_t>(148) * 4 * kQ2Threads * 2) {
    const int g = grid_for2(items, kQ2Threads * 2, 148 * 4);
    if (launch_pdl(quant_lean_kernel, g, kQ2Threads, 0, st, x, ldx, nchunks, n, q,
                   static_cast<DynWs*>(ws), scale_out, zp_out) != cudaSuccess)
      return MIXDQ_ERR_CUDA;
    return MIXDQ_OK;
  }
  const int g1 = grid_for2(items, kQ2Threads, 148 * 4);
  if (launch_pdl(minmax_rows_kernel, g1, kQ2Threads, 0, st, x, ldx, nchunks, n,
                 static_cast<DynWs*>(ws)) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  const int g2 = grid_for2(items, kQ2Threads, 148 * 8);
  if (launch_pdl(quant_rows_premm_kernel, g2, kQ2Threads, 0, st, x, ldx, nchunks, n, q,
                 static_cast<DynWs*>(ws), g1, scale_out, zp_out,
                 static_cast<unsigned long long*>(nullptr), 0) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return MIXDQ_OK;
}

// pass 2 alone on a dense tensor: the producer's CTAs stored `nparts` partials into ws->partial
int mixdq_q2_premm(const __half* x, int64_t numel, int8_t* q, float* scale_out, float* zp_out,
                   void* ws, int nparts, unsigned long long* zero_words, int zero_n,
                   cudaStream_t st) {
  const int64_t items = numel >> 3;
  if (items > kMaxItems || nparts < 1 || nparts > kMaxPartials) return MIXDQ_ERR_UNSUPPORTED;
  const unsigned int n = static_cast<unsigned int>(items);
  const int g2 = grid_for2(items, kQ2Threads, 148 * 8);
  if (launch_pdl(quant_rows_premm_kernel, g2, kQ2Threads, 0, st, x, static_cast<int64_t>(0), n, n,
                 q, static_cast<DynWs*>(ws), nparts, scale_out, zp_out, zero_words, zero_n) !=
      cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return MIXDQ_OK;
}

// LayerNorm -> fp16 y (caller's buffer) + per-CTA min/max, then pass 2 on y
int mixdq_q2_ln(const __half* x, int64_t ldx, int M, int C, const __half* gamma,
                const __half* beta, float eps, int8_t* q, __half* y, float* scale_out,
                float* zp_out, void* ws, cudaStream_t st) {
  if (static_cast<int64_t>(M) * (C >> 3) > kMaxItems) return MIXDQ_ERR_UNSUPPORTED;
  DynWs* w = static_cast<DynWs*>(ws);
  cudaError_t e;
  int g1;
  // one-cluster form: one row per warp, 16 warps per CTA
  if (mixdq_quant_mode() == 3) {
    static int ncl = -1;
    if (ncl < 0) {
      const int a = max_cluster_ctas(ln_quant_cluster_kernel<5>, kClThreads, 0);
      const int b = max_cluster_ctas(ln_quant_cluster_kernel<8>, kClThreads, 0);
      ncl = a < b ? a : b;
    }
    if (ncl > 0 && M <= ncl * (kClThreads / 32)) {
      const int g = (M + kClThreads / 32 - 1) / (kClThreads / 32);
      e = (C <= 5 * 256)
              ? launch_cluster_pdl(ln_quant_cluster_kernel<5>, g, kClThreads, 0, st, x, ldx, M, C,
                                   gamma, beta, eps, q, y, scale_out, zp_out)
              : launch_cluster_pdl(ln_quant_cluster_kernel<8>, g, kClThreads, 0, st, x, ldx, M, C,
                                   gamma, beta, eps, q, y, scale_out, zp_out);
      return e == cudaSuccess ? MIXDQ_OK : MIXDQ_ERR_CUDA;
    }
  }
  // one row per warp, the whole tensor in registers, lean barrier: M <= 148 CTAs x 8 warps
  if (mixdq_quant_mode() == 2 && M <= 148 * 8) {
    if (M <= 296) {
      g1 = (M + 1) / 2;
      e = (C <= 5 * 256)
              ? launch_pdl(ln_quant_lean_kernel<5, 64>, g1, 64, 0, st, x, ldx, M, C, gamma, beta, eps,
                           q, y, w, scale_out, zp_out)
              : launch_pdl(ln_quant_lean_kernel<8, 64>, g1, 64, 0, st, x, ldx, M, C, gamma, beta, eps,
                           q, y, w, scale_out, zp_out);
    } else {
      g1 = (M + 7) / 8;
      e = (C <= 5 * 256)
              ? launch_pdl(ln_quant_lean_kernel<5, 256>, g1, 256, 0, st, x, ldx, M, C, gamma, beta,
                           eps, q, y, w, scale_out, zp_out)
              : launch_pdl(ln_quant_lean_kernel<8, 256>, g1, 256, 0, st, x, ldx, M, C, gamma, beta,
                           eps, q, y, w, scale_out, zp_out);
    }
    return e == cudaSuccess ? MIXDQ_OK : MIXDQ_ERR_CUDA;   // (y, if given, was written too)
  }
  if (y == nullptr) return MIXDQ_ERR_UNSUPPORTED;          // the two-pass form needs the scratch
  if (M <= 512) {
    g1 = (M + 1) / 2;
    e = (C <= 5 * 256)
            ? launch_pdl(ln_minmax_kernel<5, 64>, g1, 64, 0, st, x, ldx, M, C, gamma, beta, eps, y, w)
            : launch_pdl(ln_minmax_kernel<8, 64>, g1, 64, 0, st, x, ldx, M, C, gamma, beta, eps, y, w);
  } else {
    g1 = (M + 7) / 8;
    if (g1 > 512) g1 = 512;
    e = (C <= 5 * 256)
            ? launch_pdl(ln_minmax_kernel<5, 256>, g1, 256, 0, st, x, ldx, M, C, gamma, beta, eps, y, w)
            : launch_pdl(ln_minmax_kernel<8, 256>, g1, 256, 0, st, x, ldx, M, C, gamma, beta, eps, y, w);
  }
  if (e != cudaSuccess) return MIXDQ_ERR_CUDA;
  return mixdq_q2_premm(y, static_cast<int64_t>(M) * C, q, scale_out, zp_out, ws, g1, nullptr, 0, st);
}
