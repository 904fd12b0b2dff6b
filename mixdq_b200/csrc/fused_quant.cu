// fused_quant.cu — producer-side fusion of the dynamic activation-quantise pass:
// the op that PRODUCES a quantized layer's input and the qdiff min-max quantisation of its result
// run as ONE kernel, so the fp16 intermediate never travels to HBM and back:
//
//   LayerNorm            -> int8   feeds attn1.to_q/k/v, attn2.to_q, ff.net.0.proj
//   GEGLU  h * gelu(g)   -> int8   feeds ff.net.2
//   GroupNorm [+ SiLU]   -> int8   feeds resnet conv1 / conv2 (NHWC) and Transformer2D.proj_in
//
// Arithmetic contract. Each kernel reproduces the UNFUSED sequence the reference runs (stock
// PyTorch fp16 op(s), then quantisation): the op is evaluated in fp32 and ROUNDED TO FP16 at
// every point where PyTorch materialises an fp16 tensor (LayerNorm output; GroupNorm output, SiLU
// output; GELU output, the h*gelu product); min/max and the codes are then taken from those fp16
// values with the qdiff formula in fp32 (quant_ws.cuh) — bit-exact given the fp16 values. The
// fp32 op itself (mean/variance summation order) is a floating-point restatement, compared with
// a stated tolerance in tests/.
//
// Structure of every kernel: phase 1 computes the fp16 values of the CTA's rows into a shared
// memory stash (up to ~200 KB per CTA) while tracking min/max; a grid barrier (all CTAs
// co-resident: grid <= 148, one CTA per SM) turns the partial min/max into (delta, z); phase 2
// quantises from the stash with 8-byte coalesced stores. Rows that do not fit the stash are
// recomputed in phase 2 (their inputs are L2-resident by then). GroupNorm has one more barrier in
// front for the per-(image, group) statistics, accumulated as fixed-point integers so the result
// does not depend on the arrival order of the CTAs.
#include "common.cuh"
#include "quant_ws.cuh"
#include "../../include/mixdq_b200.h"

namespace mixdq {

constexpr int kFqThreads = 512;
constexpr int kFqWarps = kFqThreads / 32;
constexpr int kFqMaxSmem = 200 * 1024;     // stash budget per CTA
constexpr int kNumSm = 148;

__device__ __forceinline__ int4 ldg16(const void* p) {
  return __ldg(reinterpret_cast<const int4*>(p));
}

__device__ __forceinline__ void unpack8(const int4& raw, float (&f)[8]) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h2[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ int4 pack8(const float (&f)[8]) {
  int4 r;
  __half2* h2 = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h2[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return r;
}

// =============================================================================================
// LayerNorm -> int8
// =============================================================================================
constexpr int kLnMaxChunks = 8;   // 16-byte chunks per lane: C <= 8 * 32 * 8 = 2048
                                  // (instantiated for 5 -> C <= 1280, every UNet here, and 8)

// y[row] (8 halves per chunk, chunks lane, lane+32, ...) of one row held by one warp.
// PyTorch: out = half(gamma * (rstd * (x - mean)) + beta), statistics in fp32, biased variance.
template <int MAXCH>
__device__ __forceinline__ void ln_row(const __half* __restrict__ xrow, int nchunks, int C,
                                       const __half* __restrict__ gamma,
                                       const __half* __restrict__ beta, float eps, int lane,
                                       int4 (&y)[MAXCH]) {
  float v[MAXCH][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      unpack8(ldg16(xrow + 8 * c), v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; ss += d * d; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / static_cast<float>(C) + eps);
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      float g[8], b[8], o[8];
      unpack8(ldg16(gamma + 8 * c), g);
      unpack8(ldg16(beta + 8 * c), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(g[j], rstd * (v[i][j] - mean), b[j]);
      y[i] = pack8(o);
    }
  }
}

// CLUSTER (all three row kernels below): the launch is ONE thread-block cluster whose stashes
// hold the whole tensor; min/max through DSMEM + the cluster barrier (quant_ws.cuh) and a
// programmatic dependency on the producer instead of a full launch gap.
template <int MAXCH, bool CLUSTER>
__global__ void __launch_bounds__(kFqThreads, 1)
ln_quant_kernel(const __half* __restrict__ x, int64_t ldx, int M, int C,
                const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps,
                int8_t* __restrict__ q, __half* __restrict__ y_out, DynWs* __restrict__ ws,
                float* __restrict__ scale_out, float* __restrict__ zp_out, int rows_per_cta,
                int stash_rows) {
  extern __shared__ int4 stash[];   // [stash_rows][C/8]
  QDbg dbg;
  dbg.begin(ws);
  pdl_launch_dependents();
  if (CLUSTER) cluster_enter();
  pdl_wait();
  dbg.waited(ws);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = C >> 3;
  const int row0 = blockIdx.x * rows_per_cta;
  const int row1 = min(M, row0 + rows_per_cta);
  float mn = 0.f, mx = 0.f;
  for (int r = row0 + warp; r < row1; r += kFqWarps) {
    int4 y[MAXCH];
    ln_row<MAXCH>(x + static_cast<int64_t>(r) * ldx, nchunks, C, gamma, beta, eps, lane, y);
    const int lr = r - row0;
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        minmax_vec8(y[i], mn, mx);
        if (lr < stash_rows) stash[lr * nchunks + c] = y[i];
        if (y_out) reinterpret_cast<int4*>(y_out + static_cast<int64_t>(r) * C)[c] = y[i];
      }
    }
  }
  float delta, z;
  dbg.stamp(2);
  if (CLUSTER) cluster_minmax_params<kFqThreads>(mn, mx, scale_out, zp_out, delta, z);
  else grid_minmax_params<kFqThreads>(ws, mn, mx, scale_out, zp_out, delta, z);
  dbg.stamp(3);
  for (int r = row0 + warp; r < row1; r += kFqWarps) {
    const int lr = r - row0;
    uint2* qrow = reinterpret_cast<uint2*>(q + static_cast<int64_t>(r) * C);
    if (lr < stash_rows) {
      for (int c = lane; c < nchunks; c += 32) qrow[c] = qdiff_vec8(stash[lr * nchunks + c], delta, z);
    } else {
      int4 y[MAXCH];
      ln_row<MAXCH>(x + static_cast<int64_t>(r) * ldx, nchunks, C, gamma, beta, eps, lane, y);
#pragma unroll
      for (int i = 0; i < MAXCH; ++i) {
        const int c = lane + 32 * i;
        if (c < nchunks) qrow[c] = qdiff_vec8(y[i], delta, z);
      }
    }
  }
  dbg.end(ws);
}

// =============================================================================================
// GEGLU -> int8      hg = [M][2*I] fp16 (first half h, second half gate)
// PyTorch: gelu = half(0.5 * g * (1 + erf(g / sqrt(2))))   (F.gelu, approximate='none', fp32 math)
//          y    = half(float(h) * float(gelu))
// =============================================================================================
__device__ __forceinline__ int4 geglu_vec8(const int4& hraw, const int4& graw) {
  float h[8], g[8], o[8];
  unpack8(hraw, h);
  unpack8(graw, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = geglu_f32(h[j], g[j]);
  return pack8(o);
}

__global__ void __launch_bounds__(kFqThreads, 1)
geglu_quant_kernel(const __half* __restrict__ hg, int64_t ld, int M, int I,
                   int8_t* __restrict__ q, __half* __restrict__ y_out, DynWs* __restrict__ ws,
                   float* __restrict__ scale_out, float* __restrict__ zp_out, int rows_per_cta,
                   int stash_rows) {
  extern __shared__ int4 stash[];   // [stash_rows][I/8]
  pdl_launch_dependents();
  pdl_wait();
  const int nchunks = I >> 3;
  const int row0 = blockIdx.x * rows_per_cta;
  const int row1 = min(M, row0 + rows_per_cta);
  const int64_t items = static_cast<int64_t>(row1 - row0) * nchunks;
  float mn = 0.f, mx = 0.f;
  for (int64_t it = threadIdx.x; it < items; it += kFqThreads) {
    const int lr = static_cast<int>(it / nchunks);
    const int c = static_cast<int>(it - static_cast<int64_t>(lr) * nchunks);
    const __half* row = hg + static_cast<int64_t>(row0 + lr) * ld;
    const int4 y = geglu_vec8(ldg16(row + 8 * c), ldg16(row + I + 8 * c));
    minmax_vec8(y, mn, mx);
    if (lr < stash_rows) stash[it] = y;
    if (y_out) reinterpret_cast<int4*>(y_out + static_cast<int64_t>(row0 + lr) * I)[c] = y;
  }
  float delta, z;
  grid_minmax_params<kFqThreads>(ws, mn, mx, scale_out, zp_out, delta, z);
  for (int64_t it = threadIdx.x; it < items; it += kFqThreads) {
    const int lr = static_cast<int>(it / nchunks);
    const int c = static_cast<int>(it - static_cast<int64_t>(lr) * nchunks);
    int4 y;
    if (lr < stash_rows) {
      y = stash[it];
    } else {
      const __half* row = hg + static_cast<int64_t>(row0 + lr) * ld;
      y = geglu_vec8(ldg16(row + 8 * c), ldg16(row + I + 8 * c));
    }
    reinterpret_cast<uint2*>(q + static_cast<int64_t>(row0 + lr) * I)[c] = qdiff_vec8(y, delta, z);
  }
}

// =============================================================================================
// plain dynamic quantisation of a row-pitched view [M][cols] (pitch ldx) -> dense int8 [M][cols]:
// channel slices of NHWC tensors (split shortcuts), attention outputs, token slices.
// =============================================================================================
template <bool CLUSTER>
__global__ void __launch_bounds__(kFqThreads, 1)
rows_quant_kernel(const __half* __restrict__ x, int64_t ldx, int M, int cols,
                  int8_t* __restrict__ q, DynWs* __restrict__ ws, float* __restrict__ scale_out,
                  float* __restrict__ zp_out, int rows_per_cta, int stash_rows) {
  extern __shared__ int4 stash[];   // [stash_rows][cols/8]
  pdl_launch_dependents();
  if (CLUSTER) cluster_enter();
  pdl_wait();
  const int nchunks = cols >> 3;
  const int row0 = blockIdx.x * rows_per_cta;
  const int row1 = min(M, row0 + rows_per_cta);
  const int64_t items = static_cast<int64_t>(row1 - row0) * nchunks;
  float mn = 0.f, mx = 0.f;
  for (int64_t it = threadIdx.x; it < items; it += kFqThreads) {
    const int lr = static_cast<int>(it / nchunks);
    const int c = static_cast<int>(it - static_cast<int64_t>(lr) * nchunks);
    const int4 y = ldg16(x + static_cast<int64_t>(row0 + lr) * ldx + 8 * c);
    minmax_vec8(y, mn, mx);
    if (lr < stash_rows) stash[it] = y;
  }
  float delta, z;
  if (CLUSTER) cluster_minmax_params<kFqThreads>(mn, mx, scale_out, zp_out, delta, z);
  else grid_minmax_params<kFqThreads>(ws, mn, mx, scale_out, zp_out, delta, z);
  for (int64_t it = threadIdx.x; it < items; it += kFqThreads) {
    const int lr = static_cast<int>(it / nchunks);
    const int c = static_cast<int>(it - static_cast<int64_t>(lr) * nchunks);
    const int4 y = (lr < stash_rows) ? stash[it]
                                     : ldg16(x + static_cast<int64_t>(row0 + lr) * ldx + 8 * c);
    reinterpret_cast<uint2*>(q + static_cast<int64_t>(row0 + lr) * cols)[c] = qdiff_vec8(y, delta, z);
  }
}

// =============================================================================================
// GroupNorm [+ SiLU] -> int8, NHWC:  x = [NB][HW][C] fp16 (pixel pitch ldx), G groups of C/G
// consecutive channels.
// PyTorch: mean/rstd per (n, group) in fp32; y = half(fma(x, a, b)), a = rstd*gamma[c],
//          b = beta[c] - mean*a; SiLU: half(y / (1 + exp(-y))) on the fp16 y.
// Each CTA works on rows (pixels) of ONE image; a lane owns the same channel chunks in every row,
// so its per-group partial sums live in registers across the CTA's rows.
// =============================================================================================
constexpr int kGnMaxChunks = 10;  // chunks per lane: C <= 10 * 32 * 8 = 2560
constexpr double kFixSum = 16777216.0;   // 2^24
constexpr double kFixSq = 4096.0;        // 2^12

__device__ __forceinline__ float silu_half(float y) {
  const float yh = __half2float(__float2half_rn(y));          // GroupNorm output tensor (fp16)
  return yh / (1.0f + expf(-yh));
}

template <bool SILU>
__device__ __forceinline__ int4 gn_vec8(const int4& raw, int c8, int cpg,
                                        const float* __restrict__ s_mean,
                                        const float* __restrict__ s_rstd,
                                        const __half* __restrict__ gamma,
                                        const __half* __restrict__ beta) {
  float v[8], g[8], b[8], o[8];
  unpack8(raw, v);
  unpack8(ldg16(gamma + c8), g);
  unpack8(ldg16(beta + c8), b);
  const int g0 = c8 / cpg;
  const int split = (g0 + 1) * cpg - c8;      // elements [0, split) belong to group g0
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int gi = j < split ? g0 : g0 + 1;
    const float a = s_rstd[gi] * g[j];
    const float bb = fmaf(-s_mean[gi], a, b[j]);
    const float y = fmaf(v[j], a, bb);
    o[j] = SILU ? silu_half(y) : y;
  }
  return pack8(o);
}

// ---- GroupNorm statistics --------------------------------------------------------------------
// An image is cut into UNITS of gn_unit_rows(C) consecutive pixel rows (4 / 2 / 1 for C <= 512 /
// 1280 / 2560: the same units whatever the batch, the grid or the kernel form). One WARP reduces one
// unit with ONE round of loads — every lane owns the 16-byte channel chunks lane, lane + 32, ...
// and adds the unit's rows in ascending order (fp32) — then the per-chunk partials meet in a
// warp-private shared-memory row and one lane per (group, quantity) adds the chunks of its group in
// ascending order; that fp32 unit partial is converted to FIXED POINT. From there on everything is
// integer addition — exact and order-independent — so a warp sums the units it walks in
// registers, the CTA adds its warps up in shared memory and issues one atomic per (image, group,
// quantity): the statistics of an image are bit-identical for every batch size, grid and
// unit-to-warp assignment (static-scale UNets are invariant under batch sharding,
// mixdq_b200/dp.py). RB x CPL = 8-10 sixteen-byte loads are in flight per lane (the first form had
// one or two and ran at ~1 TB/s on batch-64 tensors); at batch 1 a warp has one unit.
__host__ __device__ inline int gn_unit_rows(int nchunks) {
  return nchunks <= 64 ? 4 : (nchunks <= 160 ? 2 : 1);
}

// per-lane constants of the (group, quantity) reduction: no division inside the unit loop
struct GnPair {
  int c_lo, c_hi;     // chunks overlapping the group
  bool live, sq, first_hi;   // first_hi: chunk c_lo starts in the previous group -> take its high part
};
__device__ __forceinline__ GnPair gn_pair(int pq, int cpg, int G) {
  GnPair p;
  const int g = pq >> 1;
  p.live = pq < 2 * G;
  p.sq = (pq & 1) != 0;
  p.c_lo = (g * cpg) >> 3;
  p.c_hi = ((g + 1) * cpg - 1) >> 3;
  p.first_hi = 8 * p.c_lo < g * cpg;
  return p;
}

template <int CPL, int RB>
__device__ __forceinline__ void gn_unit_stats(const __half* __restrict__ ximg, int64_t ldx, int r0,
                                              int r1, int nchunks, int cpg,
                                              float4* __restrict__ wpart, int lane,
                                              const GnPair (&pr)[2], long long (&fx)[2]) {
  int4 raw[RB][CPL];
#pragma unroll
  for (int u = 0; u < RB; ++u)
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = lane + 32 * i;
      if (r0 + u < r1 && c < nchunks)
        raw[u][i] = ldg16(ximg + static_cast<int64_t>(r0 + u) * ldx + 8 * c);
    }
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      const int c8 = 8 * c;
      const int split = (c8 / cpg + 1) * cpg - c8;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        if (r0 + u < r1) {
          float v[8];
          unpack8(raw[u][i], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < split) { a0 += v[j]; a1 = fmaf(v[j], v[j], a1); }
            else           { a2 += v[j]; a3 = fmaf(v[j], v[j], a3); }
          }
        }
      }
      wpart[c] = make_float4(a0, a1, a2, a3);
    }
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (pr[h].live) {
      const float4 f = wpart[pr[h].c_lo];
      float t = pr[h].first_hi ? (pr[h].sq ? f.w : f.z) : (pr[h].sq ? f.y : f.x);
      for (int c = pr[h].c_lo + 1; c <= pr[h].c_hi; ++c) {
        const float4 w4 = wpart[c];
        t += pr[h].sq ? w4.y : w4.x;
      }
      // power-of-two scaling is exact in fp32: same value as the fp64 product
      fx[h] += __float2ll_rn(t * (pr[h].sq ? static_cast<float>(kFixSq) : static_cast<float>(kFixSum)));
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void gn_unit_stats_any(const __half* __restrict__ ximg, int64_t ldx,
                                                  int r0, int r1, int nchunks, int cpg,
                                                  float4* __restrict__ wpart, int lane,
                                                  const GnPair (&pr)[2], long long (&fx)[2]) {
  if (nchunks <= 64) gn_unit_stats<2, 4>(ximg, ldx, r0, r1, nchunks, cpg, wpart, lane, pr, fx);
  else if (nchunks <= 160) gn_unit_stats<5, 2>(ximg, ldx, r0, r1, nchunks, cpg, wpart, lane, pr, fx);
  else gn_unit_stats<kGnMaxChunks, 1>(ximg, ldx, r0, r1, nchunks, cpg, wpart, lane, pr, fx);
}

// A warp's fixed-point sums for image n: into ITS row of the CTA's shared-memory table when the
// image is one of the kGnSlots the CTA's contiguous range starts with (plain 64-bit adds, no
// atomics), else straight to the workspace. Thousands of warps adding to the 2 x G words of ONE
// image serialise in L2, so the CTA adds its warps up first (gn_publish_slots).
constexpr int kGnSlots = 2;
__device__ __forceinline__ void gn_flush_stats(DynWs* __restrict__ ws, unsigned long long* s_fx,
                                               int n_base, int n, int G, int warp, int lane,
                                               long long (&fx)[2]) {
  const int slot = n - n_base;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int pq = lane + 32 * h;
    if (pq < 2 * G && fx[h] != 0) {
      if (slot >= 0 && slot < kGnSlots)
        s_fx[(warp * kGnSlots + slot) * 64 + pq] += static_cast<unsigned long long>(fx[h]);
      else
        atomicAdd(&ws->gsum[(n * G) * 2 + pq], static_cast<unsigned long long>(fx[h]));
    }
    fx[h] = 0;
  }
}
// after a __syncthreads(): the CTA's table -> one atomic per (image, group, quantity)
__device__ __forceinline__ void gn_publish_slots(DynWs* __restrict__ ws,
                                                 const unsigned long long* s_fx, int n_base, int NB,
                                                 int G, int warps) {
  for (int i = threadIdx.x; i < kGnSlots * 64; i += blockDim.x) {
    const int slot = i >> 6, pq = i & 63;
    unsigned long long v = 0ull;
    for (int w = 0; w < warps; ++w) v += s_fx[(w * kGnSlots + slot) * 64 + pq];
    if (v != 0ull && pq < 2 * G && n_base + slot < NB)
      atomicAdd(&ws->gsum[((n_base + slot) * G) * 2 + pq], v);
  }
}

// statistics kernel of the barrier-free form: every warp of the launch walks a CONTIGUOUS range of
// the NB x units_per_image units; blockDim.x / 32 warps per CTA — few units are spread over many
// CTAs of few warps, so that at batch 1 every SM pulls a few KB instead of 8 SMs pulling 80 KB
// (dynamic shared memory: warps x nchunks float4)
template <int CPL, int RB>
__global__ void __launch_bounds__(kFqThreads, 1)
gn_stats_kernel(const __half* __restrict__ x, int64_t ldx, int NB, int HW, int C, int G,
                DynWs* __restrict__ ws, int units_per_warp) {
  extern __shared__ float4 gn_part[];
  __shared__ unsigned long long s_fx[kFqWarps * kGnSlots * 64];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int nchunks = C >> 3, cpg = C / G;
  const int ur = gn_unit_rows(nchunks);
  const int upi = (HW + ur - 1) / ur;
  const long long total = static_cast<long long>(NB) * upi;
  const long long cta_u0 = static_cast<long long>(blockIdx.x) * wpc * units_per_warp;
  const int n_base = static_cast<int>(cta_u0 / upi);
  const GnPair pr[2] = {gn_pair(lane, cpg, G), gn_pair(lane + 32, cpg, G)};
  for (int i = threadIdx.x; i < wpc * kGnSlots * 64; i += blockDim.x) s_fx[i] = 0ull;
  __syncthreads();
  pdl_wait();
  long long u = cta_u0 + static_cast<long long>(warp) * units_per_warp;
  const long long u_end = u + units_per_warp < total ? u + units_per_warp : total;
  float4* wpart = gn_part + warp * nchunks;
  long long fx[2] = {0, 0};
  int cur_n = -1;
  for (; u < u_end; ++u) {
    const int n = static_cast<int>(u / upi);
    const int r0 = static_cast<int>(u - static_cast<long long>(n) * upi) * ur;
    if (n != cur_n) {
      if (cur_n >= 0) gn_flush_stats(ws, s_fx, n_base, cur_n, G, warp, lane, fx);
      cur_n = n;
    }
    gn_unit_stats<CPL, RB>(x + static_cast<int64_t>(n) * HW * ldx, ldx, r0, min(HW, r0 + ur), nchunks,
                           cpg, wpart, lane, pr, fx);
  }
  if (cur_n >= 0) gn_flush_stats(ws, s_fx, n_base, cur_n, G, warp, lane, fx);
  __syncthreads();
  gn_publish_slots(ws, s_fx, n_base, NB, G, wpc);
}

// MODE 0: everything in one kernel (two grid barriers).  MODE 2: the apply kernel of the
// barrier-free form (gn_stats_kernel -> this: normalise -> fp16 y + min/max published for
// quant2.cu's single-pass quantiser, or int8 with static scales), chained by programmatic
// dependent launch.
template <bool SILU, int MODE>
__global__ void __launch_bounds__(kFqThreads, 1)
gn_quant_kernel(const __half* __restrict__ x, int64_t ldx, int NB, int HW, int C, int G,
                const __half* __restrict__ gamma, const __half* __restrict__ beta, float eps,
                int8_t* __restrict__ q, __half* __restrict__ y_out, DynWs* __restrict__ ws,
                float* __restrict__ scale_out, float* __restrict__ zp_out, int ctas_per_image,
                int rows_per_cta, int stash_rows, int8_t* __restrict__ qs,
                const float* __restrict__ s_inv, const float* __restrict__ s_zp) {
  // qs != nullptr (MODE 2 only): STATIC activation scales — the apply pass quantises with the
  // consumer's checkpoint parameters instead of writing fp16 + min/max for a quantise pass
  extern __shared__ int4 stash[];   // [stash_rows][C/8]
  __shared__ float s_mean[32], s_rstd[32];
  __shared__ int s_last;
  __shared__ unsigned long long s_fx0[(MODE == 0 ? kFqWarps * kGnSlots * 64 : 1)];   // MODE 0 statistics
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = C >> 3;
  const int cpg = C / G;
  {
  const int vb = blockIdx.x;
  const int n = vb / ctas_per_image;
  const int row0 = (vb - n * ctas_per_image) * rows_per_cta;
  const int row1 = min(HW, row0 + rows_per_cta);
  const __half* ximg = x + static_cast<int64_t>(n) * HW * ldx;

  if (MODE != 2) {
  // ---- phase 0: per-(n, group) sum / sum of squares (MODE 0 only; the barrier-free form runs
  //      gn_stats_kernel): this CTA's rows = whole units (rows_per_cta is a multiple of 4) ----
  {
    const int ur = gn_unit_rows(nchunks);
    float4* wpart = reinterpret_cast<float4*>(stash) + warp * nchunks;   // stash not needed yet
    const GnPair pr[2] = {gn_pair(lane, cpg, G), gn_pair(lane + 32, cpg, G)};
    for (int i = threadIdx.x; i < kFqWarps * kGnSlots * 64; i += kFqThreads) s_fx0[i] = 0ull;
    __syncthreads();
    long long fx[2] = {0, 0};
    for (int r0 = row0 + warp * ur; r0 < row1; r0 += kFqWarps * ur)
      gn_unit_stats_any(ximg, ldx, r0, min(row1, r0 + ur), nchunks, cpg, wpart, lane, pr, fx);
    gn_flush_stats(ws, s_fx0, n, n, G, warp, lane, fx);
    __syncthreads();
    gn_publish_slots(ws, s_fx0, n, NB, G, kFqWarps);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&ws->counter2, 1u);
    unsigned int spins = 0;
    while (ld_acquire_u32(&ws->counter2) < gridDim.x) {
      __nanosleep(20);
      if (++spins > (1u << 24)) __trap();
    }
  }
  __syncthreads();
  }  // MODE != 2
  if (threadIdx.x < G) {
    const long long fs = static_cast<long long>(__ldcg(&ws->gsum[(n * G + threadIdx.x) * 2]));
    const long long fq = static_cast<long long>(__ldcg(&ws->gsum[(n * G + threadIdx.x) * 2 + 1]));
    const double cnt = static_cast<double>(HW) * cpg;
    const double mean = static_cast<double>(fs) / kFixSum / cnt;
    double var = static_cast<double>(fq) / kFixSq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = rsqrtf(static_cast<float>(var) + eps);
  }
  __syncthreads();
  // the last CTA to have read the statistics re-zeroes them for the next call (MODE 2 with
  // dynamic scales: the quantise pass that follows clears them instead — nobody waits here)
  const bool rezero = (MODE != 2) || (qs != nullptr);   // static apply pass: no quantise pass follows
  if (!rezero) s_last = 0;
  if (rezero && threadIdx.x == 0) s_last = (atomicAdd(&ws->done2, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (rezero && s_last) {
    for (int i = threadIdx.x; i < NB * G * 2; i += kFqThreads) ws->gsum[i] = 0ull;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      ws->done2 = 0;
      __threadfence();
      if (MODE == 0) st_release_u32(&ws->counter2, 0u);
    }
  }

  // ---- phase 1: normalise (+SiLU) -> fp16 -> stash, min/max ----
  float mn = 0.f, mx = 0.f;
  const float q_inv = (MODE == 2 && qs) ? __ldg(s_inv) : 0.f, q_zp = (MODE == 2 && qs) ? __ldg(s_zp) : 0.f;
  {
    // flat (row, chunk) items of this CTA, four 16-byte loads in flight per thread (a warp-per-row
    // loop left each lane with one or two dependent load -> compute -> store chains per row)
    const int items = (row1 - row0) * nchunks;
    constexpr int U = 4;
    for (int it0 = threadIdx.x; it0 < items; it0 += U * kFqThreads) {
      int4 raw[U];
      int lr[U], c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int it = it0 + u * kFqThreads;
        lr[u] = it / nchunks;
        c[u] = it - lr[u] * nchunks;
        if (it < items)
          raw[u] = ldg16(ximg + static_cast<int64_t>(row0 + lr[u]) * ldx + 8 * c[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int it = it0 + u * kFqThreads;
        if (it >= items) break;
        const int4 y = gn_vec8<SILU>(raw[u], 8 * c[u], cpg, s_mean, s_rstd, gamma, beta);
        if (MODE == 2 && qs != nullptr) {
          reinterpret_cast<uint2*>(qs + (static_cast<int64_t>(n) * HW + row0 + lr[u]) * C)[c[u]] =
              static_quant8(y, q_inv, q_zp);
          continue;
        }
        minmax_vec8(y, mn, mx);
        if (MODE == 0 && lr[u] < stash_rows) stash[it] = y;
        if (y_out)
          reinterpret_cast<int4*>(y_out + (static_cast<int64_t>(n) * HW + row0 + lr[u]) * C)[c[u]] = y;
      }
    }
  }
  if (MODE == 2 && qs != nullptr) return;
  if (MODE == 2) {
    // store this CTA's min / max partial (quant2.cu protocol) and stop: quantisation is the next
    // kernel's job
    __shared__ float s_mn[kFqWarps], s_mx[kFqWarps];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kFqWarps; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
      ws->partial[blockIdx.x] = make_float2(mn, mx);
    }
    return;
  }
  float delta, z;
  grid_minmax_params<kFqThreads>(ws, mn, mx, scale_out, zp_out, delta, z);
  // ---- phase 2: quantise ----
  for (int r = row0 + warp; r < row1; r += kFqWarps) {
    const int lr = r - row0;
    uint2* qrow = reinterpret_cast<uint2*>(q + (static_cast<int64_t>(n) * HW + r) * C);
    const __half* xrow = ximg + static_cast<int64_t>(r) * ldx;
    for (int c = lane; c < nchunks; c += 32) {
      const int4 y = (lr < stash_rows)
                         ? stash[lr * nchunks + c]
                         : gn_vec8<SILU>(ldg16(xrow + 8 * c), 8 * c, cpg, s_mean, s_rstd, gamma, beta);
      qrow[c] = qdiff_vec8(y, delta, z);
    }
  }
  }
}

template <typename K>
static int set_smem(K kern, int bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) ==
                 cudaSuccess
             ? 0
             : -1;
}

}  // namespace mixdq

using namespace mixdq;

// quant2.cu / quant.cu
int mixdq_q2_rows(const __half* x, int64_t ldx, int64_t M, int cols, int8_t* q, float* scale_out,
                  float* zp_out, void* ws, cudaStream_t st, int n_bits = 8);
int mixdq_q2_premm(const __half* x, int64_t numel, int8_t* q, float* scale_out, float* zp_out,
                   void* ws, int nparts, unsigned long long* zero_words, int zero_n,
                   cudaStream_t st);
int mixdq_q2_ln(const __half* x, int64_t ldx, int M, int C, const __half* gamma,
                const __half* beta, float eps, int8_t* q, __half* y, float* scale_out,
                float* zp_out, void* ws, cudaStream_t st, int8_t* qs = nullptr,
                const float* s_inv = nullptr, const float* s_zp = nullptr);
bool mixdq_two_pass_enabled();

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#define MIXDQ_CHECK_LAUNCH()                                   \
  do {                                                         \
    if (cudaGetLastError() != cudaSuccess) return MIXDQ_ERR_CUDA; \
  } while (0)

// rows of `row_bytes` fp16-stash bytes each, distributed over <= 148 co-resident CTAs
static void plan_rows(int64_t rows, int64_t row_bytes, int min_rows_per_cta, int* grid,
                      int* rows_per_cta, int* stash_rows, int* smem) {
  int64_t g = (rows + min_rows_per_cta - 1) / min_rows_per_cta;
  if (g > kNumSm) g = kNumSm;
  if (g < 1) g = 1;
  const int64_t rpc = (rows + g - 1) / g;
  g = (rows + rpc - 1) / rpc;
  int64_t srows = kFqMaxSmem / row_bytes;
  if (srows > rpc) srows = rpc;
  *grid = static_cast<int>(g);
  *rows_per_cta = static_cast<int>(rpc);
  *stash_rows = static_cast<int>(srows);
  *smem = static_cast<int>(srows * row_bytes);
}

// one cluster of <= ncl CTAs whose stashes hold ALL rows; false if the tensor does not fit
static bool plan_rows_cluster(int64_t rows, int64_t row_bytes, int min_rows_per_cta, int ncl,
                              int* grid, int* rows_per_cta, int* smem) {
  if (ncl <= 0 || !cluster_enabled() || rows * row_bytes > 2 * kClusterMaxElems) return false;
  int64_t g = (rows + min_rows_per_cta - 1) / min_rows_per_cta;
  if (g > ncl) g = ncl;
  if (g < 1) g = 1;
  const int64_t rpc = (rows + g - 1) / g;
  g = (rows + rpc - 1) / rpc;
  if (rpc * row_bytes > kFqMaxSmem) return false;
  *grid = static_cast<int>(g);
  *rows_per_cta = static_cast<int>(rpc);
  *smem = static_cast<int>(rpc * row_bytes);
  return true;
}

extern "C" int mixdq_ln_quant_i8_dynamic(const mixdq_half_t* x, int64_t ldx, int M, int C,
                                         const mixdq_half_t* gamma, const mixdq_half_t* beta,
                                         float eps, int8_t* q, mixdq_half_t* y_out,
                                         float* scale_out, float* zp_out, void* ws,
                                         mixdq_stream_t stream) {
  // q == nullptr (with y_out): LayerNorm only, the caller quantises y with static parameters
  if (M <= 0 || C <= 0 || !x || !gamma || !beta || !ws || ldx < C ||
      (q ? (!scale_out || !zp_out) : !y_out))
    return MIXDQ_ERR_INVALID_ARG;
  if ((C & 7) || (ldx & 7) || !al16(x) || !al16(gamma) || !al16(beta) || (q && !al16(q)) ||
      (y_out && !al16(y_out)))
    return MIXDQ_ERR_ALIGNMENT;
  if (C > kLnMaxChunks * 256) return MIXDQ_ERR_UNSUPPORTED;
  if (q == nullptr)
    return mixdq_q2_ln(reinterpret_cast<const __half*>(x), ldx, M, C,
                       reinterpret_cast<const __half*>(gamma),
                       reinterpret_cast<const __half*>(beta), eps, nullptr,
                       reinterpret_cast<__half*>(y_out), nullptr, nullptr, ws,
                       static_cast<cudaStream_t>(stream));
  static bool attr = false;
  if (!attr) {
    if (set_smem(ln_quant_kernel<5, false>, kFqMaxSmem) ||
        set_smem(ln_quant_kernel<8, false>, kFqMaxSmem) ||
        set_smem(ln_quant_kernel<5, true>, kFqMaxSmem) ||
        set_smem(ln_quant_kernel<8, true>, kFqMaxSmem))
      return MIXDQ_ERR_CUDA;
    attr = true;
  }
  int grid, rpc, srows, smem;
  static int ncl = -1;
  if (ncl < 0) {
    const int a = max_cluster_ctas(ln_quant_kernel<5, true>, kFqThreads, kFqMaxSmem);
    const int b = max_cluster_ctas(ln_quant_kernel<8, true>, kFqThreads, kFqMaxSmem);
    ncl = a < b ? a : b;
  }
  if (plan_rows_cluster(M, static_cast<int64_t>(C) * 2, 1, ncl, &grid, &rpc, &smem)) {
    auto kc = (C <= 5 * 256) ? ln_quant_kernel<5, true> : ln_quant_kernel<8, true>;
    if (launch_cluster_pdl(kc, grid, kFqThreads, smem, static_cast<cudaStream_t>(stream),
                           reinterpret_cast<const __half*>(x), ldx, M, C,
                           reinterpret_cast<const __half*>(gamma),
                           reinterpret_cast<const __half*>(beta), eps, q,
                           reinterpret_cast<__half*>(y_out), static_cast<DynWs*>(ws), scale_out,
                           zp_out, rpc, rpc) != cudaSuccess)
      return MIXDQ_ERR_CUDA;
    return MIXDQ_OK;
  }
  // LayerNorm -> fp16 + min/max, then the single-pass quantiser: needs the caller's y buffer
  if (mixdq_two_pass_enabled()) {
    const int rc = mixdq_q2_ln(reinterpret_cast<const __half*>(x), ldx, M, C,
                               reinterpret_cast<const __half*>(gamma),
                               reinterpret_cast<const __half*>(beta), eps, q,
                               reinterpret_cast<__half*>(y_out), scale_out, zp_out, ws,
                               static_cast<cudaStream_t>(stream));
    if (rc != MIXDQ_ERR_UNSUPPORTED) return rc;
  }
  // one row per warp; >= 2 rows per CTA keeps the grid <= 128 CTAs for the batch-1 blocks
  plan_rows(M, static_cast<int64_t>(C) * 2, 2, &grid, &rpc, &srows, &smem);
  auto kern = (C <= 5 * 256) ? ln_quant_kernel<5, false> : ln_quant_kernel<8, false>;
  if (launch_pdl(kern, grid, kFqThreads, smem, static_cast<cudaStream_t>(stream),
                 reinterpret_cast<const __half*>(x), ldx, M, C,
                 reinterpret_cast<const __half*>(gamma), reinterpret_cast<const __half*>(beta), eps,
                 q, reinterpret_cast<__half*>(y_out), static_cast<DynWs*>(ws), scale_out, zp_out,
                 rpc, srows) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return MIXDQ_OK;
}

extern "C" int mixdq_geglu_quant_i8_dynamic(const mixdq_half_t* hg, int64_t ld, int M, int I,
                                            int8_t* q, mixdq_half_t* y_out, float* scale_out,
                                            float* zp_out, void* ws, mixdq_stream_t stream) {
  if (M <= 0 || I <= 0 || !hg || !q || !scale_out || !zp_out || !ws || ld < 2 * static_cast<int64_t>(I))
    return MIXDQ_ERR_INVALID_ARG;
  if ((I & 7) || (ld & 7) || !al16(hg) || !al16(q) || (y_out && !al16(y_out)))
    return MIXDQ_ERR_ALIGNMENT;
  static bool attr = false;
  if (!attr) { if (set_smem(geglu_quant_kernel, kFqMaxSmem)) return MIXDQ_ERR_CUDA; attr = true; }
  int grid, rpc, srows, smem;
  plan_rows(M, static_cast<int64_t>(I) * 2, 1, &grid, &rpc, &srows, &smem);
  if (launch_pdl(geglu_quant_kernel, grid, kFqThreads, smem, static_cast<cudaStream_t>(stream),
                 reinterpret_cast<const __half*>(hg), ld, M, I, q, reinterpret_cast<__half*>(y_out),
                 static_cast<DynWs*>(ws), scale_out, zp_out, rpc, srows) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return MIXDQ_OK;
}

extern "C" int mixdq_quant_i8_dynamic_rows(const mixdq_half_t* x, int64_t ldx, int M, int cols,
                                           int8_t* q, float* scale_out, float* zp_out, void* ws,
                                           mixdq_stream_t stream) {
  if (M <= 0 || cols <= 0 || !x || !q || !scale_out || !zp_out || !ws || ldx < cols)
    return MIXDQ_ERR_INVALID_ARG;
  if ((cols & 7) || (ldx & 7) || !al16(x) || !al16(q)) return MIXDQ_ERR_ALIGNMENT;
  static bool attr = false;
  if (!attr) {
    if (set_smem(rows_quant_kernel<false>, kFqMaxSmem) || set_smem(rows_quant_kernel<true>, kFqMaxSmem))
      return MIXDQ_ERR_CUDA;
    attr = true;
  }
  int grid, rpc, srows, smem;
  static int ncl = -1;
  if (ncl < 0) ncl = max_cluster_ctas(rows_quant_kernel<true>, kFqThreads, kFqMaxSmem);
  if (plan_rows_cluster(M, static_cast<int64_t>(cols) * 2, 1, ncl, &grid, &rpc, &smem)) {
    if (launch_cluster_pdl(rows_quant_kernel<true>, grid, kFqThreads, smem,
                           static_cast<cudaStream_t>(stream), reinterpret_cast<const __half*>(x),
                           ldx, M, cols, q, static_cast<DynWs*>(ws), scale_out, zp_out, rpc,
                           rpc) != cudaSuccess)
      return MIXDQ_ERR_CUDA;
    return MIXDQ_OK;
  }
  if (mixdq_two_pass_enabled()) {
    const int rc = mixdq_q2_rows(reinterpret_cast<const __half*>(x), ldx, M, cols, q, scale_out,
                                 zp_out, ws, static_cast<cudaStream_t>(stream));
    if (rc != MIXDQ_ERR_UNSUPPORTED) return rc;
  }
  // >= 16 KB of fp16 (two 16-byte vectors per thread) per CTA
  int min_rows = static_cast<int>((8192 + cols - 1) / cols);
  if (min_rows < 1) min_rows = 1;
  plan_rows(M, static_cast<int64_t>(cols) * 2, min_rows, &grid, &rpc, &srows, &smem);
  if (launch_pdl(rows_quant_kernel<false>, grid, kFqThreads, smem,
                 static_cast<cudaStream_t>(stream), reinterpret_cast<const __half*>(x), ldx, M,
                 cols, q, static_cast<DynWs*>(ws), scale_out, zp_out, rpc, srows) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return MIXDQ_OK;
}

// qs != nullptr: STATIC scales — statistics kernel, then an apply kernel that quantises with
// (*s_inv, *s_zp) straight from its registers (q, y_out, scale_out, zp_out unused)
static int gn_entry(const mixdq_half_t* x, int64_t ldx, int NB, int HW, int C, int G,
                    const mixdq_half_t* gamma, const mixdq_half_t* beta, float eps, int silu,
                    int8_t* q, mixdq_half_t* y_out, float* scale_out, float* zp_out, void* ws,
                    mixdq_stream_t stream, int8_t* qs, const float* s_inv, const float* s_zp) {
  // q == nullptr (with y_out): normalise only, the caller quantises y with static parameters
  if (NB <= 0 || HW <= 0 || C <= 0 || G <= 0 || !x || !gamma || !beta || !ws || ldx < C ||
      (qs ? (!s_inv || !s_zp) : (q ? (!scale_out || !zp_out) : !y_out)))
    return MIXDQ_ERR_INVALID_ARG;
  if ((C & 7) || (ldx & 7) || !al16(x) || !al16(gamma) || !al16(beta) || (q && !al16(q)) ||
      (y_out && !al16(y_out)) || (qs && !al16(qs)))
    return MIXDQ_ERR_ALIGNMENT;
  const int cpg = C / G;
  // a 16-byte chunk of 8 channels must touch at most two groups
  const bool two_groups = (C % G == 0) && (cpg >= 8 || cpg == 4);
  if (!two_groups || G > 32 || C > kGnMaxChunks * 256 || NB > kNumSm ||
      static_cast<int64_t>(NB) * G > kMaxStatGroups)
    return MIXDQ_ERR_UNSUPPORTED;
  static bool attr = false;
  if (!attr) {
    if (set_smem(gn_quant_kernel<true, 0>, kFqMaxSmem) || set_smem(gn_quant_kernel<false, 0>, kFqMaxSmem) ||
        set_smem(gn_stats_kernel<2, 4>, kFqMaxSmem) || set_smem(gn_stats_kernel<5, 2>, kFqMaxSmem) ||
        set_smem(gn_stats_kernel<kGnMaxChunks, 1>, kFqMaxSmem) ||
        set_smem(gn_quant_kernel<true, 2>, kFqMaxSmem) || set_smem(gn_quant_kernel<false, 2>, kFqMaxSmem))
      return MIXDQ_ERR_CUDA;
    attr = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool three_kernels = (qs || (y_out && mixdq_two_pass_enabled())) &&
                             static_cast<int64_t>(NB) * HW * (C >> 3) < (1ll << 31);
  int cpi = kNumSm / NB;                       // CTAs per image, all co-resident
  const int max_useful = (HW + 3) / 4;         // >= 4 rows per CTA
  if (cpi > max_useful) cpi = max_useful;
  if (cpi < 1) cpi = 1;
  int rpc = (HW + cpi - 1) / cpi;
  rpc = (rpc + 3) / 4 * 4;                     // whole statistics units (4 / 2 / 1 rows) per CTA
  cpi = (HW + rpc - 1) / rpc;
  int64_t srows = kFqMaxSmem / (static_cast<int64_t>(C) * 2);
  if (srows > rpc) srows = rpc;
  int smem = static_cast<int>(srows * C * 2);
  const int scratch = kFqWarps * (C / 8) * 16;   // statistics partials [warps][chunks] float4
  if (smem < scratch) smem = scratch;
  if (three_kernels) {
    // statistics kernel -> apply kernel (fp16 y + min/max, or int8) -> single-pass quantiser. No
    // CTA waits for another one, so the grids need not be co-resident.
    auto k2 = silu ? gn_quant_kernel<true, 2> : gn_quant_kernel<false, 2>;
    const __half* xh = reinterpret_cast<const __half*>(x);
    const __half* gh = reinterpret_cast<const __half*>(gamma);
    const __half* bh = reinterpret_cast<const __half*>(beta);
    __half* yh = reinterpret_cast<__half*>(y_out);
    DynWs* w = static_cast<DynWs*>(ws);
    // statistics: one warp per unit, contiguous unit ranges per warp (<= a full GPU of warps),
    // the warps spread over as many CTAs as there are SMs
    const int ur = gn_unit_rows(C / 8);
    const long long units = static_cast<long long>(NB) * ((HW + ur - 1) / ur);
    long long warps = static_cast<long long>(kNumSm) * kFqWarps;
    if (warps > units) warps = units;
    const int upw = static_cast<int>((units + warps - 1) / warps);
    warps = (units + upw - 1) / upw;
    int g1 = warps < kNumSm ? static_cast<int>(warps) : kNumSm;
    const int wpc = static_cast<int>((warps + g1 - 1) / g1);          // <= kFqWarps
    g1 = static_cast<int>((warps + wpc - 1) / wpc);
    auto k1 = (C / 8 <= 64) ? gn_stats_kernel<2, 4>
              : (C / 8 <= 160) ? gn_stats_kernel<5, 2> : gn_stats_kernel<kGnMaxChunks, 1>;
    if (launch_pdl(k1, g1, 32 * wpc, wpc * (C / 8) * 16, st, xh, ldx, NB, HW, C, G, w, upw) !=
        cudaSuccess)
      return MIXDQ_ERR_CUDA;
    if (launch_pdl(k2, NB * cpi, kFqThreads, 0, st, xh, ldx, NB, HW, C, G, gh, bh, eps, q, yh, w,
                   scale_out, zp_out, cpi, rpc, 0, qs, s_inv, s_zp) != cudaSuccess)
      return MIXDQ_ERR_CUDA;
    if (qs != nullptr) return MIXDQ_OK;   // the apply kernel's last CTA cleared the accumulators
    if (q == nullptr) {
      // GroupNorm only (static-scale callers): the quantise pass that normally re-zeroes the
      // statistics accumulators does not run, so clear them here (a memset node under capture)
      return cudaMemsetAsync(w->gsum, 0, sizeof(unsigned long long) * NB * G * 2, st) == cudaSuccess
                 ? MIXDQ_OK : MIXDQ_ERR_CUDA;
    }
    return mixdq_q2_premm(yh, static_cast<int64_t>(NB) * HW * C, q, scale_out, zp_out, ws,
                          NB * cpi, w->gsum, NB * G * 2, st);
  }
  if (q == nullptr || qs != nullptr) return MIXDQ_ERR_UNSUPPORTED;
  auto gk = silu ? gn_quant_kernel<true, 0> : gn_quant_kernel<false, 0>;
  if (launch_pdl(gk, NB * cpi, kFqThreads, smem, st, reinterpret_cast<const __half*>(x), ldx, NB, HW,
                 C, G, reinterpret_cast<const __half*>(gamma), reinterpret_cast<const __half*>(beta),
                 eps, q, reinterpret_cast<__half*>(y_out), static_cast<DynWs*>(ws), scale_out,
                 zp_out, cpi, rpc, static_cast<int>(srows), static_cast<int8_t*>(nullptr), static_cast<const float*>(nullptr),
                   static_cast<const float*>(nullptr)) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return MIXDQ_OK;
}

extern "C" int mixdq_gn_quant_i8_dynamic(const mixdq_half_t* x, int64_t ldx, int NB, int HW, int C,
                                         int G, const mixdq_half_t* gamma,
                                         const mixdq_half_t* beta, float eps, int silu, int8_t* q,
                                         mixdq_half_t* y_out, float* scale_out, float* zp_out,
                                         void* ws, mixdq_stream_t stream) {
  return gn_entry(x, ldx, NB, HW, C, G, gamma, beta, eps, silu, q, y_out, scale_out, zp_out, ws,
                  stream, nullptr, nullptr, nullptr);
}

extern "C" int mixdq_gn_quant_i8_static(const mixdq_half_t* x, int64_t ldx, int NB, int HW, int C,
                                        int G, const mixdq_half_t* gamma, const mixdq_half_t* beta,
                                        float eps, int silu, const float* scale_inv,
                                        const float* zp, int8_t* q, void* ws,
                                        mixdq_stream_t stream) {
  if (!q) return MIXDQ_ERR_INVALID_ARG;
  return gn_entry(x, ldx, NB, HW, C, G, gamma, beta, eps, silu, nullptr, nullptr, nullptr, nullptr,
                  ws, stream, q, scale_inv, zp);
}

extern "C" int mixdq_ln_quant_i8_static(const mixdq_half_t* x, int64_t ldx, int M, int C,
                                        const mixdq_half_t* gamma, const mixdq_half_t* beta,
                                        float eps, const float* scale_inv, const float* zp,
                                        int8_t* q, void* ws, mixdq_stream_t stream) {
  if (M <= 0 || C <= 0 || !x || !gamma || !beta || !ws || ldx < C || !q || !scale_inv || !zp)
    return MIXDQ_ERR_INVALID_ARG;
  if ((C & 7) || (ldx & 7) || !al16(x) || !al16(gamma) || !al16(beta) || !al16(q))
    return MIXDQ_ERR_ALIGNMENT;
  if (C > kLnMaxChunks * 256) return MIXDQ_ERR_UNSUPPORTED;
  return mixdq_q2_ln(reinterpret_cast<const __half*>(x), ldx, M, C,
                     reinterpret_cast<const __half*>(gamma), reinterpret_cast<const __half*>(beta),
                     eps, nullptr, nullptr, nullptr, nullptr, ws, static_cast<cudaStream_t>(stream),
                     q, scale_inv, zp);
}
