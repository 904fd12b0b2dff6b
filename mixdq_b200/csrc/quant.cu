// quant.cu — K-Q: per-tensor activation quantisation fp16 -> int8 (HBM-bound, 3 B/element).
//
//   static  : q = clamp(lrintf(fmaf(x, 1/delta, zp)), -128, 127)      reference CUDA formula
//             (reference csrc/quant_dequant/quantize_kernel.cu:20-24, nvcc contracts x*s+z to FMA)
//   dynamic : q = clamp(rint(x / delta) + z, 0, 255) - 128             qdiff formula, fp32, true
//             division (reference quant_utils/qdiff/quantizer/base_quantizer.py:122-128)
//
// Layout variants: flat dense, 3-D strided view -> dense rows, NCHW(strided) -> NHWC transpose.
// All loads are 16-byte (8 halves), all stores 8-byte (8 codes) on the vector paths.
#include "common.cuh"
#include "quant_ws.cuh"
#include "../../include/mixdq_b200.h"

namespace mixdq {

enum QuantMode { kStaticFma = 0, kDynamicDiv = 1 };

struct QParams {
  float a;  // static: 1/delta      dynamic: delta
  float b;  // static: zp (-128 shifted)   dynamic: z (unshifted, in [0,255])
  float inv;  // dynamic: 1/delta (set by quant_vec8 / the callers of quant_one)
  int lo, hi; // static: code range ([-128, 127]; [0, 15] for 4-bit activations)
};

template <int MODE>
__device__ __forceinline__ int quant_one(float x, QParams p) {
  if (MODE == kStaticFma) {
    int v = __float2int_rn(__fmaf_rn(x, p.a, p.b));
    return min(max(v, p.lo), p.hi);
  } else {
    return qdiff_code(x, p.a, p.inv, p.b);
  }
}

template <int MODE>
__device__ __forceinline__ QParams load_qparams(const float* a, const float* b) {
  QParams p;
  p.a = __ldg(a);
  p.b = __ldg(b);
  p.inv = 0.0f;
  p.lo = -128;
  p.hi = 127;
  if (MODE == kDynamicDiv) {
    p.b = p.b + 128.0f;  // stored zero point is z-128; exact in fp32
    p.inv = __frcp_rn(p.a);
  }
  return p;
}

template <int MODE>
__device__ __forceinline__ uint2 quant_vec8(const int4& raw, QParams p) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
  if (MODE == kDynamicDiv) p.inv = __frcp_rn(p.a);
  int q[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h2[i]);
    q[2 * i] = quant_one<MODE>(f.x, p);
    q[2 * i + 1] = quant_one<MODE>(f.y, p);
  }
  uint2 out;
  out.x = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((q[3] & 0xff) << 24);
  out.y = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((q[7] & 0xff) << 24);
  return out;
}

__device__ __forceinline__ int4 ld_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---- flat dense ---------------------------------------------------------------------------
constexpr int kQuantThreads = 256;
constexpr int kQuantUnroll = 4;  // 16-byte loads in flight per thread

template <int MODE>
__global__ void __launch_bounds__(kQuantThreads)
quant_flat_kernel(const __half* __restrict__ x, int8_t* __restrict__ q, int64_t numel,
                  const float* __restrict__ pa, const float* __restrict__ pb, int lo, int hi) {
  pdl_launch_dependents();   // the GEMM/conv that consumes q may start its weight prefetch now
  QParams p = load_qparams<MODE>(pa, pb);
  p.lo = lo;
  p.hi = hi;
  const int64_t nvec = numel >> 3;
  const int4* xv = reinterpret_cast<const int4*>(x);
  uint2* qv = reinterpret_cast<uint2*>(q);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kQuantThreads;
  int64_t i = static_cast<int64_t>(blockIdx.x) * kQuantThreads + threadIdx.x;
  for (; i + (kQuantUnroll - 1) * stride < nvec; i += kQuantUnroll * stride) {
    int4 r[kQuantUnroll];
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) r[u] = ld_stream16(xv + i + u * stride);
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) qv[i + u * stride] = quant_vec8<MODE>(r[u], p);
  }
  for (; i < nvec; i += stride) qv[i] = quant_vec8<MODE>(ld_stream16(xv + i), p);
  // scalar tail (numel % 8)
  const int64_t tail0 = nvec << 3;
  const int64_t t = tail0 + static_cast<int64_t>(blockIdx.x) * kQuantThreads + threadIdx.x;
  if (t < numel) q[t] = static_cast<int8_t>(quant_one<MODE>(__half2float(x[t]), p));
}

// Unaligned base pointers (never produced by torch allocations, but legal through the C ABI).
template <int MODE>
__global__ void __launch_bounds__(kQuantThreads)
quant_flat_scalar_kernel(const __half* __restrict__ x, int8_t* __restrict__ q, int64_t numel,
                         const float* __restrict__ pa, const float* __restrict__ pb, int lo,
                         int hi) {
  QParams p = load_qparams<MODE>(pa, pb);
  p.lo = lo;
  p.hi = hi;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kQuantThreads;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kQuantThreads + threadIdx.x; i < numel;
       i += stride)
    q[i] = static_cast<int8_t>(quant_one<MODE>(__half2float(x[i]), p));
}

// ---- 3-D strided view [d0][d1][cols] -> dense rows -----------------------------------------
template <int MODE, bool VEC>
__global__ void __launch_bounds__(kQuantThreads)
quant_strided_kernel(const __half* __restrict__ x, int8_t* __restrict__ q, int64_t rows,
                     int64_t d1, int64_t cols, int64_t s0, int64_t s1, int64_t out_pitch,
                     const float* __restrict__ pa, const float* __restrict__ pb) {
  pdl_launch_dependents();
  const QParams p = load_qparams<MODE>(pa, pb);
  const int64_t cpr = VEC ? (cols >> 3) : cols;  // work items per row
  const int64_t total = rows * cpr;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kQuantThreads;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kQuantThreads + threadIdx.x; i < total;
       i += stride) {
    const int64_t row = i / cpr;
    const int64_t c = i - row * cpr;
    const int64_t i0 = row / d1;
    const int64_t i1 = row - i0 * d1;
    const __half* src = x + i0 * s0 + i1 * s1;
    int8_t* dst = q + row * out_pitch;
    if (VEC) {
      int4 r = ld_stream16(reinterpret_cast<const int4*>(src) + c);
      reinterpret_cast<uint2*>(dst)[c] = quant_vec8<MODE>(r, p);
    } else {
      dst[c] = static_cast<int8_t>(quant_one<MODE>(__half2float(src[c]), p));
    }
  }
}

// ---- NCHW (arbitrary strides) -> NHWC, channel range [c0, c1) ------------------------------
constexpr int kTrC = 64;   // channels per tile
constexpr int kTrP = 64;   // pixels per tile
template <int MODE>
__global__ void __launch_bounds__(256)
quant_nchw2nhwc_kernel(const __half* __restrict__ x, int8_t* __restrict__ q, int C0, int Csel,
                       int H, int W, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                       const float* __restrict__ pa, const float* __restrict__ pb) {
  __shared__ int8_t tile[kTrP][kTrC + 4];
  pdl_launch_dependents();
  const QParams p = load_qparams<MODE>(pa, pb);
  const int HW = H * W;
  const int n = blockIdx.z;
  const int pix0 = blockIdx.x * kTrP;
  const int ch0 = blockIdx.y * kTrC;
  const int tp = threadIdx.x & (kTrP - 1);
  const int tc = threadIdx.x >> 6;  // 0..3
  const int pix = pix0 + tp;
  if (pix < HW) {
    const int h = pix / W, w = pix - h * W;
    const __half* src = x + n * sn + h * sh + w * sw;
#pragma unroll 4
    for (int i = 0; i < kTrC / 4; ++i) {
      const int c = ch0 + tc + 4 * i;
      if (c < Csel)
        tile[tp][tc + 4 * i] =
            static_cast<int8_t>(quant_one<MODE>(__half2float(src[(C0 + c) * sc]), p));
    }
  }
  __syncthreads();
  // store: 16 threads x 4 B cover one pixel's 64 channels
  const int lane16 = threadIdx.x & 15;
  const int prow = threadIdx.x >> 4;  // 0..15
  const bool vec_ok = (Csel & 3) == 0;
#pragma unroll
  for (int i = 0; i < kTrP / 16; ++i) {
    const int pl = prow + 16 * i;
    const int pp = pix0 + pl;
    const int c = ch0 + lane16 * 4;
    if (pp < HW && c < Csel) {
      int8_t* dst = q + (static_cast<int64_t>(n) * HW + pp) * Csel + c;
      if (vec_ok) {
        *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(&tile[pl][lane16 * 4]);
      } else {
        for (int j = 0; j < 4 && c + j < Csel; ++j) dst[j] = tile[pl][lane16 * 4 + j];
      }
    }
  }
}

// ---- dynamic: ONE kernel = min/max reduction -> grid barrier -> quantise ---------------------
// All CTAs are co-resident (grid <= 2 per SM), so a flag-based grid barrier is safe: every CTA
// publishes its partial min/max, the last one to arrive computes (delta, z) and raises the flag,
// the others spin on it with an acquire load. Each thread keeps its first kDynCache 16-byte
// vectors in registers, so tensors up to 296*256*kDynCache*8 elements (4.8 M: every layer of the
// batch-1 SDXL step) are read from memory exactly once; larger tensors are re-read (L2 hits).
constexpr int kDynCache = 8;

// CLUSTER: the launch is one thread-block cluster (<= 16 CTAs of NT threads) and carries a
// programmatic dependency: min/max go through DSMEM + the cluster barrier, and the kernel's launch
// latency hides behind the tail of the producer (griddepcontrol.wait before the first read).
template <int NT, bool CLUSTER>
__global__ void __launch_bounds__(NT)
quant_dynamic_fused_kernel(const __half* __restrict__ x, int8_t* __restrict__ q, int64_t numel,
                           DynWs* __restrict__ ws, float* __restrict__ scale_out,
                           float* __restrict__ zp_out) {
  constexpr int kQuantThreads = NT;   // shadows the file-level constant inside this kernel
  QDbg dbg;
  dbg.begin(ws);
  pdl_launch_dependents();
  if (CLUSTER) cluster_enter();
  pdl_wait();
  dbg.waited(ws);
  float mn = 0.0f, mx = 0.0f;  // qdiff clamps x_min <= 0 <= x_max (base_quantizer.py:155-158)
  const int64_t nvec = numel >> 3;
  const int4* xv = reinterpret_cast<const int4*>(x);
  uint2* qv = reinterpret_cast<uint2*>(q);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kQuantThreads;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * kQuantThreads + threadIdx.x;
  int4 cache[kDynCache];
#pragma unroll
  for (int u = 0; u < kDynCache; ++u) {
    const int64_t i = i0 + u * stride;
    if (i < nvec) { cache[u] = ld_stream16(xv + i); }
  }
#pragma unroll
  for (int u = 0; u < kDynCache; ++u)
    if (i0 + u * stride < nvec) minmax_vec8(cache[u], mn, mx);
  for (int64_t i = i0 + kDynCache * stride; i < nvec; i += stride) minmax_vec8(__ldg(xv + i), mn, mx);
  const int64_t t = (nvec << 3) + i0;   // scalar tail (numel % 8)
  float tailv = 0.0f;
  if (t < numel) {
    tailv = __half2float(x[t]);
    mn = fminf(mn, tailv);
    mx = fmaxf(mx, tailv);
  }
  QParams p;
  dbg.stamp(2);
  if (CLUSTER) cluster_minmax_params<NT>(mn, mx, scale_out, zp_out, p.a, p.b);
  else grid_minmax_params<NT>(ws, mn, mx, scale_out, zp_out, p.a, p.b);
  p.inv = __frcp_rn(p.a);
  dbg.stamp(3);
#pragma unroll
  for (int u = 0; u < kDynCache; ++u) {
    const int64_t i = i0 + u * stride;
    if (i < nvec) qv[i] = quant_vec8<kDynamicDiv>(cache[u], p);
  }
  for (int64_t i = i0 + kDynCache * stride; i < nvec; i += stride)
    qv[i] = quant_vec8<kDynamicDiv>(__ldg(xv + i), p);
  if (t < numel) q[t] = static_cast<int8_t>(quant_one<kDynamicDiv>(tailv, p));
  dbg.end(ws);
}

static inline int grid_for(int64_t items, int per_block, int max_blocks) {
  int64_t g = (items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

}  // namespace mixdq

using namespace mixdq;

// quant2.cu: the two-kernel (min/max pass + quantise pass) implementations
int mixdq_q2_rows(const __half* x, int64_t ldx, int64_t M, int cols, int8_t* q, float* scale_out,
                  float* zp_out, void* ws, cudaStream_t st, int n_bits = 8);
int mixdq_q2_premm(const __half* x, int64_t numel, int8_t* q, float* scale_out, float* zp_out,
                   void* ws, int nparts, unsigned long long* zero_words, int zero_n,
                   cudaStream_t st);
// 1 (default) = min/max pass + quantise pass chained by programmatic dependent launch; 0 = the
// first-generation single-kernel quantisers with the counter barrier, kept as the fallback for
// callers without a scratch buffer and as the A/B reference of the parity tests. Two further
// one-kernel designs (tagged-partial barrier, one-cluster DSMEM) were measured slower inside the
// whole-UNet graph and removed (profiles/README.md section 3 keeps the numbers).
// MIXDQ_QUANT_MODE / mixdq_debug_set_two_pass select the mode.
static int g_two_pass = -1;
int mixdq_quant_mode() {
  if (g_two_pass < 0) {
    const char* e = getenv("MIXDQ_QUANT_MODE");
    g_two_pass = (e && e[0] == '0') ? 0 : 1;
  }
  return g_two_pass;
}
bool mixdq_two_pass_enabled() { return mixdq_quant_mode() != 0; }
extern "C" void mixdq_debug_set_two_pass(int mode) { g_two_pass = mode <= 0 ? 0 : 1; }

extern "C" void mixdq_debug_set_cluster(int on) { cluster_mode_flag() = on ? 1 : 0; }
// profiling: point the workspace at a stamp buffer (or NULL) and restart the launch sequence;
// stream-ordered device writes, no synchronisation
extern "C" int mixdq_debug_set_quant_timing_buffer(void* ws, void* dev_ptr, mixdq_stream_t stream) {
  if (!ws) return MIXDQ_ERR_INVALID_ARG;
  DynWs* w = static_cast<DynWs*>(ws);
  static unsigned long long* h_ptr[64];
  static unsigned int h_zero = 0;
  static int slot = 0;
  unsigned long long** hp = &h_ptr[slot++ & 63];
  *hp = static_cast<unsigned long long*>(dev_ptr);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemcpyAsync(&w->qdbg, hp, sizeof(*hp), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(&w->qdbg_seq, &h_zero, sizeof(h_zero), cudaMemcpyHostToDevice, st) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return MIXDQ_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }

#define MIXDQ_CHECK_LAUNCH()                                   \
  do {                                                         \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) return MIXDQ_ERR_CUDA;             \
  } while (0)

template <int MODE>
static int launch_flat(const __half* x, int64_t numel, const float* a, const float* b, int8_t* q,
                       cudaStream_t st, int lo = -128, int hi = 127) {
  if (numel == 0) return MIXDQ_OK;
  const int max_blocks = 148 * 8;
  if (aligned16(x) && aligned8(q)) {
    int grid = grid_for(numel >> 3, kQuantThreads * kQuantUnroll, max_blocks);
    quant_flat_kernel<MODE><<<grid, kQuantThreads, 0, st>>>(x, q, numel, a, b, lo, hi);
  } else {
    int grid = grid_for(numel, kQuantThreads, max_blocks);
    quant_flat_scalar_kernel<MODE><<<grid, kQuantThreads, 0, st>>>(x, q, numel, a, b, lo, hi);
  }
  MIXDQ_CHECK_LAUNCH();
  return MIXDQ_OK;
}

extern "C" int mixdq_quant_i8_static(const mixdq_half_t* x, int64_t numel, const float* scale_inv,
                                     const float* zp, int8_t* q, mixdq_stream_t stream) {
  if (numel < 0 || (numel > 0 && (!x || !q)) || !scale_inv || !zp) return MIXDQ_ERR_INVALID_ARG;
  return launch_flat<kStaticFma>(reinterpret_cast<const __half*>(x), numel, scale_inv, zp, q,
                                 static_cast<cudaStream_t>(stream));
}

// A1 with an explicit code range: q = clamp(lrintf(x * scale_inv + zp), lo, hi). The 4-bit
// activation layers use [0, 15] with the UNSHIFTED zero point of the PTQ checkpoint.
extern "C" int mixdq_quant_i8_static_range(const mixdq_half_t* x, int64_t numel,
                                           const float* scale_inv, const float* zp, int lo,
                                           int hi, int8_t* q, mixdq_stream_t stream) {
  if (numel < 0 || (numel > 0 && (!x || !q)) || !scale_inv || !zp || lo < -128 || hi > 127 ||
      lo > hi)
    return MIXDQ_ERR_INVALID_ARG;
  return launch_flat<kStaticFma>(reinterpret_cast<const __half*>(x), numel, scale_inv, zp, q,
                                 static_cast<cudaStream_t>(stream), lo, hi);
}

extern "C" int mixdq_quant_i8_static_strided(const mixdq_half_t* x, int64_t d0, int64_t d1,
                                             int64_t cols, int64_t s0, int64_t s1,
                                             const float* scale_inv, const float* zp, int8_t* q,
                                             int64_t out_pitch, mixdq_stream_t stream) {
  if (d0 < 0 || d1 < 0 || cols < 0 || out_pitch < cols || !scale_inv || !zp)
    return MIXDQ_ERR_INVALID_ARG;
  const int64_t rows = d0 * d1;
  if (rows == 0 || cols == 0) return MIXDQ_OK;
  if (!x || !q) return MIXDQ_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* xh = reinterpret_cast<const __half*>(x);
  // fully dense view -> flat kernel
  if (s1 == cols && (s0 == d1 * cols || d0 == 1) && out_pitch == cols)
    return launch_flat<kStaticFma>(xh, rows * cols, scale_inv, zp, q, st);
  const bool vec = (cols % 8 == 0) && (s0 % 8 == 0) && (s1 % 8 == 0) && (out_pitch % 8 == 0) &&
                   aligned16(x) && aligned8(q);
  const int64_t items = vec ? rows * (cols >> 3) : rows * cols;
  int grid = grid_for(items, kQuantThreads, 148 * 16);
  if (vec)
    quant_strided_kernel<kStaticFma, true><<<grid, kQuantThreads, 0, st>>>(
        xh, q, rows, d1, cols, s0, s1, out_pitch, scale_inv, zp);
  else
    quant_strided_kernel<kStaticFma, false><<<grid, kQuantThreads, 0, st>>>(
        xh, q, rows, d1, cols, s0, s1, out_pitch, scale_inv, zp);
  MIXDQ_CHECK_LAUNCH();
  return MIXDQ_OK;
}

extern "C" int mixdq_quant_i8_nchw2nhwc(const mixdq_half_t* x, int N, int C, int H, int W,
                                        const int64_t xstride[4], int c_begin, int c_end,
                                        const float* scale_inv, const float* zp, int8_t* q_nhwc,
                                        mixdq_stream_t stream) {
  if (N < 0 || C <= 0 || H < 0 || W < 0 || !xstride || c_begin < 0 || c_end > C ||
      c_begin >= c_end || !scale_inv || !zp)
    return MIXDQ_ERR_INVALID_ARG;
  if (N == 0 || H == 0 || W == 0) return MIXDQ_OK;
  if (!x || !q_nhwc || N > 65535) return MIXDQ_ERR_INVALID_ARG;
  const int Csel = c_end - c_begin;
  const int HW = H * W;
  dim3 grid((HW + kTrP - 1) / kTrP, (Csel + kTrC - 1) / kTrC, N);
  if (grid.y > 65535) return MIXDQ_ERR_UNSUPPORTED;
  quant_nchw2nhwc_kernel<kStaticFma><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), q_nhwc, c_begin, Csel, H, W, xstride[0], xstride[1],
      xstride[2], xstride[3], scale_inv, zp);
  MIXDQ_CHECK_LAUNCH();
  return MIXDQ_OK;
}

extern "C" int64_t mixdq_quant_dynamic_ws_bytes(void) { return sizeof(DynWs); }

extern "C" int mixdq_quant_i8_dynamic(const mixdq_half_t* x, int64_t numel, float* scale_out,
                                      float* zp_out, int8_t* q, void* ws,
                                      mixdq_stream_t stream) {
  if (numel <= 0 || !x || !q || !scale_out || !zp_out || !ws) return MIXDQ_ERR_INVALID_ARG;
  if (!aligned16(x) || !aligned8(q)) return MIXDQ_ERR_ALIGNMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* xh = reinterpret_cast<const __half*>(x);
  // tiny tensors (embeddings, pooled vectors): one cluster, DSMEM + hardware barrier. Anything
  // larger needs the arithmetic throughput of the whole chip (measured: 0.33 M elements on one
  // 16-CTA cluster took 7.5 us, issue-bound) and takes the counter barrier below.
  constexpr int kClThreads = 512;
  static int ncl = -1;
  if (ncl < 0) ncl = max_cluster_ctas(quant_dynamic_fused_kernel<kClThreads, true>, kClThreads, 0);
  const int ncl_now = cluster_enabled() ? ncl : 0;
  if (ncl_now > 0 && numel <= kClusterMaxElems) {
    int grid = grid_for(numel >> 3, kClThreads, ncl_now);
    if (launch_cluster_pdl(quant_dynamic_fused_kernel<kClThreads, true>, grid, kClThreads, 0, st,
                           xh, q, numel, static_cast<DynWs*>(ws), scale_out, zp_out) != cudaSuccess)
      return MIXDQ_ERR_CUDA;
    return MIXDQ_OK;
  }
  // two short kernels (min/max, then quantise) beat one kernel with a grid barrier (quant2.cu)
  if (mixdq_two_pass_enabled() && (numel & 7) == 0 && numel < (1ll << 31)) {
    const int rc = mixdq_q2_rows(xh, numel, 1, static_cast<int>(numel), q, scale_out, zp_out, ws, st);
    if (rc != MIXDQ_ERR_UNSUPPORTED) return rc;
  }
  // co-resident grid: at most 2 CTAs per SM (see quant_dynamic_fused_kernel)
  int grid = grid_for(numel >> 3, kQuantThreads * 2, 148 * 2);
  if (launch_pdl(quant_dynamic_fused_kernel<kQuantThreads, false>, grid, kQuantThreads, 0, st, xh,
                 q, numel, static_cast<DynWs*>(ws), scale_out, zp_out) != cudaSuccess)
    return MIXDQ_ERR_CUDA;
  return MIXDQ_OK;
}

extern "C" int mixdq_quant_i8_premm(const mixdq_half_t* x, int64_t numel, int8_t* q,
                                    float* scale_out, float* zp_out, void* ws,
                                    mixdq_stream_t stream) {
  if (numel <= 0 || !x || !q || !scale_out || !zp_out || !ws) return MIXDQ_ERR_INVALID_ARG;
  if ((numel & 7) || !aligned16(x) || !aligned8(q)) return MIXDQ_ERR_ALIGNMENT;
  return mixdq_q2_premm(reinterpret_cast<const __half*>(x), numel, q, scale_out, zp_out, ws,
                        partial_count_slot(ws), nullptr, 0, static_cast<cudaStream_t>(stream));
}

// A10 with a code range: n_bits = 8 -> codes - 128 (same as mixdq_quant_i8_dynamic_rows), n_bits = 4
// -> codes 0..15 stored as they are, *zp_out = z (the 4-bit activation layers of the act_7.xx bit
// configs, which the reference gates to fp16, nn/Linear.py:28-36). Row-pitched view [M][cols].
extern "C" int mixdq_quant_i8_dynamic_bits(const mixdq_half_t* x, int64_t ldx, int64_t M,
                                           int64_t cols, int n_bits, float* scale_out,
                                           float* zp_out, int8_t* q, void* ws,
                                           mixdq_stream_t stream) {
  if (M <= 0 || cols <= 0 || ldx < cols || !x || !q || !scale_out || !zp_out || !ws)
    return MIXDQ_ERR_INVALID_ARG;
  if (n_bits != 8 && n_bits != 4) return MIXDQ_ERR_UNSUPPORTED;
  if ((cols & 7) || (ldx & 7) || !aligned16(x) || !aligned8(q) || cols > (1ll << 30))
    return MIXDQ_ERR_ALIGNMENT;
  return mixdq_q2_rows(reinterpret_cast<const __half*>(x), ldx, M, static_cast<int>(cols), q,
                       scale_out, zp_out, ws, static_cast<cudaStream_t>(stream), n_bits);
}
