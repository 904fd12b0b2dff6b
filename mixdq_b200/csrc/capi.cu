// capi.cu — extern "C" entry points of the contraction kernels: argument validation, TMA tensor-map
// construction, tile-shape heuristic, launch. Declared in include/mixdq_b200.h.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "../../include/mixdq_b200.h"
#include "simt.h"
#include "persist.h"
#include "tc_kernel.cuh"
#include "quant_ws.cuh"

using namespace mixdq;

// ------------------------------------------------------------------------------------------
// misc state
// ------------------------------------------------------------------------------------------
static thread_local const char* g_last_path = "none";
static int g_force_simt = 0;

extern "C" int mixdq_abi_version(void) { return MIXDQ_ABI_VERSION; }
extern "C" const char* mixdq_last_path(void) { return g_last_path; }
extern "C" void mixdq_force_simt(int on) { g_force_simt = on; }

extern "C" const char* mixdq_strerror(int code) {
  switch (code) {
    case MIXDQ_OK: return "success";
    case MIXDQ_ERR_INVALID_ARG: return "invalid argument (null pointer or non-positive size)";
    case MIXDQ_ERR_ALIGNMENT:
      return "Int8 kernel with input or output alignment not to 4 is not supported.";
    case MIXDQ_ERR_UNSUPPORTED: return "unsupported configuration";
    case MIXDQ_ERR_CUDA: return "CUDA kernel failed";
    case MIXDQ_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

// ------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point fetched at run time: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// rank-R uint8 tensor, dims/box innermost first, strides in bytes for dims 1..R-1.
static bool make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides, const uint32_t* box,
                      const uint32_t* elem_strides = nullptr, bool swizzle = true) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, static_cast<cuuint32_t>(rank),
                   const_cast<void*>(base), reinterpret_cast<const cuuint64_t*>(dims),
                   reinterpret_cast<const cuuint64_t*>(strides),
                   reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static bool make_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows,
                         uint64_t pitch_bytes, uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {pitch_bytes};
  uint32_t box[2] = {BLOCK_K, box_rows};
  return make_tmap(m, base, 2, dims, strides, box);
}
// fp16 output matrix [rows][cols] (row pitch ld halves) for the TMA-store epilogue of the
// persistent kernel: 16-column x 32-row boxes in the SWIZZLE_32B shared-memory layout
static bool make_tmap_out(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows,
                          uint64_t ld) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// packed 4-bit weights [rows][k/2 bytes]: 64-byte (one k-block) wide boxes, no swizzle — the
// converter warps read the tile linearly and write the swizzled int8 layout themselves
static bool make_tmap_2d_w4(CUtensorMap* m, const void* base, uint64_t k, uint64_t rows,
                            uint32_t box_rows) {
  uint64_t dims[2] = {k / 2, rows};
  uint64_t strides[1] = {k / 2};
  uint32_t box[2] = {BLOCK_K / 2, box_rows};
  return make_tmap(m, base, 2, dims, strides, box, nullptr, false);
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
extern int g_use_pdl_fwd;
template <int BN, int STAGES, int KIND, bool W4 = false>
static int launch_tc(dim3 grid, const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& a1,
                     const CUtensorMap& w1, const TcParams& p, cudaStream_t st) {
  using L = TcSmem<BN, STAGES, KIND, W4>;
  static bool attr_set = false;
  auto kern = tc_i8_kernel<BN, STAGES, KIND, W4>;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) !=
        cudaSuccess)
      return MIXDQ_ERR_CUDA;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = grid.z;   // split-K ranks of one tile form a cluster
  // programmatic dependent launch: prologue + weight prefetch overlap the preceding kernel
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl_fwd ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, w, a1, w1, p);
  return e == cudaSuccess ? MIXDQ_OK : MIXDQ_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------
// Tile-shape heuristic: a cost model fitted to in-kernel %globaltimer stamps on B200
// (tools/phase_timing.py, tools/sweep_shapes.py; numbers in ns):
//   * ~1400 from the end of the preceding kernel to the first stage landing (PDL wait + TMA);
//   * per 128-byte k-block: 112 for BN <= 64 (two issuing warps: 54 cycles per MMA, the pipe's
//     floor), 195 for BN = 128 (one issuer, two k-blocks = 8 tcgen05.mma per barrier round:
//     ~52-64 cycles per MMA — the issuing thread is blocked for each MMA's duration — plus ~230
//     cycles of wait/fence/commit per stage), 420 for BN = 256 (one k-block per stage, ~130-160
//     cycles per MMA); it was 340 with one issuer and one k-block per stage;
//   * epilogue, no split: ~300 + 6.5 per tile column (dequant and ~26 B/clk/SM of stores overlap);
//   * epilogue, split-K: 8.6 per column to write the INT32 partial tile, ~900 for the cluster
//     barrier, 350 + 0.4 per owned element to sum the partials, 300 to store;
//   * clusters of 4 / 8 CTAs with one CTA per SM fit ~132 / ~120 CTAs per wave.
// Batch-1 layers with K <= 2048 come out unsplit with 64-wide tiles, long-K layers (ff.net.2,
// 3x3 convolutions at 16x16 / 32x32) split 4-8 ways with 128/256-wide tiles, large-M layers get
// 128x256 tiles.
// ------------------------------------------------------------------------------------------
static int g_force_bn = -1;
static int g_force_splits = 0;
int g_use_pdl_fwd = 1;
extern "C" void mixdq_debug_set_pdl(int on) { g_use_pdl_fwd = on; }

// split-K exchange workspaces, registered by the host side: one per (device, stream). A launch on
// stream S only ever uses the workspace registered for S (no registration -> no split-K), so two
// streams can never corrupt each other's partial tiles.
struct WsEntry { int device; cudaStream_t stream; int32_t* ptr; int64_t bytes; };
static WsEntry g_ws_tab[256];
static int g_ws_n = 0;
static std::mutex g_ws_mu;
extern "C" int mixdq_set_workspace(int device, mixdq_stream_t stream, void* ptr, int64_t bytes) {
  if (device < 0 || bytes < 0) return MIXDQ_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(g_ws_mu);
  for (int i = 0; i < g_ws_n; ++i) {
    if (g_ws_tab[i].device == device && g_ws_tab[i].stream == st) {
      if (ptr) { g_ws_tab[i].ptr = static_cast<int32_t*>(ptr); g_ws_tab[i].bytes = bytes; }
      else { g_ws_tab[i] = g_ws_tab[--g_ws_n]; }
      return MIXDQ_OK;
    }
  }
  if (!ptr) return MIXDQ_OK;
  if (g_ws_n >= 256) return MIXDQ_ERR_WORKSPACE;
  g_ws_tab[g_ws_n++] = WsEntry{device, st, static_cast<int32_t*>(ptr), bytes};
  return MIXDQ_OK;
}
static int64_t current_ws(cudaStream_t st, int32_t** ptr) {
  *ptr = nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  std::lock_guard<std::mutex> lock(g_ws_mu);
  for (int i = 0; i < g_ws_n; ++i)
    if (g_ws_tab[i].device == dev && g_ws_tab[i].stream == st) {
      *ptr = g_ws_tab[i].ptr;
      return g_ws_tab[i].bytes;
    }
  return 0;
}
// 0 when `stream` is not being captured into a CUDA graph, else the capture sequence's unique id.
extern "C" int mixdq_stream_capture_id(mixdq_stream_t stream, unsigned long long* id_out) {
  if (!id_out) return MIXDQ_ERR_INVALID_ARG;
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  if (cudaStreamGetCaptureInfo(static_cast<cudaStream_t>(stream), &status, &id) != cudaSuccess) {
    cudaGetLastError();   // e.g. the legacy default stream while another stream captures
    *id_out = 0ull;
    return MIXDQ_OK;
  }
  *id_out = (status == cudaStreamCaptureStatusActive) ? id : 0ull;
  return MIXDQ_OK;
}
extern "C" void mixdq_debug_set_persist(int mode, int cluster) { persist_set_mode(mode, cluster); }
extern "C" void mixdq_debug_set_persist_bn(int bn) { persist_force_bn(bn); }
extern "C" void mixdq_debug_set_conv_halo(int on) { persist_set_halo(on); }
extern "C" void mixdq_debug_force_bn(int bn) { g_force_bn = bn; }
extern "C" void mixdq_debug_force_splits(int s) { g_force_splits = s; }
// MIXDQ_A_PREFETCH=1 enables an L2 prefetch of the first A tile before the dependency wait.
// Measured neutral on B200 (7.878 vs 7.872 ms per batch-1 step), hence off by default.
// weight k-blocks beyond the ring -> L2 before the dependency wait: same-box A/B, twice each:
// batch 1 7.3437 / 7.3435 -> 7.3267 / 7.3283 ms/step, batch 8 16.827 / 16.828 -> 16.840 / 16.839
constexpr bool kWPrefetchDefault = true;
static int g_a_prefetch = -1;
static int a_prefetch_flag() {
  if (g_a_prefetch < 0) {
    const char* e = getenv("MIXDQ_A_PREFETCH");
    g_a_prefetch = (e && e[0] == '1') ? 1 : 0;
    // bit 1: weight k-blocks beyond the ring are L2-prefetched before the dependency wait
    const char* w = getenv("MIXDQ_W_PREFETCH");
    if (w ? (w[0] != '0') : kWPrefetchDefault) g_a_prefetch |= 2;
  }
  return g_a_prefetch;
}
static int g_dbg_mode = 0;
extern "C" void mixdq_debug_set_mode(int mode) { g_dbg_mode = mode; }
static unsigned long long* g_dbg = nullptr;
extern "C" void mixdq_debug_set_timing_buffer(void* dev_ptr) {
  g_dbg = static_cast<unsigned long long*>(dev_ptr);
}

static inline bool valid_bn(int bn) {
  return bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256;
}

static void pick_tile(int m_tiles, int N, int total_kb, bool allow_split, cudaStream_t st,
                      int* bn_out, int* splits_out) {
  int32_t* ws_ptr = nullptr;
  const int64_t ws_bytes = allow_split ? current_ws(st, &ws_ptr) : 0;
  if (g_force_bn < 0) {
    const char* e = getenv("MIXDQ_FORCE_BN");
    g_force_bn = e ? atoi(e) : 0;
    const char* f = getenv("MIXDQ_FORCE_SPLITS");
    if (f) g_force_splits = atoi(f);
  }
  const int kNumSm = 148;
  double best = 1e30;
  int best_bn = 32, best_s = 1;
  const int cands[5] = {256, 128, 64, 32, 16};
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    if (valid_bn(g_force_bn) && bn != g_force_bn) continue;
    if (!valid_bn(g_force_bn) && bn > 16 && bn / 2 >= N) continue;  // tile twice as wide as N
    const long tiles = static_cast<long>(m_tiles) * ((N + bn - 1) / bn);
    for (int s = 1; s <= 8; s *= 2) {
      if (g_force_splits > 0 && s != g_force_splits) continue;
      if (s > 1 && (!allow_split || s > total_kb)) continue;
      if (s > 1 && tiles * s * (128L * bn * 4) > ws_bytes) continue;   // needs the workspace
      const long ctas = tiles * s;
      const int cap = (s <= 2) ? kNumSm : (s == 4 ? 132 : 120);
      const long waves = (ctas + cap - 1) / cap;
      const int kb_per = (total_kb + s - 1) / s;
      const double main = kb_per * (bn == 256 ? 420.0 : bn == 128 ? 195.0 : 112.0);
      const double epi = (s == 1) ? 300.0 + 6.5 * bn
                                  : 8.6 * bn + 900.0 + 350.0 + 0.4 * (128.0 / s) * bn + 300.0;
      const double t = waves * (1400.0 + main + epi);
      if (t < best) { best = t; best_bn = bn; best_s = s; }
    }
  }
  if (best >= 1e30) {  // forced combination not admissible: fall back to no split
    best_bn = valid_bn(g_force_bn) ? g_force_bn : 32;
    best_s = 1;
  }
  *bn_out = best_bn;
  *splits_out = best_s;
}

static void read_force_env() {
  if (g_force_bn < 0) {
    const char* e = getenv("MIXDQ_FORCE_BN");
    g_force_bn = e ? atoi(e) : 0;
    const char* f = getenv("MIXDQ_FORCE_SPLITS");
    if (f) g_force_splits = atoi(f);
  }
}
// persistent kernel (tc_persist.cuh) for multi-wave problems, unless a tile shape is forced
static int persist_bn(int m_tiles, int N, int num_kb, int kind) {
  read_force_env();
  if (g_force_bn > 0 || g_force_splits > 0) return 0;
  return persist_pick_bn(m_tiles, N, num_kb, kind);
}

template <int KIND, bool W4 = false>
static int dispatch_tc(int bn, dim3 grid, const CUtensorMap& a, const CUtensorMap& w,
                       const CUtensorMap& a1, const CUtensorMap& w1, const TcParams& p,
                       cudaStream_t st) {
  switch (bn) {
    case 256: return launch_tc<256, 4, KIND, W4>(grid, a, w, a1, w1, p, st);
    case 128: return launch_tc<128, 3, KIND, W4>(grid, a, w, a1, w1, p, st);   // 2 k-blocks per stage
    case 64: return launch_tc<64, 4, KIND, W4>(grid, a, w, a1, w1, p, st);
    case 32: return launch_tc<32, 4, KIND, W4>(grid, a, w, a1, w1, p, st);
    case 16: return launch_tc<16, 4, KIND, W4>(grid, a, w, a1, w1, p, st);
    default: return MIXDQ_ERR_UNSUPPORTED;
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------
// GEMM (A2)
// ------------------------------------------------------------------------------------------
static int gemm_common(const int8_t* A, int64_t lda, const int8_t* W, const float* p_scale,
                       const float* p_bias0, const float* a_scale, const float* a_zp,
                       const mixdq_half_t* bias, const mixdq_half_t* residual, int64_t ldr,
                       mixdq_half_t* D, int64_t ldd, int M, int N, int K,
                       int32_t* acc_out, cudaStream_t st, bool w4 = false) {
  if (M < 0 || N <= 0 || K <= 0 || !W || !p_scale || !p_bias0) return MIXDQ_ERR_INVALID_ARG;
  if (M == 0) return MIXDQ_OK;
  if (!A || !D || lda < K || ldd < N || (residual && ldr < N)) return MIXDQ_ERR_INVALID_ARG;
  if ((K & (w4 ? 31 : 3)) || (N & 3)) return MIXDQ_ERR_ALIGNMENT;

  const bool tc_ok = !g_force_simt && (K % 16 == 0) && (N % 8 == 0) && (lda % 16 == 0) &&
                     (ldd % 8 == 0) && al16(A) && al16(W) && al16(D) &&
                     (!acc_out || al16(acc_out)) && (!residual || (al16(residual) && ldr % 8 == 0));
  if (!tc_ok) {
    if (residual) return MIXDQ_ERR_ALIGNMENT;   // the fused tail exists on the tcgen05 path only
    SimtGemmArgs g{};
    g.A = A; g.lda = lda; g.W = W; g.K = K;
    g.scale = p_scale; g.bias0 = p_bias0; g.a_scale = a_scale; g.a_zp = a_zp;
    g.bias = reinterpret_cast<const __half*>(bias);
    g.D = reinterpret_cast<__half*>(D); g.ldd = ldd; g.M = M; g.N = N; g.acc_out = acc_out;
    g.w4 = w4 ? 1 : 0;
    g_last_path = w4 ? "simt-w4" : "simt";
    return simt_gemm_launch(g, st) == 0 ? MIXDQ_OK : MIXDQ_ERR_CUDA;
  }

  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int pbn = persist_bn(m_tiles, N, num_kb, KIND_GEMM);
  const int pcs = pbn ? persist_cluster_size(m_tiles) : 1;
  int bn = pbn, splits = 1;
  if (!pbn) pick_tile(m_tiles, N, num_kb, true, st, &bn, &splits);
  CUtensorMap tmA, tmW;
  if (!make_tmap_2d(&tmA, A, K, M, lda, BLOCK_M)) return MIXDQ_ERR_CUDA;
  if (w4 ? !make_tmap_2d_w4(&tmW, W, K, N, bn / pcs) : !make_tmap_2d(&tmW, W, K, N, K, bn / pcs))
    return MIXDQ_ERR_CUDA;
  TcParams p{};
  p.dbg = g_dbg;
  p.dbg_mode = g_dbg_mode;
  p.a_prefetch = a_prefetch_flag();
  p.splits = splits;
  current_ws(st, &p.ws);
  p.M = M; p.N = N; p.num_kb = num_kb;
  p.scale = p_scale; p.bias0 = p_bias0; p.a_scale = a_scale; p.a_zp = a_zp;
  p.bias = reinterpret_cast<const __half*>(bias);
  p.D = reinterpret_cast<__half*>(D); p.ldd = ldd; p.acc_out = acc_out;
  p.residual = reinterpret_cast<const __half*>(residual); p.ldr = ldr;
  if (pbn) {
    p.tiles_m = m_tiles; p.tiles_n = (N + pbn - 1) / pbn;
    CUtensorMap tmD;
    if (!make_tmap_out(&tmD, D, N, M, ldd)) return MIXDQ_ERR_CUDA;
    p.d_tma = 1; p.d_cols = N;
    g_last_path = w4 ? "tcgen05-w4-persist" : "tcgen05-persist";
    return persist_launch(KIND_GEMM, pbn, w4, pcs, tmA, tmW, tmD, p, st);
  }
  dim3 grid(m_tiles, (N + bn - 1) / bn, splits);
  if (w4) {
    g_last_path = splits > 1 ? "tcgen05-w4-splitk" : "tcgen05-w4";
    return dispatch_tc<KIND_GEMM, true>(bn, grid, tmA, tmW, tmA, tmW, p, st);
  }
  g_last_path = splits > 1 ? "tcgen05-splitk" : "tcgen05";
  return dispatch_tc<KIND_GEMM>(bn, grid, tmA, tmW, tmA, tmW, p, st);
}

extern "C" int mixdq_gemm_w8a8_f16(const int8_t* A, int64_t lda, const int8_t* W,
                                   const float* bias0, const float* scale,
                                   const mixdq_half_t* bias, mixdq_half_t* D, int64_t ldd, int M,
                                   int N, int K, int32_t* acc_out, mixdq_stream_t stream) {
  return gemm_common(A, lda, W, scale, bias0, nullptr, nullptr, bias, nullptr, 0, D, ldd, M, N, K,
                     acc_out, static_cast<cudaStream_t>(stream));
}

extern "C" int mixdq_gemm_w8a8_f16_dyn(const int8_t* A, int64_t lda, const int8_t* W,
                                       const float* w_scale, const float* wsum,
                                       const float* a_scale, const float* a_zp,
                                       const mixdq_half_t* bias, mixdq_half_t* D, int64_t ldd,
                                       int M, int N, int K, int32_t* acc_out,
                                       mixdq_stream_t stream) {
  if (!a_scale || !a_zp) return MIXDQ_ERR_INVALID_ARG;
  return gemm_common(A, lda, W, w_scale, wsum, a_scale, a_zp, bias, nullptr, 0, D, ldd, M, N, K,
                     acc_out, static_cast<cudaStream_t>(stream));
}

// ff.net.0.proj + GEGLU (KIND_GEGLU): W / w_scale / wsum / bias rows interleaved in groups of
// 16 value rows followed by their 16 gate rows; Y = [M][N2/2] fp16; min/max -> ws (DynWs::mm)
static int geglu_common(const int8_t* A, int64_t lda, const int8_t* W_il,
                        const float* w_scale_il, const float* wsum_il, const float* a_scale,
                        const float* a_zp, const mixdq_half_t* bias_il, mixdq_half_t* Y,
                        int64_t ldy, int M, int N2, int K, void* ws, mixdq_stream_t stream,
                        bool w4, int8_t* Q = nullptr, int64_t ldq = 0,
                        const float* q_inv = nullptr, const float* q_zp = nullptr) {
  // Q != nullptr: static scales of the consumer — int8 codes [M][N2/2] (row pitch ldq) instead of
  // the fp16 Y + min/max partials
  if (M < 0 || N2 <= 0 || K <= 0 || !W_il || !w_scale_il || !wsum_il || !a_scale || !a_zp || !ws)
    return MIXDQ_ERR_INVALID_ARG;
  if (M == 0) return MIXDQ_OK;
  if (Q) {
    if (!A || lda < K || ldq < N2 / 2 || !q_inv || !q_zp) return MIXDQ_ERR_INVALID_ARG;
    if ((ldq % 16) || !al16(Q)) return MIXDQ_ERR_ALIGNMENT;
    Y = reinterpret_cast<mixdq_half_t*>(Q);   // never written; keeps the pointer checks below simple
    ldy = ldq;
  }
  if (!A || !Y || lda < K || ldy < N2 / 2) return MIXDQ_ERR_INVALID_ARG;
  if ((K % (w4 ? 32 : 16)) || (N2 % 32) || (lda % 16) || (ldy % 8) || !al16(A) || !al16(W_il) ||
      !al16(Y))
    return MIXDQ_ERR_ALIGNMENT;
  if (g_force_simt) return MIXDQ_ERR_UNSUPPORTED;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  // tile width: same wave / mainloop model as pick_tile, with the GEGLU epilogue's ALU cost
  // (~14 ns per tile column: dequant + erff on 8 warps) and a 160-wide tile that puts the
  // batch-1 projection (2 x 64 tiles) on one wave of 128 CTAs
  int bn = 32;
  {
    if (g_force_bn < 0) { int d0, d1; pick_tile(1, 32, 1, false, nullptr, &d0, &d1); }   // reads the env
    const int cands[5] = {256, 160, 128, 64, 32};
    const double main_ns[5] = {440.0, 280.0, 195.0, 112.0, 112.0};
    double best = 1e30;
    for (int i = 0; i < 5; ++i) {
      const int c = cands[i];
      if ((valid_bn(g_force_bn) || g_force_bn == 160) && c != g_force_bn) continue;
      if (c > 32 && c / 2 >= N2) continue;
      const long tiles = static_cast<long>(m_tiles) * ((N2 + c - 1) / c);
      const long waves = (tiles + 147) / 148;
      const double t = waves * (1400.0 + num_kb * main_ns[i] + 300.0 + 14.0 * c);
      if (t < best) { best = t; bn = c; }
    }
  }
  const int pbn = persist_bn(m_tiles, N2, num_kb, KIND_GEGLU);
  const int pcs = pbn ? persist_cluster_size(m_tiles) : 1;
  if (pbn) bn = pbn;
  CUtensorMap tmA, tmW;
  if (!make_tmap_2d(&tmA, A, K, M, lda, BLOCK_M)) return MIXDQ_ERR_CUDA;
  if (w4 ? !make_tmap_2d_w4(&tmW, W_il, K, N2, bn / pcs)
         : !make_tmap_2d(&tmW, W_il, K, N2, K, bn / pcs))
    return MIXDQ_ERR_CUDA;
  TcParams p{};
  p.dbg = g_dbg;
  p.dbg_mode = g_dbg_mode;
  p.a_prefetch = a_prefetch_flag();
  p.splits = 1;
  p.M = M; p.N = N2; p.num_kb = num_kb;
  p.scale = w_scale_il; p.bias0 = wsum_il; p.a_scale = a_scale; p.a_zp = a_zp;
  p.bias = reinterpret_cast<const __half*>(bias_il);
  p.D = reinterpret_cast<__half*>(Y); p.ldd = ldy;
  p.mm_partial = static_cast<DynWs*>(ws)->partial;   // address arithmetic only (device pointer)
  p.q_out = Q; p.ldq = ldq; p.q_inv = q_inv; p.q_zp = q_zp;
  if (pbn) {
    // one min / max partial per persistent CTA
    p.tiles_m = m_tiles; p.tiles_n = (N2 + pbn - 1) / pbn;
    const long groups = static_cast<long>((m_tiles + pcs - 1) / pcs) * p.tiles_n;
    int sms = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    long clusters = sms / pcs;
    if (groups < clusters) clusters = groups;
    partial_count_slot(ws) = static_cast<int>(clusters * pcs);
    CUtensorMap tmD;
    if (Q) {
      tmD = tmA;                               // unused: the codes leave through plain stores
      p.d_tma = 0; p.d_cols = N2 / 2;
    } else {
      if (!make_tmap_out(&tmD, Y, N2 / 2, M, ldy)) return MIXDQ_ERR_CUDA;
      p.d_tma = 1; p.d_cols = N2 / 2;
    }
    g_last_path = w4 ? "tcgen05-w4-geglu-persist" : "tcgen05-geglu-persist";
    return persist_launch(KIND_GEGLU, pbn, w4, pcs, tmA, tmW, tmD, p, static_cast<cudaStream_t>(stream));
  }
  dim3 grid(m_tiles, (N2 + bn - 1) / bn, 1);
  if (static_cast<int64_t>(grid.x) * grid.y > kMaxPartials) return MIXDQ_ERR_UNSUPPORTED;
  partial_count_slot(ws) = static_cast<int>(grid.x * grid.y);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (w4) {
    g_last_path = "tcgen05-w4-geglu";
    switch (bn) {
      case 256: return launch_tc<256, 4, KIND_GEGLU, true>(grid, tmA, tmW, tmA, tmW, p, st);
      case 160: return launch_tc<160, 5, KIND_GEGLU, true>(grid, tmA, tmW, tmA, tmW, p, st);
      case 128: return launch_tc<128, 3, KIND_GEGLU, true>(grid, tmA, tmW, tmA, tmW, p, st);
      case 64: return launch_tc<64, 4, KIND_GEGLU, true>(grid, tmA, tmW, tmA, tmW, p, st);
      default: return launch_tc<32, 4, KIND_GEGLU, true>(grid, tmA, tmW, tmA, tmW, p, st);
    }
  }
  g_last_path = "tcgen05-geglu";
  switch (bn) {
    case 256: return launch_tc<256, 4, KIND_GEGLU>(grid, tmA, tmW, tmA, tmW, p, st);
    case 160: return launch_tc<160, 5, KIND_GEGLU>(grid, tmA, tmW, tmA, tmW, p, st);
    case 128: return launch_tc<128, 3, KIND_GEGLU>(grid, tmA, tmW, tmA, tmW, p, st);
    case 64: return launch_tc<64, 4, KIND_GEGLU>(grid, tmA, tmW, tmA, tmW, p, st);
    default: return launch_tc<32, 4, KIND_GEGLU>(grid, tmA, tmW, tmA, tmW, p, st);
  }
}

extern "C" int mixdq_gemm_w8a8_geglu_f16_dyn(const int8_t* A, int64_t lda, const int8_t* W_il,
                                             const float* w_scale_il, const float* wsum_il,
                                             const float* a_scale, const float* a_zp,
                                             const mixdq_half_t* bias_il, mixdq_half_t* Y,
                                             int64_t ldy, int M, int N2, int K, void* ws,
                                             mixdq_stream_t stream) {
  return geglu_common(A, lda, W_il, w_scale_il, wsum_il, a_scale, a_zp, bias_il, Y, ldy, M, N2, K,
                      ws, stream, false);
}

extern "C" int mixdq_gemm_w4a8_geglu_f16_dyn(const int8_t* A, int64_t lda,
                                             const uint8_t* W_il_packed, const float* w_scale_il,
                                             const float* wsum_il, const float* a_scale,
                                             const float* a_zp, const mixdq_half_t* bias_il,
                                             mixdq_half_t* Y, int64_t ldy, int M, int N2, int K,
                                             void* ws, mixdq_stream_t stream) {
  return geglu_common(A, lda, reinterpret_cast<const int8_t*>(W_il_packed), w_scale_il, wsum_il,
                      a_scale, a_zp, bias_il, Y, ldy, M, N2, K, ws, stream, true);
}

extern "C" int mixdq_gemm_geglu_i8_static(const int8_t* A, int64_t lda, const void* W_il,
                                          int w_bits, const float* w_scale_il,
                                          const float* wsum_il, const float* a_scale,
                                          const float* a_zp, const mixdq_half_t* bias_il,
                                          const float* q_scale_inv, const float* q_zp, int8_t* Q,
                                          int64_t ldq, int M, int N2, int K, void* ws,
                                          mixdq_stream_t stream) {
  if (!Q || (w_bits != 8 && w_bits != 4)) return MIXDQ_ERR_INVALID_ARG;
  return geglu_common(A, lda, static_cast<const int8_t*>(W_il), w_scale_il, wsum_il, a_scale, a_zp,
                      bias_il, nullptr, 0, M, N2, K, ws, stream, w_bits == 4, Q, ldq, q_scale_inv,
                      q_zp);
}

extern "C" int mixdq_gemm_w8a8_f16_dyn_res(const int8_t* A, int64_t lda, const int8_t* W,
                                           const float* w_scale, const float* wsum,
                                           const float* a_scale, const float* a_zp,
                                           const mixdq_half_t* bias, const mixdq_half_t* residual,
                                           int64_t ldr, mixdq_half_t* D, int64_t ldd, int M, int N,
                                           int K, int32_t* acc_out, mixdq_stream_t stream) {
  if (!a_scale || !a_zp) return MIXDQ_ERR_INVALID_ARG;
  return gemm_common(A, lda, W, w_scale, wsum, a_scale, a_zp, bias, residual, ldr, D, ldd, M, N, K,
                     acc_out, static_cast<cudaStream_t>(stream));
}

// W4A8: packed signed 4-bit weights [N][K/2] (even k in the high nibble), unpacked on the fly by the
// converter warps of tc_i8_kernel<..., W4 = true>
extern "C" int mixdq_gemm_w4a8_f16(const int8_t* A, int64_t lda, const uint8_t* W_packed,
                                   const float* bias0, const float* scale,
                                   const mixdq_half_t* bias, mixdq_half_t* D, int64_t ldd, int M,
                                   int N, int K, int32_t* acc_out, mixdq_stream_t stream) {
  return gemm_common(A, lda, reinterpret_cast<const int8_t*>(W_packed), scale, bias0, nullptr,
                     nullptr, bias, nullptr, 0, D, ldd, M, N, K, acc_out,
                     static_cast<cudaStream_t>(stream), true);
}

extern "C" int mixdq_gemm_w4a8_f16_dyn_res(const int8_t* A, int64_t lda, const uint8_t* W_packed,
                                           const float* w_scale, const float* wsum,
                                           const float* a_scale, const float* a_zp,
                                           const mixdq_half_t* bias, const mixdq_half_t* residual,
                                           int64_t ldr, mixdq_half_t* D, int64_t ldd, int M, int N,
                                           int K, int32_t* acc_out, mixdq_stream_t stream) {
  if (!a_scale || !a_zp) return MIXDQ_ERR_INVALID_ARG;
  return gemm_common(A, lda, reinterpret_cast<const int8_t*>(W_packed), w_scale, wsum, a_scale,
                     a_zp, bias, residual, ldr, D, ldd, M, N, K, acc_out,
                     static_cast<cudaStream_t>(stream), true);
}

// ------------------------------------------------------------------------------------------
// conv (A3 + A4)
// ------------------------------------------------------------------------------------------
static int conv_common(const int8_t* x, int64_t x_cpitch, const int8_t* w, const float* scale,
                       const float* wsum_krs, const float* bias0_k, const float* zp,
                       const float* a_scale, const mixdq_half_t* bias,
                       const mixdq_half_t* chan_add, int64_t ldca, const mixdq_half_t* residual,
                       mixdq_half_t* y, int N, int H, int W, int C, int K, int R, int S,
                       int stride, int pad, int32_t* acc_out, cudaStream_t st, bool w4 = false) {
  if (N < 0 || H <= 0 || W <= 0 || C <= 0 || K <= 0 || R <= 0 || S <= 0 || stride <= 0 ||
      pad < 0 || !w || !scale)
    return MIXDQ_ERR_INVALID_ARG;
  if (pad > 0 ? (!wsum_krs || !zp) : !bias0_k) return MIXDQ_ERR_INVALID_ARG;
  if (x_cpitch < C) return MIXDQ_ERR_INVALID_ARG;
  if ((C & 3) || (K & 3)) return MIXDQ_ERR_ALIGNMENT;
  const int P = (H + 2 * pad - R) / stride + 1;
  const int Q = (W + 2 * pad - S) / stride + 1;
  if (P <= 0 || Q <= 0) return MIXDQ_ERR_INVALID_ARG;
  if (N == 0) return MIXDQ_OK;
  if (!x || !y) return MIXDQ_ERR_INVALID_ARG;

  // stride 2 (the down-samplers): the A box is fetched with TMA element strides (traversal
  // stride 2 along W and H), so the tile rows are still consecutive OUTPUT pixels
  const bool geom_ok = (stride == 1 || stride == 2) &&
                       (pad == 0 || (pad == 1 && R == 3 && S == 3));
  const bool tc_ok = !g_force_simt && geom_ok && (C % (w4 ? 32 : 16) == 0) && (x_cpitch % 16 == 0) &&
                     (K % 8 == 0) && al16(x) && al16(w) && al16(y) && (!acc_out || al16(acc_out)) &&
                     (!chan_add || (al16(chan_add) && ldca % 8 == 0 && ldca >= K)) &&
                     (!residual || al16(residual));
  if (!tc_ok) {
    // dynamic scalars, the fused tails and packed 4-bit weights exist on the tcgen05 path only
    if (a_scale || chan_add || residual || w4) return MIXDQ_ERR_ALIGNMENT;
    SimtConvArgs c{};
    c.x = x; c.x_cpitch = x_cpitch; c.w = w; c.scale = scale;
    c.wsum_krs = pad > 0 ? wsum_krs : nullptr; c.bias0_k = bias0_k; c.zp = zp;
    c.bias = reinterpret_cast<const __half*>(bias); c.y = reinterpret_cast<__half*>(y);
    c.N = N; c.H = H; c.W = W; c.C = C; c.K = K; c.R = R; c.S = S; c.stride = stride; c.pad = pad;
    c.P = P; c.Q = Q; c.acc_out = acc_out;
    g_last_path = "simt";
    return simt_conv_launch(c, st) == 0 ? MIXDQ_OK : MIXDQ_ERR_CUDA;
  }

  // A box: boxN x boxH x boxW output pixels (<= 128 rows of the UMMA tile)
  // (output rows wider than one tile — latents beyond 1024 px — are tiled along q as well)
  int boxW = Q < BLOCK_M ? Q : BLOCK_M, boxH = BLOCK_M / boxW;
  if (boxH > P) boxH = P;
  int boxN = 1;
  if (boxH == P && boxW == Q) {
    boxN = BLOCK_M / (boxW * boxH); if (boxN > N) boxN = N; if (boxN < 1) boxN = 1;
  }
  const int tilesQ = (Q + boxW - 1) / boxW, tilesP = (P + boxH - 1) / boxH,
            tilesN = (N + boxN - 1) / boxN;
  const int m_tiles = tilesQ * tilesP * tilesN;
  const int kb_per_tap = (C + BLOCK_K - 1) / BLOCK_K;
  int pbn = persist_bn(m_tiles, K, R * S * kb_per_tap, KIND_CONV);
  // 3x3 convolutions on the persistent CTA-pair kernel: 160-wide tiles with one haloed A box for
  // the three vertical taps (fewer operand bytes per MMA than any plain tile width); it has its
  // own, lower size threshold
  read_force_env();
  const bool shape_forced = g_force_bn > 0 || g_force_splits > 0;
  const bool halo = !shape_forced && (pbn || persist_halo_wanted(m_tiles)) &&
                    persist_halo_ok(160, w4, persist_cluster_size(m_tiles), R, S, pad, stride, boxW,
                                    boxH, boxN);
  if (halo) pbn = 160;
  const int pcs = pbn ? persist_cluster_size(m_tiles) : 1;
  int bn = pbn, splits = 1;
  if (!pbn) pick_tile(m_tiles, K, R * S * kb_per_tap, true, st, &bn, &splits);

  CUtensorMap tmA, tmW;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(N)};
    uint64_t strides[3] = {static_cast<uint64_t>(x_cpitch), static_cast<uint64_t>(x_cpitch) * W,
                           static_cast<uint64_t>(x_cpitch) * W * H};
    uint32_t box[4] = {BLOCK_K, static_cast<uint32_t>(boxW * stride),
                       static_cast<uint32_t>((halo ? boxH + 2 : boxH) * stride),
                       static_cast<uint32_t>(boxN)};
    uint32_t estr[4] = {1, static_cast<uint32_t>(stride), static_cast<uint32_t>(stride), 1};
    if (!make_tmap(&tmA, x, 4, dims, strides, box, estr)) return MIXDQ_ERR_CUDA;
  }
  {
    // KRSC (W4: C/2 packed bytes per tap, 64-byte boxes, unswizzled)
    const uint64_t cb = w4 ? C / 2 : C;
    uint64_t dims[3] = {cb, static_cast<uint64_t>(R) * S, static_cast<uint64_t>(K)};
    uint64_t strides[2] = {cb, cb * R * S};
    uint32_t box[3] = {static_cast<uint32_t>(w4 ? BLOCK_K / 2 : BLOCK_K), 1,
                       static_cast<uint32_t>(bn / pcs)};
    if (!make_tmap(&tmW, w, 3, dims, strides, box, nullptr, !w4)) return MIXDQ_ERR_CUDA;
  }
  TcParams p{};
  p.dbg = g_dbg;
  p.dbg_mode = g_dbg_mode;
  p.a_prefetch = a_prefetch_flag();
  p.splits = splits;
  current_ws(st, &p.ws);
  p.M = N * P * Q; p.N = K;
  p.kb_per_tap = kb_per_tap;
  p.num_kb = R * S * p.kb_per_tap;
  p.S = S; p.pad = pad; p.stride = stride; p.NB = N; p.H = H; p.W = W; p.P = P; p.Q = Q;
  p.boxW = boxW; p.boxH = boxH; p.boxN = boxN; p.tilesQ = tilesQ; p.tilesP = tilesP;
  p.a_tx_bytes = static_cast<uint32_t>(boxW) * (halo ? boxH + 2 : boxH) * boxN * BLOCK_K;
  p.has_table = pad > 0 ? 1 : 0;
  p.scale = scale; p.bias0 = pad > 0 ? wsum_krs : bias0_k; p.a_zp = zp; p.a_scale = a_scale;
  p.bias = reinterpret_cast<const __half*>(bias);
  p.D = reinterpret_cast<__half*>(y); p.ldd = K; p.acc_out = acc_out;
  p.chan_add = reinterpret_cast<const __half*>(chan_add); p.ldca = ldca;
  p.rows_per_img = static_cast<int64_t>(P) * Q;
  p.residual = reinterpret_cast<const __half*>(residual); p.ldr = K;
  if (pbn) {
    p.tiles_m = m_tiles; p.tiles_n = (K + pbn - 1) / pbn;
    // TMA-store epilogue: the 128 rows of every tile must be 128 CONTIGUOUS output pixels
    // (whole output rows, no ragged tile inside an image); else the staged copy-out
    const bool rows_contig = (boxW == Q) && (boxW * boxH * boxN == BLOCK_M) &&
                             (boxN > 1 ? (boxH == P) : (P % boxH == 0));
    CUtensorMap tmD;
    if (!make_tmap_out(&tmD, y, K, static_cast<uint64_t>(N) * P * Q, K)) return MIXDQ_ERR_CUDA;
    p.d_tma = rows_contig ? 1 : 0; p.d_cols = K;
    if (halo) {
      g_last_path = "tcgen05-persist-halo";
      return persist_launch_conv_halo(tmA, tmW, tmD, p, st);
    }
    g_last_path = w4 ? "tcgen05-w4-persist" : "tcgen05-persist";
    return persist_launch(KIND_CONV, pbn, w4, pcs, tmA, tmW, tmD, p, st);
  }
  dim3 grid(m_tiles, (K + bn - 1) / bn, splits);
  if (w4) {
    g_last_path = splits > 1 ? "tcgen05-w4-splitk" : "tcgen05-w4";
    return dispatch_tc<KIND_CONV, true>(bn, grid, tmA, tmW, tmA, tmW, p, st);
  }
  g_last_path = splits > 1 ? "tcgen05-splitk" : "tcgen05";
  return dispatch_tc<KIND_CONV>(bn, grid, tmA, tmW, tmA, tmW, p, st);
}

// W4A8 convolutions: w_krsc_packed = uint8 [K][R][S][C/2], two signed 4-bit codes per byte along
// C, even c in the high nibble. tcgen05 path only (C % 32 == 0, K % 8 == 0, 1x1 or 3x3/pad 1,
// stride 1 or 2): MIXDQ_ERR_ALIGNMENT otherwise — such layers keep one code per int8.
extern "C" int mixdq_conv_w4a8_f16(const int8_t* x, int64_t x_cpitch, const uint8_t* w_packed,
                                   const float* scale, const float* wsum_krs,
                                   const float* bias0_k, const float* zp,
                                   const mixdq_half_t* bias, mixdq_half_t* y, int N, int H, int W,
                                   int C, int K, int R, int S, int stride, int pad,
                                   int32_t* acc_out, mixdq_stream_t stream) {
  return conv_common(x, x_cpitch, reinterpret_cast<const int8_t*>(w_packed), scale, wsum_krs,
                     bias0_k, zp, nullptr, bias, nullptr, 0, nullptr, y, N, H, W, C, K, R, S,
                     stride, pad, acc_out, static_cast<cudaStream_t>(stream), true);
}

extern "C" int mixdq_conv_w4a8_f16_dyn(const int8_t* x, int64_t x_cpitch, const uint8_t* w_packed,
                                       const float* w_scale, const float* wsum_krs,
                                       const float* wsum_k, const float* a_scale,
                                       const float* a_zp, const mixdq_half_t* bias,
                                       const mixdq_half_t* chan_add, int64_t ldca,
                                       const mixdq_half_t* residual, mixdq_half_t* y, int N, int H,
                                       int W, int C, int K, int R, int S, int stride, int pad,
                                       int32_t* acc_out, mixdq_stream_t stream) {
  if (!a_scale || !a_zp) return MIXDQ_ERR_INVALID_ARG;
  return conv_common(x, x_cpitch, reinterpret_cast<const int8_t*>(w_packed), w_scale, wsum_krs,
                     wsum_k, a_zp, a_scale, bias, chan_add, ldca, residual, y, N, H, W, C, K, R, S,
                     stride, pad, acc_out, static_cast<cudaStream_t>(stream), true);
}

extern "C" int mixdq_conv_w8a8_f16(const int8_t* x, int64_t x_cpitch, const int8_t* w,
                                   const float* scale, const float* wsum_krs,
                                   const float* bias0_k, const float* zp,
                                   const mixdq_half_t* bias, mixdq_half_t* y, int N, int H, int W,
                                   int C, int K, int R, int S, int stride, int pad,
                                   int32_t* acc_out, mixdq_stream_t stream) {
  return conv_common(x, x_cpitch, w, scale, wsum_krs, bias0_k, zp, nullptr, bias, nullptr, 0,
                     nullptr, y, N, H, W, C, K, R, S, stride, pad, acc_out,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int mixdq_conv_w8a8_f16_dyn(const int8_t* x, int64_t x_cpitch, const int8_t* w,
                                       const float* w_scale, const float* wsum_krs,
                                       const float* wsum_k, const float* a_scale,
                                       const float* a_zp, const mixdq_half_t* bias,
                                       const mixdq_half_t* chan_add, int64_t ldca,
                                       const mixdq_half_t* residual, mixdq_half_t* y, int N, int H,
                                       int W, int C, int K, int R, int S, int stride, int pad,
                                       int32_t* acc_out, mixdq_stream_t stream) {
  if (!a_scale || !a_zp) return MIXDQ_ERR_INVALID_ARG;
  return conv_common(x, x_cpitch, w, w_scale, wsum_krs, wsum_k, a_zp, a_scale, bias, chan_add, ldca,
                     residual, y, N, H, W, C, K, R, S, stride, pad, acc_out,
                     static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------
// split 1x1 shortcut (A6)
// ------------------------------------------------------------------------------------------
static int split_common(const int8_t* xa, int64_t lda, const int8_t* wa, int Ca,
                        const float* bias0_a, const float* scale_a, const float* a_scale_a,
                        const float* a_zp_a, const int8_t* xb, int64_t ldb, const int8_t* wb,
                        int Cb, const float* bias0_b, const float* scale_b,
                        const float* a_scale_b, const float* a_zp_b, const mixdq_half_t* bias,
                        const mixdq_half_t* residual, int64_t ldr, mixdq_half_t* y, int64_t ldy,
                        int M, int K, cudaStream_t st) {
  if (M < 0 || K <= 0 || Ca <= 0 || Cb <= 0 || !wa || !wb || !bias0_a || !scale_a || !bias0_b ||
      !scale_b)
    return MIXDQ_ERR_INVALID_ARG;
  if (M == 0) return MIXDQ_OK;
  if (!xa || !xb || !y || lda < Ca || ldb < Cb || ldy < K) return MIXDQ_ERR_INVALID_ARG;
  if ((Ca & 3) || (Cb & 3) || (K & 3)) return MIXDQ_ERR_ALIGNMENT;
  const bool tc_ok = !g_force_simt && (Ca % 16 == 0) && (Cb % 16 == 0) && (K % 8 == 0) &&
                     (lda % 16 == 0) && (ldb % 16 == 0) && (ldy % 8 == 0) && al16(xa) && al16(xb) &&
                     al16(wa) && al16(wb) && al16(y);
  if (!tc_ok) {
    if (a_scale_a || a_scale_b || residual) return MIXDQ_ERR_ALIGNMENT;  // tcgen05 path only
    SimtGemmArgs g{};
    g.A = xa; g.lda = lda; g.W = wa; g.K = Ca;
    g.A1 = xb; g.lda1 = ldb; g.W1 = wb; g.K1 = Cb;
    g.scale = scale_a; g.bias0 = bias0_a; g.scale1 = scale_b; g.bias0_1 = bias0_b;
    g.bias = reinterpret_cast<const __half*>(bias);
    g.D = reinterpret_cast<__half*>(y); g.ldd = ldy; g.M = M; g.N = K;
    g_last_path = "simt";
    return simt_gemm_launch(g, st) == 0 ? MIXDQ_OK : MIXDQ_ERR_CUDA;
  }
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  int bn, splits_unused;
  pick_tile(m_tiles, K, (Ca + BLOCK_K - 1) / BLOCK_K + (Cb + BLOCK_K - 1) / BLOCK_K, false, st,
            &bn, &splits_unused);
  if (bn > 128) bn = 128;  // two accumulators: 2 x BN TMEM columns, keep smem params small
  CUtensorMap tmA, tmW, tmA1, tmW1;
  if (!make_tmap_2d(&tmA, xa, Ca, M, lda, BLOCK_M) || !make_tmap_2d(&tmW, wa, Ca, K, Ca, bn) ||
      !make_tmap_2d(&tmA1, xb, Cb, M, ldb, BLOCK_M) || !make_tmap_2d(&tmW1, wb, Cb, K, Cb, bn))
    return MIXDQ_ERR_CUDA;
  TcParams p{};
  p.dbg = g_dbg;
  p.dbg_mode = g_dbg_mode;
  p.a_prefetch = a_prefetch_flag();
  p.splits = 1;
  p.M = M; p.N = K;
  p.num_kb = (Ca + BLOCK_K - 1) / BLOCK_K;
  p.num_kb1 = (Cb + BLOCK_K - 1) / BLOCK_K;
  p.scale = scale_a; p.bias0 = bias0_a; p.scale1 = scale_b; p.bias0_1 = bias0_b;
  p.a_scale = a_scale_a; p.a_zp = a_zp_a; p.a_scale1 = a_scale_b; p.a_zp1 = a_zp_b;
  p.bias = reinterpret_cast<const __half*>(bias);
  p.D = reinterpret_cast<__half*>(y); p.ldd = ldy;
  p.residual = reinterpret_cast<const __half*>(residual); p.ldr = ldr;
  dim3 grid(m_tiles, (K + bn - 1) / bn);
  g_last_path = "tcgen05";
  return dispatch_tc<KIND_SPLIT>(bn, grid, tmA, tmW, tmA1, tmW1, p, st);
}

extern "C" int mixdq_conv1x1_split_w8a8_f16(const int8_t* xa, int64_t lda, const int8_t* wa, int Ca,
                                            const float* bias0_a, const float* scale_a,
                                            const int8_t* xb, int64_t ldb, const int8_t* wb, int Cb,
                                            const float* bias0_b, const float* scale_b,
                                            const mixdq_half_t* bias, mixdq_half_t* y, int64_t ldy,
                                            int M, int K, mixdq_stream_t stream) {
  return split_common(xa, lda, wa, Ca, bias0_a, scale_a, nullptr, nullptr, xb, ldb, wb, Cb, bias0_b,
                      scale_b, nullptr, nullptr, bias, nullptr, 0, y, ldy, M, K,
                      static_cast<cudaStream_t>(stream));
}

extern "C" int mixdq_conv1x1_split_w8a8_f16_dyn(
    const int8_t* xa, int64_t lda, const int8_t* wa, int Ca, const float* wsum_a,
    const float* w_scale_a, const float* a_scale_a, const float* a_zp_a, const int8_t* xb,
    int64_t ldb, const int8_t* wb, int Cb, const float* wsum_b, const float* w_scale_b,
    const float* a_scale_b, const float* a_zp_b, const mixdq_half_t* bias,
    const mixdq_half_t* residual, int64_t ldr, mixdq_half_t* y, int64_t ldy, int M, int K,
    mixdq_stream_t stream) {
  if (!a_scale_a || !a_zp_a || !a_scale_b || !a_zp_b) return MIXDQ_ERR_INVALID_ARG;
  return split_common(xa, lda, wa, Ca, wsum_a, w_scale_a, a_scale_a, a_zp_a, xb, ldb, wb, Cb, wsum_b,
                      w_scale_b, a_scale_b, a_zp_b, bias, residual, ldr, y, ldy, M, K,
                      static_cast<cudaStream_t>(stream));
}
