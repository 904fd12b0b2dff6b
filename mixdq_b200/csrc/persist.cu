// persist.cu — instantiations and launcher of the persistent tcgen05 contraction kernels.
#include <stdlib.h>

#include "../../include/mixdq_b200.h"
#include "persist.h"
#include "tc_persist.cuh"

namespace mixdq {

static int g_persist_mode = -1;     // env MIXDQ_PERSIST: 0 = never, 1 = heuristic (default),
                                    // 2 = whenever the shape is supported (tests)
static int g_persist_cs = -1;       // env MIXDQ_PERSIST_CS: 1 / 2 (default 2: cta_group::2 CTA pairs)
static int g_persist_bn = 0;        // tuning hook: force this tile width (0 = cost model)
static void read_env() {
  if (g_persist_mode < 0) {
    const char* e = getenv("MIXDQ_PERSIST");
    g_persist_mode = e ? atoi(e) : 1;
    const char* c = getenv("MIXDQ_PERSIST_CS");
    g_persist_cs = c ? atoi(c) : 2;
    if (g_persist_cs != 1) g_persist_cs = 2;
  }
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// When, and with which tile width. Fitted to tools/tune_persist.py (every SDXL layer shape at
// batch 1 and 8, one-tile-per-CTA kernel vs the persistent kernel at each width, B200):
//   * the persistent kernel wins once the problem has >= ~120 tiles of 128 x 128 (a wave of work for
//     every SM); below that the one-tile kernel with its narrow tiles / split-K is faster;
//   * time ~ 2.5 us + rounds x max(k-blocks x t_kb, t_epi) + t_epi, where per k-block a CTA of a
//     pair is bound by the L2 -> SM fabric (~84 GB/s per SM when all 148 stream: 16 KB of A + its
//     half of W) rather than by the tensor pipe: t_kb = 0.38 / 0.31 / 0.29 us for 256 / 160 / 128
//     wide tiles (measured 0.39 us on M=8192 N=K=2560), and the epilogue drains a 128 x 256 tile
//     in ~2.9 us (1.9 / 1.5 us for 160 / 128);
//   * 128-wide tiles never win for convolutions.
int persist_pick_bn(int m_tiles, int N, int num_kb, int kind) {
  read_env();
  if (g_persist_mode == 0) return 0;
  const int sms = num_sms();
  // (M = 256 problems: the batch-1 GEGLU projection measured 9.9 us persistent vs 9.3 us one-tile
  // inside the UNet graph although it wins in isolation)
  // (one m-tile and >= ~4 waves of weight tiles — SDXL's hoisted cross-attention K / V projection,
  // M = 77, N = 166 400, K = 2048, 341 MB of weights — is a pure weight stream: persistent single
  // CTAs walk contiguous tile ranges without the wave tail, 87 -> 63 us = 3.9 -> 5.4 TB/s,
  // tools/kv_gemm_bench.py)
  const long tiles128 = static_cast<long>(m_tiles) * ((N + 127) / 128);
  const bool weight_stream = m_tiles == 1 && tiles128 >= 600;
  if (g_persist_mode == 1 && !weight_stream && (m_tiles < 4 || tiles128 < 120))
    return 0;
  const int cands[3] = {256, 160, 128};
  const double t_kb[3] = {0.38, 0.31, 0.29};
  const double t_epi[3] = {2.9, 1.9, 1.5};
  double best = 1e30;
  int best_bn = 0;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (g_persist_bn > 0 && bn != g_persist_bn) continue;
    if (kind == KIND_GEGLU && (bn == 128 || (N % bn))) continue;   // whole 32-column GEGLU groups
    if (kind == KIND_CONV && bn == 128 && g_persist_bn == 0) continue;
    const long tiles = static_cast<long>(m_tiles) * ((N + bn - 1) / bn);
    const long rounds = (tiles + sms - 1) / sms;
    const double main = num_kb * t_kb[i];
    const double t = rounds * (main > t_epi[i] ? main : t_epi[i]) + t_epi[i];
    if (t < best) { best = t; best_bn = bn; }
  }
  return best_bn;
}

void persist_force_bn(int bn) { g_persist_bn = bn; }
static int g_halo = -1;             // env MIXDQ_CONV_HALO / mixdq_debug_set_conv_halo: 0 = off
void persist_set_halo(int on) { g_halo = on ? 1 : 0; }

void persist_set_mode(int mode, int cs) {
  read_env();
  if (mode >= 0) g_persist_mode = mode;
  if (cs == 1 || cs == 2) g_persist_cs = cs;
}

int persist_cluster_size(int m_tiles) {
  read_env();
  return (g_persist_cs == 2 && m_tiles >= 2) ? 2 : 1;
}

template <int BN, int STAGES, int KIND, bool W4, int CS, bool HALO = false>
static int launch(const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& d, const TcParams& p,
                  cudaStream_t st) {
  using L = TpSmem<BN, STAGES, KIND, W4, CS, HALO>;
  auto kern = tc_i8_persist_kernel<BN, STAGES, KIND, W4, CS, HALO>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) !=
        cudaSuccess)
      return MIXDQ_ERR_CUDA;
    attr_set = true;
  }
  const int m_groups = (p.tiles_m + CS - 1) / CS;
  const long groups = static_cast<long>(m_groups) * p.tiles_n;
  long clusters = num_sms() / CS;
  if (groups < clusters) clusters = groups;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * CS));
  cfg.blockDim = dim3(W4 ? TP_THREADS_W4 : TP_THREADS);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, a, w, d, p) == cudaSuccess ? MIXDQ_OK : MIXDQ_ERR_CUDA;
}

// ring depth: what fits 227 KB next to the staging tiles and the per-column operands. A CTA of a
// pair (CS = 2) stages only half of the W rows, so its ring is 1.3-1.5x deeper.
template <int BN, int KIND, int CS>
struct TpStages {
  static constexpr bool CONV = KIND == KIND_CONV;          // + the 16-class border table
  static constexpr int value =
      CS == 1 ? (BN == 256 ? 3 : BN == 160 ? (CONV ? 4 : 5) : (CONV ? 5 : 6))
              : (BN == 256 ? 5 : BN == 160 ? 6 : 7);
};

template <int KIND, bool W4, int CS>
static int by_bn(int bn, const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& d,
                 const TcParams& p, cudaStream_t st) {
  switch (bn) {
    case 256: return launch<256, TpStages<256, KIND, CS>::value, KIND, W4, CS>(a, w, d, p, st);
    case 160: return launch<160, TpStages<160, KIND, CS>::value, KIND, W4, CS>(a, w, d, p, st);
    case 128: if constexpr (KIND != KIND_GEGLU)
                return launch<128, TpStages<128, KIND, CS>::value, KIND, W4, CS>(a, w, d, p, st);
              return MIXDQ_ERR_UNSUPPORTED;
    default: return MIXDQ_ERR_UNSUPPORTED;
  }
}

template <int KIND>
static int by_flags(int bn, bool w4, int cs, const CUtensorMap& a, const CUtensorMap& w,
                    const CUtensorMap& d, const TcParams& p, cudaStream_t st) {
  if (w4) return cs == 2 ? by_bn<KIND, true, 2>(bn, a, w, d, p, st) : by_bn<KIND, true, 1>(bn, a, w, d, p, st);
  return cs == 2 ? by_bn<KIND, false, 2>(bn, a, w, d, p, st) : by_bn<KIND, false, 1>(bn, a, w, d, p, st);
}

// 3x3 / pad 1 / stride 1 convolutions on CTA pairs with 160-wide tiles: one haloed A box per
// (s, channel block) serves the three vertical taps (tc_persist.cuh, TpSmem::HALO). Env
// MIXDQ_CONV_HALO=0 switches it off (A/B timing).
bool persist_halo_ok(int bn, bool w4, int cs, int R, int S, int pad, int stride, int boxW, int boxH,
                     int boxN) {
  if (g_halo < 0) { const char* e = getenv("MIXDQ_CONV_HALO"); g_halo = e ? atoi(e) : 1; }
  if (g_persist_bn > 0 && g_persist_bn != 160) return false;      // another width is being forced
  return g_halo && bn == 160 && !w4 && cs == 2 && R == 3 && S == 3 && pad == 1 && stride == 1 &&
         boxN == 1 && (boxW % 8) == 0 && (boxH + 2) * boxW <= 2 * BLOCK_M;
}

// HALO convolutions win from 4 m-tiles on (tools/tune_persist.py conv3x3, batch 1..8: 9.9 vs 11.9 us
// at M=4096 N=320, 15.0 vs 41.8 us at M=4096 N=640, 24.3 vs 72.0 us at M=2048 N=1280 K=11520; the
// M=256 layers stay on the one-tile split-K kernel, 13.9 vs 22.6 us)
bool persist_halo_wanted(int m_tiles) {
  read_env();
  return g_persist_mode == 2 || (g_persist_mode == 1 && m_tiles >= 4);
}

int persist_launch_conv_halo(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmD,
                             TcParams p, cudaStream_t st) {
  return launch<160, 3, KIND_CONV, false, 2, true>(tmA, tmW, tmD, p, st);
}

int persist_launch(int kind, int bn, bool w4, int cs, const CUtensorMap& tmA,
                   const CUtensorMap& tmW, const CUtensorMap& tmD, TcParams p, cudaStream_t st) {
  switch (kind) {
    case KIND_GEMM: return by_flags<KIND_GEMM>(bn, w4, cs, tmA, tmW, tmD, p, st);
    case KIND_CONV: return by_flags<KIND_CONV>(bn, w4, cs, tmA, tmW, tmD, p, st);
    case KIND_GEGLU: return by_flags<KIND_GEGLU>(bn, w4, cs, tmA, tmW, tmD, p, st);
    default: return MIXDQ_ERR_UNSUPPORTED;
  }
}

}  // namespace mixdq
