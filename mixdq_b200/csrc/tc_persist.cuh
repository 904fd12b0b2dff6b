// tc_persist.cuh — persistent form of the tcgen05 INT8 contraction for TENSOR-BOUND problems
// (batch >= 8 UNet layers, the per-layer sweep): many output tiles per SM.
//
// tc_i8_kernel (tc_kernel.cuh) runs one 128 x BN tile per CTA: its prologue (barrier / TMEM setup,
// first operands ~1.4 us) and its epilogue (TMEM -> dequant -> fp16 -> global, 1-2 us) are not
// overlapped with any tensor work, which caps it at 25-55 % of the INT8 peak on multi-wave
// problems (profiles/r01_tops_sweep.txt). Here each CTA is PERSISTENT and walks a strided list of
// tiles with three decoupled pipelines:
//
//   warp 0      TMA producer   : A / W k-blocks of tile after tile through ONE continuous
//                                STAGES-deep ring (the ring never drains between tiles)
//   warp 1      MMA issuer     : tcgen05.mma kind::i8 128 x BN x 32 into one of TWO TMEM
//                                accumulator slots (2 x 256 columns = the whole TMEM)
//   warps 2..9  epilogue       : drain slot s (tcgen05.ld -> dequant -> fp16 -> per-warp staging
//                                -> coalesced 16 B stores, fused tails) WHILE the issuer fills
//                                slot s^1 with the next tile
//   [W4] warps 10..13 converter: expand the packed 4-bit weight k-blocks in place (see
//                                tc_kernel.cuh); the epilogue warps are busy here, so the nibble
//                                expansion has warps of its own
//
// CTA pair (CS = 2): two CTAs of one TPC run ONE 256 x BN tile with tcgen05.mma.cta_group::2. Each
// CTA stages its own 128 A rows and HALF of the W rows (BN / 2) — the tensor cores of the pair
// read the other half from the peer's shared memory — so a k-block costs 16 KB (A) + BN / 2 x 128 B
// of L2 -> SM traffic and of ring space per SM instead of 16 KB + BN x 128 B: 1.5x fewer operand
// bytes for BN = 256 and a 6-deep instead of a 4-deep ring (the L2 -> SM fabric and the ring depth,
// not the tensor pipe, bound the single-CTA form: tools/persist_modes.py). Protocol (as CUTLASS's
// 2-SM kernels): both CTAs' TMA loads complete on the LEADER's full barrier (cp.async.bulk.tensor
// .cta_group::2), the leader's single MMA thread issues for both SMs and its tcgen05.commit
// multicasts to the empty / accumulator-full barriers of both CTAs; both CTAs' epilogue warps
// arrive on the leader's accumulator-empty barrier. A first version of CS = 2 only multicast the
// W tile to two independent cta_group::1 CTAs: same speed as CS = 1 (profiles/README.md).
//
// Same operand layouts, tensor maps, KIND semantics (GEMM / CONV / GEGLU), epilogue arithmetic and
// TcParams as tc_i8_kernel; p.tiles_m / p.tiles_n describe the tile grid.
#pragma once
#include "tc_kernel.cuh"

namespace mixdq {

// 16 epilogue warps = 4 per TMEM lane quarter (a warp may only read the 32 lanes 32 * (warp % 4)),
// each draining a quarter of the tile's columns in 16-column chunks. With 8 warps (2 per scheduler)
// the epilogue's dependent chain (tcgen05.ld -> I2F -> dequant -> staging -> store) issued one
// instruction per ~4 cycles and took 5.2 us per 128 x 256 tile, more than the tile's MMAs for
// K <= 2560 (ncu: tensor pipe 22 % active on M=8192 N=5120 K=640, profiles/README.md).
constexpr int TP_EPI_WARPS = 16;
constexpr int TP_EPI_PARTS = TP_EPI_WARPS / 4;             // column partitions of a tile
constexpr int TP_THREADS = 32 * (2 + TP_EPI_WARPS);        // 576
constexpr int TP_CONV_WARPS = 4;
constexpr int TP_THREADS_W4 = TP_THREADS + 32 * TP_CONV_WARPS;   // 448
constexpr int TP_SLOT_COLS = 256;                           // TMEM columns per accumulator slot

template <int BN, int STAGES, int KIND, bool W4, int CS = 1, bool HALO = false>
struct TpSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K;
  static constexpr int W_ROWS = BN / CS;                     // W rows staged by this CTA
  static constexpr int W_BYTES = W_ROWS * BLOCK_K;
  // HALO (3x3 / pad 1 / stride 1 convolutions): a ring stage holds ONE A box with a one-row halo
  // above and below — (boxH + 2) x boxW pixels <= 2 x 128 rows — and the W tiles of the THREE
  // vertical taps (r = 0, 1, 2) of one (s, channel block): the taps read the same pixels shifted
  // by whole image rows, i.e. by boxW x 128 B = a multiple of the 1024 B swizzle atom, so they are
  // three descriptor offsets into one box instead of three boxes through the L2 -> SM fabric
  static constexpr int A_STAGE = HALO ? 2 * A_BYTES : A_BYTES;
  static constexpr int W_STAGE = HALO ? 3 * W_BYTES : W_BYTES;
  // accumulator columns per chunk (GEGLU: 16 value + 16 gate columns -> 16 outputs)
  static constexpr int CH = (KIND == KIND_GEGLU) ? 32 : 16;
  // fp16 staging row: 16 outputs = two 16-byte halves, stored XOR-swizzled by bit 2 of the row so
  // that both the row-per-lane writes and the two-lanes-per-row reads are bank-conflict free
  static constexpr int OUT_PITCH = 32;
  // one warp: two staging tiles of 32 rows (HALO: one — its ring needs the space, and its long
  // K loops hide a store that waits for the previous one to have read the tile)
  static constexpr int OUT_BUFS = HALO ? 1 : 2;
  static constexpr int OUT_WARP = OUT_BUFS * 32 * OUT_PITCH;
  static constexpr int TAB_PITCH = BN + 4;
  static constexpr int TAB_FLOATS = (KIND == KIND_CONV) ? 16 * TAB_PITCH : 0;
  static constexpr int PARAM_FLOATS = 3 * BN + TAB_FLOATS;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = OFF_A + STAGES * A_STAGE;
  static constexpr int OFF_OUT = OFF_W + STAGES * W_STAGE;
  static constexpr int OFF_PARAM = OFF_OUT + TP_EPI_WARPS * OUT_WARP;
  static constexpr int OFF_BAR = OFF_PARAM + ((PARAM_FLOATS * 4 + 15) / 16) * 16;
  static constexpr int NUM_BARS = (W4 ? 3 : 2) * STAGES + 4;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int OFF_MM = OFF_TMEM + 16;               // GEGLU: per-warp min / max
  static constexpr int TOTAL = OFF_MM + TP_EPI_WARPS * 8;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static_assert(DYN_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

// ---- CTA-pair (cta_group::2) forms ------------------------------------------------------------
// TMA loads whose completion bytes go to an mbarrier of EITHER CTA of the pair (`bar_cluster_addr`
// is a shared::cluster address, e.g. the leader's full barrier obtained with mapa)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m,
                                                 uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* m,
                                                 uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m,
                                                 uint32_t bar_cluster_addr, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
// one 256 x N x 32 MMA across the pair: issued by ONE thread of the leader CTA; descriptors are
// shared-memory offsets valid in both CTAs, D is the same TMEM address in both
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit of the pair's MMAs arriving on the mbarrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// plain arrive on an mbarrier given by its shared::cluster address (own or peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// explicit shared-memory accesses (generic ld/st through a char* pays the generic-address path)
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w) : "memory");
}
// TMA store of one box from shared memory (tile mode, clipped at the tensor bounds) as a bulk
// async-group; wait_group.read<N>: at most N groups may still be READING shared memory
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// profiling stamps (p.dbg != nullptr): 16 slots of %globaltimer per CTA; tile dbg_tl = dbg_mode >> 4
// of every CTA (and the one after it)
#define TP_DBG(cond, slot)                                                        \
  do {                                                                            \
    if (p.dbg != nullptr && (cond)) p.dbg[blockIdx.x * 16 + (slot)] = gtime_ns(); \
  } while (0)

template <int BN, int STAGES, int KIND, bool W4, int CS, bool HALO = false>
__global__ void __launch_bounds__(W4 ? TP_THREADS_W4 : TP_THREADS, 1)
tc_i8_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmD, const TcParams p) {
  using L = TpSmem<BN, STAGES, KIND, W4, CS, HALO>;
  static_assert(!HALO || (KIND == KIND_CONV && !W4), "HALO: int8 3x3 convolutions only");
  static_assert(KIND == KIND_GEMM || KIND == KIND_CONV || KIND == KIND_GEGLU, "unsupported kind");
  static_assert(BN % 32 == 0 && BN / L::CH >= TP_EPI_PARTS && BN <= TP_SLOT_COLS,
                "tile width (every epilogue warp owns >= 1 chunk)");
  static_assert(CS == 1 || CS == 2, "cluster size along M");
  static_assert(CS == 1 || BN % 32 == 0, "W halves");
  constexpr bool PAIR = CS == 2;
  constexpr uint32_t IDESC = umma_idesc_i8(PAIR ? 2 * BLOCK_M : BLOCK_M, BN);
  constexpr int CH = L::CH;
  constexpr int NCH = BN / CH;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem + L::OFF_A;
  uint8_t* sW = smem + L::OFF_W;
  float* s_scale = reinterpret_cast<float*>(smem + L::OFF_PARAM);
  float* s_bias0 = s_scale + BN;
  float* s_bias = s_bias0 + BN;
  float* s_table = s_bias + BN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint64_t* raw_full = tmem_empty + 2;           // [STAGES], W4 only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t dbg_tl = static_cast<uint32_t>(p.dbg_mode) >> 4;   // which tile of a CTA is stamped
  TP_DBG(threadIdx.x == 0, 11);                                      // kernel entry
  pdl_launch_dependents();

  // ---- tile schedule: groups of CS vertically adjacent tiles, m fastest; every cluster takes a
  //      CONTIGUOUS range of groups, so that consecutive tiles of a CTA share their n-tile and the
  //      per-column epilogue operands (two named barriers + dependent L2 loads, ~1.5 us that no
  //      pipeline hides) are reloaded once or twice per CTA instead of once per tile
  const int crank = (CS > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int cid = blockIdx.x / CS;
  const int nclusters = gridDim.x / CS;
  int g_begin, g_end;
  const int m_groups = (p.tiles_m + CS - 1) / CS;
  const int total_groups = m_groups * p.tiles_n;
  {
    const int base = total_groups / nclusters, rem = total_groups - base * nclusters;
    g_begin = cid * base + (cid < rem ? cid : rem);
    g_end = g_begin + base + (cid < rem ? 1 : 0);
  }
  // ring iterations per tile: k-blocks, or (HALO) the 3 x kb_per_tap (s, channel block) pairs,
  // each covering the three vertical taps
  const int num_kb = HALO ? 3 * p.kb_per_tap : p.num_kb;
  const bool leader = crank == 0;
  // the leader's barriers as shared::cluster addresses (identity for the leader itself)
  auto leader_addr = [&](const void* bar) {
    return PAIR ? dsmem_map(smem_u32(bar), 0u) : smem_u32(bar);
  };

  struct Tile { int m0, tn0, tp0, tq0, n_tile0; };
  auto tile_of = [&](int g) {
    Tile t;
    const int nt = g / m_groups;
    const int mt = (g - nt * m_groups) * CS + crank;     // may be >= tiles_m (odd tile count): an
    t.n_tile0 = nt * BN;                                 // all-out-of-bounds tile, nothing stored
    t.m0 = mt * BLOCK_M; t.tn0 = t.tp0 = t.tq0 = 0;
    if (KIND == KIND_CONV) {
      int x = mt;
      const int tq = x % p.tilesQ; x /= p.tilesQ;
      const int tp = x % p.tilesP; x /= p.tilesP;
      t.tq0 = tq * p.boxW; t.tp0 = tp * p.boxH; t.tn0 = x * p.boxN;
    }
    return t;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if (p.d_tma) tma_prefetch_desc(&tmD);
    for (int i = 0; i < STAGES; ++i) {
      // full: the producer's expect_tx arrival (+ the converter warps of BOTH CTAs for W4); in
      // pair mode only the leader's full barriers are used
      mbar_init(&full_bar[i], W4 ? 1 + CS * TP_CONV_WARPS : 1);
      mbar_init(&empty_bar[i], 1);                 // one tcgen05.commit (multicast to both CTAs)
      if (W4) mbar_init(&raw_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], CS * TP_EPI_WARPS);   // pair: both CTAs' epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(tmem_slot, 2 * TP_SLOT_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, 2 * TP_SLOT_COLS); tmem_relinquish(); }
  }
  __syncwarp();                          // lane 0 of warp 0 rejoins before the CTA barrier
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();          // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t a_bytes = (KIND == KIND_CONV) ? p.a_tx_bytes : static_cast<uint32_t>(L::A_BYTES);
    constexpr int WK = W4 ? BLOCK_K / 2 : BLOCK_K;       // bytes of K per k-block in memory
    constexpr int WROWS = L::W_ROWS;                     // W rows this CTA stages
    constexpr uint32_t W_TX = W4 ? L::W_BYTES / 2 : L::W_BYTES;   // W bytes landing per CTA and k-block
    auto load_w = [&](const Tile& t, int kb, int stage) {
      // W4: packed rows (64 B) land in the upper half of the slot, unswizzled, and complete on
      // this CTA's own raw_full barrier (the converter warps signal the leader's full barrier)
      uint8_t* w_dst = sW + stage * L::W_STAGE + (W4 ? L::W_BYTES / 2 : 0);
      const int row0 = t.n_tile0 + crank * WROWS;
      if (HALO) {
        // iteration kb = (s, channel block): the W tiles of taps (0, s), (1, s), (2, s)
        const int s = kb / p.kb_per_tap;
        const int c0 = (kb - s * p.kb_per_tap) * WK;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          if (PAIR) tma_load_3d_pair(w_dst + r * L::W_BYTES, &tmW, leader_addr(&full_bar[stage]), c0, r * 3 + s, row0);
          else tma_load_3d(w_dst + r * L::W_BYTES, &tmW, &full_bar[stage], c0, r * 3 + s, row0);
        }
      } else if (KIND == KIND_CONV) {
        const int tap = kb / p.kb_per_tap;
        const int c0 = (kb - tap * p.kb_per_tap) * WK;
        if (W4 || !PAIR) tma_load_3d(w_dst, &tmW, W4 ? &raw_full[stage] : &full_bar[stage], c0, tap, row0);
        else tma_load_3d_pair(w_dst, &tmW, leader_addr(&full_bar[stage]), c0, tap, row0);
      } else {
        if (W4 || !PAIR) tma_load_2d(w_dst, &tmW, W4 ? &raw_full[stage] : &full_bar[stage], kb * WK, row0);
        else tma_load_2d_pair(w_dst, &tmW, leader_addr(&full_bar[stage]), kb * WK, row0);
      }
    };
    auto load_a = [&](const Tile& t, int kb, int stage) {
      uint8_t* a_dst = sA + stage * L::A_STAGE;
      if (HALO) {
        // (boxH + 2) x boxW box starting one image row above the tile, shifted by s - 1 pixels
        const int s = kb / p.kb_per_tap;
        const int c0 = (kb - s * p.kb_per_tap) * BLOCK_K;
        if (PAIR) tma_load_4d_pair(a_dst, &tmA, leader_addr(&full_bar[stage]), c0, t.tq0 - 1 + s, t.tp0 - 1, t.tn0);
        else tma_load_4d(a_dst, &tmA, &full_bar[stage], c0, t.tq0 - 1 + s, t.tp0 - 1, t.tn0);
      } else if (KIND == KIND_CONV) {
        const int tap = kb / p.kb_per_tap;
        const int c0 = (kb - tap * p.kb_per_tap) * BLOCK_K;
        const int r = tap / p.S, s = tap - r * p.S;
        if (PAIR) tma_load_4d_pair(a_dst, &tmA, leader_addr(&full_bar[stage]), c0,
                                   t.tq0 * p.stride - p.pad + s, t.tp0 * p.stride - p.pad + r, t.tn0);
        else tma_load_4d(a_dst, &tmA, &full_bar[stage], c0, t.tq0 * p.stride - p.pad + s,
                         t.tp0 * p.stride - p.pad + r, t.tn0);
      } else {
        if (PAIR) tma_load_2d_pair(a_dst, &tmA, leader_addr(&full_bar[stage]), kb * BLOCK_K, t.m0);
        else tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BLOCK_K, t.m0);
      }
    };
    // bytes the LEADER's full barrier of a stage expects: A (+ unpacked W) of every CTA of the pair
    const uint32_t full_tx = static_cast<uint32_t>(CS) * (a_bytes + (W4 ? 0u : (HALO ? 3u : 1u) * W_TX));
    // The first ring-full of WEIGHT k-blocks does not depend on the preceding kernel: issue it
    // before the programmatic-dependency wait (activations follow after it).
    uint32_t it = 0;                                     // k-block iterations issued so far
    int pre = 0;
    // profiling only (results are garbage): bit1 = no TMA loads, bit0 = no MMA issue
    const bool skip_tma = !W4 && (p.dbg_mode & 2) != 0;   // (bits 4.. select the stamped tile)
    if (g_begin < g_end && !skip_tma) {
      const Tile t0 = tile_of(g_begin);
      pre = num_kb < STAGES ? num_kb : STAGES;
      if (elect_one()) {
        for (int i = 0; i < pre; ++i) {
          if (W4) mbar_expect_tx(&raw_full[i], W_TX);
          else if (leader) mbar_expect_tx(&full_bar[i], full_tx);
          load_w(t0, i, i);
        }
      }
      __syncwarp();
    }
    TP_DBG(lane == 0, 12);                               // setup done, first weights in flight
    pdl_wait();
    TP_DBG(lane == 0, 13);                               // dependency wait passed
    for (int g = g_begin; g < g_end; ++g) {
      const Tile t = tile_of(g);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int stage = it % STAGES;
        const bool early_w = (g == g_begin) && (kb < pre);   // W already in flight
        if (it >= static_cast<uint32_t>(STAGES)) mbar_wait(&empty_bar[stage], ((it / STAGES) & 1u) ^ 1u);
        if (elect_one()) {
          if (skip_tma) {
            if (leader) mbar_arrive(&full_bar[stage]);
          } else {
            if (early_w) {
              if (W4 && leader) mbar_expect_tx(&full_bar[stage], full_tx);
            } else {
              if (W4) mbar_expect_tx(&raw_full[stage], W_TX);
              if (leader) mbar_expect_tx(&full_bar[stage], full_tx);
              load_w(t, kb, stage);
            }
            load_a(t, kb, stage);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: the leader CTA's only) =====================
    if (leader) {
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
    const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW));
    uint32_t it = 0, tl = 0;
    for (int g = g_begin; g < g_end; ++g, ++tl) {
      const uint32_t slot = tl & 1u;
      if (tl >= 2) mbar_wait(&tmem_empty[slot], ((tl >> 1) & 1u) ^ 1u);   // epilogue drained it
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + slot * TP_SLOT_COLS;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int stage = it % STAGES;
        TP_DBG(lane == 0 && kb == 0 && (tl == dbg_tl || tl == dbg_tl + 1), tl == dbg_tl ? 7 : 9);   // tile's first wait
        mbar_wait(&full_bar[stage], (it / STAGES) & 1u);
        TP_DBG(lane == 0 && kb == 0 && tl == dbg_tl, 10);                             // first stage landed
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = a_desc0 + static_cast<uint64_t>(stage * (L::A_STAGE >> 4));
          const uint64_t w_desc = w_desc0 + static_cast<uint64_t>(stage * (L::W_STAGE >> 4));
          if (!(p.dbg_mode & 1)) {
            if (HALO) {
              // tap r reads the box from image row r on: + r * boxW rows of 128 B (1024 B aligned)
              const uint64_t a_tap = static_cast<uint64_t>((p.boxW * BLOCK_K) >> 4);
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                  const uint64_t ad = a_desc + r * a_tap + static_cast<uint64_t>(k * (UMMA_K >> 4));
                  const uint64_t wd = w_desc + static_cast<uint64_t>(r * (L::W_BYTES >> 4) + k * (UMMA_K >> 4));
                  if (PAIR) umma_i8_pair(d_tmem, ad, wd, IDESC, (kb | r | k) ? 1u : 0u);
                  else umma_i8(d_tmem, ad, wd, IDESC, (kb | r | k) ? 1u : 0u);
                }
            } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              if (PAIR) umma_i8_pair(d_tmem, a_desc + static_cast<uint64_t>(k * (UMMA_K >> 4)),
                                     w_desc + static_cast<uint64_t>(k * (UMMA_K >> 4)), IDESC,
                                     (kb | k) ? 1u : 0u);
              else umma_i8(d_tmem, a_desc + static_cast<uint64_t>(k * (UMMA_K >> 4)),
                           w_desc + static_cast<uint64_t>(k * (UMMA_K >> 4)), IDESC, (kb | k) ? 1u : 0u);
            }
          }
          if (PAIR) {
            umma_commit_pair(&empty_bar[stage], 0x3);                // frees the slot in BOTH CTAs
            if (kb == num_kb - 1) umma_commit_pair(&tmem_full[slot], 0x3);
          } else {
            umma_commit(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit(&tmem_full[slot]);    // accumulator complete
          }
        }
        __syncwarp();
        TP_DBG(lane == 0 && kb == num_kb - 1 && tl == dbg_tl, 8);                     // last MMA issued
      }
    }
    }
  } else if (warp < 2 + TP_EPI_WARPS) {
    // ===================== epilogue warps =====================
    // Each warp drains 32 tile rows (lane == row) x its quarter of the columns in 16-output chunks:
    // tcgen05.ld -> dequant (+ the fused elementwise tails, applied in the same row-per-lane
    // layout: 32 B = one sector per row and tail) -> fp16 -> a 1 KB staging tile in the layout of a
    // SWIZZLE_32B TMA box -> ONE cp.async.bulk.tensor store per chunk issued by lane 0 (the TMA
    // unit clips at the tensor edges, so no per-row / per-column predicates and no address
    // arithmetic remain in the warps). Two staging tiles per warp: the store of chunk c drains
    // while chunk c + 1 is computed. p.d_tma == 0 (output rows of a tile not contiguous: conv
    // geometries with ragged tiles): staged copy-out with 16-byte stores instead.
    pdl_wait();                            // dynamic-quantisation scalars come from the predecessor
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;       // 0..511
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int epart = ew >> 2;             // which quarter of the tile's columns
    const bool has_bias = p.bias != nullptr;
    const uint32_t stage_u32 = smem_u32(smem + L::OFF_OUT + ew * L::OUT_WARP);   // 2 x 1 KB
    const uint32_t s_scale_u32 = smem_u32(s_scale), s_bias0_u32 = smem_u32(s_bias0),
                   s_bias_u32 = smem_u32(s_bias), s_table_u32 = smem_u32(s_table);
    const int row = quarter * 32 + lane;
    const int sw = (lane >> 2) & 1;        // SWIZZLE_32B: 16-byte half ^= bit 7 of the address
    const bool d_tma = p.d_tma != 0;
    int cur_n = -1;
    float mn = 0.f, mx = 0.f;              // GEGLU: running min / max of this CTA's outputs
    const bool to_q = (KIND == KIND_GEGLU) && p.q_out != nullptr;   // GEGLU -> int8, static scales
    const float q_inv = to_q ? __ldg(p.q_inv) : 0.f, q_zp = to_q ? __ldg(p.q_zp) : 0.f;
    uint32_t tl = 0, nstore = 0;           // nstore: chunks staged so far (staging tile = parity)
    for (int g = g_begin; g < g_end; ++g, ++tl) {
      const Tile t = tile_of(g);
      const uint32_t slot = tl & 1u;
      // ---- per-column operands of this n-tile (reloaded only when the n-tile changes) ----
      if (t.n_tile0 != cur_n) {
        named_bar_sync(1, 32 * TP_EPI_WARPS);          // everyone is done with the old values
        for (int j = et; j < BN; j += 32 * TP_EPI_WARPS) {
          const int n = t.n_tile0 + j;
          const bool ok = n < p.N;
          float sc = 0.f, b0 = 0.f, bs = 0.f;
          if (ok) {
            if (p.a_scale != nullptr) {
              sc = __fmul_rn(__ldg(p.scale + n), __ldcg(p.a_scale));
              b0 = __fmul_rn(__ldg(p.bias0 + n), __ldcg(p.a_zp));
            } else {
              sc = __ldg(p.scale + n);
              if (!(KIND == KIND_CONV) || !p.has_table) b0 = __ldg(p.bias0 + n);
            }
            if (has_bias) bs = __half2float(p.bias[n]);
          }
          s_scale[j] = sc; s_bias0[j] = b0; s_bias[j] = bs;
          if (KIND == KIND_CONV && p.has_table) {
            float w9[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) w9[q] = ok ? __ldg(p.bias0 + static_cast<int64_t>(n) * 9 + q) : 0.f;
            const float zp = __ldcg(p.a_zp);
#pragma unroll
            for (int rc = 0; rc < 4; ++rc)
#pragma unroll
              for (int sc4 = 0; sc4 < 4; ++sc4) {
                float acc = 0.f;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                  for (int s = 0; s < 3; ++s) {
                    const bool rv = !((rc & 1) && r == 0) && !((rc & 2) && r == 2);
                    const bool sv = !((sc4 & 1) && s == 0) && !((sc4 & 2) && s == 2);
                    if (rv && sv) acc = __fadd_rn(acc, w9[r * 3 + s]);
                  }
                s_table[(rc * 4 + sc4) * L::TAB_PITCH + j] = __fmul_rn(acc, zp);
              }
          }
        }
        named_bar_sync(1, 32 * TP_EPI_WARPS);
        cur_n = t.n_tile0;
      }
      const RowInfo ri = row_info<KIND>(p, row, t.m0, t.tn0, t.tp0, t.tq0);
      // first output row of this warp's 32-row slab (TMA stores: rows of a tile are contiguous)
      const RowInfo r0 = row_info<KIND>(p, quarter * 32, t.m0, t.tn0, t.tp0, t.tq0);
      const int c_lo = (NCH * epart) / TP_EPI_PARTS;
      const int c_hi = (NCH * (epart + 1)) / TP_EPI_PARTS;
      const uint32_t t_base = tmem_base + slot * TP_SLOT_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
      // per-row operand pointers of the fused tails
      const bool has_ca = (KIND != KIND_GEGLU) && p.chan_add != nullptr;
      const bool has_rs = (KIND != KIND_GEGLU) && p.residual != nullptr;
      const __half* ca_row = has_ca && ri.ok ? p.chan_add + (ri.out_row / p.rows_per_img) * p.ldca : nullptr;
      const __half* rs_row = has_rs && ri.ok ? p.residual + ri.out_row * p.ldr : nullptr;
      const uint32_t b0_u32 = (KIND == KIND_CONV && p.has_table)
                                  ? s_table_u32 + static_cast<uint32_t>(ri.cls * L::TAB_PITCH) * 4u
                                  : s_bias0_u32;

      // 8 accumulators -> 8 halves: three separately rounded fp32 operations, then RN to fp16
      auto dequant8 = [&](int col, const int32_t* a, __half* h) {
#pragma unroll
        for (int j0 = 0; j0 < 8; j0 += 4) {
          const uint32_t off = static_cast<uint32_t>(col + j0) * 4u;
          const float4 sc = lds_f4(s_scale_u32 + off);
          const float4 b0 = lds_f4(b0_u32 + off);
          float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_bias) bs = lds_f4(s_bias_u32 + off);
          const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
          const float b0v[4] = {b0.x, b0.y, b0.z, b0.w};
          const float bsv[4] = {bs.x, bs.y, bs.z, bs.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f = dequant_f32(a[j0 + j], b0v[j], scv[j]);
            if (has_bias) f = __fadd_rn(f, bsv[j]);
            h[j0 + j] = __float2half_rn(f);
          }
        }
      };
      // stage 16 halves of this lane's row and send the 32 x 16 tile out
      auto store_chunk = [&](const __half* y, int out_col) {
        const uint32_t buf = stage_u32 + (L::OUT_BUFS == 2 ? (nstore & 1u) * 1024u : 0u);
        if (d_tma) {
          // the store that last read this staging tile (two chunks ago; HALO: the previous one)
          // must be done reading
          if (lane == 0) bulk_wait_read<L::OUT_BUFS - 1>();
          __syncwarp();
          sts_v4(buf + lane * 32 + (sw << 4), reinterpret_cast<const uint4*>(y)[0]);
          sts_v4(buf + lane * 32 + ((sw ^ 1) << 4), reinterpret_cast<const uint4*>(y)[1]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmD, buf, out_col, static_cast<int>(r0.out_row));
            bulk_commit();
          }
        } else {
          __syncwarp();                                // previous copy-out has read this tile
          if (ri.ok) {
            sts_v4(buf + lane * 32 + (sw << 4), reinterpret_cast<const uint4*>(y)[0]);
            sts_v4(buf + lane * 32 + ((sw ^ 1) << 4), reinterpret_cast<const uint4*>(y)[1]);
          }
          __syncwarp();
          // two lanes per row, 16 rows per instruction
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = i * 16 + (lane >> 1);
            const RowInfo ro = row_info<KIND>(p, quarter * 32 + r, t.m0, t.tn0, t.tp0, t.tq0);
            const int col = out_col + (lane & 1) * 8;
            if (ro.ok && col + 8 <= p.d_cols)
              *reinterpret_cast<uint4*>(p.D + ro.out_row * p.ldd + col) =
                  lds_v4(buf + r * 32 + ((((lane & 1) ^ (r >> 2)) & 1) << 4));
          }
        }
        ++nstore;
      };

      TP_DBG(threadIdx.x == 64 && tl == dbg_tl, 0);       // epilogue of tile 2 starts waiting
      if (KIND == KIND_GEGLU) {
        // 32 accumulator columns = 16 value + 16 gate columns of the same 16 outputs
        mbar_wait(&tmem_full[slot], (tl >> 1) & 1u);
        TP_DBG(threadIdx.x == 64 && tl == dbg_tl, 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_base + c * 32, reinterpret_cast<uint32_t(&)[32]>(v));
          tmem_ld_wait();
          if (c == c_hi - 1) {                         // this warp has read its part of the slot
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_addr(&tmem_empty[slot]));
          }
          if (W4) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = static_cast<uint32_t>(static_cast<int32_t>(v[j]) >> 4);
          }
          __align__(16) __half y[16];
          {
            __align__(16) __half h[32];
#pragma unroll
            for (int j8 = 0; j8 < 32; j8 += 8)
              dequant8(c * 32 + j8, reinterpret_cast<const int32_t*>(v) + j8, h + j8);
            const bool live = ri.ok && (t.n_tile0 + c * 32 + 32 <= p.N);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              y[j] = geglu_half(h[j], h[16 + j]);
              const float f = live ? __half2float(y[j]) : 0.f;
              mn = fminf(mn, f);
              mx = fmaxf(mx, f);
            }
          }
          if (to_q) {
            // static scales of the consumer: 16 codes of this lane's row straight to global memory
            if (ri.ok && (t.n_tile0 + c * 32 + 32 <= p.N)) {
              const uint2 lo = static_quant8(reinterpret_cast<const int4*>(y)[0], q_inv, q_zp);
              const uint2 hi = static_quant8(reinterpret_cast<const int4*>(y)[1], q_inv, q_zp);
              *reinterpret_cast<uint4*>(p.q_out + ri.out_row * p.ldq + ((t.n_tile0 + c * 32) >> 1)) =
                  make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
          } else {
            store_chunk(y, (t.n_tile0 + c * 32) >> 1);
          }
        }
      } else {
        // operands of the fused tails, fetched one chunk ahead (the first chunk's while the MMAs
        // of this tile are still running)
        uint4 t_ca[2], t_rs[2];
        auto fetch_tail = [&](int c) {
          const int col = t.n_tile0 + c * CH;
          if (c < c_hi) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (col + i * 8 + 8 > p.N) continue;
              if (ca_row != nullptr) t_ca[i] = __ldcg(reinterpret_cast<const uint4*>(ca_row + col + i * 8));
              if (rs_row != nullptr) t_rs[i] = __ldcg(reinterpret_cast<const uint4*>(rs_row + col + i * 8));
            }
          }
        };
        if (has_ca || has_rs) fetch_tail(c_lo);
        mbar_wait(&tmem_full[slot], (tl >> 1) & 1u);
        TP_DBG(threadIdx.x == 64 && tl == dbg_tl, 1);       // accumulator ready
        tc_fence_after();
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          uint32_t v[16];
          tmem_ld_32x16(t_base + c * CH, v);
          tmem_ld_wait();
          TP_DBG(threadIdx.x == 64 && tl == dbg_tl && c == c_lo, 2);   // first chunk in registers
          if (c == c_hi - 1) {                         // this warp has read its part of the slot
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_addr(&tmem_empty[slot]));
          }
          if (W4) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = static_cast<uint32_t>(static_cast<int32_t>(v[j]) >> 4);
          }
          if (p.acc_out != nullptr && ri.ok) {         // raw accumulators (parity tests)
#pragma unroll
            for (int j4 = 0; j4 < CH; j4 += 4) {
              const int col = t.n_tile0 + c * CH + j4;
              if (col + 4 <= p.N)
                *reinterpret_cast<int4*>(p.acc_out + ri.out_row * p.N + col) =
                    make_int4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
            }
          }
          __align__(16) __half y[16];
#pragma unroll
          for (int j8 = 0; j8 < CH; j8 += 8)
            dequant8(c * CH + j8, reinterpret_cast<const int32_t*>(v) + j8, y + j8);
          if (has_ca || has_rs) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (t.n_tile0 + c * CH + i * 8 + 8 > p.N) continue;
              uint4 o = reinterpret_cast<const uint4*>(y)[i];
              if (ca_row != nullptr) o = add_half8(o, t_ca[i]);
              if (rs_row != nullptr) o = add_half8(o, t_rs[i]);
              reinterpret_cast<uint4*>(y)[i] = o;
            }
            fetch_tail(c + 1);
          }
          TP_DBG(threadIdx.x == 64 && tl == dbg_tl && c == c_lo, 3);   // first chunk dequantised
          store_chunk(y, t.n_tile0 + c * CH);
          TP_DBG(threadIdx.x == 64 && tl == dbg_tl && c == c_lo, 4);   // first chunk staged / sent
        }
      }
      TP_DBG(threadIdx.x == 64 && (tl == dbg_tl || tl == dbg_tl + 1), tl == dbg_tl ? 5 : 6);   // tile drained
    }
    if (d_tma && lane == 0) bulk_wait_all();           // every store has completed
    __syncwarp();
    tc_fence_before();
    if (KIND == KIND_GEGLU) {
      // one min / max partial per CTA (consumed by quant_rows_premm_kernel, quant2.cu)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      float* s_mm = reinterpret_cast<float*>(smem + L::OFF_MM);
      if (lane == 0) { s_mm[ew * 2] = mn; s_mm[ew * 2 + 1] = mx; }
      named_bar_sync(1, 32 * TP_EPI_WARPS);
      if (threadIdx.x == 64 && p.q_out == nullptr) {
#pragma unroll
        for (int w = 0; w < TP_EPI_WARPS; ++w) { mn = fminf(mn, s_mm[2 * w]); mx = fmaxf(mx, s_mm[2 * w + 1]); }
        p.mm_partial[blockIdx.x] = make_float2(mn, mx);
      }
    }
  } else {
    // ===================== W4 converter warps (10..13) =====================
    if constexpr (W4) {
      const int ct = threadIdx.x - 32 * (2 + TP_EPI_WARPS);       // 0..127
      constexpr int NT = 32 * TP_CONV_WARPS;
      constexpr int CPS = L::W_ROWS * 4;                           // 16-byte packed chunks per k-block
      constexpr int PT = (CPS + NT - 1) / NT;
      uint32_t it = 0;
      for (int g = g_begin; g < g_end; ++g) {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % STAGES;
          mbar_wait(&raw_full[stage], (it / STAGES) & 1u);
          uint8_t* wst = sW + stage * L::W_STAGE;
          uint4 pk[PT];
#pragma unroll
          for (int j = 0; j < PT; ++j) {
            const int c = ct + j * NT;
            if (c < CPS) pk[j] = *reinterpret_cast<const uint4*>(wst + L::W_BYTES / 2 + c * 16);
          }
          named_bar_sync(2, NT);               // every packed chunk is in registers: overwrite
#pragma unroll
          for (int j = 0; j < PT; ++j) {
            const int c = ct + j * NT;
            if (c < CPS) {
              const int row = c >> 2, cpos = c & 3;
              const uint32_t pw[4] = {pk[j].x, pk[j].y, pk[j].z, pk[j].w};
              uint32_t o[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t e = pw[q] & 0xF0F0F0F0u;
                const uint32_t d = (pw[q] << 4) & 0xF0F0F0F0u;
                o[2 * q] = __byte_perm(e, d, 0x5140);
                o[2 * q + 1] = __byte_perm(e, d, 0x7362);
              }
              uint8_t* rowp = wst + row * 128;
              const int sw = row & 7;
              *reinterpret_cast<uint4*>(rowp + (((2 * cpos) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(rowp + (((2 * cpos + 1) ^ sw) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_addr(&full_bar[stage]));
        }
      }
    }
  }

  __syncwarp();
  __syncthreads();
  TP_DBG(threadIdx.x == 0, 14);                          // all roles done
  if (PAIR) cluster_sync_all();        // no CTA exits (or frees TMEM) while the pair still runs
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 2 * TP_SLOT_COLS);
    else tmem_dealloc(tmem_base, 2 * TP_SLOT_COLS);
  }
}

}  // namespace mixdq
