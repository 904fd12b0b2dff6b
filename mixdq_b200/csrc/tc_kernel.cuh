// tc_kernel.cuh — K-G / K-C: INT8 x INT8 -> INT32 contraction on the 5th-gen tensor cores.
//
//   tcgen05.mma.cta_group::1.kind::i8, 128 x BN x 32 per instruction, s32 accumulators in TMEM,
//   A (activations) and W (weights) tiles staged by TMA into 128B-swizzled shared memory through
//   a STAGES-deep mbarrier ring, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM
//   allocator + single-thread MMA issuer, warps 2..9 = epilogue (tcgen05.ld -> dequant -> fp16;
//   two warps per TMEM lane quarter, each taking half of the tile's columns).
//
//   KIND_GEMM : A is a 2-D [M][K] matrix                       (reference qlinear, A2)
//   KIND_CONV : A is the NHWC activation tensor; each k-block is one (r,s) filter tap x 128
//               input channels fetched as a shifted 4-D TMA box with out-of-bounds zero fill —
//               implicit GEMM, no im2col buffer                 (reference qconv2d, A3)
//               the zero-point border correction of activation_zero_point_propagate (A4) is a
//               16-class x BN table built in shared memory by the epilogue warps
//   KIND_SPLIT: two K-phases (two A/W operand pairs) into two TMEM accumulators, combined in the
//               epilogue exactly like the reference's two fp16 convs + fp16 add (A6)
//   KIND_GEGLU: KIND_GEMM whose weight rows are interleaved 16 value / 16 gate columns
//               (ff.net.0.proj): the epilogue rounds the linear output to fp16, evaluates
//               half(h * half(gelu(gate))) like the stock GEGLU module, stores [M][N/2] fp16 and
//               one min / max partial per CTA, so the quantiser that follows is a single pass
//
// W4 (template flag): the weight operand is PACKED signed 4-bit codes (two per byte, even k in
//   the high nibble — the reference's only nibble convention, nn/utils.py:26-28). TMA moves the
//   packed tile (64 bytes per row and k-block: half the HBM / L2 bytes) into the upper half of
//   the stage's W slot; the eight epilogue warps — idle during the mainloop — expand it in place
//   into the 128B-swizzled int8 layout the MMA reads (5 ALU instructions per 8 codes: each
//   nibble is moved to the HIGH half of its byte, i.e. the int8 value 16*w, no sign extension
//   needed) and signal the stage's full barrier. The accumulators are therefore exactly 16x the
//   true ones and the epilogue shifts them right by 4 before the unchanged dequant arithmetic.
//
// Split-K over a thread-block cluster (p.splits > 1, cluster dims (1,1,splits)).
//   Measured on B200 (tools/phase_timing.py): with both operands in shared memory one
//   tcgen05.mma (M=128, K=32) occupies the tensor pipe ~160 cycles whatever N is — the A slab
//   (128 rows x 32 B) is streamed from shared memory at one row per cycle — so a CTA's mainloop
//   costs ~5 cycles per byte of K regardless of the tile width. Batch-1 UNet layers (M = 256)
//   have far fewer output tiles than SMs and long K, hence: wide tiles (BN up to 256) for work
//   per A-read, and the K range split across the CTAs of a cluster so that every SM streams a
//   K-slice of A and W. The partial INT32 accumulators are reduce-scattered through an
//   L2-resident global workspace (coalesced stores, one cluster barrier, coalesced .cg loads; a
//   first version exchanged them with st.shared::cluster and measured ~8 B/clk — 4.8 us per
//   128x128 tile): each CTA owns 128/splits rows of the tile, sums the partials exactly in
//   int32 and applies the dequant epilogue — no atomics, bit-exact.
//
// Epilogue: TMEM rows are owned by lanes (lane == tile row), so storing straight to global memory
//   is one 32-line scatter per instruction (measured 4.8 us for a 128x256 tile). Results are
//   therefore staged as fp16 in shared memory (aliased onto the drained stage ring) and copied
//   out with consecutive threads writing consecutive 16-byte chunks of a row.
//
// Programmatic dependent launch: the kernel is launched with programmatic stream serialization;
//   barrier/TMEM setup and the weight (W) loads of the first STAGES k-blocks are issued before
//   griddepcontrol.wait, so they overlap the tail of the preceding kernel (normally the
//   activation-quantize kernel, which triggers launch_dependents at its entry).
#pragma once
#include "common.cuh"

namespace mixdq {

enum { KIND_GEMM = 0, KIND_CONV = 1, KIND_SPLIT = 2, KIND_GEGLU = 3 };

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 128;  // int8 elements = bytes = one 128B swizzle row
constexpr int UMMA_K = 32;
constexpr int TC_THREADS = 352;      // warp 0 TMA, warp 1 MMA/TMEM, warps 2..9 epilogue,
                                     // warp 10 second MMA issuer (TcDual kernels only)
constexpr int EPI_THREADS = 256;

struct TcParams {
  // problem
  int M, N;             // GEMM rows (conv: informational), output channels
  int num_kb;           // k-blocks of phase 0
  int num_kb1;          // k-blocks of phase 1 (KIND_SPLIT only)
  int splits;           // split-K factor == cluster size along z (1 = no cluster reduction)
  // conv geometry (KIND_CONV)
  int kb_per_tap, S, pad, stride;
  int NB, H, W, P, Q;   // batch, input H/W, output P/Q
  int boxW, boxH, boxN; // pixels covered by one A box: boxN x boxH x boxW (<= 128 rows)
  int tilesQ, tilesP;   // tiles along q and p (tiles along n = gridDim.x / (tilesQ*tilesP))
  int tiles_m, tiles_n; // persistent kernel (tc_persist.cuh): tile grid walked by each CTA
  int d_tma;            // persistent kernel: 1 = outputs leave through TMA stores (tmD)
  int d_cols;           // persistent kernel: columns of the output matrix (N, or N / 2 for GEGLU)
  uint32_t a_tx_bytes;  // bytes one A box delivers
  int has_table;        // 1: pad > 0 -> border table from wsum_krs * zp ; 0: per-channel bias0
  // epilogue operands
  const float* scale;     // [N]   (static)  or w_scale[N] (dyn)
  const float* bias0;     // [N]   (static)  or wsum[N]   (dyn)   ; conv pad>0: wsum_krs [N][R*S]
  const float* a_scale;   // dyn: device scalar, else nullptr
  const float* a_zp;      // dyn / conv-table: device scalar zero point
  const __half* bias;     // [N] or nullptr
  const float* scale1;    // KIND_SPLIT second half
  const float* bias0_1;
  const float* a_scale1;  // KIND_SPLIT dyn: device scalars of the second half, else nullptr
  const float* a_zp1;
  // optional fused elementwise tail (each a separate fp16 op in the reference model, so each
  // rounds to fp16): D = half(D + chan_add[row / rows_per_img][col] (row pitch ldca)); D = half(D + residual[row][col])
  const __half* chan_add;
  int64_t ldca;
  int64_t rows_per_img;
  const __half* residual;
  int64_t ldr;
  __half* D;
  int64_t ldd;
  int32_t* acc_out;       // optional raw accumulator dump [rows][N]
  float2* mm_partial;     // KIND_GEGLU: per-CTA {min(0, min), max(0, max)} of the fp16 output
  // KIND_GEGLU with STATIC scales of the consumer (ff.net.2): the fp16 GEGLU values are quantised
  // in the epilogue, q_out[row][col] = sat8(rint(fma(y, *q_inv, *q_zp))), and D / mm_partial unused
  int8_t* q_out;
  int64_t ldq;
  const float* q_inv;
  const float* q_zp;
  int32_t* ws;            // split-K exchange workspace: [CTA][BN/4][128][4] int32 (splits > 1)
  unsigned long long* dbg; // optional phase timestamps (globaltimer ns), 8 slots per CTA
  int dbg_mode;            // profiling only: bit0 = skip MMA issue, bit1 = skip TMA loads
  int a_prefetch;          // bit 0: L2-prefetch the first A tile before the dependency wait;
                           // bit 1: L2-prefetch the weight k-blocks the ring cannot hold
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// every CTA stamps its own 16 slots: dbg[(linear CTA id) * 16 + slot] (buffer sized by the caller)
#define MIXDQ_DBG(slot)                                                                   \
  do {                                                                                    \
    if (p.dbg != nullptr)                                                                 \
      p.dbg[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = \
          gtime_ns();                                                                     \
  } while (0)

// k-blocks (128 bytes of K) per ring stage. The thread that issues tcgen05.mma is blocked for the
// duration of each MMA (measured: issue rate == completion rate), so the per-stage bookkeeping
// (mbarrier try_wait, fence, elect, commit: ~230 cycles) is NOT overlapped with the tensor pipe;
// two k-blocks = 8 MMAs per barrier round halve that overhead (442 -> ~325 cycles per k-block for
// BN <= 128). BN = 256 keeps one k-block per stage: its stages are 48 KB already.
template <int BN>
struct TcKsub { static constexpr int value = (BN <= 128) ? 2 : 1; };

// Two MMA-issuing warps, alternating ring stages, each with its OWN TMEM accumulator (summed
// exactly in the epilogue: int32 addition is associative). The issuing thread is blocked while its
// MMA executes, so one warp's barrier bookkeeping overlaps the other warp's MMAs and the tensor
// pipe stays busy. Narrow tiles only: the 160-wide GEGLU tile was tried (2 x 160 TMEM columns, 4
// stages instead of 5) and gained nothing — its mainloop streams 13 MB of weights and is bound by
// the loads, not by the issuer.
template <int BN, int KIND>
struct TcDual { static constexpr bool value = (BN <= 64) && (KIND != 2 /*KIND_SPLIT*/); };

template <int BN, int STAGES, int KIND, bool W4 = false>
struct TcSmem {
  static constexpr int KSUB = TcKsub<BN>::value;
  static constexpr int A_SUB = BLOCK_M * BLOCK_K;
  static constexpr int W_SUB = BN * BLOCK_K;
  static constexpr int A_BYTES = KSUB * A_SUB;
  static constexpr int W_BYTES = KSUB * W_SUB;
  static constexpr int TAB_PITCH = BN + 4;   // float4-aligned rows, classes land in different banks
  static constexpr int TAB_FLOATS = (KIND == KIND_CONV) ? 16 * TAB_PITCH : 0;
  static constexpr int PARAM_FLOATS = (KIND == KIND_SPLIT ? 5 : 3) * BN + TAB_FLOATS;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = OFF_A + STAGES * A_BYTES;
  static constexpr int OFF_PARAM = OFF_W + STAGES * W_BYTES;
  static constexpr int OFF_BAR = OFF_PARAM + ((PARAM_FLOATS * 4 + 15) / 16) * 16;
  static constexpr int NUM_BARS = (W4 ? 3 : 2) * STAGES + 1;   // W4: + packed-tile-landed barriers
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024 B alignment
  // fp16 output staging tile: 128 rows x (BN halves + 16 B pad), aliased onto the drained ring
  static constexpr int OUT_PITCH = BN * 2 + 16;
  static constexpr int OUT_BYTES = BLOCK_M * OUT_PITCH;
  static_assert(OUT_BYTES + 80 <= OFF_PARAM, "staging tile (+ min/max scratch) must fit in the stage ring");
};

// Per-row geometry of the output tile (which output row / border class a tile row maps to).
struct RowInfo {
  bool ok;
  int64_t out_row;
  int cls;
};

template <int KIND>
__device__ __forceinline__ RowInfo row_info(const TcParams& p, int row, int m0, int tn0, int tp0,
                                            int tq0) {
  RowInfo r;
  r.cls = 0;
  if (KIND == KIND_CONV) {
    const int per_img = p.boxH * p.boxW;
    const int dn = row / per_img;
    const int rem = row - dn * per_img;
    const int dh = rem / p.boxW, dw = rem - dh * p.boxW;
    const int n = tn0 + dn, pp = tp0 + dh, qq = tq0 + dw;
    r.ok = (dn < p.boxN) && (n < p.NB) && (pp < p.P) && (qq < p.Q);
    r.out_row = (static_cast<int64_t>(n) * p.P + pp) * p.Q + qq;
    if (p.has_table) {
      const int h0 = pp * p.stride - p.pad, w0 = qq * p.stride - p.pad;
      const int rc = (h0 < 0 ? 1 : 0) | (h0 + 2 >= p.H ? 2 : 0);
      const int sc4 = (w0 < 0 ? 1 : 0) | (w0 + 2 >= p.W ? 2 : 0);
      r.cls = rc * 4 + sc4;
    }
  } else {
    r.ok = (m0 + row) < p.M;
    r.out_row = m0 + row;
  }
  return r;
}

// 8 halves + 8 halves, each sum computed in fp32 and rounded to fp16 (what `a + b` on two fp16
// tensors does in PyTorch)
__device__ __forceinline__ uint4 add_half8(const uint4& a, const uint4& b) {
  uint4 r;
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __half22float2(pa[i]), fb = __half22float2(pb[i]);
    pr[i] = __floats2half2_rn(__fadd_rn(fa.x, fb.x), __fadd_rn(fa.y, fb.y));
  }
  return r;
}
// the fused elementwise tail on 8 consecutive output columns of one output row
__device__ __forceinline__ uint4 epilogue_tail(const TcParams& p, uint4 v, int64_t out_row, int col) {
  if (p.chan_add != nullptr) {
    const int64_t img = out_row / p.rows_per_img;
    v = add_half8(v, __ldcg(reinterpret_cast<const uint4*>(p.chan_add + img * p.ldca + col)));
  }
  if (p.residual != nullptr)
    v = add_half8(v, __ldcg(reinterpret_cast<const uint4*>(p.residual + out_row * p.ldr + col)));
  return v;
}

template <int BN, int STAGES, int KIND, bool W4 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
             const TcParams p) {
  using L = TcSmem<BN, STAGES, KIND, W4>;
  static_assert(!W4 || KIND != KIND_SPLIT, "W4 split shortcuts run as two convolutions");
  constexpr bool DUAL = TcDual<BN, KIND>::value;
  static_assert(!DUAL || (STAGES % 2 == 0), "dual issue alternates stages: even ring depth");
  constexpr int TMEM_COLS_USED = ((KIND == KIND_SPLIT || DUAL) ? 2 : 1) * BN;
  constexpr uint32_t TMEM_COLS = TMEM_COLS_USED <= 32 ? 32 : TMEM_COLS_USED <= 64 ? 64
                               : TMEM_COLS_USED <= 128 ? 128 : TMEM_COLS_USED <= 256 ? 256 : 512;
  constexpr uint32_t IDESC = umma_idesc_i8(BLOCK_M, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem + L::OFF_A;
  uint8_t* sW = smem + L::OFF_W;
  float* s_scale = reinterpret_cast<float*>(smem + L::OFF_PARAM);
  float* s_bias0 = s_scale + BN;
  float* s_bias = s_bias0 + BN;
  float* s_extra = s_bias + BN;  // conv: border table ; split: scale1, bias0_1
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* raw_full = tmem_full_bar + 1;     // W4 only: packed weight tile of the stage landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile0 = blockIdx.y * BN;
  if (threadIdx.x == 0) MIXDQ_DBG(0);   // kernel entry
  // let the next PDL-enabled kernel of the stream start its own prologue right away
  pdl_launch_dependents();

  // K range of this CTA (split-K rank == blockIdx.z == rank in the (1,1,splits) cluster)
  const int splits = (KIND == KIND_SPLIT || KIND == KIND_GEGLU) ? 1 : p.splits;
  const int krank = (splits > 1) ? static_cast<int>(blockIdx.z) : 0;
  const int total_kb = p.num_kb + (KIND == KIND_SPLIT ? p.num_kb1 : 0);
  const int kb_begin = (splits > 1) ? (krank * total_kb) / splits : 0;
  const int kb_end = (splits > 1) ? ((krank + 1) * total_kb) / splits : total_kb;
  // ring stage si holds k-blocks kb_begin + si*KSUB .. (+KSUB-1); the last one may be short
  const int nkb = kb_end - kb_begin;
  const int nst = (nkb + L::KSUB - 1) / L::KSUB;
  const int npre = nst < STAGES ? nst : STAGES;   // stages whose weights are issued before the wait
  auto nsub_of = [&](int si) { const int rem = nkb - si * L::KSUB; return rem < L::KSUB ? rem : L::KSUB; };

  // tile origin
  int m0 = 0, tn0 = 0, tp0 = 0, tq0 = 0;
  if (KIND == KIND_CONV) {
    int t = blockIdx.x;
    const int tq = t % p.tilesQ; t /= p.tilesQ;
    const int tp = t % p.tilesP; t /= p.tilesP;
    tq0 = tq * p.boxW; tp0 = tp * p.boxH; tn0 = t * p.boxN;
  } else {
    m0 = blockIdx.x * BLOCK_M;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if (KIND == KIND_SPLIT) { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmW1); }
    for (int i = 0; i < STAGES; ++i) {
      // W4: the stage is full once the A bytes landed AND the 8 converter warps expanded W
      mbar_init(&full_bar[i], W4 ? 1 + EPI_THREADS / 32 : 1);
      mbar_init(&empty_bar[i], 1);
      if (W4) mbar_init(&raw_full[i], 1);
    }
    mbar_init(tmem_full_bar, DUAL ? 2 : 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  __syncwarp();                          // lane 0 of warp 0 rejoins before the CTA barrier
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) MIXDQ_DBG(1);   // setup done (barriers, TMEM)

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop with warp-uniform control flow and ONE elected lane issues
    // (elect.sync): inside `if (lane == 0)` the compiler cannot prove uniformity and wraps every
    // uniform-datapath instruction in a waterfall loop.
    {
      const uint32_t a_bytes = (KIND == KIND_CONV) ? p.a_tx_bytes : static_cast<uint32_t>(L::A_SUB);
      // W4: the packed tile (BN rows x 64 bytes, unswizzled) lands in the UPPER half of the
      // sub-slot and completes on raw_full; the converter warps expand it in place
      constexpr int WK = W4 ? BLOCK_K / 2 : BLOCK_K;           // bytes of K per k-block in memory
      auto load_w = [&](int kb, int stage, int u) {
        uint8_t* w_dst = sW + stage * L::W_BYTES + u * L::W_SUB + (W4 ? L::W_SUB / 2 : 0);
        uint64_t* bar = W4 ? &raw_full[stage] : &full_bar[stage];
        if (KIND == KIND_CONV) {
          const int tap = kb / p.kb_per_tap;
          const int c0 = (kb - tap * p.kb_per_tap) * WK;
          tma_load_3d(w_dst, &tmW, bar, c0, tap, n_tile0);
        } else if (KIND == KIND_SPLIT && kb >= p.num_kb) {
          tma_load_2d(w_dst, &tmW1, bar, (kb - p.num_kb) * WK, n_tile0);
        } else {
          tma_load_2d(w_dst, &tmW, bar, kb * WK, n_tile0);
        }
      };
      auto load_a = [&](int kb, int stage, int u) {
        uint8_t* a_dst = sA + stage * L::A_BYTES + u * L::A_SUB;
        if (KIND == KIND_CONV) {
          const int tap = kb / p.kb_per_tap;
          const int c0 = (kb - tap * p.kb_per_tap) * BLOCK_K;
          const int r = tap / p.S, s = tap - r * p.S;
          tma_load_4d(a_dst, &tmA, &full_bar[stage], c0, tq0 * p.stride - p.pad + s,
                      tp0 * p.stride - p.pad + r, tn0);
        } else if (KIND == KIND_SPLIT && kb >= p.num_kb) {
          tma_load_2d(a_dst, &tmA1, &full_bar[stage], (kb - p.num_kb) * BLOCK_K, m0);
        } else {
          tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BLOCK_K, m0);
        }
      };
      const bool skip = W4 ? false : (p.dbg_mode & 2) != 0;
      // prologue: weights of the first ring-full of stages do not depend on the preceding
      // kernel -> issue them before the programmatic-dependency wait
      if (!skip && elect_one()) {
        for (int i = 0; i < npre; ++i) {
          const int ns = nsub_of(i);
          if (W4) mbar_expect_tx(&raw_full[i], ns * (L::W_SUB / 2));
          else mbar_expect_tx(&full_bar[i], ns * (a_bytes + L::W_SUB));
          for (int u = 0; u < ns; ++u) load_w(kb_begin + i * L::KSUB + u, i, u);
        }
        // WEIGHT k-blocks beyond the ring: pull them from HBM into L2 now, while the preceding
        // (short) kernel still runs — inside the UNet graph every weight byte is cold, and the
        // k-blocks the ring cannot hold would otherwise start their HBM trip only when a slot
        // frees up after the dependency wait (ff.net.2: 32 of its 40 k-blocks)
        if (p.a_prefetch & 2) {
          for (int si = npre; si < nst; ++si) {
            const int ns = nsub_of(si);
            for (int u = 0; u < ns; ++u) {
              const int kb = kb_begin + si * L::KSUB + u;
              if (KIND == KIND_CONV) {
                const int tap = kb / p.kb_per_tap;
                tma_prefetch_3d(&tmW, (kb - tap * p.kb_per_tap) * WK, tap, n_tile0);
              } else if (KIND == KIND_SPLIT && kb >= p.num_kb) {
                tma_prefetch_2d(&tmW1, (kb - p.num_kb) * WK, n_tile0);
              } else {
                tma_prefetch_2d(&tmW, kb * WK, n_tile0);
              }
            }
          }
        }
        // warm the path of the first A tile (L2 prefetch: its contents are not consumed)
        if (p.a_prefetch & 1) {
          const int kb = kb_begin;
          if (KIND == KIND_CONV) {
            const int tap = kb / p.kb_per_tap;
            const int c0 = (kb - tap * p.kb_per_tap) * BLOCK_K;
            const int r = tap / p.S, sx = tap - r * p.S;
            tma_prefetch_4d(&tmA, c0, tq0 * p.stride - p.pad + sx, tp0 * p.stride - p.pad + r, tn0);
          } else if (KIND == KIND_SPLIT && kb >= p.num_kb) {
            tma_prefetch_2d(&tmA1, (kb - p.num_kb) * BLOCK_K, m0);
          } else {
            tma_prefetch_2d(&tmA, kb * BLOCK_K, m0);
          }
        }
      }
      __syncwarp();
      if (lane == 0) MIXDQ_DBG(2);         // first TMA (weights) issued
      pdl_wait();                          // activations / quantisation scalars are ready
      if (elect_one()) {
        for (int i = 0; i < npre; ++i) {
          if (skip) { mbar_arrive(&full_bar[i]); continue; }
          const int ns = nsub_of(i);
          if (W4) mbar_expect_tx(&full_bar[i], ns * a_bytes);
          for (int u = 0; u < ns; ++u) load_a(kb_begin + i * L::KSUB + u, i, u);
        }
      }
      __syncwarp();
      int stage = (npre == STAGES) ? 0 : npre;
      uint32_t phase = (npre == STAGES) ? 1 : 0;
      for (int si = npre; si < nst; ++si) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          if (skip) {
            mbar_arrive(&full_bar[stage]);
          } else {
            const int ns = nsub_of(si);
            if (W4) {
              mbar_expect_tx(&raw_full[stage], ns * (L::W_SUB / 2));
              mbar_expect_tx(&full_bar[stage], ns * a_bytes);
            } else {
              mbar_expect_tx(&full_bar[stage], ns * (a_bytes + L::W_SUB));
            }
            for (int u = 0; u < ns; ++u) {
              load_w(kb_begin + si * L::KSUB + u, stage, u);
              load_a(kb_begin + si * L::KSUB + u, stage, u);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) MIXDQ_DBG(3);         // last TMA issued
    }
  } else if (warp == 1 || warp == 10) {
    // ===================== MMA issuer(s) (whole warp loops, one elected lane issues) =========
    // warp 1 alone, or (DUAL) warp 1 on even ring stages + warp 10 on odd ones, each into its own
    // accumulator
    const int which = (warp == 10) ? 1 : 0;
    if (which == 0 || DUAL) {
      // descriptor of stage 0 / k-slice 0; stages and k-slices advance the 16-byte-unit start
      // address field (the low word) by constants
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW));
      constexpr int STEP = DUAL ? 2 : 1;
      bool first = true;
      for (int si = which; si < nst; si += STEP) {
        const int stage = si % STAGES;
        const uint32_t phase = static_cast<uint32_t>(si / STAGES) & 1u;
        mbar_wait(&full_bar[stage], phase);
        if (si == 0 && lane == 0) MIXDQ_DBG(4);  // first stage landed
        tc_fence_after();
        if (elect_one()) {
          if (p.dbg_mode & 1) {
            mbar_arrive(&empty_bar[stage]);
          } else {
#pragma unroll
            for (int u = 0; u < L::KSUB; ++u) {
              const int kb = kb_begin + si * L::KSUB + u;
              if (kb < kb_end) {
                uint32_t d_tmem = tmem_base + (DUAL ? which * BN : 0);
                int kb_in_phase = kb - kb_begin;
                if (KIND == KIND_SPLIT && kb >= p.num_kb) { d_tmem += BN; kb_in_phase = kb - p.num_kb; }
                const uint64_t a_desc = a_desc0 + static_cast<uint64_t>(stage * (L::A_BYTES >> 4) + u * (L::A_SUB >> 4));
                const uint64_t w_desc = w_desc0 + static_cast<uint64_t>(stage * (L::W_BYTES >> 4) + u * (L::W_SUB >> 4));
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                  const bool overwrite = DUAL ? (first && u == 0 && k == 0) : ((kb_in_phase | k) == 0);
                  umma_i8(d_tmem, a_desc + static_cast<uint64_t>(k * (UMMA_K >> 4)),
                          w_desc + static_cast<uint64_t>(k * (UMMA_K >> 4)), IDESC,
                          overwrite ? 0u : 1u);
                }
              }
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          }
        }
        __syncwarp();
        first = false;
      }
      if (elect_one()) umma_commit(tmem_full_bar);   // this issuer's accumulator is complete
      __syncwarp();
      if (which == 0 && lane == 0) MIXDQ_DBG(5);     // last MMA issued
    }
  } else {
    // ============ epilogue warps 2..9: stage the per-column operands, wait for the MMAs ========
    const int et = threadIdx.x - 64;  // 0..255
    // W4: expand the packed weight tiles of ring stages [s0, s1) in place (see the header)
    auto convert_stages = [&](int s0, int s1) {
      constexpr int CPS = BN * 4;                        // 16-byte packed chunks per k-block
      constexpr int PT = (L::KSUB * CPS + EPI_THREADS - 1) / EPI_THREADS;
#pragma unroll 1
      for (int si = s0; si < s1; ++si) {
        const int stage = si % STAGES;
        mbar_wait(&raw_full[stage], static_cast<uint32_t>(si / STAGES) & 1u);
        const int nch = nsub_of(si) * CPS;
        uint8_t* wst = sW + stage * L::W_BYTES;
        uint4 pk[PT];
#pragma unroll
        for (int j = 0; j < PT; ++j) {
          const int i = et + j * EPI_THREADS;
          if (i < nch) {
            const int u = i / CPS, c = i - u * CPS;
            pk[j] = *reinterpret_cast<const uint4*>(wst + u * L::W_SUB + L::W_SUB / 2 + c * 16);
          }
        }
        epi_bar_sync();                    // every packed chunk is in registers: overwrite
#pragma unroll
        for (int j = 0; j < PT; ++j) {
          const int i = et + j * EPI_THREADS;
          if (i < nch) {
            const int u = i / CPS, c = i - u * CPS;
            const int row = c >> 2, cpos = c & 3;
            const uint32_t pw[4] = {pk[j].x, pk[j].y, pk[j].z, pk[j].w};
            uint32_t o[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t e = pw[q] & 0xF0F0F0F0u;          // even k: already in the high half
              const uint32_t d = (pw[q] << 4) & 0xF0F0F0F0u;   // odd k: low nibble moved up
              o[2 * q] = __byte_perm(e, d, 0x5140);            // k0 k1 k2 k3
              o[2 * q + 1] = __byte_perm(e, d, 0x7362);        // k4 k5 k6 k7
            }
            uint8_t* rowp = wst + u * L::W_SUB + row * 128;
            const int sw = row & 7;                            // 128B swizzle: chunk ^= row % 8
            *reinterpret_cast<uint4*>(rowp + (((2 * cpos) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(rowp + (((2 * cpos + 1) ^ sw) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
      }
    };
    if (W4) convert_stages(0, npre);       // weights do not depend on the predecessor
    pdl_wait();                            // dynamic-quantisation scalars come from the predecessor
    for (int j = et; j < BN; j += EPI_THREADS) {
      const int n = n_tile0 + j;
      const bool ok = n < p.N;
      float sc = 0.f, b0 = 0.f, bs = 0.f;
      if (ok) {
        if (p.a_scale != nullptr) {           // dynamic: fold the activation scalars here
          sc = __fmul_rn(__ldg(p.scale + n), __ldcg(p.a_scale));
          b0 = __fmul_rn(__ldg(p.bias0 + n), __ldcg(p.a_zp));
        } else {
          sc = __ldg(p.scale + n);
          if (!(KIND == KIND_CONV) || !p.has_table) b0 = __ldg(p.bias0 + n);
        }
        if (p.bias != nullptr) bs = __half2float(p.bias[n]);
      }
      s_scale[j] = sc; s_bias0[j] = b0; s_bias[j] = bs;
      if (KIND == KIND_SPLIT) {
        float s1 = ok ? __ldg(p.scale1 + n) : 0.f;
        float b1 = ok ? __ldg(p.bias0_1 + n) : 0.f;
        if (p.a_scale1 != nullptr) {
          s1 = __fmul_rn(s1, __ldcg(p.a_scale1));
          b1 = __fmul_rn(b1, __ldcg(p.a_zp1));
        }
        s_extra[j] = s1;
        s_extra[BN + j] = b1;
      }
      if (KIND == KIND_CONV && p.has_table) {
        // 3x3 / pad 1: class = rcls*4 + scls, bit0 = first tap cut, bit1 = last tap cut
        float w9[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w9[t] = ok ? __ldg(p.bias0 + static_cast<int64_t>(n) * 9 + t) : 0.f;
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
            s_extra[(rc * 4 + sc4) * L::TAB_PITCH + j] = __fmul_rn(acc, zp);
          }
      }
    }
    epi_bar_sync();                        // named barrier among the 256 epilogue threads
    if (W4) convert_stages(npre, nst);
  }
  // every epilogue path waits for the accumulators itself (after prefetching what it can)
  auto wait_accumulators = [&]() {
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) MIXDQ_DBG(6);  // accumulators ready
  };

  __syncwarp();
  const bool is_epi = warp >= 2 && warp < 10;
  // second accumulator in use? (DUAL kernels whose K range spans more than one ring stage)
  const bool dual_used = DUAL && nst > 1;
  const bool has_bias = p.bias != nullptr;
  const int quarter = warp & 3;          // TMEM lane quarter this warp may access
  const int ehalf = (warp - 2) >> 2;     // which half of the tile's columns this epilogue warp takes
  uint8_t* stage_out = smem;             // fp16 staging tile (ring is drained once tmem_full fired)

  // dequantise NV (4 or 8) consecutive columns of one row; returns packed halves
  auto dequant_pack = [&](int cls, int col, const int32_t* a0, const int32_t* a1, __half* h,
                          int nv) {
#pragma unroll
    for (int j0 = 0; j0 < 8; j0 += 4) {
      if (j0 >= nv) break;
      const float4 sc = *reinterpret_cast<const float4*>(s_scale + col + j0);
      const float* b0src = (KIND == KIND_CONV && p.has_table)
                               ? (s_extra + cls * L::TAB_PITCH + col + j0)
                               : (s_bias0 + col + j0);
      const float4 b0 = *reinterpret_cast<const float4*>(b0src);
      float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_bias) bs = *reinterpret_cast<const float4*>(s_bias + col + j0);
      const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
      const float b0v[4] = {b0.x, b0.y, b0.z, b0.w};
      const float bsv[4] = {bs.x, bs.y, bs.z, bs.w};
      float sc1v[4], b01v[4];
      if (KIND == KIND_SPLIT) {
        const float4 s1 = *reinterpret_cast<const float4*>(s_extra + col + j0);
        const float4 b1 = *reinterpret_cast<const float4*>(s_extra + BN + col + j0);
        sc1v[0] = s1.x; sc1v[1] = s1.y; sc1v[2] = s1.z; sc1v[3] = s1.w;
        b01v[0] = b1.x; b01v[1] = b1.y; b01v[2] = b1.z; b01v[3] = b1.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f = dequant_f32(a0[j0 + j], b0v[j], scv[j]);
        if (has_bias) f = __fadd_rn(f, bsv[j]);
        if (KIND == KIND_SPLIT) {
          // reference: two fp16 conv outputs added in fp16 (nn/Conv2d.py:346)
          // (torch adds halves in fp32 opmath and rounds once more to fp16)
          const float f1 = dequant_f32(a1[j0 + j], b01v[j], sc1v[j]);
          h[j0 + j] = __float2half_rn(__fadd_rn(__half2float(__float2half_rn(f)),
                                               __half2float(__float2half_rn(f1))));
        } else {
          h[j0 + j] = __float2half_rn(f);
        }
      }
    }
  };

  int rows_here = BLOCK_M;               // tile rows whose results this CTA stages and stores
  int row_base = 0;
  if (KIND == KIND_GEGLU) {
    if constexpr (BN >= 32) {
    if (is_epi) {
      // ---- 32 accumulator columns = 16 value + 16 gate columns of the same 16 outputs ----
      const int row = quarter * 32 + lane;
      const RowInfo ri = row_info<KIND>(p, row, m0, tn0, tp0, tq0);
      constexpr int NCH = BN / 32;                 // may be odd (BN = 160): 3 + 2 chunks
      const int c_lo = ehalf ? (NCH + 1) / 2 : 0;
      const int c_hi = ehalf ? NCH : (NCH + 1) / 2;
      RowInfo ro[2];                               // rows this lane copies out (2 lanes per row)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        ro[i] = row_info<KIND>(p, quarter * 32 + i * 16 + (lane >> 1), m0, tn0, tp0, tq0);
      float mn = 0.f, mx = 0.f;
      const bool to_q = p.q_out != nullptr;
      const float q_inv = to_q ? __ldg(p.q_inv) : 0.f, q_zp = to_q ? __ldg(p.q_zp) : 0.f;
      wait_accumulators();
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32;
        tmem_ld_32x32(taddr, reinterpret_cast<uint32_t(&)[32]>(v));
        if (DUAL && dual_used) {
          uint32_t v2[32];
          tmem_ld_32x32(taddr + BN, reinterpret_cast<uint32_t(&)[32]>(v2));
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        } else {
          tmem_ld_wait();
        }
        if (W4) {                      // codes were expanded as 16*w: exact arithmetic shift
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = static_cast<uint32_t>(static_cast<int32_t>(v[j]) >> 4);
        }
        const bool cols_ok = n_tile0 + c * 32 + 32 <= p.N;
        if (ri.ok && cols_ok) {
          __align__(16) __half h[32];
#pragma unroll
          for (int j8 = 0; j8 < 32; j8 += 8)
            dequant_pack(0, c * 32 + j8, reinterpret_cast<const int32_t*>(v) + j8,
                         reinterpret_cast<const int32_t*>(v) + j8, h + j8, 8);
          __align__(16) __half y[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            y[j] = geglu_half(h[j], h[16 + j]);
            const float f = __half2float(y[j]);
            mn = fminf(mn, f);
            mx = fmaxf(mx, f);
          }
          if (to_q) {
            // 16 codes = 16 bytes of this lane's row, straight to global memory
            const uint2 lo = static_quant8(reinterpret_cast<const int4*>(y)[0], q_inv, q_zp);
            const uint2 hi = static_quant8(reinterpret_cast<const int4*>(y)[1], q_inv, q_zp);
            *reinterpret_cast<uint4*>(p.q_out + ri.out_row * p.ldq + ((n_tile0 + c * 32) >> 1)) =
                make_uint4(lo.x, lo.y, hi.x, hi.y);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(stage_out + row * L::OUT_PITCH + c * 32);
            dst[0] = reinterpret_cast<const uint4*>(y)[0];
            dst[1] = reinterpret_cast<const uint4*>(y)[1];
          }
        }
        __syncwarp();
        if (cols_ok && !to_q) {
          const int64_t ocol = ((n_tile0 + c * 32) >> 1) + (lane & 1) * 8;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (!ro[i].ok) continue;
            const int r = quarter * 32 + i * 16 + (lane >> 1);
            *reinterpret_cast<uint4*>(p.D + ro[i].out_row * p.ldd + ocol) =
                *reinterpret_cast<const uint4*>(stage_out + r * L::OUT_PITCH + c * 32 + (lane & 1) * 16);
          }
        }
      }
      tc_fence_before();
      // min / max of this CTA's outputs: warp -> CTA (shared memory past the staging tile) -> one
      // partial per CTA, reduced by the quantise pass that follows (quant2.cu)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      }
      float* s_mm = reinterpret_cast<float*>(smem + ((L::OUT_BYTES + 15) & ~15));
      if (lane == 0) { s_mm[(warp - 2) * 2] = mn; s_mm[(warp - 2) * 2 + 1] = mx; }
      epi_bar_sync();
      if (threadIdx.x == 64 && !to_q) {
#pragma unroll
        for (int w = 0; w < 8; ++w) { mn = fminf(mn, s_mm[2 * w]); mx = fmaxf(mx, s_mm[2 * w + 1]); }
        p.mm_partial[blockIdx.y * gridDim.x + blockIdx.x] = make_float2(mn, mx);
      }
    }
    }
  } else if (splits == 1) {
    if (is_epi) {
      // ---- TMEM (lane == row) -> dequant -> fp16 -> staging tile -> global, chunk by chunk.
      //      Each warp owns 32 rows x its column half and copies a chunk out (4 lanes x 16 B per
      //      row, 8 rows per instruction) right after staging it, so the stores of chunk c drain
      //      while chunk c+1 is being dequantised. Only __syncwarp is needed.
      const int row = quarter * 32 + lane;
      const RowInfo ri = row_info<KIND>(p, row, m0, tn0, tp0, tq0);
      constexpr int CH = (BN >= 128) ? 32 : 16;   // narrow tiles: 16-column chunks keep all 8 epilogue warps busy
      constexpr int NCH = BN / CH;
      constexpr int LPR = CH / 8;                  // lanes (16-byte chunks) per row of a chunk
      constexpr int RPI = 32 / LPR;                // rows per copy-out instruction
      const int c_lo = (NCH >= 2) ? ehalf * (NCH / 2) : 0;
      const int c_hi = (NCH >= 2) ? c_lo + NCH / 2 : (ehalf == 0 ? 1 : 0);
      RowInfo ro[LPR];                             // rows this lane copies out
      int64_t ca_off[LPR];                         // chan_add row offset of those rows
#pragma unroll
      for (int i = 0; i < LPR; ++i) {
        ro[i] = row_info<KIND>(p, quarter * 32 + i * RPI + lane / LPR, m0, tn0, tp0, tq0);
        ca_off[i] = (p.chan_add != nullptr && ro[i].ok) ? (ro[i].out_row / p.rows_per_img) * p.ldca : 0;
      }
      // operands of the fused elementwise tail, fetched one chunk ahead: the first chunk's while
      // the MMAs are still running, chunk c+1's while chunk c is dequantised (they come from L2,
      // ~1 us away right after a kernel boundary)
      const bool has_tail = (p.chan_add != nullptr) || (p.residual != nullptr);
      uint4 t_ca[LPR], t_rs[LPR];
      auto fetch_tail = [&](int c) {
        const int ccol = n_tile0 + c * CH + (lane % LPR) * 8;
        if (c < c_hi && ccol + 8 <= p.N) {
#pragma unroll
          for (int i = 0; i < LPR; ++i) {
            if (!ro[i].ok) continue;
            if (p.chan_add != nullptr)
              t_ca[i] = __ldcg(reinterpret_cast<const uint4*>(p.chan_add + ca_off[i] + ccol));
            if (p.residual != nullptr)
              t_rs[i] = __ldcg(reinterpret_cast<const uint4*>(p.residual + ro[i].out_row * p.ldr + ccol));
          }
        }
      };
      if (has_tail) fetch_tail(c_lo);
      wait_accumulators();
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[CH], v1[CH];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * CH;
        const bool second = (KIND == KIND_SPLIT) || (DUAL && dual_used);
        if (CH == 32) {
          tmem_ld_32x32(taddr, reinterpret_cast<uint32_t(&)[32]>(v));
          if (second) tmem_ld_32x32(taddr + BN, reinterpret_cast<uint32_t(&)[32]>(v1));
        } else {
          tmem_ld_32x16(taddr, reinterpret_cast<uint32_t(&)[16]>(v));
          if (second) tmem_ld_32x16(taddr + BN, reinterpret_cast<uint32_t(&)[16]>(v1));
        }
        tmem_ld_wait();
        if (threadIdx.x == 64 && c == c_lo) MIXDQ_DBG(12);  // first accumulator chunk in registers
        if (DUAL && dual_used) {       // the two issuers' accumulators: exact int32 sum
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] += v1[j];
        }
        if (W4) {                      // codes were expanded as 16*w: exact arithmetic shift
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = static_cast<uint32_t>(static_cast<int32_t>(v[j]) >> 4);
        }
        if (ri.ok) {
#pragma unroll
          for (int j8 = 0; j8 < CH; j8 += 8) {
            const int col = c * CH + j8;
            if (p.acc_out != nullptr && n_tile0 + col + 8 <= p.N) {
              int32_t* arow = p.acc_out + ri.out_row * p.N + n_tile0 + col;
              *reinterpret_cast<int4*>(arow) = make_int4(v[j8], v[j8 + 1], v[j8 + 2], v[j8 + 3]);
              *reinterpret_cast<int4*>(arow + 4) = make_int4(v[j8 + 4], v[j8 + 5], v[j8 + 6], v[j8 + 7]);
            }
            __align__(16) __half h[8];
            dequant_pack(ri.cls, col, reinterpret_cast<const int32_t*>(v) + j8,
                         reinterpret_cast<const int32_t*>(v1) + j8, h, 8);
            *reinterpret_cast<uint4*>(stage_out + row * L::OUT_PITCH + col * 2) =
                *reinterpret_cast<const uint4*>(h);
          }
        }
        __syncwarp();
        if (threadIdx.x == 64 && c == c_lo) MIXDQ_DBG(13);  // first chunk dequantised and staged
        const int ccol = c * CH + (lane % LPR) * 8;
        uint4 o[LPR];
        const bool col_ok = n_tile0 + ccol + 8 <= p.N;
        if (col_ok) {
#pragma unroll
          for (int i = 0; i < LPR; ++i) {
            if (!ro[i].ok) continue;
            const int r = quarter * 32 + i * RPI + lane / LPR;
            o[i] = *reinterpret_cast<const uint4*>(stage_out + r * L::OUT_PITCH + ccol * 2);
            if (p.chan_add != nullptr) o[i] = add_half8(o[i], t_ca[i]);
            if (p.residual != nullptr) o[i] = add_half8(o[i], t_rs[i]);
          }
        }
        if (has_tail) fetch_tail(c + 1);           // next chunk's operands (registers are free now)
        if (col_ok) {
#pragma unroll
          for (int i = 0; i < LPR; ++i) {
            if (!ro[i].ok) continue;
            *reinterpret_cast<uint4*>(p.D + ro[i].out_row * p.ldd + n_tile0 + ccol) = o[i];
          }
        }
      }
      tc_fence_before();
      if (threadIdx.x == 64) MIXDQ_DBG(8);  // epilogue of this warp done
    }
  } else {
    // ===================== split-K: reduce-scatter through the L2-resident workspace =========
    // workspace slab of one CTA: [BN/4][128 rows][4] int32 -> lanes (rows) write adjacent 16 B
    const int64_t tile_lin = static_cast<int64_t>(blockIdx.y) * gridDim.x + blockIdx.x;
    int32_t* ws_tile = p.ws + tile_lin * splits * (BLOCK_M * BN);
    if (is_epi) {
      const int row = quarter * 32 + lane;
      int32_t* my = ws_tile + static_cast<int64_t>(krank) * (BLOCK_M * BN);
      constexpr int CH = (BN >= 128) ? 32 : 16;   // narrow tiles: 16-column chunks keep all 8 epilogue warps busy
      constexpr int NCH = BN / CH;
      const int c_lo = (NCH >= 2) ? ehalf * (NCH / 2) : 0;
      const int c_hi = (NCH >= 2) ? c_lo + NCH / 2 : (ehalf == 0 ? 1 : 0);
      wait_accumulators();
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[CH];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * CH;
        if (CH == 32) tmem_ld_32x32(taddr, reinterpret_cast<uint32_t(&)[32]>(v));
        else tmem_ld_32x16(taddr, reinterpret_cast<uint32_t(&)[16]>(v));
        if (DUAL && dual_used) {
          uint32_t v2[CH];
          if (CH == 32) tmem_ld_32x32(taddr + BN, reinterpret_cast<uint32_t(&)[32]>(v2));
          else tmem_ld_32x16(taddr + BN, reinterpret_cast<uint32_t(&)[16]>(v2));
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] += v2[j];
        } else {
          tmem_ld_wait();
        }
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          const int c4 = (c * CH + j) >> 2;
          *reinterpret_cast<int4*>(my + (static_cast<int64_t>(c4) * BLOCK_M + row) * 4) =
              make_int4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      tc_fence_before();
    }
    __syncwarp();
    if (threadIdx.x == 64) MIXDQ_DBG(8);    // partial tile written
    // all partial tiles of the cluster are in the workspace (release/acquire at cluster scope)
    cluster_sync_all();
    if (threadIdx.x == 64) MIXDQ_DBG(9);    // cluster barrier passed
    rows_here = BLOCK_M / splits;
    row_base = krank * rows_here;
    if (is_epi) {
      // ---- phase A': owner sums the partials of its rows (coalesced .cg loads, all `splits`
      //      loads of an item in flight together) ----
      const int et = threadIdx.x - 64;
      constexpr int C4 = BN / 4;
      const int n_items = rows_here * C4;
      // two items per iteration: 2 x splits independent 16-byte L2 loads in flight per thread
      for (int it0 = et; it0 < n_items; it0 += 2 * EPI_THREADS) {
        int4 x[2][8];
        RowInfo ri2[2];
        int rl[2], cc[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int it = it0 + u * EPI_THREADS;
          const bool live = it < n_items;
          const int c4 = live ? it / rows_here : 0;
          rl[u] = live ? it - c4 * rows_here : 0;
          cc[u] = c4 * 4;
          ri2[u] = row_info<KIND>(p, row_base + rl[u], m0, tn0, tp0, tq0);
          ri2[u].ok = ri2[u].ok && live;
          const int4* src = reinterpret_cast<const int4*>(
              ws_tile + (static_cast<int64_t>(c4) * BLOCK_M + row_base + rl[u]) * 4);
#pragma unroll
          for (int s = 0; s < 8; ++s)
            x[u][s] = (ri2[u].ok && s < splits)
                          ? __ldcg(src + static_cast<int64_t>(s) * (BLOCK_M * BN / 4))
                          : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (!ri2[u].ok) continue;
          int32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            acc[0] += x[u][s].x; acc[1] += x[u][s].y; acc[2] += x[u][s].z; acc[3] += x[u][s].w;
          }
          if (W4) { acc[0] >>= 4; acc[1] >>= 4; acc[2] >>= 4; acc[3] >>= 4; }
          const int col = cc[u];
          if (p.acc_out != nullptr && n_tile0 + col + 4 <= p.N)
            *reinterpret_cast<int4*>(p.acc_out + ri2[u].out_row * p.N + n_tile0 + col) =
                make_int4(acc[0], acc[1], acc[2], acc[3]);
          __align__(8) __half h[8];
          dequant_pack(ri2[u].cls, col, acc, acc, h, 4);
          *reinterpret_cast<uint2*>(stage_out + rl[u] * L::OUT_PITCH + col * 2) =
              *reinterpret_cast<const uint2*>(h);
        }
      }
      if (threadIdx.x == 64) MIXDQ_DBG(10);   // partials summed
    }
  }

  if (is_epi && splits > 1) {
    // ---- phase B: staging tile -> global, consecutive threads along a row (coalesced) ----
    epi_bar_sync();
    if (threadIdx.x == 64) MIXDQ_DBG(11);     // staging complete
    const int et = threadIdx.x - 64;
    constexpr int G = BN / 8;            // 16-byte chunks per row
    for (int g = et; g < rows_here * G; g += EPI_THREADS) {
      const int row_local = g / G;
      const int col = (g - row_local * G) * 8;
      if (n_tile0 + col + 8 > p.N) continue;
      const RowInfo ri = row_info<KIND>(p, row_base + row_local, m0, tn0, tp0, tq0);
      if (!ri.ok) continue;
      *reinterpret_cast<uint4*>(p.D + ri.out_row * p.ldd + n_tile0 + col) = epilogue_tail(
          p, *reinterpret_cast<const uint4*>(stage_out + row_local * L::OUT_PITCH + col * 2),
          ri.out_row, n_tile0 + col);
    }
  }
  if (threadIdx.x == 64) MIXDQ_DBG(7);    // epilogue done

  __syncwarp();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace mixdq
