// tc_kernel.cuh — K-G / K-C: INT8 x INT8 -> INT32 contraction on the 5th-gen tensor cores.
//
//   tcgen05.mma.cta_group::1.kind::i8, 128 x BN x 32 per instruction, s32 accumulators in TMEM,
//   A (activations) and W (weights) tiles staged by TMA into 128B-swizzled shared memory through
//   a STAGES-deep mbarrier ring, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM
//   allocator + single-thread MMA issuer, warps 2..5 = epilogue (tcgen05.ld -> dequant -> fp16).
//
//   KIND_GEMM : A is a 2-D [M][K] matrix                       (reference qlinear, A2)
//   KIND_CONV : A is the NHWC activation tensor; each k-block is one (r,s) filter tap x 128
//               input channels fetched as a shifted 4-D TMA box with out-of-bounds zero fill —
//               implicit GEMM, no im2col buffer                 (reference qconv2d, A3)
//               the zero-point border correction of activation_zero_point_propagate (A4) is a
//               16-class x BN table built in shared memory by the epilogue warps
//   KIND_SPLIT: two K-phases (two A/W operand pairs) into two TMEM accumulators, combined in the
//               epilogue exactly like the reference's two fp16 convs + fp16 add (A6)
#pragma once
#include "common.cuh"

namespace mixdq {

enum { KIND_GEMM = 0, KIND_CONV = 1, KIND_SPLIT = 2 };

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 128;  // int8 elements = bytes = one 128B swizzle row
constexpr int UMMA_K = 32;
constexpr int TC_THREADS = 192;

struct TcParams {
  // problem
  int M, N;             // GEMM rows (conv: informational), output channels
  int num_kb;           // k-blocks of phase 0
  int num_kb1;          // k-blocks of phase 1 (KIND_SPLIT only)
  // conv geometry (KIND_CONV)
  int kb_per_tap, S, pad;
  int NB, H, W, P, Q;   // batch, input H/W, output P/Q
  int boxW, boxH, boxN; // pixels covered by one A box: boxN x boxH x boxW (<= 128 rows)
  int tilesQ, tilesP;   // tiles along q and p (tiles along n = gridDim.x / (tilesQ*tilesP))
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
  __half* D;
  int64_t ldd;
  int32_t* acc_out;       // optional raw accumulator dump [rows][N]
};

template <int BN, int STAGES, int KIND>
struct TcSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K;
  static constexpr int W_BYTES = BN * BLOCK_K;
  static constexpr int TAB_PITCH = BN + 1;
  static constexpr int TAB_FLOATS = (KIND == KIND_CONV) ? 16 * TAB_PITCH : 0;
  static constexpr int PARAM_FLOATS = (KIND == KIND_SPLIT ? 5 : 3) * BN + TAB_FLOATS;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_W = OFF_A + STAGES * A_BYTES;
  static constexpr int OFF_PARAM = OFF_W + STAGES * W_BYTES;
  static constexpr int OFF_BAR = OFF_PARAM + ((PARAM_FLOATS * 4 + 15) / 16) * 16;
  static constexpr int NUM_BARS = 2 * STAGES + 1;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024 B alignment
};

template <int BN, int STAGES, int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
             const TcParams p) {
  using L = TcSmem<BN, STAGES, KIND>;
  constexpr int TMEM_COLS_USED = (KIND == KIND_SPLIT ? 2 : 1) * BN;
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile0 = blockIdx.y * BN;

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
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int total_kb = p.num_kb + (KIND == KIND_SPLIT ? p.num_kb1 : 0);
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* a_dst = sA + stage * L::A_BYTES;
        uint8_t* w_dst = sW + stage * L::W_BYTES;
        if (KIND == KIND_GEMM) {
          mbar_expect_tx(&full_bar[stage], L::A_BYTES + L::W_BYTES);
          tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BLOCK_K, m0);
          tma_load_2d(w_dst, &tmW, &full_bar[stage], kb * BLOCK_K, n_tile0);
        } else if (KIND == KIND_SPLIT) {
          mbar_expect_tx(&full_bar[stage], L::A_BYTES + L::W_BYTES);
          if (kb < p.num_kb) {
            tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BLOCK_K, m0);
            tma_load_2d(w_dst, &tmW, &full_bar[stage], kb * BLOCK_K, n_tile0);
          } else {
            const int k1 = kb - p.num_kb;
            tma_load_2d(a_dst, &tmA1, &full_bar[stage], k1 * BLOCK_K, m0);
            tma_load_2d(w_dst, &tmW1, &full_bar[stage], k1 * BLOCK_K, n_tile0);
          }
        } else {
          const int tap = kb / p.kb_per_tap;
          const int c0 = (kb - tap * p.kb_per_tap) * BLOCK_K;
          const int r = tap / p.S, s = tap - r * p.S;
          mbar_expect_tx(&full_bar[stage], p.a_tx_bytes + L::W_BYTES);
          tma_load_4d(a_dst, &tmA, &full_bar[stage], c0, tq0 - p.pad + s, tp0 - p.pad + r, tn0);
          tma_load_3d(w_dst, &tmW, &full_bar[stage], c0, tap, n_tile0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int total_kb = p.num_kb + (KIND == KIND_SPLIT ? p.num_kb1 : 0);
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + stage * L::A_BYTES);
        const uint32_t w_addr = smem_u32(sW + stage * L::W_BYTES);
        uint32_t d_tmem = tmem_base;
        int kb_in_phase = kb;
        if (KIND == KIND_SPLIT && kb >= p.num_kb) { d_tmem += BN; kb_in_phase = kb - p.num_kb; }
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          umma_i8(d_tmem, umma_desc_sw128(a_addr + k * UMMA_K), umma_desc_sw128(w_addr + k * UMMA_K),
                  IDESC, (kb_in_phase | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);        // accumulators complete
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int et = threadIdx.x - 64;  // 0..127
    // stage per-column epilogue operands in shared memory while the mainloop runs
    for (int j = et; j < BN; j += 128) {
      const int n = n_tile0 + j;
      const bool ok = n < p.N;
      float sc = 0.f, b0 = 0.f, bs = 0.f;
      if (ok) {
        if (p.a_scale != nullptr) {           // dynamic: fold the activation scalars here
          sc = __fmul_rn(__ldg(p.scale + n), __ldg(p.a_scale));
          b0 = __fmul_rn(__ldg(p.bias0 + n), __ldg(p.a_zp));
        } else {
          sc = __ldg(p.scale + n);
          if (!(KIND == KIND_CONV) || !p.has_table) b0 = __ldg(p.bias0 + n);
        }
        if (p.bias != nullptr) bs = __half2float(p.bias[n]);
      }
      s_scale[j] = sc; s_bias0[j] = b0; s_bias[j] = bs;
      if (KIND == KIND_SPLIT) {
        s_extra[j] = ok ? __ldg(p.scale1 + n) : 0.f;
        s_extra[BN + j] = ok ? __ldg(p.bias0_1 + n) : 0.f;
      }
      if (KIND == KIND_CONV && p.has_table) {
        // 3x3 / pad 1: class = rcls*4 + scls, bit0 = first tap cut, bit1 = last tap cut
        float w9[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w9[t] = ok ? __ldg(p.bias0 + static_cast<int64_t>(n) * 9 + t) : 0.f;
        const float zp = __ldg(p.a_zp);
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
    // named barrier among the 128 epilogue threads (id 1)
    asm volatile("bar.sync 1, 128;" ::: "memory");

    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;   // row of the 128-row tile
    bool row_ok;
    int64_t out_row;                       // row index into D / acc_out
    int cls = 0;
    if (KIND == KIND_CONV) {
      const int per_img = p.boxH * p.boxW;
      const int dn = row / per_img;
      const int rem = row - dn * per_img;
      const int dh = rem / p.boxW, dw = rem - dh * p.boxW;
      const int n = tn0 + dn, pp = tp0 + dh, qq = tq0 + dw;
      row_ok = (dn < p.boxN) && (n < p.NB) && (pp < p.P) && (qq < p.Q);
      out_row = (static_cast<int64_t>(n) * p.P + pp) * p.Q + qq;
      if (p.has_table) {
        const int h0 = pp - p.pad, w0 = qq - p.pad;
        const int rc = (h0 < 0 ? 1 : 0) | (h0 + 2 >= p.H ? 2 : 0);
        const int sc4 = (w0 < 0 ? 1 : 0) | (w0 + 2 >= p.W ? 2 : 0);
        cls = rc * 4 + sc4;
      }
    } else {
      row_ok = (m0 + row) < p.M;
      out_row = m0 + row;
    }

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    const bool has_bias = p.bias != nullptr;
    constexpr int CH = (BN >= 32) ? 32 : 16;
#pragma unroll 1
    for (int c = 0; c < BN / CH; ++c) {
      uint32_t v[CH], v1[CH];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * CH;
      if (CH == 32) {
        tmem_ld_32x32(taddr, reinterpret_cast<uint32_t(&)[32]>(v));
        if (KIND == KIND_SPLIT) tmem_ld_32x32(taddr + BN, reinterpret_cast<uint32_t(&)[32]>(v1));
      } else {
        tmem_ld_32x16(taddr, reinterpret_cast<uint32_t(&)[16]>(v));
        if (KIND == KIND_SPLIT) tmem_ld_32x16(taddr + BN, reinterpret_cast<uint32_t(&)[16]>(v1));
      }
      tmem_ld_wait();
      if (row_ok) {
        const int ncol0 = n_tile0 + c * CH;
        if (p.acc_out != nullptr) {
          int32_t* arow = p.acc_out + out_row * p.N + ncol0;
#pragma unroll
          for (int j = 0; j < CH; j += 4)
            if (ncol0 + j + 4 <= p.N)
              *reinterpret_cast<int4*>(arow + j) =
                  make_int4(static_cast<int>(v[j]), static_cast<int>(v[j + 1]),
                            static_cast<int>(v[j + 2]), static_cast<int>(v[j + 3]));
        }
        __half* drow = p.D + out_row * p.ldd + ncol0;
#pragma unroll
        for (int j8 = 0; j8 < CH; j8 += 8) {
          if (ncol0 + j8 + 8 <= p.N) {
            __align__(16) __half h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int col = c * CH + j8 + j;
              float b0;
              if (KIND == KIND_CONV) b0 = p.has_table ? s_extra[cls * L::TAB_PITCH + col] : s_bias0[col];
              else b0 = s_bias0[col];
              float f = dequant_f32(static_cast<int32_t>(v[j8 + j]), b0, s_scale[col]);
              if (has_bias) f = __fadd_rn(f, s_bias[col]);
              if (KIND == KIND_SPLIT) {
                // reference: two fp16 conv outputs added in fp16 (nn/Conv2d.py:346)
                const float f1 = dequant_f32(static_cast<int32_t>(v1[j8 + j]), s_extra[BN + col], s_extra[col]);
                // (torch adds halves in fp32 opmath and rounds once more to fp16)
                h[j] = __float2half_rn(__fadd_rn(__half2float(__float2half_rn(f)),
                                                 __half2float(__float2half_rn(f1))));
              } else {
                h[j] = __float2half_rn(f);
              }
            }
            *reinterpret_cast<uint4*>(drow + j8) = *reinterpret_cast<const uint4*>(h);
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace mixdq
