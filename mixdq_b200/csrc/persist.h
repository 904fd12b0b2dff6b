// persist.h — launcher of the persistent tcgen05 kernels (tc_persist.cuh), see persist.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_kernel.cuh"

namespace mixdq {

// Tile width the persistent kernel would use for an (m_tiles x N, num_kb k-blocks) problem, or 0
// when the problem is too small to keep every SM busy for more than one tile (then the
// one-tile-per-CTA kernel with its split-K / narrow-tile heuristic is the better fit).
int persist_pick_bn(int m_tiles, int N, int num_kb, int kind);
// test / tuning hook: mode 0 = never, 1 = heuristic, 2 = whenever supported; cs = 1 / 2 (else kept)
void persist_set_mode(int mode, int cs);
// tuning hook: restrict the cost model to one tile width (0 = free choice)
void persist_force_bn(int bn);
// cluster size along M (1 or 2) for the chosen configuration
int persist_cluster_size(int m_tiles);
// tmW must have been encoded with box rows = bn / cs; tmD (used when p.d_tma != 0) is the fp16
// output as a 2-D tensor {columns, rows} with 16 x 32 boxes, SWIZZLE_32B. Returns 0 or a negative
// MIXDQ_ERR_* code.
int persist_launch(int kind, int bn, bool w4, int cs, const CUtensorMap& tmA,
                   const CUtensorMap& tmW, const CUtensorMap& tmD, TcParams p, cudaStream_t st);
// HALO form of the 3x3 convolution (see persist.cu): applicable? / launch. tmA must then have been
// encoded with a {128, boxW, boxH + 2, 1} box and p.a_tx_bytes = (boxH + 2) * boxW * 128.
void persist_set_halo(int on);      // test / tuning hook
bool persist_halo_ok(int bn, bool w4, int cs, int R, int S, int pad, int stride, int boxW, int boxH,
                     int boxN);
bool persist_halo_wanted(int m_tiles);   // problem large enough for the HALO form (its own threshold)
int persist_launch_conv_halo(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmD,
                             TcParams p, cudaStream_t st);

}  // namespace mixdq
