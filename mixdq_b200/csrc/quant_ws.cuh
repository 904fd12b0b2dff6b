// quant_ws.cuh — device workspace and grid-wide reductions shared by the dynamic-quantisation
// kernels (quant.cu: plain tensor; fused_quant.cu: LayerNorm / GEGLU / GroupNorm producers).
//
// All of these kernels run with every CTA co-resident (grid <= SM count x resident CTAs per SM), so
// a counter-based grid barrier is safe: each CTA folds its partial min/max into two device words
// with integer atomicMax (exact and order-independent: -min and max are non-negative floats, whose
// bit patterns order like integers), bumps an arrival counter and spins on it with acquire loads;
// the last CTA to have READ the result zeroes the words again. The workspace is zero-initialised
// ONCE by the caller; every kernel leaves it zeroed on exit. Kernels that share a workspace must
// be stream-ordered (they are: one workspace per stream on the host side). Tensors small enough
// for one thread-block cluster skip the workspace altogether (cluster_minmax_params below).
#pragma once
#include <stdlib.h>
#include "common.cuh"

namespace mixdq {

constexpr int kMaxPartials = 4096;   // CTAs of one producer launch
constexpr int kMaxStatGroups = 4096; // (image, group) pairs of one GroupNorm launch

struct DynWs {
  // ---- min/max barrier ----
  unsigned int counter;   // arrivals
  unsigned int flag;      // raised by the last arriver once scale/zp are published
  unsigned int done;      // CTAs that have consumed the flag (last one resets the workspace)
  unsigned int pad0;
  // ---- GroupNorm statistics barrier ----
  unsigned int counter2;
  unsigned int flag2;
  unsigned int done2;
  unsigned int pad1;
  // ---- min/max published by a PRODUCER's epilogue (tc_i8_kernel<KIND_GEGLU>) ----
  unsigned int mm[2];     // bit patterns of -min (>= +0) and max (>= +0), atomicMax'ed as ints
  unsigned int mm_done;   // consumer CTAs that have read mm (the last one zeroes all three)
  unsigned int pad2;
  unsigned int epoch;     // launches of the tagged-partial barrier so far (quant2.cu lean kernels)
  unsigned int gmm1;      // unused
  unsigned int qdbg_seq;  // profiling: launches stamped so far (see QDbg)
  unsigned int pad3;
  unsigned long long* qdbg;   // profiling: stamp buffer or NULL
  float2 partial[kMaxPartials];
  // fixed-point (integer => order-independent, deterministic) sum / sum of squares per (n, group)
  unsigned long long gsum[2 * kMaxStatGroups];
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void spin_until_set(const unsigned int* flag) {
  unsigned int spins = 0;
  while (ld_acquire_u32(flag) == 0u) {
    __nanosleep(32);
    if (++spins > (1u << 24)) __trap();   // protocol bug: fail instead of hanging the device
  }
}

// Profiling aid: per-CTA %globaltimer stamps of the quantiser kernels, kept in registers of
// thread 0 and flushed at kernel end into ws->qdbg[(launch_seq * kQdbgMaxCtas + cta) * 8 + slot]
// (slot 0 entry, 1 dependency wait passed, 2 values loaded / min-max taken, 3 barrier passed,
// 4 done). launch_seq = ws->qdbg_seq, read after the dependency wait and bumped by CTA 0 at its
// end, so back-to-back launches (PDL, CUDA graphs) land in consecutive regions without any host
// involvement. ws->qdbg == NULL (the default) disables everything but one predicated load.
constexpr int kQdbgMaxCtas = 1024;
struct QDbg {
  unsigned long long t[5];
  unsigned long long* buf;
  unsigned int seq;
  __device__ __forceinline__ void begin(const DynWs* ws) {
    buf = nullptr;
    if (threadIdx.x == 0) {
      buf = *reinterpret_cast<unsigned long long* const volatile*>(&ws->qdbg);
      stamp(0);
    }
  }
  __device__ __forceinline__ void stamp(int i) {
    if (threadIdx.x == 0 && buf != nullptr) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t[i]));
  }
  __device__ __forceinline__ void waited(const DynWs* ws) {   // call right after pdl_wait()
    if (threadIdx.x == 0 && buf != nullptr) {
      seq = *reinterpret_cast<const volatile unsigned int*>(&ws->qdbg_seq);
      stamp(1);
    }
  }
  __device__ __forceinline__ void end(DynWs* ws) {
    if (threadIdx.x == 0 && buf != nullptr) {
      stamp(4);
      unsigned long long* dst = buf + (static_cast<size_t>(seq) * kQdbgMaxCtas + blockIdx.x) * 8;
#pragma unroll
      for (int i = 0; i < 5; ++i) dst[i] = t[i];
      if (blockIdx.x == 0) ws->qdbg_seq = seq + 1;
    }
  }
};

// qdiff asymmetric 8-bit min-max parameters (base_quantizer.py:155-190), fp32:
//   delta = max((x_max - x_min) / 255, 1e-6),  z = rint(-x_min / delta)
__device__ __forceinline__ void qdiff_params(float mn, float mx, float& delta, float& z,
                                             float qmax = 255.0f) {
  delta = __fdiv_rn(__fsub_rn(mx, mn), qmax);
  if (delta < 1e-6f) delta = 1e-6f;
  z = rintf(__fdiv_rn(-mn, delta));
}

// Grid-wide min/max -> (delta, z). Called by ALL threads of every CTA (blockDim.x = NT, a multiple
// of 32, <= 1024) with the thread's partial min (<= 0) and max (>= 0). On return every thread
// holds delta and the UNSHIFTED zero point z in [0, 255]; *scale_out = delta, *zp_out = z - 128.
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int NT>
__device__ __forceinline__ void grid_minmax_params(DynWs* __restrict__ ws, float mn, float mx,
                                                   float* __restrict__ scale_out,
                                                   float* __restrict__ zp_out, float& delta,
                                                   float& z) {
  constexpr int NW = NT / 32;
  __shared__ float smn[NW], smx[NW];
  __shared__ float s_delta, s_z;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    mn = lane < NW ? smn[lane] : 0.0f;
    mx = lane < NW ? smx[lane] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    // publish this CTA's partial and arrive (release: the partial is visible before the count);
    // three dependent L2 trips in all: arrive -> poll -> read the partials. Every CTA reduces the
    // partials itself (same order everywhere -> identical result), nobody waits for a "last" CTA.
    if (lane == 0) {
      ws->partial[blockIdx.x] = make_float2(mn, mx);
      red_release_add_u32(&ws->counter, 1u);
      unsigned int spins = 0;
      while (ld_acquire_u32(&ws->counter) < gridDim.x) {
        if (++spins > (1u << 26)) __trap();   // protocol bug: fail instead of hanging the device
      }
    }
    __syncwarp();
    mn = 0.0f; mx = 0.0f;
    for (int i = lane; i < static_cast<int>(gridDim.x); i += 32) {
      const float2 v = __ldcg(&ws->partial[i]);
      mn = fminf(mn, v.x);
      mx = fmaxf(mx, v.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
      float d, zz;
      qdiff_params(mn, mx, d, zz);
      s_delta = d;
      s_z = zz;
      if (blockIdx.x == 0) { *scale_out = d; *zp_out = zz - 128.0f; }
      // the last CTA to have read the partials re-arms the counter for the next call
      if (atomicAdd(&ws->done, 1u) == gridDim.x - 1) {
        ws->done = 0u;
        __threadfence();
        st_release_u32(&ws->counter, 0u);
      }
    }
  }
  __syncthreads();
  delta = s_delta;
  z = s_z;
}

// ------------------------------------------------------------------------------------------
// Cluster variant: the whole launch is ONE thread-block cluster (gridDim.x == cluster size, <= 16
// CTAs), so the min/max exchange goes through distributed shared memory and the hardware cluster
// barrier instead of global atomics + a spin flag (measured in-graph: the flag barrier costs
// ~4 us of the 8-10 us these kernels took at batch 1; tensors of the batch-1 transformer blocks
// are 0.6 MB and fit the stash of one cluster). No workspace is touched.
// cluster_enter() must be called by all threads at kernel entry: it arrives (without blocking) at
// the barrier whose completion tells that every CTA of the cluster has started, which DSMEM
// stores require.
// ------------------------------------------------------------------------------------------
constexpr int kMaxClusterCtas = 16;
constexpr int64_t kClusterMaxElems = 65536;   // larger tensors use every SM + the counter barrier

__device__ __forceinline__ void cluster_enter() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}

template <int NT>
__device__ __forceinline__ void cluster_minmax_params(float mn, float mx,
                                                      float* __restrict__ scale_out,
                                                      float* __restrict__ zp_out, float& delta,
                                                      float& z) {
  constexpr int NW = NT / 32;
  __shared__ float smn[NW], smx[NW];
  __shared__ float2 s_part[kMaxClusterCtas];   // slot r is written by CTA r of the cluster
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  // every CTA of the cluster is running (pairs with cluster_enter)
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  const uint32_t nc = gridDim.x;
  if (warp == 0) {
    mn = lane < NW ? smn[lane] : 0.0f;
    mx = lane < NW ? smx[lane] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (static_cast<uint32_t>(lane) < nc) {
      const uint32_t dst = dsmem_map(smem_u32(&s_part[cluster_ctarank()]), static_cast<uint32_t>(lane));
      asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(dst), "f"(mn), "f"(mx) : "memory");
    }
  }
  cluster_sync_all();   // release the DSMEM stores / acquire the peers'
  mn = 0.0f; mx = 0.0f;
  for (uint32_t i = 0; i < nc; ++i) {
    const float2 v = s_part[i];
    mn = fminf(mn, v.x);
    mx = fmaxf(mx, v.y);
  }
  qdiff_params(mn, mx, delta, z);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *scale_out = delta;
    *zp_out = z - 128.0f;
  }
}

// rint(RN(x / delta)) — the reference rounds the correctly rounded fp32 QUOTIENT
// (torch.round(x / delta), base_quantizer.py:186) — without an IEEE division per element:
// t = x * (1/delta) is within 3 ulp (< 5e-5 for |t| <= 256: |x| <= 255 delta by construction) of
// the quotient, so rint(t) is the answer unless t sits within 1e-4 of a rounding boundary; only
// those elements (~0.02 %) take the exact division. Bit-identical to the division everywhere (tests/test_gpu_ops.py).
__device__ __forceinline__ float qdiff_round_quot(float x, float delta, float inv_delta) {
  const float t = __fmul_rn(x, inv_delta);
  float r = rintf(t);
  if (fabsf(__fsub_rn(t, r)) > 0.4999f) r = rintf(__fdiv_rn(x, delta));
  return r;
}

// one qdiff code: clamp(rint(x / delta) + z, 0, 255) - 128
__device__ __forceinline__ int qdiff_code(float x, float delta, float inv_delta, float z) {
  float r = __fadd_rn(qdiff_round_quot(x, delta, inv_delta), z);
  r = fminf(fmaxf(r, 0.0f), 255.0f);
  return static_cast<int>(r) - 128;
}

// 8 halves (one 16-byte vector) -> 8 codes (one 8-byte vector)
__device__ __forceinline__ uint2 qdiff_vec8(const int4& raw, float delta, float z) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
  const float inv = __frcp_rn(delta);
  int q[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h2[i]);
    q[2 * i] = qdiff_code(f.x, delta, inv, z);
    q[2 * i + 1] = qdiff_code(f.y, delta, inv, z);
  }
  uint2 out;
  out.x = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((q[3] & 0xff) << 24);
  out.y = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((q[7] & 0xff) << 24);
  return out;
}

__device__ __forceinline__ void minmax_vec8(const int4& raw, float& mn, float& mx) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h2[j]);
    mn = fminf(mn, fminf(f.x, f.y));
    mx = fmaxf(mx, fmaxf(f.x, f.y));
  }
}

// ------------------------------------------------------------------------------------------
// host side: the largest single-cluster launch (16 = non-portable size, else 8) that this device
// can schedule for `kern` with `threads` threads and `smem` dynamic bytes per CTA; 0 = use the
// flag-barrier grid path. MIXDQ_NO_CLUSTER=1 / mixdq_debug_set_cluster(0) force 0 (A/B runs).
// ------------------------------------------------------------------------------------------
inline int& cluster_mode_flag() {
  static int mode = -1;   // -1 = read the environment on first use
  return mode;
}
inline bool cluster_enabled() {
  int& mode = cluster_mode_flag();
  if (mode < 0) {
    const char* e = getenv("MIXDQ_NO_CLUSTER");
    mode = (e && e[0] == '1') ? 0 : 1;
  }
  return mode != 0;
}
template <typename K>
inline int max_cluster_ctas(K kern, int threads, int smem) {
  (void)cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int cands[2] = {16, 8};
  for (int i = 0; i < 2; ++i) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cands[i]);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cands[i]; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n >= 1) return cands[i];
  }
  (void)cudaGetLastError();
  return 0;
}

// Host-side note "the last min/max producer launched on workspace W wrote N partials", consumed
// by the quantise pass that follows it on the same stream (separate C-ABI calls, e.g. the GEGLU
// GEMM and mixdq_quant_i8_premm). One definition shared by all translation units.
inline int& partial_count_slot(const void* ws) {
  static const void* keys[32] = {nullptr};
  static int vals[32] = {0};
  static int next_victim = 0;
  for (int i = 0; i < 32; ++i) {
    if (keys[i] == ws) return vals[i];
    if (keys[i] == nullptr) { keys[i] = ws; return vals[i]; }
  }
  // more than 32 live workspaces (one per device x stream): recycle round-robin — a producer
  // always writes its entry right before its consumer reads it
  const int v = next_victim++ & 31;
  keys[v] = ws;
  vals[v] = 0;
  return vals[v];
}

// Experiment hook: MIXDQ_CARVEOUT=<percent> pins the shared-memory carve-out preference of the
// quantiser kernels (so that the SM's L1/shared split need not change between them and the
// tcgen05 kernels, which use ~200 KB of shared memory).
template <typename K>
inline void apply_carveout_pref(K kern) {
  static int pct = -2;
  if (pct == -2) {
    const char* e = getenv("MIXDQ_CARVEOUT");
    pct = e ? atoi(e) : -1;
  }
  if (pct >= 0) (void)cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

// plain grid launch with a programmatic dependency on the preceding kernel of the stream (the
// kernel calls griddepcontrol.wait before its first global read)
template <typename K, typename... Args>
inline cudaError_t launch_pdl(K kern, int ctas, int threads, int smem, cudaStream_t st,
                              Args... args) {
  static bool pref = false;          // one instance per kernel signature is enough for the hook
  if (!pref) { apply_carveout_pref(kern); pref = true; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// one-cluster launch with a programmatic dependency on the preceding kernel of the stream
template <typename K, typename... Args>
inline cudaError_t launch_cluster_pdl(K kern, int ctas, int threads, int smem, cudaStream_t st,
                                      Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

}  // namespace mixdq
