/*
 * mixdq_b200.h — C ABI of the B200-native (sm_100a) MixDQ quantized-UNet hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types. Every entry point
 *   - enqueues work on the caller's stream and returns immediately (no allocation, no sync,
 *     CUDA-graph capturable) — the conventions of the reference extension, which launches on
 *     at::cuda::getCurrentCUDAStream() (reference kernels/mixdq_extension/csrc/quant_dequant/
 *     quantize_kernel.cu:40, qlinear/cutlassGemm_withBias_optimalAlignment.cu:208);
 *   - returns 0 on success or a negative MIXDQ_ERR_* code (see mixdq_strerror);
 *   - takes DEVICE pointers unless the parameter is documented as host.
 *
 * Each function cites the reference interface it replaces (paths relative to the reference
 * repository root, thu-nics/MixDQ @ 4f6b32ad).
 */
#ifndef MIXDQ_B200_H_
#define MIXDQ_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIXDQ_ABI_VERSION 2

/* error codes */
#define MIXDQ_OK                 0
#define MIXDQ_ERR_INVALID_ARG   -1   /* null pointer / non-positive size                      */
#define MIXDQ_ERR_ALIGNMENT     -2   /* reference: "Int8 kernel with input or output alignment
                                        not to 4 is not supported." (qlinear.cc:130-133)      */
#define MIXDQ_ERR_UNSUPPORTED   -3   /* e.g. dilation != 1 (op/qconv2d.py:120-123)            */
#define MIXDQ_ERR_CUDA          -4   /* a CUDA runtime/driver call failed (launch, tensor map) */
#define MIXDQ_ERR_WORKSPACE     -5   /* workspace too small                                    */

typedef void* mixdq_stream_t;      /* a cudaStream_t */
typedef uint16_t mixdq_half_t;     /* IEEE binary16 bit pattern (__half)  */

int         mixdq_abi_version(void);
/* Register a device scratch buffer (caller-owned, >= 16-byte aligned) for launches on `stream` of
   CUDA device `device`: split-K launches exchange their partial INT32 tiles through it (it stays
   L2-resident). A launch only ever uses the workspace registered for ITS stream, so concurrent
   streams cannot corrupt each other's partial tiles; without a workspace for the stream the
   contraction kernels never split K. NULL / 0 unregisters. Thread-safe. */
int         mixdq_set_workspace(int device, mixdq_stream_t stream, void* ptr, int64_t bytes);
/* *id_out = 0 when `stream` is not being captured into a CUDA graph, else the unique id of the
   capture sequence. Host-side caches of device results key on it: a value computed eagerly must
   not be reused inside a capture (its producing kernels would be missing from the graph). */
int         mixdq_stream_capture_id(mixdq_stream_t stream, unsigned long long* id_out);
const char* mixdq_strerror(int code);
/* Name of the kernel family the last call on this thread dispatched to ("tcgen05", "simt", ...).
   Test/diagnostic aid only. */
const char* mixdq_last_path(void);
/* Force every GEMM/conv onto the portable SIMT kernels (1) or restore heuristics (0). Test aid:
   lets the tcgen05 path be cross-checked against an independent implementation on the device. */
void        mixdq_force_simt(int on);
/* Force the output-tile width of the tcgen05 kernels (16/32/64/128/256; 0 = heuristic). Tuning and
   test aid (also settable through the MIXDQ_FORCE_BN environment variable). */
void        mixdq_debug_force_bn(int bn);
/* Persistent form of the tcgen05 kernels (tensor-bound, multi-wave problems): mode 0 = never,
   1 = heuristic (default; env MIXDQ_PERSIST), 2 = whenever the shape is supported (tests);
   cluster = 1 / 2 CTAs sharing a weight tile through TMA multicast (default 2; env
   MIXDQ_PERSIST_CS); other values leave the setting unchanged. Results are identical. */
void        mixdq_debug_set_persist(int mode, int cluster);
/* Tuning hook: restrict the persistent kernel's tile-width choice to `bn` (128 / 160 / 256; 0 =
   cost model). */
void        mixdq_debug_set_persist_bn(int bn);
/* 3x3 convolutions on the persistent kernel: one haloed A box for the three vertical taps
   (default on; 0 = plain per-tap boxes, for A/B timing and the parity tests of both forms) */
void        mixdq_debug_set_conv_halo(int on);
/* Force the split-K factor (cluster size) of the tcgen05 kernels (1/2/4/8; 0 = heuristic). */
void        mixdq_debug_force_splits(int splits);
/* Give the tcgen05 kernels a device buffer of 8 uint64 PER CTA of the largest grid launched:
   every CTA writes %globaltimer (ns) at {entry, setup done, first TMA, last TMA, first stage
   landed, last MMA issued, accumulators ready, epilogue done}. NULL disables. Profiling aid. */
void        mixdq_debug_set_timing_buffer(void* dev_ptr);
/* Enable (1, default) / disable (0) programmatic dependent launch of the tcgen05 kernels. */
void        mixdq_debug_set_pdl(int on);
/* Profiling only (results are garbage): bit0 = skip the MMA issue, bit1 = skip the TMA loads. */
void        mixdq_debug_set_mode(int mode);
/* Enable (1, default) / disable (0) the single-cluster variants of the dynamic quantisers (DSMEM +
   hardware cluster barrier + programmatic dependent launch); 0 = always the flag-barrier grid
   kernels. Results are identical either way; A/B timing and test aid (env MIXDQ_NO_CLUSTER=1). */
void        mixdq_debug_set_cluster(int on);
/* Form of the dynamic quantisers: 1 (default) = a min/max pass + a quantise pass chained by
   programmatic dependent launch; 0 = first-generation single kernels with the counter barrier
   (fallback for callers without a scratch buffer, A/B reference of the parity tests). Identical
   results in both modes (env MIXDQ_QUANT_MODE). */
void        mixdq_debug_set_two_pass(int mode);
/* Point the dynamic-quantisation workspace `ws` at a device buffer of
   launches x 1024 CTAs x 8 uint64 (or NULL = off): every quantiser launch that uses `ws` then
   stores, per CTA, %globaltimer (ns) at {entry, dependency wait passed, values loaded, barrier
   passed, done} into the region of its launch sequence number. Stream-ordered. Profiling aid. */
int         mixdq_debug_set_quant_timing_buffer(void* ws, void* dev_ptr, mixdq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A1  static per-tensor activation quantisation, fp16 -> int8
 *     q = (int8) clamp(lrintf(fmaf(x, *scale_inv, *zp)), -128, 127)
 * replaces  quantize_per_tensor_to_int8 / quantize_per_tensor_to_int8_vectorized
 *           (csrc/quant_dequant/quantize.cc:9-53; kernels quantize_kernel.cu:10-27,
 *            quantize_kernel_vectorized.cu:29-72).  scale_inv / zp are fp32 device scalars
 *            (zp already shifted by -128, nn/utils.py:428).
 * The strided form reads a 3-D view [d0][d1][cols] whose innermost dimension is contiguous
 * (element strides s0, s1) and writes a dense [d0*d1][cols] int8 array with row pitch
 * `out_pitch` elements (>= cols). It covers channel slices of NHWC tensors (split shortcuts,
 * nn/Conv2d.py:313-318) and the BOS token slice x[:,1:,:] (nn/Linear.py:180).
 * ---------------------------------------------------------------------------------------- */
int mixdq_quant_i8_static(const mixdq_half_t* x, int64_t numel,
                          const float* scale_inv, const float* zp,
                          int8_t* q, mixdq_stream_t stream);

int mixdq_quant_i8_static_strided(const mixdq_half_t* x, int64_t d0, int64_t d1, int64_t cols,
                                  int64_t s0, int64_t s1,
                                  const float* scale_inv, const float* zp,
                                  int8_t* q, int64_t out_pitch, mixdq_stream_t stream);

/* A1 + the int8 NCHW -> NHWC copy of qconv2d.cc:91-92 fused: reads fp16 logical [N][C][H][W]
 * with arbitrary element strides `xstride[4]` (host array), takes channels [c_begin, c_end),
 * writes dense int8 NHWC [N][H][W][c_end-c_begin]. */
int mixdq_quant_i8_nchw2nhwc(const mixdq_half_t* x, int N, int C, int H, int W,
                             const int64_t xstride[4], int c_begin, int c_end,
                             const float* scale_inv, const float* zp,
                             int8_t* q_nhwc, mixdq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A10 dynamic per-tensor activation quantisation (qdiff min-max, asymmetric, 8 bit)
 *     x_min = min(min(x),0), x_max = max(max(x),0), delta = max((x_max-x_min)/255, 1e-6)
 *     z = rint(-x_min/delta); q = clamp(rint(x/delta) + z, 0, 255) - 128
 * restates  BaseQuantizer.init_quant_params / forward
 *           (quant_utils/qdiff/quantizer/base_quantizer.py:155-190, 122-128) in fp32.
 * Outputs: q int8 (dense, same element order as x), *scale_out = delta,
 *          *zp_out = z - 128 (the "kernel format" zero point of nn/utils.py:428).
 * `ws` is a device workspace of at least mixdq_quant_dynamic_ws_bytes() bytes that must be
 * zero-initialised ONCE by the caller; the kernels leave it zeroed again on exit.
 * ---------------------------------------------------------------------------------------- */
int64_t mixdq_quant_dynamic_ws_bytes(void);
int mixdq_quant_i8_dynamic(const mixdq_half_t* x, int64_t numel,
                           float* scale_out, float* zp_out, int8_t* q,
                           void* ws, mixdq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A2  W8A8 linear:  D[m,n] = half( (float(sum_k A[m,k]*W[n,k]) - bias0[n]) * scale[n]
 *                                  (+ float(bias[n])) )     each step one fp32 RN operation
 * replaces  qlinear_w8_a8_ohalf (csrc/qlinear/qlinear.cc:14-137) and the four CUTLASS
 *           instantiations cutlassGemm_{withBias,noBias}_{optimal,small}Alignment.cu
 *           (epilogue order: cutlassGemm_withBias_optimalAlignment.cu:39-95).
 * A int8 [M][K] row pitch lda; W int8 [N][K] dense; D fp16 [M][N] row pitch ldd.
 * K % 4 == 0 and N % 4 == 0 required (MIXDQ_ERR_ALIGNMENT otherwise, as the reference).
 * `acc_out` (nullable, int32 [M][N] dense) additionally receives the raw INT32 accumulators —
 * test hook for the bit-exact parity check; production callers pass NULL.
 * ---------------------------------------------------------------------------------------- */
int mixdq_gemm_w8a8_f16(const int8_t* A, int64_t lda, const int8_t* W,
                        const float* bias0, const float* scale, const mixdq_half_t* bias,
                        mixdq_half_t* D, int64_t ldd, int M, int N, int K,
                        int32_t* acc_out, mixdq_stream_t stream);

/* Dynamic-scale variant: the epilogue forms scale[n] = w_scale[n] * (*a_scale) and
 * bias0[n] = wsum[n] * (*a_zp) itself from the device scalars a dynamic quantise call produced
 * (the arguments the reference signature carries but never reads, qlinear.cc:25-73). */
int mixdq_gemm_w8a8_f16_dyn(const int8_t* A, int64_t lda, const int8_t* W,
                            const float* w_scale, const float* wsum,
                            const float* a_scale, const float* a_zp, const mixdq_half_t* bias,
                            mixdq_half_t* D, int64_t ldd, int M, int N, int K,
                            int32_t* acc_out, mixdq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * W4A8 (north star "W4A8 layers unpacked on the fly"; the reference has no 4-bit kernel: its
 * 4-/2-bit layers fall back to fp16, nn/Linear.py:28-36, nn/Conv2d.py:37-61). Arithmetic = the
 * qdiff 4-bit symmetric per-channel weight quantiser (base_quantizer.py:119-185) feeding the
 * same integer identity and epilogue as A2 / A3.
 * W_packed uint8 [N][K/2] (conv: KRSC [K][R][S][C/2]): two signed 4-bit two's-complement codes
 * per byte, EVEN k (c) in the HIGH nibble — the nibble order of nn/utils.py:26-28. The packed
 * tiles travel through TMA as they are (half the weight bytes of W8) and are expanded to int8
 * inside the tcgen05 kernel. K % 32 == 0 (conv: C % 32 == 0) required.
 * mixdq_gemm_w4a8_f16 falls back to a portable kernel when the tcgen05 alignment conditions do
 * not hold; the other entry points return MIXDQ_ERR_ALIGNMENT (keep such layers as W8 codes).
 * ---------------------------------------------------------------------------------------- */
int mixdq_gemm_w4a8_f16(const int8_t* A, int64_t lda, const uint8_t* W_packed,
                        const float* bias0, const float* scale, const mixdq_half_t* bias,
                        mixdq_half_t* D, int64_t ldd, int M, int N, int K,
                        int32_t* acc_out, mixdq_stream_t stream);
/* dynamic activation scalars + optional fused residual add (see mixdq_gemm_w8a8_f16_dyn_res) */
int mixdq_gemm_w4a8_f16_dyn_res(const int8_t* A, int64_t lda, const uint8_t* W_packed,
                                const float* w_scale, const float* wsum,
                                const float* a_scale, const float* a_zp, const mixdq_half_t* bias,
                                const mixdq_half_t* residual, int64_t ldr,
                                mixdq_half_t* D, int64_t ldd, int M, int N, int K,
                                int32_t* acc_out, mixdq_stream_t stream);
/* GEGLU projection with packed 4-bit weights (see mixdq_gemm_w8a8_geglu_f16_dyn; rows interleaved
   BEFORE packing) */
int mixdq_gemm_w4a8_geglu_f16_dyn(const int8_t* A, int64_t lda, const uint8_t* W_il_packed,
                                  const float* w_scale_il, const float* wsum_il,
                                  const float* a_scale, const float* a_zp,
                                  const mixdq_half_t* bias_il, mixdq_half_t* Y, int64_t ldy,
                                  int M, int N2, int K, void* ws, mixdq_stream_t stream);
/* convolutions, static / dynamic (see mixdq_conv_w8a8_f16 / mixdq_conv_w8a8_f16_dyn) */
int mixdq_conv_w4a8_f16(const int8_t* x_nhwc, int64_t x_cpitch, const uint8_t* w_krsc_packed,
                        const float* scale, const float* wsum_krs, const float* bias0_k,
                        const float* zp, const mixdq_half_t* bias, mixdq_half_t* y_nhwc,
                        int N, int H, int W, int C, int K, int R, int S, int stride, int pad,
                        int32_t* acc_out, mixdq_stream_t stream);
int mixdq_conv_w4a8_f16_dyn(const int8_t* x_nhwc, int64_t x_cpitch, const uint8_t* w_krsc_packed,
                            const float* w_scale, const float* wsum_krs, const float* wsum_k,
                            const float* a_scale, const float* a_zp, const mixdq_half_t* bias,
                            const mixdq_half_t* chan_add, int64_t ldca,
                            const mixdq_half_t* residual, mixdq_half_t* y_nhwc, int N, int H,
                            int W, int C, int K, int R, int S, int stride, int pad,
                            int32_t* acc_out, mixdq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A3 + A4  W8A8 conv2d fprop, NHWC, cross-correlation, dilation 1, square stride/padding,
 *          with the activation zero-point propagation folded into the epilogue:
 *   y[n,p,q,k] = half( (float(acc) - zpw) * scale[k] (+ float(bias[k])) )
 *   zpw = (pad>0) ? float(sum_{(r,s) in bounds} wsum_krs[k,r,s]) * (*zp) : bias0_k[k]
 * replaces  qconv2d_w8_a8_ohalf (csrc/qconv2d/qconv2d.cc:28-206), the eight CUTLASS conv
 *           instantiations (cutlassConv2d_*.cu) and activation_zero_point_propagate
 *           (conv_act_zero_point_propagate.cu:10-83) — the fp32 [N,P,Q,K] correction tensor is
 *           never materialised.
 * x int8 NHWC [N][H][W][C] with pixel pitch `x_cpitch` elements (>= C; lets a channel slice of
 * a wider NHWC tensor be consumed in place); w int8 KRSC [K][R][S][C] dense; y fp16 NHWC
 * [N][P][Q][K] dense. C % 4 == 0 and K % 4 == 0 required.
 * wsum_krs fp32 [K][R][S] is required iff pad > 0, bias0_k fp32 [K] iff pad == 0
 * (qconv2d.cc:117-124).
 * ---------------------------------------------------------------------------------------- */
int mixdq_conv_w8a8_f16(const int8_t* x_nhwc, int64_t x_cpitch, const int8_t* w_krsc,
                        const float* scale, const float* wsum_krs, const float* bias0_k,
                        const float* zp, const mixdq_half_t* bias, mixdq_half_t* y_nhwc,
                        int N, int H, int W, int C, int K, int R, int S, int stride, int pad,
                        int32_t* acc_out, mixdq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A6  split 1x1 shortcut convolution (the nine up-block conv_shortcut layers): the two channel
 * halves carry independent activation (scale, zp) and weight scales; the reference runs two
 * convs and adds their fp16 results (nn/Conv2d.py:312-347). Fused here as
 *   y = half( half_rn(t0) + half_rn(t1) )  with  t0 = (acc0-bias0_a[k])*scale_a[k] + bias[k],
 *                                                t1 = (acc1-bias0_b[k])*scale_b[k]
 * i.e. exactly the reference's two fp16 roundings followed by an fp16 add, in one kernel with
 * two accumulators. xa/xb are int8 [M][Ca] / [M][Cb] with row pitches lda/ldb.
 * ---------------------------------------------------------------------------------------- */
int mixdq_conv1x1_split_w8a8_f16(const int8_t* xa, int64_t lda, const int8_t* wa, int Ca,
                                 const float* bias0_a, const float* scale_a,
                                 const int8_t* xb, int64_t ldb, const int8_t* wb, int Cb,
                                 const float* bias0_b, const float* scale_b,
                                 const mixdq_half_t* bias, mixdq_half_t* y, int64_t ldy,
                                 int M, int K, mixdq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dynamic-scale variants with the fused elementwise tail. The reference model follows many
 * quantized layers with an fp16 elementwise add (`x + attn(...)`, `x + ff(...)`, resnet
 * `h + temb[:, :, None, None]` and `x + h`; diffusers modules around nn/Linear.py / nn/Conv2d.py
 * outputs). Each such add is applied in the epilogue exactly as PyTorch would apply it to the
 * layer's fp16 output: D = half(float(D) + float(other)), one rounding per add.
 *   residual  fp16 [rows][N] with row pitch ldr (conv: dense NHWC [N*P*Q][K])
 *   chan_add  fp16 [images][K] with row pitch ldca (conv only: per-image channel vector,
 *             applied before residual)
 * Dynamic mode: `w_scale`/`wsum*` are the per-output-channel weight scale and weight code sums;
 * the epilogue forms scale = w_scale * (*a_scale) and the zero-point correction from *a_zp
 * (device scalars written by a dynamic quantise call).
 * These entry points require the tcgen05 path (16-byte aligned pointers, K % 16 == 0 (C for
 * conv), N % 8 == 0): MIXDQ_ERR_ALIGNMENT otherwise.
 * ---------------------------------------------------------------------------------------- */
int mixdq_gemm_w8a8_f16_dyn_res(const int8_t* A, int64_t lda, const int8_t* W,
                                const float* w_scale, const float* wsum,
                                const float* a_scale, const float* a_zp, const mixdq_half_t* bias,
                                const mixdq_half_t* residual, int64_t ldr,
                                mixdq_half_t* D, int64_t ldd, int M, int N, int K,
                                int32_t* acc_out, mixdq_stream_t stream);

/* GEGLU feed-forward input projection (diffusers GEGLU = Linear(d, 2*inner) then
   hidden * gelu(gate); the reference runs it as QuantizedLinear nn/Linear.py:147-192 followed by
   stock fp16 ops) with the GEGLU evaluated in the GEMM epilogue. The caller passes the weight
   ROWS (and w_scale / wsum / bias entries) interleaved in groups of 16: rows 32g..32g+15 are
   value rows 16g..16g+15, rows 32g+16..32g+31 the gate rows inner+16g..inner+16g+15, so that 32
   adjacent accumulator columns hold both operands of 16 outputs. Y = fp16 [M][N2/2] (row pitch
   ldy): half(h * half(gelu(g))) of the fp16-rounded linear outputs — identical to
   mixdq_gemm_w8a8_f16_dyn followed by the stock GEGLU. The tensor's min(0, min Y) / max(0, max Y)
   are folded into the dynamic-quantisation workspace `ws`, to be consumed by exactly one
   mixdq_quant_i8_premm call on the same stream (which leaves the workspace clean again).
   N2 % 32 == 0, K % 16 == 0; MIXDQ_ERR_UNSUPPORTED when the tcgen05 path is disabled. */
int mixdq_gemm_w8a8_geglu_f16_dyn(const int8_t* A, int64_t lda, const int8_t* W_il,
                                  const float* w_scale_il, const float* wsum_il,
                                  const float* a_scale, const float* a_zp,
                                  const mixdq_half_t* bias_il, mixdq_half_t* Y, int64_t ldy,
                                  int M, int N2, int K, void* ws, mixdq_stream_t stream);
/* A10 (qdiff min-max, base_quantizer.py:155-190) of a dense fp16 tensor whose min / max were
   published into `ws` by the producing kernel (see above): single pass, no grid barrier.
   numel % 8 == 0. Writes *scale_out = delta, *zp_out = z - 128, q = codes. */
int mixdq_quant_i8_premm(const mixdq_half_t* x, int64_t numel, int8_t* q, float* scale_out,
                         float* zp_out, void* ws, mixdq_stream_t stream);

/* wsum_krs fp32 [K][R][S] iff pad > 0; wsum_k fp32 [K] (sum over taps and channels) iff pad == 0 */
int mixdq_conv_w8a8_f16_dyn(const int8_t* x_nhwc, int64_t x_cpitch, const int8_t* w_krsc,
                            const float* w_scale, const float* wsum_krs, const float* wsum_k,
                            const float* a_scale, const float* a_zp, const mixdq_half_t* bias,
                            const mixdq_half_t* chan_add, int64_t ldca,
                            const mixdq_half_t* residual, mixdq_half_t* y_nhwc, int N, int H, int W, int C, int K, int R, int S,
                            int stride, int pad, int32_t* acc_out, mixdq_stream_t stream);

int mixdq_conv1x1_split_w8a8_f16_dyn(const int8_t* xa, int64_t lda, const int8_t* wa, int Ca,
                                     const float* wsum_a, const float* w_scale_a,
                                     const float* a_scale_a, const float* a_zp_a,
                                     const int8_t* xb, int64_t ldb, const int8_t* wb, int Cb,
                                     const float* wsum_b, const float* w_scale_b,
                                     const float* a_scale_b, const float* a_zp_b,
                                     const mixdq_half_t* bias, const mixdq_half_t* residual,
                                     int64_t ldr, mixdq_half_t* y, int64_t ldy, int M, int K,
                                     mixdq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Producer-side fusion of the dynamic quantise pass: the stock fp16 op that
 * produces a quantized layer's input and the A10 quantisation of its result as ONE kernel.
 * Each reproduces the unfused PyTorch sequence, rounding to fp16 wherever PyTorch materialises
 * an fp16 tensor, then applies the A10 formula to those fp16 values (same outputs as
 * mixdq_quant_i8_dynamic: q dense int8, *scale_out = delta, *zp_out = z - 128).
 *   ln     : y = half(gamma * (rstd * (x - mean)) + beta), rows of C, biased variance, fp32 stats
 *            (torch.nn.LayerNorm feeding attn1.to_q/k/v, attn2.to_q, ff.net.0.proj)
 *   geglu  : hg = [M][2*I]; y = half(float(h) * float(half(gelu(g))))  exact-erf GELU
 *            (diffusers GEGLU between ff.net.0.proj and ff.net.2)
 *   gn     : NHWC x [NB][HW][C] (pixel pitch ldx), G groups; y = half(fma(x, a, b)),
 *            a = rstd*gamma, b = beta - mean*a; silu != 0: y = half(y / (1 + exp(-y)))
 *            (torch.nn.GroupNorm [+ SiLU] feeding resnet conv1/conv2 and Transformer2D.proj_in)
 * `y_out` (nullable, dense fp16) additionally receives the fp16 values that were quantised —
 * test hook / for an un-quantised consumer. `ws` as for mixdq_quant_i8_dynamic.
 * Restrictions: C % 8 == 0, 16-byte aligned pointers; ln: C <= 2048; gn: C <= 2560, G <= 32,
 * C/G >= 8 or == 4, NB <= 148 (MIXDQ_ERR_UNSUPPORTED otherwise: run the ops unfused).
 * ---------------------------------------------------------------------------------------- */
/* A10 on a row-pitched view: x fp16 [M][cols] with row pitch ldx -> dense int8 [M][cols]
 * (channel slices of NHWC tensors for the split shortcuts nn/Conv2d.py:313-318, token slices
 * nn/Linear.py:180, strided attention outputs). cols % 8 == 0, ldx % 8 == 0. */
int mixdq_quant_i8_dynamic_rows(const mixdq_half_t* x, int64_t ldx, int M, int cols, int8_t* q,
                                float* scale_out, float* zp_out, void* ws, mixdq_stream_t stream);
/* A10 with a code range (N3: the 4-bit activation layers of kernels/cfgs/act/act_7.xx.yaml, which
   the reference gates to fp16, nn/Linear.py:28-36 / nn/Conv2d.py:38-46). n_bits = 8: identical to
   mixdq_quant_i8_dynamic_rows. n_bits = 4: delta = (max - min) / 15, z = round(-min / delta),
   q = clamp(round(x / delta) + z, 0, 15) stored as it is (no -128 shift), *zp_out = z: the codes
   feed the same int8 kernels. cols % 8 == 0, ldx % 8 == 0. */
int mixdq_quant_i8_dynamic_bits(const mixdq_half_t* x, int64_t ldx, int64_t M, int64_t cols,
                                int n_bits, float* scale_out, float* zp_out, int8_t* q, void* ws,
                                mixdq_stream_t stream);
/* out[0] = min(0, min x), out[1] = max(0, max x) of a dense fp16 tensor (numel % 8 == 0): the
   clamped range the qdiff quantizers start from (base_quantizer.py:155-158), used by the PTQ
   calibration (mixdq_b200/ptq.py; reference scripts/ptq.py:126-155). */
int mixdq_minmax_f16(const mixdq_half_t* x, int64_t numel, float* out, void* ws,
                     mixdq_stream_t stream);
/* A1 with an explicit code range: q = clamp(lrintf(x * scale_inv + zp), lo, hi), flat dense.
   [lo, hi] = [0, 15] with the unshifted zero point for static 4-bit activations. */
int mixdq_quant_i8_static_range(const mixdq_half_t* x, int64_t numel, const float* scale_inv,
                                const float* zp, int lo, int hi, int8_t* q,
                                mixdq_stream_t stream);
/* mixdq_ln_quant_i8_dynamic / mixdq_gn_quant_i8_dynamic with q == NULL (y_out required): normalise
   only — y_out receives the fp16 LayerNorm / GroupNorm[+SiLU] output and nothing is quantised
   (static-scale callers quantise y_out with mixdq_quant_i8_static and the layer's checkpoint
   parameters; scale_out / zp_out may then be NULL). */
int mixdq_ln_quant_i8_dynamic(const mixdq_half_t* x, int64_t ldx, int M, int C,
                              const mixdq_half_t* gamma, const mixdq_half_t* beta, float eps,
                              int8_t* q, mixdq_half_t* y_out, float* scale_out, float* zp_out,
                              void* ws, mixdq_stream_t stream);
int mixdq_geglu_quant_i8_dynamic(const mixdq_half_t* hg, int64_t ld, int M, int I, int8_t* q,
                                 mixdq_half_t* y_out, float* scale_out, float* zp_out, void* ws,
                                 mixdq_stream_t stream);
int mixdq_gn_quant_i8_dynamic(const mixdq_half_t* x, int64_t ldx, int NB, int HW, int C, int G,
                              const mixdq_half_t* gamma, const mixdq_half_t* beta, float eps,
                              int silu, int8_t* q, mixdq_half_t* y_out, float* scale_out,
                              float* zp_out, void* ws, mixdq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Producers with STATIC (PTQ-checkpoint) scales of the consumer: the reference quantises the
 * fp16 output of the producing op with the layer's stored parameters
 * (mixdq_extension.op.quantize_per_tensor_to_int8 at nn/Linear.py:154-176, nn/Conv2d.py:282-347;
 * formula csrc/quant_dequant/quantize_kernel.cu:20-24). Here the producer applies that formula to
 * its own fp16-rounded values before they leave the SM: q = sat8(rint(fma(y, *scale_inv, *zp))),
 * bit-identical to the dynamic-path producer followed by mixdq_quant_i8_static on its y_out.
 *   ln    : one kernel                     (LayerNorm -> to_q/k/v, attn2.to_q, ff.net.0.proj)
 *   gn    : statistics kernel + apply kernel (GroupNorm[+SiLU] -> resnet convs, proj_in)
 *   geglu : the GEGLU GEMM (rows interleaved as for mixdq_gemm_w8a8_geglu_f16_dyn; w_bits = 8:
 *           int8 codes, 4: packed nibbles) writes Q [M][N2/2] int8 (row pitch ldq, % 16) -> ff.net.2
 * Same shape restrictions as the dynamic entry points; q dense.
 * ---------------------------------------------------------------------------------------- */
int mixdq_ln_quant_i8_static(const mixdq_half_t* x, int64_t ldx, int M, int C,
                             const mixdq_half_t* gamma, const mixdq_half_t* beta, float eps,
                             const float* scale_inv, const float* zp, int8_t* q, void* ws,
                             mixdq_stream_t stream);
int mixdq_gn_quant_i8_static(const mixdq_half_t* x, int64_t ldx, int NB, int HW, int C, int G,
                             const mixdq_half_t* gamma, const mixdq_half_t* beta, float eps,
                             int silu, const float* scale_inv, const float* zp, int8_t* q,
                             void* ws, mixdq_stream_t stream);
int mixdq_gemm_geglu_i8_static(const int8_t* A, int64_t lda, const void* W_il, int w_bits,
                               const float* w_scale_il, const float* wsum_il,
                               const float* a_scale, const float* a_zp,
                               const mixdq_half_t* bias_il, const float* q_scale_inv,
                               const float* q_zp, int8_t* Q, int64_t ldq, int M, int N2, int K,
                               void* ws, mixdq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MIXDQ_B200_H_ */
