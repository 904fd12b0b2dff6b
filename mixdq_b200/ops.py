"""Host side of the op layer: the reference's `mixdq_extension._C` functions, re-implemented as a
thin Python shim over the C ABI (include/mixdq_b200.h).

Same names, argument order, checks and error behaviour as the reference's pybind module
(reference kernels/mixdq_extension/csrc/main.cpp:9-13, quant_dequant/quantize.cc:9-62,
qlinear/qlinear.cc:14-234, qconv2d/qconv2d.cc:28-235): outputs are allocated here with
torch.empty* on the caching allocator, work is enqueued on the current CUDA stream without any
synchronisation (CUDA-graph capturable), inputs are never mutated, violations raise RuntimeError.
PyTorch is plumbing only (device memory + streams); all arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

_c_int64_4 = ctypes.c_int64 * 4


def _check(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ---- launch accounting (bench.py: `gpu_launches`, kernel-family replay for the roofline) --------
_launch_count = 0
_recorder = None   # when a list: every launch appends (family, algo_bytes, algo_ops, replay, keep)


def launch_count() -> int:
    """Number of kernels of this library launched (or captured) so far in this process."""
    return _launch_count


def start_recording() -> list:
    global _recorder
    _recorder = []
    return _recorder


def stop_recording() -> list:
    global _recorder
    rec, _recorder = _recorder, None
    return rec


_workspaces = {}
WORKSPACE_BYTES = 32 << 20


def _ensure_workspace(device: torch.device) -> None:
    """Split-K scratch (stays L2-resident) of the CURRENT stream of `device`: allocated once per
    (device, stream) through torch and registered with the library, which never allocates and
    only uses a stream's own workspace (two streams cannot corrupt each other's partial tiles)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(device).cuda_stream
    if (idx, stream) not in _workspaces:
        ws = torch.empty(WORKSPACE_BYTES, dtype=torch.uint8, device=torch.device("cuda", idx))
        _lib.check(_lib.load().mixdq_set_workspace(idx, stream, ws.data_ptr(), ws.numel()))
        _workspaces[(idx, stream)] = ws


def prepare_stream(device: torch.device) -> None:
    """Allocate the per-stream workspaces of the current stream of `device` now. Call it on a
    capture stream BEFORE the capture begins, so the scratch buffers live outside the graph's
    private memory pool (mixdq.cuda_graph_opt does)."""
    _ensure_workspace(device)
    _dynamic_workspace(device)


def _capture_id(device: torch.device) -> int:
    """0 outside CUDA-graph capture, else the id of the capture the current stream belongs to."""
    return _lib.capture_id(torch.cuda.current_stream(device).cuda_stream)


def _version_of(t: torch.Tensor):
    """In-place-write counter of `t`, or None for inference tensors (which do not track one)."""
    try:
        return t._version
    except RuntimeError:
        return None


def _launch(family: str, fn, args: tuple, ref: torch.Tensor, kernels: int = 1, keep=(),
            algo_bytes: int = 0, algo_ops: int = 0) -> None:
    """Call one C-ABI entry point (stream appended as the last argument)."""
    global _launch_count
    if family[0] in "gc":          # gemm / conv families may split K
        _ensure_workspace(ref.device)
    _lib.check(fn(*args, _stream(ref)))
    _launch_count += kernels
    if _recorder is not None:
        dev = ref.device

        def replay(fn=fn, args=args, dev=dev):
            _lib.check(fn(*args, torch.cuda.current_stream(dev).cuda_stream))
        _recorder.append((family, algo_bytes, algo_ops, replay, keep, kernels))


def _gemm_bytes(M, N, K_in, NK_bytes):
    """activations once + weights once + fp16 output + 10 B/channel of epilogue
    vectors."""
    return M * K_in + NK_bytes + 2 * M * N + 10 * N


def _is_dense(t: torch.Tensor) -> bool:
    """non-overlapping and dense in some dimension order (what empty_like preserves)."""
    if t.is_contiguous():
        return True
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return True
    dims = sorted(range(t.dim()), key=lambda d: (t.stride(d), t.size(d)))
    expect = 1
    for d in dims:
        if t.size(d) == 1:
            continue
        if t.stride(d) != expect:
            return False
        expect *= t.size(d)
    return True


class _DeviceGuard:
    __slots__ = ("dev", "prev")

    def __init__(self, t: torch.Tensor):
        self.dev = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.dev is not None and cur != self.dev:
            self.prev = cur
            torch.cuda.set_device(self.dev)

    def __exit__(self, *a):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


# ---------------------------------------------------------------------------------------------
# A1  quantize_per_tensor_to_int8[_vectorized]
# ---------------------------------------------------------------------------------------------
def _check_quant_args(input, scale_inv, zero_point):
    _check(input.device.type == "cuda", "input should be on CUDA")
    _check(input.device == scale_inv.device, "input and scale should be on the same device")
    _check(input.device == zero_point.device,
           "input and zero_point should be on the same device")
    _check(input.dtype == torch.float16, "input should be fp16")
    _check(scale_inv.dtype == torch.float32, "scale_inv should be fp32")
    _check(zero_point.dtype == torch.float32, "zero_point should be fp32")


def quantize_per_tensor_to_int8(input: torch.Tensor, scale_inv: torch.Tensor,
                                zero_point: torch.Tensor) -> torch.Tensor:
    """q = int8(clamp(lrintf(x * scale_inv + zero_point), -128, 127)); returns empty_like(input).

    Reference: quantize.cc:9-30 (+ kernel quantize_kernel.cu:10-27). Unlike the reference, which
    walks `data_ptr()` linearly over `numel` whatever the strides are, a non-dense view
    (x[:, :split], x[:, 1:, :] at batch > 1) is read through its strides and the result is the
    logically correct dense tensor.
    """
    _check_quant_args(input, scale_inv, zero_point)
    lib = _lib.load()
    with _DeviceGuard(input):
        if _is_dense(input):
            out = torch.empty_like(input, dtype=torch.int8)
            _launch("quant", lib.mixdq_quant_i8_static,
                    (input.data_ptr(), input.numel(), scale_inv.data_ptr(), zero_point.data_ptr(),
                     out.data_ptr()), input, keep=(input, scale_inv, zero_point, out),
                    algo_bytes=3 * input.numel())
            return out
        return _quantize_view(input, scale_inv, zero_point)


# one kernel family serves both reference entry points
quantize_per_tensor_to_int8_vectorized = quantize_per_tensor_to_int8


def _quantize_view(x: torch.Tensor, scale_inv, zero_point) -> torch.Tensor:
    """Non-dense views. NHWC channel slices and [B, T', K] token slices go through the strided
    kernel in place; anything else is densified first."""
    lib = _lib.load()
    keep = (x, scale_inv, zero_point)
    if x.dim() == 4 and x.stride(1) == 1 and x.stride(3) >= x.size(1) \
            and x.stride(2) == x.size(3) * x.stride(3) and x.stride(0) == x.size(2) * x.stride(2):
        # channel slice of an NHWC tensor -> dense NHWC int8 (logical NCHW, channels_last)
        n, c, h, w = x.shape
        out = torch.empty((n, c, h, w), dtype=torch.int8, device=x.device,
                          memory_format=torch.channels_last)
        _launch("quant", lib.mixdq_quant_i8_static_strided,
                (x.data_ptr(), 1, n * h * w, c, 0, x.stride(3), scale_inv.data_ptr(),
                 zero_point.data_ptr(), out.data_ptr(), c), x, keep=keep + (out,),
                algo_bytes=3 * out.numel())
        return out
    if x.dim() == 3 and x.stride(2) == 1:
        b, t, k = x.shape
        out = torch.empty((b, t, k), dtype=torch.int8, device=x.device)
        _launch("quant", lib.mixdq_quant_i8_static_strided,
                (x.data_ptr(), b, t, k, x.stride(0), x.stride(1), scale_inv.data_ptr(),
                 zero_point.data_ptr(), out.data_ptr(), k), x, keep=keep + (out,),
                algo_bytes=3 * out.numel())
        return out
    xc = x.contiguous()
    out = torch.empty_like(xc, dtype=torch.int8)
    _launch("quant", lib.mixdq_quant_i8_static,
            (xc.data_ptr(), xc.numel(), scale_inv.data_ptr(), zero_point.data_ptr(),
             out.data_ptr()), xc, keep=(xc, scale_inv, zero_point, out), algo_bytes=3 * xc.numel())
    return out


def quantize_to_nhwc(input: torch.Tensor, scale_inv: torch.Tensor, zero_point: torch.Tensor,
                     c_begin: int = 0, c_end: Optional[int] = None) -> torch.Tensor:
    """Fused quantize + layout: fp16 [N,C,H,W] (any strides) channels [c_begin,c_end) -> int8
    channels_last. Replaces quantize + the int8 `.contiguous(ChannelsLast)` copy of
    qconv2d.cc:91-92, and the `x[:, :split]` slicing of nn/Conv2d.py:313-318."""
    _check_quant_args(input, scale_inv, zero_point)
    _check(input.dim() == 4, "input should be 4-D")
    n, c, h, w = input.shape
    c_end = c if c_end is None else c_end
    _check(0 <= c_begin < c_end <= c, "bad channel range")
    lib = _lib.load()
    csel = c_end - c_begin
    out = torch.empty((n, csel, h, w), dtype=torch.int8, device=input.device,
                      memory_format=torch.channels_last)
    keep = (input, scale_inv, zero_point, out)
    with _DeviceGuard(input):
        nhwc = input.stride(1) == 1 and input.stride(2) == w * input.stride(3) \
            and input.stride(0) == h * input.stride(2)
        if nhwc:
            base = input.data_ptr() + 2 * c_begin
            _launch("quant", lib.mixdq_quant_i8_static_strided,
                    (base, 1, n * h * w, csel, 0, input.stride(3), scale_inv.data_ptr(),
                     zero_point.data_ptr(), out.data_ptr(), csel), input, keep=keep,
                    algo_bytes=3 * out.numel())
        else:
            strides = _c_int64_4(*input.stride())
            _launch("quant", lib.mixdq_quant_i8_nchw2nhwc,
                    (input.data_ptr(), n, c, h, w, strides, c_begin, c_end, scale_inv.data_ptr(),
                     zero_point.data_ptr(), out.data_ptr()), input, keep=keep + (strides,),
                    algo_bytes=3 * out.numel())
    return out


# ---------------------------------------------------------------------------------------------
# A10  dynamic (qdiff min-max) quantisation
# ---------------------------------------------------------------------------------------------
_dyn_ws = {}


def _dynamic_workspace(device: torch.device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _dyn_ws.get(key)
    if ws is None:
        nbytes = _lib.load().mixdq_quant_dynamic_ws_bytes()
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _dyn_ws[key] = ws
    return ws


# Layers that consume the SAME tensor (attn1.to_q/to_k/to_v share the normalised hidden state;
# all 140 attn2.to_k/to_v of the SDXL UNet share `encoder_hidden_states`) would each recompute the
# identical min/max and codes: a tiny identity-keyed cache returns the first result instead.
# Keys are the tensor OBJECT (held strongly, so its storage cannot be recycled) and its version
# counter (bumped by any in-place write).
DYNAMIC_QUANT_CACHE = True
_dyn_cache = []          # [(tensor, version, capture id, (q, scale, zp))], most recent last
_DYN_CACHE_SLOTS = 3


def _dyn_kernels(numel: int) -> int:
    """Kernels one dynamic quantisation launches: tiny tensors take the one-cluster kernel, the
    rest a min/max pass + a quantise pass (csrc/quant2.cu)."""
    return 1 if numel <= 65536 else 2


def clear_dynamic_quant_cache() -> None:
    _dyn_cache.clear()


def quantize_per_tensor_dynamic(input: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """qdiff asymmetric 8-bit min-max quantisation of one tensor (base_quantizer.py:155-190).
    Returns (q int8, scale fp32[], zero_point fp32[] (shifted by -128)).

    The identity cache is keyed on (tensor object, version counter, CUDA-graph capture id): a
    result computed eagerly is never reused inside a capture (its kernels would be missing from
    the graph and every replay would read the stale codes), nor across two captures. Inference
    tensors carry no version counter and are not cached."""
    _check(input.device.type == "cuda", "input should be on CUDA")
    _check(input.dtype == torch.float16, "input should be fp16")
    ver = _version_of(input) if DYNAMIC_QUANT_CACHE else None
    cap = _capture_id(input.device) if ver is not None else 0
    if ver is not None:
        for ref, v, c, res in _dyn_cache:
            if ref is input and v == ver and c == cap:
                return res
    lib = _lib.load()
    x = input if _is_dense(input) else input.contiguous()
    out = torch.empty_like(x, dtype=torch.int8)
    qp = torch.empty(2, dtype=torch.float32, device=x.device)
    with _DeviceGuard(x):
        ws = _dynamic_workspace(x.device)
        _launch("quant_dyn", lib.mixdq_quant_i8_dynamic,
                (x.data_ptr(), x.numel(), qp.data_ptr(), qp.data_ptr() + 4, out.data_ptr(),
                 ws.data_ptr()), x, kernels=_dyn_kernels(x.numel()), keep=(x, qp, out, ws),
                algo_bytes=3 * x.numel())
    res = (out, qp[0], qp[1])
    if ver is not None:
        _dyn_cache.append((input, ver, cap, res))
        if len(_dyn_cache) > _DYN_CACHE_SLOTS:
            _dyn_cache.pop(0)
    return res


def quantize_per_tensor_dynamic_bits(input: torch.Tensor, n_bits: int
                                     ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """qdiff asymmetric n-bit min-max quantisation (base_quantizer.py:155-190) in the kernel format.
    n_bits = 8: same as quantize_per_tensor_dynamic. n_bits = 4 (N3, the 4-bit activation layers of
    kernels/cfgs/act/act_7.xx.yaml that the reference runs in fp16, nn/Linear.py:28-36): codes
    0..15 stored in int8 as they are, zero_point = z (unshifted) — consumed by the same int8
    kernels. Returns (q int8, scale fp32[], zero_point fp32[])."""
    if n_bits == 8:
        return quantize_per_tensor_dynamic(input)
    _check(n_bits == 4, "activation bit width should be 8 or 4")
    _check(input.device.type == "cuda", "input should be on CUDA")
    _check(input.dtype == torch.float16, "input should be fp16")
    x = input if _is_dense(input) else input.contiguous()
    _check(x.numel() % 8 == 0 and x.numel() > 0, "4-bit activation quantisation needs numel % 8 == 0")
    lib = _lib.load()
    out = torch.empty_like(x, dtype=torch.int8)
    qp = torch.empty(2, dtype=torch.float32, device=x.device)
    with _DeviceGuard(x):
        ws = _dynamic_workspace(x.device)
        _launch("quant_dyn", lib.mixdq_quant_i8_dynamic_bits,
                (x.data_ptr(), x.numel(), 1, x.numel(), 4, qp.data_ptr(), qp.data_ptr() + 4,
                 out.data_ptr(), ws.data_ptr()), x, kernels=2, keep=(x, qp, out, ws),
                algo_bytes=3 * x.numel())
    return out, qp[0], qp[1]


def quantize_per_tensor_to_int4_codes(input: torch.Tensor, scale_inv: torch.Tensor,
                                      zero_point: torch.Tensor) -> torch.Tensor:
    """Static 4-bit activation codes: q = clamp(lrintf(x * scale_inv + zero_point), 0, 15) with
    the UNSHIFTED zero point of the PTQ checkpoint (zero_point_list[1]), one code per int8."""
    _check_quant_args(input, scale_inv, zero_point)
    x = input if _is_dense(input) else input.contiguous()
    lib = _lib.load()
    out = torch.empty_like(x, dtype=torch.int8)
    with _DeviceGuard(x):
        _launch("quant", lib.mixdq_quant_i8_static_range,
                (x.data_ptr(), x.numel(), scale_inv.data_ptr(), zero_point.data_ptr(), 0, 15,
                 out.data_ptr()), x, keep=(x, scale_inv, zero_point, out),
                algo_bytes=3 * x.numel())
    return out


# ---------------------------------------------------------------------------------------------
# A2  qlinear_w8_a8_ohalf
# ---------------------------------------------------------------------------------------------
def _check_linear_common(input_int8, weight_int8, weight_scale, input_scale, input_zero_point,
                         weight_sum_by_input_channels, bias):
    dev = input_int8.device
    _check(dev.type == "cuda", "Input should be on GPU.")
    _check(dev == weight_int8.device, "input and weight_int8 should be on the same device.")
    _check(dev == weight_scale.device, "input and weight_scale should be on the same device.")
    _check(dev == input_scale.device, "input and input_scale should be on the same device.")
    _check(dev == input_zero_point.device,
           "input and input_zero_point should be on the same device.")
    _check(dev == weight_sum_by_input_channels.device,
           "input and input_zero_point should be on the same device.")
    if bias is not None:
        _check(dev == bias.device, "input and bias should be on the same device.")
    _check(input_int8.dtype == torch.int8, "input_int8 should be int8 type")
    _check(weight_int8.dtype == torch.int8, "weight_int8 should be int8 type")
    _check(weight_scale.dtype == torch.float32,
           "Currently only support weight_scale with float32 type")
    _check(input_scale.dtype == torch.float32,
           "Currently only support input_scale with float32 type")
    _check(input_zero_point.dtype == torch.float32,
           "Currently only support input_zero_point with float32 type")
    _check(weight_sum_by_input_channels.dtype == torch.float32,
           "Currently only support weight_sum_by_input_channels with float32 type")
    if bias is not None:
        _check(bias.dtype == torch.float16, "Currently only support bias with float16 type")


def qlinear_w8_a8_ohalf(input_int8, weight_int8, weight_scale, input_scale, input_zero_point,
                        weight_sum_by_input_channels, scale, bias0, bias=None,
                        _acc_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """D = half((float(A @ W^T) - bias0) * scale [+ bias]); reference qlinear.cc:14-137."""
    _check_linear_common(input_int8, weight_int8, weight_scale, input_scale, input_zero_point,
                         weight_sum_by_input_channels, bias)
    N, K = weight_int8.shape[0], weight_int8.shape[1]
    _check(weight_scale.numel() == N,
           "The size of the weight_scale vector should be equal to output_channels.")
    _check(weight_sum_by_input_channels.numel() == N,
           "The size of weight_sum_by_input_channels should equal output_channels.")
    if bias is not None:
        _check(bias.numel() == N,
               "The size of the bias vector should be equal to output_channels.")
    _check(input_int8.size(-1) == K,
           f"The last dimension of input and weight should match, got {input_int8.size(-1)} "
           f"and {K}.")
    _check(scale.dtype == torch.float32 and bias0.dtype == torch.float32
           and scale.numel() == N and bias0.numel() == N,
           "scale and bias0 should be float32 vectors of size output_channels.")
    a = input_int8.contiguous()
    w = weight_int8.contiguous()
    M = a.numel() // K if K else 0
    out = torch.empty((*input_int8.shape[:-1], N), dtype=torch.float16, device=a.device)
    lib = _lib.load()
    with _DeviceGuard(a):
        _launch("gemm", lib.mixdq_gemm_w8a8_f16,
                (a.data_ptr(), K, w.data_ptr(), bias0.data_ptr(), scale.data_ptr(), _ptr(bias),
                 out.data_ptr(), N, M, N, K, _ptr(_acc_out)), a,
                keep=(a, w, bias0, scale, bias, out, _acc_out),
                algo_bytes=_gemm_bytes(M, N, K, N * K), algo_ops=2 * M * N * K)
    return out


def qlinear_w8_a8_ohalf_dynamic(input_int8, weight_int8, weight_scale, input_scale,
                                input_zero_point, weight_sum_by_input_channels, bias=None,
                                _acc_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Dynamic-scale variant: scale = weight_scale*input_scale and bias0 = wsum*input_zp are
    formed in the kernel epilogue from device scalars."""
    _check_linear_common(input_int8, weight_int8, weight_scale, input_scale, input_zero_point,
                         weight_sum_by_input_channels, bias)
    N, K = weight_int8.shape
    a = input_int8.contiguous()
    w = weight_int8.contiguous()
    M = a.numel() // K
    out = torch.empty((*input_int8.shape[:-1], N), dtype=torch.float16, device=a.device)
    lib = _lib.load()
    with _DeviceGuard(a):
        _launch("gemm", lib.mixdq_gemm_w8a8_f16_dyn,
                (a.data_ptr(), K, w.data_ptr(), weight_scale.data_ptr(),
                 weight_sum_by_input_channels.data_ptr(), input_scale.data_ptr(),
                 input_zero_point.data_ptr(), _ptr(bias), out.data_ptr(), N, M, N, K,
                 _ptr(_acc_out)), a,
                keep=(a, w, weight_scale, weight_sum_by_input_channels, input_scale,
                      input_zero_point, bias, out, _acc_out),
                algo_bytes=_gemm_bytes(M, N, K, N * K), algo_ops=2 * M * N * K)
    return out


def qlinear_w4_a8_ohalf(input_int8, weight_packed, scale, bias0, bias=None,
                        _acc_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """W4A8: weight_packed uint8 [N, K/2], even k in the high nibble, signed 4-bit codes."""
    _check(input_int8.dtype == torch.int8, "input_int8 should be int8 type")
    _check(weight_packed.dtype == torch.uint8, "weight_packed should be uint8 type")
    N, K = weight_packed.shape[0], weight_packed.shape[1] * 2
    _check(input_int8.size(-1) == K, "The last dimension of input and weight should match")
    a = input_int8.contiguous()
    w = weight_packed.contiguous()
    M = a.numel() // K
    out = torch.empty((*input_int8.shape[:-1], N), dtype=torch.float16, device=a.device)
    lib = _lib.load()
    with _DeviceGuard(a):
        _launch("gemm_w4", lib.mixdq_gemm_w4a8_f16,
                (a.data_ptr(), K, w.data_ptr(), bias0.data_ptr(), scale.data_ptr(), _ptr(bias),
                 out.data_ptr(), N, M, N, K, _ptr(_acc_out)), a,
                keep=(a, w, bias0, scale, bias, out, _acc_out),
                algo_bytes=_gemm_bytes(M, N, K, N * K // 2), algo_ops=2 * M * N * K)
    return out


def _is_w4(weight: torch.Tensor) -> bool:
    """Packed 4-bit weights are uint8 (two codes per byte); int8 tensors are W8 codes."""
    return weight.dtype == torch.uint8


def qlinear_fp_reference(input: torch.Tensor, weight: torch.Tensor,
                         bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Debug fp16 GEMM input[M,K] @ weight[K,N] (qlinear.cc:141-204; the reference ignores bias).
    A library GEMM: not part of the quantized hot path."""
    _check(input.dtype == torch.float16 and weight.dtype == torch.float16,
           "input and weight should be fp16")
    return torch.matmul(input, weight)


# ---------------------------------------------------------------------------------------------
# A3 + A4  qconv2d_w8_a8_ohalf
# ---------------------------------------------------------------------------------------------
def _nhwc_pitch(x: torch.Tensor) -> Optional[int]:
    """channel pitch if x (logical NCHW) is an NHWC tensor or a channel slice of one."""
    n, c, h, w = x.shape
    if x.stride(1) != 1 and c != 1:
        return None
    pitch = x.stride(3)
    if pitch < c or x.stride(2) != w * pitch or x.stride(0) != h * w * pitch:
        return None
    return pitch


def qconv2d_w8_a8_ohalf(input_int8, weight_int8, weight_scale, input_scale, input_zero_point,
                        scale, weight_sum_by_input_channels=None, bias0=None, bias=None,
                        stride: Optional[int] = 1, padding: Optional[int] = 0,
                        dilation: Optional[int] = 1,
                        _acc_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """INT8 NHWC conv2d fprop with fused dequant; reference qconv2d.cc:28-206. Returns fp16
    logical [N,K,P,Q] in channels_last memory."""
    dev = input_int8.device
    _check(dev.type == "cuda", "Input should be on GPU.")
    _check(dev == weight_int8.device, "input and weight_int8 should be on the same device.")
    _check(dev == weight_scale.device, "input and weight_scale should be on the same device.")
    _check(dev == input_scale.device, "input and input_scale should be on the same device.")
    _check(dev == input_zero_point.device,
           "input and input_zero_point should be on the same device.")
    if weight_sum_by_input_channels is not None:
        _check(dev == weight_sum_by_input_channels.device,
               "input and weight_sum_by_input_channels should be on the same device.")
    if bias0 is not None:
        _check(dev == bias0.device, "input and bias0 should be on the same device.")
    if bias is not None:
        _check(dev == bias.device, "input and bias should be on the same device.")
    _check(input_int8.dtype == torch.int8, "input_int8 should be int8 type")
    w4 = _is_w4(weight_int8)     # extension: packed 4-bit weights, uint8 [K, C/2, R, S]
    _check(w4 or weight_int8.dtype == torch.int8, "weight_int8 should be int8 type")
    _check(weight_scale.dtype == torch.float32,
           "Currently only support weight_scale with float32 type")
    _check(input_scale.dtype == torch.float32,
           "Currently only support input_scale with float32 type")
    _check(input_zero_point.dtype == torch.float32,
           "Currently only support input_zero_point with float32 type")
    if weight_sum_by_input_channels is not None:
        _check(weight_sum_by_input_channels.dtype == torch.float32,
               "Currently only support weight_sum_by_input_channels with float32 type")
    if bias0 is not None:
        _check(bias0.dtype == torch.float32, "Currently only support bias0 with float32 type")
    if bias is not None:
        _check(bias.dtype == torch.float16, "Currently only support bias with float16 type")

    stride = 1 if stride is None else int(stride)
    padding = 0 if padding is None else int(padding)
    dilation = 1 if dilation is None else int(dilation)
    _check(dilation == 1, "dilation > 1 is not supported")  # reference: "has bugs" op/qconv2d.py:120

    n, c, h, w = input_int8.shape
    k, _, r, s = weight_int8.shape
    _check(weight_int8.shape[1] * (2 if w4 else 1) == c,
           "input and weight channel counts should match")
    p = (h + 2 * padding - dilation * (r - 1) - 1) // stride + 1
    q = (w + 2 * padding - dilation * (s - 1) - 1) // stride + 1
    _check(weight_scale.numel() == k,
           "The size of the weight_scale vector should be equal to output_channels.")
    if padding == 0:
        _check(bias0 is not None and bias0.numel() == k,
               "The size of bias0 should equal output_channels.")
    else:
        _check(weight_sum_by_input_channels is not None
               and weight_sum_by_input_channels.numel() == k * r * s,
               "The size of weight_sum_by_input_channels should equal K*R*S.")
    if bias is not None:
        _check(bias.numel() == k,
               "The size of the bias vector should be equal to output_channels.")

    pitch = _nhwc_pitch(input_int8)
    x = input_int8
    if pitch is None:
        x = input_int8.contiguous(memory_format=torch.channels_last)
        pitch = c
    wt = weight_int8.contiguous(memory_format=torch.channels_last)
    wsum = None
    if padding > 0:
        wsum = weight_sum_by_input_channels.contiguous()
    out = torch.empty((n, k, p, q), dtype=torch.float16, device=dev,
                      memory_format=torch.channels_last)
    lib = _lib.load()
    with _DeviceGuard(x):
        _launch("conv_w4" if w4 else "conv",
                lib.mixdq_conv_w4a8_f16 if w4 else lib.mixdq_conv_w8a8_f16,
                (x.data_ptr(), pitch, wt.data_ptr(), scale.data_ptr(), _ptr(wsum),
                 _ptr(bias0) if padding == 0 else None, input_zero_point.data_ptr(), _ptr(bias),
                 out.data_ptr(), n, h, w, c, k, r, s, stride, padding, _ptr(_acc_out)), x,
                keep=(x, wt, scale, wsum, bias0, input_zero_point, bias, out, _acc_out),
                algo_bytes=_gemm_bytes(n * p * q, k, c * h * w // max(p * q, 1),
                                       k * c * r * s // (2 if w4 else 1)),
                algo_ops=2 * n * p * q * k * c * r * s)
    return out


def qconv1x1_split_w8_a8_ohalf(xa_int8, wa_int8, scale_a, bias0_a, xb_int8, wb_int8, scale_b,
                               bias0_b, bias=None) -> torch.Tensor:
    """Fused split shortcut (nn/Conv2d.py:312-347): two 1x1 convs over channel halves with
    independent quantisation parameters, summed as the reference does. x* are int8 logical
    [N,C*,H,W] NHWC (or NHWC channel slices); w* int8 [K,C*,1,1]."""
    n, ca, h, w = xa_int8.shape
    cb = xb_int8.shape[1]
    k = wa_int8.shape[0]
    pa, pb = _nhwc_pitch(xa_int8), _nhwc_pitch(xb_int8)
    if pa is None:
        xa_int8 = xa_int8.contiguous(memory_format=torch.channels_last); pa = ca
    if pb is None:
        xb_int8 = xb_int8.contiguous(memory_format=torch.channels_last); pb = cb
    wa = wa_int8.reshape(k, ca).contiguous()
    wb = wb_int8.reshape(k, cb).contiguous()
    out = torch.empty((n, k, h, w), dtype=torch.float16, device=xa_int8.device,
                      memory_format=torch.channels_last)
    lib = _lib.load()
    with _DeviceGuard(xa_int8):
        m = n * h * w
        _launch("conv_split", lib.mixdq_conv1x1_split_w8a8_f16,
                (xa_int8.data_ptr(), pa, wa.data_ptr(), ca, bias0_a.data_ptr(), scale_a.data_ptr(),
                 xb_int8.data_ptr(), pb, wb.data_ptr(), cb, bias0_b.data_ptr(), scale_b.data_ptr(),
                 _ptr(bias), out.data_ptr(), k, m, k), xa_int8,
                keep=(xa_int8, wa, bias0_a, scale_a, xb_int8, wb, bias0_b, scale_b, bias, out),
                algo_bytes=_gemm_bytes(m, k, ca + cb, k * (ca + cb)) + 8 * k,
                algo_ops=2 * m * k * (ca + cb))
    return out


# ---------------------------------------------------------------------------------------------
# dynamic-scale layers with the fused elementwise tail, and producer-fused quantisation
# (include/mixdq_b200.h "Dynamic-scale variants" / "Producer-side fusion")
# ---------------------------------------------------------------------------------------------
def _qp_pair(device):
    qp = torch.empty(2, dtype=torch.float32, device=device)
    return qp, qp[0], qp[1]


def qlinear_dynamic_fused(input_int8, weight_int8, weight_scale, input_scale, input_zero_point,
                          weight_sum, bias=None, residual=None,
                          _acc_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qlinear_w8_a8_ohalf_dynamic followed by `+ residual` (fp16, same shape as the output)
    inside the epilogue: out = half(float(half(linear)) + float(residual)). `weight_int8` may be
    a PACKED 4-bit weight (uint8 [N, K/2]): it then runs as W4A8, unpacked inside the kernel."""
    w4 = _is_w4(weight_int8)
    N, K = weight_int8.shape[0], weight_int8.shape[1] * (2 if w4 else 1)
    a = input_int8 if input_int8.is_contiguous() else input_int8.contiguous()
    M = a.numel() // K
    out = torch.empty((*input_int8.shape[:-1], N), dtype=torch.float16, device=a.device)
    res_ptr, ldr = None, 0
    if residual is not None:
        _check(residual.dtype == torch.float16 and residual.numel() == M * N,
               "residual should be fp16 of the output's shape")
        r2 = residual.reshape(M, N)
        if r2.stride(1) != 1:
            r2 = r2.contiguous()
        res_ptr, ldr = r2.data_ptr(), r2.stride(0) if M > 1 else N
        residual = r2
    lib = _lib.load()
    with _DeviceGuard(a):
        _launch("gemm_w4" if w4 else "gemm",
                lib.mixdq_gemm_w4a8_f16_dyn_res if w4 else lib.mixdq_gemm_w8a8_f16_dyn_res,
                (a.data_ptr(), K, weight_int8.data_ptr(), weight_scale.data_ptr(),
                 weight_sum.data_ptr(), input_scale.data_ptr(), input_zero_point.data_ptr(),
                 _ptr(bias), res_ptr, ldr, out.data_ptr(), N, M, N, K, _ptr(_acc_out)), a,
                keep=(a, weight_int8, weight_scale, weight_sum, input_scale, input_zero_point,
                      bias, residual, out, _acc_out),
                algo_bytes=_gemm_bytes(M, N, K, N * K // (2 if w4 else 1))
                + (2 * M * N if residual is not None else 0),
                algo_ops=2 * M * N * K)
    return out


def qconv2d_dynamic_fused(input_int8, weight_int8, weight_scale, input_scale, input_zero_point,
                          wsum_krs=None, wsum_k=None, bias=None, stride: int = 1, padding: int = 0,
                          chan_add=None, residual=None,
                          _acc_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Dynamic-scale conv (the activation scalars are folded in the epilogue) with the optional
    fused tails: `+ chan_add[:, :, None, None]` (fp16 [N, K]) then `+ residual` (fp16 NHWC
    [N, K, P, Q]). input_int8 / weight_int8 must be channels_last. `weight_int8` may be a PACKED
    4-bit weight (uint8 [K, C/2, R, S] channels_last = KRS(C/2) in memory)."""
    w4 = _is_w4(weight_int8)
    n, c, h, w = input_int8.shape
    k, _, r, s = weight_int8.shape
    p = (h + 2 * padding - r) // stride + 1
    q = (w + 2 * padding - s) // stride + 1
    pitch = _nhwc_pitch(input_int8)
    x = input_int8
    if pitch is None:
        x = input_int8.contiguous(memory_format=torch.channels_last)
        pitch = c
    wt = weight_int8.contiguous(memory_format=torch.channels_last)
    out = torch.empty((n, k, p, q), dtype=torch.float16, device=x.device,
                      memory_format=torch.channels_last)
    if residual is not None:
        _check(residual.shape == out.shape and residual.dtype == torch.float16,
               "residual should be fp16 of the output's shape")
        if not residual.is_contiguous(memory_format=torch.channels_last):
            residual = residual.contiguous(memory_format=torch.channels_last)
    ldca = 0
    if chan_add is not None:
        _check(chan_add.dtype == torch.float16 and tuple(chan_add.shape) == (n, k),
               "chan_add should be fp16 [N, K]")
        if chan_add.stride(1) != 1 or chan_add.stride(0) % 8 != 0 or chan_add.data_ptr() % 16:
            chan_add = chan_add.contiguous()
        ldca = chan_add.stride(0) if n > 1 else k
    lib = _lib.load()
    with _DeviceGuard(x):
        _launch("conv_w4" if w4 else "conv",
                lib.mixdq_conv_w4a8_f16_dyn if w4 else lib.mixdq_conv_w8a8_f16_dyn,
                (x.data_ptr(), pitch, wt.data_ptr(), weight_scale.data_ptr(), _ptr(wsum_krs),
                 _ptr(wsum_k), input_scale.data_ptr(), input_zero_point.data_ptr(), _ptr(bias),
                 _ptr(chan_add), ldca, _ptr(residual), out.data_ptr(), n, h, w, c, k, r, s, stride,
                 padding, _ptr(_acc_out)), x,
                keep=(x, wt, weight_scale, wsum_krs, wsum_k, input_scale, input_zero_point, bias,
                      chan_add, residual, out, _acc_out),
                algo_bytes=_gemm_bytes(n * p * q, k, c * h * w // max(p * q, 1),
                                       k * c * r * s // (2 if w4 else 1))
                + (2 * n * p * q * k if residual is not None else 0),
                algo_ops=2 * n * p * q * k * c * r * s)
    return out


def qconv1x1_split_dynamic_fused(xa_int8, wa_int8, w_scale_a, wsum_a, a_scale_a, a_zp_a,
                                 xb_int8, wb_int8, w_scale_b, wsum_b, a_scale_b, a_zp_b,
                                 bias=None, residual=None) -> torch.Tensor:
    """Dynamic split shortcut: two channel halves with their own dynamic (scale, zp), one kernel,
    two accumulators, combined as the reference combines its two fp16 convs."""
    n, ca, h, w = xa_int8.shape
    cb = xb_int8.shape[1]
    k = wa_int8.shape[0]
    pa, pb = _nhwc_pitch(xa_int8), _nhwc_pitch(xb_int8)
    if pa is None:
        xa_int8 = xa_int8.contiguous(memory_format=torch.channels_last); pa = ca
    if pb is None:
        xb_int8 = xb_int8.contiguous(memory_format=torch.channels_last); pb = cb
    wa = wa_int8.reshape(k, ca)
    wb = wb_int8.reshape(k, cb)
    out = torch.empty((n, k, h, w), dtype=torch.float16, device=xa_int8.device,
                      memory_format=torch.channels_last)
    if residual is not None and not residual.is_contiguous(memory_format=torch.channels_last):
        residual = residual.contiguous(memory_format=torch.channels_last)
    lib = _lib.load()
    with _DeviceGuard(xa_int8):
        m = n * h * w
        _launch("conv_split", lib.mixdq_conv1x1_split_w8a8_f16_dyn,
                (xa_int8.data_ptr(), pa, wa.data_ptr(), ca, wsum_a.data_ptr(), w_scale_a.data_ptr(),
                 a_scale_a.data_ptr(), a_zp_a.data_ptr(),
                 xb_int8.data_ptr(), pb, wb.data_ptr(), cb, wsum_b.data_ptr(), w_scale_b.data_ptr(),
                 a_scale_b.data_ptr(), a_zp_b.data_ptr(), _ptr(bias), _ptr(residual), k,
                 out.data_ptr(), k, m, k), xa_int8,
                keep=(xa_int8, wa, wsum_a, w_scale_a, a_scale_a, a_zp_a, xb_int8, wb, wsum_b,
                      w_scale_b, a_scale_b, a_zp_b, bias, residual, out),
                algo_bytes=_gemm_bytes(m, k, ca + cb, k * (ca + cb)) + 8 * k,
                algo_ops=2 * m * k * (ca + cb))
    return out


def layernorm_quantize_dynamic(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                               eps: float, return_y: bool = False):
    """LayerNorm over the last dim + qdiff dynamic quantisation: a LayerNorm kernel that also
    publishes min/max, then the single-pass quantiser (the fp16 LayerNorm output lives in a
    scratch tensor that stays in L2). Returns (q int8 [..., C], scale, zero_point[, y fp16])."""
    _check(x.dtype == torch.float16 and weight.dtype == torch.float16
           and bias.dtype == torch.float16, "layernorm_quantize_dynamic expects fp16 tensors")
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    # LayerNorm pass + quantise pass meet in an fp16 scratch tensor (stays in L2)
    y = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    qp, sc, zp = _qp_pair(x.device)
    lib = _lib.load()
    with _DeviceGuard(x2):
        ws = _dynamic_workspace(x2.device)
        _launch("ln_quant", lib.mixdq_ln_quant_i8_dynamic,
                (x2.data_ptr(), x2.stride(0) if M > 1 else C, M, C, weight.data_ptr(),
                 bias.data_ptr(), float(eps), q.data_ptr(), _ptr(y), qp.data_ptr(),
                 qp.data_ptr() + 4, ws.data_ptr()), x2, keep=(x2, weight, bias, q, y, qp, ws),
                kernels=_dyn_kernels(M * C), algo_bytes=3 * M * C)
    return (q, sc, zp, y) if return_y else (q, sc, zp)


def geglu_quantize_dynamic(hg: torch.Tensor, return_y: bool = False):
    """GEGLU (h * gelu(gate), halves of the last dim) + dynamic quantisation in one kernel."""
    _check(hg.dtype == torch.float16, "geglu_quantize_dynamic expects fp16")
    I2 = hg.shape[-1]
    I = I2 // 2
    h2 = hg.reshape(-1, I2)
    if h2.stride(1) != 1:
        h2 = h2.contiguous()
    M = h2.shape[0]
    q = torch.empty((*hg.shape[:-1], I), dtype=torch.int8, device=hg.device)
    y = torch.empty((*hg.shape[:-1], I), dtype=torch.float16, device=hg.device) if return_y else None
    qp, sc, zp = _qp_pair(hg.device)
    lib = _lib.load()
    with _DeviceGuard(h2):
        ws = _dynamic_workspace(h2.device)
        _launch("geglu_quant", lib.mixdq_geglu_quant_i8_dynamic,
                (h2.data_ptr(), h2.stride(0) if M > 1 else I2, M, I, q.data_ptr(), _ptr(y),
                 qp.data_ptr(), qp.data_ptr() + 4, ws.data_ptr()), h2, keep=(h2, q, y, qp, ws),
                algo_bytes=5 * M * I)
    return (q, sc, zp, y) if return_y else (q, sc, zp)


def geglu_interleave_index(inner: int, device=None) -> torch.Tensor:
    """Row order of the GEGLU projection for mixdq_gemm_w8a8_geglu_f16_dyn: groups of 16 value
    rows followed by their 16 gate rows. Returns the gather index [2*inner] into the stock
    [value rows | gate rows] order."""
    _check(inner % 16 == 0, "GEGLU epilogue fusion needs inner_dim % 16 == 0")
    g = torch.arange(inner // 16, device=device).view(-1, 1, 1) * 16
    j = torch.arange(16, device=device).view(1, 1, -1)
    half = torch.tensor([0, inner], device=device).view(1, 2, 1)
    return (g + half + j).reshape(-1)


def qlinear_geglu_quantize_dynamic(input_int8, weight_il, weight_scale_il, input_scale,
                                   input_zero_point, weight_sum_il, bias_il=None,
                                   return_y: bool = False):
    """ff.net.0.proj (dynamic W8A8, rows interleaved by `geglu_interleave_index`) with the GEGLU in
    the GEMM epilogue, then the single-pass quantiser fed by the epilogue's min/max:
    2 kernels for Linear -> GEGLU -> quantise. Returns (q int8 [..., inner], scale, zp[, y]).
    `weight_il` may be PACKED 4-bit (uint8 [N2, K/2], rows interleaved before packing)."""
    w4 = _is_w4(weight_il)
    N2, K = weight_il.shape[0], weight_il.shape[1] * (2 if w4 else 1)
    I = N2 // 2
    a = input_int8 if input_int8.is_contiguous() else input_int8.contiguous()
    M = a.numel() // K
    y = torch.empty((*input_int8.shape[:-1], I), dtype=torch.float16, device=a.device)
    q = torch.empty((*input_int8.shape[:-1], I), dtype=torch.int8, device=a.device)
    qp, sc, zp = _qp_pair(a.device)
    lib = _lib.load()
    with _DeviceGuard(a):
        ws = _dynamic_workspace(a.device)
        _launch("gemm_geglu_w4" if w4 else "gemm_geglu",
                lib.mixdq_gemm_w4a8_geglu_f16_dyn if w4 else lib.mixdq_gemm_w8a8_geglu_f16_dyn,
                (a.data_ptr(), K, weight_il.data_ptr(), weight_scale_il.data_ptr(),
                 weight_sum_il.data_ptr(), input_scale.data_ptr(), input_zero_point.data_ptr(),
                 _ptr(bias_il), y.data_ptr(), I, M, N2, K, ws.data_ptr()), a,
                keep=(a, weight_il, weight_scale_il, weight_sum_il, input_scale,
                      input_zero_point, bias_il, y, ws),
                algo_bytes=M * K + N2 * K // (2 if w4 else 1) + 2 * M * I + 10 * N2,
                algo_ops=2 * M * N2 * K)
        _launch("quant_premm", lib.mixdq_quant_i8_premm,
                (y.data_ptr(), y.numel(), q.data_ptr(), qp.data_ptr(), qp.data_ptr() + 4,
                 ws.data_ptr()), y, keep=(y, q, qp, ws), algo_bytes=3 * y.numel())
    return (q, sc, zp, y) if return_y else (q, sc, zp)


def groupnorm_quantize_dynamic(x: torch.Tensor, num_groups: int, weight: torch.Tensor,
                               bias: torch.Tensor, eps: float, silu: bool, return_y: bool = False):
    """GroupNorm [+ SiLU] + dynamic quantisation in one kernel. x: fp16 logical [N,C,H,W] in
    channels_last memory. Returns (q int8 [N,C,H,W] channels_last, scale, zero_point[, y]).
    Raises RuntimeError("unsupported configuration") for group shapes the kernel does not take."""
    _check(x.dtype == torch.float16 and x.dim() == 4, "groupnorm_quantize_dynamic expects fp16 4-D")
    n, c, h, w = x.shape
    if _nhwc_pitch(x) != c:
        x = x.contiguous(memory_format=torch.channels_last)
    q = torch.empty((n, c, h, w), dtype=torch.int8, device=x.device,
                    memory_format=torch.channels_last)
    # scratch of the three-kernel form (statistics, apply + min/max, quantise); stays in L2
    y = torch.empty((n, c, h, w), dtype=torch.float16, device=x.device,
                    memory_format=torch.channels_last)
    qp, sc, zp = _qp_pair(x.device)
    lib = _lib.load()
    with _DeviceGuard(x):
        ws = _dynamic_workspace(x.device)
        _launch("gn_quant", lib.mixdq_gn_quant_i8_dynamic,
                (x.data_ptr(), c, n, h * w, c, num_groups, weight.data_ptr(), bias.data_ptr(),
                 float(eps), 1 if silu else 0, q.data_ptr(), _ptr(y), qp.data_ptr(),
                 qp.data_ptr() + 4, ws.data_ptr()), x, keep=(x, weight, bias, q, y, qp, ws),
                kernels=3, algo_bytes=3 * x.numel())
    return (q, sc, zp, y) if return_y else (q, sc, zp)


# ---- static-scale producers (fused blocks with a PTQ checkpoint) --------------------------------
# The normalisation kernels of the dynamic path write their fp16 result into a tensor; with static
# (checkpoint) activation parameters that tensor is quantised by the reference's own formula
# (quantize_per_tensor_to_int8: one FMA, round, saturate) — no min/max, no partials.
def layernorm_fp16(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float
                   ) -> torch.Tensor:
    """LayerNorm over the last dim, fp16 in / fp16 out, by the LayerNorm kernel of the fused path."""
    _check(x.dtype == torch.float16 and weight.dtype == torch.float16
           and bias.dtype == torch.float16, "layernorm_fp16 expects fp16 tensors")
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    y = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lib = _lib.load()
    with _DeviceGuard(x2):
        ws = _dynamic_workspace(x2.device)
        _launch("ln", lib.mixdq_ln_quant_i8_dynamic,
                (x2.data_ptr(), x2.stride(0) if M > 1 else C, M, C, weight.data_ptr(),
                 bias.data_ptr(), float(eps), None, y.data_ptr(), None, None, ws.data_ptr()), x2,
                keep=(x2, weight, bias, y, ws), algo_bytes=4 * M * C)
    return y


def groupnorm_fp16(x: torch.Tensor, num_groups: int, weight: torch.Tensor, bias: torch.Tensor,
                   eps: float, silu: bool) -> torch.Tensor:
    """GroupNorm [+ SiLU], fp16 channels_last in / out, by the statistics + apply kernels of the
    fused path (two launches and a memset of the statistics accumulators)."""
    _check(x.dtype == torch.float16 and x.dim() == 4, "groupnorm_fp16 expects fp16 4-D")
    n, c, h, w = x.shape
    if _nhwc_pitch(x) != c:
        x = x.contiguous(memory_format=torch.channels_last)
    y = torch.empty((n, c, h, w), dtype=torch.float16, device=x.device,
                    memory_format=torch.channels_last)
    lib = _lib.load()
    with _DeviceGuard(x):
        ws = _dynamic_workspace(x.device)
        _launch("gn", lib.mixdq_gn_quant_i8_dynamic,
                (x.data_ptr(), c, n, h * w, c, num_groups, weight.data_ptr(), bias.data_ptr(),
                 float(eps), 1 if silu else 0, None, y.data_ptr(), None, None, ws.data_ptr()), x,
                keep=(x, weight, bias, y, ws), kernels=2, algo_bytes=4 * x.numel())
    return y


def qlinear_geglu_fp16(input_int8, weight_il, weight_scale_il, input_scale, input_zero_point,
                       weight_sum_il, bias_il=None) -> torch.Tensor:
    """ff.net.0.proj with the GEGLU in the GEMM epilogue -> fp16 [..., inner] (rows of the weight
    interleaved by `geglu_interleave_index`); the first half of qlinear_geglu_quantize_dynamic."""
    w4 = _is_w4(weight_il)
    N2, K = weight_il.shape[0], weight_il.shape[1] * (2 if w4 else 1)
    I = N2 // 2
    a = input_int8 if input_int8.is_contiguous() else input_int8.contiguous()
    M = a.numel() // K
    y = torch.empty((*input_int8.shape[:-1], I), dtype=torch.float16, device=a.device)
    lib = _lib.load()
    with _DeviceGuard(a):
        ws = _dynamic_workspace(a.device)
        _launch("gemm_geglu_w4" if w4 else "gemm_geglu",
                lib.mixdq_gemm_w4a8_geglu_f16_dyn if w4 else lib.mixdq_gemm_w8a8_geglu_f16_dyn,
                (a.data_ptr(), K, weight_il.data_ptr(), weight_scale_il.data_ptr(),
                 weight_sum_il.data_ptr(), input_scale.data_ptr(), input_zero_point.data_ptr(),
                 _ptr(bias_il), y.data_ptr(), I, M, N2, K, ws.data_ptr()), a,
                keep=(a, weight_il, weight_scale_il, weight_sum_il, input_scale,
                      input_zero_point, bias_il, y, ws),
                algo_bytes=M * K + N2 * K // (2 if w4 else 1) + 2 * M * I + 10 * N2,
                algo_ops=2 * M * N2 * K)
    return y


def layernorm_quantize_static(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                              eps: float, scale_inv: torch.Tensor, zero_point: torch.Tensor
                              ) -> torch.Tensor:
    """LayerNorm + the reference's static quantiser in ONE kernel: int8 [..., C], bit-identical to
    quantize_per_tensor_to_int8(layernorm_fp16(x, ...), scale_inv, zero_point)."""
    _check(x.dtype == torch.float16 and weight.dtype == torch.float16
           and bias.dtype == torch.float16, "layernorm_quantize_static expects fp16 tensors")
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    lib = _lib.load()
    with _DeviceGuard(x2):
        ws = _dynamic_workspace(x2.device)
        _launch("ln_quant", lib.mixdq_ln_quant_i8_static,
                (x2.data_ptr(), x2.stride(0) if M > 1 else C, M, C, weight.data_ptr(),
                 bias.data_ptr(), float(eps), scale_inv.data_ptr(), zero_point.data_ptr(),
                 q.data_ptr(), ws.data_ptr()), x2,
                keep=(x2, weight, bias, scale_inv, zero_point, q, ws), algo_bytes=3 * M * C)
    return q


def groupnorm_quantize_static(x: torch.Tensor, num_groups: int, weight: torch.Tensor,
                              bias: torch.Tensor, eps: float, silu: bool, scale_inv: torch.Tensor,
                              zero_point: torch.Tensor) -> torch.Tensor:
    """GroupNorm [+ SiLU] + the reference's static quantiser (statistics kernel, then an apply
    kernel that emits the codes): int8 [N,C,H,W] channels_last, bit-identical to
    quantize_per_tensor_to_int8(groupnorm_fp16(x, ...), scale_inv, zero_point)."""
    _check(x.dtype == torch.float16 and x.dim() == 4, "groupnorm_quantize_static expects fp16 4-D")
    n, c, h, w = x.shape
    if _nhwc_pitch(x) != c:
        x = x.contiguous(memory_format=torch.channels_last)
    q = torch.empty((n, c, h, w), dtype=torch.int8, device=x.device,
                    memory_format=torch.channels_last)
    lib = _lib.load()
    with _DeviceGuard(x):
        ws = _dynamic_workspace(x.device)
        _launch("gn_quant", lib.mixdq_gn_quant_i8_static,
                (x.data_ptr(), c, n, h * w, c, num_groups, weight.data_ptr(), bias.data_ptr(),
                 float(eps), 1 if silu else 0, scale_inv.data_ptr(), zero_point.data_ptr(),
                 q.data_ptr(), ws.data_ptr()), x,
                keep=(x, weight, bias, scale_inv, zero_point, q, ws), kernels=2,
                algo_bytes=5 * x.numel())
    return q


def qlinear_geglu_quantize_static(input_int8, weight_il, weight_scale_il, input_scale,
                                  input_zero_point, weight_sum_il, bias_il, scale_inv: torch.Tensor,
                                  zero_point: torch.Tensor) -> torch.Tensor:
    """ff.net.0.proj with GEGLU AND the static quantisation of its result (for ff.net.2) in the
    GEMM epilogue: one kernel, int8 [..., inner]; bit-identical to
    quantize_per_tensor_to_int8(qlinear_geglu_fp16(...), scale_inv, zero_point)."""
    w4 = _is_w4(weight_il)
    N2, K = weight_il.shape[0], weight_il.shape[1] * (2 if w4 else 1)
    I = N2 // 2
    a = input_int8 if input_int8.is_contiguous() else input_int8.contiguous()
    M = a.numel() // K
    q = torch.empty((*input_int8.shape[:-1], I), dtype=torch.int8, device=a.device)
    lib = _lib.load()
    with _DeviceGuard(a):
        ws = _dynamic_workspace(a.device)
        _launch("gemm_geglu_w4" if w4 else "gemm_geglu", lib.mixdq_gemm_geglu_i8_static,
                (a.data_ptr(), K, weight_il.data_ptr(), 4 if w4 else 8,
                 weight_scale_il.data_ptr(), weight_sum_il.data_ptr(), input_scale.data_ptr(),
                 input_zero_point.data_ptr(), _ptr(bias_il), scale_inv.data_ptr(),
                 zero_point.data_ptr(), q.data_ptr(), I, M, N2, K, ws.data_ptr()), a,
                keep=(a, weight_il, weight_scale_il, weight_sum_il, input_scale,
                      input_zero_point, bias_il, scale_inv, zero_point, q, ws),
                algo_bytes=M * K + N2 * K // (2 if w4 else 1) + M * I + 10 * N2,
                algo_ops=2 * M * N2 * K)
    return q


def tensor_minmax(x: torch.Tensor) -> torch.Tensor:
    """fp32 [2] = (min(0, min x), max(0, max x)) of a dense fp16 tensor (numel % 8 == 0), by the
    min/max pass of the dynamic quantiser + a one-CTA reduction of its partials (PTQ calibration,
    mixdq_b200.ptq)."""
    _check(x.dtype == torch.float16 and x.device.type == "cuda", "tensor_minmax expects CUDA fp16")
    xd = x if _is_dense(x) else x.contiguous()
    _check(xd.numel() % 8 == 0 and xd.numel() > 0, "tensor_minmax needs numel % 8 == 0")
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    with _DeviceGuard(xd):
        ws = _dynamic_workspace(xd.device)
        _launch("minmax", lib.mixdq_minmax_f16,
                (xd.data_ptr(), xd.numel(), out.data_ptr(), ws.data_ptr()), xd, kernels=2,
                keep=(xd, out, ws), algo_bytes=2 * xd.numel())
    return out


def quantize_rows_dynamic(x2: torch.Tensor):
    """A10 on a 2-D row-pitched fp16 view [M, cols] (stride(1) == 1) -> dense int8 [M, cols]."""
    _check(x2.dtype == torch.float16 and x2.dim() == 2 and x2.stride(1) == 1,
           "quantize_rows_dynamic expects a 2-D fp16 view with unit inner stride")
    M, cols = x2.shape
    q = torch.empty((M, cols), dtype=torch.int8, device=x2.device)
    qp, sc, zp = _qp_pair(x2.device)
    lib = _lib.load()
    with _DeviceGuard(x2):
        ws = _dynamic_workspace(x2.device)
        _launch("quant_dyn", lib.mixdq_quant_i8_dynamic_rows,
                (x2.data_ptr(), x2.stride(0) if M > 1 else cols, M, cols, q.data_ptr(),
                 qp.data_ptr(), qp.data_ptr() + 4, ws.data_ptr()), x2, keep=(x2, q, qp, ws),
                kernels=_dyn_kernels(M * cols), algo_bytes=3 * M * cols)
    return q, sc, zp


def quantize_nhwc_slice_dynamic(x: torch.Tensor, c0: int, c1: int):
    """A10 on channels [c0, c1) of a channels_last fp16 tensor [N,C,H,W] -> dense int8
    channels_last [N, c1-c0, H, W] (no slicing copy)."""
    n, c, h, w = x.shape
    _check(_nhwc_pitch(x) == c, "quantize_nhwc_slice_dynamic expects a dense channels_last tensor")
    rows = x.permute(0, 2, 3, 1).reshape(n * h * w, c)[:, c0:c1]
    q, sc, zp = quantize_rows_dynamic(rows)
    return q.view(n, h, w, c1 - c0).permute(0, 3, 1, 2), sc, zp
