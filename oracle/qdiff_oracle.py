"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference's algorithm for the quantized-UNet hot path. Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
this module, and only as the checker / reported baseline — never the product path (the product
path is the CUDA library behind include/mixdq_b200.h and raises if that library is missing).

PARITY PINNED: the functions below are checked against outputs of the reference itself, generated
in the build container by importing the reference's Python (`oracle/make_golden.py`, fixtures under
`tests/golden/`): the qdiff fake-quant leaf modules (`QuantLayer` / `BaseQuantizer`), the reference
`QuantizedLinear.from_float` / `QuantizedConv2d.from_float` run against the shipped
`kernels/output/new_ckpt.pth`, and `torch.quantize_per_tensor` / `torch.quantize_per_channel`
(the third-party arithmetic the reference calls, PyTorch >= 2.2.1 per kernels/requirements.txt:3).
What is NOT pinned (no reference artefact exists for it here): whole-UNet outputs — `diffusers`
is absent and the reference holds no numeric whole-model fixture.

Each function cites the reference file:line it follows (paths relative to the reference root).
Arithmetic is fp32 on CPU (torch CPU tensors == numpy semantics: IEEE fp32, one rounding per
elementwise op, no FMA contraction), integer work exact in int64.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------
# qdiff quantizers (quant_utils/qdiff/quantizer/base_quantizer.py)
# ---------------------------------------------------------------------------------------------
EPS_DELTA = 1.0e-6  # base_quantizer.py:178


def act_qparams_minmax(x: torch.Tensor, n_bits: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
    """Asymmetric per-tensor min-max init. base_quantizer.py:131-190 with channel_wise=False,
    sym=False, scale_method='min_max'; first call of a running-stat quantizer (:160-165) uses
    the batch min/max directly. Returns (delta, zero_point) fp32 0-d tensors."""
    x = x.detach().float().reshape(-1)
    x_min = torch.clamp(x.min(), max=0.0)          # :155-156  x_min[x_min>0] = 0
    x_max = torch.clamp(x.max(), min=0.0)          # :157-158
    n_levels = 2 ** n_bits                         # :142 (sym False)
    delta = (x_max - x_min) / (n_levels - 1)       # :177
    if delta < EPS_DELTA:                          # :179-181
        delta = torch.full_like(delta, EPS_DELTA)
    zero_point = torch.round(-x_min / delta)       # :187
    return delta, zero_point


def act_fake_quant(x: torch.Tensor, delta: torch.Tensor, zero_point: torch.Tensor,
                   n_bits: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
    """BaseQuantizer.forward for an asymmetric quantizer. base_quantizer.py:119-128.
    Returns (codes in [0, 2^b-1] as fp32, dequantised x_hat)."""
    n_levels = 2 ** n_bits
    x_int = torch.round(x / delta) + zero_point            # :122 (round_ste == round in fwd)
    x_quant = torch.clamp(x_int, 0, n_levels - 1)           # :127
    x_dequant = (x_quant - zero_point) * delta              # :128
    return x_quant, x_dequant


def weight_qparams_minmax(w: torch.Tensor, n_bits: int = 8) -> torch.Tensor:
    """Symmetric per-output-channel min-max init. base_quantizer.py:147-185 with
    channel_wise=True, sym=True: delta_c = max|w_c| / (2^(b-1)-1); zero_point = 0.
    Returns delta of shape [Cout]."""
    w = w.detach().float()
    wf = w.reshape(w.shape[0], -1)
    x_min = torch.clamp(wf.min(dim=-1)[0], max=0.0)
    x_max = torch.clamp(wf.max(dim=-1)[0], min=0.0)
    n_levels = 2 ** (n_bits - 1) - 1                        # :142 (sym True)
    x_absmax = torch.maximum(x_min.abs(), x_max.abs())      # :174
    delta = x_absmax / n_levels                             # :176
    if delta.min() < EPS_DELTA:                             # :179-181 (fills ALL channels)
        delta = torch.full_like(delta, EPS_DELTA)
    return delta


def weight_fake_quant(w: torch.Tensor, delta: torch.Tensor, n_bits: int = 8
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """BaseQuantizer.forward for a symmetric quantizer. base_quantizer.py:119-128:
    clamp(round(w/delta), -n_levels-1, n_levels) with n_levels = 2^(b-1)-1, i.e. [-2^(b-1), 2^(b-1)-1].
    Returns (integer codes as fp32, dequantised w_hat)."""
    n_levels = 2 ** (n_bits - 1) - 1
    shape = [-1] + [1] * (w.dim() - 1)
    d = delta.reshape(shape)
    x_int = torch.round(w.float() / d)
    x_quant = torch.clamp(x_int, -n_levels - 1, n_levels)   # :124-125
    return x_quant, x_quant * d


def fake_quant_layer(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                     w_bits: int = 8, a_bits: int = 8, split: int = 0,
                     stride: int = 1, padding: int = 0,
                     act_params=None, act_params_0=None) -> torch.Tensor:
    """QuantLayer.forward with weight_quant = act_quant = True. quant_layer.py:63-103.
    Linear if weight.dim()==2 else conv2d. `split` > 0: the two input-channel ranges are
    quantised independently (activation AND weight) and concatenated (:74-88).
    act_params: optional pre-computed (delta, zero_point) — static-scale mode; else dynamic."""
    def qa(t, params):
        d, z = params if params is not None else act_qparams_minmax(t, a_bits)
        return act_fake_quant(t.float(), d, z, a_bits)[1]

    def qw(t):
        return weight_fake_quant(t, weight_qparams_minmax(t, w_bits), w_bits)[1]

    if split:
        x_hat = torch.cat([qa(x[:, :split], act_params), qa(x[:, split:], act_params_0)], dim=1)
        w_hat = torch.cat([qw(weight[:, :split]), qw(weight[:, split:])], dim=1)
    else:
        x_hat = qa(x, act_params)
        w_hat = qw(weight)
    b = None if bias is None else bias.float()
    if weight.dim() == 2:
        return F.linear(x_hat, w_hat, b)
    return F.conv2d(x_hat, w_hat, b, stride=stride, padding=padding)


# ---------------------------------------------------------------------------------------------
# kernel-path arithmetic (kernels/mixdq_extension)
# ---------------------------------------------------------------------------------------------
def quantize_static_kernel(x: torch.Tensor, scale_inv: float, zero_point: float) -> torch.Tensor:
    """The reference CUDA quantize kernel: int8(clamp(lrintf(x*scale_inv + zp), -128, 127)) with the
    multiply-add contracted to one FMA by nvcc. csrc/quant_dequant/quantize_kernel.cu:20-24.
    x is fp16 (11-bit significand), scale_inv fp32 (24-bit): the product and the sum are exact in
    fp64, so rounding the fp64 result once to fp32 reproduces fmaf bit for bit."""
    xe = x.detach().to(torch.float64)
    exact = xe * float(np.float32(scale_inv)) + float(np.float32(zero_point))
    f = exact.to(torch.float32)
    return torch.clamp(torch.round(f), -128, 127).to(torch.int8)


def quantize_dynamic_kernel(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """qdiff activation quantisation emitted in the kernel format: int8 code = q - 128, zero point
    z - 128 (nn/utils.py:428). Returns (int8 codes, delta, zp_shifted)."""
    delta, z = act_qparams_minmax(x, 8)
    q, _ = act_fake_quant(x.float(), delta, z, 8)
    return (q - 128).to(torch.int8), delta, z - 128


def quantize_weight_per_channel(w: torch.Tensor, scales: torch.Tensor, n_bits: int = 8
                                ) -> torch.Tensor:
    """torch.quantize_per_channel(w.float(), scales, zp=0, axis=0, qint8).int_repr() as called by
    nn/Linear.py:116-121 / nn/Conv2d.py:157-162: q = clamp(nearbyint(w * (1/s)), -128, 127)
    (PyTorch's CPU kernel multiplies by the fp32 reciprocal). For n_bits < 8 the clamp range is the
    signed n-bit range (W4 codes stored one per int8 before packing)."""
    lo, hi = -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1
    shape = [-1] + [1] * (w.dim() - 1)
    inv = (1.0 / scales.float()).reshape(shape)
    return torch.clamp(torch.round(w.float() * inv), lo, hi).to(torch.int8)


def int_accumulate_linear(a_int8: torch.Tensor, w_int8: torch.Tensor) -> torch.Tensor:
    """Exact INT32 accumulators of A[M,K] @ W[N,K]^T (int64 arithmetic)."""
    return (a_int8.to(torch.int64) @ w_int8.to(torch.int64).t())


def int_accumulate_conv(x_int8: torch.Tensor, w_int8: torch.Tensor, stride: int, padding: int
                        ) -> torch.Tensor:
    """Exact INT32 accumulators of the cross-correlation; zero padding in the INTEGER domain
    (what the reference kernel does; the zero-point correction only covers in-bounds taps).
    fp64 conv is exact here: |acc| <= K*128*128 << 2^53."""
    return F.conv2d(x_int8.double(), w_int8.double(), stride=stride, padding=padding).to(torch.int64)


def zero_point_propagate(wsum_krs: torch.Tensor, zp: float, H: int, W: int, stride: int,
                         padding: int) -> torch.Tensor:
    """activation_zero_point_propagate: bias0[p,q,k] = float(sum_{(r,s) in bounds} wsum[k,r,s]) * zp.
    csrc/qconv2d/conv_act_zero_point_propagate.cu:10-51. Returns fp32 [K, P, Q]."""
    K, R, S = wsum_krs.shape
    P = (H + 2 * padding - R) // stride + 1
    Q = (W + 2 * padding - S) // stride + 1
    ones = torch.ones(1, 1, H, W, dtype=torch.float64)
    acc = F.conv2d(ones, wsum_krs.double().reshape(K, 1, R, S), stride=stride, padding=padding)[0]
    assert acc.shape == (K, P, Q)
    return acc.float() * torch.tensor(zp, dtype=torch.float32)


def kernel_epilogue(acc: torch.Tensor, bias0: torch.Tensor, scale: torch.Tensor,
                    bias: Optional[torch.Tensor]) -> torch.Tensor:
    """D = half((float(acc) - bias0) * scale [+ float(bias)]), three separately rounded fp32 ops
    (CUTLASS EVT: minus, multiplies, plus; cutlassGemm_withBias_optimalAlignment.cu:39-95).
    bias0/scale/bias must broadcast against acc."""
    f = (acc.to(torch.float32) - bias0.float()) * scale.float()
    if bias is not None:
        f = f + bias.float()
    return f.to(torch.float16)


def qlinear_kernel(a_int8, w_int8, bias0, scale, bias=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """qlinear_w8_a8_ohalf (csrc/qlinear/qlinear.cc:14-137). Returns (fp16 out, int accumulators)."""
    K = w_int8.shape[1]
    acc = int_accumulate_linear(a_int8.reshape(-1, K), w_int8)
    out = kernel_epilogue(acc, bias0[None, :], scale[None, :], None if bias is None else bias[None, :])
    return out.reshape(*a_int8.shape[:-1], w_int8.shape[0]), acc


def qconv2d_kernel(x_int8, w_int8, scale, wsum_krs, bias0_k, zp, bias, stride, padding
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """qconv2d_w8_a8_ohalf (csrc/qconv2d/qconv2d.cc:28-206): per-pixel bias0 when padding > 0,
    per-channel bias0 otherwise. Returns (fp16 out [N,K,P,Q], int accumulators [N,K,P,Q])."""
    acc = int_accumulate_conv(x_int8, w_int8, stride, padding)
    if padding > 0:
        K, _, R, S = w_int8.shape
        b0 = zero_point_propagate(wsum_krs.reshape(K, R, S), float(zp), x_int8.shape[2],
                                  x_int8.shape[3], stride, padding)[None]
    else:
        b0 = bias0_k.float()[None, :, None, None]
    out = kernel_epilogue(acc, b0, scale[None, :, None, None],
                          None if bias is None else bias[None, :, None, None])
    return out, acc


def split_shortcut_kernel(out_a: torch.Tensor, out_b: torch.Tensor) -> torch.Tensor:
    """`output + output_0` of two fp16 conv results (nn/Conv2d.py:346): torch adds halves in fp32
    and rounds to fp16."""
    return (out_a.float() + out_b.float()).to(torch.float16)


# ---- from_float (nn/Linear.py:57-140, nn/Conv2d.py:91-244, nn/utils.py:412-458) -------------
def ckpt_qparams(ckpt: Dict, module_name: str, kind: str, n_bit: int, suffix: str = ""
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """get_quant_para (nn/utils.py:412-458): bit_idx = log2(bits)-1; activation zero point is
    shifted by -128 (uint8 -> int8). kind in {'weight','act'}; suffix '_0' for the split twin."""
    bit_idx = int(math.log2(n_bit) - 1)
    key = module_name + ('.weight_quantizer' if kind == 'weight' else '.act_quantizer') + suffix
    scales = ckpt[key]['delta_list'][bit_idx]
    zp = ckpt[key]['zero_point_list'][bit_idx]
    if kind == 'act':
        zp = zp - 128
    return scales.float(), zp.float()


def from_float_buffers(weight: torch.Tensor, w_scales: torch.Tensor, a_scale: torch.Tensor,
                       a_zp: torch.Tensor, padding: int = 0) -> Dict[str, torch.Tensor]:
    """The buffer set QuantizedLinear / QuantizedConv2d.from_float derive for one (half-)layer:
    weight_int, weight_sum_by_input_channels | bias0, scale, act_scales_inv."""
    w_int = quantize_weight_per_channel(weight, w_scales)
    out = {"weight_int": w_int, "scale": w_scales.float() * a_scale.float(),
           "act_scales_inv": 1 / a_scale.float()}
    if weight.dim() == 2:
        wsum = w_int.float().sum(dim=1)                               # Linear.py:125
        out["weight_sum_by_input_channels"] = wsum
        out["bias0"] = wsum * a_zp.float()                            # Linear.py:131-132
    elif padding == 0:
        out["bias0"] = w_int.float().sum(dim=[1, 2, 3]) * a_zp.float()   # Conv2d.py:166-170
    else:
        out["weight_sum_by_input_channels"] = w_int.float().sum(dim=1, keepdim=True)  # :173-176
    return out


# ---- W4 packing (north star; nibble order of nn/utils.py:26-28: even index -> high nibble) ----
def pack_int4(codes: torch.Tensor) -> torch.Tensor:
    """int8 codes in [-8,7], last dim even -> uint8 [..., K/2]; even k in the HIGH nibble,
    two's-complement nibbles."""
    c = codes.to(torch.int16)
    hi = (c[..., 0::2] & 0xF) << 4
    lo = c[..., 1::2] & 0xF
    return (hi | lo).to(torch.uint8)


def unpack_int4(packed: torch.Tensor) -> torch.Tensor:
    p = packed.to(torch.int16)
    hi = (p >> 4) & 0xF
    lo = p & 0xF
    both = torch.stack([hi, lo], dim=-1).reshape(*packed.shape[:-1], packed.shape[-1] * 2)
    return torch.where(both >= 8, both - 16, both).to(torch.int8)


# ---------------------------------------------------------------------------------------------
# Stock fp16 ops AROUND the quantized leaves (the model graph the reference quantizes comes from
# diffusers / torch.nn, reference kernels/mixdq.py:4,35-41). On the GPU the model runs in fp16:
# each op computes in fp32 and materialises an fp16 tensor. The fused producer kernels
# (include/mixdq_b200.h "Producer-side fusion") and the fused epilogue tails must reproduce these
# sequences; the restatements below are fp32 CPU math with the fp16 roundings at the same points.
# ---------------------------------------------------------------------------------------------
def layernorm_fp16(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float
                   ) -> torch.Tensor:
    """torch.nn.LayerNorm on an fp16 tensor: fp32 statistics, one rounding to fp16."""
    return F.layer_norm(x.float(), (x.shape[-1],), weight.float(), bias.float(), eps).half()


def geglu_fp16(hg: torch.Tensor) -> torch.Tensor:
    """diffusers GEGLU on fp16: h, gate = chunk(2); gelu(gate) (exact erf) -> fp16; h * gelu -> fp16."""
    h, gate = hg.chunk(2, dim=-1)
    g = F.gelu(gate.float()).half()
    return (h.float() * g.float()).half()


def groupnorm_fp16(x: torch.Tensor, groups: int, weight: torch.Tensor, bias: torch.Tensor,
                   eps: float, silu: bool) -> torch.Tensor:
    """torch.nn.GroupNorm on fp16 (fp32 statistics, fp16 result) [+ torch.nn.SiLU on that fp16
    tensor (fp32 math, fp16 result)]."""
    y = F.group_norm(x.float(), groups, weight.float(), bias.float(), eps).half()
    if silu:
        y = F.silu(y.float()).half()
    return y


def add_fp16(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """`a + b` on two fp16 tensors: fp32 add, one rounding."""
    return (a.float() + b.float()).half()
