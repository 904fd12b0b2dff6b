"""QuantizedConv2d — drop-in for the reference's kernels/mixdq_extension/nn/Conv2d.py:16-347.

Same constructor / `from_float(float_mod, split=0, ckpt=None)` / buffer names (incl. the `*_0`
twins of split shortcuts) / `_get_name()`. The forward differs in how many kernels it launches:
  reference : quantize -> int8 NCHW->NHWC copy -> zero-point-propagate -> conv   (+ 2nd set + add
              for split shortcuts: nn/Conv2d.py:312-347)
  here      : fused quantize+layout kernel -> implicit-GEMM conv with the border correction in
              its epilogue; a split shortcut is two quantize launches + ONE dual-accumulator
              kernel.
Weights are stored KRSC (channels_last) so no per-call layout conversion happens.
"""
from __future__ import annotations

import logging

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.ao.quantization import QConfig

from .. import ops
from .utils import (create_qparams_from_dtype, minmax_weight_scales, pack_int4, quantize_weight,
                    unpack_int4, QParam)

__all__ = ["QuantizedConv2d"]

_Q8 = (torch.qint8, torch.quint8)
_Q4 = (torch.quint4x2,)


def _w_ok(q):
    return (q is not None and q.dtype in _Q8 + _Q4 and q.qscheme == torch.per_channel_affine
            and bool(torch.all(q.zero_points == 0.0).item()))


def _a_ok(q):
    return q is not None and q.dtype in _Q8 and q.qscheme == torch.per_tensor_affine


class QuantizedConv2d(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride, padding,
                 dilation, groups=1, bias=True, device=None, w_qparams=None, w_qparams_0=None,
                 a_qparams=None, a_qparams_0=None, module_name=None, split=0,
                 dynamic: bool = False) -> None:
        super().__init__()
        self.module_name = module_name
        self.split = split
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.device = device
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.groups = groups
        self.dynamic = bool(dynamic)

        geometry_ok = (len(set(stride)) == 1 and len(set(padding)) == 1
                       and len(set(dilation)) == 1 and dilation[0] == 1 and groups == 1)
        acts_ok = self.dynamic or (_a_ok(a_qparams) and (split == 0 or _a_ok(a_qparams_0)))
        self.valid_for_acceleration = (_w_ok(w_qparams) and (split == 0 or _w_ok(w_qparams_0))
                                       and acts_ok and geometry_ok)
        if self.valid_for_acceleration and (in_channels % 4 != 0 or out_channels % 4 != 0
                                            or (split and split % 4 != 0)):
            logging.warning("Linear layer with in_features = "
                            f"{in_channels} and out_features = {out_channels} cannot use "
                            "quantized kernel due to misalignment. Falling back to FP kernels")
            self.valid_for_acceleration = False
        self.w_bits = 8
        if self.valid_for_acceleration:
            self.w_bits = 4 if w_qparams.dtype in _Q4 else 8
            self._register_qparams("", w_qparams, a_qparams, device)
            if split != 0:
                self._register_qparams("_0", w_qparams_0, a_qparams_0, device)

    def _register_qparams(self, sfx, w_q, a_q, device):
        self.register_buffer("weight_scales" + sfx, w_q.scales.to(device).float())
        self.register_buffer("weight_zero_points" + sfx, w_q.zero_points.to(device).float())
        if not self.dynamic:
            self.register_buffer("act_scales" + sfx, a_q.scales.to(device).float())
            self.register_buffer("act_zero_points" + sfx, a_q.zero_points.to(device).float())
            self.register_buffer("act_scales_inv" + sfx, 1 / getattr(self, "act_scales" + sfx))

    # ------------------------------------------------------------------------------------
    @classmethod
    def from_float(cls, float_mod, split=0, ckpt=None):
        assert hasattr(float_mod, "qconfig") and isinstance(float_mod.qconfig, QConfig)
        w_dtype = float_mod.qconfig.weight().dtype
        act_dtype = float_mod.qconfig.activation().dtype
        weight = float_mod.weight.detach()
        device = weight.device
        n_out = weight.shape[0]
        w_bit = getattr(float_mod, "w_bit", 8)
        w_bit_eff = 4 if w_bit == 2 else w_bit
        dynamic = ckpt is None
        w_q = w_q0 = a_q = a_q0 = None
        use_dynamic = False
        if dynamic:
            if w_dtype in _Q8 + _Q4:
                def mk(w):
                    s = minmax_weight_scales(w, w_bit_eff)
                    return QParam(qscheme=torch.per_channel_affine, dtype=w_dtype, scales=s,
                                  zero_points=torch.zeros_like(s), axis=0)
                if split:
                    w_q, w_q0 = mk(weight[:, :split]), mk(weight[:, split:])
                else:
                    w_q = mk(weight)
            use_dynamic = hasattr(float_mod, "a_bit") and act_dtype in _Q8
        else:
            pair = create_qparams_from_dtype(dtype=w_dtype, device=device, is_channel_wise=True,
                                             num_kernels=n_out, ckpt=ckpt,
                                             module_name=float_mod.module_name,
                                             quant_type="weight", bit_width=w_bit_eff, split=split)
            if pair is not None:
                w_q, w_q0 = pair
            if hasattr(float_mod, "a_bit"):
                pair = create_qparams_from_dtype(dtype=act_dtype, device=device,
                                                 is_channel_wise=False, num_kernels=n_out,
                                                 ckpt=ckpt, module_name=float_mod.module_name,
                                                 quant_type="act", bit_width=float_mod.a_bit,
                                                 split=split)
                if pair is not None:
                    a_q, a_q0 = pair

        new_mod = cls(float_mod.in_channels, float_mod.out_channels, float_mod.kernel_size,
                      float_mod.stride, float_mod.padding, float_mod.dilation, float_mod.groups,
                      float_mod.bias is not None, device=device, w_qparams=w_q,
                      w_qparams_0=w_q0, a_qparams=a_q, a_qparams_0=a_q0,
                      module_name=float_mod.module_name, split=split, dynamic=use_dynamic)

        if new_mod.valid_for_acceleration:
            pad0 = float_mod.padding[0] == 0

            def quantise(w, sfx):
                scales = getattr(new_mod, "weight_scales" + sfx)
                if new_mod.w_bits == 8 and not dynamic:
                    w_int = torch.quantize_per_channel(
                        w.float(), scales, getattr(new_mod, "weight_zero_points" + sfx), axis=0,
                        dtype=w_dtype).int_repr()
                else:
                    w_int = quantize_weight(w, scales, new_mod.w_bits, exact_division=dynamic)
                # KRSC storage: what the implicit-GEMM kernel consumes (qconv2d.cc:94-95).
                # 4-bit layers the tcgen05 kernel can take are stored PACKED (KRS(C/2), even c in
                # the high nibble) as `weight_int4`; the rest keep one code per int8.
                w_int = w_int.contiguous(memory_format=torch.channels_last)
                if new_mod.w_bits == 4 and new_mod._w4_packable(w.shape[1]):
                    new_mod.register_buffer("weight_int4" + sfx, pack_int4(w_int, dim=1))
                else:
                    new_mod.register_buffer("weight_int" + sfx, w_int)
                if pad0:
                    wsum = w_int.float().sum(dim=[1, 2, 3])
                    if new_mod.dynamic:
                        new_mod.register_buffer("weight_sum_per_output_channel" + sfx, wsum)
                    else:
                        new_mod.register_buffer(
                            "bias0" + sfx, wsum * getattr(new_mod, "act_zero_points" + sfx))
                        # the fused block forwards fold the activation scalars inside the kernel
                        # (bias0[n] = wsum[n] * zp): not part of the reference's state_dict
                        new_mod.register_buffer("weight_sum_per_output_channel" + sfx, wsum,
                                                persistent=False)
                    setattr(new_mod, "weight_sum_by_input_channels" + sfx, None)
                else:
                    new_mod.register_buffer("weight_sum_by_input_channels" + sfx,
                                            w_int.float().sum(dim=1, keepdim=True))
                    setattr(new_mod, "bias0" + sfx, None)
                if not new_mod.dynamic:
                    new_mod.register_buffer(
                        "scale" + sfx, scales * getattr(new_mod, "act_scales" + sfx))

            if split == 0:
                quantise(weight, "")
            else:
                quantise(weight[:, :split, ...], "")
                quantise(weight[:, split:, ...], "_0")
        else:
            new_mod.register_buffer("weight", weight)
        if float_mod.bias is not None:
            new_mod.register_buffer("bias", float_mod.bias.detach())
        else:
            new_mod.bias = None
        return new_mod

    def _get_name(self):
        if self.valid_for_acceleration:
            return "QuantizedConv2dW8A8" if self.w_bits == 8 else "QuantizedConv2dW4A8"
        return "QuantizedConv2dFPFallback"

    def _w4_packable(self, c_in: int) -> bool:
        """geometry / alignment the packed-weight tcgen05 convolution takes (capi.cu conv_common)"""
        stride, pad, k = self.stride[0], self.padding[0], self.kernel_size[0]
        return (c_in % 32 == 0 and self.out_channels % 8 == 0 and stride in (1, 2)
                and self.kernel_size[0] == self.kernel_size[1]
                and (pad == 0 or (pad == 1 and k == 3)))

    def _weight(self, sfx: str) -> torch.Tensor:
        """the layer's weight operand: packed uint8 (`weight_int4*`) or int8 codes"""
        w = getattr(self, "weight_int4" + sfx, None)
        return w if w is not None else getattr(self, "weight_int" + sfx)

    # ------------------------------------------------------------------------------------
    def forward_fallback(self, x: torch.Tensor):
        """Non-fp16 input: float conv on the DEQUANTISED weight — warns once per module."""
        if not getattr(self, "_warned_fallback", False):
            self._warned_fallback = True
            logging.warning(f"{self._get_name()} {self.module_name}: input dtype {x.dtype} is not "
                            "fp16; running F.conv2d on the dequantised weight (no int8 kernel)")

        def deq(sfx):
            w = self._weight(sfx)
            if w.dtype == torch.uint8:
                w = unpack_int4(w, dim=1)
            w = w.float() * getattr(self, "weight_scales" + sfx)[:, None, None, None]
            return w.to(x.dtype)
        bias = self.bias.to(x.dtype) if self.bias is not None else None
        args = (self.stride, self.padding, self.dilation, self.groups)
        if self.split == 0:
            return F.conv2d(x, deq(""), bias, *args)
        return (F.conv2d(x[:, :self.split], deq(""), bias, *args)
                + F.conv2d(x[:, self.split:], deq("_0"), None, *args))

    def _half(self, x, sfx, c0, c1, bias):
        """quantise channels [c0,c1) of x and run one conv; returns fp16 channels_last."""
        stride, pad = self.stride[0], self.padding[0]
        w_int = self._weight(sfx)
        if self.dynamic:
            xs = x if (c0 == 0 and c1 == x.shape[1]) else x[:, c0:c1]
            ksel = c1 - c0
            tc_ok = (ksel % 16 == 0 and self.out_channels % 8 == 0 and stride in (1, 2)
                     and (pad == 0 or (pad == 1 and self.kernel_size[0] == 3)))
            if tc_ok:
                # activation scalars are folded inside the kernel epilogue; channel slices of an
                # NHWC tensor are quantised in place (no slicing copy)
                if xs is x:
                    x_int, a_scale, a_zp = ops.quantize_per_tensor_dynamic(
                        x.contiguous(memory_format=torch.channels_last))
                else:
                    x_int, a_scale, a_zp = ops.quantize_nhwc_slice_dynamic(
                        x.contiguous(memory_format=torch.channels_last), c0, c1)
                return ops.qconv2d_dynamic_fused(
                    x_int, w_int, getattr(self, "weight_scales" + sfx), a_scale, a_zp,
                    getattr(self, "weight_sum_by_input_channels" + sfx) if pad > 0 else None,
                    getattr(self, "weight_sum_per_output_channel" + sfx) if pad == 0 else None,
                    bias, stride, pad)
            x_int, a_scale, a_zp = ops.quantize_per_tensor_dynamic(
                xs.contiguous(memory_format=torch.channels_last))
            scale = getattr(self, "weight_scales" + sfx) * a_scale
            bias0 = None
            if pad == 0:
                bias0 = getattr(self, "weight_sum_per_output_channel" + sfx) * a_zp
            return ops.qconv2d_w8_a8_ohalf(
                x_int, w_int, getattr(self, "weight_scales" + sfx), a_scale, a_zp, scale,
                getattr(self, "weight_sum_by_input_channels" + sfx), bias0, bias, stride, pad)
        x_int = ops.quantize_to_nhwc(x, getattr(self, "act_scales_inv" + sfx),
                                     getattr(self, "act_zero_points" + sfx), c0, c1)
        return ops.qconv2d_w8_a8_ohalf(
            x_int, w_int, getattr(self, "weight_scales" + sfx), getattr(self, "act_scales" + sfx),
            getattr(self, "act_zero_points" + sfx), getattr(self, "scale" + sfx),
            getattr(self, "weight_sum_by_input_channels" + sfx), getattr(self, "bias0" + sfx),
            bias, stride, pad)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.valid_for_acceleration:
            return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation,
                            self.groups)
        if x.dtype != torch.float16:
            return self.forward_fallback(x)
        C = x.shape[1]
        if self.split == 0:
            return self._half(x, "", 0, C, self.bias)
        packed = getattr(self, "weight_int4", None) is not None   # W4: two convs + fp16 add
        fused = (not self.dynamic and not packed and self.kernel_size[0] == 1
                 and self.kernel_size[1] == 1 and self.padding[0] == 0 and self.stride[0] == 1)
        if fused:
            xa = ops.quantize_to_nhwc(x, self.act_scales_inv, self.act_zero_points, 0, self.split)
            xb = ops.quantize_to_nhwc(x, self.act_scales_inv_0, self.act_zero_points_0,
                                      self.split, C)
            return ops.qconv1x1_split_w8_a8_ohalf(xa, self.weight_int, self.scale, self.bias0,
                                                  xb, self.weight_int_0, self.scale_0,
                                                  self.bias0_0, self.bias)
        fused_dyn = (self.dynamic and not packed and self.kernel_size[0] == 1 and self.kernel_size[1] == 1
                     and self.padding[0] == 0 and self.stride[0] == 1 and self.split % 16 == 0
                     and (C - self.split) % 16 == 0 and self.out_channels % 8 == 0)
        if fused_dyn:
            xc = x.contiguous(memory_format=torch.channels_last)
            xa, sa, za = ops.quantize_nhwc_slice_dynamic(xc, 0, self.split)
            xb, sb, zb = ops.quantize_nhwc_slice_dynamic(xc, self.split, C)
            return ops.qconv1x1_split_dynamic_fused(
                xa, self.weight_int, self.weight_scales, self.weight_sum_per_output_channel, sa, za,
                xb, self.weight_int_0, self.weight_scales_0, self.weight_sum_per_output_channel_0,
                sb, zb, self.bias)
        out = self._half(x, "", 0, self.split, self.bias)
        out_0 = self._half(x, "_0", self.split, C, None)   # bias applied once (Conv2d.py:337-339)
        return out + out_0
