"""QuantizedLinear — drop-in for the reference's kernels/mixdq_extension/nn/Linear.py:17-194.

Same constructor / `from_float(float_mod, split=0, ckpt=None)` / buffer names / `_get_name()`
/ BOS special case, so `quantize.convert` and state_dicts are interchangeable. Differences:
  * the forward runs on the sm_100a kernels behind include/mixdq_b200.h (through mixdq_b200.ops);
  * 4-bit weights (`torch.quint4x2` qconfig, the reference's FP fallback, nn/Linear.py:28-36) run
    as true W4A8: `weight_int4` holds the packed codes (uint8 [N, K/2], even k in the high nibble,
    nn/utils.py:26-28), which the tcgen05 kernel unpacks on the fly;
  * `ckpt=None` selects dynamic mode: qdiff min-max weight scales computed here, activations
    quantised per call from their own min/max (reference base_quantizer.py:155-190);
  * 4-bit ACTIVATIONS (`a_bit: 4` in kernels/cfgs/act/act_7.xx.yaml — every such layer of the
    shipped configs is a Linear; the reference runs them in fp16, nn/Linear.py:28-36) are
    quantised to codes 0..15 kept one per int8, with the unshifted zero point, and run on the
    same int8 kernels (`a_bits = 4`, `_get_name()` -> "...A4").
"""
from __future__ import annotations

import logging

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.ao.quantization import QConfig

from .. import ops
from .utils import (create_qparams_from_dtype, minmax_weight_scales, pack_int4, quantize_weight,
                    QParam)

__all__ = ["QuantizedLinear"]

_W8 = (torch.qint8, torch.quint8)
_W4 = (torch.quint4x2,)


def _weight_kind(w_qparams):
    if w_qparams is None or w_qparams.qscheme != torch.per_channel_affine:
        return None
    if not bool(torch.all(w_qparams.zero_points == 0.0).item()):
        return None
    if w_qparams.dtype in _W8:
        return "w8"
    if w_qparams.dtype in _W4:
        return "w4"
    return None


def _act_ok(a_qparams):
    return (a_qparams is not None and a_qparams.dtype in _W8 + _W4
            and a_qparams.qscheme == torch.per_tensor_affine)


class QuantizedLinear(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None,
                 w_qparams=None, a_qparams=None, module_name=None, dynamic: bool = False,
                 a_bits: int = 8) -> None:
        super().__init__()
        assert a_bits in (4, 8)
        self.a_bits = a_bits
        self.module_name = module_name
        self.in_features = in_features
        self.out_features = out_features
        self.device = device
        self.dynamic = bool(dynamic)
        self.w_kind = _weight_kind(w_qparams)
        self.valid_for_acceleration = self.w_kind is not None and (self.dynamic or _act_ok(a_qparams))
        k_align = 32 if self.w_kind == "w4" else 4
        if self.valid_for_acceleration and (in_features % k_align != 0 or out_features % 4 != 0):
            logging.warning("Linear layer with in_features = "
                            f"{in_features} and out_features = {out_features} cannot use "
                            "quantized kernel due to misalignment. Falling back to FP kernels")
            self.valid_for_acceleration = False
        if self.valid_for_acceleration:
            self.register_buffer("weight_scales", w_qparams.scales.to(device).float())
            self.register_buffer("weight_zero_points", w_qparams.zero_points.to(device).float())
            if not self.dynamic:
                self.register_buffer("act_scales", a_qparams.scales.to(device).float())
                self.register_buffer("act_zero_points", a_qparams.zero_points.to(device).float())
                self.register_buffer("act_scales_inv", 1 / self.act_scales)

    # ------------------------------------------------------------------------------------
    @classmethod
    def from_float(cls, float_mod, split=0, ckpt=None):
        assert hasattr(float_mod, "qconfig") and isinstance(float_mod.qconfig, QConfig)
        w_dtype = float_mod.qconfig.weight().dtype
        act_dtype = float_mod.qconfig.activation().dtype
        device = float_mod.weight.device
        weight = float_mod.weight.detach()
        n_out = weight.shape[0]
        w_bit = getattr(float_mod, "w_bit", 8)
        w_bit_eff = 4 if w_bit == 2 else w_bit   # 2-bit layers are promoted to 4 bit
        dynamic = ckpt is None

        if dynamic:
            if w_dtype in _W8 + _W4:
                scales = minmax_weight_scales(weight, w_bit_eff)
                w_qparams = QParam(qscheme=torch.per_channel_affine, dtype=w_dtype, scales=scales,
                                   zero_points=torch.zeros_like(scales), axis=0)
            else:
                w_qparams = None
            a_qparams = None
            use_dynamic = hasattr(float_mod, "a_bit") and act_dtype in _W8 + _W4
        else:
            w_pair = create_qparams_from_dtype(dtype=w_dtype, device=device, is_channel_wise=True,
                                               num_kernels=n_out, ckpt=ckpt,
                                               module_name=float_mod.module_name,
                                               quant_type="weight", bit_width=w_bit_eff,
                                               split=split)
            w_qparams = w_pair[0] if w_pair is not None else None
            a_qparams = None
            if hasattr(float_mod, "a_bit"):
                a_pair = create_qparams_from_dtype(dtype=act_dtype, device=device,
                                                   is_channel_wise=False, num_kernels=n_out,
                                                   ckpt=ckpt, module_name=float_mod.module_name,
                                                   quant_type="act", bit_width=float_mod.a_bit,
                                                   split=split)
                a_qparams = a_pair[0] if a_pair is not None else None
                if a_qparams is not None and act_dtype in _W4:
                    # 4-bit codes stay unsigned 0..15: undo the uint8 -> int8 shift of get_quant_para
                    a_qparams = a_qparams._replace(zero_points=a_qparams.zero_points + 128)
            use_dynamic = False

        new_mod = cls(float_mod.in_features, float_mod.out_features, float_mod.bias is not None,
                      device=device, w_qparams=w_qparams, a_qparams=a_qparams,
                      module_name=float_mod.module_name, dynamic=use_dynamic,
                      a_bits=4 if act_dtype in _W4 else 8)

        name = float_mod.module_name or ""
        if "attn2" in name and ("to_k" in name or "to_v" in name) and hasattr(float_mod, "bos"):
            new_mod.bos = float_mod.bos
            new_mod.register_buffer("bos_pre_computed", float_mod.bos_pre_computed)

        if new_mod.valid_for_acceleration:
            if new_mod.w_kind == "w8":
                if dynamic:
                    weight_int = quantize_weight(weight, new_mod.weight_scales, 8, exact_division=True)
                else:
                    weight_int = torch.quantize_per_channel(
                        weight.float(), new_mod.weight_scales, new_mod.weight_zero_points,
                        axis=0, dtype=w_qparams.dtype).int_repr()
                new_mod.register_buffer("weight_int", weight_int)
            else:
                weight_int = quantize_weight(weight, new_mod.weight_scales, 4,
                                             exact_division=dynamic)
                new_mod.register_buffer("weight_int4", pack_int4(weight_int))
            wsum = weight_int.float().sum(dim=1)
            new_mod.register_buffer("weight_sum_by_input_channels", wsum)
            if not new_mod.dynamic:
                new_mod.register_buffer("scale", new_mod.weight_scales * new_mod.act_scales)
                new_mod.register_buffer("bias0", wsum * new_mod.act_zero_points)
        else:
            new_mod.register_buffer("weight", weight)
        if float_mod.bias is not None:
            new_mod.register_buffer("bias", float_mod.bias.detach())
        else:
            new_mod.bias = None
        return new_mod

    def _get_name(self):
        if self.valid_for_acceleration:
            return f"QuantizedLinear{'W8' if self.w_kind == 'w8' else 'W4'}A{self.a_bits}"
        return "QuantizedLinearFPFallback"

    # ------------------------------------------------------------------------------------
    def _dequantized_weight(self, dtype):
        if self.w_kind == "w8":
            w = self.weight_int.float()
        else:
            from .utils import unpack_int4
            w = unpack_int4(self.weight_int4).float()
        return (w * self.weight_scales[:, None]).to(dtype)

    def forward_fallback(self, x):
        """Non-fp16 input: the int8 kernels only take fp16 activations, so the layer runs as a
        float op on the DEQUANTISED weight — never silently: warns once per module."""
        if not getattr(self, "_warned_fallback", False):
            self._warned_fallback = True
            logging.warning(f"{self._get_name()} {self.module_name}: input dtype {x.dtype} is not "
                            "fp16; running F.linear on the dequantised weight (no int8 kernel)")
        return F.linear(x, self._dequantized_weight(x.dtype),
                        self.bias.to(x.dtype) if self.bias is not None else None)

    def _qlinear(self, x: torch.Tensor) -> torch.Tensor:
        if self.dynamic:
            x_int, a_scale, a_zp = ops.quantize_per_tensor_dynamic_bits(x, self.a_bits)
            if self.w_kind == "w8":
                return ops.qlinear_w8_a8_ohalf_dynamic(
                    x_int, self.weight_int, self.weight_scales, a_scale, a_zp,
                    self.weight_sum_by_input_channels, self.bias)
            # packed 4-bit weights, unpacked inside the tcgen05 kernel; the activation scalars
            # are folded in its epilogue
            return ops.qlinear_dynamic_fused(x_int, self.weight_int4, self.weight_scales, a_scale,
                                             a_zp, self.weight_sum_by_input_channels, self.bias)
        if self.a_bits == 4:
            x_int = ops.quantize_per_tensor_to_int4_codes(x, self.act_scales_inv,
                                                          self.act_zero_points)
        else:
            x_int = ops.quantize_per_tensor_to_int8(x, self.act_scales_inv, self.act_zero_points)
        if self.w_kind == "w8":
            return ops.qlinear_w8_a8_ohalf(
                x_int, self.weight_int, self.weight_scales, self.act_scales,
                self.act_zero_points, self.weight_sum_by_input_channels, self.scale, self.bias0,
                self.bias)
        return ops.qlinear_w4_a8_ohalf(x_int, self.weight_int4, self.scale, self.bias0, self.bias)

    # ---- GEGLU-interleaved stored layout (mixdq_b200.fused.GegluLinear) ---------------------
    # When the block-level fusion evaluates the GEGLU in this projection's epilogue the rows are
    # STORED interleaved (16 value rows, their 16 gate rows, ...). The module's own forward and
    # its state_dict keep the stock row order by undoing the permutation on the way out.
    def _stock_order(self, t: torch.Tensor, dim: int) -> torch.Tensor:
        if getattr(self, "geglu_interleaved", False):
            return t.index_select(dim, self.geglu_inverse_index.to(t.device))
        return t

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        if getattr(self, "geglu_interleaved", False):
            for k in ("weight_int", "weight_int4", "weight_scales", "weight_zero_points",
                      "weight_sum_by_input_channels", "scale", "bias0", "bias"):
                if prefix + k in destination:
                    destination[prefix + k] = self._stock_order(destination[prefix + k], 0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.valid_for_acceleration:
            return F.linear(x, self.weight, self.bias)
        if x.dtype != torch.float16:
            return self._stock_order(self.forward_fallback(x), -1)
        if not getattr(self, "bos", False):
            return self._stock_order(self._qlinear(x), -1)
        # BOS-aware cross-attention K/V: the first text token bypasses quantisation and takes a
        # pre-computed fp16 output (reference nn/Linear.py:178-194).
        out_rest = self._qlinear(x[:, 1:, :])
        out_first = self.bos_pre_computed.expand(x.shape[0], -1, -1)
        return torch.cat([out_first, out_rest], dim=1)
