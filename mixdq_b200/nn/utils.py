"""Quantisation-parameter plumbing for the quantized modules.

Mirrors the interface of the reference's kernels/mixdq_extension/nn/utils.py (QParam :64-68,
create_qparams_from_dtype :79-135, get_quant_para :412-458, uint4 helpers :13-52) so that
`from_float` reads the same PTQ checkpoint format:

    ckpt[name + '.weight_quantizer'] = {'delta_list': f16 [3, Cout], 'zero_point_list': f16 [3, Cout]}
    ckpt[name + '.act_quantizer']    = {'delta_list': f16 [3],       'zero_point_list': f16 [3]}
    (+ '..._0' twins for the second channel range of a split shortcut); index 0/1/2 = 2/4/8 bit.

Additions over the reference: min-max weight parameters computed from the weights when no
checkpoint is given (dynamic mode), and signed-int4 packing for the W4A8 kernels.
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import Optional, Tuple

import torch

dtype_to_bw = {
    torch.qint8: 8,
    torch.quint8: 8,
    torch.quint4x2: 4,
    torch.quint2x4: 2,
    torch.float16: 16,
}


class QParam(namedtuple("QParam", ["qscheme", "dtype", "scales", "zero_points", "axis"],
                        defaults=[torch.per_tensor_affine, torch.quint8, 1.0, 0.0, 0])):
    """(qscheme, dtype, scales, zero_points, axis) — field-compatible with the reference tuple."""

    @property
    def zp_float(self):
        return self.scales * self.zero_points


def bit_index(n_bit: int) -> int:
    """Row of delta_list / zero_point_list holding the n_bit parameters (2/4/8 -> 0/1/2)."""
    return int(math.log2(n_bit) - 1)


def get_quant_para(ckpt, n_bit, module_name, quant_type, split=0, device=None):
    """Look up (scales, zero_point, scales_0, zero_point_0) for one module.

    Activation zero points are stored for uint8 codes; the kernels use int8 codes, hence the
    -128 shift (reference nn/utils.py:428). `split > 0` additionally returns the '_0' twin that
    quantises input channels [split:]."""
    if quant_type not in ("weight", "act"):
        raise ValueError(f"unknown quant_type {quant_type}")
    idx = bit_index(n_bit)
    key = f"{module_name}.{'weight' if quant_type == 'weight' else 'act'}_quantizer"
    shift = 128 if quant_type == "act" else 0

    def fetch(k):
        if k not in ckpt:
            raise KeyError(f"{k} not found in the quantisation checkpoint")
        entry = ckpt[k]
        return entry["delta_list"][idx].to(device), (entry["zero_point_list"][idx] - shift).to(device)

    scales, zero_point = fetch(key)
    if split > 0:
        scales_0, zero_point_0 = fetch(key + "_0")
        return scales, zero_point, scales_0, zero_point_0
    return scales, zero_point, None, None


def create_qparams_from_dtype(dtype, device, is_channel_wise=False, num_kernels=None, ckpt=None,
                              module_name=None, bit_width=0, quant_type=None, split=0):
    """Build the QParam pair (main, second-half-or-None) of a module from the checkpoint."""
    if dtype == torch.float16:
        return None
    if dtype not in (torch.qint8, torch.quint8, torch.quint4x2):
        raise ValueError(f"Unsupported quantize dtype {dtype}")
    scales, zps, scales_0, zps_0 = get_quant_para(ckpt, bit_width, module_name, quant_type,
                                                  split=split, device=device)
    if is_channel_wise:
        assert num_kernels is not None
        kw = dict(qscheme=torch.per_channel_affine, dtype=dtype, axis=0)
    else:
        kw = dict(qscheme=torch.per_tensor_affine, dtype=dtype)
    main = QParam(scales=scales, zero_points=zps, **kw)
    second = QParam(scales=scales_0, zero_points=zps_0, **kw) if split > 0 else None
    return main, second


def minmax_weight_scales(weight: torch.Tensor, n_bits: int) -> torch.Tensor:
    """Symmetric per-output-channel min-max scales, the qdiff weight quantizer's init
    (reference base_quantizer.py:147-185, sym=True): delta_c = max|w_c| / (2^(b-1) - 1)."""
    w = weight.detach().float().reshape(weight.shape[0], -1)
    delta = w.abs().amax(dim=1) / (2 ** (n_bits - 1) - 1)
    if delta.min() < 1e-6:
        delta = torch.full_like(delta, 1e-6)
    return delta


def quantize_weight(weight: torch.Tensor, scales: torch.Tensor, n_bits: int = 8,
                    exact_division: bool = False) -> torch.Tensor:
    """Per-output-channel symmetric weight codes as int8.

    exact_division=False reproduces torch.quantize_per_channel (multiply by the fp32 reciprocal,
    round half to even, clamp) that the reference's from_float uses (nn/Linear.py:116-121);
    exact_division=True reproduces the qdiff quantizer round(w / delta)."""
    lo, hi = -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1
    shape = [-1] + [1] * (weight.dim() - 1)
    s = scales.float().reshape(shape)
    w = weight.detach().float()
    q = torch.round(w / s) if exact_division else torch.round(w * (1.0 / s))
    return torch.clamp(q, lo, hi).to(torch.int8)


def pack_int4(codes: torch.Tensor, dim: int = -1) -> torch.Tensor:
    """int8 codes in [-8, 7] -> uint8 with `dim` halved: even index in the HIGH nibble (the
    reference's only packing convention, nn/utils.py:26-28 — last dim, or dim 1 for 4-D conv
    weights, :20-24), two's-complement nibbles. A channels_last 4-D input packed along dim 1 stays
    channels_last, i.e. KRS(C/2) in memory — what the W4 convolution kernel consumes."""
    assert codes.shape[dim] % 2 == 0
    c = codes.to(torch.int16).movedim(dim, -1)
    packed = (((c[..., 0::2] & 0xF) << 4) | (c[..., 1::2] & 0xF)).to(torch.uint8).movedim(-1, dim)
    if codes.dim() == 4 and codes.is_contiguous(memory_format=torch.channels_last):
        return packed.contiguous(memory_format=torch.channels_last)
    return packed.contiguous()


def unpack_int4(packed: torch.Tensor, dim: int = -1) -> torch.Tensor:
    p = packed.to(torch.int16).movedim(dim, -1)
    both = torch.stack([(p >> 4) & 0xF, p & 0xF], dim=-1)
    both = both.reshape(*p.shape[:-1], p.shape[-1] * 2)
    out = torch.where(both >= 8, both - 16, both).to(torch.int8).movedim(-1, dim)
    if packed.dim() == 4 and packed.is_contiguous(memory_format=torch.channels_last):
        return out.contiguous(memory_format=torch.channels_last)
    return out.contiguous()
