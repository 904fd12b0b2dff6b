"""Block-level fusion of the quantized UNet (dynamic W8A8): fewer, fatter kernels per step.

The reference swaps only the nn.Linear / nn.Conv2d leaves (kernels/quantize.py:606-669) and leaves
every op between two quantized layers to stock PyTorch; at batch 1 the step is then dominated by
~1 700 short kernels. `fuse_unet` goes one level up — it re-binds the
`forward` of the diffusers-named blocks (the reference patches block forwards the same way for its
oracle: quant_utils/qdiff/models/quant_block_forward_func.py:54-66) so that

  * the op that produces a quantized layer's input is fused with the dynamic quantisation of its
    result: LayerNorm -> int8 (to_q/k/v, ff.net.0.proj), GroupNorm[+SiLU] -> int8 (resnet convs,
    proj_in), GEGLU -> int8 (ff.net.2)                        [csrc/fused_quant.cu];
  * layers that consume the SAME tensor run as ONE contraction over N-concatenated weights (one
    per weight kind when W8 and packed-W4 layers share an input):
    attn1.to_q/to_k/to_v; every attn2.to_k/to_v of the UNet (they all read
    `encoder_hidden_states`); every resnet `time_emb_proj` (they all read silu(temb));
  * the fp16 elementwise op that follows a layer runs in its epilogue: residual adds
    (`x + attn`, `x + ff`, `proj_out + res`, resnet `x + h`) and `h + temb[:, :, None, None]`.

Arithmetic is unchanged: with dynamic per-tensor scales the quantized codes of a shared input are
identical for all its consumers, per-output-channel weight scales make N-concatenation exact, and
every fused elementwise op keeps its own fp16 rounding (tests/test_gpu_fused.py,
tests/test_gpu_modules.py::test_fused_unet_matches_unfused). A block is fused only if all the
layers involved are dynamic W8A8 / W4A8 `QuantizedLinear` / `QuantizedConv2d` on the tcgen05 path;
anything else (static checkpoints, fp16-protected layers) keeps the stock forward.

Call it after the model sits on its CUDA device: concatenated weights become the storage of the
member layers' `weight_int` buffers (views), which a later `.to(device)` would split again.
"""
from __future__ import annotations

import types
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .nn.conv2d import QuantizedConv2d
from .nn.linear import QuantizedLinear


# ---------------------------------------------------------------------------------------------
# eligibility
# ---------------------------------------------------------------------------------------------
def _lin_ok(m) -> bool:
    if not (isinstance(m, QuantizedLinear) and m.valid_for_acceleration):
        return False
    if m.a_bits != 8:
        return False                       # 4-bit activation layers run module by module
    k_align = 32 if m.w_kind == "w4" else 16
    return m.in_features % k_align == 0 and m.out_features % 8 == 0


def _lin_weight(m: QuantizedLinear) -> torch.Tensor:
    """int8 codes [N, K] (W8) or packed uint8 [N, K/2] (W4): ops dispatch on the dtype"""
    return m.weight_int if m.w_kind == "w8" else m.weight_int4


def _set_lin_weight(m: QuantizedLinear, w: torch.Tensor) -> None:
    if m.w_kind == "w8":
        m.weight_int = w
    else:
        m.weight_int4 = w


def _conv_ok(m) -> bool:
    if not (isinstance(m, QuantizedConv2d) and m.valid_for_acceleration):
        return False
    pad, stride, k = m.padding[0], m.stride[0], m.kernel_size[0]
    geom = stride in (1, 2) and (pad == 0 or (pad == 1 and k == 3))
    if getattr(m, "weight_int4", None) is not None and m.split != 0:
        return False                       # packed W4 split shortcut: two convolutions, unfused
    return geom and m.in_channels % 16 == 0 and m.out_channels % 8 == 0 and \
        (m.split == 0 or (m.split % 16 == 0 and k == 1 and pad == 0 and stride == 1))


def _gn_ok(norm: nn.GroupNorm) -> bool:
    c, g = norm.num_channels, norm.num_groups
    cpg = c // g
    return (c % g == 0 and (cpg >= 8 or cpg == 4) and g <= 32 and c <= 2560 and c % 8 == 0
            and norm.weight is not None and norm.weight.dtype == torch.float16)


def _ln_ok(norm: nn.LayerNorm) -> bool:
    return (norm.weight is not None and norm.weight.dtype == torch.float16
            and norm.normalized_shape[-1] % 8 == 0 and norm.normalized_shape[-1] <= 2048)


# ---------------------------------------------------------------------------------------------
# activation-quantisation policy of a consumer layer
#   dynamic: per-tensor min-max of the tensor itself (qdiff, base_quantizer.py:155-190)
#   static : the layer's PTQ-checkpoint parameters and the reference's own formula
#            (quantize_per_tensor_to_int8, quantize.cc:9-30). The GEMM / conv entry points take the
#            activation scalars from device memory and form scale[n] = w_scale[n] * a_scale and
#            bias0[n] = wsum[n] * a_zp in fp32 — the same two products `from_float` stores as the
#            `scale` / `bias0` buffers (nn/Linear.py:125-132) — so the static path is bit-identical
#            to the unfused static modules wherever the activation tensor is.
# ---------------------------------------------------------------------------------------------
def _act_key(m, sfx: str = ""):
    """None for a dynamic layer, else its static (delta, zero point) as a hashable key"""
    if m.dynamic:
        return None
    return (float(getattr(m, "act_scales" + sfx)), float(getattr(m, "act_zero_points" + sfx)))


def _same_act(mods) -> bool:
    return len({_act_key(m) for m in mods}) == 1


def _static_args(m, sfx: str = ""):
    return (getattr(m, "act_scales_inv" + sfx), getattr(m, "act_scales" + sfx),
            getattr(m, "act_zero_points" + sfx))


def _q_ln(x, norm, cons):
    """LayerNorm -> int8 for consumer `cons`: (codes, scale, zero point)"""
    if cons.dynamic:
        return ops.layernorm_quantize_dynamic(x, norm.weight, norm.bias, norm.eps)
    inv, sc, zp = _static_args(cons)
    return ops.layernorm_quantize_static(x, norm.weight, norm.bias, norm.eps, inv, zp), sc, zp


def _q_gn(x, norm, silu: bool, cons):
    """GroupNorm [+ SiLU] -> int8 channels_last for consumer `cons`"""
    if cons.dynamic:
        return ops.groupnorm_quantize_dynamic(x, norm.num_groups, norm.weight, norm.bias, norm.eps,
                                              silu=silu)
    inv, sc, zp = _static_args(cons)
    return ops.groupnorm_quantize_static(x, norm.num_groups, norm.weight, norm.bias, norm.eps,
                                         silu, inv, zp), sc, zp


def _q_act(x, cons):
    """plain tensor ([B, T, C] tokens, dense or row-pitched; [B, K]; NHWC images) -> int8"""
    if cons.dynamic:
        return _quant_tokens(x) if x.dim() == 3 else ops.quantize_per_tensor_dynamic(x)
    inv, sc, zp = _static_args(cons)
    return ops.quantize_per_tensor_to_int8(x, inv, zp), sc, zp


# ---------------------------------------------------------------------------------------------
# N-concatenated linears
# ---------------------------------------------------------------------------------------------
class CatLinear:
    """Several dynamic linears with the same in_features, run as one GEMM per weight kind (W8
    codes / packed W4: N-concatenation needs one storage format). The members' buffers become
    views of the concatenated storage (no second copy of the weights)."""

    def __init__(self, mods: List[QuantizedLinear]):
        assert len({m.in_features for m in mods}) == 1
        self.mods = mods
        self.sizes = [m.out_features for m in mods]
        has_bias = [m.bias is not None for m in mods]
        assert all(has_bias) or not any(has_bias)
        self.parts = []                      # one (weight, scales, wsum, bias, n) per weight kind
        self.where = []                      # member i -> (part index, column offset)
        slot = {}
        for kind in ("w8", "w4"):
            members = [m for m in mods if m.w_kind == kind]
            if not members:
                continue
            weight = torch.cat([_lin_weight(m) for m in members], dim=0)
            scales = torch.cat([m.weight_scales for m in members])
            wsum = torch.cat([m.weight_sum_by_input_channels for m in members])
            bias = torch.cat([m.bias for m in members]) if all(has_bias) else None
            off = 0
            for m in members:
                n = m.out_features
                _set_lin_weight(m, weight[off:off + n])
                m.weight_scales = scales[off:off + n]
                m.weight_sum_by_input_channels = wsum[off:off + n]
                if bias is not None:
                    m.bias = bias[off:off + n]
                slot[id(m)] = (len(self.parts), off)
                off += n
            self.parts.append((weight, scales, wsum, bias, off))
        self.where = [slot[id(m)] for m in mods]
        self.n_total = sum(self.sizes)
        # single-kind views kept for callers / tests that look at the concatenated operands
        self.weight_int, self.weight_scales, self.wsum, self.bias, _ = self.parts[0]

    def run_parts(self, q8, scale, zp):
        """one output tensor per weight kind"""
        return [ops.qlinear_dynamic_fused(q8, w, sc, scale, zp, ws, b)
                for (w, sc, ws, b, _) in self.parts]

    def run(self, q8, scale, zp):
        """[..., n_total] in member order (single weight kind: the GEMM output itself)"""
        outs = self.run_parts(q8, scale, zp)
        if len(outs) == 1:
            return outs[0]
        return torch.cat([self.slice_of(outs, i) for i in range(len(self.mods))], dim=-1)

    def slice_of(self, outs, i: int) -> torch.Tensor:
        part, off = self.where[i]
        return outs[part][..., off:off + self.sizes[i]]


class GegluLinear:
    """ff.net.0.proj with the GEGLU in its epilogue: the projection's rows / scales / sums / bias
    re-ordered in place (16 value rows, then their 16 gate rows), so one accumulator chunk holds
    both operands of 16 outputs. The interleaved order becomes the module's STORED layout —
    `mod.geglu_interleaved = True` — instead of a second copy of a 13 MB weight per block; the
    module's own (unfused) forward and its state_dict undo the permutation on the way out."""

    def __init__(self, mod: QuantizedLinear):
        inner = mod.out_features // 2
        idx = ops.geglu_interleave_index(inner, mod.weight_scales.device)
        if not getattr(mod, "geglu_interleaved", False):
            _set_lin_weight(mod, _lin_weight(mod).index_select(0, idx).contiguous())
            mod.weight_scales = mod.weight_scales.index_select(0, idx).contiguous()
            mod.weight_sum_by_input_channels = \
                mod.weight_sum_by_input_channels.index_select(0, idx).contiguous()
            if mod.bias is not None:
                mod.bias = mod.bias.index_select(0, idx).contiguous()
            for k in ("scale", "bias0", "weight_zero_points"):      # static-mode per-row buffers
                if getattr(mod, k, None) is not None:
                    setattr(mod, k, getattr(mod, k).index_select(0, idx).contiguous())
            mod.geglu_interleaved = True
            mod.register_buffer("geglu_inverse_index", torch.argsort(idx), persistent=False)
        self.mod = mod

    def run(self, q8, scale, zp, cons=None):
        """-> int8 GEGLU output quantised for `cons` (ff.net.2)"""
        m = self.mod
        if cons is None or cons.dynamic:
            return ops.qlinear_geglu_quantize_dynamic(q8, _lin_weight(m), m.weight_scales, scale,
                                                      zp, m.weight_sum_by_input_channels, m.bias)
        inv, sc, z2 = _static_args(cons)
        return ops.qlinear_geglu_quantize_static(q8, _lin_weight(m), m.weight_scales, scale, zp,
                                                 m.weight_sum_by_input_channels, m.bias,
                                                 inv, z2), sc, z2


class SharedInputGroup:
    """Layers of the whole UNet that consume one tensor (all attn2.to_k/to_v <- encoder hidden
    states; all resnet time_emb_proj <- silu(temb)): quantise once, one GEMM, hand out column
    slices.

    The result is cached for the duration of ONE UNet forward: `fuse_unet` registers a forward
    pre-hook on the UNet that calls `reset()`, and the key also carries the tensor's identity
    (held strongly, so the id cannot be recycled), its version counter (where the tensor tracks
    one) and the CUDA-graph capture id of the current stream. A value computed eagerly (e.g. in
    the warm-up passes of `cuda_graph_opt`) is therefore never reused inside a capture — its
    kernels would be missing from the graph and every replay would see the first prompt's K/V."""

    def __init__(self, mods: List[QuantizedLinear], pre=None, bos: bool = False):
        self.cat = CatLinear(mods)
        self.index = {id(m): i for i, m in enumerate(mods)}
        self.pre = pre
        self.bos = bos
        if bos:
            # pre-computed first-token rows, one [1, 1, N_part] tensor per weight kind
            self.bos_rows = [torch.cat([m.bos_pre_computed for m in mods if m.w_kind == kind],
                                       dim=-1)
                             for kind in ("w8", "w4") if any(m.w_kind == kind for m in mods)]
        self._key = None
        self._out = None

    def reset(self) -> None:
        self._key = None
        self._out = None

    def get(self, mod, x: torch.Tensor) -> torch.Tensor:
        key = (x, ops._version_of(x), ops._capture_id(x.device))
        if self._key is None or self._key[0] is not x or self._key[1:] != key[1:]:
            xin = self.pre(x) if self.pre is not None else x
            if self.bos:
                xin = xin[:, 1:, :]
            q8, s, z = _q_act(xin, self.cat.mods[0])
            outs = self.cat.run_parts(q8, s, z)
            if self.bos:
                outs = [torch.cat([rows.expand(o.shape[0], -1, -1), o], dim=1)
                        for rows, o in zip(self.bos_rows, outs)]
            self._key = key
            self._out = outs
        return self.cat.slice_of(self._out, self.index[id(mod)])


def _run_linear(m: QuantizedLinear, q8, s, z, residual=None):
    return ops.qlinear_dynamic_fused(q8, _lin_weight(m), m.weight_scales, s, z,
                                     m.weight_sum_by_input_channels, m.bias, residual)


def _run_conv(m: QuantizedConv2d, q8, s, z, chan_add=None, residual=None):
    pad = m.padding[0]
    return ops.qconv2d_dynamic_fused(
        q8, m._weight(""), m.weight_scales, s, z,
        m.weight_sum_by_input_channels if pad > 0 else None,
        m.weight_sum_per_output_channel if pad == 0 else None,
        m.bias, m.stride[0], pad, chan_add, residual)


def _quant_tokens(x: torch.Tensor):
    """dynamic quantisation of [B, T, C] fp16, dense or row-pitched."""
    if x.is_contiguous():
        return ops.quantize_per_tensor_dynamic(x)
    b, t, c = x.shape
    if x.stride(2) == 1 and x.stride(0) == t * x.stride(1):
        q, s, z = ops.quantize_rows_dynamic(x.reshape(b * t, c))
        return q.view(b, t, c), s, z
    return ops.quantize_per_tensor_dynamic(x.contiguous())


# ---------------------------------------------------------------------------------------------
# fused block forwards (attribute names = diffusers names)
# ---------------------------------------------------------------------------------------------
def _heads(t: torch.Tensor, heads: int):
    b, n, c = t.shape
    return t.view(b, n, heads, c // heads).transpose(1, 2)


import os as _os
# An own cross-attention kernel (mma.sync, min/max partials of its output) was measured slower than
# the library SDPA + a separate min/max pass inside the whole-UNet graph (13.2 us vs 5.7 + 2.4 us,
# profiles/README.md section 3) and removed in round 2.
_SDPA_BACKEND = None      # None = PyTorch's own choice; set by MIXDQ_SDPA_BACKEND (tuning aid)


def _sdpa_ctx():
    """Optional pin of the library attention backend (flash | efficient | cudnn | math) for A/B
    timing; attention itself is stock PyTorch in the reference and stays a library call here."""
    global _SDPA_BACKEND
    if _SDPA_BACKEND is None:
        import os
        _SDPA_BACKEND = os.environ.get("MIXDQ_SDPA_BACKEND", "").lower() or "default"
    if _SDPA_BACKEND == "default":
        return None
    from torch.nn.attention import SDPBackend, sdpa_kernel
    table = {"flash": SDPBackend.FLASH_ATTENTION, "efficient": SDPBackend.EFFICIENT_ATTENTION,
             "cudnn": SDPBackend.CUDNN_ATTENTION, "math": SDPBackend.MATH}
    return sdpa_kernel(table[_SDPA_BACKEND])


def _attention(q, k, v, heads):
    ctx = _sdpa_ctx()
    if ctx is None:
        o = F.scaled_dot_product_attention(_heads(q, heads), _heads(k, heads), _heads(v, heads))
    else:
        with ctx:
            o = F.scaled_dot_product_attention(_heads(q, heads), _heads(k, heads), _heads(v, heads))
    b, h, t, d = o.shape
    return o.transpose(1, 2).reshape(b, t, h * d)


def _is_native(m) -> bool:
    """True for the in-repo skeleton classes (mixdq_b200.unet); anything else with a diffusers
    class name is called / answered with the diffusers conventions (keyword context, tuple or
    output-class returns)."""
    return type(m).__module__.startswith("mixdq_b200.")


# call-time arguments of the diffusers block forwards that the fused forwards do not implement:
# when any of them carries a value the original forward runs instead
_UNSUPPORTED_KW = ("attention_mask", "encoder_attention_mask", "timestep", "class_labels",
                   "added_cond_kwargs")


def _needs_stock_forward(args, kwargs, n_pos_ok: int) -> bool:
    if len(args) > n_pos_ok:
        return any(a is not None for a in args[n_pos_ok:])
    for k in _UNSUPPORTED_KW:
        if kwargs.get(k) is not None:
            return True
    cak = kwargs.get("cross_attention_kwargs")
    return bool(cak)


def _ctx_of(self, args, kwargs):
    """encoder_hidden_states of a block call: keyword, or positional (index 0 for the in-repo
    skeleton `blk(x, ctx)`, index 1 for diffusers `blk(x, attention_mask, ctx, ...)`)."""
    if "encoder_hidden_states" in kwargs:
        return kwargs["encoder_hidden_states"]
    if _is_native(self):
        return args[0] if args else None
    return args[1] if len(args) > 1 else None


def fused_transformer_block_forward(self, hidden_states, *args, **kwargs):
    """BasicTransformerBlock: 14 kernels instead of ~30 (see module docstring)."""
    f = self._mixdq_fused
    if not _is_native(self) and (_needs_stock_forward(args, kwargs, 2)
                                 or (args and args[0] is not None)):
        return f["orig_forward"](hidden_states, *args, **kwargs)
    ctx = _ctx_of(self, args, kwargs)
    x = hidden_states
    c = x.shape[-1]
    # --- self-attention ---
    q8, s, z = _q_ln(x, self.norm1, self.attn1.to_q)
    cat = f["qkv"]
    outs = cat.run_parts(q8, s, z)
    o = _attention(cat.slice_of(outs, 0), cat.slice_of(outs, 1), cat.slice_of(outs, 2),
                   self.attn1.heads)
    o8, s, z = _q_act(o, self.attn1.to_out[0])
    x = _run_linear(self.attn1.to_out[0], o8, s, z, residual=x)
    # --- cross-attention ---
    q8, s, z = _q_ln(x, self.norm2, self.attn2.to_q)
    q = _run_linear(self.attn2.to_q, q8, s, z)
    kv = f["kv"]
    kk, vv = kv.get(self.attn2.to_k, ctx), kv.get(self.attn2.to_v, ctx)
    heads = self.attn2.heads
    o = _attention(q, kk, vv, heads)
    o8, s, z = _q_act(o, self.attn2.to_out[0])
    x = _run_linear(self.attn2.to_out[0], o8, s, z, residual=x)
    # --- feed-forward ---
    q8, s, z = _q_ln(x, self.norm3, self.ff.net[0].proj)
    ff2 = self.ff.net[2]
    if f.get("ffproj") is not None:
        g8, s, z = f["ffproj"].run(q8, s, z, ff2)
    else:
        hg = _run_linear(self.ff.net[0].proj, q8, s, z)
        if ff2.dynamic:
            g8, s, z = ops.geglu_quantize_dynamic(hg)
        else:
            h, gate = hg.chunk(2, dim=-1)
            g8, s, z = _q_act(h * F.gelu(gate), ff2)
    return _run_linear(ff2, g8, s, z, residual=x)


def fused_transformer2d_forward(self, hidden_states, *args, **kwargs):
    """Transformer2DModel: GroupNorm -> int8 feeds proj_in; proj_out adds the residual.
    diffusers conventions are honoured for non-skeleton classes: `encoder_hidden_states` as the
    first positional or a keyword, `return_dict` (default True) selects the module's output class
    or a 1-tuple; calls carrying masks / class labels / cross_attention_kwargs run the original
    forward."""
    native = _is_native(self)
    if not native and _needs_stock_forward(args, kwargs, 1):
        return self._mixdq_fused["orig_forward"](hidden_states, *args, **kwargs)
    ctx = kwargs["encoder_hidden_states"] if "encoder_hidden_states" in kwargs else \
        (args[0] if args else None)
    x = hidden_states
    b, c, h, w = x.shape
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    q8, s, z = _q_gn(x, self.norm, False, self.proj_in)
    y = _run_linear(self.proj_in, q8.permute(0, 2, 3, 1).reshape(b, h * w, c), s, z)
    for blk in self.transformer_blocks:
        y = blk(y, ctx) if native else blk(y, encoder_hidden_states=ctx)
    o8, s, z = _q_act(y, self.proj_out)
    res = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    out = _run_linear(self.proj_out, o8, s, z, residual=res)
    out = out.reshape(b, h, w, c).permute(0, 3, 1, 2)
    if native:
        return out
    if not kwargs.get("return_dict", True):
        return (out,)
    import sys
    out_cls = getattr(sys.modules.get(type(self).__module__), "Transformer2DModelOutput", None)
    return out_cls(sample=out) if out_cls is not None else (out,)


def fused_resnet_forward(self, input_tensor, temb, *args, **kwargs):
    """ResnetBlock2D: GroupNorm+SiLU -> int8 feeds both convs; conv1 adds the time embedding,
    conv2 adds the (shortcut of the) input."""
    f = self._mixdq_fused
    x = input_tensor
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    n1, n2 = self.norm1, self.norm2
    h8, s, z = _q_gn(x, n1, True, self.conv1)
    t = f["temb"].get(self.time_emb_proj, temb)                  # [B, K] fp16 (column slice)
    h = _run_conv(self.conv1, h8, s, z, chan_add=t)
    h8, s, z = _q_gn(h, n2, True, self.conv2)
    sc = self.conv_shortcut
    if sc is None:
        res = x
    elif sc.split == 0:
        x8, xs, xz = _q_act(x, sc)
        res = _run_conv(sc, x8, xs, xz)
    else:
        c = x.shape[1]
        if sc.dynamic:
            xa, sa, za = ops.quantize_nhwc_slice_dynamic(x, 0, sc.split)
            xb, sb, zb = ops.quantize_nhwc_slice_dynamic(x, sc.split, c)
        else:
            (ia, sa, za), (ib, sb, zb) = _static_args(sc), _static_args(sc, "_0")
            xa = ops.quantize_to_nhwc(x, ia, za, 0, sc.split)
            xb = ops.quantize_to_nhwc(x, ib, zb, sc.split, c)
        res = ops.qconv1x1_split_dynamic_fused(
            xa, sc.weight_int, sc.weight_scales, sc.weight_sum_per_output_channel, sa, za,
            xb, sc.weight_int_0, sc.weight_scales_0, sc.weight_sum_per_output_channel_0, sb, zb,
            sc.bias)
    return _run_conv(self.conv2, h8, s, z, residual=res)


# ---------------------------------------------------------------------------------------------
# the pass
# ---------------------------------------------------------------------------------------------
def _is(m, name: str) -> bool:
    return type(m).__name__ == name


def _plain(obj, **expected) -> bool:
    """every attribute named in `expected` is absent or equal to the expected value"""
    return all(getattr(obj, k, v) == v for k, v in expected.items())


def _attn_plain(attn) -> bool:
    """diffusers Attention features the fused forward ignores must be off."""
    head = attn.to_q.out_features // attn.heads
    scale = getattr(attn, "scale", head ** -0.5)
    return (_plain(attn, group_norm=None, spatial_norm=None, norm_cross=None, norm_q=None,
                   norm_k=None, residual_connection=False, rescale_output_factor=1.0,
                   upcast_attention=False, upcast_softmax=False, added_kv_proj_dim=None,
                   only_cross_attention=False)
            and abs(float(scale) - head ** -0.5) < 1e-6)


def _block_ok(blk) -> bool:
    if not _is_native(blk):
        if not _plain(blk, only_cross_attention=False, use_ada_layer_norm=False,
                      use_ada_layer_norm_zero=False, use_ada_layer_norm_single=False,
                      use_ada_layer_norm_continuous=False, use_layer_norm=True,
                      pos_embed=None, _chunk_size=None, fuser=None):
            return False
        if getattr(blk, "norm_type", "layer_norm") != "layer_norm":
            return False
        if not (_attn_plain(blk.attn1) and _attn_plain(blk.attn2)):
            return False
    try:
        lins = [blk.attn1.to_q, blk.attn1.to_k, blk.attn1.to_v, blk.attn1.to_out[0],
                blk.attn2.to_q, blk.attn2.to_k, blk.attn2.to_v, blk.attn2.to_out[0],
                blk.ff.net[0].proj, blk.ff.net[2]]
        norms = [blk.norm1, blk.norm2, blk.norm3]
    except AttributeError:
        return False
    if not all(_lin_ok(m) for m in lins) or not all(isinstance(n, nn.LayerNorm) and _ln_ok(n) for n in norms):
        return False
    # one mode per block; q / k / v read ONE quantised tensor, so static parameters must agree
    if len({m.dynamic for m in lins}) != 1 or not _same_act(lins[:3]) or not _same_act(lins[5:7]):
        return False
    bos = [bool(getattr(m, "bos", False)) for m in (blk.attn2.to_k, blk.attn2.to_v)]
    return bos[0] == bos[1]


def _resnet_ok(res) -> bool:
    if not _is_native(res) and not _plain(res, output_scale_factor=1.0, up=False, down=False,
                                          time_embedding_norm="default", upsample=None,
                                          downsample=None):
        return False
    try:
        ok = (_conv_ok(res.conv1) and res.conv1.split == 0 and _conv_ok(res.conv2)
              and res.conv2.split == 0 and _lin_ok(res.time_emb_proj)
              and isinstance(res.norm1, nn.GroupNorm) and _gn_ok(res.norm1) and _gn_ok(res.norm2))
    except AttributeError:
        return False
    sc = getattr(res, "conv_shortcut", None)
    mods = [res.conv1, res.conv2, res.time_emb_proj] + ([sc] if sc is not None else [])
    return ok and (sc is None or _conv_ok(sc)) and len({m.dynamic for m in mods}) == 1 \
        and isinstance(getattr(res, "nonlinearity", None), nn.SiLU)


def _kv_key(blk):
    """attn2.to_k / to_v layers that may share one quantised context tensor and one GEMM: same
    context width, same BOS mode and — static scales — the same activation parameters"""
    k = blk.attn2.to_k
    return (k.in_features, bool(getattr(k, "bos", False)), _act_key(k))


def _temb_key(r):
    return (r.time_emb_proj.in_features, _act_key(r.time_emb_proj))


def fuse_unet(unet: nn.Module, verbose: bool = False) -> dict:
    """Re-bind the forwards of every eligible block of `unet` (in place). Returns a summary
    {"transformer_blocks": n, "transformer2d": n, "resnets": n, "kv_layers": n, "temb_layers": n}.
    Idempotent; blocks that are not eligible keep their stock forward."""
    if getattr(unet, "_mixdq_fused_summary", None) is not None:
        return unet._mixdq_fused_summary
    blocks = [m for m in unet.modules() if _is(m, "BasicTransformerBlock") and _block_ok(m)]
    resnets = [m for m in unet.modules() if _is(m, "ResnetBlock2D") and _resnet_ok(m)]
    t2ds = [m for m in unet.modules() if _is(m, "Transformer2DModel")
            and hasattr(m, "proj_in") and _lin_ok(m.proj_in) and _lin_ok(m.proj_out)
            and isinstance(m.norm, nn.GroupNorm) and _gn_ok(m.norm)
            and (_is_native(m) or _plain(m, is_input_continuous=True, use_linear_projection=True,
                                         is_input_vectorized=False, is_input_patches=False))]
    summary = {"transformer_blocks": len(blocks), "transformer2d": len(t2ds),
               "resnets": len(resnets), "kv_layers": 0, "temb_layers": 0}
    # one K/V group per (context width, BOS mode)
    kv_groups = {}
    for blk in blocks:
        key = _kv_key(blk)
        kv_groups.setdefault(key, []).extend([blk.attn2.to_k, blk.attn2.to_v])
    kv_objs = {k: SharedInputGroup(v, bos=k[1]) for k, v in kv_groups.items()}
    for blk in blocks:
        key = _kv_key(blk)
        proj = blk.ff.net[0].proj
        blk._mixdq_fused = {"orig_forward": blk.forward,
                            "qkv": CatLinear([blk.attn1.to_q, blk.attn1.to_k, blk.attn1.to_v]),
                            "kv": kv_objs[key],
                            "ffproj": GegluLinear(proj) if proj.out_features % 32 == 0 else None}
        blk.forward = types.MethodType(fused_transformer_block_forward, blk)
        summary["kv_layers"] += 2
    for m in t2ds:
        m._mixdq_fused = {"orig_forward": m.forward}
        m.forward = types.MethodType(fused_transformer2d_forward, m)
    if resnets:
        temb_groups = {}
        for r in resnets:
            temb_groups.setdefault(_temb_key(r), []).append(r.time_emb_proj)
        temb_objs = {k: SharedInputGroup(v, pre=F.silu) for k, v in temb_groups.items()}
        for r in resnets:
            r._mixdq_fused = {"temb": temb_objs[_temb_key(r)]}
            r.forward = types.MethodType(fused_resnet_forward, r)
            summary["temb_layers"] += 1
    # the shared-input results live for ONE UNet forward (see SharedInputGroup)
    groups = list(kv_objs.values()) + (list(temb_objs.values()) if resnets else [])
    unet._mixdq_shared_groups = groups

    def _reset_shared(_module, _args, _kwargs=None):
        for g in groups:
            g.reset()
    unet.register_forward_pre_hook(_reset_shared)
    unet._mixdq_fused_summary = summary
    if verbose:
        print(f"mixdq fuse_unet: {summary}")
    return summary
