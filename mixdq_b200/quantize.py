"""Module surgery: swap float nn.Linear / nn.Conv2d leaves for their quantized counterparts.

Keeps the call signature of the reference's kernels/quantize.py `convert` (:527-564) — itself a
fork of torch.ao.quantization.convert whose only functional additions are threading `ckpt` to
`from_float` and assigning a channel `split` to the up-block `conv_shortcut` layers (:631-648).
This is a compact re-implementation of exactly that behaviour:

  * every child that carries a non-None `.qconfig` and whose type is in `mapping` is replaced by
    `mapping[type].from_float(child, split=..., ckpt=ckpt)`; forward (pre-)hooks and device
    affinity are preserved (reference :650-668);
  * the split of a shortcut is taken from the float module's `.split` attribute when the caller
    (mixdq.register_qconfig_from_input_files) derived it from the architecture, otherwise from
    the reference's SDXL list in traversal order. The reference keeps the list position in a
    module-global counter that is never reset (a second convert() in one process walks off the
    list, :64,640-642); here the position is local to one convert() call.
"""
from __future__ import annotations

import copy
from typing import Dict, Optional

import torch
import torch.nn as nn

# SDXL-Turbo up-block shortcut splits in module-traversal order (reference kernels/quantize.py:61)
_SPLIT = [1280, 1280, 1280, 1280, 640, 640, 640, 320, 320]


def _default_mapping():
    from .nn.conv2d import QuantizedConv2d
    from .nn.linear import QuantizedLinear
    return {nn.Linear: QuantizedLinear, nn.Conv2d: QuantizedConv2d}


def derive_up_block_splits(unet: nn.Module) -> Dict[str, int]:
    """{module name: split} for every `up_blocks.*.resnets.*.conv_shortcut`.

    The split is the channel count of the hidden state before it is concatenated with the skip
    connection (reference quant_block_forward_func.py:96-102: `split = hidden_states.size(1)`):
    the previous resnet's output channels, or for the first resnet of a block the output channels
    of the previous up block (the mid block for up_blocks.0)."""
    splits: Dict[str, int] = {}
    up_blocks = getattr(unet, "up_blocks", None)
    if up_blocks is None:
        return splits

    def out_ch(resnet):
        return resnet.conv2.out_channels

    prev = None
    mid = getattr(unet, "mid_block", None)
    if mid is not None and hasattr(mid, "resnets"):
        prev = out_ch(mid.resnets[-1])
    elif hasattr(unet, "down_blocks"):
        prev = out_ch(unet.down_blocks[-1].resnets[-1])
    for b, block in enumerate(up_blocks):
        for i, resnet in enumerate(block.resnets):
            hidden = prev
            if getattr(resnet, "conv_shortcut", None) is not None and hidden is not None:
                splits[f"up_blocks.{b}.resnets.{i}.conv_shortcut"] = hidden
            prev = out_ch(resnet)
    return splits


class _ConvertState:
    def __init__(self):
        self.shortcut_idx = 0


def _split_for(mod, state: _ConvertState) -> int:
    name = getattr(mod, "module_name", "") or ""
    if "up_blocks" in name and "conv_shortcut" in name:
        if getattr(mod, "split", None) is not None:
            return int(mod.split)
        if state.shortcut_idx >= len(_SPLIT):
            raise RuntimeError("more up-block shortcuts than the SDXL split list covers; set "
                               "`.split` on the float modules (derive_up_block_splits)")
        s = _SPLIT[state.shortcut_idx]
        state.shortcut_idx += 1
        return s
    return 0


def swap_module(mod, mapping, custom_module_class_mapping=None, ckpt=None, _state=None):
    """Return the quantized counterpart of `mod` (or `mod` itself if it has none)."""
    state = _state or _ConvertState()
    if getattr(mod, "qconfig", None) is None or type(mod) not in mapping:
        return mod
    qcls = mapping[type(mod)]
    new_mod = qcls.from_float(mod, split=_split_for(mod, state), ckpt=ckpt)
    for hook in mod._forward_pre_hooks.values():
        new_mod.register_forward_pre_hook(hook)
    for hook in mod._forward_hooks.values():
        new_mod.register_forward_hook(hook)
    devices = {p.device for p in mod.parameters()} | {b.device for b in mod.buffers()}
    assert len(devices) <= 1, \
        f"swap_module only works with cpu or single-device CUDA modules, but got devices {devices}"
    if devices:
        new_mod.to(next(iter(devices)))
    return new_mod


def _convert(module, mapping, ckpt, state):
    for name, child in list(module.named_children()):
        _convert(child, mapping, ckpt, state)
        module._modules[name] = swap_module(child, mapping, None, ckpt=ckpt, _state=state)
    return module


def _remove_qconfig(module):
    for m in module.modules():
        if hasattr(m, "qconfig"):
            try:
                del m.qconfig
            except AttributeError:
                pass


def convert(module, mapping=None, inplace=False, remove_qconfig=True, is_reference=False,
            convert_custom_config_dict=None, ckpt=None):
    """Convert submodules of `module` according to `mapping` via `from_float`.

    Same parameters as the reference (kernels/quantize.py:527-529); `is_reference` and
    `convert_custom_config_dict` are accepted for signature compatibility and unused by the
    MixDQ flow."""
    if mapping is None:
        mapping = _default_mapping()
    if not inplace:
        module = copy.deepcopy(module)
    _convert(module, mapping, ckpt, _ConvertState())
    if remove_qconfig:
        _remove_qconfig(module)
    return module
