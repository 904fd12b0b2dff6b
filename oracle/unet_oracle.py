"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (see the header of qdiff_oracle.py for the rules).

Whole-UNet fake-quant driver: wraps every nn.Linear / nn.Conv2d leaf of a float UNet in an
`OracleQuantLayer`, the restatement of the reference's qdiff `QuantLayer`
(quant_utils/qdiff/models/quant_layer.py:14-115) driven the way scripts/quant_txt2img.py:175-196
drives it: quantizers initialise from the first forward, i.e. dynamic per-tensor min-max
activation quantisation and per-channel min-max weight quantisation of that input.

Restated model-level semantics (the reference implements them in diffusers-dependent files that
cannot be imported here):
  * split shortcuts — up-block `conv_shortcut` layers quantise input channels [0:split) and
    [split:) independently, split = hidden-state channels before the skip concat
    (quant_block_forward_func.py:96-102, quant_block.py:163-166, quant_layer.py:63-89);
  * BOS — cross-attention `to_k` / `to_v` pass the first text token through the un-quantised
    weight and quantise only tokens [1:] (quant_block.py:585-625);
  * protected layers — layers without an activation bit-width run fully un-quantised
    (quant_txt2img.py:223-226, quant_model.py:268-278; kernel path: nn/Linear.py:155-156).
PARITY UNPINNED for the whole-UNet output: no reference artefact pins it; the leaf
arithmetic it is built from is pinned by tests/test_oracle_golden.py.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import qdiff_oracle as O


class OracleQuantLayer(nn.Module):
    def __init__(self, org: nn.Module, name: str, w_bits: int = 8, a_bits: Optional[int] = 8,
                 split: int = 0, bos: bool = False, static_act=None, static_act_0=None):
        super().__init__()
        self.name = name
        self.weight = org.weight.detach().float()
        self.bias = None if org.bias is None else org.bias.detach().float()
        self.is_conv = isinstance(org, nn.Conv2d)
        self.stride = org.stride[0] if self.is_conv else 1
        self.padding = org.padding[0] if self.is_conv else 0
        self.w_bits, self.a_bits, self.split, self.bos = w_bits, a_bits, split, bos
        self.static_act, self.static_act_0 = static_act, static_act_0
        self.record = None   # optional dict filled with the last call's qparams (for tests)

    def _fq(self, x):
        return O.fake_quant_layer(x, self.weight, self.bias, self.w_bits, self.a_bits, self.split,
                                  self.stride, self.padding, self.static_act, self.static_act_0)

    def forward(self, x):
        x = x.float()
        if self.a_bits is None:      # protected: no weight, no activation quantisation
            if self.is_conv:
                return F.conv2d(x, self.weight, self.bias, stride=self.stride, padding=self.padding)
            return F.linear(x, self.weight, self.bias)
        if self.bos:
            first = F.linear(x[:, :1, :], self.weight, self.bias)
            return torch.cat([first, self._fq(x[:, 1:, :])], dim=1)
        return self._fq(x)


def wrap_unet(unet: nn.Module, w_bits: Dict[str, int], a_bits: Dict[str, int],
              splits: Dict[str, int], bos: bool = False) -> nn.Module:
    """Replace leaves in place (the UNet must be an fp32 CPU copy)."""
    for name, mod in list(unet.named_modules()):
        if not isinstance(mod, (nn.Linear, nn.Conv2d)) or name not in w_bits:
            continue
        wb = w_bits[name]
        wb = 4 if wb == 2 else wb
        is_bos = bos and "attn2" in name and ("to_k" in name or "to_v" in name)
        layer = OracleQuantLayer(mod, name, wb, a_bits.get(name), splits.get(name, 0), is_bos)
        parent_name, _, leaf = name.rpartition(".")
        parent = unet.get_submodule(parent_name) if parent_name else unet
        parent._modules[leaf] = layer
    return unet
