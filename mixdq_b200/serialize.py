"""N2 — on-disk format of a QUANTISED UNet.

The reference keeps only the PTQ parameters on disk (`kernels/convert_ckpt.py:22-46` -> new_ckpt.pth:
delta / zero_point lists per layer) and re-quantises every weight from the fp16 UNet at every load
(`nn/Linear.py:116-132`, `nn/Conv2d.py:157-242`): the fp16 model (5.0 GB for SDXL) has to exist on
the GPU first. Here the quantised model itself is serialised:

  file = torch.save({
      "format": "mixdq_b200.quantized_unet", "version": 1, "meta": {...caller's...},
      "leaves":  {module name: {"kind": "linear" | "conv2d", "attrs": {plain attributes}}},
      "tensors": {module name + "." + buffer: tensor},      # int8 codes, PACKED int4 (two codes
                                                            # per byte, even k in the high nibble,
                                                            # nn/utils.py:26-28), fp32 scales /
                                                            # sums, fp16 bias, BOS rows
      "other":   {state_dict entries of everything that is not a quantised leaf},
  })

Buffers are written in their STORED layout: a GEGLU projection whose rows were interleaved for the
fused epilogue (`geglu_interleaved`) stays interleaved in the file — one copy of every weight, in
the layout the kernels consume. `load_quantized_unet` rebuilds the modules on a skeleton created on
the META device: no fp16 weight is ever materialised, the GPU holds the 2.5 GB of int8 / int4
buffers (SDXL, all W8) and nothing else (tests/test_gpu_realsize.py checks the resident bytes).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .nn.conv2d import QuantizedConv2d
from .nn.linear import QuantizedLinear

FORMAT = "mixdq_b200.quantized_unet"
VERSION = 1
_PLAIN = (int, float, bool, str, type(None))


def _plain_attr(v) -> bool:
    if isinstance(v, _PLAIN):
        return True
    return isinstance(v, tuple) and all(isinstance(e, _PLAIN) for e in v)


def _leaf_record(m: nn.Module) -> dict:
    attrs = {k: v for k, v in m.__dict__.items()
             if not k.startswith("_") and k != "training" and _plain_attr(v)}
    return {"kind": "linear" if isinstance(m, QuantizedLinear) else "conv2d", "attrs": attrs}


def quantized_state(unet: nn.Module, meta: Optional[dict] = None) -> dict:
    """The serialisable description of a quantised UNet (CPU tensors, no aliasing)."""
    leaves: Dict[str, dict] = {}
    tensors: Dict[str, torch.Tensor] = {}
    for name, m in unet.named_modules():
        if isinstance(m, (QuantizedLinear, QuantizedConv2d)):
            leaves[name] = _leaf_record(m)
            # non-persistent buffers that are operands (not derived indices) travel too, and come
            # back non-persistent, so the reference-format state_dict is unchanged
            leaves[name]["non_persistent"] = sorted(
                b for b in m._non_persistent_buffers_set if b != "geglu_inverse_index")
            for bname, buf in m._buffers.items():
                if buf is None or bname == "geglu_inverse_index":
                    continue
                # members of an N-concatenated group are views of one storage: store each leaf's
                # own rows (clone), the loader re-concatenates when it fuses
                t = buf.detach().to("cpu")
                if buf.dim() == 4 and buf.is_contiguous(memory_format=torch.channels_last):
                    t = t.contiguous(memory_format=torch.channels_last)
                else:
                    t = t.contiguous()
                tensors[f"{name}.{bname}"] = t.clone() if t.data_ptr() == buf.data_ptr() else t
    prefixes = tuple(n + "." for n in leaves)
    other = {k: v.detach().to("cpu").clone() for k, v in unet.state_dict().items()
             if not k.startswith(prefixes)}
    return {"format": FORMAT, "version": VERSION, "meta": dict(meta or {}), "leaves": leaves,
            "tensors": tensors, "other": other}


def save_quantized_unet(unet: nn.Module, path, meta: Optional[dict] = None) -> dict:
    """Write `unet` (after `quantize_unet`, fused or not) to `path`. Returns a size summary."""
    state = quantized_state(unet, meta)
    torch.save(state, str(path))
    nbytes = lambda d: sum(t.numel() * t.element_size() for t in d.values())
    return {"leaves": len(state["leaves"]), "quantized_bytes": nbytes(state["tensors"]),
            "other_bytes": nbytes(state["other"])}


def _build_leaf(rec: dict, tensors: Dict[str, torch.Tensor], prefix: str, device) -> nn.Module:
    cls = QuantizedLinear if rec["kind"] == "linear" else QuantizedConv2d
    m = cls.__new__(cls)
    nn.Module.__init__(m)
    for k, v in rec["attrs"].items():
        setattr(m, k, v)
    m.device = device
    for key, t in tensors.items():
        if key.startswith(prefix) and "." not in key[len(prefix):]:
            bname = key[len(prefix):]
            m.register_buffer(bname, t.to(device),
                              persistent=bname not in rec.get("non_persistent", ()))
    if getattr(m, "geglu_interleaved", False):
        from . import ops
        idx = ops.geglu_interleave_index(m.out_features // 2, device)
        m.register_buffer("geglu_inverse_index", torch.argsort(idx), persistent=False)
    return m


def _set_submodule(root: nn.Module, name: str, new: nn.Module) -> None:
    parent = root
    parts = name.split(".")
    for p in parts[:-1]:
        parent = getattr(parent, p)
    parent._modules[parts[-1]] = new


def load_quantized_unet(skeleton: nn.Module, path_or_state, device, fuse: Optional[bool] = None
                        ) -> nn.Module:
    """Turn `skeleton` — the float architecture, ideally built under `torch.device("meta")` so that
    no weight exists yet — into the quantised model stored at `path_or_state`, on `device`.

    Every quantised leaf is rebuilt from its stored buffers (no `from_float`, no fp16 weight);
    the remaining parameters / buffers are taken from the file. `fuse` as in `quantize_unet`
    (default: dynamic mode on a CUDA device)."""
    state = path_or_state if isinstance(path_or_state, dict) else \
        torch.load(str(path_or_state), map_location="cpu", weights_only=False)
    if state.get("format") != FORMAT:
        raise ValueError("not a mixdq_b200 quantised-UNet file")
    if state.get("version") != VERSION:
        raise ValueError(f"unsupported file version {state.get('version')} (this build reads {VERSION})")
    device = torch.device(device)
    modules = dict(skeleton.named_modules())
    missing = [n for n in state["leaves"] if n not in modules]
    if missing:
        raise RuntimeError(f"{len(missing)} stored layers do not exist in the skeleton, e.g. {missing[0]}")
    for name, rec in state["leaves"].items():
        _set_submodule(skeleton, name, _build_leaf(rec, state["tensors"], name + ".", device))
    # everything else: materialise from the file (parameters stay parameters)
    other = state["other"]
    for mname, mod in skeleton.named_modules():
        if isinstance(mod, (QuantizedLinear, QuantizedConv2d)):
            continue
        for pname, p in list(mod._parameters.items()):
            if p is None:
                continue
            key = f"{mname}.{pname}" if mname else pname
            if key not in other:
                raise RuntimeError(f"parameter {key} is missing from the file")
            mod._parameters[pname] = nn.Parameter(other[key].to(device), requires_grad=False)
        for bname, b in list(mod._buffers.items()):
            if b is None or bname in mod._non_persistent_buffers_set:
                continue
            key = f"{mname}.{bname}" if mname else bname
            if key in other:
                mod._buffers[bname] = other[key].to(device)
    if fuse is None:
        fuse = device.type == "cuda"
    if fuse:
        from .fused import fuse_unet
        fuse_unet(skeleton)
    return skeleton
