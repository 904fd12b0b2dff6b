"""Model-level API: bit-config registration, `quantize_unet`, whole-UNet CUDA graph, ComfyUI nodes.

Mirrors the reference's kernels/mixdq.py (register_qconfig_from_input_files :43-143,
convert_to_quantized :146-154, quantize_unet :158-160, cuda_graph_opt :188-290,
NODE_CLASS_MAPPINGS :779-791) — same function names, argument meaning and error behaviour — on
top of the sm_100a kernels. The diffusers pipeline itself stays external (imported lazily by the
ComfyUI nodes); everything below works on any UNet whose leaf names follow the diffusers naming
(e.g. mixdq_b200.unet.UNet2DConditionModel).
"""
from __future__ import annotations

import functools
import json
import threading
from pathlib import Path
from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn
from torch.ao.quantization import PlaceholderObserver, QConfig

from . import ops
from .nn.conv2d import QuantizedConv2d
from .nn.linear import QuantizedLinear
from .quantize import convert, derive_up_block_splits

_CFG_DB = Path(__file__).resolve().parent / "cfgs" / "bit_configs.json"

# bits -> storage dtype tag of the qconfig (reference kernels/mixdq.py:49-53; 2 bit is treated as 4)
bw_to_dtype = {8: torch.qint8, 4: torch.quint4x2, 2: torch.quint4x2}


def nvtx_decorator(forward_func, name=None):
    """Wrap a module forward in an NVTX range (reference kernels/mixdq.py:18-33)."""
    def wrapper(self, *args, **kwargs):
        name_ = name if name is not None else f"Forward {self.__class__.__name__}"
        torch.cuda.nvtx.range_push(name_)
        try:
            return forward_func(self, *args, **kwargs)
        finally:
            torch.cuda.nvtx.range_pop()
    return wrapper


def _strip_prefix(name: str) -> str:
    pos = name.find("model.")
    return name[pos + 6:] if pos >= 0 else name


def load_bit_config(spec) -> Dict[str, int]:
    """`spec` is a YAML file in the reference format (`model.<layer>: bits` per line), a dict, or
    the id of a packaged MixDQ config such as "weight/weight_8.00.yaml" / "act/act_8.00.yaml"
    (the reference's kernels/cfgs/**, shipped as one compact JSON)."""
    if isinstance(spec, dict):
        raw = spec
    else:
        p = Path(str(spec))
        if p.exists():
            import yaml
            raw = yaml.safe_load(p.read_text())
        else:
            db = json.loads(_CFG_DB.read_text())
            key = "/".join(p.parts[-2:])
            if key not in db["configs"]:
                raise FileNotFoundError(f"bit config {spec} not found (packaged: "
                                        f"{sorted(db['configs'])})")
            raw = {n: b for n, b in zip(db["names"], db["configs"][key]) if b}
    return {_strip_prefix(k): int(v) for k, v in raw.items()}


def register_qconfig_from_input_files(unet, args, bos, bos_dict):
    """Attach `.qconfig`, `.module_name`, `.w_bit`, `.a_bit`, `.bos*` (and `.split`) to the float
    leaves named in `args.w_config` / `args.a_config`. A name in a config that maps to no module
    raises RuntimeError, like the reference (:97-101, :139-143)."""
    w_bits = load_bit_config(args.w_config)
    modules = dict(unet.named_modules())
    splits = derive_up_block_splits(unet)

    missing = [n for n in w_bits if n not in modules]
    if missing:
        for n in missing:
            print(f"{n} not found in UNet!")
        raise RuntimeError("Not all keys in weight yaml map to a module in UNet.")
    for name, bits in w_bits.items():
        mod = modules[name]
        assert not hasattr(mod, "qconfig") or mod.qconfig is None
        mod.qconfig = QConfig(activation=PlaceholderObserver.with_args(dtype=torch.float16),
                              weight=PlaceholderObserver.with_args(dtype=bw_to_dtype[bits]))
        mod.module_name = name
        mod.w_bit = bits
        if name in splits:
            mod.split = splits[name]
        if "attn2" in name and ("to_k" in name or "to_v" in name):
            mod.bos = bos
            mod.bos_pre_computed = bos_dict[name] if bos_dict is not None else None

    if getattr(args, "a_config", None) is None:
        return
    a_bits = load_bit_config(args.a_config)
    missing = [n for n in a_bits if n not in modules]
    if missing:
        for n in missing:
            print(f"{n} not found in UNet!")
        raise RuntimeError("Not all keys in act yaml map to a module in UNet.")
    for name, bits in a_bits.items():
        mod = modules[name]
        act = PlaceholderObserver.with_args(dtype=bw_to_dtype[bits])
        if getattr(mod, "qconfig", None):
            mod.qconfig = QConfig(weight=mod.qconfig.weight, activation=act)
        else:
            mod.qconfig = QConfig(activation=act,
                                  weight=PlaceholderObserver.with_args(dtype=torch.float16))
            mod.module_name = name
        mod.a_bit = bits


def convert_to_quantized(unet, ckpt):
    convert(unet, mapping={nn.Linear: QuantizedLinear, nn.Conv2d: QuantizedConv2d},
            inplace=True, ckpt=ckpt)


def quantize_unet(unet, args, ckpt, bos, bos_dict, fuse: Optional[bool] = None):
    """Quantize `unet` in place. `ckpt` is the PTQ checkpoint dict in the kernel format
    (reference kernels/convert_ckpt.py:22-46) or None for dynamic activation quantisation with
    min-max weight scales.

    `fuse` (extension; the reference signature ends at `bos_dict`): also re-bind the block
    forwards to the fused kernels (mixdq_b200.fused.fuse_unet), for dynamic AND static (ckpt)
    activation scales. Default: when the model already sits on a CUDA device. The module tree,
    names and buffers of the quantized leaves are the same either way."""
    register_qconfig_from_input_files(unet, args, bos=bos, bos_dict=bos_dict)
    convert_to_quantized(unet, ckpt)
    if fuse is None:
        p = next(iter(unet.buffers()), None)
        fuse = p is not None and p.device.type == "cuda"
    if fuse:
        from .fused import fuse_unet
        fuse_unet(unet)
    return unet


def compute_bos_dict(unet, encoder_hidden_states) -> Dict[str, torch.Tensor]:
    """`bos_pre_computed[name]` = fp16 K/V projection of the first (BOS) text token, [1,1,Cout].
    The reference ships these for the real SDXL-Turbo weights (kernels/bos_pre_computed.pt); with
    other weights they are recomputed from the float layers."""
    out = {}
    first = encoder_hidden_states[:1, :1, :]
    with torch.no_grad():
        for name, mod in unet.named_modules():
            if isinstance(mod, nn.Linear) and "attn2" in name and ("to_k" in name or "to_v" in name):
                out[name] = torch.nn.functional.linear(first.to(mod.weight.dtype), mod.weight,
                                                       mod.bias).detach()
    return out


# ---------------------------------------------------------------------------------------------
# whole-UNet CUDA graph (reference kernels/mixdq.py:188-290)
# ---------------------------------------------------------------------------------------------
def _sig(arg):
    if isinstance(arg, torch.Tensor):
        scalar = arg.item() if arg.device.type == "cpu" and arg.numel() == 1 else None
        return (arg.device.type, arg.device.index, arg.dtype, tuple(arg.shape), scalar)
    if isinstance(arg, (str, int, float, bytes, bool)):
        return arg
    if isinstance(arg, (tuple, list)):
        return tuple(_sig(a) for a in arg)
    if isinstance(arg, dict):
        return tuple(sorted(((_sig(k), _sig(v)) for k, v in arg.items()), key=lambda kv: str(kv[0])))
    return type(arg)


def _clone(arg):
    if isinstance(arg, torch.Tensor):
        return arg.detach().clone(memory_format=torch.preserve_format)
    if isinstance(arg, tuple):
        return tuple(_clone(a) for a in arg)
    if isinstance(arg, list):
        return [_clone(a) for a in arg]
    if isinstance(arg, dict):
        return {k: _clone(v) for k, v in arg.items()}
    if arg is None or isinstance(arg, (str, int, float, bytes, bool)):
        return arg
    raise ValueError(f"Unknown argument type {arg}")


def _copy_into(dst, src):
    if isinstance(src, torch.Tensor):
        dst.copy_(src)
    elif isinstance(src, (tuple, list)):
        for d, s in zip(dst, src):
            _copy_into(d, s)
    elif isinstance(src, dict):
        for k, v in src.items():
            _copy_into(dst[k], v)


def _first_cuda_device(obj) -> torch.device:
    if isinstance(obj, torch.Tensor):
        return obj.device if obj.device.type == "cuda" else None
    if isinstance(obj, (tuple, list)):
        for o in obj:
            d = _first_cuda_device(o)
            if d is not None:
                return d
    if isinstance(obj, dict):
        return _first_cuda_device(list(obj.values()))
    return None


def _invalidate_host_caches(unet) -> None:
    ops.clear_dynamic_quant_cache()
    for g in getattr(unet, "_mixdq_shared_groups", ()):
        g.reset()


def cuda_graph_opt(unet, args=None, warmup: int = 3):
    """Replace `unet.forward` by a version that captures one CUDA graph per argument signature
    and replays it: inputs are copied into static buffers, the graph is replayed, the static
    output is returned. Thread-safe capture, as the reference."""
    lock = threading.Lock()
    cache = {}
    wrapped = unet.forward

    @functools.wraps(wrapped)
    def forward_with_cuda_graph(*a, **kw):
        key = (_sig(a), _sig(kw))
        if key not in cache:
            with lock:
                if key not in cache:
                    sa, skw = _clone((a, kw))
                    dev = _first_cuda_device((sa, skw))
                    # warm-up AND capture run on a stream of this graph's own, whose scratch
                    # buffers (split-K exchange, dynamic-quantisation workspace) are allocated
                    # before the capture begins — outside the graph's private memory pool, and
                    # not shared with any other graph or stream (replays may run concurrently)
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.no_grad(), torch.cuda.stream(side):
                        ops.prepare_stream(dev)
                        for _ in range(warmup):
                            wrapped(*sa, **skw)
                    torch.cuda.current_stream(dev).wait_stream(side)
                    # nothing computed eagerly may be reused inside the capture: its kernels
                    # would be missing from the graph (the caches also key on the capture id)
                    _invalidate_host_caches(unet)
                    graph = torch.cuda.CUDAGraph()
                    with torch.no_grad(), torch.cuda.graph(graph, stream=side):
                        static_out = wrapped(*sa, **skw)
                    _invalidate_host_caches(unet)
                    cache[key] = ((sa, skw), graph, static_out, side)
        (sa, skw), graph, static_out, _ = cache[key]
        _copy_into((sa, skw), (a, kw))
        graph.replay()
        return static_out

    forward_with_cuda_graph.__self__ = unet
    forward_with_cuda_graph._cached = cache
    unet.forward = forward_with_cuda_graph
    return unet


# ---------------------------------------------------------------------------------------------
# ComfyUI nodes — same node ids / I/O types as the reference plugin (kernels/mixdq.py:536-791)
# ---------------------------------------------------------------------------------------------
_DEFAULT_QUERY = "A cinematic shot of a baby racoon wearing an intricate italian priest robe."


def _require_diffusers():
    try:
        from diffusers import StableDiffusionXLPipeline  # noqa: F401
        return StableDiffusionXLPipeline
    except Exception as e:  # pragma: no cover - diffusers is an external dependency
        raise RuntimeError("the ComfyUI nodes need `diffusers` for the SDXL pipeline "
                           "(text encoders, scheduler, VAE); only the UNet hot path lives in "
                           "mixdq_b200") from e


def _node_args(weight_mode: str, act_mode: str):
    w = "weight/weight_8.00.yaml" if weight_mode.startswith("W8") else "weight/weight_5.02.yaml"
    a = None if act_mode.startswith("None") else (
        "act/act_8.00.yaml" if act_mode.startswith("W8") else "act/act_7.84.yaml")
    return SimpleNamespace(w_config=w, a_config=a, bos=False)


def _run_pipeline(pipeline, query):
    import time
    t0 = time.perf_counter()
    image = pipeline(prompt=[query], guidance_scale=0.0, num_inference_steps=1,
                     output_type="pil").images[0]
    dt = time.perf_counter() - t0
    import numpy as np
    arr = torch.from_numpy(np.asarray(image.convert("RGB")).astype("float32") / 255.0)[None]
    mem = torch.cuda.max_memory_allocated() / 2 ** 20
    return arr, f"cost time: {dt:.3f} s, peak memory: {mem:.1f} MB"


class load_modelpipeline:
    @classmethod
    def INPUT_TYPES(cls):
        return {"required": {"model_path": ("STRING", {"default": "stabilityai/sdxl-turbo"})}}
    RETURN_TYPES = ("PIPELINE",)
    RETURN_NAMES = ("org_pipeline",)
    FUNCTION = "load"
    CATEGORY = "MixDQ"

    def load(self, model_path):
        cls = _require_diffusers()
        return (cls.from_pretrained(model_path, torch_dtype=torch.float16, variant="fp16"),)


class OriginGen:
    @classmethod
    def INPUT_TYPES(cls):
        return {"required": {"org_pipeline": ("PIPELINE",),
                             "query": ("STRING", {"default": _DEFAULT_QUERY, "multiline": True})}}
    RETURN_TYPES = ("IMAGE", "STRING",)
    RETURN_NAMES = ("org_image", "org_efficiency",)
    FUNCTION = "generate"
    CATEGORY = "MixDQ"

    def generate(self, org_pipeline, query):
        org_pipeline.to("cuda")
        return _run_pipeline(org_pipeline, query)


class Mixdq:
    @classmethod
    def INPUT_TYPES(cls):
        return {"required": {
            "org_pipeline": ("PIPELINE",),
            "query": ("STRING", {"default": _DEFAULT_QUERY, "multiline": True}),
            "weight_mode": (["W8-bit(Recommended)", "W5.02-bit(W4A8 kernels)"],),
            "act_mode": (["W8-bit(Recommended)", "W7.84-bit(Unsupported)", "None(Unsupported)"],)}}
    RETURN_TYPES = ("IMAGE", "STRING",)
    RETURN_NAMES = ("quant_image", "quant_efficiency",)
    FUNCTION = "mixdq_quant"
    CATEGORY = "MixDQ"
    ckpt_path = "./custom_nodes/MixDQ/kernels/output/new_ckpt.pth"
    bos_path = "./custom_nodes/MixDQ/kernels/bos_pre_computed.pt"

    def mixdq_quant(self, org_pipeline, query, weight_mode, act_mode):
        torch.cuda.empty_cache()
        args = _node_args(weight_mode, act_mode)
        ckpt = torch.load(self.ckpt_path, map_location="cpu")
        bos_dict = torch.load(self.bos_path, map_location="cpu")
        quantize_unet(org_pipeline.unet, args, ckpt, args.bos, bos_dict)
        org_pipeline.to("cuda")
        return _run_pipeline(org_pipeline, query)


class MixdqIntegral:
    @classmethod
    def INPUT_TYPES(cls):
        d = Mixdq.INPUT_TYPES()
        return d
    RETURN_TYPES = ("IMAGE", "IMAGE", "STRING", "STRING",)
    RETURN_NAMES = ("quant_image", "org_image", "quant_efficiency", "org_efficiency",)
    FUNCTION = "run_both"
    CATEGORY = "MixDQ"

    def run_both(self, org_pipeline, query, weight_mode, act_mode):
        org_pipeline.to("cuda")
        org_img, org_txt = _run_pipeline(org_pipeline, query)
        q_img, q_txt = Mixdq().mixdq_quant(org_pipeline, query, weight_mode, act_mode)
        return (q_img, org_img, q_txt, org_txt)


NODE_CLASS_MAPPINGS = {
    "Mixdq": Mixdq,
    "LoadPipe": load_modelpipeline,
    "OrgGen": OriginGen,
    "MixdqIntegral": MixdqIntegral,
}

NODE_DISPLAY_NAME_MAPPINGS = {
    "Mixdq": "MixdqQuant",
    "LoadPipe": "LoadPipeline",
    "SDXL-Turbo": "OrgGen",
    "MixdqIntegral": "MixdqIntegral",
}
