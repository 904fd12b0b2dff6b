"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own Python.

TEST INFRASTRUCTURE. Runs only in the build container (needs /root/reference, which does not
exist on the GPU box); the fixtures it writes are committed and are what the tests read.

  python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]

What is executed, unmodified, from the reference tree:
  * quant_utils/qdiff/models/quant_layer.py  QuantLayer  (+ quantizer/base_quantizer.py) —
    imported through a namespace stub that skips qdiff/__init__.py (it eagerly imports diffusers,
    absent here);
  * kernels/mixdq_extension/nn/{Linear,Conv2d,utils}.py  from_float — with `mixdq_extension._C`
    stubbed (the CUDA extension is only touched in forward(), never in from_float);
  * kernels/output/new_ckpt.pth, kernels/bos_pre_computed.pt, kernels/cfgs/**.yaml (data);
  * torch.quantize_per_tensor as used by the reference's op self-test (op/quant.py:7-30).
"""
from __future__ import annotations

import argparse
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import yaml


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def import_qdiff(ref: Path):
    pkg = types.ModuleType("qdiff")
    pkg.__path__ = [str(ref / "quant_utils" / "qdiff")]
    sys.modules["qdiff"] = pkg
    from qdiff.models.quant_layer import QuantLayer  # noqa
    return QuantLayer


def import_ref_nn(ref: Path):
    pkg = types.ModuleType("mixdq_extension")
    pkg.__path__ = [str(ref / "kernels" / "mixdq_extension")]
    sys.modules["mixdq_extension"] = pkg
    stub = types.ModuleType("mixdq_extension._C")
    for name in ("quantize_per_tensor_to_int8", "quantize_per_tensor_to_int8_vectorized",
                 "qlinear_w8_a8_ohalf", "qlinear_fp_reference", "qconv2d_w8_a8_ohalf"):
        setattr(stub, name, None)
    sys.modules["mixdq_extension._C"] = stub
    pkg._C = stub
    from mixdq_extension.nn.Linear import QuantizedLinear  # noqa
    from mixdq_extension.nn.Conv2d import QuantizedConv2d  # noqa
    return QuantizedLinear, QuantizedConv2d


def wq_cfg(n_bits):
    return AttrDict(n_bits=n_bits, sym=True, channel_wise=True, scale_method="min_max",
                    round_mode="nearest", mixed_precision=[2, 4, 8])


def aq_cfg(n_bits):
    return AttrDict(n_bits=n_bits, channel_wise=False, scale_method="min_max",
                    round_mode="nearest_ste", running_stat=True, mixed_precision=[2, 4, 8])


def run_quant_layer(QuantLayer, mod, x, w_bits, a_bits, split=0):
    layer = QuantLayer(mod, wq_cfg(w_bits), aq_cfg(a_bits))
    for q in (layer.weight_quantizer, layer.act_quantizer):
        q.module_name = "golden"
    layer.set_quant_state(True, True)
    with torch.no_grad():
        if split:
            y = layer(x, split=split)
        else:
            y = layer(x)
    out = {"y": y}
    out["w_delta_list"] = layer.weight_quantizer.delta_list.reshape(3, -1)
    out["a_delta_list"] = layer.act_quantizer.delta_list.reshape(3)
    out["a_zp_list"] = layer.act_quantizer.zero_point_list.reshape(3)
    if split:
        out["w_delta_list_0"] = layer.weight_quantizer_0.delta_list.reshape(3, -1)
        out["a_delta_list_0"] = layer.act_quantizer_0.delta_list.reshape(3)
        out["a_zp_list_0"] = layer.act_quantizer_0.zero_point_list.reshape(3)
    return out


def golden_qdiff(QuantLayer, out: Path):
    g = torch.Generator().manual_seed(1234)
    cases = {}
    # linear
    lin = nn.Linear(64, 48)
    x = torch.randn(2, 5, 64, generator=g) * 1.3 + 0.2
    for wb in (8, 4):
        r = run_quant_layer(QuantLayer, lin, x, wb, 8)
        cases[f"linear_w{wb}a8"] = dict(x=x, weight=lin.weight.detach(), bias=lin.bias.detach(), **r)
    # conv 3x3 p1
    conv = nn.Conv2d(16, 24, 3, padding=1)
    xc = torch.randn(2, 16, 9, 9, generator=g)
    r = run_quant_layer(QuantLayer, conv, xc, 8, 8)
    cases["conv3x3_w8a8"] = dict(x=xc, weight=conv.weight.detach(), bias=conv.bias.detach(), **r)
    # conv 3x3 stride 2
    conv2 = nn.Conv2d(16, 8, 3, stride=2, padding=1)
    r = run_quant_layer(QuantLayer, conv2, xc, 8, 8)
    cases["conv3x3s2_w8a8"] = dict(x=xc, weight=conv2.weight.detach(), bias=conv2.bias.detach(), **r)
    # 1x1 split shortcut
    sc = nn.Conv2d(24, 16, 1)
    xs = torch.cat([torch.randn(2, 8, 6, 6, generator=g) * 3.0,
                    torch.randn(2, 16, 6, 6, generator=g) * 0.5 + 1.0], dim=1)
    r = run_quant_layer(QuantLayer, sc, xs, 8, 8, split=8)
    cases["conv1x1_split8_w8a8"] = dict(x=xs, weight=sc.weight.detach(), bias=sc.bias.detach(), **r)
    flat = {}
    for cname, d in cases.items():
        for k, v in d.items():
            flat[f"{cname}.{k}"] = v.detach().numpy()
    np.savez_compressed(out / "qdiff_quant_layer.npz", **flat)
    print("qdiff cases:", list(cases))


def golden_from_float(ref: Path, out: Path):
    from torch.ao.quantization import QConfig, PlaceholderObserver
    QuantizedLinear, QuantizedConv2d = import_ref_nn(ref)
    ckpt = torch.load(ref / "kernels" / "output" / "new_ckpt.pth", map_location="cpu")
    g = torch.Generator().manual_seed(4321)

    def prep(mod, name, w_bit=8, a_bit=8):
        mod.qconfig = QConfig(activation=PlaceholderObserver.with_args(dtype=torch.qint8),
                              weight=PlaceholderObserver.with_args(dtype=torch.qint8))
        mod.module_name = name
        mod.w_bit = w_bit
        mod.a_bit = a_bit
        return mod

    flat = {}

    def dump(tag, float_mod, qmod):
        flat[f"{tag}.weight"] = float_mod.weight.detach().numpy()
        if float_mod.bias is not None:
            flat[f"{tag}.bias"] = float_mod.bias.detach().numpy()
        for k, v in qmod.state_dict().items():
            flat[f"{tag}.buf.{k}"] = v.numpy()
        flat[f"{tag}.valid"] = np.array(qmod.valid_for_acceleration)

    # Linear with real ckpt scales (N=1280), small synthetic K
    lin = nn.Linear(32, 1280)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(1280, 32, generator=g) * 0.05)
    name = "time_embedding.linear_1"
    q = QuantizedLinear.from_float(prep(lin, name), ckpt=ckpt)
    dump("linear", lin, q)
    flat["linear.name"] = np.array(name)
    # conv_in: real shape 4 -> 320, 3x3 p1
    conv = nn.Conv2d(4, 320, 3, padding=1)
    name = "conv_in"
    q = QuantizedConv2d.from_float(prep(conv, name), ckpt=ckpt)
    dump("conv_p1", conv, q)
    flat["conv_p1.name"] = np.array(name)
    # 1x1 shortcut without split
    c1 = nn.Conv2d(32, 640, 1)
    name = "down_blocks.1.resnets.0.conv_shortcut"
    q = QuantizedConv2d.from_float(prep(c1, name), ckpt=ckpt)
    dump("conv_p0", c1, q)
    flat["conv_p0.name"] = np.array(name)
    # split shortcut (Cout=320), small synthetic Cin = 48, split 16
    cs = nn.Conv2d(48, 320, 1)
    name = "up_blocks.2.resnets.1.conv_shortcut"
    q = QuantizedConv2d.from_float(prep(cs, name), split=16, ckpt=ckpt)
    dump("conv_split", cs, q)
    flat["conv_split.name"] = np.array(name)
    flat["conv_split.split"] = np.array(16)
    # the raw ckpt entries those modules read (so tests need not ship the 19 MB checkpoint)
    keys = ["time_embedding.linear_1", "conv_in", "down_blocks.1.resnets.0.conv_shortcut",
            "up_blocks.2.resnets.1.conv_shortcut"]
    for k in keys:
        for suffix in (".weight_quantizer", ".act_quantizer", ".weight_quantizer_0",
                       ".act_quantizer_0"):
            if k + suffix in ckpt:
                for f in ("delta_list", "zero_point_list"):
                    flat[f"ckpt.{k}{suffix}.{f}"] = ckpt[k + suffix][f].numpy()
    np.savez_compressed(out / "ref_from_float.npz", **flat)

    # 8-bit activation / weight-scale summary of the whole checkpoint (pins the reader + skeleton)
    summary = {}
    for k, v in ckpt.items():
        if k.endswith(".act_quantizer") or k.endswith(".act_quantizer_0"):
            summary[k] = [float(v["delta_list"][2]), float(v["zero_point_list"][2])]
        else:
            summary[k] = [int(v["delta_list"].shape[1])]
    (out / "ckpt_summary.json").write_text(json.dumps(summary, indent=0, sort_keys=True))
    print("ckpt entries:", len(ckpt))


def golden_known_answer(out: Path):
    """op/quant.py:7-30 — the reference's own quantize self-test target."""
    g = torch.Generator().manual_seed(7)
    t = torch.rand(1024, generator=g).half()
    tf = t.float()
    zero_point = torch.round((tf.max() + tf.min()) / 2)
    scale = (tf.max() - tf.min()) / 255
    ref = torch.quantize_per_tensor(tf, scale, zero_point, torch.qint8).int_repr()
    # a second, wider case with negative values and saturation
    t2 = (torch.randn(4096, generator=g) * 3).half()
    scale2 = torch.tensor(0.0323)
    zp2 = torch.tensor(2.0)
    ref2 = torch.quantize_per_tensor(t2.float(), scale2, zp2, torch.qint8).int_repr()
    # per-channel weights
    w = torch.randn(24, 40, generator=g) * 0.1
    ws = w.abs().amax(dim=1) / 127
    wq = torch.quantize_per_channel(w, ws, torch.zeros(24), 0, torch.qint8).int_repr()
    np.savez_compressed(out / "torch_quantize_known_answer.npz",
                        x=t.numpy(), scale=scale.numpy(), zp=zero_point.numpy(), q=ref.numpy(),
                        x2=t2.numpy(), scale2=scale2.numpy(), zp2=zp2.numpy(), q2=ref2.numpy(),
                        w=w.numpy(), ws=ws.numpy(), wq=wq.numpy())


def golden_configs(ref: Path, out: Path):
    cfg = {}
    for sub in ("weight", "act"):
        for f in sorted((ref / "kernels" / "cfgs" / sub).glob("*.yaml")):
            d = yaml.safe_load(f.read_text())
            cfg[f"{sub}/{f.name}"] = {k.replace("model.", "", 1): v for k, v in d.items()}
    names = sorted(cfg["weight/uniform_8.yaml"])
    compact = {"names": names, "configs": {}}
    for k, d in cfg.items():
        compact["configs"][k] = [d.get(n, 0) for n in names]   # 0 = layer absent from the file
    (out / "bit_configs.json").write_text(json.dumps(compact))
    bos = torch.load(ref / "kernels" / "bos_pre_computed.pt", map_location="cpu")
    (out / "bos_shapes.json").write_text(json.dumps({k: list(v.shape) for k, v in bos.items()},
                                                    sort_keys=True))
    print("configs:", list(cfg), "bos entries:", len(bos))


def golden_ptq(QuantLayer, out: Path):
    """The calibration flow of scripts/ptq.py:126-155 on single layers: weight quantizers
    initialised from the weights (one forward with weight_quant only), then the activation
    quantizers' running statistics over several calibration batches (momentum 0.95,
    base_quantizer.py:41,160-171 — updated once per bit width of `mixed_precision` per forward)."""
    g = torch.Generator().manual_seed(4321)
    flat = {}

    def calibrate(name, mod, batches, split=0):
        layer = QuantLayer(mod, wq_cfg(8), aq_cfg(8))
        qs = [layer.weight_quantizer, layer.act_quantizer]
        if split:
            # the '_0' twins are created by the first split forward
            pass
        for q in qs:
            q.module_name = "golden"
        kw = dict(split=split) if split else {}
        with torch.no_grad():
            layer.set_quant_state(True, False)
            layer(batches[0], **kw)
            layer.weight_quantizer.init_done = True
            if split:
                layer.weight_quantizer_0.init_done = True
            layer.set_quant_state(True, True)
            for xb in batches:
                layer(xb, **kw)
        flat[f"{name}.weight"] = mod.weight.detach().numpy()
        for i, xb in enumerate(batches):
            flat[f"{name}.x{i}"] = xb.numpy()
        flat[f"{name}.n_batches"] = np.array(len(batches))
        flat[f"{name}.w_delta_list"] = layer.weight_quantizer.delta_list.reshape(3, -1).numpy()
        flat[f"{name}.a_delta_list"] = layer.act_quantizer.delta_list.reshape(3).numpy()
        flat[f"{name}.a_zp_list"] = layer.act_quantizer.zero_point_list.reshape(3).numpy()
        if split:
            flat[f"{name}.w_delta_list_0"] = layer.weight_quantizer_0.delta_list.reshape(3, -1).numpy()
            flat[f"{name}.a_delta_list_0"] = layer.act_quantizer_0.delta_list.reshape(3).numpy()
            flat[f"{name}.a_zp_list_0"] = layer.act_quantizer_0.zero_point_list.reshape(3).numpy()

    lin = nn.Linear(64, 48)
    calibrate("linear", lin, [torch.randn(2, 7, 64, generator=g) * (1.0 + 0.3 * i) + 0.1 * i
                              for i in range(4)])
    conv = nn.Conv2d(16, 24, 3, padding=1)
    calibrate("conv", conv, [torch.randn(2, 16, 9, 9, generator=g) * (0.8 + 0.2 * i) for i in range(3)])
    sc = nn.Conv2d(24, 16, 1)
    calibrate("split", sc, [torch.cat([torch.randn(2, 8, 6, 6, generator=g) * (2.0 + i),
                                       torch.randn(2, 16, 6, 6, generator=g) * 0.5 + 1.0], dim=1)
                            for i in range(3)], split=8)
    np.savez_compressed(out / "ptq_running_stat.npz", **flat)
    print("ptq cases: linear, conv, split")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=str(Path(__file__).resolve().parent.parent / "tests" / "golden"))
    a = ap.parse_args()
    ref, out = Path(a.ref), Path(a.out)
    out.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)
    QuantLayer = import_qdiff(ref)
    golden_qdiff(QuantLayer, out)
    golden_ptq(QuantLayer, out)
    golden_known_answer(out)
    golden_configs(ref, out)
    golden_from_float(ref, out)
    print("wrote", sorted(p.name for p in out.iterdir()))


if __name__ == "__main__":
    main()
