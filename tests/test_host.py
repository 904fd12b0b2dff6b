"""Host-side logic on CPU: from_float buffers against the reference's (golden), bit-config
registration, convert(), split derivation, the UNet skeleton's layer inventory."""
import json
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn
from torch.ao.quantization import PlaceholderObserver, QConfig

from mixdq_b200 import mixdq, quantize
from mixdq_b200.nn import QuantizedConv2d, QuantizedLinear
from mixdq_b200.nn import utils as U
from mixdq_b200.unet import build_unet


def _load(path):
    z = np.load(path, allow_pickle=False)
    return {k: torch.from_numpy(z[k]) if z[k].dtype.kind != "U" else str(z[k]) for k in z.files}


@pytest.fixture(scope="module")
def ref(golden_dir):
    return _load(golden_dir / "ref_from_float.npz")


def _ckpt(g):
    ck = {}
    for k in g:
        if k.startswith("ckpt."):
            name, field = k[5:].rsplit(".", 1)
            ck.setdefault(name, {})[field] = g[k]
    return ck


def _prep(mod, name, w_dtype=torch.qint8, a_dtype=torch.qint8, w_bit=8, a_bit=8):
    mod.qconfig = QConfig(activation=PlaceholderObserver.with_args(dtype=a_dtype),
                          weight=PlaceholderObserver.with_args(dtype=w_dtype))
    mod.module_name = name
    mod.w_bit = w_bit
    if a_bit is not None:
        mod.a_bit = a_bit
    return mod


def _float_mod(g, tag, cls, *a, **kw):
    m = cls(*a, **kw)
    with torch.no_grad():
        m.weight.copy_(g[tag + ".weight"])
        m.bias.copy_(g[tag + ".bias"])
    return m


@pytest.mark.parametrize("tag", ["linear", "conv_p1", "conv_p0", "conv_split"])
def test_from_float_state_dict_equals_reference(ref, tag):
    """Same buffer names and bit-identical values as the reference's from_float run on the
    shipped new_ckpt.pth (SURVEY §8(b) 'Module buffers')."""
    ck = _ckpt(ref)
    name = ref[tag + ".name"]
    if tag == "linear":
        fm = _float_mod(ref, tag, nn.Linear, 32, 1280)
        q = QuantizedLinear.from_float(_prep(fm, name), ckpt=ck)
    elif tag == "conv_p1":
        fm = _float_mod(ref, tag, nn.Conv2d, 4, 320, 3, padding=1)
        q = QuantizedConv2d.from_float(_prep(fm, name), ckpt=ck)
    elif tag == "conv_p0":
        fm = _float_mod(ref, tag, nn.Conv2d, 32, 640, 1)
        q = QuantizedConv2d.from_float(_prep(fm, name), ckpt=ck)
    else:
        fm = _float_mod(ref, tag, nn.Conv2d, 48, 320, 1)
        q = QuantizedConv2d.from_float(_prep(fm, name), split=int(ref[tag + ".split"]), ckpt=ck)
    assert q.valid_for_acceleration == bool(ref[tag + ".valid"])
    want = {k[len(tag) + 5:]: v for k, v in ref.items() if k.startswith(tag + ".buf.")}
    got = q.state_dict()
    assert set(got) == set(want)
    for k, v in want.items():
        assert got[k].dtype == v.dtype, k
        assert torch.equal(got[k], v), k
    assert q._get_name() in ("QuantizedLinearW8A8", "QuantizedConv2dW8A8")


def test_fp_fallback_gates(ref):
    ck = _ckpt(ref)
    name = ref["linear.name"]
    # activation left fp16 (layer absent from the act yaml) -> FP fallback, as the reference
    fm = _float_mod(ref, "linear", nn.Linear, 32, 1280)
    q = QuantizedLinear.from_float(_prep(fm, name, a_dtype=torch.float16, a_bit=None), ckpt=ck)
    assert not q.valid_for_acceleration and q._get_name() == "QuantizedLinearFPFallback"
    assert set(q.state_dict()) == {"weight", "bias"}
    x = torch.randn(3, 32)
    assert torch.equal(q(x), torch.nn.functional.linear(x, fm.weight, fm.bias))
    # 4-bit activations (N3): the reference gates them to fp16 (nn/Linear.py:28-36); here they
    # run on the int8 kernels with codes 0..15 and the UNSHIFTED 4-bit zero point of the ckpt
    q = QuantizedLinear.from_float(_prep(fm, name, a_dtype=torch.quint4x2, a_bit=4), ckpt=ck)
    assert q.valid_for_acceleration and q.a_bits == 4 and q._get_name() == "QuantizedLinearW8A4"
    zl = ck[name + ".act_quantizer"]["zero_point_list"]
    dl = ck[name + ".act_quantizer"]["delta_list"]
    assert float(q.act_zero_points) == float(zl[1]) and 0 <= float(q.act_zero_points) <= 15
    assert float(q.act_scales) == float(dl[1])
    assert torch.equal(q.bias0, q.weight_sum_by_input_channels * q.act_zero_points)
    # misaligned features -> warning + fallback (nn/Linear.py:37-43)
    odd = nn.Linear(30, 1280)
    q = QuantizedLinear.from_float(_prep(odd, name), ckpt=ck)
    assert not q.valid_for_acceleration


def test_w4_linear_packs_signed_nibbles(ref):
    ck = _ckpt(ref)
    name = ref["linear.name"]
    fm = _float_mod(ref, "linear", nn.Linear, 32, 1280)
    q = QuantizedLinear.from_float(_prep(fm, name, w_dtype=torch.quint4x2, w_bit=4), ckpt=ck)
    assert q.valid_for_acceleration and q._get_name() == "QuantizedLinearW4A8"
    assert q.weight_int4.shape == (1280, 16) and q.weight_int4.dtype == torch.uint8
    codes = U.unpack_int4(q.weight_int4)
    assert codes.min() >= -8 and codes.max() <= 7
    # 4-bit scales are row 1 of delta_list (bit_idx = log2(4)-1)
    assert torch.equal(q.weight_scales, ck[name + ".weight_quantizer"]["delta_list"][1].float())
    assert torch.equal(q.weight_sum_by_input_channels, codes.float().sum(1))


def test_get_quant_para_shift_and_index(ref):
    ck = _ckpt(ref)
    s, z, s0, z0 = U.get_quant_para(ck, 8, "conv_in", "act")
    assert s0 is None and z0 is None
    assert float(z) == float(ck["conv_in.act_quantizer"]["zero_point_list"][2]) - 128
    s4, z4, _, _ = U.get_quant_para(ck, 4, "conv_in", "act")
    assert float(s4) == float(ck["conv_in.act_quantizer"]["delta_list"][1])
    with pytest.raises(KeyError):
        U.get_quant_para(ck, 8, "no.such.layer", "weight")


def test_sdxl_skeleton_matches_reference_inventory(golden_dir):
    """794 quantizable leaves, names == the reference's YAML keys, Cout == ckpt entries."""
    cfg = json.loads((golden_dir / "bit_configs.json").read_text())
    summary = json.loads((golden_dir / "ckpt_summary.json").read_text())
    with torch.device("meta"):
        unet = build_unet("sdxl-turbo")
    layers = unet.quantizable_layers()
    assert sorted(n for n, _ in layers) == cfg["names"]
    assert sum(isinstance(m, nn.Linear) for _, m in layers) == 743
    assert sum(isinstance(m, nn.Conv2d) for _, m in layers) == 51
    for n, m in layers:
        assert summary[n + ".weight_quantizer"][0] == m.weight.shape[0]
    assert sum(m.weight.numel() for _, m in layers) == 2_565_857_280 or \
        abs(sum(m.weight.numel() for _, m in layers) / 1e9 - 2.566) < 2e-3
    # split list of the reference (kernels/quantize.py:61), in traversal order
    splits = quantize.derive_up_block_splits(unet)
    assert list(splits.values()) == quantize._SPLIT
    assert set(k + ".act_quantizer_0" for k in splits) == {k for k in summary if k.endswith("act_quantizer_0")}
    # BOS layers
    bos = json.loads((golden_dir / "bos_shapes.json").read_text())
    attn2_kv = [n for n, _ in layers if "attn2" in n and ("to_k" in n or "to_v" in n)]
    assert sorted(attn2_kv) == sorted(bos)
    mods = dict(layers)
    for n, shp in bos.items():
        assert shp == [1, 1, mods[n].weight.shape[0]]


def test_sd_turbo_skeleton_counts():
    with torch.device("meta"):
        unet = build_unet("sd-turbo")
    layers = unet.quantizable_layers()
    assert len(layers) == 282
    assert abs(sum(m.weight.numel() for _, m in layers) / 1e9 - 0.866) < 2e-3
    assert list(quantize.derive_up_block_splits(unet).values()) == \
        [1280] * 7 + [640, 640, 640, 320, 320]


def test_packaged_bit_configs():
    w8 = mixdq.load_bit_config("weight/weight_8.00.yaml")
    assert len(w8) == 794 and sorted(set(w8.values())) == [4, 8]
    assert sum(v == 4 for v in w8.values()) == 4
    a8 = mixdq.load_bit_config("./cfgs/act/act_8.00.yaml")
    assert len(a8) == 785 and "conv_in" not in a8 and "conv_out" not in a8
    w5 = mixdq.load_bit_config("weight/weight_5.02.yaml")
    assert (sum(v == 8 for v in w5.values()), sum(v == 4 for v in w5.values()),
            sum(v == 2 for v in w5.values())) == (401, 246, 147)
    with pytest.raises(FileNotFoundError):
        mixdq.load_bit_config("weight/nope.yaml")


def test_yaml_file_config_with_model_prefix(tmp_path):
    p = tmp_path / "w.yaml"
    p.write_text("model.conv_in: 8\nmodel.conv_out: 4\n")
    assert mixdq.load_bit_config(str(p)) == {"conv_in": 8, "conv_out": 4}


def _tiny_quantized(dynamic=True, a_config="uniform"):
    unet = build_unet("tiny", seed=1)
    names = [n for n, _ in unet.quantizable_layers()]
    args = SimpleNamespace(w_config={"model." + n: 8 for n in names},
                           a_config=None if a_config is None else {"model." + n: 8 for n in names})
    return unet, names, args


def test_quantize_unet_tiny_dynamic_cpu():
    unet, names, args = _tiny_quantized()
    mixdq.quantize_unet(unet, args, ckpt=None, bos=False, bos_dict=None)
    mods = dict(unet.named_modules())
    for n in names:
        assert isinstance(mods[n], (QuantizedLinear, QuantizedConv2d)), n
        assert mods[n].valid_for_acceleration, n
        assert not hasattr(mods[n], "qconfig")
    sc = mods["up_blocks.0.resnets.0.conv_shortcut"]
    assert sc.split == 128 and sc.weight_int.shape[1] == 128 and sc.weight_int_0.shape[1] == 128
    assert mods["up_blocks.1.resnets.1.conv_shortcut"].split == 64
    # weights are buffers, not parameters
    assert all(not isinstance(m, (nn.Linear, nn.Conv2d)) for m in unet.modules())
    # non-fp16 input takes the dequantised fallback and stays close to the float layer
    q = mods["time_embedding.linear_1"]
    x = torch.randn(2, q.in_features)
    ref_unet = build_unet("tiny", seed=1)
    want = dict(ref_unet.named_modules())["time_embedding.linear_1"](x)
    assert torch.allclose(q(x), want, atol=2e-2)


def test_register_qconfig_unknown_name_raises():
    unet, names, args = _tiny_quantized()
    args.w_config["model.not_a_layer"] = 8
    with pytest.raises(RuntimeError, match="weight yaml"):
        mixdq.register_qconfig_from_input_files(unet, args, bos=False, bos_dict=None)
    unet, names, args = _tiny_quantized()
    args.a_config["model.not_a_layer"] = 8
    with pytest.raises(RuntimeError, match="act yaml"):
        mixdq.register_qconfig_from_input_files(unet, args, bos=False, bos_dict=None)


def test_convert_twice_does_not_walk_off_split_list():
    """The reference keeps the split position in a never-reset global (kernels/quantize.py:64), so a
    second convert() in one process indexes past the 9-entry list; here every convert() starts
    fresh, and an architecture-derived `.split` attribute takes precedence over the list."""
    seen = []

    class Recorder(nn.Module):
        @classmethod
        def from_float(cls, mod, split=0, ckpt=None):
            seen.append((mod.module_name, split))
            return cls()

    def tree():
        root = nn.Module()
        root.up_blocks = nn.ModuleList()
        for b in range(3):
            blk = nn.Module()
            blk.resnets = nn.ModuleList()
            for i in range(3):
                r = nn.Module()
                r.conv_shortcut = nn.Conv2d(4, 4, 1)
                r.conv_shortcut.qconfig = object()
                r.conv_shortcut.module_name = f"up_blocks.{b}.resnets.{i}.conv_shortcut"
                blk.resnets.append(r)
            root.up_blocks.append(blk)
        return root

    for _ in range(2):
        seen.clear()
        quantize.convert(tree(), mapping={nn.Conv2d: Recorder}, inplace=True, remove_qconfig=False)
        assert [s for _, s in seen] == quantize._SPLIT
    t = tree()
    t.up_blocks[0].resnets[0].conv_shortcut.split = 96
    seen.clear()
    quantize.convert(t, mapping={nn.Conv2d: Recorder}, inplace=True, remove_qconfig=False)
    assert seen[0][1] == 96 and seen[1][1] == quantize._SPLIT[0]


def test_bos_registration():
    unet, names, args = _tiny_quantized()
    ehs = torch.randn(1, 77, unet.cfg.cross_attention_dim)
    bos_dict = mixdq.compute_bos_dict(unet, ehs)
    assert all(v.shape[:2] == (1, 1) for v in bos_dict.values()) and len(bos_dict) == 8
    mixdq.quantize_unet(unet, args, ckpt=None, bos=True, bos_dict=bos_dict)
    m = dict(unet.named_modules())["mid_block.attentions.0.transformer_blocks.0.attn2.to_k"]
    assert m.bos is True and "bos_pre_computed" in m.state_dict()


def test_node_mappings_match_reference_workflow():
    assert set(mixdq.NODE_CLASS_MAPPINGS) == {"Mixdq", "LoadPipe", "OrgGen", "MixdqIntegral"}
    assert mixdq.Mixdq.RETURN_TYPES == ("IMAGE", "STRING") and mixdq.Mixdq.FUNCTION == "mixdq_quant"
    assert "org_pipeline" in mixdq.Mixdq.INPUT_TYPES()["required"]
    import mixdq_extension.op.qconv2d as qc
    import mixdq_extension.op.qlinear as ql
    import mixdq_extension.op.quant as qq
    assert callable(qc.qconv2d) and callable(ql.qlinear) and callable(qq.quantize_per_tensor)


def test_quantized_unet_file_round_trip(tmp_path):
    """N2: the quantised model (int8 / packed-int4 codes, scales, sums, BOS rows, stored layout) is
    written to disk and rebuilt on a META skeleton without any float weight; every buffer,
    attribute and the reference-format state_dict survive the round trip."""
    from mixdq_b200 import serialize
    from mixdq_b200.unet import UNet2DConditionModel, tiny_config
    unet, names, args = _tiny_quantized()
    # mixed precision: a few W4 layers, one 4-bit-activation layer, BOS rows on the K/V layers
    for i, n in enumerate(names):
        if "ff.net" in n or "attn1.to_q" in n:
            args.w_config["model." + n] = 4
        if n.endswith("attn2.to_out.0"):
            args.a_config["model." + n] = 4
    ehs = torch.randn(1, 77, unet.cfg.cross_attention_dim)
    bos_dict = mixdq.compute_bos_dict(unet, ehs)
    mixdq.quantize_unet(unet, args, ckpt=None, bos=True, bos_dict=bos_dict)
    path = tmp_path / "tiny_quantized.pt"
    info = serialize.save_quantized_unet(unet, path, meta={"model": "tiny", "w": "mixed"})
    assert info["leaves"] == len(names) and info["quantized_bytes"] > 0
    with torch.device("meta"):
        skel = UNet2DConditionModel(tiny_config())
    loaded = serialize.load_quantized_unet(skel, path, "cpu")
    a, b = dict(unet.named_modules()), dict(loaded.named_modules())
    kinds = set()
    for n in names:
        assert type(a[n]) is type(b[n]) and a[n]._get_name() == b[n]._get_name(), n
        kinds.add(a[n]._get_name())
        assert set(a[n]._buffers) == set(b[n]._buffers), n
        for k, t in a[n]._buffers.items():
            if t is not None:
                assert torch.equal(t, b[n]._buffers[k]) and t.dtype == b[n]._buffers[k].dtype, (n, k)
        for k in ("dynamic", "a_bits", "w_kind", "w_bits", "split", "bos", "valid_for_acceleration",
                  "in_features", "out_features", "kernel_size", "stride", "padding"):
            assert getattr(a[n], k, None) == getattr(b[n], k, None), (n, k)
    assert {"QuantizedLinearW4A8", "QuantizedLinearW8A4", "QuantizedLinearW8A8",
            "QuantizedConv2dW8A8"} <= kinds
    sa, sb = unet.state_dict(), loaded.state_dict()
    assert set(sa) == set(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    assert not any(p.is_meta for p in loaded.parameters()) and not any(x.is_meta for x in loaded.buffers())
    # a file of another format / version is refused
    with pytest.raises(ValueError):
        serialize.load_quantized_unet(skel, {"format": "something else"}, "cpu")


def test_repo_root_is_a_comfyui_plugin():
    """repo-root __init__.py re-exports the node mappings like the reference's (__init__.py:1-3)"""
    import importlib.util
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("mixdq_plugin_root", root / "__init__.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert set(mod.NODE_CLASS_MAPPINGS) == {"Mixdq", "LoadPipe", "OrgGen", "MixdqIntegral"}
    assert mod.__all__ == ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]


def test_shipping_config_on_the_sdxl_skeleton():
    """weight_8.00.yaml + act_8.00.yaml (the ComfyUI node's default, kernels/mixdq.py:560-570) on the
    SDXL-Turbo skeleton: 794 weight entries of which 4 are 4-bit, 785 activation entries — the 9
    layers missing from the act config keep fp16 activations and therefore run as FP fallbacks
    (conv_in, conv_out, down_blocks.0.resnets.0.conv2, up_blocks.2.resnets.2
    .conv_shortcut and five ff.net.2 layers)."""
    from mixdq_b200.unet import UNet2DConditionModel, sdxl_turbo_config
    with torch.device("meta"):
        unet = UNet2DConditionModel(sdxl_turbo_config())
    args = SimpleNamespace(w_config="weight/weight_8.00.yaml", a_config="act/act_8.00.yaml")
    mixdq.register_qconfig_from_input_files(unet, args, bos=False, bos_dict=None)
    leaves = dict(unet.quantizable_layers())
    assert len(leaves) == 794
    w4 = sorted(n for n, m in leaves.items() if m.qconfig.weight().dtype == torch.quint4x2)
    fp_act = sorted(n for n, m in leaves.items() if m.qconfig.activation().dtype == torch.float16)
    assert len(w4) == 4 and all(leaves[n].w_bit == 4 for n in w4)
    assert len(fp_act) == 9 and not any(hasattr(leaves[n], "a_bit") for n in fp_act)
    expect = {"conv_in", "conv_out", "down_blocks.0.resnets.0.conv2",
              "up_blocks.2.resnets.2.conv_shortcut",
              "up_blocks.0.attentions.0.transformer_blocks.0.ff.net.2"}
    assert expect <= set(fp_act)
    assert sum(1 for n in fp_act if n.startswith("down_blocks.2.attentions.1.transformer_blocks")
               and n.endswith("ff.net.2")) == 4
    # the 9 up-block shortcuts carry their channel split (kernels/quantize.py:61)
    splits = [leaves[n].split for n in leaves if "up_blocks" in n and n.endswith("conv_shortcut")]
    assert splits == [1280, 1280, 1280, 1280, 640, 640, 640, 320, 320]
    # from_float gates: a protected layer becomes an FP fallback, a 4-bit layer W4A8 (checked on
    # one small layer of each kind, materialised on the CPU)
    for name, want in ((fp_act[0], "FPFallback"), (w4[0], "W4A8")):
        src = leaves[name]
        if isinstance(src, nn.Linear):
            real = nn.Linear(src.in_features, src.out_features, bias=src.bias is not None).half()
            qcls = QuantizedLinear
        else:
            real = nn.Conv2d(src.in_channels, src.out_channels, src.kernel_size, src.stride,
                             src.padding, bias=src.bias is not None).half()
            qcls = QuantizedConv2d
        for k in ("qconfig", "module_name", "w_bit", "a_bit", "split"):
            if hasattr(src, k):
                setattr(real, k, getattr(src, k))
        q = qcls.from_float(real, split=getattr(src, "split", 0) or 0, ckpt=None)
        assert q._get_name().endswith(want), (name, q._get_name())


def test_ptq_running_statistics_match_reference_quantizers(golden_dir):
    """N4: mixdq_b200.ptq restates the calibration arithmetic of the reference's BaseQuantizer
    (running min / max with momentum 0.95 updated once per bit width per forward, delta, zero
    point; per-channel symmetric weight scales) — pinned on vectors produced by the reference's own
    quantizers driven like scripts/ptq.py:126-155 (oracle/make_golden.py::golden_ptq)."""
    import numpy as np
    from mixdq_b200 import ptq
    z = np.load(golden_dir / "ptq_running_stat.npz")
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    for name, split in (("linear", 0), ("conv", 0), ("split", 8)):
        w = g[f"{name}.weight"]
        n = int(g[f"{name}.n_batches"])
        stats = [ptq.ActRunningStat(), ptq.ActRunningStat()]
        for i in range(n):
            x = g[f"{name}.x{i}"]
            parts = (x[:, :split], x[:, split:]) if split else (x,)
            for st, part in zip(stats, parts):
                st.observe(*ptq.tensor_minmax(part))
        sfxs = ("", "_0") if split else ("",)
        wparts = (w[:, :split], w[:, split:]) if split else (w,)
        for st, sfx, wp in zip(stats, sfxs, wparts):
            assert torch.equal(st.delta_list, g[f"{name}.a_delta_list{sfx}"]), (name, sfx)
            assert torch.equal(st.zero_point_list, g[f"{name}.a_zp_list{sfx}"]), (name, sfx)
            assert torch.equal(ptq.weight_delta_list(wp), g[f"{name}.w_delta_list{sfx}"]), (name, sfx)


def test_ptq_calibrate_produces_a_kernel_format_checkpoint():
    """calibrate() on the tiny UNet (CPU, float): every layer gets the entries `from_float` looks
    up (incl. the `_0` twins of the split shortcuts), in the dtype / shapes of new_ckpt.pth
    (kernels/convert_ckpt.py:22-46), and the static quantisation it drives runs."""
    from mixdq_b200 import ptq
    unet = build_unet("tiny", seed=5)
    names = [n for n, _ in unet.quantizable_layers()]
    batches = [unet.example_inputs(2, "cpu", torch.float32, seed=s) for s in (1, 2, 3)]
    ck = ptq.calibrate(unet, batches)
    splits = quantize.derive_up_block_splits(unet)
    for n in names:
        cout = dict(unet.named_modules())[n].weight.shape[0]
        for key in [n + ".weight_quantizer", n + ".act_quantizer"] + \
                ([n + ".weight_quantizer_0", n + ".act_quantizer_0"] if n in splits else []):
            e = ck[key]
            assert e["delta_list"].dtype == torch.float16 and e["zero_point_list"].dtype == torch.float16
            want = (3, cout) if "weight" in key else (3,)
            assert tuple(e["delta_list"].shape) == want and tuple(e["zero_point_list"].shape) == want
        a = ck[n + ".act_quantizer"]
        assert (a["delta_list"] > 0).all() and (a["zero_point_list"][2] >= 0) and (a["zero_point_list"][2] <= 255)
    assert len(ck) == 2 * len(names) + 2 * len(splits)
    args = SimpleNamespace(w_config={"model." + n: 8 for n in names},
                           a_config={"model." + n: 8 for n in names})
    mixdq.quantize_unet(unet.half(), args, ckpt=ck, bos=False, bos_dict=None)
    q = dict(unet.named_modules())["mid_block.attentions.0.transformer_blocks.0.attn1.to_q"]
    assert q.valid_for_acceleration and not q.dynamic
    assert float(q.act_scales) == float(ck[q.module_name + ".act_quantizer"]["delta_list"][2])
