"""The oracle (oracle/qdiff_oracle.py) against fixtures produced by the REFERENCE's own Python
(oracle/make_golden.py): qdiff QuantLayer outputs, reference from_float buffers, and the
torch.quantize_* known answers of the reference's op self-tests. CPU only."""
import json

import numpy as np
import pytest
import torch

from oracle import qdiff_oracle as O


def _load(path):
    z = np.load(path, allow_pickle=False)
    return {k: torch.from_numpy(z[k]) if z[k].dtype.kind != "U" else str(z[k]) for k in z.files}


@pytest.fixture(scope="module")
def qdiff(golden_dir):
    return _load(golden_dir / "qdiff_quant_layer.npz")


CASES = [("linear_w8a8", 8, 0, 1, 0), ("linear_w4a8", 4, 0, 1, 0), ("conv3x3_w8a8", 8, 0, 1, 1),
         ("conv3x3s2_w8a8", 8, 0, 2, 1), ("conv1x1_split8_w8a8", 8, 8, 1, 0)]


@pytest.mark.parametrize("case,w_bits,split,stride,pad", CASES)
def test_fake_quant_layer_matches_reference_quantlayer(qdiff, case, w_bits, split, stride, pad):
    g = {k.split(".", 1)[1]: v for k, v in qdiff.items() if k.startswith(case + ".")}
    y = O.fake_quant_layer(g["x"], g["weight"], g["bias"], w_bits=w_bits, a_bits=8, split=split,
                           stride=stride, padding=pad)
    # same fp32 operations in the same order (bit-identical on the generating CPU; the
    # tolerance only absorbs a different BLAS summation order on another host)
    assert torch.allclose(y, g["y"], rtol=1e-5, atol=2e-6), (y - g["y"]).abs().max()


@pytest.mark.parametrize("case,w_bits,split,stride,pad", CASES)
def test_qparams_match_reference_lists(qdiff, case, w_bits, split, stride, pad):
    g = {k.split(".", 1)[1]: v for k, v in qdiff.items() if k.startswith(case + ".")}
    halves = [(g["x"], g["weight"], "")] if not split else [
        (g["x"][:, :split], g["weight"][:, :split], ""), (g["x"][:, split:], g["weight"][:, split:], "_0")]
    for x, w, sfx in halves:
        for idx, bits in enumerate((2, 4, 8)):
            d, z = O.act_qparams_minmax(x, bits)
            assert d.item() == g["a_delta_list" + sfx][idx].item()
            assert z.item() == g["a_zp_list" + sfx][idx].item()
            wd = O.weight_qparams_minmax(w, bits)
            assert torch.equal(wd, g["w_delta_list" + sfx][idx])


@pytest.mark.parametrize("case,w_bits,split,stride,pad", CASES)
def test_integer_identity_reproduces_fake_quant(qdiff, case, w_bits, split, stride, pad):
    """(acc - z*wsum) * s_a*s_w + b on the integer codes == the fake-quant fp32 output, to fp32
    rounding; after fp16 rounding inside the north star's tolerance (1e-2 / cos 0.9999)."""
    g = {k.split(".", 1)[1]: v for k, v in qdiff.items() if k.startswith(case + ".")}
    x, w, b = g["x"], g["weight"], g["bias"]
    conv = w.dim() == 4
    halves = [(x, w)] if not split else [(x[:, :split], w[:, :split]), (x[:, split:], w[:, split:])]
    total = None
    for i, (xh, wh) in enumerate(halves):
        q, s_a, zp = O.quantize_dynamic_kernel(xh)
        s_w = O.weight_qparams_minmax(wh, w_bits)
        w_int = O.weight_fake_quant(wh, s_w, w_bits)[0].to(torch.int8)
        scale = s_w * s_a
        bias = b if i == 0 else None
        if conv:
            wsum = w_int.float().sum(dim=1)
            if pad:
                out, _ = O.qconv2d_kernel(q, w_int, scale, wsum, None, zp, bias, stride, pad)
            else:
                out, _ = O.qconv2d_kernel(q, w_int, scale, None, wsum.sum(dim=[1, 2]) * zp, zp, bias,
                                          stride, pad)
        else:
            wsum = w_int.float().sum(dim=1)
            out, _ = O.qlinear_kernel(q, w_int, wsum * zp, scale, bias)
        total = out if total is None else O.split_shortcut_kernel(total, out)
    ref = g["y"]
    err = (total.float() - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(total.float().flatten(), ref.flatten(), dim=0).item()
    assert err <= 1e-2 and cos >= 0.9999, (err, cos)


def test_static_quantize_formula_vs_torch_quantize_per_tensor(golden_dir):
    """Reference op self-test (op/quant.py:7-30): the kernel formula against
    torch.quantize_per_tensor. They agree except on fp32 near-ties of x/scale."""
    g = _load(golden_dir / "torch_quantize_known_answer.npz")
    q = O.quantize_static_kernel(g["x"], 1.0 / g["scale"].item(), g["zp"].item())
    assert torch.equal(q, g["q"])
    q2 = O.quantize_static_kernel(g["x2"], (1.0 / g["scale2"]).item(), g["zp2"].item())
    assert (q2 != g["q2"]).sum().item() <= 2
    assert (q2.int() - g["q2"].int()).abs().max().item() <= 1
    wq = O.quantize_weight_per_channel(g["w"], g["ws"])
    assert torch.equal(wq, g["wq"])


def test_from_float_buffers_match_reference(golden_dir):
    g = _load(golden_dir / "ref_from_float.npz")

    def ckpt_entry(name, sfx):
        return {"delta_list": g[f"ckpt.{name}{sfx}.delta_list"],
                "zero_point_list": g[f"ckpt.{name}{sfx}.zero_point_list"]}

    def ckpt_for(name, split=False):
        c = {name + ".weight_quantizer": ckpt_entry(name, ".weight_quantizer"),
             name + ".act_quantizer": ckpt_entry(name, ".act_quantizer")}
        if split:
            c[name + ".weight_quantizer_0"] = ckpt_entry(name, ".weight_quantizer_0")
            c[name + ".act_quantizer_0"] = ckpt_entry(name, ".act_quantizer_0")
        return c

    for tag, pad in (("linear", 0), ("conv_p1", 1), ("conv_p0", 0)):
        name = g[tag + ".name"]
        ck = ckpt_for(name)
        ws, wz = O.ckpt_qparams(ck, name, "weight", 8)
        a_s, a_z = O.ckpt_qparams(ck, name, "act", 8)
        assert torch.equal(ws, g[tag + ".buf.weight_scales"])
        assert torch.equal(a_s, g[tag + ".buf.act_scales"])
        assert torch.equal(a_z, g[tag + ".buf.act_zero_points"])
        bufs = O.from_float_buffers(g[tag + ".weight"], ws, a_s, a_z, padding=pad)
        for k, v in bufs.items():
            assert torch.equal(v, g[f"{tag}.buf.{k}"]), (tag, k)
    # split shortcut: both halves
    name = g["conv_split.name"]
    split = int(g["conv_split.split"])
    ck = ckpt_for(name, split=True)
    w = g["conv_split.weight"]
    for sfx, wh in (("", w[:, :split]), ("_0", w[:, split:])):
        ws, _ = O.ckpt_qparams(ck, name, "weight", 8, sfx)
        a_s, a_z = O.ckpt_qparams(ck, name, "act", 8, sfx)
        bufs = O.from_float_buffers(wh, ws, a_s, a_z, padding=0)
        assert torch.equal(bufs["weight_int"], g["conv_split.buf.weight_int" + sfx])
        assert torch.equal(bufs["scale"], g["conv_split.buf.scale" + sfx])
        assert torch.equal(bufs["bias0"], g["conv_split.buf.bias0" + sfx])


def test_ckpt_summary_known_values(golden_dir):
    """conv_in 8-bit activation (delta, zp) = (0.0323, 130) in the shipped ckpt."""
    s = json.loads((golden_dir / "ckpt_summary.json").read_text())
    d, z = s["conv_in.act_quantizer"]
    assert abs(d - 0.0323) < 1e-4 and z == 130.0
    assert sum(k.endswith(".weight_quantizer") for k in s) == 794
    assert sum(k.endswith("_0") for k in s) == 18


def test_int4_pack_roundtrip():
    codes = torch.randint(-8, 8, (7, 64), dtype=torch.int8)
    p = O.pack_int4(codes)
    assert p.shape == (7, 32) and p.dtype == torch.uint8
    assert torch.equal(O.unpack_int4(p), codes)
    # even index -> high nibble
    assert int(p[0, 0]) == ((int(codes[0, 0]) & 0xF) << 4 | (int(codes[0, 1]) & 0xF))
