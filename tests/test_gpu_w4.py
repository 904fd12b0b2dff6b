"""W4A8 on the tensor cores (north star "W4A8 layers unpacked on the fly"): packed 4-bit weights go
through TMA as they are and are expanded to int8 inside tc_i8_kernel<..., W4 = true>. The reference
has no 4-bit kernel (nn/Linear.py:28-36 gates 4-bit layers to fp16), so the oracle is the integer
identity of op/qlinear.py:66-83 / op/qconv2d.py:65-99 on the UNPACKED codes — INT32 accumulators
and fp16 outputs bit-exact — and the qdiff 4-bit fake-quant path at the north-star tolerance.
Nibble order: even k in the high nibble (nn/utils.py:26-28)."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn
from torch.ao.quantization import PlaceholderObserver, QConfig

from oracle import qdiff_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from mixdq_b200 import build
    build.build()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from mixdq_b200 import ops as _ops
    return _ops


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int16)


def _w4_linear(ops, dev, M, N, K, bias=True, dynamic=False, residual=False, seed=0):
    from mixdq_b200 import _lib
    g = torch.Generator().manual_seed(seed * 131 + M + N + K)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g)
    codes = torch.randint(-8, 8, (N, K), dtype=torch.int8, generator=g)
    packed = O.pack_int4(codes)
    w_scale = 0.001 + 0.01 * torch.rand(N, generator=g)
    a_scale, a_zp = torch.tensor(0.0371), torch.tensor(-11.0)
    wsum = codes.float().sum(1)
    b = torch.randn(N, generator=g).half() if bias else None
    res = torch.randn(M, N, generator=g).half() if residual else None
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    if dynamic:
        out = ops.qlinear_dynamic_fused(a.to(dev), packed.to(dev), w_scale.to(dev), a_scale.to(dev),
                                        a_zp.to(dev), wsum.to(dev), None if b is None else b.to(dev),
                                        None if res is None else res.to(dev), _acc_out=acc)
    else:
        out = ops.qlinear_w4_a8_ohalf(a.to(dev), packed.to(dev), (w_scale * a_scale).to(dev),
                                      (wsum * a_zp).to(dev), None if b is None else b.to(dev),
                                      _acc_out=acc)
    torch.cuda.synchronize()
    path = _lib.last_path()
    ref, ref_acc = O.qlinear_kernel(a, codes, wsum * a_zp, w_scale * a_scale, b)
    if res is not None:
        ref = (ref.float() + res.float()).half()
    assert torch.equal(acc.cpu().long(), ref_acc), "INT32 accumulators differ"
    assert torch.equal(bits(out), bits(ref)), "fp16 outputs differ"
    return path


# every SDXL linear shape with K % 32 == 0 + ragged M / N / K tails
W4_SHAPES = [(256, 1280, 1280), (256, 10240, 1280), (256, 1280, 5120), (1024, 5120, 640),
             (1024, 640, 640), (1024, 640, 2560), (77, 1280, 2048), (77, 640, 2048),
             (1, 1280, 1280), (1, 320, 1280), (1, 1280, 2816), (1, 1280, 320),
             (300, 200, 352), (129, 264, 64), (128, 16, 32), (5, 8, 32)]


@pytest.mark.parametrize("M,N,K", W4_SHAPES)
def test_w4_linear_runs_on_tcgen05_bit_exact(ops, dev, M, N, K):
    assert _w4_linear(ops, dev, M, N, K, bias=(M % 2 == 0)).startswith("tcgen05-w4")


@pytest.mark.parametrize("bn,splits", [(16, 1), (16, 2), (32, 1), (32, 4), (64, 1), (64, 2), (64, 8),
                                       (128, 1), (128, 4), (128, 8), (256, 1), (256, 2), (256, 8)])
def test_w4_linear_every_tile_width_and_split(ops, dev, bn, splits):
    """each BN instantiation (one / two k-blocks per ring stage, one / two MMA issuers) x split-K
    cluster size, incl. M / N / K tails and K ranges longer than the ring (phase wrap-around)"""
    from mixdq_b200 import _lib
    lib = _lib.load()
    lib.mixdq_debug_force_bn(bn)
    lib.mixdq_debug_force_splits(splits)
    try:
        want = "tcgen05-w4-splitk" if splits > 1 else "tcgen05-w4"
        assert _w4_linear(ops, dev, 260, 328, 1280) == want
        assert _w4_linear(ops, dev, 128, 512, 5120, bias=False, dynamic=True) == want
        if splits <= 4:
            assert _w4_linear(ops, dev, 77, 640, 416 + 96) == want
    finally:
        lib.mixdq_debug_force_bn(0)
        lib.mixdq_debug_force_splits(0)


def test_w4_linear_dynamic_with_residual(ops, dev):
    assert _w4_linear(ops, dev, 256, 1280, 5120, dynamic=True, residual=True).startswith("tcgen05-w4")
    assert _w4_linear(ops, dev, 1024, 640, 640, dynamic=True, residual=True, bias=False).startswith("tcgen05-w4")


def test_w4_linear_extreme_codes_do_not_overflow(ops, dev):
    """accumulators are carried as 16x their value inside the kernel: K = 5120 at the extreme codes
    (a = -128, w = -8) is 5120 * 128 * 128 = 8.4e7 << 2^31"""
    M, N, K = 128, 64, 5120
    a = torch.full((M, K), -128, dtype=torch.int8)
    codes = torch.full((N, K), -8, dtype=torch.int8)
    codes[1::2] = 7
    one, zero = torch.ones(N), torch.zeros(N)
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    ops.qlinear_w4_a8_ohalf(a.to(dev), O.pack_int4(codes).to(dev), (one * 1e-6).to(dev), zero.to(dev),
                            None, _acc_out=acc)
    want = a.long() @ codes.long().t()
    assert torch.equal(acc.cpu().long(), want)


def test_w4_linear_simt_fallback_still_matches(ops, dev):
    """misaligned activation pitch -> portable kernel, same arithmetic"""
    from mixdq_b200 import _lib
    lib = _lib.load()
    lib.mixdq_force_simt(1)
    try:
        assert _w4_linear(ops, dev, 33, 24, 64) == "simt-w4"
    finally:
        lib.mixdq_force_simt(0)


W4_CONVS = [  # (n,h,w,c,k,r,s,pad,stride): SDXL resnet convs that are 4 bit in weight_5.02.yaml
    (1, 64, 64, 320, 320, 3, 3, 1, 1), (1, 32, 32, 640, 640, 3, 3, 1, 1),
    (1, 16, 16, 1280, 1280, 3, 3, 1, 1), (1, 16, 16, 2560, 1280, 3, 3, 1, 1),
    (1, 32, 32, 960, 640, 3, 3, 1, 1), (2, 16, 16, 640, 1280, 1, 1, 0, 1),
    (1, 32, 32, 640, 640, 3, 3, 1, 2), (3, 7, 9, 96, 40, 3, 3, 1, 1), (1, 5, 5, 32, 8, 1, 1, 0, 1),
]


@pytest.mark.parametrize("n,h,w,c,k,r,s,pad,stride", W4_CONVS)
@pytest.mark.parametrize("dynamic", [False, True])
def test_w4_conv_bit_exact(ops, dev, n, h, w, c, k, r, s, pad, stride, dynamic):
    from mixdq_b200 import _lib
    from mixdq_b200.nn.utils import pack_int4
    g = torch.Generator().manual_seed(n * h * w + c + k)
    x = torch.randint(-128, 128, (n, c, h, w), dtype=torch.int8, generator=g)
    codes = torch.randint(-8, 8, (k, c, r, s), dtype=torch.int8, generator=g)
    packed = pack_int4(codes.contiguous(memory_format=torch.channels_last), dim=1)
    assert packed.shape == (k, c // 2, r, s) and packed.is_contiguous(memory_format=torch.channels_last)
    w_scale = 0.001 + 0.01 * torch.rand(k, generator=g)
    a_scale, a_zp = torch.tensor(0.123), torch.tensor(7.0)
    b = torch.rand(k, generator=g).half()
    wsum_krs = codes.float().sum(dim=1, keepdim=True) if pad > 0 else None
    wsum_k = codes.float().sum(dim=[1, 2, 3]) if pad == 0 else None
    P = (h + 2 * pad - r) // stride + 1
    Q = (w + 2 * pad - s) // stride + 1
    acc = torch.empty(n * P * Q, k, dtype=torch.int32, device=dev)
    xd = x.to(dev).contiguous(memory_format=torch.channels_last)
    if dynamic:
        out = ops.qconv2d_dynamic_fused(xd, packed.to(dev), w_scale.to(dev), a_scale.to(dev), a_zp.to(dev),
                                        None if wsum_krs is None else wsum_krs.to(dev),
                                        None if wsum_k is None else wsum_k.to(dev), b.to(dev),
                                        stride, pad, _acc_out=acc)
    else:
        out = ops.qconv2d_w8_a8_ohalf(xd, packed.to(dev), w_scale.to(dev), a_scale.to(dev), a_zp.to(dev),
                                      (w_scale * a_scale).to(dev),
                                      None if wsum_krs is None else wsum_krs.to(dev),
                                      None if wsum_k is None else (wsum_k * a_zp).to(dev), b.to(dev),
                                      stride, pad, 1, _acc_out=acc)
    torch.cuda.synchronize()
    assert _lib.last_path().startswith("tcgen05-w4")
    ref, ref_acc = O.qconv2d_kernel(x, codes, w_scale * a_scale, wsum_krs,
                                    None if wsum_k is None else wsum_k * a_zp, a_zp, b, stride, pad)
    got_acc = acc.cpu().view(n, P, Q, k).permute(0, 3, 1, 2).long()
    assert torch.equal(got_acc, ref_acc), "INT32 accumulators differ"
    assert torch.equal(bits(out.contiguous()), bits(ref)), "fp16 outputs differ"


@pytest.mark.parametrize("M,inner,K", [(256, 5120, 1280), (1024, 2560, 640), (77, 64, 96)])
def test_w4_geglu_projection(ops, dev, M, inner, K):
    """packed-W4 ff.net.0.proj with the GEGLU in the epilogue == the W8 kernel on the unpacked
    codes (same epilogue code, so bit-identical), which tests/test_gpu_fused.py pins to PyTorch."""
    g = torch.Generator().manual_seed(inner + K)
    x = torch.randn(M, K, generator=g).half()
    codes = torch.randint(-8, 8, (2 * inner, K), dtype=torch.int8, generator=g)
    w_scale = 0.002 + 0.01 * torch.rand(2 * inner, generator=g)
    wsum = codes.float().sum(1)
    b = torch.randn(2 * inner, generator=g).half()
    idx = ops.geglu_interleave_index(inner)
    c_il, s_il, ws_il, b_il = codes[idx], w_scale[idx], wsum[idx], b[idx]
    q8, s, z = ops.quantize_per_tensor_dynamic(x.to(dev))
    q4, s4, z4, y4 = ops.qlinear_geglu_quantize_dynamic(
        q8, O.pack_int4(c_il).to(dev), s_il.to(dev), s, z, ws_il.to(dev), b_il.to(dev), return_y=True)
    from mixdq_b200 import _lib
    assert _lib.last_path() != "simt"
    q8_, s8, z8, y8 = ops.qlinear_geglu_quantize_dynamic(
        q8, c_il.to(dev), s_il.to(dev), s, z, ws_il.to(dev), b_il.to(dev), return_y=True)
    assert torch.equal(bits(y4), bits(y8)) and torch.equal(q4, q8_)
    assert s4.item() == s8.item() and z4.item() == z8.item()


def _prep(mod, name, w_bit):
    dt = torch.qint8 if w_bit == 8 else torch.quint4x2
    mod.qconfig = QConfig(activation=PlaceholderObserver.with_args(dtype=torch.qint8),
                          weight=PlaceholderObserver.with_args(dtype=dt))
    mod.module_name = name
    mod.w_bit = w_bit
    mod.a_bit = 8
    return mod


@pytest.mark.parametrize("w_bit", [4, 2])
def test_w4_modules_vs_fake_quant(dev, w_bit):
    """QuantizedLinear / QuantizedConv2d with a 4-bit (or 2-bit, promoted to 4) qconfig, dynamic
    activations: packed buffers, tcgen05-w4 path, output within the north-star tolerance of the
    qdiff fake-quant path (4-bit weight quantiser, 8-bit activation quantiser) on the CPU."""
    from mixdq_b200 import _lib
    from mixdq_b200.nn import QuantizedConv2d, QuantizedLinear
    torch.manual_seed(0)
    lin = nn.Linear(640, 1280).half()
    q = QuantizedLinear.from_float(_prep(lin, "lin", w_bit), ckpt=None).to(dev)
    assert q._get_name() == "QuantizedLinearW4A8" and q.weight_int4.shape == (1280, 320)
    assert q.weight_int4.dtype == torch.uint8 and not hasattr(q, "weight_int")
    x = torch.randn(2, 77, 640).half()
    y = q(x.to(dev))
    assert _lib.last_path().startswith("tcgen05-w4")
    ref = O.fake_quant_layer(x.float(), lin.weight.float(), lin.bias.float(), w_bits=4, a_bits=8)
    err = (y.float().cpu() - ref).abs().max().item() / ref.abs().max().item()
    cos = torch.nn.functional.cosine_similarity(y.float().cpu().flatten(), ref.flatten(), dim=0).item()
    assert err <= 1e-2 and cos >= 0.9999, (err, cos)

    conv = nn.Conv2d(64, 128, 3, padding=1).half()
    qc = QuantizedConv2d.from_float(_prep(conv, "conv", w_bit), ckpt=None).to(dev)
    assert qc._get_name() == "QuantizedConv2dW4A8" and qc.weight_int4.shape == (128, 32, 3, 3)
    xc = torch.randn(2, 64, 16, 16).half()
    yc = qc(xc.to(dev).contiguous(memory_format=torch.channels_last))
    assert _lib.last_path().startswith("tcgen05-w4")
    refc = O.fake_quant_layer(xc.float(), conv.weight.float(), conv.bias.float(), w_bits=4, a_bits=8,
                              stride=1, padding=1)
    err = (yc.float().cpu() - refc).abs().max().item() / refc.abs().max().item()
    cos = torch.nn.functional.cosine_similarity(yc.float().cpu().flatten(), refc.flatten(), dim=0).item()
    assert err <= 1e-2 and cos >= 0.9999, (err, cos)
    # conv_in-like layer the packed kernel cannot take (C = 4) keeps one code per int8
    c4 = nn.Conv2d(4, 64, 3, padding=1).half()
    q4 = QuantizedConv2d.from_float(_prep(c4, "conv_in", w_bit), ckpt=None).to(dev)
    assert hasattr(q4, "weight_int") and q4.weight_int.abs().max() <= 8
    q4(torch.randn(1, 4, 8, 8).half().to(dev))


def test_mixed_precision_unet_fused_matches_unfused(dev):
    """tiny SDXL-topology UNet with a weight_5.02-like mix (W4 cross-attention K/V/Q, W4 feed-forward,
    some W4 resnet convs, W8 elsewhere): the W4 blocks ARE fused (no fallback), and the fused UNet
    reproduces the leaf-by-leaf quantized UNet block by block."""
    from mixdq_b200 import mixdq, ops
    from mixdq_b200.fused import fuse_unet
    from mixdq_b200.unet import build_unet
    unet = build_unet("tiny", seed=3).half()
    names = [n for n, _ in unet.quantizable_layers()]

    def wbits(n):
        if "attn2.to_k" in n or "attn2.to_q" in n or "ff.net" in n:
            return 4
        if "attn2.to_v" in n or "attn2.to_out" in n:
            return 2
        if n.endswith("conv1") and "down_blocks" in n:
            return 4
        return 8
    w_bits = {n: wbits(n) for n in names}
    a_bits = {n: 8 for n in names}
    inputs = unet.example_inputs(2, "cpu", torch.float16, seed=1)
    mixdq.quantize_unet(unet, SimpleNamespace(w_config=w_bits, a_config=a_bits), ckpt=None, bos=False,
                        bos_dict=None, fuse=False)
    unet = unet.to(dev).to(memory_format=torch.channels_last)
    kw = {k: v.to(dev) for k, v in inputs.items()}
    kinds = ("BasicTransformerBlock", "ResnetBlock2D", "Transformer2DModel")
    blocks = [(n, m) for n, m in unet.named_modules() if type(m).__name__ in kinds]
    rec = {}

    def hook(name):
        def f(m, inp, out):
            rec[name] = ([t.detach().clone() for t in inp], out.detach().clone())
        return f
    handles = [m.register_forward_hook(hook(n)) for n, m in blocks]
    with torch.no_grad():
        plain = unet(**kw)[0].clone()
        sd_before = {k: v.clone() for k, v in unet.state_dict().items()}
        for h in handles:
            h.remove()
        summary = fuse_unet(unet)
        assert summary["transformer_blocks"] == 4 and summary["resnets"] == 8, summary
        rec_fam = ops.start_recording()
        fused = unet(**kw)[0].clone()
        ops.stop_recording()
        fams = {r[0] for r in rec_fam}
        assert {"gemm_w4", "gemm_geglu_w4", "conv_w4", "gemm", "conv"} <= fams, fams
        for n, m in blocks:
            inp, out = rec[n]
            got = m(*inp)
            err = (got.float() - out.float()).abs().max().item() / out.float().abs().max().item()
            cos = torch.nn.functional.cosine_similarity(got.float().flatten(), out.float().flatten(), dim=0).item()
            assert err <= 1e-2 and cos >= 0.9999, (n, err, cos)
        # the GEGLU-interleaved stored layout is invisible from outside: state_dict and the
        # module's own forward keep the stock row order
        sd_after = unet.state_dict()
        assert sd_before.keys() == sd_after.keys()
        for k in sd_before:
            assert torch.equal(sd_before[k], sd_after[k]), k
        proj = unet.mid_block.attentions[0].transformer_blocks[0].ff.net[0].proj
        assert proj.geglu_interleaved
        xin = torch.randn(2, 16, proj.in_features, device=dev).half()
        y_il = proj(xin)
    err = (fused.float() - plain.float()).abs().max().item() / plain.float().abs().max().item()
    assert err <= 8e-2, err
