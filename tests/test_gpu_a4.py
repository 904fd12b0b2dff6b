"""N3 — 4-bit activation layers (kernels/cfgs/act/act_7.xx.yaml, `a_bit: 4`). The reference gates
them to fp16 (nn/Linear.py:28-36), so the arithmetic is the qdiff asymmetric quantiser at
n_bits = 4 (base_quantizer.py:155-190): delta = (max - min) / 15, z = round(-min / delta),
q = clamp(round(x / delta) + z, 0, 15). Codes stay unsigned (one per int8, no -128 shift) and feed
the same int8 kernels; the integer identity of op/qlinear.py:66-83 then holds with zp = z.
Codes / (delta, z) / INT32 accumulators bit-exact; fp16 outputs at the north-star tolerance against
the fake-quant float path."""
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


def _lib_path():
    from mixdq_b200 import _lib
    return _lib.load().mixdq_last_path().decode()


def _oracle_codes(x, n_bits):
    delta, z = O.act_qparams_minmax(x, n_bits)
    q, _ = O.act_fake_quant(x.float(), delta, z, n_bits)
    return q, delta, z


@pytest.mark.parametrize("numel", [8, 4096, 77 * 2048, 256 * 1280, 4096 * 640, 8 * 1024 * 2560])
def test_dynamic_a4_codes_bit_exact(ops, dev, numel):
    g = torch.Generator().manual_seed(numel)
    x = (torch.randn(numel, generator=g) * 1.7 + 0.2).half()
    q, s, z = ops.quantize_per_tensor_dynamic_bits(x.to(dev), 4)
    qr, dr, zr = _oracle_codes(x, 4)
    assert q.dtype == torch.int8
    assert torch.equal(s.cpu(), dr.reshape(())) and torch.equal(z.cpu(), zr.reshape(()))
    assert torch.equal(q.cpu().to(torch.int64), qr.to(torch.int64))
    assert int(q.min()) >= 0 and int(q.max()) <= 15


def test_dynamic_a4_one_sided_and_8bit_alias(ops, dev):
    x = torch.rand(4096).half() + 0.5          # min clamps to 0 (base_quantizer.py:155-158)
    q, s, z = ops.quantize_per_tensor_dynamic_bits(x.to(dev), 4)
    qr, dr, zr = _oracle_codes(x, 4)
    assert float(z) == 0.0 and torch.equal(q.cpu().to(torch.int64), qr.to(torch.int64))
    # n_bits = 8 is the ordinary dynamic quantiser
    q8, s8, z8 = ops.quantize_per_tensor_dynamic_bits(x.to(dev), 8)
    r8, rs, rz = O.quantize_dynamic_kernel(x)
    assert torch.equal(q8.cpu(), r8) and torch.equal(s8.cpu(), rs) and torch.equal(z8.cpu(), rz)


def test_large_tensor_batched_quantise_pass(ops, dev):
    """tensors of several waves take the four-vectors-per-thread form of the quantise pass
    (csrc/quant2.cu quant_rows_premm_kernel<4>): same codes as the oracle, 8 and 4 bit"""
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(8 * 4096, 640, generator=g) * 2.1 - 0.4).half()
    q8, s8, z8 = ops.quantize_per_tensor_dynamic(x.to(dev))
    r8, rs, rz = O.quantize_dynamic_kernel(x)
    assert torch.equal(s8.cpu(), rs) and torch.equal(z8.cpu(), rz) and torch.equal(q8.cpu(), r8)
    q4, s4, z4 = ops.quantize_per_tensor_dynamic_bits(x.to(dev), 4)
    qr, dr, zr = _oracle_codes(x, 4)
    assert torch.equal(q4.cpu().to(torch.int64), qr.to(torch.int64))
    # row-pitched view (a column slice) through the same pass
    wide = (torch.randn(4096, 1920, generator=g) * 3).half().to(dev)
    qv, sv, zv = ops.quantize_rows_dynamic(wide[:, 640:])
    rv, rsv, rzv = O.quantize_dynamic_kernel(wide[:, 640:].cpu().contiguous())
    assert torch.equal(qv.cpu(), rv) and torch.equal(sv.cpu(), rsv)


def test_static_a4_codes(ops, dev):
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(3, 77, 640, generator=g) * 2).half()
    delta, zp = torch.tensor(0.5459), torch.tensor(8.0)      # conv_in 4-bit entry of new_ckpt.pth
    inv = (1 / delta)
    q = ops.quantize_per_tensor_to_int4_codes(x.to(dev), inv.to(dev), zp.to(dev))
    assert q.shape == x.shape and int(q.min()) >= 0 and int(q.max()) <= 15
    # the kernel contracts x * s + z into one FMA like the reference's (quantize_kernel.cu:20-24):
    # fp16 x fp32 + 8 is exact in fp64, so rounding it to fp32 IS the correctly rounded FMA
    fma32 = (x.double() * inv.double() + 8.0).float()
    ref = torch.clamp(torch.round(fma32), 0, 15)
    assert torch.equal(q.cpu().float(), ref)


def _float_linear(K, N, seed, w_bit, a_bit, name="blk.attn2.to_out.0"):
    torch.manual_seed(seed)
    fm = nn.Linear(K, N).half()
    wd = torch.qint8 if w_bit == 8 else torch.quint4x2
    ad = torch.qint8 if a_bit == 8 else torch.quint4x2
    fm.qconfig = QConfig(activation=PlaceholderObserver.with_args(dtype=ad),
                         weight=PlaceholderObserver.with_args(dtype=wd))
    fm.module_name, fm.w_bit, fm.a_bit = name, w_bit, a_bit
    return fm


@pytest.mark.parametrize("M,N,K,w_bit", [(256, 1280, 1280, 8), (1024, 640, 640, 8),
                                         (77, 1280, 2048, 4), (256, 1280, 5120, 4)])
def test_linear_w8a4_w4a4_dynamic(ops, dev, M, N, K, w_bit):
    from mixdq_b200.nn.linear import QuantizedLinear
    from mixdq_b200.nn.utils import unpack_int4
    fm = _float_linear(K, N, M + N, w_bit, 4).to(dev)
    qm = QuantizedLinear.from_float(fm, ckpt=None)
    assert qm.valid_for_acceleration and qm.dynamic and qm.a_bits == 4
    assert qm._get_name() == f"QuantizedLinearW{w_bit}A4"
    g = torch.Generator().manual_seed(K)
    x = (torch.randn(1, M, K, generator=g) * 1.5).half()
    y = qm(x.to(dev))
    assert "tcgen05" in _lib_path()
    # integer identity on the oracle's codes: bit-exact fp16
    qa, da, za = _oracle_codes(x, 4)
    w_codes = (qm.weight_int if w_bit == 8 else unpack_int4(qm.weight_int4)).cpu()
    acc = O.int_accumulate_linear(qa.reshape(M, K).to(torch.int8), w_codes)
    wsum = w_codes.float().sum(1)
    ws = qm.weight_scales.cpu()
    ref = O.kernel_epilogue(acc, wsum * za.reshape(()), ws * da.reshape(()), qm.bias.cpu())
    assert torch.equal(y.cpu().reshape(M, N), ref)
    # fake-quant float path at the north-star tolerance (max-abs <= 1e-2 of the output range,
    # cosine >= 0.9999)
    fq = O.fake_quant_layer(x.float().reshape(M, K), fm.weight.detach().cpu().float(),
                            fm.bias.detach().cpu().float(), w_bits=w_bit, a_bits=4)
    out = y.cpu().float().reshape(M, N)
    cos = torch.nn.functional.cosine_similarity(out.flatten(), fq.flatten(), dim=0).item()
    assert cos >= 0.9999, cos
    assert (out - fq).abs().max().item() <= 1e-2 * max(fq.abs().max().item(), 1.0)


def test_a4_layers_of_act_7_77_on_the_unet(dev):
    """kernels/cfgs/act/act_7.77.yaml on the SDXL skeleton: its 66 `a_bit: 4` layers (all Linear)
    become ...A4 modules, none falls back to fp16 for being 4-bit"""
    from types import SimpleNamespace
    from mixdq_b200 import mixdq
    from mixdq_b200.unet import UNet2DConditionModel, sdxl_turbo_config
    from mixdq_b200.nn.linear import QuantizedLinear
    a_bits = mixdq.load_bit_config("act/act_7.77.yaml")
    four = sorted(n for n, b in a_bits.items() if b == 4)
    assert len(four) == 66
    with torch.device("meta"):
        unet = UNet2DConditionModel(sdxl_turbo_config()).half()
    mods = dict(unet.named_modules())
    assert all(isinstance(mods[n], nn.Linear) for n in four)
    # convert just those layers (materialised one at a time: no 5 GB model needed)
    for n in four[:6]:
        m = mods[n]
        fm = _float_linear(m.in_features, m.out_features, 1, 8, 4, name=n).to(dev)
        qm = QuantizedLinear.from_float(fm, ckpt=None)
        assert qm._get_name() == "QuantizedLinearW8A4"
