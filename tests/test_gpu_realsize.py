"""Parity at the REAL SDXL-Turbo block sizes: one 1280-wide transformer block at
256 tokens (the 16x16 level: 60 of the UNet's 70 blocks), the 1280@16x16 and 320@64x64 resnets and
a split-shortcut up-block resnet, against the qdiff fake-quant oracle (oracle/unet_oracle.py wraps
every Linear / Conv2d leaf in the fp32 restatement of QuantLayer.forward, quant_layer.py:63-103).

  * teacher-forced: every quantised leaf, fed the ORACLE's input of that leaf, reproduces the
    oracle's output inside the north-star tolerance (max-abs <= 1e-2 of the output range,
    cosine >= 0.9999) — on the module path AND inside the fused block forwards;
  * free-running: the block output of the fused int8 path stays as close to the oracle as the
    plain fp16 block does (quantisation-noise-sized bound, as in tests/test_gpu_modules.py).

Plus: static (checkpoint) activation scales make the UNet invariant under batch sharding — the
claim behind the data-parallel mode (mixdq_b200/dp.py:9-10, BASELINE config 4)."""
import copy
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

from oracle import unet_oracle as UO

pytestmark = pytest.mark.gpu
TOL_ABS, TOL_COS = 1e-2, 0.9999


@pytest.fixture(scope="module")
def dev():
    from mixdq_b200 import build
    build.build()
    return torch.device("cuda:0")


def _stats(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)
    cos = torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()
    return err, cos


def _leaf_names(block):
    return [n for n, m in block.named_modules() if isinstance(m, (nn.Linear, nn.Conv2d))]


def _quantize(block, names, dev, splits=None, fuse=True):
    from mixdq_b200 import mixdq
    q = copy.deepcopy(block).half()
    for n, s in (splits or {}).items():
        q.get_submodule(n).split = s
    args = SimpleNamespace(w_config={n: 8 for n in names}, a_config={n: 8 for n in names})
    q = q.to(dev).to(memory_format=torch.channels_last)
    mixdq.quantize_unet(q, args, ckpt=None, bos=False, bos_dict=None, fuse=fuse)
    return q.eval()


def _oracle(block, names, splits=None):
    ref = copy.deepcopy(block).float()
    UO.wrap_unet(ref, {n: 8 for n in names}, {n: 8 for n in names}, splits or {}, bos=False)
    return ref.eval()


def _teacher_forced(qblock, ref, names, ref_call, dev):
    """run the oracle once, recording every leaf's (input, output); check the GPU leaves on them"""
    rec = {}

    def hook(name):
        def f(m, inp, out):
            rec[name] = (inp[0].detach(), out.detach())
        return f
    handles = [ref.get_submodule(n).register_forward_hook(hook(n)) for n in names]
    with torch.no_grad():
        ref_out = ref_call(ref)
    for h in handles:
        h.remove()
    assert len(rec) == len(names)
    worst = (0.0, 1.0, None)
    with torch.no_grad():
        for n in names:
            xi, yo = rec[n]
            xin = xi.half().to(dev)
            if xin.dim() == 4:
                xin = xin.contiguous(memory_format=torch.channels_last)
            err, cos = _stats(qblock.get_submodule(n)(xin), yo)
            assert err <= TOL_ABS and cos >= TOL_COS, (n, err, cos)
            if err > worst[0]:
                worst = (err, cos, n)
    return ref_out, worst


def test_sdxl_transformer_block_1280_at_256_tokens(dev):
    from mixdq_b200.unet import BasicTransformerBlock
    torch.manual_seed(11)
    blk = BasicTransformerBlock(1280, 2048, 64)
    names = _leaf_names(blk)
    assert len(names) == 10                  # q,k,v,out x2 + ff.net.0.proj + ff.net.2
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 256, 1280, generator=g).half()
    ctx = torch.randn(1, 77, 2048, generator=g).half()
    ref = _oracle(blk, names)
    q = _quantize(blk, names, dev)
    assert getattr(q, "_mixdq_fused", None) is not None, "the real-size block must take the fused path"
    ref_out, worst = _teacher_forced(q, ref, names, lambda r: r(x.float(), ctx.float()), dev)
    with torch.no_grad():
        got = q(x.to(dev), ctx.to(dev))
        fp = copy.deepcopy(blk).half().to(dev)(x.to(dev), ctx.to(dev))
    err_q, cos_q = _stats(got, ref_out)
    err_fp, _ = _stats(fp, ref_out)
    # free-running: fused int8 block vs oracle, bounded by the distance of the fp16 block itself
    assert cos_q >= 0.9995 and err_q <= max(2.5 * err_fp, TOL_ABS), (err_q, cos_q, err_fp, worst)


@pytest.mark.parametrize("cin,cout,hw", [(1280, 1280, 16), (320, 320, 64), (640, 1280, 16)])
def test_sdxl_resnets_real_size(dev, cin, cout, hw):
    from mixdq_b200.unet import ResnetBlock2D
    torch.manual_seed(cin + hw)
    blk = ResnetBlock2D(cin, cout, 1280, 32)
    names = _leaf_names(blk)
    g = torch.Generator().manual_seed(hw)
    x = torch.randn(1, cin, hw, hw, generator=g).half()
    temb = torch.randn(1, 1280, generator=g).half()
    ref = _oracle(blk, names)
    q = _quantize(blk, names, dev)
    ref_out, worst = _teacher_forced(q, ref, names, lambda r: r(x.float(), temb.float()), dev)
    xin = x.to(dev).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        got = q(xin, temb.to(dev))
        fp = copy.deepcopy(blk).half().to(dev).to(memory_format=torch.channels_last)(xin, temb.to(dev))
    err_q, cos_q = _stats(got, ref_out)
    err_fp, _ = _stats(fp, ref_out)
    assert cos_q >= 0.9995 and err_q <= max(2.5 * err_fp, TOL_ABS), (err_q, cos_q, err_fp, worst)


def test_up_block_resnet_with_split_shortcut_real_size(dev):
    """up_blocks.0.resnets.0 of SDXL: 2560 -> 1280 at 16x16, conv_shortcut split at 1280 — the two
    halves of the concatenated input are quantised independently (quant_layer.py:74-88)"""
    from mixdq_b200.unet import ResnetBlock2D
    torch.manual_seed(3)
    blk = ResnetBlock2D(2560, 1280, 1280, 32)
    names = _leaf_names(blk)
    splits = {"conv_shortcut": 1280}
    g = torch.Generator().manual_seed(9)
    x = torch.cat([torch.randn(1, 1280, 16, 16, generator=g), 3 * torch.randn(1, 1280, 16, 16, generator=g)],
                  dim=1).half()
    temb = torch.randn(1, 1280, generator=g).half()
    ref = _oracle(blk, names, splits)
    # the module name must look like an up-block shortcut for `convert` to hand the split over
    q = copy.deepcopy(blk).half()
    q.conv_shortcut.split = 1280
    from mixdq_b200 import mixdq
    args = SimpleNamespace(w_config={n: 8 for n in names}, a_config={n: 8 for n in names})
    q = q.to(dev).to(memory_format=torch.channels_last)
    mixdq.register_qconfig_from_input_files(q, args, bos=False, bos_dict=None)
    q.conv_shortcut.module_name = "up_blocks.0.resnets.0.conv_shortcut"
    mixdq.convert_to_quantized(q, None)
    assert q.conv_shortcut.split == 1280 and q.conv_shortcut.weight_int_0.shape[1] == 1280
    _teacher_forced(q.eval(), ref, names, lambda r: r(x.float(), temb.float()), dev)


def test_static_scales_are_invariant_under_batch_sharding(dev):
    """mixdq_b200/dp.py:9-10: with static (PTQ checkpoint) activation scales every sample's step is
    independent of the rest of the batch, so sharding the batch over ranks changes nothing:
    batch 4 in one piece == two shards of 2 == four shards of 1, bit for bit. (Dynamic per-tensor
    scales take min/max over the local shard and do NOT have this property — also checked.)"""
    import bench
    from mixdq_b200 import dp
    fp16 = bench.build_fp16_unet("tiny", dev, seed=2)
    inputs = fp16.example_inputs(4, dev, torch.float16, seed=7)
    static = bench.quantize_copy(fp16, "static")
    with torch.no_grad():
        whole = static(**inputs)[0]
        for world in (2, 4):
            parts = [static(**dp.shard_inputs(inputs, 4, world, r))[0] for r in range(world)]
            assert torch.equal(torch.cat(parts, dim=0), whole), world
        dyn = bench.quantize_copy(fp16, "dynamic")
        whole_d = dyn(**inputs)[0]
        parts_d = torch.cat([dyn(**dp.shard_inputs(inputs, 4, 2, r))[0] for r in range(2)], dim=0)
    err, cos = _stats(parts_d, whole_d)
    assert cos >= 0.99 and not torch.equal(parts_d, whole_d), (err, cos)


def test_quantized_file_loads_without_a_float_model(dev, tmp_path):
    """N2: quantise (fused) -> save -> rebuild on a META skeleton straight onto the GPU: identical
    outputs, the stored GEGLU-interleaved layout is kept as it is (no second copy), and the device
    never holds more than the quantised model + its non-quantised parameters."""
    import bench
    from mixdq_b200 import serialize
    from mixdq_b200.unet import UNet2DConditionModel, tiny_config
    fp16 = bench.build_fp16_unet("tiny", dev, seed=4)
    inputs = fp16.example_inputs(2, dev, torch.float16, seed=3)
    q = bench.quantize_copy(fp16, "dynamic", fuse=True)
    with torch.no_grad():
        want = q(**inputs)[0].clone()
    path = tmp_path / "q.pt"
    serialize.save_quantized_unet(q, path)
    ff = [m for n, m in q.named_modules() if n.endswith("ff.net.0.proj")]
    assert ff and all(getattr(m, "geglu_interleaved", False) for m in ff)
    del q, fp16
    torch.cuda.empty_cache()
    base = torch.cuda.memory_allocated(dev)
    with torch.device("meta"):
        skel = UNet2DConditionModel(tiny_config())
    loaded = serialize.load_quantized_unet(skel, path, dev).to(memory_format=torch.channels_last).eval()
    resident = torch.cuda.memory_allocated(dev) - base
    assert resident <= 1.05 * bench.module_bytes(loaded) + (1 << 20), (resident, bench.module_bytes(loaded))
    assert getattr(loaded, "_mixdq_fused_summary", None) is not None
    with torch.no_grad():
        got = loaded(**inputs)[0]
    assert torch.equal(got, want)


def test_ptq_calibration_on_the_gpu(dev):
    """N4: calibrate the fp16 UNet on the GPU (min / max of every layer input by the library's
    reduction kernel), quantise with the resulting checkpoint (static scales, fused blocks) and
    stay close to the fp16 UNet; the min/max kernel agrees with torch exactly."""
    import bench
    from mixdq_b200 import mixdq, ops, ptq
    g = torch.Generator().manual_seed(0)
    for shape in ((8,), (256, 1280), (2, 320, 64, 64), (77 * 2048,)):
        x = (torch.randn(*shape, generator=g) * 3 - 0.5).half().to(dev)
        mm = ops.tensor_minmax(x).cpu()
        assert float(mm[0]) == min(float(x.min()), 0.0) and float(mm[1]) == max(float(x.max()), 0.0)
    pos = torch.rand(4096, generator=g).half().to(dev) + 1
    assert float(ops.tensor_minmax(pos)[0]) == 0.0              # range always contains zero
    fp16 = bench.build_fp16_unet("tiny", dev, seed=6)
    batches = [fp16.example_inputs(2, dev, torch.float16, seed=s) for s in (11, 12, 13)]
    ck = ptq.calibrate(fp16, batches)
    # same statistics as the torch path on the CPU copy of the inputs (fp16 min/max is exact)
    name = "mid_block.attentions.0.transformer_blocks.0.attn1.to_q"
    assert ck[name + ".act_quantizer"]["delta_list"].shape == (3,)
    names = [n for n, _ in fp16.quantizable_layers()]
    args = SimpleNamespace(w_config={n: 8 for n in names}, a_config={n: 8 for n in names})
    q = copy.deepcopy(fp16)
    mixdq.quantize_unet(q, args, ckpt=ck, bos=False, bos_dict=None)
    q = q.to(memory_format=torch.channels_last).eval()
    assert getattr(q, "_mixdq_fused_summary", None) is not None
    test_in = fp16.example_inputs(2, dev, torch.float16, seed=14)
    with torch.no_grad():
        want = fp16(**test_in)[0]
        got = q(**test_in)[0]
    err, cos = _stats(got, want)
    assert cos >= 0.99 and err <= 0.15, (err, cos)
