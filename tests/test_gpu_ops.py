"""Parity of the CUDA kernels (through the op layer / C ABI) against the CPU oracle.
Integer, byte and index results are compared bit-exactly; fp16 outputs bit-exactly against the
oracle's restatement of the reference epilogue. Includes the reference's own op self-tests
(op/quant.py:7-61, op/qlinear.py:29-108, op/qconv2d.py:25-119) with their shapes and tolerances."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import qdiff_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from mixdq_b200 import build
    build.build()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops(dev):
    from mixdq_b200 import ops as _ops
    return _ops


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int16)


# ------------------------------------------------------------------------------------------- A1
def test_reference_quant_selftest(ops, dev):
    """op/quant.py:7-30: both entry points == torch.quantize_per_tensor on rand(1024)."""
    from mixdq_extension.op.quant import quantize_per_tensor, quantize_per_tensor_vectorized
    torch.manual_seed(0)
    t = torch.rand((1024,), dtype=torch.float16, device=dev)
    tf = t.float()
    zero_point = torch.round((tf.max() + tf.min()) / 2)
    scale = (tf.max() - tf.min()) / 255
    q1 = quantize_per_tensor(t, 1 / scale, zero_point)
    q2 = quantize_per_tensor_vectorized(t, 1 / scale, zero_point)
    ref = torch.quantize_per_tensor(tf.cpu(), scale.cpu(), zero_point.cpu(), torch.qint8).int_repr()
    assert torch.equal(q1.cpu(), ref) and torch.equal(q1, q2)
    assert q1.dtype == torch.int8 and q1.shape == t.shape


@pytest.mark.parametrize("numel", [1, 7, 8, 1024, 77 * 2048, 1000003, 4096 * 960])
def test_quant_static_flat_bit_exact(ops, dev, numel):
    g = torch.Generator().manual_seed(numel)
    x = (torch.randn(numel, generator=g) * 2.5).half()
    sinv, zp = torch.tensor(1 / 0.0323), torch.tensor(2.0)
    q = ops.quantize_per_tensor_to_int8(x.to(dev), sinv.to(dev), zp.to(dev))
    assert torch.equal(q.cpu(), O.quantize_static_kernel(x, sinv.item(), zp.item()))


def test_quant_static_saturates_and_handles_golden(ops, dev, golden_dir):
    z = np.load(golden_dir / "torch_quantize_known_answer.npz")
    x = torch.from_numpy(z["x2"])
    sinv = (1.0 / torch.from_numpy(z["scale2"])).float()
    zp = torch.from_numpy(z["zp2"]).float()
    q = ops.quantize_per_tensor_to_int8(x.to(dev), sinv.to(dev), zp.to(dev)).cpu()
    assert torch.equal(q, O.quantize_static_kernel(x, sinv.item(), zp.item()))
    assert q.min() == -128 and q.max() == 127
    assert (q != torch.from_numpy(z["q2"])).sum() <= 2      # vs torch.quantize_per_tensor


def test_quant_empty_and_layout_preserved(ops, dev):
    s, z = torch.tensor(3.0, device=dev), torch.tensor(1.0, device=dev)
    assert ops.quantize_per_tensor_to_int8(torch.empty(0, 8, dtype=torch.half, device=dev), s, z).numel() == 0
    x = torch.randn(2, 32, 5, 7, device=dev).half().contiguous(memory_format=torch.channels_last)
    q = ops.quantize_per_tensor_to_int8(x, s, z)
    assert q.stride() == x.stride()          # empty_like semantics (quantize.cc:26)
    assert torch.equal(q.cpu(), O.quantize_static_kernel(x.cpu(), 3.0, 1.0))


def test_quant_error_behaviour(ops, dev):
    s, z = torch.tensor(3.0, device=dev), torch.tensor(1.0, device=dev)
    with pytest.raises(RuntimeError, match="input should be fp16"):
        ops.quantize_per_tensor_to_int8(torch.zeros(8, device=dev), s, z)
    with pytest.raises(RuntimeError, match="input should be on CUDA"):
        ops.quantize_per_tensor_to_int8(torch.zeros(8).half(), s, z)
    with pytest.raises(RuntimeError, match="scale_inv should be fp32"):
        ops.quantize_per_tensor_to_int8(torch.zeros(8, device=dev).half(), s.half(), z)


def test_quant_strided_views(ops, dev):
    """Views the reference kernel mis-reads at batch > 1 are handled by stride."""
    g = torch.Generator().manual_seed(3)
    s, z = torch.tensor(17.3), torch.tensor(-3.0)
    x = torch.randn(3, 96, 10, 6, generator=g).half()
    xd = x.to(dev).contiguous(memory_format=torch.channels_last)
    for sl in (slice(0, 32), slice(32, 96)):
        q = ops.quantize_per_tensor_to_int8(xd[:, sl], s.to(dev), z.to(dev))
        assert torch.equal(q.cpu(), O.quantize_static_kernel(x[:, sl], s.item(), z.item()))
    x3 = torch.randn(4, 77, 64, generator=g).half()
    q = ops.quantize_per_tensor_to_int8(x3.to(dev)[:, 1:, :], s.to(dev), z.to(dev))
    assert torch.equal(q.cpu(), O.quantize_static_kernel(x3[:, 1:, :], s.item(), z.item()))
    # arbitrary non-dense view
    xt = torch.randn(16, 40, generator=g).half()
    q = ops.quantize_per_tensor_to_int8(xt.to(dev).t()[:, ::2], s.to(dev), z.to(dev))
    assert torch.equal(q.cpu(), O.quantize_static_kernel(xt.t()[:, ::2], s.item(), z.item()))


@pytest.mark.parametrize("shape,c0,c1", [((2, 96, 16, 16), 32, 96), ((1, 320, 64, 64), 0, 320),
                                         ((3, 20, 7, 5), 4, 16), ((1, 4, 9, 9), 0, 4)])
@pytest.mark.parametrize("nhwc_in", [False, True])
def test_quant_to_nhwc(ops, dev, shape, c0, c1, nhwc_in):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(*shape, generator=g).half()
    s, z = torch.tensor(9.1), torch.tensor(4.0)
    xd = x.to(dev)
    if nhwc_in:
        xd = xd.contiguous(memory_format=torch.channels_last)
    q = ops.quantize_to_nhwc(xd, s.to(dev), z.to(dev), c0, c1)
    assert q.is_contiguous(memory_format=torch.channels_last) or q.shape[1] == 1
    assert torch.equal(q.cpu(), O.quantize_static_kernel(x[:, c0:c1], s.item(), z.item()))


@pytest.mark.parametrize("numel", [5, 4096 * 320 + 3, 77 * 2048, 2 * 1280 * 16 * 16])
def test_quant_dynamic_matches_qdiff(ops, dev, numel):
    """codes, delta and zero point bit-exact against the qdiff restatement (fp32, true division)."""
    g = torch.Generator().manual_seed(numel)
    x = (torch.randn(numel, generator=g) * 1.7 + 0.3).half()
    q, s, z = ops.quantize_per_tensor_dynamic(x.to(dev))
    qr, sr, zr = O.quantize_dynamic_kernel(x)
    assert s.item() == sr.item() and z.item() == zr.item()
    assert torch.equal(q.cpu(), qr)
    # second call reuses the self-resetting workspace
    q2, s2, z2 = ops.quantize_per_tensor_dynamic((x * 2).to(dev))
    qr2, sr2, zr2 = O.quantize_dynamic_kernel(x * 2)
    assert torch.equal(q2.cpu(), qr2) and s2.item() == sr2.item() and z2.item() == zr2.item()


def test_quant_dynamic_one_sided_and_constant(ops, dev):
    for x in (torch.rand(4096).half() + 1.0, -torch.rand(4096).half() - 0.5, torch.zeros(64).half()):
        q, s, z = ops.quantize_per_tensor_dynamic(x.to(dev))
        qr, sr, zr = O.quantize_dynamic_kernel(x)
        assert torch.equal(q.cpu(), qr) and s.item() == sr.item() and z.item() == zr.item()


@pytest.mark.parametrize("lim", [0.37, 1.0, 3.3, 17.0, 900.0])
def test_quant_dynamic_every_fp16_value(ops, dev, lim):
    """Every finite fp16 value of [-lim, 0.71*lim] (all rounding boundaries that exist for the
    resulting delta): the kernels' reciprocal-multiply + exact fix-up equals the true division."""
    allh = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16).view(torch.float16)
    x = allh[torch.isfinite(allh) & (allh >= -lim) & (allh <= 0.71 * lim)].contiguous()
    x = torch.cat([x, x.flip(0)])[: (x.numel() * 2) // 64 * 64]
    q, s, z = ops.quantize_per_tensor_dynamic(x.to(dev))
    qr, sr, zr = O.quantize_dynamic_kernel(x)
    assert s.item() == sr.item() and z.item() == zr.item()
    assert torch.equal(q.cpu(), qr)
    q2, s2, z2 = ops.quantize_rows_dynamic(x.to(dev).view(8, -1))
    assert torch.equal(q2.cpu().view(-1), qr) and s2.item() == sr.item()


def test_quant_cuda_graph_replay(ops, dev):
    """The (commented-out) graph test of op/quant.py:32-61: capture once, replay on new data."""
    x = torch.rand(4096, device=dev).half()
    s = torch.tensor(100.0, device=dev)
    z = torch.tensor(-50.0, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            out = ops.quantize_per_tensor_to_int8(x, s, z)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = ops.quantize_per_tensor_to_int8(x, s, z)
    x2 = torch.rand(4096).half()
    x.copy_(x2); s.fill_(200.0); z.fill_(-100.0)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), O.quantize_static_kernel(x2, 200.0, -100.0))


# ------------------------------------------------------------------------------------------- A2
def test_reference_qlinear_selftest(ops, dev):
    """op/qlinear.py:29-108 run_test(64, 8, 16): same inputs, formulae and tolerances."""
    from mixdq_extension.op.qlinear import qlinear, quantize_per_tensor, quantize_per_tensor_vectorized
    torch.manual_seed(42)
    nsamples, ic, oc = 64, 8, 16
    input_fp16 = 6 * torch.rand(nsamples, ic, dtype=torch.float16, device=dev) - 3
    weight_int = torch.randint(-3, 3, (oc, ic), dtype=torch.int8, device=dev)
    input_scale = torch.scalar_tensor(0.123, dtype=torch.float32, device=dev)
    input_zp = torch.scalar_tensor(5.00, dtype=torch.float32, device=dev)
    weight_scale = 0.1 + torch.rand((oc,), dtype=torch.float32, device=dev)
    bias = torch.rand((oc,), device=dev, dtype=torch.float16)
    input_int = quantize_per_tensor(input_fp16, input_scale, input_zp).to(torch.int8)
    input_int_2 = quantize_per_tensor_vectorized(input_fp16, input_scale, input_zp).to(torch.int8)
    torch.testing.assert_close(input_int, input_int_2, atol=1., rtol=1e-2)
    output = qlinear(input_int, weight_int, weight_scale, input_scale, input_zp,
                     weight_int.float().sum(dim=1), weight_scale * input_scale,
                     weight_int.float().sum(dim=1) * input_zp, bias)
    infused_scale = weight_scale * input_scale
    offset = weight_scale * weight_int.to(torch.int32).sum(dim=1)
    offset *= input_zp * input_scale
    int_gemm_out = torch.matmul(input_int.to(torch.float32), weight_int.to(torch.float32).transpose(0, 1))
    reference_int = (int_gemm_out * infused_scale - offset + bias.float()).to(torch.float16)
    weight_fp = weight_int.to(torch.float32) * weight_scale[:, None]
    input_fp = (input_int.to(torch.float32) - input_zp) * input_scale
    reference_fp = (torch.matmul(input_fp, weight_fp.transpose(0, 1)) + bias.float()).half()
    torch.testing.assert_close(output, reference_int, atol=1e-4, rtol=1e-2)
    torch.testing.assert_close(output, reference_fp, rtol=1e-2, atol=1e-2)
    # fp16 debug GEMM vs torch.matmul (op/qlinear.py:85-95)
    a = 0.158 * torch.rand((nsamples, ic), device=dev, dtype=torch.float16)
    w = 0.158 ** torch.rand((ic, oc), device=dev, dtype=torch.float16)
    torch.testing.assert_close(ops.qlinear_fp_reference(a, w, bias), torch.matmul(a, w),
                               rtol=1e-4, atol=1e-2)


def _linear_case(ops, dev, M, N, K, bias=True, seed=0, lead=None, force_simt=False):
    from mixdq_b200 import _lib
    g = torch.Generator().manual_seed(seed * 7919 + M + N + K)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g)
    w = torch.randint(-128, 128, (N, K), dtype=torch.int8, generator=g)
    w_scale = 0.001 + 0.01 * torch.rand(N, generator=g)
    a_scale, a_zp = torch.tensor(0.0371), torch.tensor(-11.0)
    wsum = w.float().sum(1)
    scale, bias0 = w_scale * a_scale, wsum * a_zp
    b = torch.randn(N, generator=g).half() if bias else None
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    a_in = a if lead is None else a.reshape(*lead, K)
    _lib.load().mixdq_force_simt(1 if force_simt else 0)
    try:
        out = ops.qlinear_w8_a8_ohalf(a_in.to(dev), w.to(dev), w_scale.to(dev), a_scale.to(dev),
                                      a_zp.to(dev), wsum.to(dev), scale.to(dev), bias0.to(dev),
                                      None if b is None else b.to(dev), _acc_out=acc)
        torch.cuda.synchronize()
        path = _lib.last_path()
    finally:
        _lib.load().mixdq_force_simt(0)
    ref, ref_acc = O.qlinear_kernel(a_in, w, bias0, scale, b)
    assert torch.equal(acc.cpu().long(), ref_acc), "INT32 accumulators differ"
    assert out.shape == ref.shape and out.dtype == torch.float16
    assert torch.equal(bits(out), bits(ref)), "fp16 outputs differ"
    return path


# SDXL-Turbo shapes + ragged/edge shapes
LINEAR_SHAPES = [(256, 1280, 1280), (256, 10240, 1280), (256, 1280, 5120), (1024, 5120, 640),
                 (1024, 640, 640), (1024, 640, 2560), (77, 1280, 2048), (77, 640, 2048),
                 (1, 1280, 1280), (1, 320, 1280), (1, 1280, 2816), (1, 1280, 320),
                 (300, 200, 336), (129, 264, 48), (128, 16, 16), (5, 8, 16)]


@pytest.mark.parametrize("M,N,K", LINEAR_SHAPES)
def test_qlinear_tcgen05_bit_exact(ops, dev, M, N, K):
    assert _linear_case(ops, dev, M, N, K, bias=(M % 2 == 0)).startswith("tcgen05")


@pytest.mark.parametrize("M,N,K", [(64, 16, 8), (33, 20, 36), (7, 4, 4), (256, 1280, 1280)])
def test_qlinear_simt_bit_exact(ops, dev, M, N, K):
    force = (K % 16 == 0)
    assert _linear_case(ops, dev, M, N, K, force_simt=force) == "simt"


def test_qlinear_batched_leading_dims_and_noncontiguous(ops, dev):
    _linear_case(ops, dev, 2 * 77, 640, 2048, lead=(2, 77))
    # non-contiguous int8 input is densified silently (qlinear.cc:75-77)
    g = torch.Generator().manual_seed(5)
    a = torch.randint(-128, 128, (64, 256), dtype=torch.int8, generator=g)
    w = torch.randint(-128, 128, (32, 128), dtype=torch.int8, generator=g)
    one, zero = torch.ones(32), torch.zeros(32)
    s1 = torch.tensor(1.0)
    out = ops.qlinear_w8_a8_ohalf(a.to(dev)[:, ::2], w.to(dev), one.to(dev), s1.to(dev), s1.to(dev),
                                  zero.to(dev), (one * 0.001).to(dev), zero.to(dev), None)
    ref, _ = O.qlinear_kernel(a[:, ::2], w, zero, one * 0.001, None)
    assert torch.equal(bits(out), bits(ref))


@pytest.mark.parametrize("bn,splits", [(16, 1), (16, 2), (32, 1), (32, 4), (64, 1), (64, 2), (64, 8),
                                       (128, 1), (128, 4), (128, 8), (256, 1), (256, 2), (256, 8)])
def test_qlinear_every_tile_width_and_split(ops, dev, bn, splits):
    """each BN instantiation of the tcgen05 kernel x split-K cluster size, incl. M/N/K tails and
    uneven k-block ranges (K = 400 -> 4 k-blocks, K = 1280 -> 10 k-blocks over 8 ranks)"""
    from mixdq_b200 import _lib
    lib = _lib.load()
    lib.mixdq_debug_force_bn(bn)
    lib.mixdq_debug_force_splits(splits)
    try:
        want = "tcgen05-splitk" if splits > 1 else "tcgen05"
        assert _linear_case(ops, dev, 260, 328, 1280) == want
        assert _linear_case(ops, dev, 128, 512, 1024, bias=False) == want
        if splits <= 4:
            assert _linear_case(ops, dev, 77, 640, 400 + 112) == want
    finally:
        lib.mixdq_debug_force_bn(0)
        lib.mixdq_debug_force_splits(0)


@pytest.mark.parametrize("bn,splits", [(64, 4), (128, 8), (256, 8), (32, 2)])
def test_qconv2d_split_k_clusters(ops, dev, bn, splits):
    from mixdq_b200 import _lib
    lib = _lib.load()
    lib.mixdq_debug_force_bn(bn)
    lib.mixdq_debug_force_splits(splits)
    try:
        assert _conv_case(ops, dev, 1, 16, 16, 1280, 1280, 3, 3, 1, 1) == "tcgen05-splitk"
        assert _conv_case(ops, dev, 2, 14, 14, 96, 200, 3, 3, 1, 1) == "tcgen05-splitk"
        assert _conv_case(ops, dev, 1, 32, 32, 640, 320, 1, 1, 0, 1).startswith("tcgen05")
    finally:
        lib.mixdq_debug_force_bn(0)
        lib.mixdq_debug_force_splits(0)


def test_qlinear_dynamic_variant(ops, dev):
    g = torch.Generator().manual_seed(9)
    M, N, K = 256, 640, 640
    x = (torch.randn(M, K, generator=g) * 1.3).half()
    w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g)
    w_scale = 0.001 + 0.01 * torch.rand(N, generator=g)
    wsum = w.float().sum(1)
    b = torch.randn(N, generator=g).half()
    q, s, z = ops.quantize_per_tensor_dynamic(x.to(dev))
    out = ops.qlinear_w8_a8_ohalf_dynamic(q, w.to(dev), w_scale.to(dev), s, z, wsum.to(dev), b.to(dev))
    qr, sr, zr = O.quantize_dynamic_kernel(x)
    ref, _ = O.qlinear_kernel(qr, w, wsum * zr, w_scale * sr, b)
    assert torch.equal(bits(out), bits(ref))


@pytest.mark.parametrize("M,N,K", [(256, 1280, 1280), (77, 640, 2048), (3, 8, 32)])
def test_qlinear_w4_packed(ops, dev, M, N, K):
    """W4A8: packed signed nibbles (even k high) == the W8 oracle on the unpacked codes."""
    g = torch.Generator().manual_seed(K)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g)
    codes = torch.randint(-8, 8, (N, K), dtype=torch.int8, generator=g)
    packed = O.pack_int4(codes)
    scale = 0.001 + 0.01 * torch.rand(N, generator=g)
    bias0 = codes.float().sum(1) * 3.0
    b = torch.randn(N, generator=g).half()
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    out = ops.qlinear_w4_a8_ohalf(a.to(dev), packed.to(dev), scale.to(dev), bias0.to(dev), b.to(dev),
                                  _acc_out=acc)
    ref, ref_acc = O.qlinear_kernel(a, codes, bias0, scale, b)
    assert torch.equal(acc.cpu().long(), ref_acc)
    assert torch.equal(bits(out), bits(ref))


def test_qlinear_error_behaviour(ops, dev):
    a = torch.zeros(4, 6, dtype=torch.int8, device=dev)
    w = torch.zeros(8, 6, dtype=torch.int8, device=dev)
    f = torch.zeros(8, device=dev)
    s = torch.tensor(1.0, device=dev)
    with pytest.raises(RuntimeError, match="alignment not to 4"):
        ops.qlinear_w8_a8_ohalf(a, w, f, s, s, f, f, f, None)
    with pytest.raises(RuntimeError, match="input_int8 should be int8 type"):
        ops.qlinear_w8_a8_ohalf(a.half(), w, f, s, s, f, f, f, None)
    with pytest.raises(RuntimeError, match="bias with float16"):
        ops.qlinear_w8_a8_ohalf(a, w, f, s, s, f, f, f, f)
    with pytest.raises(RuntimeError, match="last dimension"):
        ops.qlinear_w8_a8_ohalf(torch.zeros(4, 8, dtype=torch.int8, device=dev), w, f, s, s, f, f, f, None)
    with pytest.raises(RuntimeError, match="weight_scale vector"):
        ops.qlinear_w8_a8_ohalf(a, w, f[:4], s, s, f, f, f, None)


def test_qlinear_full_size_checksum(ops, dev):
    """BASELINE config 3 size (B=8 GEGLU projection, 2048 x 10240 x 1280): a checksum of checksums
    — sum_m acc[m, n] == sum_k (sum_m A[m, k]) * W[n, k] — needs O(MK + NK) CPU work only."""
    g = torch.Generator().manual_seed(1)
    M, N, K = 2048, 10240, 1280
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g)
    w = torch.randint(-128, 128, (N, K), dtype=torch.int8, generator=g)
    one, zero = torch.ones(N), torch.zeros(N)
    s1 = torch.tensor(1.0)
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    ops.qlinear_w8_a8_ohalf(a.to(dev), w.to(dev), one.to(dev), s1.to(dev), s1.to(dev), zero.to(dev),
                            one.to(dev), zero.to(dev), None, _acc_out=acc)
    col = acc.long().sum(dim=0).cpu()
    want = (w.long() * a.long().sum(dim=0)[None, :]).sum(dim=1)
    assert torch.equal(col, want)
    row = acc.long().sum(dim=1).cpu()
    want_r = (a.long() * w.long().sum(dim=0)[None, :]).sum(dim=1)
    assert torch.equal(row, want_r)


# ------------------------------------------------------------------------------------- A3 + A4
def _conv_case(ops, dev, n, h, w, c, k, r, s, pad, stride, bias=True, small_vals=False, force_simt=False,
               seed=0):
    from mixdq_b200 import _lib
    g = torch.Generator().manual_seed(seed * 31 + n * h * w + c + k)
    lo, hi = (-3, 3) if small_vals else (-128, 128)
    x = torch.randint(lo, hi, (n, c, h, w), dtype=torch.int8, generator=g)
    wt = torch.randint(lo, hi, (k, c, r, s), dtype=torch.int8, generator=g)
    w_scale = 0.1 + torch.rand(k, generator=g) if small_vals else 0.001 + 0.01 * torch.rand(k, generator=g)
    a_scale = torch.tensor(0.123)
    a_zp = torch.tensor(2.345 if small_vals else 7.0)
    scale = w_scale * a_scale
    b = torch.rand(k, generator=g).half() if bias else None
    wsum = wt.float().sum(dim=1, keepdim=True) if pad > 0 else None
    bias0 = wt.float().sum(dim=[1, 2, 3]) * a_zp if pad == 0 else None
    P = (h + 2 * pad - r) // stride + 1
    Q = (w + 2 * pad - s) // stride + 1
    acc = torch.empty(n * P * Q, k, dtype=torch.int32, device=dev)
    _lib.load().mixdq_force_simt(1 if force_simt else 0)
    try:
        from mixdq_extension.op.qconv2d import qconv2d  # noqa: F401  (reference entry point exists)
        out = ops.qconv2d_w8_a8_ohalf(
            x.to(dev).contiguous(memory_format=torch.channels_last),
            wt.to(dev).contiguous(memory_format=torch.channels_last),
            w_scale.to(dev), a_scale.to(dev), a_zp.to(dev), scale.to(dev),
            None if wsum is None else wsum.to(dev), None if bias0 is None else bias0.to(dev),
            None if b is None else b.to(dev), stride, pad, 1, _acc_out=acc)
        torch.cuda.synchronize()
        path = _lib.last_path()
    finally:
        _lib.load().mixdq_force_simt(0)
    ref, ref_acc = O.qconv2d_kernel(x, wt, scale, wsum, bias0, a_zp, b, stride, pad)
    got_acc = acc.cpu().view(n, P, Q, k).permute(0, 3, 1, 2).long()
    assert torch.equal(got_acc, ref_acc), "INT32 accumulators differ"
    assert out.shape == (n, k, P, Q) and out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(bits(out.contiguous()), bits(ref)), "fp16 outputs differ"
    return path


REF_CONV_CASES = [  # op/qconv2d.py:104-119 (n,h,w,c,k,r,s,pad,stride,bias)
    (1, 14, 14, 512, 1024, 3, 3, 1, 1, True), (1, 14, 14, 512, 1024, 3, 3, 1, 2, True),
    (1, 14, 14, 512, 1024, 3, 3, 0, 1, False), (1, 14, 14, 512, 1024, 3, 3, 0, 1, True),
    (1, 14, 14, 512, 1024, 3, 3, 0, 2, True),
    (1, 7, 7, 4, 320, 3, 3, 1, 1, True), (1, 7, 7, 4, 320, 3, 3, 0, 1, True),
    (1, 7, 7, 4, 320, 3, 3, 1, 2, True), (1, 7, 7, 4, 320, 3, 3, 0, 2, True),
    (1, 7, 7, 320, 4, 3, 3, 1, 1, True), (1, 7, 7, 320, 4, 3, 3, 0, 1, True),
    (1, 7, 7, 320, 4, 3, 3, 1, 2, True), (1, 7, 7, 320, 4, 3, 3, 0, 2, True),
]


@pytest.mark.parametrize("n,h,w,c,k,r,s,pad,stride,bias", REF_CONV_CASES)
def test_reference_qconv2d_selftest_cases(ops, dev, n, h, w, c, k, r, s, pad, stride, bias):
    """the reference's 13 distinct self-test configurations, its value ranges (randint(-3,3),
    input_scale 0.123, input_zp 2.345), checked bit-exactly instead of at fp16 tolerance"""
    _conv_case(ops, dev, n, h, w, c, k, r, s, pad, stride, bias, small_vals=True)


SDXL_CONVS = [  # (n,h,w,c,k,r,s,pad,stride)
    (1, 64, 64, 320, 320, 3, 3, 1, 1), (1, 32, 32, 640, 640, 3, 3, 1, 1),
    (1, 16, 16, 1280, 1280, 3, 3, 1, 1), (1, 16, 16, 2560, 1280, 3, 3, 1, 1),
    (1, 32, 32, 320, 640, 3, 3, 1, 1), (1, 64, 64, 960, 320, 3, 3, 1, 1),
    (1, 32, 32, 320, 640, 1, 1, 0, 1), (1, 16, 16, 640, 1280, 1, 1, 0, 1),
    (2, 8, 8, 1280, 1280, 3, 3, 1, 1), (3, 16, 16, 640, 640, 3, 3, 1, 1),
    (2, 64, 64, 320, 320, 3, 3, 1, 1), (1, 1, 1, 64, 64, 3, 3, 1, 1), (1, 2, 3, 32, 16, 3, 3, 1, 1),
    (1, 64, 64, 320, 320, 3, 3, 1, 2), (1, 32, 32, 640, 640, 3, 3, 1, 2),   # SDXL downsamplers
    (2, 14, 14, 512, 1024, 3, 3, 1, 2), (1, 15, 13, 64, 64, 3, 3, 1, 2), (3, 14, 14, 128, 64, 3, 3, 0, 2),
    # output rows wider than one 128-pixel tile (latents beyond 1024 px): tiled along q as well
    (1, 3, 160, 32, 32, 3, 3, 1, 1), (2, 2, 131, 64, 16, 1, 1, 0, 1), (1, 5, 300, 16, 32, 3, 3, 1, 2),
]


@pytest.mark.parametrize("n,h,w,c,k,r,s,pad,stride", SDXL_CONVS)
def test_qconv2d_tcgen05_bit_exact(ops, dev, n, h, w, c, k, r, s, pad, stride):
    assert _conv_case(ops, dev, n, h, w, c, k, r, s, pad, stride).startswith("tcgen05")


@pytest.mark.parametrize("n,h,w,c,k,r,s,pad,stride", [
    (1, 64, 64, 4, 320, 3, 3, 1, 1), (1, 64, 64, 320, 4, 3, 3, 1, 1),       # conv_in / conv_out
    (3, 9, 13, 4, 64, 3, 3, 1, 1), (2, 5, 3, 4, 2048, 3, 3, 1, 1), (2, 6, 6, 4, 24, 3, 3, 1, 1),   # 4-channel kernel: ragged q groups, K groups
    (2, 9, 9, 64, 8, 3, 3, 1, 1), (3, 10, 7, 48, 8, 3, 3, 1, 2), (2, 8, 8, 1280, 8, 3, 3, 1, 1),   # few output channels: K = 8 form, weights beyond the shared-memory budget
    (1, 9, 9, 32, 32, 5, 5, 2, 1)])
def test_qconv2d_other_geometries_bit_exact(ops, dev, n, h, w, c, k, r, s, pad, stride):
    _conv_case(ops, dev, n, h, w, c, k, r, s, pad, stride)


def test_qconv2d_simt_cross_check(ops, dev):
    assert _conv_case(ops, dev, 1, 16, 16, 128, 128, 3, 3, 1, 1, force_simt=True) == "simt"


def test_qconv2d_nchw_input_is_converted(ops, dev):
    """int8 NCHW input: silently converted like qconv2d.cc:91-95"""
    g = torch.Generator().manual_seed(2)
    x = torch.randint(-128, 128, (1, 64, 8, 8), dtype=torch.int8, generator=g)
    wt = torch.randint(-128, 128, (32, 64, 3, 3), dtype=torch.int8, generator=g)
    ws = torch.full((32,), 0.01)
    s, z = torch.tensor(0.05), torch.tensor(3.0)
    out = ops.qconv2d_w8_a8_ohalf(x.to(dev), wt.to(dev), ws.to(dev), s.to(dev), z.to(dev),
                                  (ws * s).to(dev), wt.float().sum(1, keepdim=True).to(dev), None,
                                  None, 1, 1, 1)
    ref, _ = O.qconv2d_kernel(x, wt, ws * s, wt.float().sum(1, keepdim=True), None, z, None, 1, 1)
    assert torch.equal(bits(out.contiguous()), bits(ref))


def test_qconv2d_error_behaviour(ops, dev):
    x = torch.zeros(1, 8, 4, 4, dtype=torch.int8, device=dev)
    w = torch.zeros(8, 8, 3, 3, dtype=torch.int8, device=dev)
    f = torch.zeros(8, device=dev)
    s = torch.tensor(1.0, device=dev)
    with pytest.raises(RuntimeError, match="K\\*R\\*S"):
        ops.qconv2d_w8_a8_ohalf(x, w, f, s, s, f, None, f, None, 1, 1, 1)
    with pytest.raises(RuntimeError, match="bias0 should equal"):
        ops.qconv2d_w8_a8_ohalf(x, w, f, s, s, f, None, None, None, 1, 0, 1)
    with pytest.raises(RuntimeError, match="dilation"):
        ops.qconv2d_w8_a8_ohalf(x, w, f, s, s, f, None, f, None, 1, 0, 2)


@pytest.mark.parametrize("n,hw,ca,cb,k", [(1, 16, 1280, 1280, 1280), (1, 32, 1280, 640, 640),
                                          (2, 64, 320, 320, 320), (1, 16, 1280, 640, 1280),
                                          (1, 5, 16, 32, 24)])
def test_split_shortcut_fused(ops, dev, n, hw, ca, cb, k):
    """A6: one dual-accumulator kernel == the reference's two fp16 convs + fp16 add"""
    g = torch.Generator().manual_seed(ca + cb)
    x = torch.randint(-128, 128, (n, ca + cb, hw, hw), dtype=torch.int8, generator=g)
    wa = torch.randint(-128, 128, (k, ca, 1, 1), dtype=torch.int8, generator=g)
    wb = torch.randint(-128, 128, (k, cb, 1, 1), dtype=torch.int8, generator=g)
    sa = (0.001 + 0.01 * torch.rand(k, generator=g)) * 0.05
    sb = (0.001 + 0.01 * torch.rand(k, generator=g)) * 0.07
    b0a = wa.float().sum(dim=[1, 2, 3]) * 3.0
    b0b = wb.float().sum(dim=[1, 2, 3]) * -9.0
    b = torch.rand(k, generator=g).half()
    xd = x.to(dev).contiguous(memory_format=torch.channels_last)
    out = ops.qconv1x1_split_w8_a8_ohalf(xd[:, :ca], wa.to(dev), sa.to(dev), b0a.to(dev),
                                         xd[:, ca:], wb.to(dev), sb.to(dev), b0b.to(dev), b.to(dev))
    o0, _ = O.qconv2d_kernel(x[:, :ca], wa, sa, None, b0a, 0.0, b, 1, 0)
    o1, _ = O.qconv2d_kernel(x[:, ca:], wb, sb, None, b0b, 0.0, None, 1, 0)
    assert torch.equal(bits(out.contiguous()), bits(O.split_shortcut_kernel(o0, o1)))
