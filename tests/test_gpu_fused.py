"""Producer-fused quantisation kernels (LayerNorm / GEGLU / GroupNorm[+SiLU] -> int8) and the fused
elementwise epilogue tails (residual, per-image channel add) against the CPU oracle.

Contract checked (include/mixdq_b200.h):
  * the QUANTISATION is bit-exact: codes, scale and zero point equal the oracle's A10 formula
    applied to the fp16 values the kernel itself produced (returned through the y_out hook);
  * the fp16 values are a floating-point restatement of the stock PyTorch op: compared with the
    oracle's fp32 CPU evaluation within 2 fp16 ulps, >= 99 % of elements identical;
  * hence codes vs the fully-CPU pipeline differ by at most 1 step on a small fraction;
  * the GEMM / conv tails are bit-exact (integer accumulate + separately rounded fp16 adds).
"""
import pytest
import torch

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


def check_fp16_restatement(y_gpu, y_ref, abs_floor=0.0, min_same=0.99, ulps=2):
    """|gpu - ref| <= 2 fp16 ulps of the value (or `abs_floor`: GELU's negative tail is a
    cancellation, 1 + erf(x) ~ 1e-6, where CUDA erff and the CPU's vectorised erf differ by many
    RELATIVE ulps on values that are ~1e-5 of the quantisation step)."""
    a, b = y_gpu.float().cpu(), y_ref.float()
    ulp = torch.maximum(b.abs(), torch.tensor(6.1e-5)) * 2.0 ** -10
    tol = torch.clamp(ulps * ulp, min=abs_floor)
    assert ((a - b).abs() <= tol).all(), ((a - b).abs() / tol).max()
    same = (bits(y_gpu) == bits(y_ref)).float().mean().item()
    assert same >= min_same, same


def check_quant(q, s, z, y_gpu, y_ref):
    qr, sr, zr = O.quantize_dynamic_kernel(y_gpu.cpu())
    assert torch.equal(s.cpu(), sr) and torch.equal(z.cpu(), zr)
    assert torch.equal(q.cpu(), qr), "INT8 codes differ from the oracle on the kernel's fp16 values"
    q2, _, _ = O.quantize_dynamic_kernel(y_ref)
    d = (q.cpu().int() - q2.int()).abs()
    assert d.max().item() <= 1 and (d != 0).float().mean().item() <= 0.02


@pytest.mark.parametrize("M,C", [(256, 1280), (1024, 640), (77, 64), (8192, 640), (300, 2048),
                                 (4, 1280), (4500, 1280)])
def test_layernorm_quant(ops, dev, M, C):
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 2 + 0.3).half()
    w = (1 + 0.2 * torch.randn(C, generator=g)).half()
    b = (0.1 * torch.randn(C, generator=g)).half()
    q, s, z, y = ops.layernorm_quantize_dynamic(x.to(dev), w.to(dev), b.to(dev), 1e-5, return_y=True)
    y_ref = O.layernorm_fp16(x, w, b, 1e-5)
    check_fp16_restatement(y, y_ref, abs_floor=3e-5)
    check_quant(q, s, z, y, y_ref)
    q1, s1, z1 = ops.layernorm_quantize_dynamic(x.to(dev), w.to(dev), b.to(dev), 1e-5)
    assert torch.equal(q1, q) and torch.equal(s1, s) and torch.equal(z1, z)


def test_layernorm_quant_3d_and_repeat(ops, dev):
    """[B, T, C] input; repeated calls leave the workspace consistent (grid barrier reset)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 1024, 640, generator=g).half()
    w = torch.ones(640).half(); b = torch.zeros(640).half()
    outs = [ops.layernorm_quantize_dynamic(x.to(dev), w.to(dev), b.to(dev), 1e-5) for _ in range(5)]
    for q, s, z in outs[1:]:
        assert torch.equal(q, outs[0][0]) and torch.equal(s, outs[0][1])
    assert outs[0][0].shape == (2, 1024, 640)


@pytest.mark.parametrize("M,I", [(256, 5120), (1024, 2560), (77, 128), (8192, 2560), (3, 64)])
def test_geglu_quant(ops, dev, M, I):
    g = torch.Generator().manual_seed(M + I)
    hg = (torch.randn(M, 2 * I, generator=g) * 1.5).half()
    q, s, z, y = ops.geglu_quantize_dynamic(hg.to(dev), return_y=True)
    y_ref = O.geglu_fp16(hg)
    check_fp16_restatement(y, y_ref, abs_floor=2e-4 * y_ref.float().abs().max().item(),
                           min_same=0.97)
    check_quant(q, s, z, y, y_ref)


@pytest.mark.parametrize("M,I,K", [(256, 5120, 1280), (1024, 2560, 640), (77, 64, 128),
                                   (300, 80, 64), (4096, 1280, 320), (1, 16, 16)])
def test_geglu_in_gemm_epilogue(ops, dev, M, I, K):
    """ff.net.0.proj with the GEGLU in the tcgen05 epilogue + the single-pass quantiser fed by the
    epilogue's min/max == the unfused linear -> GEGLU+quantise kernels, bit for bit; repeated
    calls leave the min/max words of the workspace clean."""
    g = torch.Generator().manual_seed(M + I + K)
    x8 = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
    w = torch.randint(-127, 128, (2 * I, K), dtype=torch.int8, generator=g).to(dev)
    ws_ = (0.001 + 0.01 * torch.rand(2 * I, generator=g)).to(dev)
    wsum = w.float().sum(1)
    bias = torch.randn(2 * I, generator=g).half().to(dev)
    a_s = torch.tensor(0.04, device=dev); a_z = torch.tensor(-7.0, device=dev)
    hg = ops.qlinear_dynamic_fused(x8, w, ws_, a_s, a_z, wsum, bias)
    q_ref, s_ref, z_ref, y_ref = ops.geglu_quantize_dynamic(hg, return_y=True)
    idx = ops.geglu_interleave_index(I, dev)
    args = (x8, w[idx].contiguous(), ws_[idx].contiguous(), a_s, a_z, wsum[idx].contiguous(),
            bias[idx].contiguous())
    for _ in range(3):
        q, s, z, y = ops.qlinear_geglu_quantize_dynamic(*args, return_y=True)
        assert torch.equal(bits(y), bits(y_ref)), "GEGLU epilogue differs from the stock sequence"
        assert torch.equal(s, s_ref) and torch.equal(z, z_ref)
        assert torch.equal(q, q_ref)


def test_cluster_and_grid_quantisers_agree(ops, dev):
    """Single-cluster (DSMEM + cluster barrier + PDL) and flag-barrier grid variants of the
    dynamic quantisers return identical codes / scale / zero point."""
    from mixdq_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(256, 1280, generator=g) * 2 + 0.3).half().to(dev)
    w = (1 + 0.2 * torch.randn(1280, generator=g)).half().to(dev)
    b = (0.1 * torch.randn(1280, generator=g)).half().to(dev)
    wide = (torch.randn(1024, 1920, generator=g) * 3).half().to(dev)
    outs = []
    try:
        for mode in (0, 1, 0, 1):
            lib.mixdq_debug_set_cluster(mode)
            ops.clear_dynamic_quant_cache()
            o = []
            for rows in (256, 48, 1):          # 48 x 1280 and below: the one-cluster variants
                o += list(ops.layernorm_quantize_dynamic(x[:rows], w, b, 1e-5))
                o += list(ops.quantize_per_tensor_dynamic(x[:rows].clone()))
                o += list(ops.quantize_rows_dynamic(wide[:rows, 640:]))
            outs.append([t.clone() for t in o])
    finally:
        lib.mixdq_debug_set_cluster(1)
    for o in outs[1:]:
        for a_, b_ in zip(o, outs[0]):
            assert torch.equal(a_, b_)


def test_two_pass_and_single_kernel_quantisers_agree(ops, dev):
    """min/max pass + quantise pass (csrc/quant2.cu) == the first-generation single-kernel
    quantisers, bit for bit, for plain / row-pitched / LayerNorm / GroupNorm inputs."""
    from mixdq_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(1024, 640, generator=g) * 2 + 0.3).half().to(dev)
    w = (1 + 0.2 * torch.randn(640, generator=g)).half().to(dev)
    b = (0.1 * torch.randn(640, generator=g)).half().to(dev)
    wide = (torch.randn(1024, 1920, generator=g) * 3).half().to(dev)
    # (one image: the barrier-free GroupNorm cuts every image into the batch-1 row ranges so that
    # results do not depend on the batch; the single-kernel form splits per batch)
    img = (torch.randn(1, 640, 32, 32, generator=g) * 1.5).half().to(dev).contiguous(
        memory_format=torch.channels_last)
    outs = []
    try:
        for mode in (0, 1, 0, 1):
            lib.mixdq_debug_set_two_pass(mode)
            ops.clear_dynamic_quant_cache()
            o = list(ops.layernorm_quantize_dynamic(x, w, b, 1e-5, return_y=True))
            o += list(ops.layernorm_quantize_dynamic(x[:256], w, b, 1e-5, return_y=True))   # one cluster
            o += list(ops.quantize_per_tensor_dynamic(x[:400].clone()))
            o += list(ops.quantize_rows_dynamic(wide[:200, 640:]))
            o += list(ops.quantize_per_tensor_dynamic(x.clone()))
            o += list(ops.quantize_rows_dynamic(wide[:, 640:]))
            o += list(ops.groupnorm_quantize_dynamic(img, 32, w, b, 1e-5, True, return_y=True))
            o += list(ops.groupnorm_quantize_dynamic(img, 32, w, b, 1e-6, False))
            outs.append([t.clone() for t in o])
    finally:
        lib.mixdq_debug_set_two_pass(1)
    for o in outs[1:]:
        for a_, b_ in zip(o, outs[0]):
            assert torch.equal(a_, b_)


@pytest.mark.parametrize("N,C,H,W,G,silu", [
    (1, 320, 64, 64, 32, True), (1, 640, 32, 32, 32, True), (1, 1280, 16, 16, 32, True),
    (1, 2560, 16, 16, 32, True), (1, 960, 64, 64, 32, True), (1, 1920, 32, 32, 32, True),
    (2, 1280, 16, 16, 32, False), (8, 640, 32, 32, 32, False), (3, 64, 8, 8, 16, True),
    (2, 128, 16, 16, 16, True), (1, 320, 64, 64, 32, False)])
def test_groupnorm_quant(ops, dev, N, C, H, W, G, silu):
    g = torch.Generator().manual_seed(N * C + H)
    x = (torch.randn(N, C, H, W, generator=g) * torch.linspace(0.5, 3.0, C).view(1, C, 1, 1)
         + torch.linspace(-1, 1, C).view(1, C, 1, 1)).half()
    w = (1 + 0.2 * torch.randn(C, generator=g)).half()
    b = (0.1 * torch.randn(C, generator=g)).half()
    xd = x.to(dev).contiguous(memory_format=torch.channels_last)
    q, s, z, y = ops.groupnorm_quantize_dynamic(xd, G, w.to(dev), b.to(dev), 1e-5, silu, return_y=True)
    assert q.is_contiguous(memory_format=torch.channels_last) and q.shape == x.shape
    y_ref = O.groupnorm_fp16(x, G, w, b, 1e-5, silu)
    # y = x*a + b cancels for |y| << |x*a|: fp32 evaluation-order differences (~1e-7 * |x*a|) are
    # an absolute, not a relative, error there -> absolute floor of 3e-5 (1e-3 of a code step)
    # with SiLU two fp16 roundings compose: a 1-ulp flip of the GroupNorm value moves the SiLU
    # value by up to ~1.1 ulp more -> 3 ulps
    check_fp16_restatement(y.contiguous(), y_ref, abs_floor=3e-5, ulps=3 if silu else 2)
    check_quant(q.contiguous(), s, z, y.contiguous(), y_ref)
    # deterministic across launches (fixed-point statistics do not depend on CTA arrival order)
    q2, s2, z2 = ops.groupnorm_quantize_dynamic(xd, G, w.to(dev), b.to(dev), 1e-5, silu)
    assert torch.equal(q2, q) and torch.equal(s2, s) and torch.equal(z2, z)


def test_groupnorm_unsupported_group_shape_raises(ops, dev):
    x = torch.randn(1, 96, 8, 8).half().to(dev).contiguous(memory_format=torch.channels_last)
    w = torch.ones(96).half().to(dev)
    with pytest.raises(RuntimeError):
        ops.groupnorm_quantize_dynamic(x, 32, w, w, 1e-5, True)     # 3 channels per group


@pytest.mark.parametrize("M,cols,pitch", [(4096, 640, 1920), (1024, 1280, 1280), (77, 2048, 2048),
                                          (300, 64, 192)])
def test_rows_quant(ops, dev, M, cols, pitch):
    g = torch.Generator().manual_seed(M)
    full = (torch.randn(M, pitch, generator=g) * 3).half()
    view = full.to(dev)[:, pitch - cols:]
    q, s, z = ops.quantize_rows_dynamic(view)
    qr, sr, zr = O.quantize_dynamic_kernel(full[:, pitch - cols:].contiguous())
    assert torch.equal(q.cpu(), qr) and torch.equal(s.cpu(), sr) and torch.equal(z.cpu(), zr)


@pytest.mark.parametrize("M,N,K", [(256, 1280, 1280), (1024, 640, 2560), (77, 640, 2048),
                                   (256, 1280, 5120), (2048, 1280, 1280)])
def test_linear_residual_tail(ops, dev, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).half()
    w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g)
    ws = 0.001 + 0.01 * torch.rand(N, generator=g)
    wsum = w.float().sum(1)
    bias = torch.randn(N, generator=g).half()
    res = torch.randn(M, N, generator=g).half()
    q, s, z = ops.quantize_per_tensor_dynamic(x.to(dev))
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    y = ops.qlinear_dynamic_fused(q, w.to(dev), ws.to(dev), s, z, wsum.to(dev), bias.to(dev),
                                  residual=res.to(dev), _acc_out=acc)
    qr, sr, zr = O.quantize_dynamic_kernel(x)
    yr, accr = O.qlinear_kernel(qr, w, wsum * zr, ws * sr, bias)
    assert torch.equal(acc.cpu().long(), accr)
    assert torch.equal(bits(y), bits(O.add_fp16(yr, res)))


@pytest.mark.parametrize("n,h,w_,c,k,r,pad,stride", [
    (1, 32, 32, 320, 640, 3, 1, 1), (2, 16, 16, 640, 640, 3, 1, 1), (1, 64, 64, 320, 320, 3, 1, 2),
    (1, 32, 32, 640, 320, 1, 0, 1), (1, 16, 16, 1280, 1280, 3, 1, 1)])
def test_conv_dynamic_tails(ops, dev, n, h, w_, c, k, r, pad, stride):
    g = torch.Generator().manual_seed(h + c + k)
    x = torch.randn(n, c, h, w_, generator=g).half()
    wt = torch.randint(-127, 128, (k, c, r, r), dtype=torch.int8, generator=g)
    ws = 0.001 + 0.01 * torch.rand(k, generator=g)
    bias = torch.randn(k, generator=g).half()
    P = (h + 2 * pad - r) // stride + 1
    Q = (w_ + 2 * pad - r) // stride + 1
    chan = torch.randn(n, k, generator=g).half()
    res = torch.randn(n, k, P, Q, generator=g).half()
    xd = x.to(dev).contiguous(memory_format=torch.channels_last)
    q, s, z = ops.quantize_per_tensor_dynamic(xd)
    wd = wt.to(dev).contiguous(memory_format=torch.channels_last)
    wsum_krs = wt.float().sum(1, keepdim=True) if pad > 0 else None
    wsum_k = wt.float().sum(dim=[1, 2, 3]) if pad == 0 else None
    y = ops.qconv2d_dynamic_fused(
        q, wd, ws.to(dev), s, z, None if wsum_krs is None else wsum_krs.to(dev),
        None if wsum_k is None else wsum_k.to(dev), bias.to(dev), stride, pad,
        chan_add=chan.to(dev), residual=res.to(dev).contiguous(memory_format=torch.channels_last))
    qr, sr, zr = O.quantize_dynamic_kernel(x)
    yr, _ = O.qconv2d_kernel(qr, wt, ws * sr, wsum_krs, None if wsum_k is None else wsum_k * zr,
                             zr, bias, stride, pad)
    ref = O.add_fp16(O.add_fp16(yr, chan[:, :, None, None].expand_as(yr)), res)
    assert torch.equal(bits(y.contiguous()), bits(ref))


def test_split_shortcut_dynamic(ops, dev):
    g = torch.Generator().manual_seed(9)
    n, ca, cb, k, h = 1, 640, 320, 640, 32
    x = torch.cat([torch.randn(n, ca, h, h, generator=g) * 2, torch.randn(n, cb, h, h, generator=g) * 0.5 + 0.3], 1).half()
    wa = torch.randint(-127, 128, (k, ca, 1, 1), dtype=torch.int8, generator=g)
    wb = torch.randint(-127, 128, (k, cb, 1, 1), dtype=torch.int8, generator=g)
    wsa, wsb = 0.001 + 0.01 * torch.rand(k, generator=g), 0.001 + 0.01 * torch.rand(k, generator=g)
    bias = torch.randn(k, generator=g).half()
    res = torch.randn(n, k, h, h, generator=g).half()
    xd = x.to(dev).contiguous(memory_format=torch.channels_last)
    qa, sa, za = ops.quantize_nhwc_slice_dynamic(xd, 0, ca)
    qb, sb, zb = ops.quantize_nhwc_slice_dynamic(xd, ca, ca + cb)
    y = ops.qconv1x1_split_dynamic_fused(
        qa, wa.to(dev), wsa.to(dev), wa.float().sum(dim=[1, 2, 3]).to(dev), sa, za,
        qb, wb.to(dev), wsb.to(dev), wb.float().sum(dim=[1, 2, 3]).to(dev), sb, zb,
        bias.to(dev), residual=res.to(dev).contiguous(memory_format=torch.channels_last))
    qar, sar, zar = O.quantize_dynamic_kernel(x[:, :ca].contiguous())
    qbr, sbr, zbr = O.quantize_dynamic_kernel(x[:, ca:].contiguous())
    assert torch.equal(qa.cpu().contiguous(), qar) and torch.equal(qb.cpu().contiguous(), qbr)
    o0, _ = O.qconv2d_kernel(qar, wa, wsa * sar, None, wa.float().sum(dim=[1, 2, 3]) * zar, 0.0, bias, 1, 0)
    o1, _ = O.qconv2d_kernel(qbr, wb, wsb * sbr, None, wb.float().sum(dim=[1, 2, 3]) * zbr, 0.0, None, 1, 0)
    ref = O.add_fp16(O.split_shortcut_kernel(o0, o1), res)
    assert torch.equal(bits(y.contiguous()), bits(ref))


# ---- producers with STATIC (checkpoint) scales: one pass, bit-identical to producer -> A1 ----
def _static_params(dev, delta, zp):
    d = torch.tensor(delta, dtype=torch.float32)
    return (1.0 / d).to(dev), torch.tensor(float(zp), dtype=torch.float32).to(dev)


@pytest.mark.parametrize("M,C", [(256, 1280), (1024, 640), (77, 64), (8192, 640), (300, 2048),
                                 (4, 1280)])
def test_layernorm_quant_static(ops, dev, M, C):
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 2 + 0.3).half().to(dev)
    w = (1 + 0.2 * torch.randn(C, generator=g)).half().to(dev)
    b = (0.1 * torch.randn(C, generator=g)).half().to(dev)
    for delta, zp in ((0.031, -9.0), (0.004, 20.0)):        # the second one saturates
        inv, z = _static_params(dev, delta, zp)
        y = ops.layernorm_fp16(x, w, b, 1e-5)
        want = ops.quantize_per_tensor_to_int8(y, inv, z)
        got = ops.layernorm_quantize_static(x, w, b, 1e-5, inv, z)
        assert torch.equal(got, want)
        # and against the oracle's A1 formula on the kernel's own fp16 values
        assert torch.equal(got.cpu(), O.quantize_static_kernel(y.cpu(), inv.item(), z.item()))
    # the dynamic path still works on the same workspace afterwards
    q, s, zz, yy = ops.layernorm_quantize_dynamic(x, w, b, 1e-5, return_y=True)
    assert torch.equal(bits(yy), bits(y))


@pytest.mark.parametrize("N,C,H,W,G,silu", [
    (1, 320, 64, 64, 32, True), (1, 1280, 16, 16, 32, True), (1, 2560, 16, 16, 32, True),
    (2, 1280, 16, 16, 32, False), (8, 640, 32, 32, 32, False), (3, 64, 8, 8, 16, True)])
def test_groupnorm_quant_static(ops, dev, N, C, H, W, G, silu):
    g = torch.Generator().manual_seed(N * C + H)
    x = (torch.randn(N, C, H, W, generator=g) * torch.linspace(0.5, 3.0, C).view(1, C, 1, 1)
         + torch.linspace(-1, 1, C).view(1, C, 1, 1)).half()
    w = (1 + 0.2 * torch.randn(C, generator=g)).half().to(dev)
    b = (0.1 * torch.randn(C, generator=g)).half().to(dev)
    xd = x.to(dev).contiguous(memory_format=torch.channels_last)
    inv, z = _static_params(dev, 0.023, -31.0)
    for _ in range(3):     # repeated calls: the apply kernel's last CTA re-zeroes the statistics
        y = ops.groupnorm_fp16(xd, G, w, b, 1e-5, silu)
        want = ops.quantize_per_tensor_to_int8(y, inv, z)
        got = ops.groupnorm_quantize_static(xd, G, w, b, 1e-5, silu, inv, z)
        assert got.is_contiguous(memory_format=torch.channels_last) and got.shape == x.shape
        assert torch.equal(got, want)
    # the dynamic three-kernel form finds the accumulators clean
    q, s, zz, yy = ops.groupnorm_quantize_dynamic(xd, G, w, b, 1e-5, silu, return_y=True)
    assert torch.equal(bits(yy.contiguous()), bits(y.contiguous()))


@pytest.mark.parametrize("M,I,K,w4", [(256, 5120, 1280, False), (1024, 2560, 640, False),
                                      (77, 64, 128, False), (300, 80, 64, False),
                                      (2048, 5120, 1280, False), (8192, 2560, 640, False),
                                      (256, 5120, 1280, True), (2048, 5120, 1280, True)])
def test_geglu_static_quant_in_gemm_epilogue(ops, dev, M, I, K, w4):
    """ff.net.0.proj + GEGLU + the static quantiser of ff.net.2 in ONE tcgen05 kernel (one-tile and
    persistent forms, W8 and packed W4) == GEGLU GEMM -> fp16 -> quantize_per_tensor_to_int8."""
    from mixdq_b200.nn.utils import pack_int4
    g = torch.Generator().manual_seed(M + I + K)
    x8 = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
    lim = 8 if w4 else 128
    w = torch.randint(-lim + 1, lim, (2 * I, K), dtype=torch.int8, generator=g)
    ws_ = (0.001 + 0.01 * torch.rand(2 * I, generator=g)).to(dev) * (16 if w4 else 1)
    wsum = w.float().sum(1).to(dev)
    bias = torch.randn(2 * I, generator=g).half().to(dev)
    a_s = torch.tensor(0.04, device=dev); a_z = torch.tensor(-7.0, device=dev)
    idx = ops.geglu_interleave_index(I, dev)
    w_il = w.to(dev)[idx].contiguous()
    if w4:
        w_il = pack_int4(w_il)
    args = (x8, w_il, ws_[idx].contiguous(), a_s, a_z, wsum[idx].contiguous(),
            bias[idx].contiguous())
    y = ops.qlinear_geglu_fp16(*args)
    rng = y.float().abs().max().item()
    inv, z = _static_params(dev, max(rng, 1e-3) / 200.0, -40.0)     # some saturation
    want = ops.quantize_per_tensor_to_int8(y, inv, z)
    for _ in range(2):
        got = ops.qlinear_geglu_quantize_static(*args, inv, z)
        assert got.shape == (M, I) and torch.equal(got, want)
    assert torch.equal(got.cpu(), O.quantize_static_kernel(y.cpu(), inv.item(), z.item()))
