"""The persistent tcgen05 kernel (csrc/tc_persist.cuh: continuous TMA ring, two TMEM accumulator
slots, clusters of 2 CTAs sharing the weight tile through TMA multicast, W4 converter warps) against
the same oracle as the one-tile-per-CTA kernel: INT32 accumulators and fp16 outputs bit-exact.
Mode 2 of mixdq_debug_set_persist forces it onto small shapes so that ragged M / N / K tails, odd
tile counts (an all-out-of-bounds tile in the last pair) and the tile -> slot / ring phase wrap-around
are checked exactly on the CPU; the heuristic (mode 1) is exercised at BASELINE config 3 sizes through
checksums."""
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
def ops():
    from mixdq_b200 import ops as _ops
    return _ops


@pytest.fixture(params=[1, 2], ids=["cs1", "cs2"])
def forced(request):
    from mixdq_b200 import _lib
    lib = _lib.load()
    lib.mixdq_debug_set_persist(2, request.param)
    yield request.param
    lib.mixdq_debug_set_persist(1, 2)


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int16)


def _linear(ops, dev, M, N, K, w4=False, bias=True, residual=False, dynamic=True):
    from mixdq_b200 import _lib
    g = torch.Generator().manual_seed(M * 3 + N * 5 + K)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g)
    lo, hi = (-8, 8) if w4 else (-128, 128)
    codes = torch.randint(lo, hi, (N, K), dtype=torch.int8, generator=g)
    w_dev = (O.pack_int4(codes) if w4 else codes).to(dev)
    w_scale = 0.001 + 0.01 * torch.rand(N, generator=g)
    a_scale, a_zp = torch.tensor(0.0371), torch.tensor(-11.0)
    wsum = codes.float().sum(1)
    b = torch.randn(N, generator=g).half() if bias else None
    res = torch.randn(M, N, generator=g).half() if residual else None
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    out = ops.qlinear_dynamic_fused(a.to(dev), w_dev, w_scale.to(dev), a_scale.to(dev), a_zp.to(dev),
                                    wsum.to(dev), None if b is None else b.to(dev),
                                    None if res is None else res.to(dev), _acc_out=acc)
    torch.cuda.synchronize()
    path = _lib.last_path()
    ref, ref_acc = O.qlinear_kernel(a, codes, wsum * a_zp, w_scale * a_scale, b)
    if res is not None:
        ref = (ref.float() + res.float()).half()
    assert torch.equal(acc.cpu().long(), ref_acc), "INT32 accumulators differ"
    assert torch.equal(bits(out), bits(ref)), "fp16 outputs differ"
    return path


# (M, N, K): tails in every dimension; tile counts 1, odd, > #SM (several tiles per CTA -> both TMEM
# slots and several ring wrap-arounds), K shorter and longer than the ring
SHAPES = [(128, 256, 128), (130, 520, 400), (777, 1280, 640), (256, 10240, 256), (2100, 2048, 96),
          (4000, 1280, 256), (384, 320, 1280), (640, 640, 2560)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_persistent_linear_bit_exact(ops, dev, forced, M, N, K):
    assert _linear(ops, dev, M, N, K, bias=(M % 2 == 0), residual=(N % 512 == 0)) == "tcgen05-persist"


@pytest.mark.parametrize("M,N,K", [(130, 520, 416), (777, 1280, 640), (4000, 1280, 256), (384, 320, 1280)])
def test_persistent_linear_w4_bit_exact(ops, dev, forced, M, N, K):
    assert _linear(ops, dev, M, N, K, w4=True, residual=True) == "tcgen05-w4-persist"


CONVS = [  # (n,h,w,c,k,r,s,pad,stride)
    (2, 64, 64, 320, 320, 3, 3, 1, 1), (8, 16, 16, 1280, 1280, 3, 3, 1, 1), (3, 32, 32, 640, 640, 3, 3, 1, 2),
    (5, 14, 14, 96, 200, 3, 3, 1, 1), (2, 32, 32, 1280, 640, 1, 1, 0, 1), (3, 7, 9, 96, 40, 3, 3, 1, 1),
]


# 3x3 / pad 1 / stride 1 on CTA pairs: the HALO form (one A box with a one-row halo for the three
# vertical taps). Ragged tiles along P (24x24, 40x40), tiles narrower than 128 rows (48x48), 8-wide rows,
# channel counts that are not a multiple of the 128-byte k-block, ragged N tiles (k = 200).
HALO_CONVS = [(2, 64, 64, 320, 320), (8, 16, 16, 1280, 1280), (4, 32, 32, 640, 320), (3, 24, 24, 96, 200),
              (2, 48, 48, 160, 160), (2, 16, 8, 320, 640), (2, 40, 40, 64, 96), (6, 16, 16, 2560, 1280)]


@pytest.mark.parametrize("n,h,w,c,k", HALO_CONVS)
@pytest.mark.parametrize("halo", [1, 0], ids=["halo", "plain"])
def test_persistent_conv3x3_halo_bit_exact(ops, dev, n, h, w, c, k, halo):
    from mixdq_b200 import _lib
    lib = _lib.load()
    lib.mixdq_debug_set_persist(2, 2)
    lib.mixdq_debug_set_conv_halo(halo)
    try:
        _conv_case(ops, dev, n, h, w, c, k, 3, 3, 1, 1, False,
                   "tcgen05-persist-halo" if halo else "tcgen05-persist")
    finally:
        lib.mixdq_debug_set_conv_halo(1)
        lib.mixdq_debug_set_persist(1, 2)


@pytest.mark.parametrize("n,h,w,c,k,r,s,pad,stride", CONVS)
@pytest.mark.parametrize("w4", [False, True])
def test_persistent_conv_bit_exact(ops, dev, forced, n, h, w, c, k, r, s, pad, stride, w4):
    _conv_case(ops, dev, n, h, w, c, k, r, s, pad, stride, w4, None)


def _conv_case(ops, dev, n, h, w, c, k, r, s, pad, stride, w4, want_path):
    from mixdq_b200 import _lib
    from mixdq_b200.nn.utils import pack_int4
    g = torch.Generator().manual_seed(n * h * w + c + k)
    x = torch.randint(-128, 128, (n, c, h, w), dtype=torch.int8, generator=g)
    lo, hi = (-8, 8) if w4 else (-128, 128)
    codes = torch.randint(lo, hi, (k, c, r, s), dtype=torch.int8, generator=g)
    wcl = codes.contiguous(memory_format=torch.channels_last)
    w_dev = (pack_int4(wcl, dim=1) if w4 else wcl).to(dev)
    w_scale = 0.001 + 0.01 * torch.rand(k, generator=g)
    a_scale, a_zp = torch.tensor(0.123), torch.tensor(7.0)
    b = torch.rand(k, generator=g).half()
    wsum_krs = codes.float().sum(dim=1, keepdim=True) if pad > 0 else None
    wsum_k = codes.float().sum(dim=[1, 2, 3]) if pad == 0 else None
    P = (h + 2 * pad - r) // stride + 1
    Q = (w + 2 * pad - s) // stride + 1
    ca = torch.randn(n, k, generator=g).half()
    res = torch.randn(n, k, P, Q, generator=g).half()
    acc = torch.empty(n * P * Q, k, dtype=torch.int32, device=dev)
    out = ops.qconv2d_dynamic_fused(x.to(dev).contiguous(memory_format=torch.channels_last), w_dev,
                                    w_scale.to(dev), a_scale.to(dev), a_zp.to(dev),
                                    None if wsum_krs is None else wsum_krs.to(dev),
                                    None if wsum_k is None else wsum_k.to(dev), b.to(dev), stride, pad,
                                    chan_add=ca.to(dev),
                                    residual=res.to(dev).contiguous(memory_format=torch.channels_last),
                                    _acc_out=acc)
    torch.cuda.synchronize()
    if want_path is None:
        assert _lib.last_path() in (("tcgen05-w4-persist",) if w4 else
                                    ("tcgen05-persist", "tcgen05-persist-halo"))
    else:
        assert _lib.last_path() == want_path
    ref, ref_acc = O.qconv2d_kernel(x, codes, w_scale * a_scale, wsum_krs,
                                    None if wsum_k is None else wsum_k * a_zp, a_zp, b, stride, pad)
    ref = (ref.float() + ca.float()[:, :, None, None]).half()
    ref = (ref.float() + res.float()).half()
    got_acc = acc.cpu().view(n, P, Q, k).permute(0, 3, 1, 2).long()
    assert torch.equal(got_acc, ref_acc), "INT32 accumulators differ"
    assert torch.equal(bits(out.contiguous()), bits(ref)), "fp16 outputs differ"


@pytest.mark.parametrize("M,inner,K,w4", [(300, 640, 320, False), (1024, 2560, 640, False),
                                         (300, 640, 320, True), (2048, 5120, 1280, True)])
def test_persistent_geglu_matches_one_tile_kernel(ops, dev, forced, M, inner, K, w4):
    """persistent GEGLU projection (+ per-CTA min/max partials -> quantise pass) == the
    one-tile-per-CTA kernel, which tests/test_gpu_fused.py pins to PyTorch"""
    from mixdq_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(inner + K)
    x = torch.randn(M, K, generator=g).half()
    lo, hi = (-8, 8) if w4 else (-128, 128)
    codes = torch.randint(lo, hi, (2 * inner, K), dtype=torch.int8, generator=g)
    w_scale = 0.002 + 0.01 * torch.rand(2 * inner, generator=g)
    wsum = codes.float().sum(1)
    b = torch.randn(2 * inner, generator=g).half()
    idx = ops.geglu_interleave_index(inner)
    c_il = codes[idx]
    w_dev = (O.pack_int4(c_il) if w4 else c_il).to(dev)
    q8, s, z = ops.quantize_per_tensor_dynamic(x.to(dev))
    args = (q8, w_dev, w_scale[idx].to(dev), s, z, wsum[idx].to(dev), b[idx].to(dev))
    qp, sp, zp, yp = ops.qlinear_geglu_quantize_dynamic(*args, return_y=True)
    lib.mixdq_debug_set_persist(0, 0)
    try:
        q1, s1, z1, y1 = ops.qlinear_geglu_quantize_dynamic(*args, return_y=True)
    finally:
        lib.mixdq_debug_set_persist(2, 0)
    assert torch.equal(bits(yp), bits(y1)) and torch.equal(qp, q1)
    assert sp.item() == s1.item() and zp.item() == z1.item()


def test_heuristic_picks_persistent_at_batch8_sizes(ops, dev):
    """BASELINE config 3 / 5 sizes under the default heuristic: the B=8 GEGLU-sized projection goes
    persistent; checked through a checksum of checksums (O(MK + NK) CPU work)."""
    from mixdq_b200 import _lib
    g = torch.Generator().manual_seed(1)
    M, N, K = 8192, 5120, 640
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g)
    w = torch.randint(-128, 128, (N, K), dtype=torch.int8, generator=g)
    one, zero = torch.ones(N), torch.zeros(N)
    s1 = torch.tensor(1.0)
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    ops.qlinear_w8_a8_ohalf(a.to(dev), w.to(dev), one.to(dev), s1.to(dev), s1.to(dev), zero.to(dev),
                            one.to(dev), zero.to(dev), None, _acc_out=acc)
    assert _lib.last_path() == "tcgen05-persist"
    col = acc.long().sum(dim=0).cpu()
    assert torch.equal(col, (w.long() * a.long().sum(dim=0)[None, :]).sum(dim=1))
    row = acc.long().sum(dim=1).cpu()
    assert torch.equal(row, (a.long() * w.long().sum(dim=0)[None, :]).sum(dim=1))
    # small problems stay on the one-tile-per-CTA kernel
    ops.qlinear_w8_a8_ohalf(a[:256].to(dev), w[:1280].to(dev), one[:1280].to(dev), s1.to(dev), s1.to(dev),
                            zero[:1280].to(dev), one[:1280].to(dev), zero[:1280].to(dev), None)
    assert _lib.last_path() in ("tcgen05", "tcgen05-splitk")
