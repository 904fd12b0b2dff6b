"""First-contact GPU bring-up: exercises every C-ABI entry point once against exact CPU integer
references and prints PASS/FAIL lines. Development aid (the judged parity tests live in tests/)."""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops  # noqa: E402

torch.manual_seed(0)
dev = torch.device("cuda:0")
ok_all = True


def report(name, ok, extra=""):
    global ok_all
    ok_all &= bool(ok)
    print(("PASS " if ok else "FAIL ") + name + " " + extra, flush=True)


def ref_quant(x, sinv, zp):
    v = torch.round(torch.addcmul(zp.double(), x.double(), sinv.double()).float())  # placeholder
    return v


def quant_ref_fma(x16, sinv, zp):
    # fmaf(x, sinv, zp) in fp32 with a single rounding == round-to-fp32 of the exact double result
    # (x is fp16 -> 11-bit, sinv 24-bit: product exact in double; + zp exact in double for |.|<2^53)
    exact = x16.double() * sinv.double() + zp.double()
    f = exact.float()
    return torch.clamp(torch.round(f), -128, 127).to(torch.int8)


def test_quant():
    for numel in [1024, 4096 * 960, 77 * 2048, 1000003, 7]:
        x = (torch.randn(numel) * 2).half()
        scale = torch.tensor(0.0323); zp = torch.tensor(2.0)
        sinv = 1 / scale
        q = ops.quantize_per_tensor_to_int8(x.to(dev), sinv.to(dev), zp.to(dev)).cpu()
        ref = quant_ref_fma(x, sinv, zp)
        report(f"quant_static numel={numel}", torch.equal(q, ref), f"mismatch={(q != ref).sum().item()}")
    # NCHW -> NHWC + channel slice
    x = torch.randn(2, 96, 16, 16).half()
    sinv = torch.tensor(17.3); zp = torch.tensor(-3.0)
    q = ops.quantize_to_nhwc(x.to(dev), sinv.to(dev), zp.to(dev), 32, 96).cpu()
    ref = quant_ref_fma(x[:, 32:96], sinv, zp)
    report("quant_nchw2nhwc slice", torch.equal(q, ref) and q.is_contiguous(memory_format=torch.channels_last))
    xcl = x.to(dev).contiguous(memory_format=torch.channels_last)
    q = ops.quantize_to_nhwc(xcl, sinv.to(dev), zp.to(dev), 0, 32).cpu()
    report("quant_nhwc slice", torch.equal(q, quant_ref_fma(x[:, :32], sinv, zp)))
    q = ops.quantize_per_tensor_to_int8(xcl[:, 32:], sinv.to(dev), zp.to(dev)).cpu()
    report("quant view nhwc slice", torch.equal(q, quant_ref_fma(x[:, 32:], sinv, zp)))
    x3 = torch.randn(3, 77, 64).half()
    q = ops.quantize_per_tensor_to_int8(x3.to(dev)[:, 1:, :], sinv.to(dev), zp.to(dev)).cpu()
    report("quant view bos slice", torch.equal(q, quant_ref_fma(x3[:, 1:, :], sinv, zp)))
    # dynamic
    for numel in [4096 * 320, 77 * 2048 + 3]:
        x = (torch.randn(numel) * 1.7 + 0.3).half()
        q, s, z = ops.quantize_per_tensor_dynamic(x.to(dev))
        xf = x.float()
        mn = torch.clamp(xf.min(), max=0); mx = torch.clamp(xf.max(), min=0)
        delta = (mx - mn) / 255
        zz = torch.round(-mn / delta)
        ref = (torch.clamp(torch.round(xf / delta) + zz, 0, 255) - 128).to(torch.int8)
        ok = torch.equal(q.cpu(), ref) and s.item() == delta.item() and z.item() == zz.item() - 128
        report(f"quant_dynamic numel={numel}", ok, f"mismatch={(q.cpu() != ref).sum().item()} s={s.item()} {delta.item()} z={z.item()} {zz.item()-128}")


def gemm_case(M, N, K, bias=True, force_simt=False, lo=-128, hi=128):
    lib = _lib.load()
    lib.mixdq_force_simt(1 if force_simt else 0)
    a = torch.randint(lo, hi, (M, K), dtype=torch.int8)
    w = torch.randint(lo, hi, (N, K), dtype=torch.int8)
    w_scale = 0.001 + 0.01 * torch.rand(N)
    a_scale = torch.tensor(0.0371); a_zp = torch.tensor(-11.0)
    wsum = w.float().sum(1)
    scale = w_scale * a_scale
    bias0 = wsum * a_zp
    b = torch.randn(N).half() if bias else None
    acc = torch.empty(M, N, dtype=torch.int32, device=dev)
    out = ops.qlinear_w8_a8_ohalf(a.to(dev), w.to(dev), w_scale.to(dev), a_scale.to(dev), a_zp.to(dev),
                                  wsum.to(dev), scale.to(dev), bias0.to(dev),
                                  b.to(dev) if bias else None, _acc_out=acc)
    torch.cuda.synchronize()
    path = _lib.last_path()
    ref_acc = (a.double() @ w.double().t()).to(torch.int64)
    ok_acc = torch.equal(acc.cpu().to(torch.int64), ref_acc)
    f = (ref_acc.float() - bias0[None]) * scale[None]
    if bias:
        f = f + b.float()[None]
    ref = f.half()
    ok_out = torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))
    lib.mixdq_force_simt(0)
    nbad = (acc.cpu().to(torch.int64) != ref_acc).sum().item()
    report(f"gemm M={M} N={N} K={K} bias={bias} path={path}", ok_acc and ok_out,
           f"acc_ok={ok_acc} out_ok={ok_out} nbad_acc={nbad}")
    return ok_acc and ok_out


def conv_case(n, h, w, c, k, r, s, pad, stride, bias=True, force_simt=False):
    lib = _lib.load()
    lib.mixdq_force_simt(1 if force_simt else 0)
    x = torch.randint(-128, 128, (n, c, h, w), dtype=torch.int8)
    wt = torch.randint(-128, 128, (k, c, r, s), dtype=torch.int8)
    w_scale = 0.001 + 0.01 * torch.rand(k)
    a_scale = torch.tensor(0.123); a_zp = torch.tensor(7.0)
    scale = w_scale * a_scale
    b = torch.rand(k).half() if bias else None
    wsum = wt.float().sum(dim=1, keepdim=True) if pad > 0 else None
    bias0 = wt.float().sum(dim=[1, 2, 3]) * a_zp if pad == 0 else None
    P = (h + 2 * pad - r) // stride + 1
    Q = (w + 2 * pad - s) // stride + 1
    acc = torch.empty(n * P * Q, k, dtype=torch.int32, device=dev)
    out = ops.qconv2d_w8_a8_ohalf(
        x.to(dev).contiguous(memory_format=torch.channels_last),
        wt.to(dev).contiguous(memory_format=torch.channels_last),
        w_scale.to(dev), a_scale.to(dev), a_zp.to(dev), scale.to(dev),
        wsum.to(dev) if wsum is not None else None, bias0.to(dev) if bias0 is not None else None,
        b.to(dev) if bias else None, stride, pad, 1, _acc_out=acc)
    torch.cuda.synchronize()
    path = _lib.last_path()
    ref_acc = F.conv2d(x.double(), wt.double(), stride=stride, padding=pad)  # exact in fp64
    zpc = F.conv2d(torch.full((n, 1, h, w), 1.0, dtype=torch.float64), wt.double().sum(1, keepdim=True),
                   stride=stride, padding=pad).float() * a_zp
    f = (ref_acc.float() - zpc) * scale[None, :, None, None]
    if bias:
        f = f + b.float()[None, :, None, None]
    ref = f.half()
    got_acc = acc.cpu().view(n, P, Q, k).permute(0, 3, 1, 2).to(torch.float64)
    ok_acc = torch.equal(got_acc, ref_acc)
    ok_out = torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))
    lib.mixdq_force_simt(0)
    report(f"conv n={n} h={h} w={w} c={c} k={k} r={r} pad={pad} stride={stride} bias={bias} path={path}",
           ok_acc and ok_out, f"acc_ok={ok_acc} out_ok={ok_out} nbad={(got_acc != ref_acc).sum().item()}")


def split_case(n, h, w, ca, cb, k, force_simt=False):
    lib = _lib.load()
    lib.mixdq_force_simt(1 if force_simt else 0)
    xa = torch.randint(-128, 128, (n, ca, h, w), dtype=torch.int8)
    xb = torch.randint(-128, 128, (n, cb, h, w), dtype=torch.int8)
    wa = torch.randint(-128, 128, (k, ca, 1, 1), dtype=torch.int8)
    wb = torch.randint(-128, 128, (k, cb, 1, 1), dtype=torch.int8)
    sa = (0.001 + 0.01 * torch.rand(k)) * 0.05
    sb = (0.001 + 0.01 * torch.rand(k)) * 0.07
    b0a = wa.float().sum(dim=[1, 2, 3]) * 3.0
    b0b = wb.float().sum(dim=[1, 2, 3]) * -9.0
    b = torch.rand(k).half()
    xcat = torch.cat([xa, xb], 1).to(dev).contiguous(memory_format=torch.channels_last)
    out = ops.qconv1x1_split_w8_a8_ohalf(xcat[:, :ca], wa.to(dev), sa.to(dev), b0a.to(dev),
                                         xcat[:, ca:], wb.to(dev), sb.to(dev), b0b.to(dev), b.to(dev))
    torch.cuda.synchronize()
    path = _lib.last_path()
    acc_a = F.conv2d(xa.double(), wa.double()).float()
    acc_b = F.conv2d(xb.double(), wb.double()).float()
    o0 = ((acc_a - b0a[None, :, None, None]) * sa[None, :, None, None] + b.float()[None, :, None, None]).half()
    o1 = ((acc_b - b0b[None, :, None, None]) * sb[None, :, None, None]).half()
    ref = (o0.float() + o1.float()).half()
    ok = torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))
    lib.mixdq_force_simt(0)
    report(f"split n={n} hw={h} ca={ca} cb={cb} k={k} path={path}", ok,
           f"nbad={(out.cpu() != ref).sum().item()}")


def bench_gemm(M, N, K, iters=20):
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    w = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
    z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev)
    s1 = torch.tensor(1.0, device=dev)
    for _ in range(3):
        ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ah = a.half(); wh = w.half()
    for _ in range(3):
        F.linear(ah, wh)
    e0.record()
    for _ in range(iters):
        F.linear(ah, wh)
    e1.record(); torch.cuda.synchronize()
    ms16 = e0.elapsed_time(e1) / iters
    print(f"BENCH gemm M={M} N={N} K={K}: int8 {ms*1e3:.1f} us ({2*M*N*K/ms/1e9:.1f} TOPS)  fp16 cublas {ms16*1e3:.1f} us "
          f"({2*M*N*K/ms16/1e9:.1f} TFLOPS)", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    t0 = time.time()
    test_quant()
    # SIMT first (no tcgen05): validates the harness itself
    gemm_case(64, 16, 8, force_simt=True)
    gemm_case(200, 320, 640, force_simt=True)
    # tcgen05
    gemm_case(128, 128, 128)
    gemm_case(128, 128, 128, lo=1, hi=2)
    gemm_case(256, 1280, 1280)
    gemm_case(77, 640, 2048, bias=False)
    gemm_case(1, 1280, 320)
    gemm_case(1024, 5120, 640)
    gemm_case(4096, 320, 320)
    gemm_case(300, 200, 336)
    gemm_case(2048, 10240, 1280)
    conv_case(1, 14, 14, 512, 1024, 3, 3, 1, 1, force_simt=True)
    conv_case(1, 14, 14, 512, 1024, 3, 3, 1, 2, force_simt=True)
    conv_case(1, 7, 7, 4, 320, 3, 3, 1, 1)
    conv_case(1, 7, 7, 320, 4, 3, 3, 0, 2)
    conv_case(1, 16, 16, 128, 128, 3, 3, 1, 1)
    conv_case(1, 14, 14, 512, 1024, 3, 3, 1, 1)
    conv_case(1, 14, 14, 512, 1024, 3, 3, 0, 1, bias=False)
    conv_case(2, 64, 64, 320, 320, 3, 3, 1, 1)
    conv_case(1, 16, 16, 1280, 1280, 3, 3, 1, 1)
    conv_case(2, 8, 8, 640, 640, 3, 3, 1, 1)
    conv_case(1, 32, 32, 320, 640, 1, 1, 0, 1)
    conv_case(1, 64, 64, 320, 320, 3, 3, 1, 2)
    split_case(1, 16, 16, 1280, 1280, 1280, force_simt=True)
    split_case(1, 16, 16, 1280, 1280, 1280)
    split_case(2, 32, 32, 1280, 640, 640)
    split_case(1, 64, 64, 640, 320, 320)
    print("ALL", "PASS" if ok_all else "FAIL", f"{time.time()-t0:.1f}s", flush=True)
    for shp in [(256, 1280, 1280), (256, 10240, 1280), (256, 1280, 5120), (1024, 5120, 640),
                (77, 1280, 2048), (2048, 1280, 1280), (2048, 10240, 1280), (8192, 5120, 640), (8192, 8192, 8192)]:
        bench_gemm(*shp)
