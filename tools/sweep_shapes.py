"""Per-shape timing of the contraction kernels inside a CUDA graph (launch latency included, as in
the real UNet step), with weights rotated through > L2-size buffers so they stream from HBM.
Development / tuning aid; writes gpurun_out/sweep.json."""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()

LINEAR = [(256, 10240, 1280, 60), (256, 1280, 1280, 372), (256, 1280, 5120, 60),
          (1024, 5120, 640, 10), (1024, 640, 640, 70), (1024, 640, 2560, 10),
          (77, 1280, 2048, 120), (77, 640, 2048, 20), (1, 1280, 1280, 9)]
CONV = [  # n,h,w,c,k,r, count
    (1, 16, 16, 1280, 1280, 3, 10), (1, 64, 64, 320, 320, 3, 7), (1, 32, 32, 640, 640, 3, 6),
    (1, 16, 16, 2560, 1280, 3, 2), (1, 32, 32, 1280, 1280, 3, 1), (1, 64, 64, 640, 640, 3, 1),
    (1, 64, 64, 640, 320, 3, 2), (1, 32, 32, 1920, 640, 3, 1), (1, 64, 64, 960, 320, 3, 1)]


def graph_time(fns, iters=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / len(fns) * 1e3  # us per launch


def sweep_linear(M, N, K, bns, batch_scale=1):
    M = M * batch_scale
    ncopy = max(4, int(300e6 // (N * K)) + 1)
    ncopy = min(ncopy, 64)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(ncopy)]
    z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev)
    s1 = torch.tensor(1.0, device=dev)
    res = {}
    # empty-ish reference: launch overhead of a trivial kernel node in the same graph setting
    tiny = torch.zeros(32, device=dev)
    res["null_kernel"] = graph_time([(lambda: tiny.add_(1.0)) for _ in range(16)])
    for bn in bns:
        if isinstance(bn, tuple):
            lib.mixdq_debug_force_bn(bn[0]); lib.mixdq_debug_force_splits(bn[1])
        else:
            lib.mixdq_debug_force_bn(0); lib.mixdq_debug_force_splits(0)
        outs = []
        fns = [(lambda w=w: outs.append(ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)))
               for w in ws]
        res[str(bn)] = graph_time(fns)
    lib.mixdq_debug_force_bn(0); lib.mixdq_debug_force_splits(0)
    ah = a.half(); whs = [w.half() for w in ws[:8]]
    outs = []
    res["fp16_cublas"] = graph_time([(lambda w=w: outs.append(F.linear(ah, w))) for w in whs])
    return res


def sweep_conv(n, h, w, c, k, r, bns):
    pad = 1 if r == 3 else 0
    ncopy = min(32, max(4, int(300e6 // (k * c * r * r)) + 1))
    x = torch.randint(-128, 128, (n, c, h, w), dtype=torch.int8, device=dev).contiguous(memory_format=torch.channels_last)
    wts = [torch.randint(-128, 128, (k, c, r, r), dtype=torch.int8, device=dev).contiguous(memory_format=torch.channels_last)
           for _ in range(ncopy)]
    o = torch.ones(k, device=dev); s1 = torch.tensor(1.0, device=dev)
    wsum = torch.zeros(k, 1, r, r, device=dev)
    res = {}
    for bn in bns:
        if isinstance(bn, tuple):
            lib.mixdq_debug_force_bn(bn[0]); lib.mixdq_debug_force_splits(bn[1])
        else:
            lib.mixdq_debug_force_bn(0); lib.mixdq_debug_force_splits(0)
        outs = []
        fns = [(lambda wt=wt: outs.append(ops.qconv2d_w8_a8_ohalf(x, wt, o, s1, s1, o, wsum if pad else None,
                                                                   None if pad else o, None, 1, pad, 1)))
               for wt in wts]
        res[str(bn)] = graph_time(fns)
    lib.mixdq_debug_force_bn(0); lib.mixdq_debug_force_splits(0)
    xh = x.half(); whs = [wt.half() for wt in wts[:4]]
    outs = []
    res["fp16_cudnn"] = graph_time([(lambda wt=wt: outs.append(F.conv2d(xh, wt, padding=pad))) for wt in whs])
    return res


if __name__ == "__main__":
    bscale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    out = {"linear": {}, "conv": {}}
    bns = ["auto", (64, 1), (128, 1), (256, 1), (64, 4), (64, 8), (128, 2), (128, 4), (128, 8),
           (256, 2), (256, 4), (256, 8)]
    for (M, N, K, cnt) in LINEAR:
        r = sweep_linear(M, N, K, bns, bscale)
        out["linear"][f"{M*bscale}x{N}x{K}"] = r
        hbm_us = (M * bscale * K + N * K + 2 * M * bscale * N) / 6.54e6
        print(f"linear M={M*bscale} N={N} K={K} x{cnt}: " + " ".join(f"{k}:{v:.1f}" for k, v in r.items())
              + f"  | hbm floor {hbm_us:.2f}us", flush=True)
    for (n, h, w, c, k, r_, cnt) in CONV:
        r = sweep_conv(n * bscale, h, w, c, k, r_, bns)
        out["conv"][f"{n*bscale}x{h}x{w}x{c}->{k}r{r_}"] = r
        hbm_us = (n * bscale * h * w * c + k * c * r_ * r_ + 2 * n * bscale * h * w * k) / 6.54e6
        print(f"conv n={n*bscale} {h}x{w} {c}->{k} r={r_} x{cnt}: " + " ".join(f"{k_}:{v:.1f}" for k_, v in r.items())
              + f"  | hbm floor {hbm_us:.2f}us", flush=True)
    json.dump(out, open(f"gpurun_out/sweep_b{bscale}.json", "w"), indent=1)
