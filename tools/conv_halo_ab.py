"""3x3 convolutions on the persistent kernel: HALO form (one A box with a one-row halo feeds the
three vertical taps) vs plain per-tap boxes, same process. Back-to-back launches inside a CUDA
graph, weights rotated through several buffers."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops  # noqa: E402
from tools.tops_sweep import graph_time, PEAK  # noqa: E402  (also prints the GEMM sweep when imported? no: guarded below)

dev = torch.device("cuda:0")
lib = _lib.load()
SHAPES = [(8, 64, 64, 320, 320), (8, 32, 32, 640, 640), (8, 16, 16, 1280, 1280), (8, 64, 64, 640, 320),
          (8, 32, 32, 1280, 640), (8, 16, 16, 2560, 1280), (32, 32, 32, 640, 640), (64, 64, 64, 320, 320),
          (64, 32, 32, 640, 640), (64, 16, 16, 1280, 1280), (64, 8, 8, 1280, 1280)]
for (n, h, w_, c, k) in SHAPES:
    x = torch.randint(-128, 128, (n, c, h, w_), dtype=torch.int8, device=dev).contiguous(
        memory_format=torch.channels_last)
    ws = [torch.randint(-127, 128, (k, c, 3, 3), dtype=torch.int8, device=dev).contiguous(
        memory_format=torch.channels_last) for _ in range(3)]
    sc = torch.ones(k, device=dev); s1 = torch.tensor(1.0, device=dev); zp = torch.tensor(3.0, device=dev)
    wsum = [w.float().sum(1, keepdim=True).contiguous() for w in ws]
    res = {}
    for halo in (1, 0, 1, 0):
        lib.mixdq_debug_set_conv_halo(halo)
        outs = []
        t = graph_time([(lambda w=w, s=s: outs.append(ops.qconv2d_w8_a8_ohalf(x, w, sc, s1, zp, sc, s, None, None, 1, 1, 1)))
                        for w, s in zip(ws, wsum)] * 2)
        res.setdefault(halo, []).append((t, lib.mixdq_last_path().decode()))
        del outs
    lib.mixdq_debug_set_conv_halo(1)
    ops_ = 2.0 * n * h * w_ * k * c * 9
    th = min(t for t, _ in res[1]); tp = min(t for t, _ in res[0])
    print(f"conv3x3 n={n:3d} {h}x{w_} c={c} k={k}: halo {th*1e6:8.1f} us {ops_/th/1e12:7.1f} TOP/s ({100*ops_/th/1e12/PEAK:4.1f} %) [{res[1][0][1]}]"
          f"   plain {tp*1e6:8.1f} us {ops_/tp/1e12:7.1f} TOP/s ({100*ops_/tp/1e12/PEAK:4.1f} %) [{res[0][0][1]}]   x{tp/th:.2f}", flush=True)
    del ws, x
