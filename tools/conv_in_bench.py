"""conv_in (C = 4 -> 320, 3x3) / conv_out (320 -> 4) at batch 1 / 8 / 64, back to back in a graph."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops
from tools.tops_sweep import graph_time
dev = torch.device("cuda:0")
lib = _lib.load()
for n in (1, 8, 64):
    for (c, k) in ((4, 320), (320, 4)):
        x = torch.randint(-128, 128, (n, c, 64, 64), dtype=torch.int8, device=dev).contiguous(memory_format=torch.channels_last)
        w = torch.randint(-127, 128, (k, c, 3, 3), dtype=torch.int8, device=dev).contiguous(memory_format=torch.channels_last)
        sc = torch.ones(k, device=dev); s1 = torch.tensor(1.0, device=dev); zp = torch.tensor(3.0, device=dev)
        wsum = w.float().sum(1, keepdim=True).contiguous()
        keep = []
        t = graph_time([lambda: keep.append(ops.qconv2d_w8_a8_ohalf(x, w, sc, s1, zp, sc, wsum, None, None, 1, 1, 1))] * 6)
        print(f"conv n={n:2d} {c}->{k}: {t*1e6:9.1f} us  ({lib.mixdq_last_path().decode()}; output {n*4096*k*2/1e6:.1f} MB)", flush=True)
