"""GroupNorm[+SiLU] -> int8 (dynamic: statistics + apply + quantise; static: statistics + apply)
chained back to back in a CUDA graph, per UNet shape. MIXDQ_B200_LIB selects the library."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import ops
from tools.tops_sweep import graph_time

dev = torch.device("cuda:0")
SHAPES = [(1, 320, 64), (1, 640, 32), (1, 1280, 16), (1, 2560, 16), (1, 1920, 32), (1, 960, 64),
          (8, 320, 64), (8, 1280, 16), (64, 320, 64), (64, 640, 32), (64, 1280, 16)]
inv = torch.tensor(30.0, device=dev); zp = torch.tensor(-3.0, device=dev)
for (n, c, h) in SHAPES:
    x = torch.randn(n, c, h, h, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    w = torch.ones(c, device=dev, dtype=torch.float16); b = torch.zeros(c, device=dev, dtype=torch.float16)
    keep = []
    td = graph_time([lambda: keep.append(ops.groupnorm_quantize_dynamic(x, 32, w, b, 1e-5, True))] * 10)
    keep.clear()
    ts = graph_time([lambda: keep.append(ops.groupnorm_quantize_static(x, 32, w, b, 1e-5, True, inv, zp))] * 10)
    keep.clear()
    print(f"gn n={n:3d} c={c:5d} {h}x{h}: dynamic {td*1e6:8.2f} us   static {ts*1e6:8.2f} us   "
          f"({x.numel()*2/1e6:.1f} MB)", flush=True)
